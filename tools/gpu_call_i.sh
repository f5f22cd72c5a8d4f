#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/i_tests.log
tail -6 gpurun_out/i_tests.log
timeout 900 python bench.py > gpurun_out/i_bench1.json 2> gpurun_out/i_bench1.err; python -c "
import json;d=json.load(open('gpurun_out/i_bench1.json'));print(d['value'],d['e2e']['value'],d['encoder_s'],d.get('e2e_predict'),d.get('reference_gpu'))"
