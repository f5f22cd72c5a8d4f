#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/e_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/e_tests.log
tail -8 gpurun_out/e_tests.log
timeout 300 python tools/encode_cloud_profile.py > gpurun_out/e_encprof.log 2>&1; tail -2 gpurun_out/e_encprof.log
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; python -c "
import json;d=json.load(open('gpurun_out/e_bench.json'));print(d['value'],d['e2e']['value'],d['encoder_s'],d.get('e2e_predict'),d['roofline_fkaconv']['ms_per_call'],d.get('reference_gpu',{}).get('value'))"
