#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke\|Error"
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/v_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/v_tests.log
tail -4 gpurun_out/v_tests.log
timeout 900 python bench.py > gpurun_out/v_bench1.json 2> gpurun_out/v_bench1.err; tail -3 gpurun_out/v_bench1.err; python -c "
import json;d=json.load(open('gpurun_out/v_bench1.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms_per_step'],d['encoder_s'],d['e2e_predict']['total_s'],d['fit']['value'],d['fit']['ms_per_step'], d['max_abs_err_vs_oracle'])"
