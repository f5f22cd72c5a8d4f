#!/bin/bash
# round-2 closing call: smoke, whole GPU suite, default bench line, launch list of the decode, config 3 line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke\|Error"
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/k_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/k_tests.log
tail -3 gpurun_out/k_tests.log
timeout 900 python bench.py > gpurun_out/k_bench1.json 2> gpurun_out/k_bench1.err; tail -3 gpurun_out/k_bench1.err; python -c "
import json;d=json.load(open('gpurun_out/k_bench1.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms_per_step'],d['encoder_s'],d['e2e_predict']['total_s'],d['fit']['value'],d['fit']['ms_per_step'], d['max_abs_err_vs_oracle'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_launches_res65.csv python bench.py --profile-run --resolution 65 --steps 2 --warmup 1 --latents random --no-cpu-baseline --no-reference-gpu --no-predict > gpurun_out/k_ncu1.log 2>&1
timeout 600 python bench.py --num-pts-local 200 --points 250000 --no-cpu-baseline --no-reference-gpu --no-predict --no-fit > gpurun_out/k_bench_cfg3.json 2> gpurun_out/k_bench_cfg3.err
python -c "
import json;d=json.load(open('gpurun_out/k_bench_cfg3.json'));print('cfg3',d['value'],d['e2e']['value'],d['encoder_s'])" || tail -5 gpurun_out/k_bench_cfg3.err
