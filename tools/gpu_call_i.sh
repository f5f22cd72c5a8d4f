#!/bin/bash
# round-2 final session: smoke, whole GPU suite, default bench line, launch list + ncu --set full of the tcgen05 decode kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke\|Error"
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/j_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/j_tests.log
tail -3 gpurun_out/j_tests.log
timeout 900 python bench.py > gpurun_out/j_bench1.json 2> gpurun_out/j_bench1.err; tail -3 gpurun_out/j_bench1.err; python -c "
import json;d=json.load(open('gpurun_out/j_bench1.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms_per_step'],d['encoder_s'],d['e2e_predict']['total_s'],d['fit']['value'],d['fit']['ms_per_step'], d['max_abs_err_vs_oracle'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j_launches_res65.csv python bench.py --profile-run --resolution 65 --steps 2 --warmup 1 --latents random --no-cpu-baseline --no-reference-gpu --no-predict > gpurun_out/j_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"projection_tc|pn_stn|pn_feat|knn_warp" -c 4 -o gpurun_out/prof_decode_r02c python bench.py --profile-run --resolution 65 --steps 1 --warmup 0 --latents random --no-cpu-baseline --no-reference-gpu --no-predict > gpurun_out/j_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
