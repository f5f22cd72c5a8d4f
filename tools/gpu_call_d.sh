#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/d_tests.log
tail -5 gpurun_out/d_tests.log
timeout 300 python tools/encode_cloud_profile.py > gpurun_out/d_encprof.log 2>&1; tail -3 gpurun_out/d_encprof.log
timeout 300 python tools/fka_bench.py > gpurun_out/d_fka_bench.log 2>&1; cat gpurun_out/d_fka_bench.log
timeout 600 python tools/pass_ablation.py > gpurun_out/d_pass_ablation.md 2> gpurun_out/d_pass_ablation.err; cat gpurun_out/d_pass_ablation.md; tail -3 gpurun_out/d_pass_ablation.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; python -c "
import json;d=json.load(open('gpurun_out/d_bench.json'));print(d['value'],d['encoder_s'],d.get('e2e_predict'),d['roofline_fkaconv']['ms_per_call'])"
