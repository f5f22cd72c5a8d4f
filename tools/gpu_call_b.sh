#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fkaconv or encoder or binary or config3 or latent or smoke or pipeline" > gpurun_out/b_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/b_tests.log
tail -5 gpurun_out/b_tests.log
timeout 300 python tools/fka_bench.py > gpurun_out/b_fka_bench.log 2>&1; cat gpurun_out/b_fka_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fka_fused -s 3 -c 1 -o gpurun_out/prof_fka_v1 python tools/fka_bench.py --one > gpurun_out/b_ncu.log 2>&1; tail -3 gpurun_out/b_ncu.log
timeout 300 python tools/encode_cloud_profile.py > gpurun_out/b_encprof.log 2>&1; tail -2 gpurun_out/b_encprof.log
