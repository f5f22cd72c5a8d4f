#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fkaconv or encoder or binary or config3 or latent or pipeline or stitch or two_same or real_cloud" > gpurun_out/c_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/c_tests.log
tail -5 gpurun_out/c_tests.log
timeout 300 python tools/fka_bench.py > gpurun_out/c_fka_bench.log 2>&1; cat gpurun_out/c_fka_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fka_ -s 9 -c 3 -o gpurun_out/prof_fka_v2 python tools/fka_bench.py --one > gpurun_out/c_ncu.log 2>&1; tail -3 gpurun_out/c_ncu.log
timeout 300 python tools/encode_cloud_profile.py > gpurun_out/c_encprof.log 2>&1; tail -2 gpurun_out/c_encprof.log
