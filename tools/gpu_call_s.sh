#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/train_gemm_bench.py 2>&1 | tail -10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 4 -o gpurun_out/prof_train_gemm_r02 -f python tools/train_gemm_bench.py --once > gpurun_out/s_ncu.log 2>&1; tail -3 gpurun_out/s_ncu.log
ncu -i gpurun_out/prof_train_gemm_r02.ncu-rep --page raw --csv > gpurun_out/prof_train_gemm_r02_raw.csv 2>/dev/null; ls -la gpurun_out/prof_train_gemm_r02*
