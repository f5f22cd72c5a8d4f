#!/bin/bash
mkdir -p gpurun_out
# config 3 (ppsurf_200nn, 250k points) and config 4's workload on one GPU
timeout 900 python bench.py --num-pts-local 200 --points 250000 --cpu-sample 16384 > gpurun_out/g_bench_cfg3.json 2> gpurun_out/g_bench_cfg3.err
python -c "
import json;d=json.load(open('gpurun_out/g_bench_cfg3.json'));print('cfg3',d['value'],d['e2e']['value'],d['encoder_s'],d.get('max_abs_err_vs_oracle'),d.get('reference_gpu',{}).get('value'))" || tail -5 gpurun_out/g_bench_cfg3.err
timeout 900 python bench.py --resolution 257 --no-cpu-baseline --no-reference-gpu --steps 2 > gpurun_out/g_bench_res257.json 2> gpurun_out/g_bench_res257.err
python -c "
import json;d=json.load(open('gpurun_out/g_bench_res257.json'));print('res257',d['value'],d['e2e']['value'],d.get('e2e_predict'))" || tail -5 gpurun_out/g_bench_res257.err
# launch list of the decode step (resolution 65, 2 steps) and of one encoder pass batch
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_launches_res65.csv python bench.py --profile-run --resolution 65 --steps 2 --warmup 1 --latents random --no-cpu-baseline --no-reference-gpu --no-predict > gpurun_out/g_ncu1.log 2>&1
# full captures: kNN of the grid queries, the PointNet kernels, projection
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_warp -c 2 -o gpurun_out/prof_knn_r02 python bench.py --profile-run --resolution 65 --steps 1 --warmup 0 --latents random --no-cpu-baseline --no-reference-gpu --no-predict > gpurun_out/g_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
