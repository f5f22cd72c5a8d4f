#!/bin/bash
mkdir -p gpurun_out
# full captures of the round-2 decode kernels on the resolution-65 grid (one launch each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"projection_tc|pn_stn|pn_feat" -c 3 -f -o gpurun_out/prof_decode_r02 python bench.py --profile-run --resolution 65 --steps 1 --warmup 0 --latents random --no-cpu-baseline --no-reference-gpu --no-predict --no-fit > gpurun_out/aa_ncu.log 2>&1
tail -2 gpurun_out/aa_ncu.log; ls -la gpurun_out/prof_decode_r02.ncu-rep
python tools/ncu_summary.py gpurun_out/prof_decode_r02.ncu-rep > gpurun_out/prof_decode_r02_summary.csv; wc -l gpurun_out/prof_decode_r02_summary.csv
