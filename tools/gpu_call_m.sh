#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/m_launches_fit.csv python bench.py --workload fit --steps 1 --warmup 3 --no-reference-gpu > gpurun_out/m_fit.log 2>&1
python tools/launch_summary.py gpurun_out/m_launches_fit.csv | head -50
