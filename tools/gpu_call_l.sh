#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload fit --steps 5 --warmup 3 > gpurun_out/l_fit1.json 2> gpurun_out/l_fit1.err; tail -5 gpurun_out/l_fit1.err; cat gpurun_out/l_fit1.json
