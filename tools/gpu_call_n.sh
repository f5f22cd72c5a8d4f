#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -15
timeout 600 python bench.py --workload fit --steps 10 --warmup 3 > gpurun_out/n_fit1.json 2> gpurun_out/n_fit1.err; tail -5 gpurun_out/n_fit1.err; cat gpurun_out/n_fit1.json
timeout 600 python bench.py --workload fit --steps 10 --warmup 3 --fit-eager --no-reference-gpu > gpurun_out/n_fit1_eager.json 2> gpurun_out/n_fit1_eager.err; tail -3 gpurun_out/n_fit1_eager.err; cat gpurun_out/n_fit1_eager.json
