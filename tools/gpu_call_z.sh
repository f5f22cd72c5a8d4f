#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --workload fit --steps 10 --warmup 3 --no-reference-gpu > gpurun_out/z_fit1.json 2> gpurun_out/z_fit1.err; tail -3 gpurun_out/z_fit1.err; python -c "
import json; d=json.load(open('gpurun_out/z_fit1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
