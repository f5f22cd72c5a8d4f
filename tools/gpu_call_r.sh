#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider -k "training_step" 2>&1 | tail -4
timeout 600 python bench.py --workload fit --steps 10 --warmup 3 --no-reference-gpu > gpurun_out/r_fit1.json 2> gpurun_out/r_fit1.err; tail -3 gpurun_out/r_fit1.err; python -c "
import json; d=json.load(open('gpurun_out/r_fit1.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_first_last_warmup'])"
timeout 500 python tools/train_profile.py bf16 2>&1 | grep -v Warn | head -24
