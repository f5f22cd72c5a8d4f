#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/ac_bench2.json 2> gpurun_out/ac_bench2.err; tail -2 gpurun_out/ac_bench2.err; python -c "
import json;d=json.load(open('gpurun_out/ac_bench2.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['encoder_s'],d.get('collective_ms'),d.get('e2e_predict'))"
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_multiprocess.py -m gpu -q --no-header -p no:cacheprovider -k "ddp or two_gpus" 2>&1 | tail -3
