#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/af_bench8.json 2> gpurun_out/af_bench8.err; tail -2 gpurun_out/af_bench8.err; python -c "
import json;d=json.load(open('gpurun_out/af_bench8.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['encoder_s'],d.get('collective_ms'),d.get('e2e_predict'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 4 --steps 5 --warmup 3 --no-predict > gpurun_out/af_bench4.json 2> gpurun_out/af_bench4.err; python -c "
import json;d=json.load(open('gpurun_out/af_bench4.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['encoder_s'])"
