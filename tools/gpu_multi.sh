#!/bin/bash
# records of the multi-GPU configurations (run under gpurun --gpus 8)
mkdir -p gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_bench$N.json 2> gpurun_out/multi_bench$N.err; python -c "
import json;d=json.load(open('gpurun_out/multi_bench$N.json'));print('decode',d['value'],d['ms_per_step'],d['e2e']['value'],d['encoder_s'],d.get('e2e_predict',{}).get('total_s'))"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --workload fit --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_fit$N.json 2> gpurun_out/multi_fit$N.err; python -c "
import json;d=json.load(open('gpurun_out/multi_fit$N.json'));print('fit',d['value'],d['ms_per_step'],d['e2e']['value'])"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus $N --resolution 257 --steps 3 --warmup 3 --no-predict > gpurun_out/multi_res257_$N.json 2> gpurun_out/multi_res257_$N.err; python -c "
import json;d=json.load(open('gpurun_out/multi_res257_$N.json'));print('res257',d['value'],d['ms_per_step'],d['e2e']['value'])"
