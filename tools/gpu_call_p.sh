#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_multiprocess.py -m gpu -q --no-header -p no:cacheprovider -k "ddp or two_gpus" 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --workload fit --gpus 2 --steps 10 --warmup 3 > gpurun_out/p_fit2.json 2> gpurun_out/p_fit2.err; tail -3 gpurun_out/p_fit2.err; cat gpurun_out/p_fit2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --workload fit --gpus 2 --steps 10 --warmup 3 --fit-eager > gpurun_out/p_fit2_eager.json 2> gpurun_out/p_fit2_eager.err; tail -3 gpurun_out/p_fit2_eager.err; cat gpurun_out/p_fit2_eager.json
