#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --resolution 257 --steps 3 --warmup 3 > gpurun_out/q_res257_8gpu.json 2> gpurun_out/q_res257_8gpu.err; tail -2 gpurun_out/q_res257_8gpu.err; cat gpurun_out/q_res257_8gpu.json | head -c 1500; echo
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --workload fit --gpus 8 --steps 10 --warmup 3 > gpurun_out/q_fit8.json 2> gpurun_out/q_fit8.err; tail -2 gpurun_out/q_fit8.err; cat gpurun_out/q_fit8.json | head -c 1500; echo
