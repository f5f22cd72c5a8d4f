#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/f_tests.log
tail -6 gpurun_out/f_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/f_bench2.json 2> gpurun_out/f_bench2.err
python -c "
import json;d=json.load(open('gpurun_out/f_bench2.json'));print(d['value'],d['e2e']['value'],d['encoder_s'],d.get('collective_ms'),d.get('e2e_predict'))" || tail -20 gpurun_out/f_bench2.err
timeout 900 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err; python -c "
import json;d=json.load(open('gpurun_out/f_bench1.json'));print(d['value'],d['e2e']['value'],d['encoder_s'],d.get('e2e_predict'))"
