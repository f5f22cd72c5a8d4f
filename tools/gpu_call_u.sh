#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/knn_shard_probe.py 2>&1 | tail -9
timeout 900 python bench.py --no-cpu-baseline --no-fit > gpurun_out/u_bench1.json 2> gpurun_out/u_bench1.err; python -c "
import json;d=json.load(open('gpurun_out/u_bench1.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms_per_step'],d['encoder_s'],d.get('e2e_predict'))"
