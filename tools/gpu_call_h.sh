#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "decode or tensor_core or full_size or from_latent or binary or config3 or stitch" > gpurun_out/h_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/h_tests.log
tail -4 gpurun_out/h_tests.log
timeout 300 python tools/tc_phase_profile.py 2>&1 | tail -9
timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu --no-predict > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; python -c "
import json;d=json.load(open('gpurun_out/h_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms_per_step'])"
