#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/predict_profile.py 2>&1 | grep -v Warn | sed -n 1,40p
timeout 200 python tools/knn_shard_probe.py 2>&1 | tail -9
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "knn or region or full_size or predict or stitch" 2>&1 | tail -4
