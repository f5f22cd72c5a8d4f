#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "knn or full_size or stitch or config3 or latent" 2>&1 | tail -4
timeout 200 python tools/knn_shard_probe.py 2>&1 | tail -9
timeout 200 python tools/knn_run_probe2.py 2>&1 | tail -4
timeout 200 python tools/decode_profile.py 100000 50 129 2>&1 | grep -v Warn | tail -8
timeout 300 python tools/predict_profile.py 2>&1 | grep "sweeps\|untimed\|knn_warp"
