#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "knn or config3 or large_patches or full_size" 2>&1 | tail -4
timeout 250 python tools/decode_profile.py 250000 200 129 2>&1 | grep -v Warn | tail -10
