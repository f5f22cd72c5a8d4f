#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/j_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/j_tests.log
tail -60 gpurun_out/j_tests.log
