#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_bf16_probe.py 2>&1 | tail -12
