#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/encoder_kernel_profile.py 2>&1 | grep "encode_cloud\|total device"
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "latent or predict or encoder or sampling or two_gpus or same_shaped" 2>&1 | tail -4
