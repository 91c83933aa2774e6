#!/bin/bash
# round 2, call L: CTA-pair conv kernel -- parity (pair == one-CTA bitwise, torch parity), A/B timing, VAE timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vae_gpu.py -x -q -m gpu -k "conv" > gpurun_out/r2l_conv_tests.log 2>&1
tail -5 gpurun_out/r2l_conv_tests.log
timeout 300 python tools/conv_ab.py --json gpurun_out/r2l_conv_ab.json > gpurun_out/r2l_conv_ab.log 2>&1
cat gpurun_out/r2l_conv_ab.log | tail -12
