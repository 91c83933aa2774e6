#!/bin/bash
# round 2, call M: VAE after CTA-pair convs + one-launch GroupNorm statistics + cache written by the norm kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_fullsize_parity_gpu.py -x -q -m gpu -k "vae or VAE or conv or gn or norm" > gpurun_out/r2m_vae_tests.log 2>&1
tail -5 gpurun_out/r2m_vae_tests.log
timeout 300 python tools/vae_timing.py > gpurun_out/r2m_vae_timing.log 2>&1
tail -12 gpurun_out/r2m_vae_timing.log
timeout 300 python tools/vae_profile.py > gpurun_out/r2m_vae_profile.json 2> gpurun_out/r2m_vae_profile.err
cut -c1-1500 gpurun_out/r2m_vae_profile.json
