#!/bin/bash
# Round-2 GPU call T (1 GPU, the last of the budget), kept as the record of the experiment: attention mode 6 (one MMA-issuer
# warp per query tile, tools/experiments/attention/two_issuer_warps.patch) against mode 5 -- bit identity + isolated timing
# at full size, the kernel parity tests for mode 6, a DiT forward with mode 6.  Result: correct but slower; not merged.
mkdir -p gpurun_out
timeout 120 python tools/attn_bench.py 5 6 > gpurun_out/r2t_attn_bench.log 2>&1
echo "attn_bench exit $?"; cat gpurun_out/r2t_attn_bench.log | cut -c1-260
timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "attention and 6" > gpurun_out/r2t_pytest_mode6.log 2>&1
echo "pytest exit $?"; tail -n 2 gpurun_out/r2t_pytest_mode6.log
AETHER_ATTENTION_MODE=6 timeout 80 python -m pytest tests/test_dit_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2t_pytest_dit_mode6.log 2>&1
echo "dit exit $?"; tail -n 2 gpurun_out/r2t_pytest_dit_mode6.log
