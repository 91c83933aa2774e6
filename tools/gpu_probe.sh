#!/bin/bash
# First-contact GPU run: each kernel family in its own process under `timeout`, so a hang in one
# does not take the others (or the box) down.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_smi.txt 2>&1
for grp in "gemm" "attention" "not gemm and not attention"; do
  name=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp" | tee -a gpurun_out/probe.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" --timeout 120 -p no:cacheprovider \
      > "gpurun_out/probe_${name}.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/probe.log
  tail -n 25 "gpurun_out/probe_${name}.log"
done
