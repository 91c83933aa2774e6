#!/bin/bash
# Round-2 GPU call A: full GPU test suite (incl. the new full-size parity tests), a short bench run through the new
# round-based workload, and the in-step A/B of attention mode 2 vs 5.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
echo "=== pytest -m gpu" | tee gpurun_out/r2a.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider -s > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2a.log
tail -n 40 gpurun_out/r2a_pytest.log
echo "=== bench (dev: 10 denoise steps per tile)" | tee -a gpurun_out/r2a.log
for mode in 5 2; do
  AETHER_ATTENTION_MODE=$mode timeout 600 python bench.py --steps 3 --warmup 3 --tile-steps 10 --no-cpu-baseline \
      --no-gpu-library-baseline > gpurun_out/r2a_bench_mode${mode}.json 2> gpurun_out/r2a_bench_mode${mode}.err
  echo "bench mode $mode exit $?" | tee -a gpurun_out/r2a.log
  tail -c 3000 gpurun_out/r2a_bench_mode${mode}.json
  tail -n 5 gpurun_out/r2a_bench_mode${mode}.err
done
echo "=== isolated attention timings" | tee -a gpurun_out/r2a.log
timeout 300 python tools/attn_bench.py 5 2 0 > gpurun_out/r2a_attn_bench.log 2>&1
cat gpurun_out/r2a_attn_bench.log
