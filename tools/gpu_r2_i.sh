#!/bin/bash
# Round-2 GPU call I (1 GPU), kept as the record of how the attention experiments were run: experimental modes 6 (lazy
# max), 7 (packed f32x2 math), 8 (both) vs mode 5 -- parity, isolated timing, in-step timing (development rounds of 10
# denoise steps).  The packed math became mode 5 and the experimental ids were removed again (git history), so today the
# dispatcher rejects 6 / 7 and this script only reproduces the mode-5 rows.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_properties_gpu.py tests/test_vae_gpu.py -m gpu -q \
    -k "attention or native or vae_decode" --timeout 600 -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1
echo "pytest exit $?"; tail -n 8 gpurun_out/r2i_pytest.log
timeout 300 python tools/attn_bench.py 5 6 7 > gpurun_out/r2i_attn_bench.log 2>&1
cat gpurun_out/r2i_attn_bench.log
for mode in 5 6 7; do
  AETHER_ATTENTION_MODE=$mode timeout 600 python bench.py --steps 3 --warmup 3 --tile-steps 10 --no-cpu-baseline \
      --no-gpu-library-baseline --no-strong-leg --no-exchange-check > gpurun_out/r2i_bench_mode${mode}.json 2> gpurun_out/r2i_bench_mode${mode}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2i_bench_mode${mode}.json"))
    print("mode ${mode}: round ms", round(d["ms_per_step"], 1), "attention ms", round(d["roofline"]["avg_launch_ms"], 4), "clock", d["clocks"]["sm_mhz"])
except Exception as e:
    print("mode ${mode}: failed", e)
PY
done
