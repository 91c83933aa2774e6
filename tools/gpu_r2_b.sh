#!/bin/bash
# Round-2 GPU call B (1 GPU): VAE native schedule (tests + timing + launch list), pipeline / full-size tests, the
# vectorised streaming kernels, compute-sanitizer on small shapes.
mkdir -p gpurun_out
echo "=== pytest (vae, pipeline, fullsize parity, kernels subset)" | tee gpurun_out/r2b.log
timeout 1500 python -m pytest tests/test_vae_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_parity_gpu.py \
    tests/test_fullsize_properties_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -s \
    > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2b.log
grep -n "rel-rms\|passed\|failed\|Error" gpurun_out/r2b_pytest.log | tail -n 30
echo "=== VAE timing" | tee -a gpurun_out/r2b.log
timeout 600 python tools/vae_timing.py > gpurun_out/r2b_vae_timing.json 2> gpurun_out/r2b_vae_timing.err
echo "vae timing exit $?" | tee -a gpurun_out/r2b.log
cat gpurun_out/r2b_vae_timing.json; tail -n 5 gpurun_out/r2b_vae_timing.err
echo "=== perf_kernels" | tee -a gpurun_out/r2b.log
timeout 600 python tools/perf_kernels.py > gpurun_out/r2b_perf_kernels.log 2>&1
echo "perf exit $?" | tee -a gpurun_out/r2b.log
tail -n 40 gpurun_out/r2b_perf_kernels.log
echo "=== compute-sanitizer (small shapes)" | tee -a gpurun_out/r2b.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_kernels_gpu.py tests/test_vae_gpu.py -m gpu -q \
      -p no:cacheprovider -k "short_sequences or gemm_cta_pair or conv3d_matches_torch or groupnorm" --timeout 800 \
      > gpurun_out/r2b_sanitizer_${tool}.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/r2b.log
  grep -n "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|Error" gpurun_out/r2b_sanitizer_${tool}.log | tail -n 8
done
