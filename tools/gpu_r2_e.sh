#!/bin/bash
# Round-2 GPU call E (1 GPU): split-tail attention (tests + isolated timing), DiT / pipeline tests on the new build,
# VAE profile after the gn_finalize / conv_yb changes, ncu captures (to_out GEMM, full-size VAE convs, attention).
mkdir -p gpurun_out
echo "=== pytest (kernels, dit, fullsize properties, vae, input)" | tee gpurun_out/r2e.log
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_dit_gpu.py tests/test_fullsize_properties_gpu.py \
    tests/test_vae_gpu.py tests/test_input_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider \
    > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2e.log
tail -n 12 gpurun_out/r2e_pytest.log
echo "=== attention isolated (split tail on/off)" | tee -a gpurun_out/r2e.log
timeout 300 python tools/attn_bench.py 5 > gpurun_out/r2e_attn_bench.log 2>&1
cat gpurun_out/r2e_attn_bench.log
echo "=== VAE timing + profile" | tee -a gpurun_out/r2e.log
timeout 600 python tools/vae_timing.py > gpurun_out/r2e_vae_timing.json 2> gpurun_out/r2e_vae_timing.err
cat gpurun_out/r2e_vae_timing.json; tail -n 3 gpurun_out/r2e_vae_timing.err
timeout 600 python tools/vae_profile.py > gpurun_out/r2e_vae_profile.json 2> /dev/null
python - <<'PY'
import json
for line in open("gpurun_out/r2e_vae_profile.json"):
    d = json.loads(line)
    top = list(d["kernels"].items())[:8]
    print(d["op"], d["wall_ms"], d["sum_kernel_ms"], [(k, v["total_ms"]) for k, v in top])
PY
echo "=== ncu captures" | tee -a gpurun_out/r2e.log
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:gemm2_kernel -s 1 -c 1 -o gpurun_out/r2_gemm2_to_out python tools/prof_one.py gemm_out 2 > /dev/null 2>&1
timeout 600 $N -k regex:conv_kernel -s 100 -c 1 -o gpurun_out/r2_vae_conv_a python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
timeout 600 $N -k regex:conv_kernel -s 160 -c 1 -o gpurun_out/r2_vae_conv_b python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
timeout 300 $N -k regex:attention_v3_kernel -s 2 -c 1 -o gpurun_out/r2_attention_mode5 python tools/prof_one.py attention 2 5 > /dev/null 2>&1
timeout 300 $N -k regex:scale_reduce -s 1 -c 1 -o gpurun_out/r2_scale_reduce python tools/prof_one.py blend 2 > /dev/null 2>&1
for f in r2_gemm2_to_out r2_vae_conv_a r2_vae_conv_b r2_attention_mode5 r2_scale_reduce; do
  if [ -f gpurun_out/$f.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_summary.txt 2>&1
    grep -n "Kernel Name\|Grid Size\|gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed\|dram_throughput.avg\|pipe_xu.avg.pct_of_peak_sustained_elapsed" gpurun_out/${f}_ncu_summary.txt | head -9
  else
    echo "$f: no report"
  fi
done
