#!/bin/bash
# Round-2 GPU call F (1 GPU): the gated-residual epilogue change (tests + per-kernel timings), full-size VAE conv captures,
# and the first REAL bench line of the round (50 denoise steps per tile, all legs).
mkdir -p gpurun_out
echo "=== pytest (gemm, dit, vae, fullsize properties)" | tee gpurun_out/r2f.log
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_dit_gpu.py tests/test_vae_gpu.py \
    tests/test_fullsize_properties_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2f.log
tail -n 6 gpurun_out/r2f_pytest.log
echo "=== perf_kernels" | tee -a gpurun_out/r2f.log
timeout 600 python tools/perf_kernels.py > gpurun_out/r2f_perf_kernels.log 2>&1
grep -n "'name': 'to_out'\|'name': 'ff2'\|'name': 'qkv'\|'name': 'ff1'" gpurun_out/r2f_perf_kernels.log
timeout 600 python tools/vae_timing.py > gpurun_out/r2f_vae_timing.json 2> /dev/null
cat gpurun_out/r2f_vae_timing.json
echo "=== ncu: full-size VAE convs (encode of 9 frames: launches 3.. are 128-ch 3x3x3 convs at 9x240x360)" | tee -a gpurun_out/r2f.log
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:conv_kernel -s 3 -c 1 -o gpurun_out/r2_vae_conv128_fullres python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
timeout 600 $N -k "regex:conv_kernel<\(int\)256>|conv_kernel<256>" -s 2 -c 1 -o gpurun_out/r2_vae_conv256_halfres python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
timeout 300 $N -k regex:gemm2_kernel -s 1 -c 1 -o gpurun_out/r2_gemm2_to_out_after python tools/prof_one.py gemm_out 2 > /dev/null 2>&1
for f in r2_vae_conv128_fullres r2_vae_conv256_halfres r2_gemm2_to_out_after; do
  if [ -f gpurun_out/$f.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_summary.txt 2>&1
    grep -n "Kernel Name\|Grid Size\|gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed\|lts__t_bytes.sum \[" gpurun_out/${f}_ncu_summary.txt | head -8
  else
    echo "$f: no report"
  fi
done
echo "=== bench (real: 50 steps per tile), K=3 W=3" | tee -a gpurun_out/r2f.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err
echo "bench exit $?" | tee -a gpurun_out/r2f.log
cat gpurun_out/r2f_bench_1gpu.json; tail -n 5 gpurun_out/r2f_bench_1gpu.err
