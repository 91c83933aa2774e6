#!/bin/bash
# Round-2 GPU call C (1 GPU): new input-side tests, VAE per-kernel profile, racecheck re-run on the CTA-pair GEMM,
# ncu --set full captures of the rewritten streaming kernels and of full-size VAE convs.
mkdir -p gpurun_out
echo "=== pytest (input, vae native)" | tee gpurun_out/r2c.log
timeout 900 python -m pytest tests/test_input_gpu.py tests/test_vae_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider \
    -k "input or native or resize or u8 or device_clip" > gpurun_out/r2c_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r2c.log
tail -n 15 gpurun_out/r2c_pytest.log
echo "=== VAE per-kernel profile" | tee -a gpurun_out/r2c.log
timeout 600 python tools/vae_profile.py > gpurun_out/r2c_vae_profile.json 2> gpurun_out/r2c_vae_profile.err
echo "vae profile exit $?" | tee -a gpurun_out/r2c.log
cat gpurun_out/r2c_vae_profile.json; tail -n 5 gpurun_out/r2c_vae_profile.err
echo "=== racecheck, CTA-pair GEMM" | tee -a gpurun_out/r2c.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -m gpu -q \
    -p no:cacheprovider -k "gemm_cta_pair" --timeout 500 > gpurun_out/r2c_sanitizer_racecheck_gemm.log 2>&1
grep -n "RACECHECK SUMMARY\|passed\|failed" gpurun_out/r2c_sanitizer_racecheck_gemm.log | tail -n 4
echo "=== ncu captures" | tee -a gpurun_out/r2c.log
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:cfg_dpm_step -s 1 -c 1 -o gpurun_out/r2_cfg_dpm_step python tools/prof_one.py step 2 > /dev/null 2>&1
timeout 300 $N -k regex:blend_crossfade -s 1 -c 1 -o gpurun_out/r2_blend_crossfade python tools/prof_one.py blend 2 > /dev/null 2>&1
timeout 300 $N -k regex:scale_reduce -s 1 -c 1 -o gpurun_out/r2_scale_reduce python tools/prof_one.py blend 2 > /dev/null 2>&1
# full-size VAE kernels: one 128->128 3x3x3 conv at 9 x 240 x 360 (up_block 3), one 256->256 (up_block 2), gn_apply
timeout 600 $N -k regex:gn_apply_kernel -s 600 -c 1 -o gpurun_out/r2_vae_gn_apply python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
timeout 600 $N -k "regex:conv_kernel<.*128" -s 330 -c 1 -o gpurun_out/r2_vae_conv128 python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
timeout 600 $N -k "regex:conv_kernel<.*256" -s 150 -c 1 -o gpurun_out/r2_vae_conv256 python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
for f in r2_cfg_dpm_step r2_blend_crossfade r2_scale_reduce r2_vae_gn_apply r2_vae_conv128 r2_vae_conv256; do
  if [ -f gpurun_out/$f.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_summary.txt 2>&1
    grep -n "Kernel Name\|Grid Size\|gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed\|dram_throughput.avg" gpurun_out/${f}_ncu_summary.txt | head -8
  else
    echo "$f: no report"
  fi
done
