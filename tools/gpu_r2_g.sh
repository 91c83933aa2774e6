#!/bin/bash
# Round-2 GPU call G (1 GPU): real bench line (50 steps per tile, K=3, W=3, all legs), pose / input GPU tests, conv<256> capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pose_blend.py tests/test_input_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/r2g_pytest.log
echo "=== bench (real: 50 steps per tile), K=3 W=3"
timeout 1800 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err
echo "bench exit $?"
cat gpurun_out/r2g_bench_1gpu.json; tail -n 5 gpurun_out/r2g_bench_1gpu.err
echo "=== reference arm"
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err
cat gpurun_out/r2g_bench_reference.json
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:conv_kernel -s 12 -c 1 -o gpurun_out/r2_vae_conv256_halfres python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
if [ -f gpurun_out/r2_vae_conv256_halfres.ncu-rep ]; then
  python tools/ncu_summary.py gpurun_out/r2_vae_conv256_halfres.ncu-rep > gpurun_out/r2_vae_conv256_halfres_ncu_summary.txt 2>&1
  grep -n "Kernel Name\|Grid Size\|gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" gpurun_out/r2_vae_conv256_halfres_ncu_summary.txt | head -7
fi
