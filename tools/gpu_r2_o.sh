#!/bin/bash
# Round-2 GPU call O (1 GPU): final state -- whole `pytest -m gpu` suite, smoke(), memcheck of the new VAE kernels on small
# shapes, ncu captures of the CTA-pair convs at the decoder's shapes, the bench line (50 steps per tile) and the reference arm.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2o_pytest.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/r2o_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/r2o_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_vae_gpu.py -m gpu -q -x \
    -p no:cacheprovider -k "conv or gn or norm or native" > gpurun_out/r2o_sanitizer_memcheck_vae.log 2>&1
echo "memcheck exit $?"; tail -n 4 gpurun_out/r2o_sanitizer_memcheck_vae.log
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:conv2_kernel -s 3 -c 1 -o gpurun_out/r2_vae_conv2_128_fullres python tools/conv_ab.py --only 4 --reps 1 > /dev/null 2>&1
timeout 400 $N -k regex:conv2_kernel -s 3 -c 1 -o gpurun_out/r2_vae_conv2_256_halfres python tools/conv_ab.py --only 2 --reps 1 > /dev/null 2>&1
for f in r2_vae_conv2_128_fullres r2_vae_conv2_256_halfres; do
  if [ -f gpurun_out/$f.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_summary.txt 2>&1
    grep -n "Kernel Name\|Grid Size\|gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" gpurun_out/${f}_ncu_summary.txt | head -7
  fi
done
echo "=== bench (50 steps per tile), default flags"
timeout 1800 python bench.py > gpurun_out/r2o_bench_1gpu.json 2> gpurun_out/r2o_bench_1gpu.err
echo "bench exit $?"
cat gpurun_out/r2o_bench_1gpu.json; tail -n 3 gpurun_out/r2o_bench_1gpu.err
echo "=== reference arm"
timeout 900 python bench.py --impl reference > gpurun_out/r2o_bench_reference.json 2> gpurun_out/r2o_bench_reference.err
cat gpurun_out/r2o_bench_reference.json
