#!/bin/bash
# End-of-round evidence run on ONE B200 (invoked through gpurun): the complete GPU test suite, the default bench line,
# the reference arm, smoke(), per-kernel timings, ncu --set full captures of the two dominant kernels and the ncu
# launch list of a short bench run.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r1_pytest_gpu.txt
timeout 900 python bench.py 2>&1 | grep -E "^\{" > gpurun_out/r1_bench_1gpu_default.json
cut -c1-300 gpurun_out/r1_bench_1gpu_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep -E "^\{" > gpurun_out/r1_bench_reference_arm.json
cut -c1-300 gpurun_out/r1_bench_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/perf_kernels.py --quick > gpurun_out/r1_perf_kernels.log 2>&1
timeout 300 python tools/attn_bench.py 5 7 8 12 > gpurun_out/r1_attn_bench.log 2>&1
timeout 300 python tools/fullsize_tasks_check.py 2 > gpurun_out/r1_fullsize_tasks.log 2>&1
tail -4 gpurun_out/r1_fullsize_tasks.log
timeout 300 python tools/vae_fullsize_check.py > gpurun_out/r1_vae_fullsize.log 2>&1
tail -3 gpurun_out/r1_vae_fullsize.log
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:attention_v3 -s 1 -c 1 -o gpurun_out/r1_attention_mode5_final python tools/prof_one.py attention 2 5 > /dev/null 2>&1
timeout 300 $N -k regex:gemm2_kernel -s 1 -c 1 -o gpurun_out/r1_gemm2_ff1 python tools/prof_one.py gemm 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aether -s 1041 -c 694 --csv \
    --log-file gpurun_out/r1_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-e2e \
    > gpurun_out/r1_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12
