#!/bin/bash
# Round-2 GPU call S (1 GPU): ncu launch list of the bench command (this library's kernels around the first two DiT forwards)
# and a fresh `--set full` capture of the final attention kernel with its stall hotspots.
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:attention_v3 -s 2 -c 1 \
    -o gpurun_out/r2_attention_mode5_final python tools/prof_one.py attention 4 5 > /dev/null 2>&1
echo "attention capture exit $?"
timeout 230 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:aether:: \
    --launch-skip 4000 -c 1300 --csv --log-file gpurun_out/r2s_bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --tile-steps 2 --no-cpu-baseline --no-gpu-library-baseline --no-strong-leg \
    --no-exchange-check > gpurun_out/r2s_bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2s_bench_launches.csv
python tools/ncu_summary.py gpurun_out/r2_attention_mode5_final.ncu-rep > gpurun_out/r2_attention_mode5_final_ncu_summary.txt 2>&1
python tools/ncu_stall_hotspots.py gpurun_out/r2_attention_mode5_final.ncu-rep 900 > gpurun_out/r2_attention_mode5_final_stall_hotspots.txt 2>&1
grep -n "gpu__time_duration.sum\|xu.avg.pct_of_peak_sustained_elapsed\|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed\|registers_per_thread \|issue_active" gpurun_out/r2_attention_mode5_final_ncu_summary.txt | head
head -3 gpurun_out/r2_attention_mode5_final_stall_hotspots.txt
