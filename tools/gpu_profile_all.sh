#!/bin/bash
# ncu captures of every kernel family (one launch each, --set full) + the launch list of a short bench run.
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:attention_v3 -s 1 -c 1 -o gpurun_out/r1_attention_mode5 python tools/prof_one.py attention 2 5 > /dev/null 2>&1
$N -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/r1_gemm_ff1 python tools/prof_one.py gemm 2 > /dev/null 2>&1
$N -k regex:ln_modulate -s 1 -c 1 -o gpurun_out/r1_ln_modulate python tools/prof_one.py ln 2 > /dev/null 2>&1
$N -k regex:qk_norm_rope -s 1 -c 1 -o gpurun_out/r1_qk_norm_rope python tools/prof_one.py qk 2 > /dev/null 2>&1
$N -k regex:small_m_linear -s 1 -c 1 -o gpurun_out/r1_adaln_gemv python tools/prof_one.py gemv 2 > /dev/null 2>&1
$N -k regex:cfg_dpm_step -s 1 -c 1 -o gpurun_out/r1_cfg_dpm_step python tools/prof_one.py step 2 > /dev/null 2>&1
$N -k regex:patchify_kernel -s 1 -c 1 -o gpurun_out/r1_patchify python tools/prof_one.py patch 2 > /dev/null 2>&1
$N -k regex:blend_crossfade -s 1 -c 1 -o gpurun_out/r1_blend_crossfade python tools/prof_one.py blend 2 > /dev/null 2>&1
# VAE kernels at full-size tiles: the 60th conv launch of a 9-frame encode+decode is a 128->128 3x3x3 conv at 240x360
$N -k regex:conv_kernel -s 300 -c 1 -o gpurun_out/r1_vae_conv python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
$N -k regex:gn_apply -s 300 -c 1 -o gpurun_out/r1_vae_gn_apply python tools/vae_fullsize_check.py 9 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 700 --csv --log-file gpurun_out/r1_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-e2e > gpurun_out/r1_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -20
