#!/bin/bash
# Round-2 GPU call H (8 GPUs): the tile-parallel bench at N = 8 exactly like the driver launches it (development rounds of
# 10 denoise steps per tile to bound the cost: 8 GPUs are charged 8x), incl. the 24-tile golden check over 8 ranks.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_gpus.txt
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
    --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 3 --tile-steps 10 > gpurun_out/r2h_bench_8gpu.json \
    2> gpurun_out/r2h_bench_8gpu.err
echo "exit $?"
cat gpurun_out/r2h_bench_8gpu.json; grep -v "^$\|OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2h_bench_8gpu.err | tail -n 8
