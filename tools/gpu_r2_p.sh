#!/bin/bash
# Round-2 GPU call P (8 GPUs): the real line at N = 8 -- 50 denoise steps per tile, K = 2 timed rounds, W = 3 -- launched like
# the driver launches it.
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
    --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2p_bench_8gpu.json \
    2> gpurun_out/r2p_bench_8gpu.err
echo "exit $?"
cat gpurun_out/r2p_bench_8gpu.json | cut -c1-400; grep -v "^$\|OMP_NUM_THREADS\|\*\*\*\*" gpurun_out/r2p_bench_8gpu.err | tail -n 6
