#!/bin/bash
# Round-2 GPU call N (2 GPUs): tile-parallel goldens incl. a non-zero blend rank, and a short bench run on the final kernels
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/multigpu_sliding.py > gpurun_out/r2n_multigpu_sliding.log 2>&1
echo "exit $?"
grep -n "sliding_\|MULTIGPU" gpurun_out/r2n_multigpu_sliding.log | tail -n 12
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --tile-steps 4 > gpurun_out/r2n_bench_2gpu.json \
    2> gpurun_out/r2n_bench_2gpu.err
echo "exit $?"
tail -c 3000 gpurun_out/r2n_bench_2gpu.json; tail -n 5 gpurun_out/r2n_bench_2gpu.err
