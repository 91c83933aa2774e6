#!/bin/bash
# Round-2 GPU call D (2 GPUs): the tile-parallel path over NCCL -- golden check on both ranks, the pytest wrapper, and a
# short bench run (4 denoise steps per tile) through torch.distributed.run exactly like the driver launches it.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2d_gpus.txt
echo "=== tools/multigpu_sliding.py" | tee gpurun_out/r2d.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/multigpu_sliding.py > gpurun_out/r2d_multigpu_sliding.log 2>&1
echo "exit $?" | tee -a gpurun_out/r2d.log
grep -n "sliding_\|MULTIGPU" gpurun_out/r2d_multigpu_sliding.log | tail -n 12
echo "=== pytest tests/test_multigpu_gpu.py" | tee -a gpurun_out/r2d.log
timeout 700 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
tail -n 3 gpurun_out/r2d_pytest.log
echo "=== bench --gpus 2 (dev: 4 denoise steps per tile)" | tee -a gpurun_out/r2d.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --tile-steps 4 > gpurun_out/r2d_bench_2gpu.json \
    2> gpurun_out/r2d_bench_2gpu.err
echo "exit $?" | tee -a gpurun_out/r2d.log
tail -c 3500 gpurun_out/r2d_bench_2gpu.json; tail -n 8 gpurun_out/r2d_bench_2gpu.err
