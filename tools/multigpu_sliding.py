"""Tile-parallel sliding-window path on N GPUs (torchrun, NCCL): every rank runs its share of the tiles in rounds, the
disparity tiles of each round travel peer to peer over NVLink to the blend rank, which streams them onto the chain; the
result (broadcast back for this check) is compared on every rank with the reference's golden blend
(tests/golden/sliding_*.npz).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_sliding.py
"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import fake_tile_outputs, subsample, synthetic_long_clip  # noqa: E402
from aether_b200.sliding_window import plan_windows, process_with_sliding_window  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ok = True
    # the last pass puts the blend rank on the LAST rank (tile 0's rgb then follows the disparity tiles there)
    for name, root in (("temporal", 0), ("horizontal", 0), ("vertical", 0), ("long", 0), ("long", world - 1)):
        g = np.load(ROOT / "tests" / "golden" / f"sliding_{name}.npz")
        t, h, w = g["thw"].tolist()
        obs = synthetic_long_clip(t, h, w)
        ran = []

        def tile_fn(tl, crop):
            ran.append(tl.k)
            return fake_tile_outputs(crop, tl.t_start, tl.h_start, tl.w_start)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        stats = {}
        rgb, disp = process_with_sliding_window(None, obs, 4, t, 3407, rank=rank, world_size=world, tile_fn=tile_fn,
                                                root=root, result_on_all_ranks=True, stats=stats)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n_tiles = len(plan_windows(t, h, w, t).tiles)
        good = (np.allclose(subsample(disp, (3, 16, 16)), g["disparity_sub"], rtol=2e-6, atol=0)
                and abs(disp.sum() - float(g["disparity_sum"])) <= 2e-6 * abs(float(g["disparity_sum"]))
                and np.array_equal(subsample(rgb, (8, 32, 32, 1)), g["rgb_sub"])
                and ran == list(range(rank, n_tiles, world)))
        ok = ok and good
        print(f"[rank {rank}/{world}] sliding_{name} (blend rank {root}): tiles run here {ran} of {n_tiles}; matches reference golden: {good}; "
              f"{dt:.2f} s; device-timed ms {dict((k, round(v, 2)) for k, v in stats.items())}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_SLIDING_OK" if flag.item() == 1 else "MULTIGPU_SLIDING_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
