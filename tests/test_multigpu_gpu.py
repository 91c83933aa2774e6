"""N > 1 on real GPUs (skipped on a single-GPU box): the tile-parallel sliding-window path with the NCCL all-gather,
launched exactly like the driver launches bench.py (torch.distributed.run, one rank per GPU, 127.0.0.1)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sliding_window_tile_parallel_nccl():
    n = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        str(ROOT / "tools" / "multigpu_sliding.py")], capture_output=True, text=True, timeout=600,
                       cwd=str(ROOT))
    assert r.returncode == 0 and "MULTIGPU_SLIDING_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
