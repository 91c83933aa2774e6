"""Time aether_attention_bf16 modes at the AetherV1 geometry (B=1, S=15076, H=48, dh=64) against torch SDPA, and
report the max deviation from SDPA on the same inputs.  usage: attn_bench.py [mode ...]   (default: 5 8)"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200 import ops  # noqa: E402

modes = [int(a) for a in sys.argv[1:]] or [5, 8]
DEV = "cuda"
B, S, H = 1, 15076, 48
g = torch.Generator(device=DEV).manual_seed(0)
qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g).bfloat16()
flops = 4.0 * B * H * S * S * 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H * 64)
sd, _ = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
print(json.dumps(dict(impl="torch_sdpa", ms=sd, tflops=flops / sd / 1e9)), flush=True)
for m in modes:
    out = ops.attention(qkv, v_fp16=m)
    err = (out.float() - ref.float()).abs().max().item()
    med, best = timeit(lambda: ops.attention(qkv, v_fp16=m))
    print(json.dumps(dict(mode=m, ms=med, ms_best=best, tflops=flops / med / 1e9, max_abs_vs_sdpa=err)), flush=True)
