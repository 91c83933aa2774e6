"""Time aether_attention_bf16 modes at the AetherV1 geometry (B=1, S=15076, H=48, dh=64) against torch SDPA, and
report the max deviation from SDPA on the same inputs.  usage: attn_bench.py [mode ...]   (default: 5)"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200 import ops  # noqa: E402

modes = [int(a) for a in sys.argv[1:]] or [5]
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
from aether_b200 import _lib  # noqa: E402
from aether_b200._lib import check, current_stream, ptr  # noqa: E402
lib = _lib.load()
for m in modes:
    if m == 2:                                          # fp16 P/V variant: V third re-encoded as fp16
        qkv.view(torch.float16)[:, :, 2] = qkv[:, :, 2].float().half()
    for split in ((True, False) if m == 5 else (False,)):
        need = lib.aether_attention_workspace_bytes(B, S, H, m) if split else 0
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=DEV)       # allocated once: the timing is the kernels only
        out = torch.empty(B, S, H * 64, dtype=torch.bfloat16, device=DEV)

        def run():
            check(lib.aether_attention_bf16_ws(ptr(qkv), ptr(out), B, S, H, 0.125, m, ptr(ws) if need else 0, need,
                                               current_stream()), "attention")
        run()
        err = (out.float() - ref.float()).abs().max().item()
        med, best = timeit(run)
        print(json.dumps(dict(mode=m, split_tail=bool(need), ms=med, ms_best=best, tflops=flops / med / 1e9,
                              max_abs_vs_sdpa=err)), flush=True)
