"""Run ONE kernel family at full shape a few times (for ncu captures).  usage: prof_one.py attention|gemm|ln"""
import sys, math, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200 import ops
which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(0)
S, D, H = 15076, 3072, 48
if which == "attention":
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    qkv = torch.randn(1, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    if mode:
        qkv.view(torch.float16)[:, :, 2] = qkv[:, :, 2].float().half()
    for _ in range(n): ops.attention(qkv, v_fp16=mode)
elif which == "gemm":
    a = torch.randn(S, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(4 * D, D, device="cuda", generator=g) / math.sqrt(D)).bfloat16()
    b = torch.zeros(4 * D, device="cuda")
    for _ in range(n): ops.gemm(a, w, b, 1)
elif which == "ln":
    x = torch.randn(1, S, D, device="cuda", generator=g).bfloat16()
    gm = torch.ones(D, device="cuda"); bt = torch.zeros(D, device="cuda"); mod = torch.randn(1, 4 * D, device="cuda")
    for _ in range(n): ops.ln_modulate(x, gm, bt, 1e-5, mod[:, :D], mod[:, D:2*D], mod[:, 2*D:3*D], mod[:, 3*D:], St=226, mod_bstride=4*D)
torch.cuda.synchronize()
