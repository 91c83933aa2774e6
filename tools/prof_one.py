"""Run ONE kernel family at full shape a few times (for ncu captures).
usage: prof_one.py attention [n] [mode] | gemm | gemm_out | ln | qk | gemv | step | patch | blend"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200 import ops  # noqa: E402

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(0)
S, D, H = 15076, 3072, 48
if which == "attention":
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    qkv = torch.randn(1, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    if mode in (1, 2):
        qkv.view(torch.float16)[:, :, 2] = qkv[:, :, 2].float().half()
    for _ in range(n):
        ops.attention(qkv, v_fp16=mode)
elif which == "gemm":
    a = torch.randn(S, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(4 * D, D, device="cuda", generator=g) / math.sqrt(D)).bfloat16()
    b = torch.zeros(4 * D, device="cuda")
    for _ in range(n):
        ops.gemm(a, w, b, 1)
elif which == "gemm_out":     # to_out: N = K = 3072, gated residual epilogue in place (the GEMM furthest from cuBLAS)
    a = torch.randn(S, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(D, D, device="cuda", generator=g) / math.sqrt(D)).bfloat16()
    b = torch.zeros(D, device="cuda")
    gv = torch.randn(1, D, device="cuda", generator=g); gt = torch.randn(1, D, device="cuda", generator=g)
    out = torch.randn(S, D, device="cuda", generator=g).bfloat16()
    for _ in range(n):
        ops.gemm(a, w, b, 2, out=out, gate_vid=gv, gate_txt=gt, S=S, St=226)
elif which == "ln":
    x = torch.randn(1, S, D, device="cuda", generator=g).bfloat16()
    gm = torch.ones(D, device="cuda"); bt = torch.zeros(D, device="cuda"); mod = torch.randn(1, 4 * D, device="cuda")
    for _ in range(n):
        ops.ln_modulate(x, gm, bt, 1e-5, mod[:, :D], mod[:, D:2 * D], mod[:, 2 * D:3 * D], mod[:, 3 * D:], St=226,
                        mod_bstride=4 * D)
elif which == "qk":
    qkv = torch.randn(1, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    vec = torch.ones(64, device="cuda")
    cos = torch.rand(S - 226, 64, device="cuda"); sin = torch.rand(S - 226, 64, device="cuda")
    for _ in range(n):
        ops.qk_norm_rope(qkv, vec, vec, vec, vec, 1e-6, cos, sin, 226)
elif which == "gemv":
    N = (12 * 42 + 2) * D
    w = torch.randn(N, 512, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(1, 512, device="cuda"); b = torch.zeros(N, device="cuda")
    for _ in range(n):
        ops.small_m_linear(x, w, b, 1)
elif which == "step":
    shape = (1, 11, 56, 60, 90)
    v = torch.randn(shape, device="cuda", generator=g).bfloat16()
    s = torch.randn(shape, device="cuda", generator=g).bfloat16()
    n1 = torch.randn(shape, device="cuda", generator=g).bfloat16()
    n2 = torch.randn(shape, device="cuda", generator=g).bfloat16()
    old = torch.randn(shape, device="cuda", generator=g)
    co = ops.dpm_coeffs(0.5, 0.8, 0.4, -0.1, 1.2, 0.2, 0.3, True)
    for _ in range(n):
        ops.cfg_dpm_step(v, s, co, n1, n2, old, 1.0)
elif which == "patch":
    lat = torch.randn(1, 11, 96, 60, 90, device="cuda").bfloat16()
    tok = torch.randn(14850, 224, device="cuda").bfloat16()
    for _ in range(n):
        ops.patchify(lat)
        ops.unpatchify(tok, 1, 11, 56, 60, 90)
elif which == "blend":
    from aether_b200.sliding_window import blend_chain
    wins = [torch.rand(41, 480, 720, device="cuda") for _ in range(2)]
    for _ in range(n):
        blend_chain(wins, [(0, 41), (8, 49)], 0)
torch.cuda.synchronize()
