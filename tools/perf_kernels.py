"""Per-kernel timing at the full 41-frame 480x720 shapes (S = 15076, D = 3072, 48 heads), CUDA events on the
launching stream, warm-up 3, L2 flushed between iterations.  Writes gpurun_out/perf_kernels.json.
cuBLAS / torch SDPA timings are printed next to ours for context only (they are not on the product path)."""
import json
import math
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200 import ops  # noqa: E402

DEV = "cuda"
PEAKS = {}
pk = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
if pk.exists():
    PEAKS = json.loads(pk.read_text())
TF_PEAK = PEAKS.get("bf16_tflops", 1590.0)
HBM_PEAK = PEAKS.get("hbm_gbs", 6650.0)

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def timeit(fn, iters=5, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    quick = "--quick" in sys.argv
    res = {"peaks": {"bf16_tflops": TF_PEAK, "hbm_gbs": HBM_PEAK}, "gemm": [], "attention": [], "hbm": []}
    S, D, H = 15076, 3072, 48
    g = torch.Generator(device=DEV).manual_seed(0)
    # ------------------------------------------------------------------ GEMMs
    shapes = [("qkv", S, 3 * D, D, 0), ("to_out", S, D, D, 2), ("ff1", S, 4 * D, D, 1), ("ff2", S, D, 4 * D, 2),
              ("qkv_b2", 2 * S, 3 * D, D, 0)]
    for name, M, N, K, epi in shapes:
        a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
        w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=DEV, generator=g)
        out = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        gate = torch.randn(1 + M // S, N, device=DEV, generator=g)
        kw = dict(gate_vid=gate, gate_txt=gate, S=S, St=226) if epi == 2 else {}
        med, best = timeit(lambda: ops.gemm(a, w, bias, epi, out=out, **kw))
        flops = 2.0 * M * N * K
        cb_med, cb_best = timeit(lambda: torch.matmul(a, w.t(), out=out))
        r = dict(name=name, M=M, N=N, K=K, epi=epi, ms=med, ms_best=best, tflops=flops / med / 1e9,
                 frac_of_measured_peak=flops / med / 1e9 / TF_PEAK, cublas_ms=cb_med, cublas_tflops=flops / cb_med / 1e9)
        print(r, flush=True)
        res["gemm"].append(r)
        del a, w, out
    # ------------------------------------------------------------------ attention
    for B in ([1] if quick else [1, 2]):
        qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g).bfloat16()
        flops = 4.0 * B * H * S * S * 64
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
        sd_med, _ = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters=3)
        for mode in (5, 0):
            if mode == 1:     # V third re-encoded as fp16 for the fp16-PV modes
                qkv.view(torch.float16)[:, :, 2] = qkv[:, :, 2].float().half()
            med, best = timeit(lambda: ops.attention(qkv, v_fp16=mode), iters=3)
            r = dict(B=B, S=S, H=H, mode=mode, ms=med, ms_best=best, tflops=flops / med / 1e9,
                     frac_of_measured_peak=flops / med / 1e9 / TF_PEAK, torch_sdpa_ms=sd_med,
                     torch_sdpa_tflops=flops / sd_med / 1e9)
            print(r, flush=True)
            res["attention"].append(r)
        del qkv
    # ------------------------------------------------------------------ HBM kernels
    x = torch.randn(1, S, D, device=DEV, generator=g).bfloat16()
    gamma = torch.ones(D, device=DEV); beta = torch.zeros(D, device=DEV)
    mod = torch.randn(1, 4 * D, device=DEV, generator=g)
    y = torch.empty_like(x)
    med, best = timeit(lambda: ops.ln_modulate(x, gamma, beta, 1e-5, mod[:, :D], mod[:, D:2 * D], mod[:, 2 * D:3 * D],
                                               mod[:, 3 * D:], St=226, mod_bstride=4 * D, out=y))
    byt = 2.0 * x.numel() * 2
    res["hbm"].append(dict(name="ln_modulate", ms=med, gbs=byt / med / 1e6, frac=byt / med / 1e6 / HBM_PEAK))
    qkv = torch.randn(1, S, 3, H, 64, device=DEV, generator=g).bfloat16()
    vec = torch.ones(64, device=DEV)
    cos = torch.rand(S - 226, 64, device=DEV); sin = torch.rand(S - 226, 64, device=DEV)
    med, best = timeit(lambda: ops.qk_norm_rope(qkv, vec, vec, vec, vec, 1e-6, cos, sin, 226))
    byt = 2.0 * qkv.numel() * 2 * 2 / 3
    res["hbm"].append(dict(name="qk_norm_rope", ms=med, gbs=byt / med / 1e6, frac=byt / med / 1e6 / HBM_PEAK))
    N_adaln = (12 * 42 + 2) * D
    w = torch.randn(N_adaln, 512, device=DEV, dtype=torch.bfloat16)
    xb = torch.randn(1, 512, device=DEV)
    bias = torch.zeros(N_adaln, device=DEV)
    med, best = timeit(lambda: ops.small_m_linear(xb, w, bias, 1))
    byt = N_adaln * 512 * 2.0
    res["hbm"].append(dict(name="adaln_gemv_B1", ms=med, gbs=byt / med / 1e6, frac=byt / med / 1e6 / HBM_PEAK))
    xb2 = torch.randn(2, 512, device=DEV)
    med, best = timeit(lambda: ops.small_m_linear(xb2, w, bias, 1))
    res["hbm"].append(dict(name="adaln_gemv_B2", ms=med, gbs=byt / med / 1e6, frac=byt / med / 1e6 / HBM_PEAK))
    lat = torch.randn(1, 11, 96, 60, 90, device=DEV).bfloat16()
    med, best = timeit(lambda: ops.patchify(lat))
    byt = 2.0 * lat.numel() * 2
    res["hbm"].append(dict(name="patchify", ms=med, gbs=byt / med / 1e6, frac=byt / med / 1e6 / HBM_PEAK))
    for r in res["hbm"]:
        print(r, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    Path("gpurun_out/perf_kernels.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
