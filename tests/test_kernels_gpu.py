"""GPU parity of every CUDA kernel, called through the C ABI (aether_b200.ops -> ctypes), against plain
fp32 torch restatements of the same op (the floating-point 'oracle' for a single kernel).

Tolerances (bf16 storage, fp32 accumulate): a result element r is accepted when
|r - ref| <= atol + rtol * |ref| with rtol = 2^-7 (two bf16 ulps) and an atol scaled to the output magnitude;
each test states its own numbers.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from aether_b200 import ops
    return ops


def _close(got, ref, rtol, atol, what):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = (err > bound).sum().item()
    assert bad == 0, (f"{what}: {bad}/{err.numel()} elements out of tolerance; max err {err.max().item():.4g}, "
                      f"ref absmax {ref.abs().max().item():.4g}")


# --------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (300, 256, 384), (1000, 768, 1024),
                                   (4276, 3072, 3072), (226, 3072, 4096), (500, 224, 3072)])
def test_gemm_bias(M, N, K):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    out = ops.gemm(a, w, bias, 0)
    ref = a.float() @ w.float().t() + bias
    _close(out, ref, 2 ** -7, 2e-2, f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K,epi", [(1025, 256, 128, 0), (1400, 320, 192, 0), (2048, 512, 64, 1), (3000, 768, 320, 1),
                                       (15076, 256, 256, 0)])
def test_gemm_cta_pair(M, N, K, epi):
    """M >= 1024 runs the cta_group::2 kernel: ragged M (second CTA of the last pair partly / fully out of range),
    ragged N (W rows zero-filled by TMA), fewer tiles than CTA pairs, both arithmetic epilogues."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g) * 0.5
    out = ops.gemm(a, w, bias, epi)
    ref = a.float() @ w.float().t() + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    _close(out, ref, 2 ** -7, 2e-2, f"gemm2 {M}x{N}x{K} epi{epi}")


def test_gemm_cta_pair_gated_residual_fp16_columns():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(66)
    B, S, St, N, K = 2, 1283, 226, 512, 256
    M = B * S
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    gate_v = torch.randn(B, N, device=DEV, generator=g)
    gate_t = torch.randn(B, N, device=DEV, generator=g)
    resid = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    out = resid.clone()
    ops.gemm(a, w, bias, 2, out=out, gate_vid=gate_v, gate_txt=gate_t, S=S, St=St)
    y = (a.float() @ w.float().t() + bias).view(B, S, N)
    gate = torch.where((torch.arange(S, device=DEV) < St)[None, :, None], gate_t[:, None, :], gate_v[:, None, :])
    _close(out.view(B, S, N), resid.float().view(B, S, N) + gate * y, 2 ** -7, 3e-2, "gemm2 gated residual")
    out16 = ops.gemm(a, w, bias, 0, f16_from_col=256)
    ref = a.float() @ w.float().t() + bias
    _close(out16[:, :256], ref[:, :256], 2 ** -7, 2e-2, "gemm2 bf16 part")
    _close(out16.view(torch.float16)[:, 256:], ref[:, 256:], 2 ** -10, 2e-3, "gemm2 fp16 part")


def test_gemm_gelu():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    M, N, K = 700, 1024, 256
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    out = ops.gemm(a, w, bias, 1)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + bias, approximate="tanh")
    _close(out, ref, 2 ** -7, 2e-2, "gemm+gelu")


def test_gemm_gated_residual():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(6)
    B, S, St, N, K = 2, 333, 18, 256, 512
    M = B * S
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    gate_v = torch.randn(B, N, device=DEV, generator=g)
    gate_t = torch.randn(B, N, device=DEV, generator=g)
    resid = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    out = resid.clone()
    ops.gemm(a, w, bias, 2, out=out, gate_vid=gate_v, gate_txt=gate_t, S=S, St=St)
    y = (a.float() @ w.float().t() + bias).view(B, S, N)
    gate = torch.where((torch.arange(S, device=DEV) < St)[None, :, None], gate_t[:, None, :], gate_v[:, None, :])
    ref = resid.float().view(B, S, N) + gate * y
    _close(out.view(B, S, N), ref, 2 ** -7, 3e-2, "gemm gated residual")


# --------------------------------------------------------------------------------------------- attention
def _attention_case(qkv_f32, v_fp16):
    """Run the kernel in either mode; the reference uses exactly the 16-bit-rounded operands the kernel sees."""
    ops = _ops()
    qkv = qkv_f32.bfloat16()
    v_ref = qkv[:, :, 2].float()
    if v_fp16 == 2:                                     # modes 0 and 5 keep V in bf16
        vh = qkv_f32[:, :, 2].half()
        qkv.view(torch.float16)[:, :, 2] = vh          # V third carries fp16 bit patterns
        v_ref = vh.float()
    out = ops.attention(qkv, v_fp16=v_fp16)
    B, S, _, H, _ = qkv.shape
    q, k = (qkv[:, :, i].float().transpose(1, 2) for i in range(2))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v_ref.transpose(1, 2))
    return out, ref.transpose(1, 2).reshape(B, S, H * 64)


@pytest.mark.parametrize("v_fp16", [0, 2, 5])
@pytest.mark.parametrize("B,S,H", [(1, 256, 2), (1, 128, 1), (2, 318, 4), (1, 1000, 2), (1, 4276, 4)])
def test_attention(B, S, H, v_fp16):
    g = torch.Generator(device=DEV).manual_seed(S + H)
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g)
    out, ref = _attention_case(qkv, v_fp16)
    _close(out, ref, 2 ** -6, 1e-2, f"attention B{B} S{S} H{H} fp16={v_fp16}")


@pytest.mark.parametrize("v_fp16", [0, 2, 5])
@pytest.mark.parametrize("S", [1, 40, 64, 65, 128, 129, 191, 193, 256, 257, 385, 500, 512, 641])
def test_attention_short_sequences(S, v_fp16):
    """1, 2, 3 ... key tiles: prologue / drain paths of the pipelined kernels, ragged last tile, empty second query tile."""
    g = torch.Generator(device=DEV).manual_seed(S)
    qkv = torch.randn(1, S, 3, 3, 64, device=DEV, generator=g)
    out, ref = _attention_case(qkv, v_fp16)
    _close(out, ref, 2 ** -6, 1e-2, f"attention short S{S} mode={v_fp16}")


@pytest.mark.parametrize("v_fp16", [0, 2, 5])
def test_attention_large_logits(v_fp16):
    """Rows whose running max grows by more than 2^8 between key tiles exercise the lazy O (and l) rescale."""
    g = torch.Generator(device=DEV).manual_seed(11)
    B, S, H = 1, 640, 2
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g)
    qkv[:, :, 0] *= 4.0
    qkv[:, 300:, 1] *= 6.0          # later keys produce much larger logits
    out, ref = _attention_case(qkv, v_fp16)
    _close(out, ref, 2 ** -6, 2e-2, f"attention large logits fp16={v_fp16}")


@pytest.mark.parametrize("v_fp16", [0, 2, 5])
def test_attention_peaked_and_flat_rows(v_fp16):
    """Nearly one-hot rows (large scale) and nearly uniform rows (tiny scale) in the same launch."""
    g = torch.Generator(device=DEV).manual_seed(12)
    B, S, H = 1, 900, 2
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g)
    qkv[:, :450, 0] *= 8.0
    qkv[:, 450:, 0] *= 0.01
    out, ref = _attention_case(qkv, v_fp16)
    _close(out, ref, 2 ** -6, 2e-2, f"attention peaked/flat fp16={v_fp16}")


@pytest.mark.parametrize("B,S,St,H,K", [(1, 300, 18, 2, 128), (2, 700, 226, 4, 256), (1, 1500, 226, 6, 192)])
def test_fused_qkv_projection_equals_gemm_then_qk_norm_rope(B, S, St, H, K):
    """The QKV GEMM with QK-LayerNorm + RoPE in its epilogue against the two separate kernels it replaces (single-CTA
    and CTA-pair kernels: rows = B*S below and above 1024).  Same roundings and summation order => identical output
    up to the rare contraction difference (<= 1 bf16 ulp on a handful of elements)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + S + H)
    x = torch.randn(B * S, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(3 * H * 64, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(3 * H * 64, device=DEV, generator=g) * 0.3
    gq, gk = (1 + 0.2 * torch.randn(64, device=DEV, generator=g) for _ in range(2))
    bq, bk = (0.1 * torch.randn(64, device=DEV, generator=g) for _ in range(2))
    cos = torch.rand(S - St, 64, device=DEV, generator=g) * 2 - 1
    sin = torch.rand(S - St, 64, device=DEV, generator=g) * 2 - 1
    ref = ops.gemm(x, w, bias, 0).view(B, S, 3, H, 64).clone()
    ops.qk_norm_rope(ref, gq, bq, gk, bk, 1e-6, cos, sin, St)
    out = ops.gemm_qkv_norm_rope(x, w, bias, B, S, St, H, gq, bq, gk, bk, 1e-6, cos, sin)
    assert torch.equal(out[:, :, 2], ref[:, :, 2])                       # V third: plain projection
    diff = (out.float() - ref.float()).abs()
    ulp = ref.float().abs().clamp_min(2 ** -6) * 2 ** -7
    assert (diff <= ulp).all(), f"max diff {diff.max().item()}"
    assert (diff > 0).float().mean().item() < 1e-3


def test_attention_unknown_variant_rejected():
    qkv = torch.zeros(1, 128, 3, 1, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="AETHER_ERR_INVALID"):
        _ops().attention(qkv, v_fp16=99)


def test_gemm_fp16_columns():
    """Columns >= f16_from_col are emitted as fp16 (V third of the fused QKV projection)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(21)
    M, N, K, split = 300, 768, 256, 512
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    out = ops.gemm(a, w, bias, 0, f16_from_col=split)
    ref = a.float() @ w.float().t() + bias
    _close(out[:, :split], ref[:, :split], 2 ** -7, 2e-2, "bf16 part")
    _close(out.view(torch.float16)[:, split:], ref[:, split:], 2 ** -10, 2e-3, "fp16 part")


# --------------------------------------------------------------------------------------------- HBM kernels
@pytest.mark.parametrize("D", [256, 3072])
def test_ln_modulate(D):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(D)
    B, S, St = 2, 77, 18
    x = (torch.randn(B, S, D, device=DEV, generator=g) * 2 + 0.5).bfloat16()
    gamma = 1 + 0.1 * torch.randn(D, device=DEV, generator=g)
    beta = 0.1 * torch.randn(D, device=DEV, generator=g)
    mod = torch.randn(B, 4 * D, device=DEV, generator=g) * 0.3
    sv, cv, st, ct = (mod[:, i * D:(i + 1) * D] for i in range(4))
    out = ops.ln_modulate(x, gamma, beta, 1e-5, sv, cv, st, ct, St=St, mod_bstride=4 * D)
    ln = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
    is_t = (torch.arange(S, device=DEV) < St)[None, :, None]
    ref = ln * (1 + torch.where(is_t, ct[:, None], cv[:, None])) + torch.where(is_t, st[:, None], sv[:, None])
    _close(out, ref, 2 ** -7, 1e-2, "ln_modulate")
    # double LayerNorm variant (norm_final -> norm_out.norm -> modulate)
    g2 = 1 + 0.1 * torch.randn(D, device=DEV, generator=g)
    b2 = 0.1 * torch.randn(D, device=DEV, generator=g)
    out2 = ops.ln_modulate(x, gamma, beta, 1e-5, sv, cv, St=0, gamma2=g2, beta2=b2, mod_bstride=4 * D)
    ref2 = torch.nn.functional.layer_norm(ln, (D,), g2, b2, 1e-5) * (1 + cv[:, None]) + sv[:, None]
    _close(out2, ref2, 2 ** -7, 1e-2, "double ln_modulate")


@pytest.mark.parametrize("D,affine", [(256, True), (3072, True), (768, False)])
def test_ln_modulate_persistent_path(D, affine):
    """rows >= 1024 take the shared-memory-parameter persistent kernel (text and video segments, 2 batch items)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(D + 1)
    B, S, St = 2, 700, 226
    x = (torch.randn(B, S, D, device=DEV, generator=g) * 1.5 - 0.2).bfloat16()
    gamma = (1 + 0.1 * torch.randn(D, device=DEV, generator=g)) if affine else None
    beta = (0.1 * torch.randn(D, device=DEV, generator=g)) if affine else None
    mod = torch.randn(B, 4 * D, device=DEV, generator=g) * 0.3
    sv, cv, st, ct = (mod[:, i * D:(i + 1) * D] for i in range(4))
    out = ops.ln_modulate(x, gamma, beta, 1e-5, sv, cv, st, ct, St=St, mod_bstride=4 * D)
    ln = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
    is_t = (torch.arange(S, device=DEV) < St)[None, :, None]
    ref = ln * (1 + torch.where(is_t, ct[:, None], cv[:, None])) + torch.where(is_t, st[:, None], sv[:, None])
    _close(out, ref, 2 ** -7, 1e-2, "ln_modulate persistent")


def test_qk_norm_rope():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(3)
    B, S, St, H = 2, 150, 18, 4
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g).bfloat16()
    gq, bq, gk, bk = (torch.randn(64, device=DEV, generator=g) * 0.2 + (1 if i % 2 == 0 else 0) for i in range(4))
    ang = torch.rand(S - St, 32, device=DEV, generator=g) * 6.28
    cos = ang.cos().repeat_interleave(2, dim=1).contiguous()
    sin = ang.sin().repeat_interleave(2, dim=1).contiguous()
    ref = qkv.clone().float()
    for idx, (gm, bt) in enumerate(((gq, bq), (gk, bk))):
        x = torch.nn.functional.layer_norm(qkv[:, :, idx].float(), (64,), gm, bt, 1e-6).bfloat16().float()
        xv = x[:, St:]
        xr, xi = xv.reshape(B, S - St, H, 32, 2).unbind(-1)
        rot = torch.stack([-xi, xr], dim=-1).flatten(3)
        x[:, St:] = xv * cos[None, :, None, :] + rot * sin[None, :, None, :]
        ref[:, :, idx] = x
    ops.qk_norm_rope(qkv, gq, bq, gk, bk, 1e-6, cos, sin, St)
    _close(qkv, ref, 2 ** -7, 1e-2, "qk_norm_rope")


@pytest.mark.parametrize("B,N,K,act", [(1, 64, 3072, 0), (2, 64, 64, 1), (2, 5000, 512, 1), (1, 777, 256, 1),
                                       (3, 100, 128, 0)])
def test_small_m_linear(B, N, K, act):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(N + K)
    x = torch.randn(B, K, device=DEV, generator=g)
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    y = ops.small_m_linear(x, w, bias, act)
    xa = torch.nn.functional.silu(x) if act else x
    ref = xa @ w.float().t() + bias
    _close(y, ref, 1e-4, 1e-4, "small_m_linear")


def test_timestep_sinusoid():
    ops = _ops()
    t = torch.tensor([999, 19], device=DEV, dtype=torch.int64)
    emb = ops.timestep_sinusoid(t, 3072, True, 0.0)
    half = 1536
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, device=DEV, dtype=torch.float32) / half)
    arg = t[:, None].float() * freq[None]
    ref = torch.cat([arg.cos(), arg.sin()], dim=-1)
    _close(emb, ref, 0, 2e-4, "timestep sinusoid")   # fp32 sin/cos of arguments up to ~1e3


def test_patchify_unpatchify():
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(9)
    B, F, Cc, H, W = 2, 3, 96, 12, 20
    x = torch.randn(B, F, Cc, H, W, device=DEV, generator=g).bfloat16()
    p = ops.patchify(x)
    ref = x.view(B, F, Cc, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(-1, Cc * 4)
    assert torch.equal(p, ref)
    Co = 56
    tok = torch.randn(B * F * (H // 2) * (W // 2), Co * 4, device=DEV, generator=g).bfloat16()
    out = ops.unpatchify(tok, B, F, Co, H, W)
    ref = tok.view(B, F, H // 2, W // 2, Co, 2, 2).permute(0, 1, 4, 2, 5, 3, 6).reshape(B, F, Co, H, W)
    assert torch.equal(out, ref)
    # full-size geometry (11 x 96 x 60 x 90) is bit-exact as well
    x = torch.randn(1, 11, 96, 60, 90, device=DEV, generator=g).bfloat16()
    assert torch.equal(ops.patchify(x), x.view(1, 11, 96, 30, 2, 45, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(-1, 384))


def test_cfg_dpm_step_bit_exact():
    """K8 against the torch expression graph of the oracle scheduler (same promotions) -- bit exact."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle.scheduler import OracleDPMScheduler
    ops = _ops()
    sch = OracleDPMScheduler()
    sch.set_timesteps(50)
    g = torch.Generator(device=DEV).manual_seed(1)
    shape = (1, 5, 56, 12, 20)
    sample = torch.randn(shape, device=DEV, generator=g).bfloat16()
    v2 = torch.randn((2,) + shape[1:], device=DEV, generator=g).bfloat16()
    n1 = torch.randn(shape, device=DEV, generator=g).bfloat16()
    n2 = torch.randn(shape, device=DEV, generator=g).bfloat16()
    old = torch.randn(shape, device=DEV, generator=g)
    guidance = 2.37
    for (t, tb, old_x0) in [(999, None, None), (979, 999, old), (499, 519, old), (19, 39, old)]:
        c = sch.coefficients(t, tb)
        vf = v2.float()
        v = vf[:1] + guidance * (vf[1:] - vf[:1])
        ref_prev, ref_x0 = sch.step(v, old_x0, t, tb, sample, noises=[n1, n2])
        second = old_x0 is not None and c["prev_t"] >= 0
        co = ops.dpm_coeffs(c["sqrt_a"], c["sqrt_1ma"], c["m1"], c["m2"], c["m3"] if second else 0.0,
                            c["m4"] if second else 0.0, c["m_noise"], second)
        prev, prev32, x0 = ops.cfg_dpm_step(v2, sample, co, n1, n2, old_x0, guidance, want_prev_f32=True)
        assert torch.equal(x0, ref_x0), f"x0 mismatch at t={t}"
        assert torch.equal(prev32, ref_prev.float()), f"prev mismatch at t={t}"
        assert torch.equal(prev, ref_prev.to(torch.bfloat16))
