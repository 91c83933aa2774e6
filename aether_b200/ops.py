"""Per-kernel Python entry points: torch tensors in, C-ABI calls out (no torch compute on this path).

Each function mirrors one `aether_*` symbol of include/aether_b200.h and validates dtype/contiguity
before passing raw pointers.  Used by the module wrappers (transformer.py, scheduler.py, ...) and by
the parity tests, which call the CUDA path through exactly this boundary.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check, current_stream, device_guard, ptr

BF16, F32, F64 = torch.bfloat16, torch.float32, torch.float64


def _need(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous CUDA {dtype}, got {t.dtype} cuda={t.is_cuda} "
                         f"contiguous={t.is_contiguous()}")


@device_guard
def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, epilogue: int = 0,
         out: Optional[torch.Tensor] = None, gate_vid=None, gate_txt=None, S: int = 0, St: int = 0,
         f16_from_col: int = -1) -> torch.Tensor:
    """out[M,N] = epi(a[M,K] @ w[N,K]^T).  epilogue 2 updates `out` in place (out = out + gate * (acc + bias)).
    Columns >= f16_from_col of the (bf16-typed) output hold fp16 bit patterns."""
    lib = _lib.require_device()
    _need(a, BF16, "a"); _need(w, BF16, "w")
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=BF16, device=a.device)
    _need(out, BF16, "out")
    gb = 0
    if epilogue == 2:
        _need(gate_vid, F32, "gate_vid"); _need(gate_txt, F32, "gate_txt")
        gb = gate_vid.stride(0) if gate_vid.dim() == 2 else N
    if bias is not None:
        _need(bias, F32, "bias")
    check(lib.aether_gemm_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K, ptr(bias),
                               epilogue, ptr(gate_vid), ptr(gate_txt), gb, S, St, f16_from_col, current_stream()),
          "gemm_bf16")
    return out


@device_guard
def attention(qkv: torch.Tensor, scale: Optional[float] = None, v_fp16: int = 5, split_tail: bool = True) -> torch.Tensor:
    """qkv [B,S,3,H,64] bf16 -> out [B,S,H*64] bf16.  `v_fp16` = kernel variant (5 default, 0 baseline, 2 fp16 P/V: the V
    third then holds fp16 bit patterns, `qkv.view(torch.float16)[:, :, 2] = v.half()`).  `split_tail` hands the kernel
    the scratch it needs to cut the last, partially filled wave of its grid along the keys (variant 5 only)."""
    lib = _lib.require_device()
    _need(qkv, BF16, "qkv")
    B, S, three, H, dh = qkv.shape
    assert three == 3 and dh == 64
    out = torch.empty(B, S, H * dh, dtype=BF16, device=qkv.device)
    sc = float(scale if scale is not None else dh ** -0.5)
    need = lib.aether_attention_workspace_bytes(B, S, H, int(v_fp16)) if split_tail else 0
    if need > 0:
        ws = torch.empty(need, dtype=torch.uint8, device=qkv.device)
        check(lib.aether_attention_bf16_ws(ptr(qkv), ptr(out), B, S, H, sc, int(v_fp16), ptr(ws), need, current_stream()),
              "attention_bf16_ws")
    else:
        check(lib.aether_attention_bf16(ptr(qkv), ptr(out), B, S, H, sc, int(v_fp16), current_stream()), "attention_bf16")
    return out


@device_guard
def ln_modulate(x, gamma, beta, eps, shift_vid, scale_vid, shift_txt=None, scale_txt=None, St=0, gamma2=None,
                beta2=None, mod_bstride=None, out=None):
    """x [B,S,D] bf16; shift/scale fp32 [B,D] (or views with batch stride mod_bstride)."""
    lib = _lib.require_device()
    _need(x, BF16, "x")
    B, S, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    if mod_bstride is None:
        mod_bstride = shift_vid.stride(0) if shift_vid.dim() == 2 else D
    check(lib.aether_ln_modulate(ptr(x), ptr(out), B, S, St, D, ptr(gamma), ptr(beta), float(eps), ptr(gamma2),
                                 ptr(beta2), ptr(shift_vid), ptr(scale_vid), ptr(shift_txt), ptr(scale_txt),
                                 mod_bstride, current_stream()), "ln_modulate")
    return out


@device_guard
def qk_norm_rope(qkv, gq, bq, gk, bk, eps=1e-6, cos=None, sin=None, St=0):
    lib = _lib.require_device()
    _need(qkv, BF16, "qkv")
    B, S, _, H, _ = qkv.shape
    for t in (gq, bq, gk, bk):
        _need(t, F32, "qk norm param")
    if cos is not None:
        _need(cos, F32, "cos"); _need(sin, F32, "sin")
        assert cos.shape == (S - St, 64)
    check(lib.aether_qk_norm_rope(ptr(qkv), B, S, St, H, ptr(gq), ptr(bq), ptr(gk), ptr(bk), float(eps), ptr(cos),
                                  ptr(sin), current_stream()), "qk_norm_rope")
    return qkv


@device_guard
def gemm_qkv_norm_rope(x, w, bias, B, S, St, H, gq, bq, gk, bk, eps=1e-6, cos=None, sin=None, f16_from_col=-1):
    """x [B*S, K] bf16, w [3*H*64, K] bf16 -> qkv [B, S, 3, H, 64] bf16 with q / k already QK-LayerNorm'ed and rotated
    (the fused epilogue of the QKV projection; same values as gemm(...) followed by qk_norm_rope(...))."""
    lib = _lib.require_device()
    _need(x, BF16, "x"); _need(w, BF16, "w")
    rows, K = x.shape
    assert rows == B * S and w.shape[0] == 3 * H * 64
    for t in (gq, bq, gk, bk):
        _need(t, F32, "qk norm param")
    if bias is not None:
        _need(bias, F32, "bias")
    if cos is not None:
        _need(cos, F32, "cos"); _need(sin, F32, "sin")
        assert cos.shape == (S - St, 64)
    qkv = torch.empty(B, S, 3, H, 64, dtype=BF16, device=x.device)
    check(lib.aether_gemm_qkv_norm_rope_bf16(ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(qkv), rows, K, ptr(bias), S, St,
                                             H, ptr(gq), ptr(bq), ptr(gk), ptr(bk), float(eps), ptr(cos), ptr(sin),
                                             f16_from_col, current_stream()), "gemm_qkv_norm_rope_bf16")
    return qkv


@device_guard
def small_m_linear(x, w, bias=None, act=0):
    lib = _lib.require_device()
    _need(x, F32, "x"); _need(w, BF16, "w")
    Bn, K = x.shape
    N = w.shape[0]
    y = torch.empty(Bn, N, dtype=F32, device=x.device)
    check(lib.aether_small_m_linear(ptr(x), ptr(w), ptr(bias), ptr(y), Bn, N, K, act, current_stream()),
          "small_m_linear")
    return y


@device_guard
def timestep_sinusoid(t, dim, flip_sin_to_cos=True, freq_shift=0.0):
    lib = _lib.require_device()
    _need(t, torch.int64, "timesteps")
    emb = torch.empty(t.shape[0], dim, dtype=F32, device=t.device)
    check(lib.aether_timestep_sinusoid(ptr(t), ptr(emb), t.shape[0], dim, int(flip_sin_to_cos), float(freq_shift),
                                       current_stream()), "timestep_sinusoid")
    return emb


@device_guard
def patchify(x):
    lib = _lib.require_device()
    _need(x, BF16, "x")
    B, F, Cc, H, W = x.shape
    out = torch.empty(B * F * (H // 2) * (W // 2), Cc * 4, dtype=BF16, device=x.device)
    check(lib.aether_patchify(ptr(x), ptr(out), B, F, Cc, H, W, current_stream()), "patchify")
    return out


@device_guard
def unpatchify(tok, B, F, Cc, H, W):
    lib = _lib.require_device()
    _need(tok, BF16, "tok")
    out = torch.empty(B, F, Cc, H, W, dtype=BF16, device=tok.device)
    check(lib.aether_unpatchify(ptr(tok), tok.stride(0), ptr(out), B, F, Cc, H, W, current_stream()), "unpatchify")
    return out


def dpm_coeffs(sqrt_alpha, sqrt_one_minus_alpha, m1, m2, m3, m4, m_noise, second_order, prediction_type=0):
    z = lambda v: 0.0 if v is None else float(v)
    return _lib.DpmCoeffs(z(sqrt_alpha), z(sqrt_one_minus_alpha), z(m1), z(m2), z(m3), z(m4), z(m_noise),
                          int(second_order), int(prediction_type))


@device_guard
def cfg_dpm_step(model_out, sample, coeffs, noise1, noise2=None, old_x0=None, guidance=1.0, want_prev_f32=False):
    """model_out [n_cfg, ...] bf16 or fp32; sample bf16 -> (prev_bf16, prev_f32 | None, x0_f32)."""
    lib = _lib.require_device()
    _need(sample, BF16, "sample"); _need(noise1, BF16, "noise1")
    N = sample.numel()
    n_cfg = model_out.numel() // N
    assert model_out.numel() == n_cfg * N and n_cfg in (1, 2) and model_out.is_contiguous()
    prev = torch.empty_like(sample)
    prev32 = torch.empty(sample.shape, dtype=F32, device=sample.device) if want_prev_f32 else None
    x0 = torch.empty(sample.shape, dtype=F32, device=sample.device)
    check(lib.aether_cfg_dpm_step(ptr(model_out), int(model_out.dtype == F32), n_cfg, float(guidance), ptr(sample),
                                  ptr(old_x0), ptr(noise1), ptr(noise2), C.byref(coeffs), ptr(prev), ptr(prev32),
                                  ptr(x0), N, current_stream()), "cfg_dpm_step")
    return prev, prev32, x0


@device_guard
def resize_bilinear_u8(frames: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """frames uint8 CUDA [T, h, w, 3] -> uint8 [T, H, W, 3]: cv2.resize(frame, (W, H)) (INTER_LINEAR), bit for bit."""
    lib = _lib.require_device()
    _need(frames, torch.uint8, "frames")
    T, h, w, c = frames.shape
    assert c == 3
    out = torch.empty(T, H, W, 3, dtype=torch.uint8, device=frames.device)
    check(lib.aether_resize_bilinear_u8(ptr(frames), ptr(out), T, h, w, H, W, current_stream()), "resize_bilinear_u8")
    return out


@device_guard
def u8_frames_to_model_input(frames: torch.Tensor) -> torch.Tensor:
    """uint8 CUDA crop [F, H, W, 3] (any frame / row strides, contiguous pixels) -> bf16 [F, 3, H, W] = 2 * x / 255 - 1."""
    lib = _lib.require_device()
    if frames.dtype != torch.uint8 or not frames.is_cuda or frames.dim() != 4 or frames.shape[3] != 3 \
            or frames.stride(3) != 1 or frames.stride(2) != 3:
        raise ValueError("u8_frames_to_model_input: expected a uint8 CUDA [F, H, W, 3] view with contiguous pixels")
    F, H, W, _ = frames.shape
    out = torch.empty(F, 3, H, W, dtype=BF16, device=frames.device)
    check(lib.aether_u8_frames_to_model_input(ptr(frames), frames.stride(0), frames.stride(1), ptr(out), F, H, W,
                                              current_stream()), "u8_frames_to_model_input")
    return out
