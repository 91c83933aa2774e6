"""`CogVideoXTransformer3DModel`-shaped module whose forward is the C-ABI call `aether_dit_forward`.

Drop-in for the `transformer` object the reference pipeline is constructed with
(/root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py:274-288) and calls at :865-875:

    transformer(hidden_states=[B,f,96,h,w] bf16, encoder_hidden_states=[B,226,4096] bf16, timestep=[B],
                ofs=None, image_rotary_emb=(cos, sin) [Sv,64] fp32, attention_kwargs=None,
                return_dict=False) -> (Tensor[B,f,56,h,w] bf16,)

and reads `.config.{patch_size, patch_size_t, sample_width, sample_height, sample_frames,
attention_head_dim, use_rotary_positional_embeddings, ofs_embed_dim}` (:308-338, :545, :722-728, :808, :815).

Parameter names equal the diffusers state-dict keys (transformer_blocks.N.attn1.to_q.weight, ...), so
`load_state_dict` of real `AetherWorldModel/AetherV1` transformer weights works unchanged; `pack()` then
re-lays them out for the kernels (fused QKV, one concatenated AdaLN matrix, fp32 vectors) and builds the
native handle.  No torch compute happens in `forward`; without the CUDA library it raises.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import DitConfig, DitLayerWeights, DitWeights, check, current_stream, device_guard, ptr

_DEFAULTS = dict(
    num_attention_heads=48, attention_head_dim=64, in_channels=96, out_channels=56, flip_sin_to_cos=True,
    freq_shift=0, time_embed_dim=512, text_embed_dim=4096, num_layers=42, sample_width=90, sample_height=60,
    sample_frames=41, patch_size=2, patch_size_t=None, temporal_compression_ratio=4, max_text_seq_length=226,
    activation_fn="gelu-approximate", timestep_activation_fn="silu", norm_elementwise_affine=True, norm_eps=1e-5,
    attention_bias=True, use_rotary_positional_embeddings=True, use_learned_positional_embeddings=False,
    ofs_embed_dim=None, ff_mult=4, spatial_interpolation_scale=1.875, temporal_interpolation_scale=1.0,
)


class _Config(SimpleNamespace):
    """Attribute + mapping access, like diffusers' FrozenDict config."""

    def __getitem__(self, k):
        return getattr(self, k)

    def get(self, k, default=None):
        return getattr(self, k, default)

    def to_dict(self):
        return dict(self.__dict__)


class _Holder(nn.Module):
    """Parameter holder (never executed)."""


def _linear(i, o, bias=True, **kw):
    m = _Holder()
    m.weight = nn.Parameter(torch.empty(o, i, **kw), requires_grad=False)
    if bias:
        m.bias = nn.Parameter(torch.empty(o, **kw), requires_grad=False)
    return m


def _norm(d, affine=True, **kw):
    m = _Holder()
    if affine:
        m.weight = nn.Parameter(torch.ones(d, **kw), requires_grad=False)
        m.bias = nn.Parameter(torch.zeros(d, **kw), requires_grad=False)
    return m


class AetherTransformer3D(nn.Module):
    def __init__(self, device=None, dtype=torch.bfloat16, **config):
        super().__init__()
        cfg = dict(_DEFAULTS)
        unknown = set(config) - set(cfg)
        if unknown:
            raise ValueError(f"unknown config keys: {sorted(unknown)}")
        cfg.update(config)
        self.config = _Config(**cfg)
        c = self.config
        if c.patch_size_t is not None:
            raise NotImplementedError("CogVideoX-1.5 (patch_size_t) geometry is not on the Aether path (reference :320)")
        if c.ofs_embed_dim is not None:
            raise NotImplementedError("ofs embedding is not used by Aether (reference :813-817)")
        if c.activation_fn != "gelu-approximate" or c.timestep_activation_fn != "silu":
            raise NotImplementedError("only gelu-approximate / silu (CogVideoX-5b family)")
        kw = dict(device=device, dtype=dtype)
        D = c.num_attention_heads * c.attention_head_dim
        T = c.time_embed_dim
        aff = c.norm_elementwise_affine
        self.patch_embed = _Holder()
        self.patch_embed.proj = _Holder()
        self.patch_embed.proj.weight = nn.Parameter(torch.empty(D, c.in_channels, c.patch_size, c.patch_size, **kw),
                                                    requires_grad=False)
        self.patch_embed.proj.bias = nn.Parameter(torch.empty(D, **kw), requires_grad=False)
        self.patch_embed.text_proj = _linear(c.text_embed_dim, D, **kw)
        self.time_embedding = _Holder()
        self.time_embedding.linear_1 = _linear(D, T, **kw)
        self.time_embedding.linear_2 = _linear(T, T, **kw)
        blocks = []
        for _ in range(c.num_layers):
            b = _Holder()
            b.norm1 = _Holder(); b.norm1.linear = _linear(T, 6 * D, **kw); b.norm1.norm = _norm(D, aff, **kw)
            b.attn1 = _Holder()
            b.attn1.norm_q = _norm(c.attention_head_dim, True, **kw)
            b.attn1.norm_k = _norm(c.attention_head_dim, True, **kw)
            b.attn1.to_q = _linear(D, D, c.attention_bias, **kw)
            b.attn1.to_k = _linear(D, D, c.attention_bias, **kw)
            b.attn1.to_v = _linear(D, D, c.attention_bias, **kw)
            b.attn1.to_out = nn.ModuleList([_linear(D, D, **kw)])
            b.norm2 = _Holder(); b.norm2.linear = _linear(T, 6 * D, **kw); b.norm2.norm = _norm(D, aff, **kw)
            b.ff = _Holder()
            n0 = _Holder(); n0.proj = _linear(D, c.ff_mult * D, **kw)
            b.ff.net = nn.ModuleList([n0, _Holder(), _linear(c.ff_mult * D, D, **kw)])
            blocks.append(b)
        self.transformer_blocks = nn.ModuleList(blocks)
        self.norm_final = _norm(D, aff, **kw)
        self.norm_out = _Holder(); self.norm_out.linear = _linear(T, 2 * D, **kw); self.norm_out.norm = _norm(D, aff, **kw)
        self.proj_out = _linear(D, c.patch_size * c.patch_size * c.out_channels, **kw)
        if c.use_learned_positional_embeddings:
            joint = c.max_text_seq_length + (c.sample_height // c.patch_size) * (c.sample_width // c.patch_size) * (
                (c.sample_frames - 1) // c.temporal_compression_ratio + 1)
            self.patch_embed.register_buffer("pos_embedding", torch.zeros(1, joint, D, **kw), persistent=True)
        self._handle = None
        self._packed = None
        self._ws = None
        self._n_layers_override = -1
        # kernel-mode switch (not a model hyper-parameter), set before pack(): attention kernel variant
        # (csrc/attention*_tcgen05.cu, DESIGN.md "Attention roofline"): 5 = decoupled S/P buffers + skewed MMA schedule
        # (default: 3.31 ms isolated / 3.86 ms inside the power-capped step at S=15076), 2 = fp16 P/V (3.21 ms isolated but
        # 3.88 ms in-step: no gain), 0 = baseline.  AETHER_ATTENTION_MODE overrides it.
        import os
        self.attention_fp16_pv = int(os.environ.get("AETHER_ATTENTION_MODE", "5"))
        # QK-LayerNorm + RoPE inside the QKV GEMM epilogue (one launch less per layer; measured slower: off).  Read here,
        # once per module, and stored in the handle's config -- the C library itself never reads the environment.
        self.fused_qkv_epilogue = os.environ.get("AETHER_FUSED_QK", "0")[:1] == "1"
        # attention variant 5 cuts the partially filled tail wave of its grid along the keys (scratch in the workspace);
        # off => one launch, and the output of a batch item no longer depends on the batch size (fp32 merge order)
        self.attention_split_tail = os.environ.get("AETHER_ATTENTION_SPLIT_TAIL", "1")[:1] != "0"

    # ------------------------------------------------------------------ torch plumbing
    @property
    def dtype(self):
        return torch.bfloat16

    @property
    def device(self):
        if self._packed is not None:
            return self._packed["w_adaln"].device
        return self.proj_out.weight.device

    @property
    def inner_dim(self):
        return self.config.num_attention_heads * self.config.attention_head_dim

    def _apply(self, fn, *a, **k):
        dev_before = self.proj_out.weight.device
        r = super()._apply(fn, *a, **k)
        if self._handle is not None and self.proj_out.weight.device != dev_before:
            self.release()                # parameters moved to another device: packed buffers are stale
        return r

    def to(self, *args, **kwargs):
        """Moving an already packed module to the device it lives on is a no-op (pack(release_unpacked=True) drops
        the unfused q/k/v weights, so a re-pack must never be triggered by `pipeline.to(same_device)`)."""
        if self._handle is not None:
            dev = None
            for a in list(args) + list(kwargs.values()):
                if isinstance(a, (str, torch.device)):
                    dev = torch.device(a)
            if dev is None or (dev.type == self.device.type and (dev.index is None or dev.index == self.device.index)):
                return self
            if getattr(self, "_released_unpacked", False):
                raise RuntimeError("this module was packed with release_unpacked=True and cannot be moved")
        return super().to(*args, **kwargs)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def release(self):
        if self._handle is not None:
            _lib.load().aether_dit_destroy(self._handle)
        self._handle = None
        self._packed = None
        self._ws = None

    # ------------------------------------------------------------------ synthetic weights (bench / smoke)
    @torch.no_grad()
    def init_synthetic_(self, seed: int = 0):
        """seeded N(0, sigma^2) weights with sigma = 1/sqrt(fan_in) capped at 0.08, small biases,
        LayerNorm gamma ~ 1 + 0.1 N, beta ~ 0.05 N.  Generated on the parameters' own device."""
        dev = self.proj_out.weight.device
        g = torch.Generator(device=dev).manual_seed(seed)
        for name, p in sorted(self.named_parameters()):
            is_norm = ".norm" in name or name.startswith("norm_final")
            if p.ndim == 1 and is_norm and "linear" not in name:
                if name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32))
            elif p.ndim == 1:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32))
            else:
                fan_in = p[0].numel()
                s = min(0.08, 1.0 / math.sqrt(fan_in))
                # chunked to bound the fp32 temporary for the 12288 x 3072 matrices
                flat = p.view(-1)
                step = 1 << 26
                for i in range(0, flat.numel(), step):
                    n = min(step, flat.numel() - i)
                    flat[i:i + n].copy_(s * torch.randn(n, generator=g, device=dev, dtype=torch.float32))
        return self

    # ------------------------------------------------------------------ packing + native handle
    @torch.no_grad()
    @device_guard
    def pack(self, release_unpacked: bool = False):
        """Lay the weights out for the kernels and create the native handle (idempotent)."""
        lib = _lib.require_device()
        if self._handle is not None:
            return self
        c = self.config
        dev = self.proj_out.weight.device
        if dev.type != "cuda":
            raise RuntimeError("AetherTransformer3D.pack(): move the module to a CUDA device first (no CPU path)")
        D = self.inner_dim
        bf = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        ones = torch.ones(D, device=dev, dtype=torch.float32)
        zeros = torch.zeros(D, device=dev, dtype=torch.float32)

        def norm_gb(m):
            return (f32(m.weight), f32(m.bias)) if hasattr(m, "weight") else (ones, zeros)

        def bias_or_zero(m, n):
            return f32(m.bias) if hasattr(m, "bias") else torch.zeros(n, device=dev, dtype=torch.float32)

        P = {}
        P["w_time1"], P["b_time1"] = bf(self.time_embedding.linear_1.weight), f32(self.time_embedding.linear_1.bias)
        P["w_time2"], P["b_time2"] = bf(self.time_embedding.linear_2.weight), f32(self.time_embedding.linear_2.bias)
        P["w_text"], P["b_text"] = bf(self.patch_embed.text_proj.weight), f32(self.patch_embed.text_proj.bias)
        P["w_patch"] = bf(self.patch_embed.proj.weight.reshape(D, -1))
        P["b_patch"] = f32(self.patch_embed.proj.bias)
        L = c.num_layers
        T = c.time_embed_dim
        w_adaln = torch.empty((12 * L + 2) * D, T, device=dev, dtype=torch.bfloat16)
        b_adaln = torch.empty((12 * L + 2) * D, device=dev, dtype=torch.float32)
        layers = (DitLayerWeights * L)()
        keep = []
        for i, blk in enumerate(self.transformer_blocks):
            r0 = i * 12 * D
            w_adaln[r0:r0 + 6 * D].copy_(blk.norm1.linear.weight); b_adaln[r0:r0 + 6 * D].copy_(blk.norm1.linear.bias)
            w_adaln[r0 + 6 * D:r0 + 12 * D].copy_(blk.norm2.linear.weight)
            b_adaln[r0 + 6 * D:r0 + 12 * D].copy_(blk.norm2.linear.bias)
            a = blk.attn1
            w_qkv = torch.cat([bf(a.to_q.weight), bf(a.to_k.weight), bf(a.to_v.weight)], dim=0).contiguous()
            b_qkv = torch.cat([bias_or_zero(a.to_q, D), bias_or_zero(a.to_k, D), bias_or_zero(a.to_v, D)]).contiguous()
            n1g, n1b = norm_gb(blk.norm1.norm)
            n2g, n2b = norm_gb(blk.norm2.norm)
            t = dict(w_qkv=w_qkv, b_qkv=b_qkv, w_out=bf(a.to_out[0].weight), b_out=f32(a.to_out[0].bias),
                     w_ff1=bf(blk.ff.net[0].proj.weight), b_ff1=f32(blk.ff.net[0].proj.bias),
                     w_ff2=bf(blk.ff.net[2].weight), b_ff2=f32(blk.ff.net[2].bias),
                     norm1_g=n1g, norm1_b=n1b, norm2_g=n2g, norm2_b=n2b,
                     qn_g=f32(a.norm_q.weight), qn_b=f32(a.norm_q.bias), kn_g=f32(a.norm_k.weight),
                     kn_b=f32(a.norm_k.bias))
            keep.append(t)
            for k, v in t.items():
                setattr(layers[i], k, v.data_ptr())
            if release_unpacked:
                for m in (a.to_q, a.to_k, a.to_v):
                    m.weight.data = torch.empty(0, device=dev, dtype=torch.bfloat16)
        r0 = 12 * L * D
        w_adaln[r0:].copy_(self.norm_out.linear.weight); b_adaln[r0:].copy_(self.norm_out.linear.bias)
        P["w_adaln"], P["b_adaln"] = w_adaln, b_adaln
        P["normf_g"], P["normf_b"] = norm_gb(self.norm_final)
        P["normo_g"], P["normo_b"] = norm_gb(self.norm_out.norm)
        P["w_proj"], P["b_proj"] = bf(self.proj_out.weight), f32(self.proj_out.bias)
        P["layers_keep"] = keep
        P["layers_arr"] = layers

        w = DitWeights()
        for k in ("w_time1", "b_time1", "w_time2", "b_time2", "w_text", "b_text", "w_patch", "b_patch", "w_adaln",
                  "b_adaln", "normf_g", "normf_b", "normo_g", "normo_b", "w_proj", "b_proj"):
            setattr(w, k, P[k].data_ptr())
        w.pos_embedding = 0
        w.layers = C.cast(layers, C.POINTER(DitLayerWeights))
        cfg = DitConfig(c.num_attention_heads, c.attention_head_dim, c.num_layers, c.in_channels, c.out_channels,
                        c.patch_size, c.time_embed_dim, c.text_embed_dim, int(c.flip_sin_to_cos), float(c.freq_shift),
                        float(c.norm_eps), c.ff_mult, int(self.attention_fp16_pv),
                        int(self.fused_qkv_epilogue and c.attention_head_dim == 64), int(self.attention_split_tail))
        h = C.c_void_p()
        check(lib.aether_dit_create(C.byref(cfg), C.byref(w), C.byref(h)), "dit_create")
        self._handle = h
        self._packed = P
        self._weights_struct = w
        self._released_unpacked = bool(release_unpacked)
        return self

    def _pos_embedding_for(self, St, F, H, W):
        """diffusers CogVideoXPatchEmbed adds a positional table when `use_learned_positional_embeddings` OR
        `not use_rotary_positional_embeddings` (the 2B-style fixed sin-cos branch): the stored table when the geometry
        equals the sample geometry and the table is learned, else a freshly computed 3-D sin-cos table (zeros on the
        text rows).  Returns bf16 [St+Sv, D], or None for the rotary-only configuration (AetherV1)."""
        c = self.config
        learned = bool(c.use_learned_positional_embeddings)
        if not learned and c.use_rotary_positional_embeddings:
            return None
        key = (St, F, H, W)
        cache = self._packed.setdefault("pos_cache", {})
        if key in cache:
            return cache[key]
        if learned and (H != c.sample_height or W != c.sample_width):
            raise ValueError("learned positional embeddings require the sample height/width (diffusers behaviour)")
        pre = (F - 1) * c.temporal_compression_ratio + 1
        if learned and pre == c.sample_frames:
            pos = self.patch_embed.pos_embedding[0, :St + F * (H // 2) * (W // 2)]
        else:
            from .posembed import sincos_3d_joint
            pos = sincos_3d_joint(self.inner_dim, W // 2, H // 2, F, St, c.spatial_interpolation_scale,
                                  c.temporal_interpolation_scale).to(self.device)
        cache[key] = pos.to(torch.bfloat16).contiguous()
        return cache[key]

    # ------------------------------------------------------------------ measurement hooks (bench.py)
    def enable_timing(self, enable: bool = True):
        if self._handle is None:
            self.pack()
        check(_lib.load().aether_dit_enable_timing(self._handle, int(enable)), "dit_enable_timing")

    def read_attention_timing(self):
        """(total ms, launches) of the attention kernel in the last forward; call after a stream synchronize."""
        ms, n = C.c_float(0), C.c_int32(0)
        check(_lib.load().aether_dit_read_timing(self._handle, C.byref(ms), C.byref(n)), "dit_read_timing")
        return ms.value, n.value

    def launches_per_forward(self, batch: int) -> int:
        """Kernels of this library launched by one forward (see csrc/dit_forward.cu): 8 per layer (LN-modulate, QKV,
        QK-norm+RoPE, attention, to_out, LN-modulate, FF1, FF2); 7 with AETHER_FUSED_QK=1 (QK-norm+RoPE in the QKV
        epilogue)."""
        per_layer = 7 if self.fused_qkv_epilogue and self.config.attention_head_dim == 64 else 8
        return 4 + 1 + 2 * batch + per_layer * self.config.num_layers + 1 + batch + 1

    # ------------------------------------------------------------------ loop form: concat / repeat / expand folded in
    @torch.no_grad()
    @device_guard
    def forward_split(self, latents: torch.Tensor, condition: torch.Tensor, encoder_hidden_states: torch.Tensor,
                      timestep: torch.Tensor, image_rotary_emb=None):
        """Reference :832-875 in one C-ABI call: `latents` [1 or B, F, C0, H, W] is broadcast over the batch of
        `condition` [B, F, C1, H, W] and concatenated along channels inside the patch-gather kernel; the prompt
        embedding [1 or B, St, dim] and the timestep [1 or B] broadcast likewise.  Returns [B, F, out, H, W] bf16."""
        lib = _lib.require_device()
        if self._handle is None:
            self.pack()
        c = self.config
        B, F, C1, H, W = condition.shape
        Bl, _, C0, _, _ = latents.shape
        assert C0 + C1 == c.in_channels and latents.is_cuda
        lat = latents.to(torch.bfloat16).contiguous()
        cond = condition.to(torch.bfloat16).contiguous()
        txt = encoder_hidden_states.to(torch.bfloat16).contiguous()
        ts = timestep.to(device=lat.device, dtype=torch.int64).reshape(-1).contiguous()
        St = txt.shape[1]
        cos = sin = None
        if image_rotary_emb is not None:
            cos = image_rotary_emb[0].to(device=lat.device, dtype=torch.float32).contiguous()
            sin = image_rotary_emb[1].to(device=lat.device, dtype=torch.float32).contiguous()
        check(lib.aether_dit_set_pos_embedding(self._handle, ptr(self._pos_embedding_for(St, F, H, W))),
              "dit_set_pos_embedding")
        need = lib.aether_dit_workspace_bytes(self._handle, B, F, H, W, St)
        if self._ws is None or self._ws.numel() < need or self._ws.device != lat.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=lat.device)
        out = torch.empty(B, F, c.out_channels, H, W, dtype=torch.bfloat16, device=lat.device)
        check(lib.aether_dit_forward_split(self._handle, ptr(lat), Bl, C0, ptr(cond), ptr(txt), txt.shape[0], ptr(ts),
                                           ts.numel(), ptr(cos), ptr(sin), ptr(out), B, F, H, W, St, ptr(self._ws),
                                           self._ws.numel(), self._n_layers_override, current_stream()),
              "dit_forward_split")
        return out

    # ------------------------------------------------------------------ forward = one C-ABI call
    @torch.no_grad()
    @device_guard
    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, timestep: torch.Tensor,
                timestep_cond=None, ofs=None, image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                attention_kwargs=None, return_dict: bool = False):
        lib = _lib.require_device()
        if self._handle is None:
            self.pack()
        c = self.config
        if not hidden_states.is_cuda:
            raise RuntimeError("AetherTransformer3D.forward: inputs must be CUDA tensors (no CPU path)")
        B, F, Cin, H, W = hidden_states.shape
        St = encoder_hidden_states.shape[1]
        assert Cin == c.in_channels and encoder_hidden_states.shape[2] == c.text_embed_dim
        hs = hidden_states.to(torch.bfloat16).contiguous()
        txt = encoder_hidden_states.to(torch.bfloat16).contiguous()
        ts = timestep.to(device=hs.device, dtype=torch.int64).reshape(-1).contiguous()
        if ts.numel() == 1 and B > 1:
            ts = ts.expand(B).contiguous()
        cos = sin = None
        if image_rotary_emb is not None:
            cos = image_rotary_emb[0].to(device=hs.device, dtype=torch.float32).contiguous()
            sin = image_rotary_emb[1].to(device=hs.device, dtype=torch.float32).contiguous()
        pos = self._pos_embedding_for(St, F, H, W)
        check(lib.aether_dit_set_pos_embedding(self._handle, ptr(pos)), "dit_set_pos_embedding")
        need = lib.aether_dit_workspace_bytes(self._handle, B, F, H, W, St)
        if self._ws is None or self._ws.numel() < need or self._ws.device != hs.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=hs.device)
        out = torch.empty(B, F, c.out_channels, H, W, dtype=torch.bfloat16, device=hs.device)
        check(lib.aether_dit_forward(self._handle, ptr(hs), ptr(txt), ptr(ts), ptr(cos), ptr(sin), ptr(out), B, F, H,
                                     W, St, ptr(self._ws), self._ws.numel(), self._n_layers_override,
                                     current_stream()), "dit_forward")
        if return_dict:
            return SimpleNamespace(sample=out)
        return (out,)
