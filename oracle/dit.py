"""Oracle (test infrastructure): fp32 restatement of `CogVideoXTransformer3DModel.forward`.

The reference calls this third-party module once per scheduler step at
/root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py:865-875 with
(hidden_states [B,f,96,h,w], encoder_hidden_states [B,226,4096], timestep [B],
ofs=None, image_rotary_emb=(cos,sin) [Sv,64], return_dict=False).

The module itself lives in `diffusers>=0.32.2` (requirements.txt:4), which is
not in /root/reference and not installable here.  PARITY UNPINNED: this file
restates the published diffusers v0.32 algorithm:
  diffusers/models/transformers/cogvideox_transformer_3d.py  (model + block)
  diffusers/models/attention_processor.py::CogVideoXAttnProcessor2_0
  diffusers/models/normalization.py::{CogVideoXLayerNormZero, AdaLayerNorm}
  diffusers/models/embeddings.py::{CogVideoXPatchEmbed, Timesteps, TimestepEmbedding}
  diffusers/models/attention.py::FeedForward (gelu-approximate)
Parameter names equal the diffusers state-dict keys so that real
`AetherWorldModel/AetherV1/transformer` safetensors load unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .rope import apply_rotary_emb


@dataclass
class DiTConfig:
    """Field names = CogVideoXTransformer3DModel.config keys.  Defaults = the
    Aether geometry (CogVideoX-5b-I2V with in=96/out=56)."""
    num_attention_heads: int = 48
    attention_head_dim: int = 64
    in_channels: int = 96
    out_channels: int = 56
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    num_layers: int = 42
    sample_width: int = 90
    sample_height: int = 60
    sample_frames: int = 41
    patch_size: int = 2
    patch_size_t: Optional[int] = None
    temporal_compression_ratio: int = 4
    max_text_seq_length: int = 226
    activation_fn: str = "gelu-approximate"
    timestep_activation_fn: str = "silu"
    norm_elementwise_affine: bool = True
    norm_eps: float = 1e-5
    attention_bias: bool = True
    use_rotary_positional_embeddings: bool = True
    use_learned_positional_embeddings: bool = False
    ofs_embed_dim: Optional[int] = None
    ff_mult: int = 4

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    def to_dict(self):
        return asdict(self)


def tiny_config(**kw) -> DiTConfig:
    """Small geometry for seconds-scale CPU parity cases (same algorithm)."""
    base = dict(num_attention_heads=4, attention_head_dim=64, num_layers=2, time_embed_dim=64,
                text_embed_dim=128, max_text_seq_length=18, sample_width=20, sample_height=12,
                sample_frames=17)
    base.update(kw)
    return DiTConfig(**base)


def timestep_sinusoid(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool, shift: float,
                      max_period: float = 10000.0) -> torch.Tensor:
    """diffusers get_timestep_embedding (scale=1): fp32 `exponent`, fp32 angle `t * exp(exponent)`, fp32 sin/cos.

    The three transcendentals are evaluated in float64 on the fp32 operand and rounded back to fp32, i.e. the
    correctly rounded fp32 result.  torch's own fp32 exp/sin/cos go through a per-ISA vector math library whose last
    bit differs between hosts, and a 1-ulp change of a frequency moves the angle of t = 999 by 6e-5 -- enough to make
    goldens generated from this oracle host-dependent (observed: MKL AVX2 vs AVX-512 paths)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - shift)
    freq = torch.exp(exponent.double()).float()
    emb = timesteps[:, None].float() * freq[None, :]
    emb = torch.cat([torch.sin(emb.double()), torch.cos(emb.double())], dim=-1).float()
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, out_dim)
        self.linear_2 = nn.Linear(out_dim, out_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class PatchEmbed(nn.Module):
    """CogVideoXPatchEmbed with patch_size_t=None (the live branch, reference :320)."""

    def __init__(self, cfg: DiTConfig):
        super().__init__()
        D = cfg.inner_dim
        self.patch_size = cfg.patch_size
        self.proj = nn.Conv2d(cfg.in_channels, D, kernel_size=cfg.patch_size, stride=cfg.patch_size, bias=True)
        self.text_proj = nn.Linear(cfg.text_embed_dim, D)

    def forward(self, text, video):
        B, f, C, H, W = video.shape
        text = self.text_proj(text)
        x = self.proj(video.reshape(B * f, C, H, W))          # [B*f, D, H/p, W/p]
        x = x.view(B, f, *x.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)   # [B, f*h*w, D]
        return torch.cat([text, x], dim=1)


class LayerNormZero(nn.Module):
    """CogVideoXLayerNormZero."""

    def __init__(self, cond_dim, dim, eps, affine):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=affine)

    def forward(self, h, e, temb):
        shift, scale, gate, eshift, escale, egate = self.linear(F.silu(temb)).chunk(6, dim=1)
        h = self.norm(h) * (1 + scale)[:, None, :] + shift[:, None, :]
        e = self.norm(e) * (1 + escale)[:, None, :] + eshift[:, None, :]
        return h, e, gate[:, None, :], egate[:, None, :]


class Attention(nn.Module):
    """diffusers Attention(qk_norm="layer_norm", eps=1e-6, bias=True, out_bias=True)
    + CogVideoXAttnProcessor2_0."""

    def __init__(self, cfg: DiTConfig):
        super().__init__()
        D = cfg.inner_dim
        self.heads = cfg.num_attention_heads
        self.dh = cfg.attention_head_dim
        self.norm_q = nn.LayerNorm(self.dh, eps=1e-6, elementwise_affine=True)
        self.norm_k = nn.LayerNorm(self.dh, eps=1e-6, elementwise_affine=True)
        self.to_q = nn.Linear(D, D, bias=cfg.attention_bias)
        self.to_k = nn.Linear(D, D, bias=cfg.attention_bias)
        self.to_v = nn.Linear(D, D, bias=cfg.attention_bias)
        self.to_out = nn.ModuleList([nn.Linear(D, D, bias=True), nn.Dropout(0.0)])

    def forward(self, h, e, rope):
        St = e.shape[1]
        x = torch.cat([e, h], dim=1)
        B, S, _ = x.shape
        q = self.to_q(x).view(B, S, self.heads, self.dh).transpose(1, 2)
        k = self.to_k(x).view(B, S, self.heads, self.dh).transpose(1, 2)
        v = self.to_v(x).view(B, S, self.heads, self.dh).transpose(1, 2)
        q = self.norm_q(q)
        k = self.norm_k(k)
        if rope is not None:
            cos, sin = rope
            q = torch.cat([q[:, :, :St], apply_rotary_emb(q[:, :, St:], cos, sin)], dim=2)
            k = torch.cat([k[:, :, :St], apply_rotary_emb(k[:, :, St:], cos, sin)], dim=2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, S, self.heads * self.dh)
        o = self.to_out[0](o)
        return o[:, St:], o[:, :St]


class GELUProj(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    """diffusers FeedForward(activation_fn="gelu-approximate", final_dropout=True):
    net = [GELU(tanh) w/ proj, Dropout, Linear, Dropout]."""

    def __init__(self, dim, mult):
        super().__init__()
        self.net = nn.ModuleList([GELUProj(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim), nn.Dropout(0.0)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class Block(nn.Module):
    """CogVideoXBlock."""

    def __init__(self, cfg: DiTConfig):
        super().__init__()
        D = cfg.inner_dim
        self.norm1 = LayerNormZero(cfg.time_embed_dim, D, cfg.norm_eps, cfg.norm_elementwise_affine)
        self.attn1 = Attention(cfg)
        self.norm2 = LayerNormZero(cfg.time_embed_dim, D, cfg.norm_eps, cfg.norm_elementwise_affine)
        self.ff = FeedForward(D, cfg.ff_mult)

    def forward(self, h, e, temb, rope):
        St = e.shape[1]
        nh, ne, gate, egate = self.norm1(h, e, temb)
        ah, ae = self.attn1(nh, ne, rope)
        h = h + gate * ah
        e = e + egate * ae
        nh, ne, gate, egate = self.norm2(h, e, temb)
        ff = self.ff(torch.cat([ne, nh], dim=1))
        h = h + gate * ff[:, St:]
        e = e + egate * ff[:, :St]
        return h, e


class AdaLayerNorm(nn.Module):
    """diffusers AdaLayerNorm(chunk_dim=1): shift first, then scale."""

    def __init__(self, cond_dim, dim, eps, affine):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 2 * dim)
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=affine)

    def forward(self, x, temb):
        shift, scale = self.linear(F.silu(temb)).chunk(2, dim=1)
        return self.norm(x) * (1 + scale[:, None, :]) + shift[:, None, :]


class OracleDiT(nn.Module):
    """CogVideoXTransformer3DModel (rotary / "5B" branch)."""

    def __init__(self, cfg: DiTConfig):
        super().__init__()
        self.config = cfg
        D = cfg.inner_dim
        self.patch_embed = PatchEmbed(cfg)
        self.time_embedding = TimestepEmbedding(D, cfg.time_embed_dim)
        self.transformer_blocks = nn.ModuleList([Block(cfg) for _ in range(cfg.num_layers)])
        self.norm_final = nn.LayerNorm(D, eps=cfg.norm_eps, elementwise_affine=cfg.norm_elementwise_affine)
        self.norm_out = AdaLayerNorm(cfg.time_embed_dim, D, cfg.norm_eps, cfg.norm_elementwise_affine)
        self.proj_out = nn.Linear(D, cfg.patch_size * cfg.patch_size * cfg.out_channels)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, hidden_states, encoder_hidden_states, timestep, ofs=None,
                image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                attention_kwargs=None, return_dict: bool = False, n_layers: Optional[int] = None):
        cfg = self.config
        B, f, C, H, W = hidden_states.shape
        t_emb = timestep_sinusoid(timestep, cfg.inner_dim, cfg.flip_sin_to_cos, cfg.freq_shift)
        emb = self.time_embedding(t_emb.to(hidden_states.dtype))
        x = self.patch_embed(encoder_hidden_states, hidden_states)
        St = encoder_hidden_states.shape[1]
        e, h = x[:, :St], x[:, St:]
        blocks = self.transformer_blocks if n_layers is None else self.transformer_blocks[:n_layers]
        for blk in blocks:
            h, e = blk(h, e, emb, image_rotary_emb)
        x = self.norm_final(torch.cat([e, h], dim=1))[:, St:]
        x = self.norm_out(x, emb)
        x = self.proj_out(x)
        p = cfg.patch_size
        out = x.reshape(B, f, H // p, W // p, -1, p, p)
        out = out.permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
        return (out,)


def seeded_init_(model: nn.Module, seed: int = 0, std: float = 0.02) -> nn.Module:
    """SURVEY.md 8(d) synthetic weights: N(0, std^2) for every weight matrix and
    bias (AdaLN linears included, so gates are live); LayerNorm gamma ~ 1 + N(0, 0.1^2),
    beta ~ N(0, 0.05^2) so that the affine paths are exercised by parity tests."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            is_norm = (".norm." in name or name.startswith("norm_final") or ".norm_q" in name or ".norm_k" in name
                       or name.endswith("norm.weight") or name.endswith("norm.bias"))
            if is_norm and p.ndim == 1:
                if name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel() if p.ndim > 1 else 1
                s = std if p.ndim == 1 else min(std * 4, 1.0 / math.sqrt(fan_in))
                p.copy_(s * torch.randn(p.shape, generator=g))
    return model
