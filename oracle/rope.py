"""Oracle (test infrastructure): 3D rotary table of the Aether pipeline.

Restates /root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py:25-144
(`get_3d_rotary_pos_embed`, with the Aether-specific `fps_factor` on the temporal
grid, :81-90) and :148-163 (`get_resize_crop_region_for_grid`), plus the
third-party helper it calls at :108-111, diffusers `get_1d_rotary_pos_embed`
(use_real=True branch).

PINNED: tests/golden/make_golden.py imports the reference file in this container
(under a diffusers shim that only supplies `get_1d_rotary_pos_embed` = the
function below) and commits its cos/sin tables; tests/test_oracle_rope.py
compares.
"""
from __future__ import annotations

import numpy as np
import torch


def get_1d_rotary_pos_embed(dim: int, pos, theta: float = 10000.0):
    """diffusers.models.embeddings.get_1d_rotary_pos_embed(use_real=True,
    linear_factor=1, ntk_factor=1, repeat_interleave_real=True)."""
    assert dim % 2 == 0
    if isinstance(pos, int):
        pos = torch.arange(pos)
    if isinstance(pos, np.ndarray):
        pos = torch.from_numpy(pos)
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    freqs = torch.outer(pos.to(torch.float32), freqs)
    cos = freqs.cos().repeat_interleave(2, dim=1).float()
    sin = freqs.sin().repeat_interleave(2, dim=1).float()
    return cos, sin


def get_resize_crop_region_for_grid(src, tgt_width, tgt_height):
    """reference :148-163."""
    tw, th = tgt_width, tgt_height
    h, w = src
    r = h / w
    if r > (th / tw):
        resize_height = th
        resize_width = int(round(th / h * w))
    else:
        resize_width = tw
        resize_height = int(round(tw / w * h))
    crop_top = int(round((th - resize_height) / 2.0))
    crop_left = int(round((tw - resize_width) / 2.0))
    return (crop_top, crop_left), (crop_top + resize_height, crop_left + resize_width)


def get_3d_rotary_pos_embed(embed_dim, crops_coords, grid_size, temporal_size,
                            theta: int = 10000, fps_factor: float = 1.0):
    """reference :25-144, grid_type == "linspace" (the CogVideoX-1.0 branch that
    is live for Aether: patch_size_t is None, :320-332)."""
    start, stop = crops_coords
    gh, gw = grid_size
    grid_h = torch.linspace(start[0], stop[0] * (gh - 1) / gh, gh, dtype=torch.float32)
    grid_w = torch.linspace(start[1], stop[1] * (gw - 1) / gw, gw, dtype=torch.float32)
    grid_t = torch.linspace(0, temporal_size * (temporal_size - 1) / temporal_size,
                            temporal_size, dtype=torch.float32) * fps_factor
    dim_t = embed_dim // 4
    dim_h = embed_dim // 8 * 3
    dim_w = embed_dim // 8 * 3
    t_cos, t_sin = get_1d_rotary_pos_embed(dim_t, grid_t, theta)
    h_cos, h_sin = get_1d_rotary_pos_embed(dim_h, grid_h, theta)
    w_cos, w_sin = get_1d_rotary_pos_embed(dim_w, grid_w, theta)

    def combine(ft, fh, fw):
        ft = ft[:, None, None, :].expand(-1, gh, gw, -1)
        fh = fh[None, :, None, :].expand(temporal_size, -1, gw, -1)
        fw = fw[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([ft, fh, fw], dim=-1).reshape(temporal_size * gh * gw, -1)

    return combine(t_cos, h_cos, w_cos), combine(t_sin, h_sin, w_sin)


def prepare_rotary_positional_embeddings(height, width, num_latent_frames, *, patch_size=2,
                                         vae_scale_factor_spatial=8, sample_height=60,
                                         sample_width=90, attention_head_dim=64,
                                         base_fps=12, fps=12):
    """reference :299-348 (`_prepare_rotary_positional_embeddings`, p_t is None)."""
    grid_h = height // (vae_scale_factor_spatial * patch_size)
    grid_w = width // (vae_scale_factor_spatial * patch_size)
    base_w = sample_width // patch_size
    base_h = sample_height // patch_size
    crops = get_resize_crop_region_for_grid((grid_h, grid_w), base_w, base_h)
    return get_3d_rotary_pos_embed(attention_head_dim, crops, (grid_h, grid_w),
                                   num_latent_frames, fps_factor=base_fps / fps)


def apply_rotary_emb(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """diffusers apply_rotary_emb(use_real=True, use_real_unbind_dim=-1): interleaved
    pairs.  x: [B, H, S, D]; cos/sin: [S, D]."""
    cos = cos[None, None]
    sin = sin[None, None]
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)
