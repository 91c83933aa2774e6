"""3-D rotary table with the Aether fps factor (once per pipeline call; tiny).

Mirrors /root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py:25-144 (`get_3d_rotary_pos_embed`,
"linspace" grid, fps_factor on the temporal grid :81-90), :148-163 (`get_resize_crop_region_for_grid`) and
:299-348 (`_prepare_rotary_positional_embeddings`, CogVideoX-1.0 branch) including the diffusers helper
`get_1d_rotary_pos_embed(use_real=True)` it calls at :108-111.
"""
from __future__ import annotations

from typing import Tuple

import torch


def get_resize_crop_region_for_grid(src, tgt_width, tgt_height):
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


def _rope_1d(dim: int, pos: torch.Tensor, theta: float):
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32, device=pos.device)[: dim // 2] / dim))
    ang = torch.outer(pos, freqs)
    return ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float()


def get_3d_rotary_pos_embed(embed_dim: int, crops_coords, grid_size, temporal_size: int, theta: int = 10000,
                            device=None, fps_factor: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    start, stop = crops_coords
    gh, gw = grid_size
    kw = dict(device=device, dtype=torch.float32)
    grid_h = torch.linspace(start[0], stop[0] * (gh - 1) / gh, gh, **kw)
    grid_w = torch.linspace(start[1], stop[1] * (gw - 1) / gw, gw, **kw)
    grid_t = torch.linspace(0, temporal_size * (temporal_size - 1) / temporal_size, temporal_size, **kw) * fps_factor
    dim_t, dim_h, dim_w = embed_dim // 4, embed_dim // 8 * 3, embed_dim // 8 * 3
    (tc, ts), (hc, hs), (wc, ws) = _rope_1d(dim_t, grid_t, theta), _rope_1d(dim_h, grid_h, theta), _rope_1d(dim_w, grid_w, theta)

    def combine(ft, fh, fw):
        ft = ft[:, None, None, :].expand(-1, gh, gw, -1)
        fh = fh[None, :, None, :].expand(temporal_size, -1, gw, -1)
        fw = fw[None, None, :, :].expand(temporal_size, gh, -1, -1)
        return torch.cat([ft, fh, fw], dim=-1).reshape(temporal_size * gh * gw, -1).contiguous()

    return combine(tc, hc, wc), combine(ts, hs, ws)


def prepare_rotary_positional_embeddings(height: int, width: int, num_latent_frames: int, *, patch_size: int,
                                         vae_scale_factor_spatial: int, sample_height: int, sample_width: int,
                                         attention_head_dim: int, base_fps: int, fps: int, device=None):
    grid_h = height // (vae_scale_factor_spatial * patch_size)
    grid_w = width // (vae_scale_factor_spatial * patch_size)
    crops = get_resize_crop_region_for_grid((grid_h, grid_w), sample_width // patch_size, sample_height // patch_size)
    return get_3d_rotary_pos_embed(attention_head_dim, crops, (grid_h, grid_w), num_latent_frames, device=device,
                                   fps_factor=base_fps / fps)
