"""3-D sin-cos positional table of diffusers' CogVideoXPatchEmbed (`_get_positional_embeddings` ->
`get_3d_sincos_pos_embed`), used only when `use_learned_positional_embeddings` is set and the latent
geometry differs from the sample geometry.  Host-side, once per geometry."""
from __future__ import annotations

import numpy as np
import torch


def _sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_3d_joint(embed_dim: int, grid_w: int, grid_h: int, frames: int, text_len: int,
                    spatial_scale: float = 1.875, temporal_scale: float = 1.0) -> torch.Tensor:
    """[text_len + frames*grid_h*grid_w, embed_dim] fp32; zeros on the text rows."""
    d_sp, d_t = 3 * embed_dim // 4, embed_dim // 4
    gh = np.arange(grid_h, dtype=np.float32) / spatial_scale
    gw = np.arange(grid_w, dtype=np.float32) / spatial_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_h, grid_w)
    emb_sp = np.concatenate([_sincos_1d(d_sp // 2, grid[0]), _sincos_1d(d_sp // 2, grid[1])], axis=1)
    emb_t = _sincos_1d(d_t, np.arange(frames, dtype=np.float32) / temporal_scale)
    emb_sp = np.repeat(emb_sp[None], frames, axis=0)
    emb_t = np.repeat(emb_t[:, None], grid_h * grid_w, axis=1)
    pos = np.concatenate([emb_t, emb_sp], axis=-1).reshape(frames * grid_h * grid_w, embed_dim)
    out = torch.zeros(text_len + pos.shape[0], embed_dim, dtype=torch.float32)
    out[text_len:] = torch.from_numpy(pos.astype(np.float32))
    return out
