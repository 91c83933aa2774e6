"""Oracle (test infrastructure): fp32 restatement of diffusers `AutoencoderKLCogVideoX`.

Reference call sites (/root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py):
  vae.encode(x).latent_dist.sample(generator)  :557-620 (via retrieve_latents :233-245)
  vae.decode(z).sample (inherited decode_latents)  :931, :936
  vae.config.{latent_channels, scaling_factor, invert_scale_latents, block_out_channels,
              temporal_compression_ratio}  :571, :843, :925-929
  vae.enable_slicing(); vae.enable_tiling()  scripts/demo.py:229-230, evaluation/.../launch_aether.py:437-438

The module lives in third-party diffusers (models/autoencoders/autoencoder_kl_cogvideox.py), absent here:
PARITY UNPINNED; this restates the published v0.32 algorithm: causal Conv3d with the
first-frame / conv_cache temporal padding, GroupNorm (encoder) / SpatialNorm3D (decoder) + SiLU resnets,
temporal avg-pool / nearest up-sampling that keeps frame 0, 3x3 spatial tiling with linear blends, and
frame batching (8 sample frames / 2 latent frames per batch, remainder folded into the first batch).
Parameter names equal the diffusers state-dict keys.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from types import SimpleNamespace
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 256, 512)
    latent_channels: int = 16
    layers_per_block: int = 3
    norm_eps: float = 1e-6
    norm_num_groups: int = 32
    temporal_compression_ratio: int = 4
    sample_height: int = 480
    sample_width: int = 720
    scaling_factor: float = 0.7
    invert_scale_latents: bool = False

    def to_dict(self):
        return asdict(self)


def tiny_vae_config(**kw) -> VAEConfig:
    base = dict(block_out_channels=(32, 64, 64, 128), layers_per_block=1, norm_num_groups=8, sample_height=96,
                sample_width=160)
    base.update(kw)
    return VAEConfig(**base)


class CausalConv3d(nn.Module):
    """CogVideoXCausalConv3d (pad_mode="constant"): time pad = first frame x (k-1) or the conv_cache."""

    def __init__(self, cin, cout, kernel_size):
        super().__init__()
        k = kernel_size
        self.kt = k
        self.pad = (k - 1) // 2
        self.conv = nn.Conv3d(cin, cout, (k, k, k), stride=1, padding=0)

    def forward(self, x, conv_cache=None):
        if self.kt > 1:
            cached = [conv_cache] if conv_cache is not None else [x[:, :, :1]] * (self.kt - 1)
            x = torch.cat(cached + [x], dim=2)
        new_cache = x[:, :, -self.kt + 1:].clone() if self.kt > 1 else None
        x = F.pad(x, (self.pad, self.pad, self.pad, self.pad), mode="constant", value=0)
        return self.conv(x), new_cache


class SpatialNorm3D(nn.Module):
    def __init__(self, f_channels, zq_channels, groups):
        super().__init__()
        self.norm_layer = nn.GroupNorm(num_channels=f_channels, num_groups=groups, eps=1e-6, affine=True)
        self.conv_y = CausalConv3d(zq_channels, f_channels, 1)
        self.conv_b = CausalConv3d(zq_channels, f_channels, 1)

    def forward(self, f, zq):
        if f.shape[2] > 1 and f.shape[2] % 2 == 1:
            f_first, f_rest = f[:, :, :1], f[:, :, 1:]
            z_first, z_rest = zq[:, :, :1], zq[:, :, 1:]
            z_first = F.interpolate(z_first, size=f_first.shape[-3:])
            z_rest = F.interpolate(z_rest, size=f_rest.shape[-3:])
            zq = torch.cat([z_first, z_rest], dim=2)
        else:
            zq = F.interpolate(zq, size=f.shape[-3:])
        y, _ = self.conv_y(zq)
        b, _ = self.conv_b(zq)
        return self.norm_layer(f) * y + b


class ResnetBlock3D(nn.Module):
    def __init__(self, cin, cout, groups, eps, spatial_norm_dim=None):
        super().__init__()
        self.cin, self.cout = cin, cout
        if spatial_norm_dim is None:
            self.norm1 = nn.GroupNorm(num_channels=cin, num_groups=groups, eps=eps)
            self.norm2 = nn.GroupNorm(num_channels=cout, num_groups=groups, eps=eps)
        else:
            self.norm1 = SpatialNorm3D(cin, spatial_norm_dim, groups)
            self.norm2 = SpatialNorm3D(cout, spatial_norm_dim, groups)
        self.spatial = spatial_norm_dim is not None
        self.conv1 = CausalConv3d(cin, cout, 3)
        self.conv2 = CausalConv3d(cout, cout, 3)
        if cin != cout:
            self.conv_shortcut = nn.Conv3d(cin, cout, kernel_size=1, stride=1, padding=0)

    def forward(self, x, zq=None, cache=None):
        cache = cache or {}
        new_cache = {}
        h = self.norm1(x, zq) if self.spatial else self.norm1(x)
        h = F.silu(h)
        h, new_cache["conv1"] = self.conv1(h, cache.get("conv1"))
        h = self.norm2(h, zq) if self.spatial else self.norm2(h)
        h = F.silu(h)
        h, new_cache["conv2"] = self.conv2(h, cache.get("conv2"))
        if self.cin != self.cout:
            x = self.conv_shortcut(x)
        return h + x, new_cache


class Downsample3D(nn.Module):
    def __init__(self, ch, compress_time):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, kernel_size=3, stride=2, padding=0)
        self.compress_time = compress_time

    def forward(self, x):
        if self.compress_time:
            B, C, T, H, W = x.shape
            x = x.permute(0, 3, 4, 1, 2).reshape(B * H * W, C, T)
            if x.shape[-1] % 2 == 1:
                first, rest = x[..., 0], x[..., 1:]
                if rest.shape[-1] > 0:
                    rest = F.avg_pool1d(rest, kernel_size=2, stride=2)
                x = torch.cat([first[..., None], rest], dim=-1)
            else:
                x = F.avg_pool1d(x, kernel_size=2, stride=2)
            x = x.reshape(B, H, W, C, x.shape[-1]).permute(0, 3, 4, 1, 2)
        x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        B, C, T, H, W = x.shape
        x = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        x = self.conv(x)
        return x.reshape(B, T, C, *x.shape[2:]).permute(0, 2, 1, 3, 4)


class Upsample3D(nn.Module):
    def __init__(self, ch, compress_time):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, kernel_size=3, stride=1, padding=1)
        self.compress_time = compress_time

    def forward(self, x):
        if self.compress_time:
            if x.shape[2] > 1 and x.shape[2] % 2 == 1:
                first, rest = x[:, :, 0], x[:, :, 1:]
                first = F.interpolate(first, scale_factor=2.0)
                rest = F.interpolate(rest, scale_factor=2.0)
                x = torch.cat([first[:, :, None], rest], dim=2)
            elif x.shape[2] > 1:
                x = F.interpolate(x, scale_factor=2.0)
            else:
                x = F.interpolate(x.squeeze(2), scale_factor=2.0)[:, :, None]
        else:
            B, C, T, H, W = x.shape
            x = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
            x = F.interpolate(x, scale_factor=2.0)
            x = x.reshape(B, T, C, *x.shape[2:]).permute(0, 2, 1, 3, 4)
        B, C, T, H, W = x.shape
        x = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
        x = self.conv(x)
        return x.reshape(B, T, *x.shape[1:]).permute(0, 2, 1, 3, 4)


class _Stage(nn.Module):
    """Down / mid / up block: `resnets` (+ `downsamplers` | `upsamplers`)."""

    def __init__(self, cin, cout, n, groups, eps, spatial_norm_dim=None, down=None, up=None):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock3D(cin if i == 0 else cout, cout, groups, eps, spatial_norm_dim)
                                      for i in range(n)])
        if down is not None:
            self.downsamplers = nn.ModuleList([Downsample3D(cout, down)])
        if up is not None:
            self.upsamplers = nn.ModuleList([Upsample3D(cout, up)])

    def forward(self, x, zq=None, cache=None):
        cache = cache or {}
        new_cache = {}
        for i, r in enumerate(self.resnets):
            x, new_cache[f"resnet_{i}"] = r(x, zq, cache.get(f"resnet_{i}"))
        if hasattr(self, "downsamplers"):
            x = self.downsamplers[0](x)
        if hasattr(self, "upsamplers"):
            x = self.upsamplers[0](x)
        return x, new_cache


class Encoder3D(nn.Module):
    def __init__(self, c: VAEConfig):
        super().__init__()
        ch = c.block_out_channels
        tlevel = int(np.log2(c.temporal_compression_ratio))
        self.conv_in = CausalConv3d(c.in_channels, ch[0], 3)
        blocks = []
        out = ch[0]
        for i in range(len(ch)):
            cin, out = out, ch[i]
            final = i == len(ch) - 1
            blocks.append(_Stage(cin, out, c.layers_per_block, c.norm_num_groups, c.norm_eps,
                                 down=None if final else (i < tlevel)))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _Stage(ch[-1], ch[-1], 2, c.norm_num_groups, c.norm_eps)
        self.norm_out = nn.GroupNorm(c.norm_num_groups, ch[-1], eps=1e-6)
        self.conv_out = CausalConv3d(ch[-1], 2 * c.latent_channels, 3)

    def forward(self, x, cache=None):
        cache = cache or {}
        nc = {}
        h, nc["conv_in"] = self.conv_in(x, cache.get("conv_in"))
        for i, b in enumerate(self.down_blocks):
            h, nc[f"down_{i}"] = b(h, None, cache.get(f"down_{i}"))
        h, nc["mid"] = self.mid_block(h, None, cache.get("mid"))
        h = F.silu(self.norm_out(h))
        h, nc["conv_out"] = self.conv_out(h, cache.get("conv_out"))
        return h, nc


class Decoder3D(nn.Module):
    def __init__(self, c: VAEConfig):
        super().__init__()
        ch = list(reversed(c.block_out_channels))
        tlevel = int(np.log2(c.temporal_compression_ratio))
        zc = c.latent_channels
        self.conv_in = CausalConv3d(zc, ch[0], 3)
        self.mid_block = _Stage(ch[0], ch[0], 2, c.norm_num_groups, c.norm_eps, spatial_norm_dim=zc)
        blocks = []
        out = ch[0]
        for i in range(len(ch)):
            cin, out = out, ch[i]
            final = i == len(ch) - 1
            blocks.append(_Stage(cin, out, c.layers_per_block + 1, c.norm_num_groups, c.norm_eps, spatial_norm_dim=zc,
                                 up=None if final else (i < tlevel)))
        self.up_blocks = nn.ModuleList(blocks)
        self.norm_out = SpatialNorm3D(ch[-1], zc, c.norm_num_groups)
        self.conv_out = CausalConv3d(ch[-1], c.out_channels, 3)

    def forward(self, z, cache=None):
        cache = cache or {}
        nc = {}
        h, nc["conv_in"] = self.conv_in(z, cache.get("conv_in"))
        h, nc["mid"] = self.mid_block(h, z, cache.get("mid"))
        for i, b in enumerate(self.up_blocks):
            h, nc[f"up_{i}"] = b(h, z, cache.get(f"up_{i}"))
        h = F.silu(self.norm_out(h, z))
        h, nc["conv_out"] = self.conv_out(h, cache.get("conv_out"))
        return h, nc


class DiagonalGaussian:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        gdev = generator.device if generator is not None else self.parameters.device
        n = torch.randn(self.mean.shape, generator=generator, device=gdev, dtype=self.parameters.dtype)
        return self.mean + self.std * n.to(self.parameters.device)

    def mode(self):
        return self.mean


class OracleVAE(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.config = cfg
        self.encoder = Encoder3D(cfg)
        self.decoder = Decoder3D(cfg)
        self.use_slicing = False
        self.use_tiling = False
        self.num_latent_frames_batch_size = 2
        self.num_sample_frames_batch_size = 8
        n = len(cfg.block_out_channels) - 1
        self.tile_sample_min_height = cfg.sample_height // 2
        self.tile_sample_min_width = cfg.sample_width // 2
        self.tile_latent_min_height = int(self.tile_sample_min_height / (2 ** n))
        self.tile_latent_min_width = int(self.tile_sample_min_width / (2 ** n))
        self.tile_overlap_factor_height = 1 / 6
        self.tile_overlap_factor_width = 1 / 5

    def enable_slicing(self):
        self.use_slicing = True

    def enable_tiling(self):
        self.use_tiling = True

    # -------------------------------------------------------------- frame batching
    @staticmethod
    def _batches(num_frames, fbs):
        nb = max(num_frames // fbs, 1)
        rem = num_frames % fbs
        return [(fbs * i + (0 if i == 0 else rem), fbs * (i + 1) + rem) for i in range(nb)]

    def _run_batched(self, net, x, fbs):
        cache = None
        outs = []
        for s, e in self._batches(x.shape[2], fbs):
            y, cache = net(x[:, :, s:e], cache)
            outs.append(y)
        return torch.cat(outs, dim=2)

    # -------------------------------------------------------------- blends
    @staticmethod
    def blend_v(a, b, extent):
        extent = min(a.shape[3], b.shape[3], extent)
        for y in range(extent):
            b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
        return b

    @staticmethod
    def blend_h(a, b, extent):
        extent = min(a.shape[4], b.shape[4], extent)
        for x in range(extent):
            b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
        return b

    def _tiled(self, net, x, fbs, tile_h, tile_w, ov_h, ov_w, blend_h, blend_w, lim_h, lim_w):
        H, W = x.shape[3], x.shape[4]
        rows = []
        for i in range(0, H, ov_h):
            row = []
            for j in range(0, W, ov_w):
                row.append(self._run_batched(net, x[:, :, :, i:i + tile_h, j:j + tile_w], fbs))
            rows.append(row)
        result_rows = []
        for i, row in enumerate(rows):
            result_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self.blend_v(rows[i - 1][j], tile, blend_h)
                if j > 0:
                    tile = self.blend_h(row[j - 1], tile, blend_w)
                result_row.append(tile[:, :, :, :lim_h, :lim_w])
            result_rows.append(torch.cat(result_row, dim=4))
        return torch.cat(result_rows, dim=3)

    # -------------------------------------------------------------- public surface
    def _encode(self, x):
        H, W = x.shape[3], x.shape[4]
        if self.use_tiling and (W > self.tile_sample_min_width or H > self.tile_sample_min_height):
            ov_h = int(self.tile_sample_min_height * (1 - self.tile_overlap_factor_height))
            ov_w = int(self.tile_sample_min_width * (1 - self.tile_overlap_factor_width))
            bl_h = int(self.tile_latent_min_height * self.tile_overlap_factor_height)
            bl_w = int(self.tile_latent_min_width * self.tile_overlap_factor_width)
            return self._tiled(self.encoder, x, self.num_sample_frames_batch_size, self.tile_sample_min_height,
                               self.tile_sample_min_width, ov_h, ov_w, bl_h, bl_w, self.tile_latent_min_height - bl_h,
                               self.tile_latent_min_width - bl_w)
        return self._run_batched(self.encoder, x, self.num_sample_frames_batch_size)

    def encode(self, x):
        if self.use_slicing and x.shape[0] > 1:
            h = torch.cat([self._encode(s) for s in x.split(1)])
        else:
            h = self._encode(x)
        return SimpleNamespace(latent_dist=DiagonalGaussian(h))

    def _decode(self, z):
        H, W = z.shape[3], z.shape[4]
        if self.use_tiling and (W > self.tile_latent_min_width or H > self.tile_latent_min_height):
            ov_h = int(self.tile_latent_min_height * (1 - self.tile_overlap_factor_height))
            ov_w = int(self.tile_latent_min_width * (1 - self.tile_overlap_factor_width))
            bl_h = int(self.tile_sample_min_height * self.tile_overlap_factor_height)
            bl_w = int(self.tile_sample_min_width * self.tile_overlap_factor_width)
            return self._tiled(self.decoder, z, self.num_latent_frames_batch_size, self.tile_latent_min_height,
                               self.tile_latent_min_width, ov_h, ov_w, bl_h, bl_w, self.tile_sample_min_height - bl_h,
                               self.tile_sample_min_width - bl_w)
        return self._run_batched(self.decoder, z, self.num_latent_frames_batch_size)

    def decode(self, z):
        if self.use_slicing and z.shape[0] > 1:
            dec = torch.cat([self._decode(s) for s in z.split(1)])
        else:
            dec = self._decode(z)
        return SimpleNamespace(sample=dec)


def seeded_vae_init_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Variance-preserving seeded weights (He-style for the conv stacks) so activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if p.ndim == 1:
                if "norm" in name and name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (1.3 / fan_in ** 0.5))
    return model
