"""`AutoencoderKLCogVideoX`-shaped VAE whose device work is the K9 kernels (conv_tcgen05.cu, vae_kernels.cu).

Drop-in for the `vae` object of the reference pipeline
(/root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py):
    vae.encode(x[B,3,T,H,W]).latent_dist.sample(generator) / .mode()     :557-620 via retrieve_latents :233-245
    vae.decode(z[B,16,t,h,w]).sample                                      :931, :936 (inherited decode_latents)
    vae.config.{latent_channels, scaling_factor, invert_scale_latents, block_out_channels,
                temporal_compression_ratio}                               :571, :843, :925-929
    vae.enable_slicing(); vae.enable_tiling()                             scripts/demo.py:229-230

The control flow (batch slicing, 3x3 spatial tiling with linear blends, frame batching with carried conv
caches, first-frame-preserving temporal resampling) restates diffusers v0.32
models/autoencoders/autoencoder_kl_cogvideox.py; parameter names equal the diffusers state-dict
keys so real CogVideoX-5b-I2V VAE weights load unchanged.  Activations live channels-last [T, H, W, C] in bf16;
every arithmetic step is a C-ABI kernel call (ops below) -- torch is used for allocation, slicing and copies.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check, current_stream, ptr, device_guard

BF16 = torch.bfloat16

_DEFAULTS = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 256, 512), latent_channels=16,
                 layers_per_block=3, act_fn="silu", norm_eps=1e-6, norm_num_groups=32, temporal_compression_ratio=4,
                 sample_height=480, sample_width=720, scaling_factor=0.7, invert_scale_latents=False,
                 use_quant_conv=False, use_post_quant_conv=False)


def _ceil(a, b):
    return (a + b - 1) // b * b


class _H(nn.Module):
    pass


def _conv_holder(cin, cout, k, dims, **kw):
    m = _H()
    m.weight = nn.Parameter(torch.empty((cout, cin) + (k,) * dims, **kw), requires_grad=False)
    m.bias = nn.Parameter(torch.empty(cout, **kw), requires_grad=False)
    return m


def _causal(cin, cout, k, **kw):
    m = _H()
    m.conv = _conv_holder(cin, cout, k, 3, **kw)
    return m


def _gn_holder(c, **kw):
    m = _H()
    m.weight = nn.Parameter(torch.ones(c, **kw), requires_grad=False)
    m.bias = nn.Parameter(torch.zeros(c, **kw), requires_grad=False)
    return m


def _spatial_norm(c, zc, **kw):
    m = _H()
    m.norm_layer = _gn_holder(c, **kw)
    m.conv_y = _causal(zc, c, 1, **kw)
    m.conv_b = _causal(zc, c, 1, **kw)
    return m


def _resnet(cin, cout, zc, **kw):
    m = _H()
    m.norm1 = _spatial_norm(cin, zc, **kw) if zc else _gn_holder(cin, **kw)
    m.norm2 = _spatial_norm(cout, zc, **kw) if zc else _gn_holder(cout, **kw)
    m.conv1 = _causal(cin, cout, 3, **kw)
    m.conv2 = _causal(cout, cout, 3, **kw)
    if cin != cout:
        m.conv_shortcut = _conv_holder(cin, cout, 1, 3, **kw)
    return m


def _stage(cin, cout, n, zc, down=None, up=None, **kw):
    m = _H()
    m.resnets = nn.ModuleList([_resnet(cin if i == 0 else cout, cout, zc, **kw) for i in range(n)])
    if down is not None:
        s = _H(); s.conv = _conv_holder(cout, cout, 3, 2, **kw)
        m.downsamplers = nn.ModuleList([s])
    if up is not None:
        s = _H(); s.conv = _conv_holder(cout, cout, 3, 2, **kw)
        m.upsamplers = nn.ModuleList([s])
    return m


class _Posterior:
    """DiagonalGaussianDistribution over channels-last moments [T, h, w, 2L]."""

    def __init__(self, moments_cl: torch.Tensor, latent_channels: int):
        self.m = moments_cl
        self.L = latent_channels

    @property
    def device(self):
        return self.m.device

    @device_guard
    def _run(self, noise):
        T, h, w, Cp = self.m.shape
        z = torch.empty(1, self.L, T, h, w, dtype=BF16, device=self.m.device)
        check(_lib.require_device().aether_posterior_sample(ptr(self.m), Cp, self.L, ptr(noise), ptr(z), T * h * w,
                                                            current_stream()), "posterior_sample")
        return z

    def sample(self, generator=None):
        T, h, w, _ = self.m.shape
        gdev = generator.device if generator is not None else self.m.device
        noise = torch.randn((1, self.L, T, h, w), generator=generator, device=gdev, dtype=BF16).to(self.m.device)
        return self._run(noise.contiguous())

    def mode(self):
        return self._run(None)


class _CatPosterior:
    def __init__(self, posts: List[_Posterior]):
        self.posts = posts

    def sample(self, generator=None):
        return torch.cat([p.sample(generator) for p in self.posts], dim=0)

    def mode(self):
        return torch.cat([p.mode() for p in self.posts], dim=0)


class AetherVAE(nn.Module):
    def __init__(self, device=None, dtype=BF16, **config):
        super().__init__()
        cfg = dict(_DEFAULTS)
        unknown = set(config) - set(cfg)
        if unknown:
            raise ValueError(f"unknown config keys: {sorted(unknown)}")
        cfg.update(config)
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        self.config = SimpleNamespace(**cfg)
        c = self.config
        if c.use_quant_conv or c.use_post_quant_conv:
            raise NotImplementedError("CogVideoX VAEs have no (post_)quant_conv")
        kw = dict(device=device, dtype=dtype)
        ch = c.block_out_channels
        tl = int(np.log2(c.temporal_compression_ratio))
        zc = c.latent_channels
        enc = _H()
        enc.conv_in = _causal(c.in_channels, ch[0], 3, **kw)
        blocks, out = [], ch[0]
        for i in range(len(ch)):
            cin, out = out, ch[i]
            blocks.append(_stage(cin, out, c.layers_per_block, 0, down=None if i == len(ch) - 1 else (i < tl), **kw))
        enc.down_blocks = nn.ModuleList(blocks)
        enc.mid_block = _stage(ch[-1], ch[-1], 2, 0, **kw)
        enc.norm_out = _gn_holder(ch[-1], **kw)
        enc.conv_out = _causal(ch[-1], 2 * zc, 3, **kw)
        self.encoder = enc
        rch = list(reversed(ch))
        dec = _H()
        dec.conv_in = _causal(zc, rch[0], 3, **kw)
        dec.mid_block = _stage(rch[0], rch[0], 2, zc, **kw)
        blocks, out = [], rch[0]
        for i in range(len(rch)):
            cin, out = out, rch[i]
            blocks.append(_stage(cin, out, c.layers_per_block + 1, zc, up=None if i == len(rch) - 1 else (i < tl), **kw))
        dec.up_blocks = nn.ModuleList(blocks)
        dec.norm_out = _spatial_norm(rch[-1], zc, **kw)
        dec.conv_out = _causal(rch[-1], c.out_channels, 3, **kw)
        self.decoder = dec
        self._down_time = [i < tl for i in range(len(ch) - 1)]
        self._up_time = [i < tl for i in range(len(ch) - 1)]

        self.use_slicing = False
        self.use_tiling = False
        self.num_latent_frames_batch_size = 2
        self.num_sample_frames_batch_size = 8
        n = len(ch) - 1
        self.tile_sample_min_height = c.sample_height // 2
        self.tile_sample_min_width = c.sample_width // 2
        self.tile_latent_min_height = int(self.tile_sample_min_height / (2 ** n))
        self.tile_latent_min_width = int(self.tile_sample_min_width / (2 ** n))
        self.tile_overlap_factor_height = 1 / 6
        self.tile_overlap_factor_width = 1 / 5
        self._packed: Optional[Dict[str, dict]] = None
        self._imaps: Dict[tuple, torch.Tensor] = {}
        self._gn_ws = None
        self._handle = None
        self._handle_tiling = False
        self._ws = None
        # True: orchestrate every kernel from Python through the per-kernel C entry points (the path of round 1, kept so
        # that each kernel stays individually testable and as the cross-check of the native schedule); False: one
        # aether_vae_encode / aether_vae_decode call per batch item.  AETHER_VAE_PER_OP=1 selects it globally.
        import os
        self.per_op = os.environ.get("AETHER_VAE_PER_OP", "0")[:1] == "1"

    # ------------------------------------------------------------------ surface
    @property
    def dtype(self):
        return BF16

    @property
    def device(self):
        return self.decoder.conv_out.conv.weight.device

    def enable_slicing(self):
        self.use_slicing = True

    def enable_tiling(self):
        self.use_tiling = True

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._packed = None
        self._imaps = {}
        self._gn_ws = None
        self._destroy_handle()
        self._ws = None
        return r

    @torch.no_grad()
    def init_synthetic_(self, seed: int = 0):
        dev = self.device
        g = torch.Generator(device=dev).manual_seed(seed)
        for name, p in sorted(self.named_parameters()):
            if p.ndim == 1:
                if "norm" in name and name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * (1.3 / fan_in ** 0.5))
        return self

    # ------------------------------------------------------------------ packing
    @torch.no_grad()
    @device_guard
    def pack(self):
        if self._packed is not None:
            return self
        _lib.require_device()
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("AetherVAE.pack(): move the module to a CUDA device first (no CPU path)")
        P: Dict[str, dict] = {}

        def conv(name, m):
            w = m.weight.detach().to(dev, torch.float32)
            if w.ndim == 4:
                w = w[:, :, None]                                  # Conv2d -> kt = 1
            cout, cin, kt, kh, kw_ = w.shape
            cin_t = _ceil(cin, 8)                                  # channel count of the activation tensor
            cin_p = _ceil(cin_t, 64)
            cout_p = _ceil(cout, 8)
            wp = torch.zeros(cout_p, kt * kh * kw_, cin_p, device=dev, dtype=torch.float32)
            wp[:cout, :, :cin] = w.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh * kw_, cin)
            b = torch.zeros(cout_p, device=dev, dtype=torch.float32)
            b[:cout] = m.bias.detach().to(dev, torch.float32)
            P[name] = dict(w=wp.reshape(cout_p, -1).to(BF16).contiguous(), b=b, k=(kt, kh, kw_), cin=cin_t, cout=cout_p)

        def gn(name, m):
            P[name] = dict(g=m.weight.detach().to(dev, torch.float32).contiguous(),
                           b=m.bias.detach().to(dev, torch.float32).contiguous())

        def spatial(name, m):
            gn(name + ".norm_layer", m.norm_layer)
            # conv_y | conv_b (1x1x1 on the latent) fused row-wise: ONE GEMM yields [N_z, 2C] = (scale | bias)
            wy, wb = m.conv_y.conv, m.conv_b.conv
            P[name + ".conv_yb"] = dict(
                w=torch.cat([wy.weight.detach().reshape(wy.weight.shape[0], -1),
                             wb.weight.detach().reshape(wb.weight.shape[0], -1)], dim=0).to(dev, BF16).contiguous(),
                b=torch.cat([wy.bias.detach(), wb.bias.detach()]).to(dev, torch.float32).contiguous())

        def resnet(name, m, spatial_norm):
            (spatial if spatial_norm else gn)(name + ".norm1", m.norm1)
            (spatial if spatial_norm else gn)(name + ".norm2", m.norm2)
            conv(name + ".conv1", m.conv1.conv)
            conv(name + ".conv2", m.conv2.conv)
            if hasattr(m, "conv_shortcut"):
                conv(name + ".conv_shortcut", m.conv_shortcut)

        def stage(name, m, spatial_norm):
            for i, r in enumerate(m.resnets):
                resnet(f"{name}.resnets.{i}", r, spatial_norm)
            if hasattr(m, "downsamplers"):
                conv(f"{name}.downsamplers.0.conv", m.downsamplers[0].conv)
            if hasattr(m, "upsamplers"):
                conv(f"{name}.upsamplers.0.conv", m.upsamplers[0].conv)

        e, d = self.encoder, self.decoder
        conv("encoder.conv_in", e.conv_in.conv)
        for i, b in enumerate(e.down_blocks):
            stage(f"encoder.down_blocks.{i}", b, False)
        stage("encoder.mid_block", e.mid_block, False)
        gn("encoder.norm_out", e.norm_out)
        conv("encoder.conv_out", e.conv_out.conv)
        conv("decoder.conv_in", d.conv_in.conv)
        stage("decoder.mid_block", d.mid_block, True)
        for i, b in enumerate(d.up_blocks):
            stage(f"decoder.up_blocks.{i}", b, True)
        spatial("decoder.norm_out", d.norm_out)
        conv("decoder.conv_out", d.conv_out.conv)
        # all conv_y | conv_b matrices of the decoder in ONE buffer (the individual entries become views of it): the
        # native schedule evaluates every SpatialNorm's 1x1x1 convs of a frame batch with a single GEMM on the latent
        yb = sorted(k for k in P if k.endswith(".conv_yb"))
        w_all = torch.cat([P[k]["w"] for k in yb], dim=0).contiguous()
        b_all = torch.cat([P[k]["b"] for k in yb]).contiguous()
        off = 0
        for k in yb:
            n = P[k]["w"].shape[0]
            P[k] = dict(w=w_all[off:off + n], b=b_all[off:off + n], col=off)
            off += n
        P["decoder.conv_yb_all"] = dict(w=w_all, b=b_all, col=0, all=True)
        self._packed = P
        self._create_handle()
        return self

    # ------------------------------------------------------------------ native handle (whole schedule in one C call)
    def _tiling_ints(self):
        """The int(...) arithmetic of diffusers tiled_encode / tiled_decode, evaluated once here and handed to the C side
        (stride between tiles, blend extent, kept extent)."""
        oh, ow = self.tile_overlap_factor_height, self.tile_overlap_factor_width
        e_bl_h, e_bl_w = int(self.tile_latent_min_height * oh), int(self.tile_latent_min_width * ow)
        d_bl_h, d_bl_w = int(self.tile_sample_min_height * oh), int(self.tile_sample_min_width * ow)
        return dict(
            enc_overlap_h=int(self.tile_sample_min_height * (1 - oh)), enc_overlap_w=int(self.tile_sample_min_width * (1 - ow)),
            enc_blend_h=e_bl_h, enc_blend_w=e_bl_w, enc_limit_h=self.tile_latent_min_height - e_bl_h,
            enc_limit_w=self.tile_latent_min_width - e_bl_w,
            dec_overlap_h=int(self.tile_latent_min_height * (1 - oh)), dec_overlap_w=int(self.tile_latent_min_width * (1 - ow)),
            dec_blend_h=d_bl_h, dec_blend_w=d_bl_w, dec_limit_h=self.tile_sample_min_height - d_bl_h,
            dec_limit_w=self.tile_sample_min_width - d_bl_w)

    def _create_handle(self):
        import ctypes as C
        lib = _lib.require_device()
        self._destroy_handle()
        c = self.config
        cfg = _lib.VaeConfig()
        for k, v in dict(
                in_channels=c.in_channels, out_channels=c.out_channels, latent_channels=c.latent_channels,
                num_blocks=len(c.block_out_channels), layers_per_block=c.layers_per_block,
                norm_num_groups=c.norm_num_groups, norm_eps=float(c.norm_eps),
                temporal_compression_ratio=c.temporal_compression_ratio, use_tiling=int(self.use_tiling),
                num_latent_frames_batch_size=self.num_latent_frames_batch_size,
                num_sample_frames_batch_size=self.num_sample_frames_batch_size,
                tile_sample_min_height=self.tile_sample_min_height, tile_sample_min_width=self.tile_sample_min_width,
                tile_latent_min_height=self.tile_latent_min_height, tile_latent_min_width=self.tile_latent_min_width,
                **self._tiling_ints()).items():
            setattr(cfg, k, v)
        items = sorted(self._packed.items())
        arr = (_lib.VaeParam * len(items))()
        keep = []
        for i, (name, p) in enumerate(items):
            nb = name.encode()
            keep.append(nb)
            arr[i].name = nb
            if "k" in p:
                arr[i].kind, arr[i].data, arr[i].bias = 0, p["w"].data_ptr(), p["b"].data_ptr()
                arr[i].kt, arr[i].kh, arr[i].kw = p["k"]
                arr[i].cin, arr[i].cout = p["cin"], p["cout"]
            elif "g" in p:
                arr[i].kind, arr[i].data, arr[i].bias = 1, p["g"].data_ptr(), p["b"].data_ptr()
            else:       # kind 2: one SpatialNorm's conv_y | conv_b (cin = its first column in the concatenation); kind 3: all
                arr[i].kind = 3 if p.get("all") else 2
                arr[i].data, arr[i].bias = p["w"].data_ptr(), p["b"].data_ptr()
                arr[i].cin, arr[i].cout = p["col"], p["w"].shape[0]
        h = C.c_void_p()
        check(lib.aether_vae_create(C.byref(cfg), arr, len(items), C.byref(h)), "vae_create")
        self._handle, self._handle_tiling = h, bool(self.use_tiling)
        self._ws = None

    def _destroy_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().aether_vae_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    def _native(self, op: int, x: torch.Tensor):
        """x: bf16 [C, T, H, W] (any T/H strides, unit W stride) -> channels-last moments (op 0) or NCTHW frames (op 1)."""
        import ctypes as C
        lib = _lib.load()
        if self._handle is None or self._handle_tiling != bool(self.use_tiling):
            self._create_handle()                       # enable_tiling() after pack(): the handle carries the flag
        if x.stride(3) != 1:
            x = x.contiguous()
        Cc, T, H, W = x.shape
        dims = (C.c_int32 * 4)()
        check(lib.aether_vae_output_shape(self._handle, op, T, H, W, dims), "vae_output_shape")
        need = lib.aether_vae_workspace_bytes(self._handle, op, T, H, W)
        if need < 0:
            raise RuntimeError("aether_vae_workspace_bytes failed")
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        To, Ho, Wo, Co = (int(v) for v in dims)
        if op == 0:
            out = torch.empty(To, Ho, Wo, Co, dtype=BF16, device=x.device)
            fn, what = lib.aether_vae_encode, "vae_encode"
        else:
            out = torch.empty(self.config.out_channels, To, Ho, Wo, dtype=BF16, device=x.device)
            fn, what = lib.aether_vae_decode, "vae_decode"
        check(fn(self._handle, ptr(x), x.stride(0), x.stride(1), x.stride(2), T, H, W, ptr(out), ptr(self._ws),
                 self._ws.numel(), current_stream()), what)
        return out

    def launches(self, op: int, T: int, H: int, W: int) -> int:
        """Kernel launches + device copies one native encode (op 0, pixels) / decode (op 1, latent size) enqueues."""
        self.pack()
        if self._handle_tiling != bool(self.use_tiling):
            self._create_handle()
        return int(_lib.load().aether_vae_launch_count(self._handle, op, T, H, W))

    # ------------------------------------------------------------------ kernel wrappers (channels-last tensors)
    def _imap(self, vals) -> torch.Tensor:
        key = tuple(int(v) for v in vals)
        t = self._imaps.get(key)
        if t is None:
            t = torch.tensor(key, dtype=torch.int32, device=self.device)
            self._imaps[key] = t
        return t

    def _conv(self, xpad, name, T_out, stride=1, pad=(1, 1), resid=None, out_hw=None):
        p = self._packed[name]
        kt, kh, kw_ = p["k"]
        T_in, H, W, Cin = xpad.shape
        assert Cin == p["cin"] and T_in == T_out + kt - 1, (name, xpad.shape, p["cin"], T_out, kt)
        Ho, Wo = out_hw if out_hw is not None else (H, W)
        y = torch.empty(T_out, Ho, Wo, p["cout"], dtype=BF16, device=xpad.device)
        check(_lib.load().aether_conv3d_bf16(ptr(xpad), T_in, H, W, Cin, ptr(p["w"]), ptr(p["b"]), ptr(resid), ptr(y),
                                             T_out, Ho, Wo, p["cout"], kt, kh, kw_, stride, pad[0], pad[1],
                                             current_stream()), f"conv3d {name}")
        return y

    def _norm(self, x, name, out, silu=True, zq=None):
        """GroupNorm (zq None) or SpatialNorm3D (zq = latent batch [Tz, hz, wz, L] channels-last) + SiLU -> `out`."""
        lib = _lib.load()
        T, H, W, C = x.shape
        G = self.config.norm_num_groups
        N = T * H * W
        if self._gn_ws is None or self._gn_ws.numel() < lib.aether_gn_workspace_floats(C):
            self._gn_ws = torch.empty(lib.aether_gn_workspace_floats(2048), dtype=torch.float32, device=x.device)
        mr = torch.empty(2 * G, dtype=torch.float32, device=x.device)
        # encoder GroupNorms use config.norm_eps; SpatialNorm3D's norm_layer and encoder.norm_out hard-code 1e-6
        eps = 1e-6 if (zq is not None or name == "encoder.norm_out") else self.config.norm_eps
        check(lib.aether_gn_stats(ptr(x), N, C, G, float(eps), ptr(self._gn_ws), ptr(mr), current_stream()), "gn_stats")
        assert out.is_contiguous() and out.shape == x.shape
        if zq is None:
            p = self._packed[name]
            check(lib.aether_gn_apply(ptr(x), ptr(out), N, C, G, ptr(mr), ptr(p["g"]), ptr(p["b"]), 0, 0, 0, 0, H, W, 1, 1,
                                      int(silu), current_stream()), "gn_apply")
            return out
        p = self._packed[name + ".norm_layer"]
        Tz, hz, wz, L = zq.shape
        z2 = zq.reshape(Tz * hz * wz, L)
        pyb = self._packed[name + ".conv_yb"]
        zyb = ops.gemm(z2, pyb["w"], pyb["b"], 0)       # 1x1x1 convs at latent resolution (commute with nearest interp.)
        zy, zb = zyb, zyb[:, C:]                        # column blocks [0, C) = conv_y, [C, 2C) = conv_b; row stride 2C
        if T > 1 and T % 2 == 1:
            tmap = [0] + [1 + ((t - 1) * (Tz - 1)) // (T - 1) for t in range(1, T)]
        else:
            tmap = [(t * Tz) // T for t in range(T)]
        check(lib.aether_gn_apply(ptr(x), ptr(out), N, C, G, ptr(mr), ptr(p["g"]), ptr(p["b"]), ptr(zy), ptr(zb), 2 * C,
                                  ptr(self._imap(tmap)), H, W, hz, wz, int(silu), current_stream()), "gn_apply")
        return out

    @staticmethod
    def _padded(T, H, W, C, dev):
        return torch.empty(T + 2, H, W, C, dtype=BF16, device=dev)

    @staticmethod
    def _fill_time_pad(buf, cache):
        """CogVideoXCausalConv3d: [conv_cache] or [first frame] * 2 in front of the input."""
        if cache is not None:
            buf[:2].copy_(cache)
        else:
            buf[0].copy_(buf[2])
            buf[1].copy_(buf[2])
        return buf[-2:].clone()

    def _causal_from(self, x, name, cache, new_cache, resid=None):
        """causal conv of a plain tensor x (conv_in): copy into a padded buffer first."""
        T, H, W, C = x.shape
        buf = self._padded(T, H, W, C, x.device)
        buf[2:].copy_(x)
        new_cache[name] = self._fill_time_pad(buf, cache.get(name))
        return self._conv(buf, name, T, resid=resid)

    def _norm_conv(self, x, norm_name, conv_name, cache, new_cache, zq=None, resid=None):
        """norm -> SiLU -> causal 3x3x3 conv (the norm writes straight into the time-padded conv input)."""
        T, H, W, C = x.shape
        buf = self._padded(T, H, W, C, x.device)
        self._norm(x, norm_name, buf[2:], True, zq)
        new_cache[conv_name] = self._fill_time_pad(buf, cache.get(conv_name))
        return self._conv(buf, conv_name, T, resid=resid)

    def _resnet(self, x, name, cache, new_cache, zq=None):
        h = self._norm_conv(x, name + ".norm1", name + ".conv1", cache, new_cache, zq)
        r = x
        if (name + ".conv_shortcut") in self._packed:
            r = self._conv(x, name + ".conv_shortcut", x.shape[0], pad=(0, 0))
        return self._norm_conv(h, name + ".norm2", name + ".conv2", cache, new_cache, zq, resid=r)

    def _downsample(self, x, name, compress_time):
        lib = _lib.load()
        T, H, W, C = x.shape
        if compress_time:
            if T % 2 == 1:
                ia = [0] + [2 * k - 1 for k in range(1, (T - 1) // 2 + 1)]
                ib = [-1] + [2 * k for k in range(1, (T - 1) // 2 + 1)]
            else:
                ia = [2 * k for k in range(T // 2)]
                ib = [2 * k + 1 for k in range(T // 2)]
            y = torch.empty(len(ia), H, W, C, dtype=BF16, device=x.device)
            check(lib.aether_avgpool_time(ptr(x), ptr(y), ptr(self._imap(ia)), ptr(self._imap(ib)), len(ia), H * W * C,
                                          current_stream()), "avgpool_time")
            x = y
        T = x.shape[0]
        Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1       # F.pad (0,1,0,1) then 3x3 stride 2
        return self._conv(x, name + ".downsamplers.0.conv", T, stride=2, pad=(0, 0), out_hw=(Ho, Wo))

    def _upsample(self, x, name, compress_time):
        lib = _lib.load()
        T, H, W, C = x.shape
        if compress_time and T > 1 and T % 2 == 1:
            tmap = [0] + [1 + (t - 1) // 2 for t in range(1, 1 + 2 * (T - 1))]
        elif compress_time and T > 1:
            tmap = [t // 2 for t in range(2 * T)]
        else:
            tmap = list(range(T))
        y = torch.empty(len(tmap), 2 * H, 2 * W, C, dtype=BF16, device=x.device)
        check(lib.aether_upsample_nearest(ptr(x), ptr(y), ptr(self._imap(tmap)), len(tmap), 2 * H, 2 * W, H, W, 2, 2, C,
                                          current_stream()), "upsample_nearest")
        return self._conv(y, name + ".upsamplers.0.conv", len(tmap))

    # ------------------------------------------------------------------ encoder / decoder on one frame batch
    def _encoder(self, x, cache):
        nc: Dict[str, torch.Tensor] = {}
        h = self._causal_from(x, "encoder.conv_in", cache, nc)
        for i, blk in enumerate(self.encoder.down_blocks):
            for j in range(len(blk.resnets)):
                h = self._resnet(h, f"encoder.down_blocks.{i}.resnets.{j}", cache, nc)
            if hasattr(blk, "downsamplers"):
                h = self._downsample(h, f"encoder.down_blocks.{i}", self._down_time[i])
        for j in range(2):
            h = self._resnet(h, f"encoder.mid_block.resnets.{j}", cache, nc)
        return self._norm_conv(h, "encoder.norm_out", "encoder.conv_out", cache, nc), nc

    def _decoder(self, z, cache):
        nc: Dict[str, torch.Tensor] = {}
        h = self._causal_from(z, "decoder.conv_in", cache, nc)
        for j in range(2):
            h = self._resnet(h, f"decoder.mid_block.resnets.{j}", cache, nc, zq=z)
        for i, blk in enumerate(self.decoder.up_blocks):
            for j in range(len(blk.resnets)):
                h = self._resnet(h, f"decoder.up_blocks.{i}.resnets.{j}", cache, nc, zq=z)
            if hasattr(blk, "upsamplers"):
                h = self._upsample(h, f"decoder.up_blocks.{i}", self._up_time[i])
        return self._norm_conv(h, "decoder.norm_out", "decoder.conv_out", cache, nc, zq=z), nc

    # ------------------------------------------------------------------ frame batching / tiling (diffusers order)
    @staticmethod
    def _batches(num_frames, fbs):
        nb = max(num_frames // fbs, 1)
        rem = num_frames % fbs
        return [(fbs * i + (0 if i == 0 else rem), fbs * (i + 1) + rem) for i in range(nb)]

    def _to_cl(self, x_ncthw, Cp):
        """[C, T, H, W] (any strides) -> channels-last [T, H, W, Cp] with zero-padded channels."""
        x = x_ncthw.contiguous()
        Cc, T, H, W = x.shape
        out = torch.empty(T, H, W, Cp, dtype=BF16, device=x.device)
        check(_lib.load().aether_ncthw_to_thwc(ptr(x), ptr(out), Cc, Cp, T * H * W, current_stream()), "ncthw_to_thwc")
        return out

    def _run_batched(self, net, x_ncthw, fbs, Cp):
        cache: Dict[str, torch.Tensor] = {}
        outs = []
        for s, e in self._batches(x_ncthw.shape[1], fbs):
            y, cache = net(self._to_cl(x_ncthw[:, s:e], Cp), cache)
            outs.append(y)
        return torch.cat(outs, dim=0) if len(outs) > 1 else outs[0]

    def _blend(self, a, b, axis, extent):
        extent = min(a.shape[axis], b.shape[axis], extent)
        check(_lib.load().aether_tile_blend(ptr(a), ptr(b), b.shape[0], a.shape[1], a.shape[2], b.shape[1], b.shape[2],
                                            b.shape[3], axis, extent, current_stream()), "tile_blend")
        return b

    def _tiled(self, net, x, fbs, Cp, tile_h, tile_w, ov_h, ov_w, bl_h, bl_w, lim_h, lim_w):
        H, W = x.shape[2], x.shape[3]
        rows = []
        for i in range(0, H, ov_h):
            rows.append([self._run_batched(net, x[:, :, i:i + tile_h, j:j + tile_w], fbs, Cp)
                         for j in range(0, W, ov_w)])
        result_rows = []
        for i, row in enumerate(rows):
            rr = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = self._blend(rows[i - 1][j], tile, 1, bl_h)
                if j > 0:
                    tile = self._blend(row[j - 1], tile, 2, bl_w)
                rr.append(tile[:, :lim_h, :lim_w])
            result_rows.append(torch.cat(rr, dim=2))
        return torch.cat(result_rows, dim=1).contiguous()

    def _encode_one(self, x):           # x: [3, T, H, W]
        c = self.config
        Cp = _ceil(c.in_channels, 8)
        H, W = x.shape[2], x.shape[3]
        if self.use_tiling and (W > self.tile_sample_min_width or H > self.tile_sample_min_height):
            ov_h = int(self.tile_sample_min_height * (1 - self.tile_overlap_factor_height))
            ov_w = int(self.tile_sample_min_width * (1 - self.tile_overlap_factor_width))
            bl_h = int(self.tile_latent_min_height * self.tile_overlap_factor_height)
            bl_w = int(self.tile_latent_min_width * self.tile_overlap_factor_width)
            return self._tiled(self._encoder, x, self.num_sample_frames_batch_size, Cp, self.tile_sample_min_height,
                               self.tile_sample_min_width, ov_h, ov_w, bl_h, bl_w, self.tile_latent_min_height - bl_h,
                               self.tile_latent_min_width - bl_w)
        return self._run_batched(self._encoder, x, self.num_sample_frames_batch_size, Cp)

    def _decode_one(self, z):           # z: [L, T, h, w]
        c = self.config
        Cp = _ceil(c.latent_channels, 8)
        H, W = z.shape[2], z.shape[3]
        if self.use_tiling and (W > self.tile_latent_min_width or H > self.tile_latent_min_height):
            ov_h = int(self.tile_latent_min_height * (1 - self.tile_overlap_factor_height))
            ov_w = int(self.tile_latent_min_width * (1 - self.tile_overlap_factor_width))
            bl_h = int(self.tile_sample_min_height * self.tile_overlap_factor_height)
            bl_w = int(self.tile_sample_min_width * self.tile_overlap_factor_width)
            return self._tiled(self._decoder, z, self.num_latent_frames_batch_size, Cp, self.tile_latent_min_height,
                               self.tile_latent_min_width, ov_h, ov_w, bl_h, bl_w, self.tile_sample_min_height - bl_h,
                               self.tile_sample_min_width - bl_w)
        return self._run_batched(self._decoder, z, self.num_latent_frames_batch_size, Cp)

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    @device_guard
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        _lib.require_device()
        self.pack()
        if not x.is_cuda:
            raise RuntimeError("AetherVAE.encode: input must be a CUDA tensor (no CPU path)")
        x = x.to(BF16)
        one = self._encode_one if self.per_op else (lambda xi: self._native(0, xi))
        posts = [_Posterior(one(xi), self.config.latent_channels) for xi in x]
        post = posts[0] if len(posts) == 1 else _CatPosterior(posts)
        return SimpleNamespace(latent_dist=post)

    @torch.no_grad()
    @device_guard
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        _lib.require_device()
        self.pack()
        if not z.is_cuda:
            raise RuntimeError("AetherVAE.decode: input must be a CUDA tensor (no CPU path)")
        z = z.to(BF16)
        if not self.per_op:
            return SimpleNamespace(sample=torch.stack([self._native(1, zi) for zi in z], dim=0))
        outs = []
        for zi in z:
            y = self._decode_one(zi)                                   # [T, H, W, 8]
            T, H, W, Cp = y.shape
            o = torch.empty(self.config.out_channels, T, H, W, dtype=BF16, device=y.device)
            check(_lib.load().aether_thwc_to_ncthw(ptr(y), ptr(o), self.config.out_channels, Cp, T * H * W,
                                                   current_stream()), "thwc_to_ncthw")
            outs.append(o)
        return SimpleNamespace(sample=torch.stack(outs, dim=0))
