"""Shared builders for tests and for tests/golden/make_golden.py (test infrastructure; may import oracle/)."""
from __future__ import annotations

import numpy as np
import torch

# tiny geometry used by every pipeline-level parity case: 17 frames of 96x160 -> 5 latent frames of 12x20
TINY = dict(height=96, width=160, num_frames=17, text_len=18, text_dim=128)


def tiny_oracle_modules(dtype=torch.bfloat16, seed=0):
    """(transformer, vae, scheduler) oracle modules for the tiny geometry, seeded."""
    from oracle.dit import OracleDiT, seeded_init_, tiny_config
    from oracle.scheduler import OracleDPMScheduler
    from oracle.vae import OracleVAE, seeded_vae_init_, tiny_vae_config
    dit = seeded_init_(OracleDiT(tiny_config()), seed=seed).eval()
    vae = seeded_vae_init_(OracleVAE(tiny_vae_config()), seed=seed + 1).eval()
    with torch.no_grad():                      # every implementation sees the same bf16-representable weights
        for m in (dit, vae):
            for p in m.parameters():
                p.copy_(p.bfloat16().float())
    vae.enable_slicing()
    vae.enable_tiling()
    return dit.to(dtype), vae.to(dtype), OracleDPMScheduler()


class _ExactDiT:
    """fp64-compute / bf16-I/O adapter around the oracle DiT: the module boundary the reference pipeline sees is
    the production one (bf16 tensors in, bf16 out, :865-875) while the arithmetic in between is float64, so the
    result does not depend on the host's bf16/fp32 GEMM kernels (oneDNN picks them per ISA).  Golden fixtures made
    with these modules are portable across x86 hosts up to rare single-ulp bf16 rounding flips."""

    def __init__(self, dit):
        self.inner = dit.double()
        self.config = dit.config
        self.dtype = torch.bfloat16

    def __call__(self, hidden_states, encoder_hidden_states, timestep, ofs=None, image_rotary_emb=None,
                 attention_kwargs=None, return_dict=False):
        rot = None if image_rotary_emb is None else tuple(r.double() for r in image_rotary_emb)
        out = self.inner(hidden_states.double(), encoder_hidden_states.double(), timestep, ofs=ofs,
                         image_rotary_emb=rot)[0]
        return (out.to(hidden_states.dtype),)


class _ExactVAE:
    """Same adapter for the VAE: encode/decode compute in float64, posterior parameters / decoded sample are handed
    back in the caller's dtype (so the posterior noise is drawn in bf16 exactly like the production path)."""

    def __init__(self, vae):
        self.inner = vae.double()
        self.config = vae.config
        self.dtype = torch.bfloat16

    def enable_slicing(self):
        self.inner.enable_slicing()

    def enable_tiling(self):
        self.inner.enable_tiling()

    def encode(self, x):
        from types import SimpleNamespace
        from oracle.vae import DiagonalGaussian
        h = self.inner.encode(x.double()).latent_dist.parameters
        return SimpleNamespace(latent_dist=DiagonalGaussian(h.to(x.dtype)))

    def decode(self, z):
        from types import SimpleNamespace
        return SimpleNamespace(sample=self.inner.decode(z.double()).sample.to(z.dtype))


def exact_oracle_modules(seed=0):
    """(transformer, vae, scheduler) for the tiny geometry: same seeded bf16-representable weights as
    `tiny_oracle_modules`, float64 arithmetic, bf16 module I/O.  Used for the PORTABLE pipeline goldens."""
    dit, vae, sched = tiny_oracle_modules(torch.float32, seed)
    return _ExactDiT(dit), _ExactVAE(vae), sched


def empty_prompt_embeds(seed=7):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(1, TINY["text_len"], TINY["text_dim"], generator=g) * 0.2).bfloat16()


def synthetic_video(num_frames=17, height=96, width=160, seed=3) -> np.ndarray:
    """float32 [F, H, W, 3] in [0, 1], smooth in space and time."""
    g = np.random.default_rng(seed)
    t = np.linspace(0, 1, num_frames)[:, None, None, None]
    y = np.linspace(0, 1, height)[None, :, None, None]
    x = np.linspace(0, 1, width)[None, None, :, None]
    ph = g.uniform(0, 6.28, size=(1, 1, 1, 3))
    v = 0.5 + 0.25 * np.sin(6.28 * (x * 2 + t) + ph) + 0.2 * np.cos(6.28 * (y * 3 - t * 0.5) + ph * 2)
    v = v + 0.03 * g.standard_normal(v.shape)
    return np.clip(v, 0, 1).astype(np.float32)


def synthetic_raymap(num_frames=17, h=12, w=20, seed=5) -> np.ndarray:
    g = np.random.default_rng(seed)
    base = g.standard_normal((1, 6, h, w)).astype(np.float32)
    drift = np.linspace(0, 1, num_frames, dtype=np.float32)[:, None, None, None]
    return (base + drift * g.standard_normal((1, 6, 1, 1)).astype(np.float32)).astype(np.float32)


# ------------------------------------------------------------------------------------------------------
# deterministic stand-in for one pipeline call of the sliding-window path: (tile ranges, crop) -> outputs.
# Disparity gets a tile-dependent gain so that the masked-LSQ scale alignment has real work to do.
# ------------------------------------------------------------------------------------------------------
def fake_tile_outputs(crop: np.ndarray, t_start: int, h_start: int, w_start: int):
    """crop float64/float32 [F, h, w, 3] -> (rgb float32 [F,h,w,3], disparity float32 [F,h,w])."""
    crop = np.asarray(crop, dtype=np.float64)
    gain = 1.0 + 0.07 * ((t_start // 8) % 5) + 0.11 * (1 if w_start > 0 else 0) + 0.13 * (1 if h_start > 0 else 0)
    disp = (0.2 + crop.mean(axis=-1)) * gain
    f = np.arange(crop.shape[0], dtype=np.float64)[:, None, None]
    disp = disp * (1.0 + 0.01 * np.sin(0.37 * (f + t_start)))
    return crop.astype(np.float32), disp.astype(np.float32)


def synthetic_long_clip(t, h, w, seed=11) -> np.ndarray:
    """float64 [1, t, h, w, 3] like launch_aether.prepare_input (cv2.resize(...)/255.0 -> float64)."""
    g = np.random.default_rng(seed)
    yy = np.linspace(0, 1, h)[None, :, None, None]
    xx = np.linspace(0, 1, w)[None, None, :, None]
    tt = np.linspace(0, 1, t)[:, None, None, None]
    ph = g.uniform(0, 6.28, size=(1, 1, 1, 3))
    v = 0.5 + 0.3 * np.sin(6.28 * (xx * 1.5 + tt * 2) + ph) * np.cos(6.28 * (yy * 2 + tt) + ph)
    return np.clip(v, 0, 1)[None].astype(np.float64)


def subsample(a: np.ndarray, steps) -> np.ndarray:
    sl = tuple(slice(None, None, s) for s in steps)
    return np.ascontiguousarray(a[sl])


PREPARE_INPUT_SIZES = [(720, 1280), (436, 1024), (480, 640), (1080, 1920), (600, 600)]


def prepare_input_frames(h, w, n=2):
    """Seeded uint8 frames for the prepare_input golden (tests/golden/make_golden.py::make_prepare_input)."""
    g = np.random.default_rng(h * 10000 + w)
    return [g.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for _ in range(n)]


def depth_eval_case(seed=0, frames=3, h=48, w=64):
    """Seeded (prediction, ground truth, custom mask) for the depth-metric golden: a smooth scene with a depth step,
    holes (gt = 0), far pixels beyond max_depth and a prediction that is an affine + noisy distortion of it."""
    g = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    gt = np.stack([4.0 + 30.0 * xx + 6.0 * np.sin(5 * yy + 0.7 * f) + 25.0 * (xx > 0.6) for f in range(frames)])
    gt[:, :3, :] = 0.0
    gt[:, -2:, -9:] = 95.0
    pred = 0.37 * gt + 1.3 + 0.4 * g.standard_normal(gt.shape)
    pred = np.abs(pred) + 0.05
    mask = g.random(gt.shape) > 0.2
    return pred.astype(np.float64), gt.astype(np.float64), mask


# ------------------------------------------------------------------------------------------------------
# synthetic camera trajectory / raymap / window outputs for the pose and point-map goldens (8f ranks 1, 2)
# ------------------------------------------------------------------------------------------------------
def _rot_xyz(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rz @ ry @ rx


def synthetic_trajectory(n, seed=0):
    """camera-to-world poses [n, 4, 4]: a smooth arc with a slow pan plus a little jitter."""
    g = np.random.default_rng(seed)
    t = np.linspace(0, 1, n)
    poses = np.zeros((n, 4, 4))
    for i, ti in enumerate(t):
        poses[i, :3, :3] = _rot_xyz(0.05 * np.sin(3 * ti), 0.6 * ti, 0.03 * np.cos(2 * ti))
        poses[i, :3, 3] = [0.8 * np.sin(1.5 * ti), 0.05 * ti, 1.2 * ti]
        poses[i, 3, 3] = 1.0
    poses[:, :3, 3] += 0.004 * g.standard_normal((n, 3))
    return poses


def raymap_from_poses(poses, h, w, focal_px, scale=1.0):
    """float32 raymap [n, 6, h, w] the way the model emits it: channels 0:3 ray directions through the pixel centres of an
    (8h x 8w) image with focal length `focal_px`, 3:6 sign(o) * log1p(|o|) of the (scaled) camera centre."""
    n = poses.shape[0]
    H, W = 8 * h, 8 * w
    ys, xs = np.meshgrid((np.arange(h) + 0.5) * 8, (np.arange(w) + 0.5) * 8, indexing="ij")
    d_cam = np.stack([(xs - W / 2) / focal_px, (ys - H / 2) / focal_px, np.ones_like(xs)], axis=0)      # [3, h, w]
    out = np.zeros((n, 6, h, w), dtype=np.float32)
    for i in range(n):
        out[i, :3] = np.einsum("ij,jhw->ihw", poses[i, :3, :3], d_cam)
        o = poses[i, :3, 3] * scale
        out[i, 3:] = (np.sign(o) * np.log1p(np.abs(o)))[:, None, None]
    return out


def fake_pose_window(t_start, frames, h=3, w=5, seed=0, total=160):
    """One window's pipeline outputs (rgb f32 [F, 8h, 8w, 3], disparity f32 [F, 8h, 8w], raymap f32 [F, 6, h, w]) cut out
    of a global synthetic scene, with a window-dependent similarity (scale + small rotation) and noise on the camera, and a
    window-dependent gain on the disparity -- what the alignment code has to undo."""
    g = np.random.default_rng(seed * 1000 + t_start)
    H, W = 8 * h, 8 * w
    traj = synthetic_trajectory(total, seed)[t_start:t_start + frames].copy()
    s = 1.0 + 0.15 * ((t_start // 8) % 3)
    Rw = _rot_xyz(0.02 * (t_start % 5), -0.03 * (t_start % 3), 0.01)
    traj[:, :3, :3] = traj[:, :3, :3] @ Rw
    traj[:, :3, 3] = s * traj[:, :3, 3] + 0.002 * g.standard_normal((frames, 3))
    ray = raymap_from_poses(traj, h, w, focal_px=0.9 * W, scale=10.0)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    f = (np.arange(frames) + t_start)[:, None, None]
    disp = (0.25 + 0.5 * xx[None] * (0.6 + 0.4 * np.sin(0.11 * f)) + 0.2 * yy[None]) * (0.8 + 0.1 * ((t_start // 8) % 4))
    disp = np.clip(disp + 0.01 * g.standard_normal(disp.shape), 0.02, 1.0)
    rgb = np.clip(np.stack([xx[None] + 0 * f, yy[None] + 0 * f, 0.5 + 0.5 * np.sin(0.2 * f) + 0 * xx[None]], axis=-1), 0, 1)
    return rgb.astype(np.float32), disp.astype(np.float32), ray.astype(np.float32)
