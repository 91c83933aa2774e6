"""Tile-parallel sliding-window evaluation path (the only multi-GPU axis of the hot path).

Mirrors /root/reference/evaluation/video_depth/launch_aether.py:81-287 (`process_with_sliding_window`):
41-frame temporal windows with stride 8, 480x720 spatial tiles (only one axis may tile), one pipeline call
per tile with a freshly seeded generator (:151-158), spatial then temporal blend with a masked-LSQ scale
(`compute_scale`, aether/utils/postprocess_utils.py:847-864) and a linear cross-fade.

What changes, B200-first:
  * the reference runs the tiles serially in one process (it can only shard whole sequences, :320-323);
    here the flattened tile list (reference order k = t_idx * n_spatial + i) is partitioned round-robin
    over the ranks of a torch.distributed group, every rank runs its tiles on its own GPU, and ONE
    all-gather (NCCL over NVLink/NVSwitch) of the fp32 disparity tiles precedes the blend;
  * the blend chain itself (sequential and order-sensitive, :166-285) is executed on the device with the K10
    kernels, in the reference's order, into one fp64 buffer in place (the reference re-allocates an
    np.ones(float64) result per step; element-wise the arithmetic is identical).
Return value = the reference's: (final_rgb fp32 [F,480,720,3] = first tile of the first window (:173-176,
:265-266), final_disparity fp64 [T,H,W]).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream


# ------------------------------------------------------------------------------------------ tiling plan
@dataclass(frozen=True)
class Tile:
    k: int                 # flattened index, reference order
    t_idx: int
    s_idx: int
    t_start: int
    t_end: int
    h_start: int
    h_end: int
    w_start: int
    w_end: int


@dataclass(frozen=True)
class Plan:
    tiles: Tuple[Tile, ...]
    n_temporal: int
    n_spatial: int
    is_horizontal: bool
    frames_per_window: int


def plan_windows(t: int, h: int, w: int, total_frames: Optional[int] = None) -> Plan:
    """launch_aether.py:87-149."""
    total_frames = t if total_frames is None else total_frames
    max_frames_per_window = 41
    while max_frames_per_window > total_frames:
        max_frames_per_window -= 8
    temporal_stride = 8
    target_h, target_w = 480, 720
    spatial_overlap_h, spatial_overlap_w = 60, 90
    h_windows = 1 if h <= target_h else math.ceil((h - target_h) / (target_h - spatial_overlap_h)) + 1
    w_windows = 1 if w <= target_w else math.ceil((w - target_w) / (target_w - spatial_overlap_w)) + 1
    assert h_windows == 1 or w_windows == 1, (h_windows, w_windows)
    spatial_stride_h = (h - target_h) // (h_windows - 1) if h_windows > 1 else 0
    spatial_stride_w = (w - target_w) // (w_windows - 1) if w_windows > 1 else 0
    t_starts = list(range(0, t - max_frames_per_window, temporal_stride))
    t_starts.append(t - max_frames_per_window)
    num_windows, stride, is_horizontal = ((w_windows, spatial_stride_w, True) if w_windows > 1
                                          else (h_windows, spatial_stride_h, False))
    tiles = []
    for t_idx, t_start in enumerate(t_starts):
        t_end = min(t_start + max_frames_per_window, t)
        if t_end < t and t_end - t_start < max_frames_per_window:
            t_start = max(0, t - max_frames_per_window)
            t_end = t
        for i in range(num_windows):
            if is_horizontal:
                h_start, h_end = 0, target_h
                w_start = int(i * stride)
                w_end = w_start + target_w
                if w_end > w:
                    w_start, w_end = w - target_w, w
            else:
                w_start, w_end = 0, target_w
                h_start = int(i * stride)
                h_end = h_start + target_h
                if h_end > h:
                    h_start, h_end = h - target_h, h
            tiles.append(Tile(len(tiles), t_idx, i, t_start, t_end, h_start, h_end, w_start, w_end))
    return Plan(tuple(tiles), len(t_starts), num_windows, is_horizontal, max_frames_per_window)


def partition_tiles(n_tiles: int, rank: int, world_size: int) -> List[int]:
    """Round-robin: consecutive tiles (which share a temporal window) land on different GPUs."""
    return list(range(rank, n_tiles, world_size))


# ------------------------------------------------------------------------------------------ exchange step
def gather_tiles(local: Sequence[Tuple[int, torch.Tensor]], n_tiles: int, rank: int, world_size: int,
                 group=None) -> List[torch.Tensor]:
    """All-gather the per-tile disparity tensors (all of identical shape) so that every rank holds the full,
    reference-ordered list.  One collective; with NCCL it runs over NVLink/NVSwitch."""
    if world_size == 1:
        out = [None] * n_tiles
        for k, d in local:
            out[k] = d
        return out
    import torch.distributed as dist
    n_max = (n_tiles + world_size - 1) // world_size
    ref = local[0][1]
    send = torch.zeros((n_max,) + tuple(ref.shape), dtype=ref.dtype, device=ref.device)
    for j, (k, d) in enumerate(local):
        assert k == rank + j * world_size
        send[j].copy_(d)
    recv = torch.empty((world_size * n_max,) + tuple(ref.shape), dtype=ref.dtype, device=ref.device)
    dist.all_gather_into_tensor(recv, send, group=group)      # rank r's block lands at rows [r*n_max, (r+1)*n_max)
    return [recv[(k % world_size) * n_max + k // world_size] for k in range(n_tiles)]


# ------------------------------------------------------------------------------------------ K10 wrappers
def _v3(t: torch.Tensor):
    assert t.dim() == 3 and t.stride(2) == 1 and t.is_cuda and t.dtype in (torch.float32, torch.float64)
    return t.data_ptr(), int(t.dtype == torch.float64), t.stride(0), t.stride(1)


def compute_scale(prediction: torch.Tensor, target: torch.Tensor) -> float:
    """compute_scale(prediction, target, ones) of postprocess_utils.py:847-864 on CUDA views [n0,n1,n2]."""
    lib = _lib.require_device()
    out = torch.zeros(2, dtype=torch.float64, device=prediction.device)
    pp, pf, p0, p1 = _v3(prediction)
    tp, tf, t0, t1 = _v3(target)
    n0, n1, n2 = prediction.shape
    assert tuple(target.shape) == (n0, n1, n2)
    check(lib.aether_scale_reduce(pp, pf, p0, p1, tp, tf, t0, t1, n0, n1, n2, out.data_ptr(), current_stream()),
          "scale_reduce")
    num, den = out.tolist()
    num32, den32 = np.float32(num), np.float32(den)     # the reference holds both sums in fp32
    return float(num32 / den32) if den32 != 0 else 0.0


def _crossfade_(dst: torch.Tensor, acc: torch.Tensor, win: torch.Tensor, scale: float, axis: int):
    lib = _lib.require_device()
    assert dst.dtype == torch.float64 and dst.shape == acc.shape == win.shape
    ap, af, a0, a1 = _v3(acc)
    wp, wf, w0, w1 = _v3(win)
    n0, n1, n2 = dst.shape
    check(lib.aether_blend_crossfade(dst.data_ptr(), dst.stride(0), dst.stride(1), ap, af, a0, a1, wp, wf, w0, w1,
                                     float(scale), n0, n1, n2, axis, current_stream()), "blend_crossfade")


def _scale_copy_(dst: torch.Tensor, src: torch.Tensor, scale: float, apply_scale: bool):
    lib = _lib.require_device()
    assert dst.dtype == torch.float64 and dst.shape == src.shape
    if dst.numel() == 0:
        return
    sp, sf, s0, s1 = _v3(src)
    n0, n1, n2 = dst.shape
    check(lib.aether_scale_copy(dst.data_ptr(), dst.stride(0), dst.stride(1), sp, sf, s0, s1, float(scale),
                                int(apply_scale), n0, n1, n2, current_stream()), "scale_copy")


def blend_chain(windows: Sequence[torch.Tensor], ranges: Sequence[Tuple[int, int]], axis: int) -> torch.Tensor:
    """The reference's sequential blend along `axis` (2 = width / 1 = height: launch_aether.py:166-252;
    0 = time: :259-285).  windows[i] covers [ranges[i][0], ranges[i][1]) on that axis.  Returns windows[0]
    itself when there is a single window (dtype preserved like the reference), else an fp64 tensor."""
    if len(windows) == 1:
        return windows[0]
    w0 = windows[0]
    full = list(w0.shape)
    full[axis] = max(r[1] for r in ranges)
    buf = torch.empty(full, dtype=torch.float64, device=w0.device)
    sl = lambda a, b: tuple(slice(a, b) if d == axis else slice(None) for d in range(3))
    _scale_copy_(buf[sl(ranges[0][0], ranges[0][1])], w0, 1.0, False)
    for idx in range(1, len(windows)):
        win = windows[idx]
        start = ranges[idx][0]
        prev_end = ranges[idx - 1][1]
        overlap = prev_end - start
        n_win = win.shape[axis]
        scale = compute_scale(win[sl(0, overlap)], buf[sl(start, prev_end)])
        _crossfade_(buf[sl(start, prev_end)], buf[sl(start, prev_end)], win[sl(0, overlap)], scale, axis)
        _scale_copy_(buf[sl(prev_end, start + n_win)], win[sl(overlap, n_win)], scale, True)
    return buf


def blend_all(disparities: Sequence[torch.Tensor], plan: Plan) -> torch.Tensor:
    """Spatial blend inside each temporal window, then the temporal chain (reference order)."""
    ns = plan.n_spatial
    temporal, t_ranges = [], []
    for ti in range(plan.n_temporal):
        tl = plan.tiles[ti * ns:(ti + 1) * ns]
        wins = [disparities[t.k] for t in tl]
        rng = [(t.w_start, t.w_end) if plan.is_horizontal else (t.h_start, t.h_end) for t in tl]
        temporal.append(blend_chain(wins, rng, 2 if plan.is_horizontal else 1))
        t_ranges.append((tl[0].t_start, tl[0].t_end))
    return blend_chain(temporal, t_ranges, 0)


# ------------------------------------------------------------------------------------------ entry point
def process_with_sliding_window(pipeline, obs_image: np.ndarray, num_inference_step: int, total_frames: int,
                                seed: int, rank: int = 0, world_size: int = 1, group=None,
                                device: Optional[torch.device] = None,
                                tile_fn: Optional[Callable] = None):
    """Same positional signature as the reference (launch_aether.py:81-83); rank/world_size/group select the
    tile-parallel mode.  `tile_fn(tile, crop) -> (rgb, disparity)` overrides the pipeline call (tests)."""
    b, t, h, w, c = obs_image.shape
    assert b == 1, "Only batch size 1 is supported"
    plan = plan_windows(t, h, w, total_frames)
    device = device or torch.device("cuda", torch.cuda.current_device())
    local = []
    rgb0 = None
    for k in partition_tiles(len(plan.tiles), rank, world_size):
        tl = plan.tiles[k]
        crop = obs_image[0, tl.t_start:tl.t_end, tl.h_start:tl.h_end, tl.w_start:tl.w_end, :]
        if tile_fn is not None:
            rgb, disp = tile_fn(tl, crop)
        else:
            rgb, disp, _ = pipeline(video=crop, num_inference_steps=num_inference_step, num_frames=tl.t_end - tl.t_start,
                                    generator=torch.Generator(device=device).manual_seed(seed), return_dict=False,
                                    fps=12)
            rgb, disp = rgb[0], disp[0]
        if k == 0:
            rgb0 = rgb
        d = torch.as_tensor(disp)
        local.append((k, d.to(device=device, dtype=torch.float32)))
    disparities = gather_tiles(local, len(plan.tiles), rank, world_size, group)
    final = blend_all(disparities, plan)
    if world_size > 1:
        import torch.distributed as dist
        shape = (plan.tiles[0].t_end - plan.tiles[0].t_start, plan.tiles[0].h_end - plan.tiles[0].h_start,
                 plan.tiles[0].w_end - plan.tiles[0].w_start, 3)
        r = torch.as_tensor(rgb0, dtype=torch.float32, device=device) if rank == 0 else torch.empty(
            shape, dtype=torch.float32, device=device)
        dist.broadcast(r, src=0, group=group)
        rgb0 = r.cpu().numpy()
    return np.asarray(rgb0), final.cpu().numpy()


# ------------------------------------------------------------------------------------------ launcher glue (host side)
def prepare_frames(images: Sequence[np.ndarray]) -> np.ndarray:
    """`prepare_input` of launch_aether.py:388-403 without the file read: every H x W x 3 uint8 frame is resized so
    that the short side of the 480 x 720 window is met exactly (aspect kept, `cv2.resize` default = bilinear) and
    scaled to [0, 1] float64.  Returns [T, new_h, new_w, 3]."""
    import cv2  # the reference's resampler; there is no substitute that reproduces its rounding
    out = []
    for img in images:
        h, w = img.shape[:2]
        aspect_ratio = w / h
        new_h, new_w = ((480, int(round(480 * aspect_ratio))) if aspect_ratio > 720 / 480
                        else (int(round(720 / aspect_ratio)), 720))
        out.append(cv2.resize(img, (new_w, new_h)) / 255.0)
    return np.stack(out)


def disparity_to_depth(disparity: np.ndarray) -> np.ndarray:
    """launch_aether.py:347: depth_maps = np.clip(1.0 / disparity_video, 0, 1e2)."""
    return np.clip(1.0 / disparity, 0, 1e2)


def evaluate_sequence(pipeline, frames: Sequence[np.ndarray], num_inference_step: int, seed: int, rank: int = 0,
                      world_size: int = 1, group=None, device: Optional[torch.device] = None):
    """One sequence of the video-depth evaluation (launch_aether.py:338-347) with the tiles spread over `world_size`
    ranks: frames -> prepare_frames -> sliding-window inference + blend -> (rgb of tile 0, disparity, depth)."""
    obs = prepare_frames(frames)[None]
    rgb, disparity = process_with_sliding_window(pipeline, obs, num_inference_step, len(frames), seed, rank=rank,
                                                 world_size=world_size, group=group, device=device)
    return rgb, disparity, disparity_to_depth(disparity)
