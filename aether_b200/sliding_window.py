"""Tile-parallel sliding-window evaluation path (the only multi-GPU axis of the hot path).

Mirrors /root/reference/evaluation/video_depth/launch_aether.py:81-287 (`process_with_sliding_window`):
41-frame temporal windows with stride 8, 480x720 spatial tiles (only one axis may tile), one pipeline call
per tile with a freshly seeded generator (:151-158), spatial then temporal blend with a masked-LSQ scale
(`compute_scale`, aether/utils/postprocess_utils.py:847-864) and a linear cross-fade.

What changes, B200-first:
  * the reference runs the tiles serially in one process (it can only shard whole sequences, :320-323);
    here the flattened tile list (reference order k = t_idx * n_spatial + i) is dealt round-robin over the
    ranks of a torch.distributed group and processed in ROUNDS: in round j rank r runs tile j * N + r on its
    own GPU (weights replicated), the per-tile disparity stays on the device (pipeline output_type="pt");
  * exchange: after each round the N - 1 fp32 disparity tiles of the round travel once, peer to peer, to the
    blend rank (NCCL send/recv over NVLink / NVSwitch, 56.7 MB per tile) -- nothing is replicated to everyone;
  * the blend chain itself (sequential and order-sensitive, :166-285) STREAMS behind the tile compute on the
    blend rank: a window is spatially blended and appended to the temporal chain as soon as its tiles have
    arrived, in the reference's order, into one fp64 buffer in place (the reference re-allocates an
    np.ones(float64) result per step; element-wise the arithmetic is identical) with the K10 kernels; the
    masked-LSQ scale of every window stays in device memory (no host synchronisation in the chain).
Return value on the blend rank = the reference's: (final_rgb fp32 [F,480,720,3] = first tile of the first
window (:173-176, :265-266), final_disparity fp64 [T,H,W]); other ranks get (None, None) unless
`result_on_all_ranks` asks for a broadcast.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, device_guard


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


def tile_owner(k: int, world_size: int) -> int:
    """Round-robin deal: tile k is run by rank k % N in round k // N (consecutive tiles -- the spatial tiles of one
    temporal window -- land on different GPUs, and the windows complete in chain order)."""
    return k % world_size


def partition_tiles(n_tiles: int, rank: int, world_size: int) -> List[int]:
    """Tiles run by `rank`, in round order."""
    return list(range(rank, n_tiles, world_size))


# ------------------------------------------------------------------------------------------ K10 wrappers
def _v3(t: torch.Tensor):
    assert t.dim() == 3 and t.stride(2) == 1 and t.is_cuda and t.dtype in (torch.float32, torch.float64)
    return t.data_ptr(), int(t.dtype == torch.float64), t.stride(0), t.stride(1)


def _new_work(device) -> torch.Tensor:
    """Device scratch of aether_scale_reduce: {sum(p*t), sum(p*p)} as fp64 + block counter + per-block partials."""
    n = _lib.require_device().aether_scale_reduce_work_bytes()
    return torch.zeros((n + 7) // 8, dtype=torch.float64, device=device)


@device_guard
def scale_sums_(work: torch.Tensor, prediction: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Enqueue compute_scale's two sums (postprocess_utils.py:847-864, mask of ones) over CUDA views [n0,n1,n2];
    the result stays in `work[:2]` on the device (no synchronisation)."""
    lib = _lib.require_device()
    pp, pf, p0, p1 = _v3(prediction)
    tp, tf, t0, t1 = _v3(target)
    n0, n1, n2 = prediction.shape
    assert tuple(target.shape) == (n0, n1, n2)
    check(lib.aether_scale_reduce(pp, pf, p0, p1, tp, tf, t0, t1, n0, n1, n2, work.data_ptr(), current_stream()),
          "scale_reduce")
    return work


def compute_scale(prediction: torch.Tensor, target: torch.Tensor) -> float:
    """compute_scale(prediction, target, ones) as a python float (synchronises; the blend chain itself never calls
    this -- it consumes the device-resident sums)."""
    work = scale_sums_(_new_work(prediction.device), prediction, target)
    num, den = work[:2].tolist()
    num32, den32 = np.float32(num), np.float32(den)     # the reference holds both sums in fp32
    return float(num32 / den32) if den32 != 0 else 0.0


@device_guard
def _crossfade_(dst: torch.Tensor, acc: torch.Tensor, win: torch.Tensor, axis: int, scale: float = 1.0,
                sums: Optional[torch.Tensor] = None):
    lib = _lib.require_device()
    assert dst.dtype == torch.float64 and dst.shape == acc.shape == win.shape
    ap, af, a0, a1 = _v3(acc)
    wp, wf, w0, w1 = _v3(win)
    n0, n1, n2 = dst.shape
    check(lib.aether_blend_crossfade(dst.data_ptr(), dst.stride(0), dst.stride(1), ap, af, a0, a1, wp, wf, w0, w1,
                                     float(scale), 0 if sums is None else sums.data_ptr(), n0, n1, n2, axis,
                                     current_stream()), "blend_crossfade")


@device_guard
def _scale_copy_(dst: torch.Tensor, src: torch.Tensor, apply_scale: bool, scale: float = 1.0,
                 sums: Optional[torch.Tensor] = None):
    lib = _lib.require_device()
    assert dst.dtype == torch.float64 and dst.shape == src.shape
    if dst.numel() == 0:
        return
    sp, sf, s0, s1 = _v3(src)
    n0, n1, n2 = dst.shape
    check(lib.aether_scale_copy(dst.data_ptr(), dst.stride(0), dst.stride(1), sp, sf, s0, s1, float(scale),
                                0 if sums is None else sums.data_ptr(), int(apply_scale), n0, n1, n2,
                                current_stream()), "scale_copy")


@device_guard
def disparity_to_depth_device(disparity: torch.Tensor) -> torch.Tensor:
    """launch_aether.py:347 on the device: fp64 depth = clip(1 / disparity, 0, 100) of a CUDA [T,H,W] tensor."""
    lib = _lib.require_device()
    out = torch.empty(disparity.shape, dtype=torch.float64, device=disparity.device)
    sp, sf, s0, s1 = _v3(disparity)
    n0, n1, n2 = disparity.shape
    check(lib.aether_disparity_to_depth(out.data_ptr(), out.stride(0), out.stride(1), sp, sf, s0, s1, n0, n1, n2,
                                        current_stream()), "disparity_to_depth")
    return out


def _axis_slices(axis: int):
    return lambda a, b: tuple(slice(a, b) if d == axis else slice(None) for d in range(3))


@device_guard
def _append_(buf: torch.Tensor, win: torch.Tensor, start: int, prev_end: int, axis: int, work: torch.Tensor):
    """One link of the reference's chain (:193-250 spatial / :268-284 temporal), in place on `buf`: `win` covers
    [start, start + n) on `axis`, the accumulation so far ends at `prev_end`.  One C call (aether_blend_link = scale
    sums, cross-fade of the overlap, scaled copy of the new part; the scale stays on the device)."""
    lib = _lib.require_device()
    assert buf.dtype == torch.float64 and buf.dim() == 3 and buf.stride(2) == 1
    wp, wf, w0, w1 = _v3(win)
    e0, e1, e2 = win.shape
    check(lib.aether_blend_link(buf.data_ptr(), buf.stride(0), buf.stride(1), wp, wf, w0, w1, e0, e1, e2, axis, start,
                                prev_end, work.data_ptr(), current_stream()), "blend_link")


def blend_chain(windows: Sequence[torch.Tensor], ranges: Sequence[Tuple[int, int]], axis: int,
                out: Optional[torch.Tensor] = None, work: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's sequential blend along `axis` (2 = width / 1 = height: launch_aether.py:166-252;
    0 = time: :259-285).  windows[i] covers [ranges[i][0], ranges[i][1]) on that axis.  Returns windows[0]
    itself when there is a single window (dtype preserved like the reference), else an fp64 tensor (`out` when
    given).  Nothing here synchronises with the host."""
    if len(windows) == 1:
        return windows[0]
    w0 = windows[0]
    full = list(w0.shape)
    full[axis] = max(r[1] for r in ranges)
    buf = out if out is not None else torch.empty(full, dtype=torch.float64, device=w0.device)
    assert list(buf.shape) == full and buf.dtype == torch.float64
    work = work if work is not None else _new_work(w0.device)
    sl = _axis_slices(axis)
    _scale_copy_(buf[sl(ranges[0][0], ranges[0][1])], w0, False)
    for idx in range(1, len(windows)):
        _append_(buf, windows[idx], ranges[idx][0], ranges[idx - 1][1], axis, work)
    return buf


class StreamingBlend:
    """The blend rank's state: windows are pushed in temporal order as their tiles arrive; each push enqueues the
    spatial blend of the window (reference :166-252) and one link of the temporal chain (:259-285) on the current
    stream.  `final` is the reference's final_disparity (fp64 [T, H, W]) once every window has been pushed."""

    def __init__(self, plan: Plan, t: int, h: int, w: int, device):
        self.plan = plan
        self.device = torch.device(device)
        self.final = torch.empty((t, h, w), dtype=torch.float64, device=self.device) if plan.n_temporal > 1 else None
        self._win = (torch.empty((plan.frames_per_window, h, w), dtype=torch.float64, device=self.device)
                     if plan.n_spatial > 1 else None)
        self._work = _new_work(self.device)
        self._pending = {}
        self._next = 0
        self._prev_end = 0
        self._single = None

    @property
    def windows_done(self) -> int:
        return self._next

    def push_tile(self, k: int, disparity: torch.Tensor):
        """disparity: fp32 CUDA [F, 480, 720] of tile k (any arrival order); completes windows in order."""
        assert disparity.dtype == torch.float32 and disparity.device == self.device
        self._pending[k] = disparity
        ns = self.plan.n_spatial
        while self._next < self.plan.n_temporal and all(self._next * ns + i in self._pending for i in range(ns)):
            tiles = self.plan.tiles[self._next * ns:(self._next + 1) * ns]
            self._push_window(tiles, [self._pending.pop(tl.k) for tl in tiles])
            self._next += 1

    def _push_window(self, tiles, disps):
        p = self.plan
        if p.n_spatial > 1:
            rng = [(tl.w_start, tl.w_end) if p.is_horizontal else (tl.h_start, tl.h_end) for tl in tiles]
            win = blend_chain(disps, rng, 2 if p.is_horizontal else 1, out=self._win, work=self._work)
        else:
            win = disps[0]
        t0, t1 = tiles[0].t_start, tiles[0].t_end
        if p.n_temporal == 1:
            self._single = win.clone() if win is self._win else win     # dtype preserved like the reference (:262-264)
        elif self._next == 0:
            _scale_copy_(self.final[t0:t1], win, False)
        else:
            _append_(self.final, win, t0, self._prev_end, 0, self._work)
        self._prev_end = t1

    def result(self) -> torch.Tensor:
        assert self._next == self.plan.n_temporal, "not every window has been pushed"
        return self._single if self.plan.n_temporal == 1 else self.final


def blend_all(disparities: Sequence[torch.Tensor], plan: Plan) -> torch.Tensor:
    """Spatial blend inside each temporal window, then the temporal chain (reference order), from a complete list."""
    d0 = disparities[0]
    t = max(tl.t_end for tl in plan.tiles)
    h = max(tl.h_end for tl in plan.tiles)
    w = max(tl.w_end for tl in plan.tiles)
    sb = StreamingBlend(plan, t, h, w, d0.device)
    for tl in plan.tiles:
        sb.push_tile(tl.k, disparities[tl.k])
    return sb.result()


# ------------------------------------------------------------------------------------------ exchange step
def exchange_round(round_tiles: Sequence[int], mine: Optional[Tuple[int, torch.Tensor]], shape, rank: int,
                   world_size: int, root: int = 0, group=None, device=None) -> List[Tuple[int, torch.Tensor]]:
    """Move the round's disparity tiles to the blend rank: every owner other than `root` sends its tile once, peer to
    peer; `root` returns [(k, tile)] for all tiles of the round (its own included), the others return [].  A rank
    without a tile in this round (last, partial round; or world_size > n_tiles) takes no part."""
    if world_size == 1:
        return [mine] if mine is not None else []
    import torch.distributed as dist
    ops, got = [], []
    if rank == root:
        for k in round_tiles:
            if tile_owner(k, world_size) == root:
                assert mine is not None and mine[0] == k
                got.append(mine)
            else:
                buf = torch.empty(shape, dtype=torch.float32, device=device)
                ops.append(dist.P2POp(dist.irecv, buf, tile_owner(k, world_size), group))
                got.append((k, buf))
    elif mine is not None:
        ops.append(dist.P2POp(dist.isend, mine[1].contiguous(), root, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return sorted(got, key=lambda kv: kv[0]) if rank == root else []


# ------------------------------------------------------------------------------------------ entry point
class TileParallelRun:
    """One sliding-window evaluation, driven round by round (`run_round(j)` for j in range(n_rounds), then
    `finish()`).  `process_with_sliding_window` is the plain loop over it; bench.py times individual rounds."""

    def __init__(self, pipeline, obs_image, num_inference_step: int, total_frames: int, seed: int, rank: int = 0,
                 world_size: int = 1, group=None, device: Optional[torch.device] = None,
                 tile_fn: Optional[Callable] = None, root: int = 0, skip_unused_rgb: bool = True,
                 stats: Optional[dict] = None):
        b, t, h, w, c = obs_image.shape
        assert b == 1, "Only batch size 1 is supported"
        self.pipeline, self.obs, self.steps, self.seed = pipeline, obs_image, num_inference_step, seed
        self.rank, self.world, self.group, self.root = rank, world_size, group, root
        self.tile_fn, self.skip_unused_rgb, self.stats = tile_fn, skip_unused_rgb, stats
        self.plan = plan_windows(t, h, w, total_frames)
        self.thw = (t, h, w)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_tiles = len(self.plan.tiles)
        tl0 = self.plan.tiles[0]
        self.tile_shape = (self.plan.frames_per_window, tl0.h_end - tl0.h_start, tl0.w_end - tl0.w_start)
        self.n_rounds = (self.n_tiles + world_size - 1) // world_size
        with torch.cuda.device(self.device):
            self.blend = StreamingBlend(self.plan, t, h, w, self.device) if rank == root else None
        self.rgb0 = None
        self._marks = []                                  # (kind, start_event, end_event)
        self._fetched = 0

    def _timed(self, kind, fn):
        if self.stats is None:
            return fn()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b_.record()
        self._marks.append((kind, a, b_))
        return r

    def _run_tile(self, k):
        tl = self.plan.tiles[k]
        crop = self.obs[0, tl.t_start:tl.t_end, tl.h_start:tl.h_end, tl.w_start:tl.w_end, :]
        if self.tile_fn is not None:
            return self.tile_fn(tl, crop)
        kw = {"decode_rgb": False} if (self.skip_unused_rgb and k != 0) else {}
        out = self.pipeline(video=crop, num_inference_steps=self.steps, num_frames=tl.t_end - tl.t_start,
                            generator=torch.Generator(device=self.device).manual_seed(self.seed), return_dict=False,
                            fps=12, output_type="pt", **kw)
        return (None if out[0] is None else out[0][0]), out[1][0]

    def run_round(self, j: int):
        """Round j: this rank's tile j * N + rank through the pipeline, the round's tiles to the blend rank, and (there)
        every window that is now complete onto the chain.  Everything is enqueued on the current stream."""
        with torch.cuda.device(self.device):
            round_tiles = list(range(j * self.world, min((j + 1) * self.world, self.n_tiles)))
            k = j * self.world + self.rank
            mine = None
            if k < self.n_tiles:
                rgb, disp = self._timed("tile_ms", lambda: self._run_tile(k))
                if k == 0:
                    self.rgb0 = rgb
                d = torch.as_tensor(disp).to(device=self.device, dtype=torch.float32)
                assert tuple(d.shape) == self.tile_shape, (tuple(d.shape), self.tile_shape)
                mine = (k, d)
            got = self._timed("collective_ms", lambda: exchange_round(round_tiles, mine, self.tile_shape, self.rank,
                                                                      self.world, self.root, self.group, self.device))
            if self.rank == self.root:
                self._timed("blend_ms", lambda: [self.blend.push_tile(kk, dd) for kk, dd in got])
            if j == 0 and self.world > 1 and self.root != tile_owner(0, self.world):
                self._move_rgb0()

    def _move_rgb0(self):
        """The rgb the reference returns is tile 0's (:173-176, :265-266); when the blend rank is not its owner the
        frames follow the disparity tile there, once."""
        import torch.distributed as dist
        owner = tile_owner(0, self.world)
        if self.rank == owner:
            r = torch.as_tensor(np.asarray(self.rgb0) if not isinstance(self.rgb0, torch.Tensor) else self.rgb0)
            dist.send(r.to(device=self.device, dtype=torch.float32).contiguous(), self.root, group=self.group)
            self.rgb0 = None
        elif self.rank == self.root:
            r = torch.empty(self.tile_shape + (3,), dtype=torch.float32, device=self.device)
            dist.recv(r, owner, group=self.group)
            self.rgb0 = r

    def collect_stats(self):
        """Fold the device timings recorded so far into `stats` (synchronises the device)."""
        if self.stats is None:
            return
        torch.cuda.synchronize(self.device)
        for kind, a, b_ in self._marks:
            self.stats[kind] = self.stats.get(kind, 0.0) + a.elapsed_time(b_)
        self._marks = []

    def finish_stats_reset(self):
        """Drop the timings recorded so far (e.g. after warm-up rounds)."""
        if self.stats is not None:
            torch.cuda.synchronize(self.device)
            self._marks = []
            self.stats.clear()

    def fetch_finalized(self, out: Optional[torch.Tensor] = None, discard: bool = False):
        """Blend rank, multi-window clips: the frames of the blended disparity that no later window can touch any more
        (everything before the start of the next window to be pushed), copied to the host since the previous call --
        lets a caller stream results out while later tiles are still being computed.  Returns (first_frame, numpy
        fp64 [n, H, W]) or None when nothing new is final.  `out`: optional pinned host tensor to stage through (used when
        it is large enough); `discard`: only advance the "already fetched" mark, copy nothing."""
        if self.blend is None or self.blend.final is None:
            return None
        m = self.blend.windows_done
        upto = self.thw[0] if m >= self.plan.n_temporal else self.plan.tiles[m * self.plan.n_spatial].t_start
        if m == 0 or upto <= self._fetched:
            return None
        lo, self._fetched = self._fetched, upto
        if discard:
            return None
        src = self.blend.final[lo:upto]
        if out is not None and out.shape[0] >= upto - lo:
            dst = out[: upto - lo]
            dst.copy_(src, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return lo, dst.numpy()
        return lo, src.cpu().numpy()

    def finish(self, result_on_all_ranks: bool = False, return_device: bool = False):
        t, h, w = self.thw
        plan, device = self.plan, self.device
        with torch.cuda.device(device):
            final = self.blend.result() if self.rank == self.root else None
            rgb0 = self.rgb0
            if self.world > 1 and result_on_all_ranks:
                import torch.distributed as dist
                dt = torch.float64 if (plan.n_temporal > 1 or plan.n_spatial > 1) else torch.float32
                if self.rank != self.root:
                    final = torch.empty((t, h, w), dtype=dt, device=device)
                dist.broadcast(final, src=self.root, group=self.group)
                r = (torch.as_tensor(np.asarray(rgb0) if not isinstance(rgb0, torch.Tensor) else rgb0,
                                     dtype=torch.float32, device=device) if self.rank == self.root
                     else torch.empty(self.tile_shape + (3,), dtype=torch.float32, device=device))
                dist.broadcast(r, src=self.root, group=self.group)
                rgb0 = r
            out_rgb = out_disp = None
            if final is not None:
                def fetch():
                    rr = rgb0.cpu().numpy() if isinstance(rgb0, torch.Tensor) else np.asarray(rgb0)
                    return rr, (final if return_device else final.cpu().numpy())
                out_rgb, out_disp = self._timed("d2h_ms", fetch)
            if self.stats is not None:
                self.collect_stats()
                self.stats["tiles_run"] = len(partition_tiles(self.n_tiles, self.rank, self.world))
                self.stats["rounds"] = self.n_rounds
        return out_rgb, out_disp


def process_with_sliding_window(pipeline, obs_image: np.ndarray, num_inference_step: int, total_frames: int,
                                seed: int, rank: int = 0, world_size: int = 1, group=None,
                                device: Optional[torch.device] = None,
                                tile_fn: Optional[Callable] = None, root: int = 0, result_on_all_ranks: bool = False,
                                skip_unused_rgb: bool = True, stats: Optional[dict] = None,
                                return_device: bool = False):
    """Same positional signature as the reference (launch_aether.py:81-83); rank/world_size/group select the
    tile-parallel mode.  `tile_fn(tile, crop) -> (rgb, disparity)` overrides the pipeline call (tests).
    `skip_unused_rgb`: the reference decodes the rgb latents of every tile and keeps only those of tile 0
    (:173-176, :265-266); with this flag the other tiles skip that VAE decode (identical outputs).
    `stats` (optional dict) receives device-timed `tile_ms`, `collective_ms` (peer-to-peer exchange incl. the wait for
    the slowest rank of each round), `blend_ms`, `d2h_ms` of this rank.
    `return_device`: return the blended disparity as a CUDA fp64 tensor instead of numpy.
    Returns (final_rgb, final_disparity) on the blend rank `root`; (None, None) elsewhere unless
    `result_on_all_ranks`."""
    run = TileParallelRun(pipeline, obs_image, num_inference_step, total_frames, seed, rank, world_size, group, device,
                          tile_fn, root, skip_unused_rgb, stats)
    for j in range(run.n_rounds):
        run.run_round(j)
    return run.finish(result_on_all_ranks, return_device)


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


def window_size(h: int, w: int) -> Tuple[int, int]:
    """(new_h, new_w) of prepare_input (launch_aether.py:392-399): short side of the 480 x 720 window met exactly."""
    aspect_ratio = w / h
    return ((480, int(round(480 * aspect_ratio))) if aspect_ratio > 720 / 480 else (int(round(720 / aspect_ratio)), 720))


def prepare_frames_device(images, device: Optional[torch.device] = None) -> torch.Tensor:
    """`prepare_input` on the device, up to (not including) its `/ 255.0`: H x W x 3 uint8 frames (a list of equally
    sized numpy arrays, or one uint8 array / tensor [T, H, W, 3]) are uploaded as they are (3 bytes per pixel) and
    resized by aether_resize_bilinear_u8, which reproduces cv2.resize's INTER_LINEAR bit for bit.  Returns a uint8
    CUDA tensor [T, new_h, new_w, 3]; `DeviceClip` turns it into the clip object the tile loop consumes."""
    from . import ops
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if isinstance(images, torch.Tensor):
        frames = images
    else:
        frames = torch.from_numpy(np.ascontiguousarray(np.stack(list(images)) if not isinstance(images, np.ndarray)
                                                       else images))
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError("prepare_frames_device expects uint8 frames [T, H, W, 3]")
    frames = frames.to(device).contiguous()
    new_h, new_w = window_size(frames.shape[1], frames.shape[2])
    return ops.resize_bilinear_u8(frames, new_h, new_w)


class DeviceClip:
    """The `[1, T, H, W, 3]` clip of launch_aether.py:401-403 held as uint8 on the GPU.  Indexing it like the reference
    indexes its float64 array (`obs[0, t0:t1, h0:h1, w0:w1, :]`, :140-147) returns a uint8 CUDA view that the pipeline
    converts to its bf16 model input in one kernel; `/ 255.0` is folded into that kernel."""

    def __init__(self, frames_u8: torch.Tensor):
        assert frames_u8.dtype == torch.uint8 and frames_u8.is_cuda and frames_u8.dim() == 4
        self.frames = frames_u8
        self.shape = (1,) + tuple(frames_u8.shape)

    def __getitem__(self, idx):
        b, ts, hs, ws, cs = idx
        assert b == 0 and cs == slice(None)
        return self.frames[ts, hs, ws]


def disparity_to_depth(disparity: np.ndarray) -> np.ndarray:
    """launch_aether.py:347: depth_maps = np.clip(1.0 / disparity_video, 0, 1e2)."""
    return np.clip(1.0 / disparity, 0, 1e2)


def evaluate_sequence(pipeline, frames: Sequence[np.ndarray], num_inference_step: int, seed: int, rank: int = 0,
                      world_size: int = 1, group=None, device: Optional[torch.device] = None,
                      device_input: bool = True, **kw):
    """One sequence of the video-depth evaluation (launch_aether.py:338-347) with the tiles spread over `world_size`
    ranks: frames -> prepare_input -> sliding-window inference + blend -> (rgb of tile 0, disparity, depth) on the
    blend rank ((None, None, None) elsewhere).  depth = clip(1 / disparity, 0, 100) is taken on the device.
    `device_input` (default): the uint8 frames are resized and fed to the tiles on the GPU (prepare_frames_device /
    DeviceClip, same values as the host path); False keeps the reference's float64 host array."""
    obs = DeviceClip(prepare_frames_device(frames, device)) if device_input else prepare_frames(frames)[None]
    rgb, disparity = process_with_sliding_window(pipeline, obs, num_inference_step, len(frames), seed, rank=rank,
                                                 world_size=world_size, group=group, device=device,
                                                 return_device=True, **kw)
    if disparity is None:
        return None, None, None
    depth = disparity_to_depth_device(disparity)
    return rgb, disparity.cpu().numpy(), depth.cpu().numpy()
