"""Oracle (test infrastructure): numpy/torch-CPU restatement of the sliding-window blend.

Restates /root/reference/aether/utils/postprocess_utils.py:847-864 (`compute_scale`) and
/root/reference/evaluation/video_depth/launch_aether.py:166-287 (spatial blend, temporal blend chain).
PINNED: tests/golden/make_golden.py runs the reference's own `compute_scale` and
`process_with_sliding_window` in this container; tests/test_oracle_blend.py compares this file with the
committed outputs (compute_scale.npz, sliding_*.npz).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def compute_scale(prediction, target, mask) -> float:
    """postprocess_utils.py:847-864."""
    if isinstance(prediction, np.ndarray):
        prediction = torch.from_numpy(prediction).float()
    if isinstance(target, np.ndarray):
        target = torch.from_numpy(target).float()
    if isinstance(mask, np.ndarray):
        mask = torch.from_numpy(mask).bool()
    numerator = torch.sum(mask * prediction * target, (1, 2))
    denominator = torch.sum(mask * prediction * prediction, (1, 2))
    scale = torch.zeros_like(numerator)
    valid = (denominator != 0).nonzero()
    scale[valid] = numerator[valid] / denominator[valid]
    return scale.item()


def _take(a, axis, lo, hi):
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(lo, hi)
    return a[tuple(sl)]


def blend_chain(windows: Sequence[np.ndarray], ranges: Sequence[Tuple[int, int]], axis: int):
    """One chain along `axis` (2: width :219-234, 1: height :235-250, 0: time :262-285)."""
    final = None
    for idx, (win, rng) in enumerate(zip(windows, ranges)):
        if idx == 0:
            final = win
            continue
        prev_end = ranges[idx - 1][1]
        start = rng[0]
        overlap = prev_end - start
        a = _take(win, axis, 0, overlap)
        b = _take(final, axis, final.shape[axis] - overlap, final.shape[axis])
        if axis == 2:
            dims = (1, -1, overlap)
        elif axis == 1:
            dims = (1, overlap, -1)
        else:
            dims = (1, -1, windows[0].shape[-1])
        scale = compute_scale(a.reshape(*dims), b.reshape(*dims), np.ones_like(b).reshape(*dims))
        aligned = scale * win
        shape = list(windows[0].shape)
        shape[axis] = rng[1]
        result = np.ones(shape)
        wshape = [1, 1, 1]
        wshape[axis] = overlap
        weight = np.linspace(1, 0, overlap).reshape(wshape)
        sl = lambda lo, hi: tuple(slice(lo, hi) if d == axis else slice(None) for d in range(3))
        result[sl(0, start)] = final[sl(0, start)]
        result[sl(prev_end, None)] = aligned[sl(prev_end - start, None)]
        result[sl(start, prev_end)] = final[sl(start, prev_end)] * weight + aligned[sl(0, overlap)] * (1 - weight)
        final = result
    return final


def blend_all(disparities: List[np.ndarray], tiles: np.ndarray, n_spatial: int, is_horizontal: bool):
    """tiles: int array [n, 6] = (t_start, t_end, h_start, h_end, w_start, w_end) in reference order."""
    n_temporal = len(tiles) // n_spatial
    temporal, t_ranges = [], []
    for ti in range(n_temporal):
        tl = tiles[ti * n_spatial:(ti + 1) * n_spatial]
        wins = disparities[ti * n_spatial:(ti + 1) * n_spatial]
        rng = [(int(t[4]), int(t[5])) if is_horizontal else (int(t[2]), int(t[3])) for t in tl]
        temporal.append(blend_chain(wins, rng, 2 if is_horizontal else 1))
        t_ranges.append((int(tl[0][0]), int(tl[0][1])))
    return blend_chain(temporal, t_ranges, 0)
