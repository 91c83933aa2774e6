"""Output side of the video-depth evaluation: the on-disk format the evaluator consumes and the
depth-metric alignment that reads it back.

Reference:
  evaluation/video_depth/launch_aether.py:349-365   depth = clip(1 / disparity, 0, 100) -> `frame_%04d.npy` per frame
                                                     (+ colourised mp4s through imageio / matplotlib, which this image
                                                     does not have: written only when both import)
  evaluation/video_depth/tools.py:179-470           depth_evaluation: validity mask, alignment of the prediction to the
                                                     ground truth (median / scale (Weiszfeld) / scale&shift least squares /
                                                     scale&shift LAD by Adam / metric), error metrics and delta thresholds
This is host-side evaluation glue (masked reductions over a few megapixels, once per sequence), restated with torch
tensor ops on whatever device the inputs live on; the numbers are pinned by tests/golden/depth_eval.npz, produced by the
reference's own function (tests/golden/make_golden.py::make_depth_eval)."""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ writers
def save_depth_frames(path: str, depth_maps: Sequence[np.ndarray]) -> int:
    """launch_aether.py:364-365: one `frame_%04d.npy` per frame (np.save of the float64 [H, W] map)."""
    os.makedirs(path, exist_ok=True)
    for i, depth_map in enumerate(depth_maps):
        np.save(f"{path}/frame_{(i):04d}.npy", depth_map)
    return len(depth_maps)


def write_sequence_outputs(path: str, rgb: np.ndarray, disparity: np.ndarray, fps: int = 24) -> Dict[str, object]:
    """launch_aether.py:347-365 for one sequence.  The depth frames are always written; the two preview videos need
    imageio (+ matplotlib's "Spectral" colour map for the disparity one) and are skipped, and reported as such, where
    those packages are missing."""
    depth_maps = np.clip(1.0 / disparity, 0, 1e2)
    out = {"depth_frames": save_depth_frames(path, depth_maps), "videos": []}
    try:
        import imageio.v3 as iio
        import matplotlib
    except Exception as e:  # pragma: no cover - depends on the image
        out["videos_skipped"] = f"{type(e).__name__}: {e}"
        return out
    pos = disparity[disparity > 0]
    lo, hi = pos.min(), pos.max()
    col = matplotlib.colormaps["Spectral"](((hi - disparity) / (hi - lo)).clip(0, 1), bytes=False)[..., 0:3]
    iio.imwrite(os.path.join(path, "pred_disparity.mp4"), (col * 255).astype(np.uint8), fps=fps)
    iio.imwrite(os.path.join(path, "pred_rgb.mp4"), (np.clip(rgb, 0, 1) * 255).astype(np.uint8), fps=fps)
    out["videos"] = ["pred_disparity.mp4", "pred_rgb.mp4"]
    return out


# ------------------------------------------------------------------------------------------------ metrics
def depth_to_disparity(depth: torch.Tensor) -> torch.Tensor:
    """tools.py:31-41: 1 / depth where depth > 0, else 0."""
    out = torch.zeros_like(depth)
    pos = depth > 0
    out[pos] = 1.0 / depth[pos]
    return out


def depth_edge(depth: torch.Tensor, rtol: float = 0.03, kernel_size: int = 3) -> torch.Tensor:
    """tools.py:123-176 (mask=None, atol=None): pixels whose 3x3 neighbourhood spans more than rtol of their depth."""
    import torch.nn.functional as F
    shape = depth.shape
    d = depth.reshape(-1, 1, *shape[-2:])
    pad = kernel_size // 2
    span = F.max_pool2d(d, kernel_size, stride=1, padding=pad) + F.max_pool2d(-d, kernel_size, stride=1, padding=pad)
    return ((span / d).nan_to_num_() > rtol).reshape(*shape)


def _lad_adam(pred: torch.Tensor, gt: torch.Tensor, s_init: float, lr: float, max_iters: int, tol: float = 1e-6):
    """tools.py:69-120: minimise sum |s * pred + t - gt| over (s, t) with Adam, stop when the loss moves < tol."""
    s = torch.tensor([s_init], requires_grad=True, device=pred.device, dtype=pred.dtype)
    t = torch.tensor([0.0], requires_grad=True, device=pred.device, dtype=pred.dtype)
    opt = torch.optim.Adam([s, t], lr=lr)
    prev = None
    with torch.enable_grad():
        for _ in range(max_iters):
            opt.zero_grad()
            loss = torch.sum(torch.abs(s * pred + t - gt))
            loss.backward()
            opt.step()
            if prev is not None and torch.abs(prev - loss) < tol:
                break
            prev = loss.item()
    return s.detach().item(), t.detach().item()


ALIGNMENTS = ("median", "scale", "scale&shiftl2", "scale&shift", "metric")


def depth_evaluation(predicted_depth_original, ground_truth_depth_original, max_depth=80, custom_mask=None,
                     post_clip_min=None, post_clip_max=None, pre_clip_min=None, pre_clip_max=None,
                     align_with_lstsq=False, align_with_lad=False, align_with_lad2=False, metric_scale=False, lr=1e-4,
                     max_iters=1000, use_gpu=False, align_with_scale=False, disp_input=False, mask_edge=False):
    """Same signature, return value and arithmetic as tools.py:179-470 (`align_with_lad`, the scipy variant the
    evaluation scripts never select -- eval_depth.py:150-215 -- raises NotImplementedError)."""
    as_t = lambda a: torch.from_numpy(a) if isinstance(a, np.ndarray) else a
    pred0, gt0, cmask = as_t(predicted_depth_original), as_t(ground_truth_depth_original), custom_mask
    if cmask is not None:
        cmask = as_t(cmask)
    if align_with_lad:
        raise NotImplementedError("align_with_lad (scipy.optimize.minimize) is not used by the evaluation scripts")
    if pred0.dim() == 3:                                   # frames are stacked along the rows
        w = pred0.shape[-1]
        pred0, gt0 = pred0.reshape(-1, w), gt0.reshape(-1, w)
        if cmask is not None:
            cmask = cmask.reshape(-1, w)
    if use_gpu:
        pred0, gt0 = pred0.cuda(), gt0.cuda()
    mask = (gt0 > 0) & (gt0 < max_depth) if max_depth is not None else gt0 > 0
    if mask_edge:
        mask = mask & (~depth_edge(gt0))
    pred, gt = pred0[mask], gt0[mask]
    if pre_clip_min is not None:
        pred = torch.clamp(pred, min=pre_clip_min)
    if pre_clip_max is not None:
        pred = torch.clamp(pred, max=pre_clip_max)
    real_gt = None
    if disp_input:                                         # align in disparity space
        real_gt = gt.clone()
        gt = 1 / (gt + 1e-8)

    s = t = None
    if metric_scale:
        kind = "metric"
    elif align_with_lstsq:
        kind = "affine"
        A = np.hstack([pred.cpu().numpy().reshape(-1, 1), np.ones((pred.numel(), 1), dtype=pred.cpu().numpy().dtype)])
        sol = np.linalg.lstsq(A, gt.cpu().numpy().reshape(-1, 1), rcond=None)[0]
        s = torch.tensor(sol[0], device=pred0.device)
        t = torch.tensor(sol[1], device=pred0.device)
        pred = s * pred + t
    elif align_with_lad2:
        kind = "affine"
        s, t = _lad_adam(pred, gt, (torch.median(gt) / torch.median(pred)).item(), lr, max_iters)
        pred = s * pred + t
    elif align_with_scale:
        kind = "scale"
        s = torch.nanmean(gt) / torch.nanmean(pred)
        for _ in range(10):                                # Weiszfeld iterations of the L1 scale fit
            wgt = 1.0 / ((s * pred - gt).abs() + 1e-8)
            s = torch.sum(wgt * pred * gt) / torch.sum(wgt * pred ** 2)
        s = s.clamp(min=1e-3).detach()
        pred = s * pred
    else:
        kind = "scale"
        s = torch.median(gt) / torch.median(pred)
        pred = pred * s

    if disp_input:
        gt = real_gt
        pred = depth_to_disparity(pred)
    if post_clip_min is not None:
        pred = torch.clamp(pred, min=post_clip_min)
    if post_clip_max is not None:
        pred = torch.clamp(pred, max=post_clip_max)
    inner = None
    if cmask is not None:
        assert cmask.shape == gt0.shape
        inner = cmask.cpu()[mask.cpu()].to(pred.device)
        pred, gt = pred[inner], gt[inner]

    abs_rel = torch.mean(torch.abs(pred - gt) / gt).item()
    sq_rel = torch.mean(((pred - gt) ** 2) / gt).item()
    rmse = torch.sqrt(torch.mean((pred - gt) ** 2)).item()
    pred = torch.clamp(pred, min=1e-5)
    log_rmse = torch.sqrt(torch.mean((torch.log(pred) - torch.log(gt)) ** 2)).item()
    ratio = torch.maximum(pred / gt, gt / pred)
    deltas = [torch.mean((ratio < thr).float()).item() for thr in (1.0, 1.25, 1.25 ** 2, 1.25 ** 3)]

    full = pred0 if kind == "metric" else (pred0 * s + t if kind == "affine" else pred0 * s)
    if disp_input:
        full = depth_to_disparity(full)
    err = torch.where(mask, torch.abs(full - gt0) / gt0, torch.zeros_like(gt0))
    gt_full = torch.where(mask, gt0, torch.zeros_like(gt0))
    n_valid = torch.sum(mask).item() if inner is None else torch.sum(inner).item()
    vals = [abs_rel, sq_rel, rmse, log_rmse] + deltas
    if n_valid == 0:
        vals = [0] * 8
    keys = ["Abs Rel", "Sq Rel", "RMSE", "Log RMSE", "δ < 1.", "δ < 1.25", "δ < 1.25^2", "δ < 1.25^3"]
    results = dict(zip(keys, vals))
    results["valid_pixels"] = n_valid
    return results, err, full, gt_full


def evaluate_depth_sequence(pred_dir: str, gt_depths: Sequence[np.ndarray], align: str = "scale&shift",
                            max_depth: float = 70, use_gpu: bool = False, mask_edge: bool = False) -> Dict[str, float]:
    """Read the `frame_%04d.npy` files written by `save_depth_frames` and evaluate them against ground-truth maps with
    the alignment selected like eval_depth.py:157-215 (`--align`)."""
    files = sorted(f for f in os.listdir(pred_dir) if f.startswith("frame_") and f.endswith(".npy"))
    pred = np.stack([np.load(os.path.join(pred_dir, f)) for f in files])
    gt = np.stack(list(gt_depths)).astype(pred.dtype)
    kw = {"scale&shift": dict(align_with_lad2=True), "scale": dict(align_with_scale=True), "metric": dict(metric_scale=True),
          "scale&shiftl2": dict(align_with_lstsq=True), "median": {}}[align]
    return depth_evaluation(pred, gt, max_depth=max_depth, use_gpu=use_gpu, post_clip_max=max_depth, mask_edge=mask_edge,
                            **kw)[0]
