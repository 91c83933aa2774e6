"""aether_b200.pose_blend against goldens produced by the reference's own files (tests/golden/make_golden.py::
make_pose_blend): raymap -> poses / point maps, pose smoothing, Sim(3) camera alignment, SLERP interpolation, the rel-pose
window blend (evaluation/rel_pose/launch_aether.py:124-250) and the demo merge (scripts/demo.py:235-422).
Same numpy / scipy / torch primitives on identical inputs: tolerance 1e-9 relative (LAPACK / BLAS summation order)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import fake_pose_window, raymap_from_poses, subsample, synthetic_trajectory

RT = dict(rtol=1e-9, atol=1e-11)


def _stats(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "pose_blend.npz")


def test_raymap_to_poses_and_pointmap(g):
    from aether_b200 import pose_blend as P
    traj = synthetic_trajectory(17, seed=3)
    ray = raymap_from_poses(traj, 12, 20, focal_px=150.0, scale=1.0)
    r = ray.copy()
    pose, fx, fy = P.raymap_to_poses(r)
    np.testing.assert_allclose(pose, g["r2p__pose"], **RT)
    np.testing.assert_allclose(fx, g["r2p__fovx"], **RT)
    np.testing.assert_allclose(fy, g["r2p__fovy"], **RT)
    assert not np.array_equal(r[:, 3:], ray[:, 3:])                 # un-log1p happened in place, like the reference
    # the synthetic raymap was built from `traj`: the recovered poses are that trajectory (sanity of the fixture itself)
    np.testing.assert_allclose(pose[:, :3, 3], traj[:, :3, 3], atol=1e-5)
    _, disp, _ = fake_pose_window(0, 17, h=12, w=20, seed=1)
    for mode in ("none", "simple", "kalman"):
        pcd = P.postprocess_pointmap(disp.copy(), ray.copy(), smooth_camera=(mode != "none"), smooth_method=mode)
        np.testing.assert_allclose(pcd["camera_pose"], g[f"pcd_{mode}__pose"], **RT)
        np.testing.assert_allclose(pcd["intrinsics"], g[f"pcd_{mode}__K"], **RT)
        np.testing.assert_allclose(_stats(pcd["pointmap"]), g[f"pcd_{mode}__points"], rtol=1e-7)
        np.testing.assert_allclose(subsample(pcd["pointmap"], (4, 16, 16, 1)), g[f"pcd_{mode}__points_sub"], rtol=1e-6,
                                   atol=1e-6)


def test_smoothing_alignment_interpolation(g):
    from aether_b200 import pose_blend as P
    noisy = synthetic_trajectory(41, seed=5)
    np.testing.assert_allclose(P.smooth_poses(noisy.copy(), 5, "gaussian"), g["smooth_gaussian"], **RT)
    np.testing.assert_allclose(P.smooth_poses(noisy.copy(), 7, "savgol"), g["smooth_savgol"], **RT)
    # filterpy is absent: the product's own Kalman class vs the oracle's (two restatements of the same equations)
    np.testing.assert_allclose(P.smooth_trajectory(noisy.copy(), 5), g["smooth_kalman"], **RT)
    static = np.repeat(noisy[:1], 12, axis=0) + 1e-4 * np.random.default_rng(0).standard_normal((12, 4, 4)) * (
        np.arange(16).reshape(4, 4) % 4 == 3)
    st = P.detect_static_sequence(static)
    np.testing.assert_allclose([float(st[0]), st[1], st[2]], g["static__flags"], **RT)
    np.testing.assert_allclose(P.adaptive_pose_smoothing(static.copy(), st[1], st[2]), g["static__smoothed"], **RT)
    a, b = synthetic_trajectory(33, seed=7)[:, :3, :4], synthetic_trajectory(33, seed=8)[:, :3, :4]
    R, T, s = P.align_camera_extrinsics(torch.from_numpy(a), torch.from_numpy(b))
    np.testing.assert_allclose(R.numpy(), g["align__R"], **RT)
    np.testing.assert_allclose(T.numpy(), g["align__T"], **RT)
    assert float(s) == pytest.approx(float(g["align__s"]), rel=1e-12)
    np.testing.assert_allclose(P.apply_transformation(torch.from_numpy(a), R, T, s).numpy(), g["align__applied"], **RT)
    got = np.stack([P.interpolate_poses(noisy[3], noisy[30], w) for w in (0.0, 0.25, 0.5, 1.0)])
    np.testing.assert_allclose(got, g["interp"], **RT)
    np.testing.assert_allclose(P.interpolate_poses(noisy[3], noisy[4], 0.3), g["interp_close"], **RT)
    np.testing.assert_allclose(got[0], noisy[30], atol=1e-12)        # weight 0 of pose1 -> pose2 (known answer)


def test_rel_pose_window_blend(g):
    from aether_b200 import pose_blend as P
    frames = np.zeros((1, 105, 24, 40, 3))
    frames[0, :, 0, 0, 0] = np.arange(105)
    assert P.pose_window_starts(105) == ([0, 32, 64], 41) and P.pose_window_starts(41) == ([0], 41)
    assert P.pose_window_starts(30)[1] == 25                          # 41 shrinks in steps of 8 until it fits

    def window_fn(clip):
        return fake_pose_window(int(clip[0, 0, 0, 0]), clip.shape[0], seed=2)
    res = P.process_video_with_sliding_window(None, frames, 4, 42, window_fn=window_fn)
    assert tuple(res["range"]) == tuple(g["relpose__range"])
    for k in ("rgb", "disparity", "focals"):
        np.testing.assert_allclose(_stats(res[k]), g[f"relpose__{k}"], rtol=1e-7)
    np.testing.assert_allclose(res["poses"], g["relpose__poses"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(subsample(res["disparity"], (5, 4, 4)), g["relpose__disparity_sub"], rtol=1e-6)
    assert res["rgb"].shape == (105, 24, 40, 3) and res["poses"].shape == (105, 4, 4)


@pytest.mark.parametrize("align", [False, True])
def test_demo_merged_reconstruction(g, align):
    from aether_b200 import pose_blend as P
    starts = P.get_window_starts(89, 41, 24)
    assert starts == g["demo__starts"].tolist() == [0, 24, 48]
    assert P.get_window_starts(41, 41, 24) == [0] and P.get_window_starts(50, 41, 24) == [0, 9]
    wins = []
    for t0 in starts:
        rgb, d, r = fake_pose_window(t0, 41, seed=4)
        wins.append(SimpleNamespace(rgb=rgb, disparity=d, raymap=r))
    args = P.demo_merge_args(width=40, height=24, smooth_camera=True, smooth_method="kalman", align_pointmaps=align)
    m_rgb, m_disp, m_pose, pts = P.blend_and_merge_window_results(wins, starts, args)
    tag = f"demo_align{int(align)}"
    np.testing.assert_allclose(_stats(m_rgb), g[f"{tag}__rgb"], rtol=1e-9)
    np.testing.assert_allclose(_stats(m_disp), g[f"{tag}__disp"], rtol=1e-7)
    np.testing.assert_allclose(m_pose, g[f"{tag}__poses"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(_stats(pts), g[f"{tag}__pts"], rtol=1e-6)
    np.testing.assert_allclose(subsample(np.asarray(pts), (6, 4, 4, 1)), g[f"{tag}__pts_sub"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(subsample(m_disp, (6, 4, 4)), g[f"{tag}__disp_sub"], rtol=1e-6)
    assert m_rgb.shape == (89, 24, 40, 3) and m_pose.shape == (89, 4, 4) and np.asarray(pts).shape == (89, 24, 40, 3)


@pytest.mark.gpu
def test_project_points_kernel_matches_host_unprojection():
    """aether_project_points (one launch per clip) vs the reference-shaped per-frame numpy `project` (golden-pinned above)."""
    from aether_b200 import pose_blend as P
    gen = np.random.default_rng(3)
    T, H, W = 7, 48, 80
    disp = gen.uniform(0.01, 1.0, (T, H, W))
    disp[0, 0, :5] = [0.0, 1e-12, 5e8, 1.0, 0.5]                     # clip(…, 1e-8, 1e8) edge cases
    poses = synthetic_trajectory(T, seed=9)
    focals = gen.uniform(60, 90, T)
    ref = P.project_clip(disp, focals, poses, W, H)
    for dt in (torch.float64, torch.float32):
        d = torch.from_numpy(disp).to("cuda", dt)
        got = P.project_clip(d, focals, poses, W, H)
        want = ref if dt == torch.float64 else P.project_clip(d.cpu().double().numpy(), focals, poses, W, H)
        assert got.dtype == torch.float64 and tuple(got.shape) == (T, H, W, 3)
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-12, atol=1e-9)
