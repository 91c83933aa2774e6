"""Input side on the device against the reference's own host code:
  * aether_resize_bilinear_u8 vs the golden of the reference's prepare_input (evaluation/video_depth/launch_aether.py:388-403,
    cv2.resize INTER_LINEAR then / 255.0 -- tests/golden/prepare_input.npz, produced by the reference's function): bit-exact;
  * aether_u8_frames_to_model_input vs the pipeline's host preprocessing (:451-512) for every pixel value and on strided crops;
  * the whole pipeline fed with a uint8 DeviceClip crop vs the float64 host crop the reference's launcher hands over: identical
    latents."""
import numpy as np
import pytest
import torch

from helpers import PREPARE_INPUT_SIZES, TINY, empty_prompt_embeds, prepare_input_frames, subsample

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_device_resize_reproduces_prepare_input_golden(golden_dir):
    from aether_b200.sliding_window import prepare_frames_device
    g = np.load(golden_dir / "prepare_input.npz")
    for h, w in PREPARE_INPUT_SIZES:
        frames = prepare_input_frames(h, w)
        got_u8 = prepare_frames_device(frames, torch.device(DEV))
        got = got_u8.cpu().numpy() / 255.0                       # the reference's own last step (float64)
        assert got.dtype == np.float64 and list(got.shape) == g[f"{h}x{w}__shape"].tolist()
        assert got.sum() == float(g[f"{h}x{w}__sum"])
        assert np.array_equal(subsample(got, (1, 24, 24, 1)), g[f"{h}x{w}__sub"])


def test_u8_model_input_equals_host_preprocessing_for_every_value():
    from aether_b200 import ops
    k = torch.arange(256, dtype=torch.uint8)
    frames = k.reshape(1, 16, 16, 1).repeat(2, 1, 1, 3).contiguous()
    frames[1] = frames[1].flip(1)
    got = ops.u8_frames_to_model_input(frames.to(DEV)).cpu()
    for ref_in in (frames.numpy().astype(np.float32) / 255.0, frames.numpy() / 255.0):     # :454 (float32) and launcher (float64)
        ref = (2.0 * torch.from_numpy(ref_in).permute(0, 3, 1, 2) - 1.0).to(torch.bfloat16)
        assert torch.equal(got, ref)


def test_u8_model_input_on_a_strided_crop():
    from aether_b200 import ops
    from aether_b200.sliding_window import DeviceClip
    g = torch.Generator().manual_seed(0)
    clip = torch.randint(0, 256, (9, 40, 61, 3), generator=g, dtype=torch.uint8).to(DEV)
    crop = DeviceClip(clip)[0, 2:7, 8:40, 5:45, :]
    assert not crop.is_contiguous()
    got = ops.u8_frames_to_model_input(crop)
    ref = (2.0 * (crop.cpu().double() / 255.0).permute(0, 3, 1, 2) - 1.0).to(torch.bfloat16)
    assert got.shape == (5, 3, 32, 40) and torch.equal(got.cpu(), ref)


def test_pipeline_on_device_clip_equals_host_float64_frames():
    """One tile through the CUDA pipeline: uint8 crop resident on the GPU vs the float64 host crop of the reference's
    launcher (identical VAE input => identical posterior, noise and latents)."""
    from test_pipeline_gpu import _cuda_pipeline
    from aether_b200.sliding_window import DeviceClip
    H, W, F = TINY["height"], TINY["width"], TINY["num_frames"]
    g = np.random.default_rng(5)
    frames = g.integers(0, 256, size=(F + 4, H, W + 24, 3), dtype=np.uint8)
    clip = DeviceClip(torch.from_numpy(frames).to(DEV))
    host = frames / 255.0
    pipe = _cuda_pipeline()
    kw = dict(task="reconstruction", height=H, width=W, num_frames=F, num_inference_steps=2, output_latents=True)
    a = pipe(video=clip[0, 2:2 + F, 0:H, 24:24 + W, :], generator=torch.Generator().manual_seed(1), **kw)
    b = pipe(video=host[2:2 + F, 0:H, 24:24 + W, :], generator=torch.Generator().manual_seed(1), **kw)
    assert torch.equal(a, b)
