"""Parity AT THE BASELINE SHAPE (BASELINE.json configs[1..3]: 41 x 480 x 720 -> 11 x 60 x 90 latents, S = 226 + 14850 =
15076 tokens, 48 heads x 64, D = 3072, text 4096, time 512) against the fp32 oracle run on the GPU box's host cores.

What tiny geometries cannot exercise in combination: the ragged last query/key tile (15076 = 117.78 x 128), the
text/video split at row 226 inside a 256-row GEMM tile, 48 heads, the CTA-pair GEMM at N = 3072 / 9216 / 12288 and
the CFG batch of 2 whose second item starts at row 15076 (not a tile boundary).

  * k transformer layers of `aether_dit_forward` (k = 2 at B = 1; k = 1 at B = 2) + patch embed + tail vs
    oracle.dit.OracleDiT with the same k layers (the oracle times one block in 3-15 s on the box's cores);
  * all 42 layers at B = 2 against two B = 1 forwards of the same items (batch independence: the only check of the
    full-depth CFG batch the CPU cannot give in minutes) -- bit-equal with the single-launch attention, bf16-noise-equal
    with the split tail wave;
  * the real-geometry VAE (block_out_channels 128/256/256/512, 3 layers per block, 240 x 360 px / 30 x 45 latent
    tiles): encode of a 17 x 160 x 432 strip and decode of 4 x 20 x 54 latents -- two horizontally overlapping tiles
    of the real tile width with their blend, two frame batches with conv caches -- vs oracle.vae (sized so that the
    fp32 oracle needs about a minute in total on the box's cores; the first version of this test used a full
    240 x 720 tile row and spent 231 s in the oracle);
  * the three pipeline tasks at 41 x 480 x 720 (reconstruction, prediction + raymap, planning; CFG batch 2 with the
    dynamic guidance schedule, reference :832-899): shapes, dtypes, ranges, determinism and the raymap un-fold.

Tolerances (bf16 storage, fp32 accumulation): DiT rel-RMS <= 1.5e-2 and max-abs <= 0.15 x max|ref| vs the fp32 oracle
(same bound as tests/test_dit_gpu.py); VAE rel-RMS <= 2.5e-2 (tests/test_vae_gpu.py).
"""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
F_LAT, H_LAT, W_LAT, ST = 11, 60, 90, 226


def _rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref)
    return (err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-20)).item(), err.abs().max().item()


# ---------------------------------------------------------------------------------------------------- DiT, k layers
def _dit_pair(num_layers, seed=0):
    from oracle.dit import DiTConfig, OracleDiT, seeded_init_
    from aether_b200.transformer import AetherTransformer3D
    cfg = DiTConfig(num_layers=num_layers)                   # AetherV1 geometry, only the depth is reduced
    oracle = seeded_init_(OracleDiT(cfg), seed=seed).eval()
    with torch.no_grad():
        for p in oracle.parameters():
            p.copy_(p.bfloat16().float())
    model = AetherTransformer3D(**cfg.to_dict())
    model.load_state_dict(oracle.state_dict(), strict=True)
    return cfg, oracle, model.to(DEV).pack()


def _dit_inputs(cfg, B, seed=1):
    from oracle.rope import prepare_rotary_positional_embeddings
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, F_LAT, cfg.in_channels, H_LAT, W_LAT, generator=g).bfloat16()
    e = (torch.randn(B, ST, cfg.text_embed_dim, generator=g) * 0.2).bfloat16()
    cos, sin = prepare_rotary_positional_embeddings(480, 720, F_LAT, sample_height=cfg.sample_height,
                                                    sample_width=cfg.sample_width)
    assert cos.shape == (F_LAT * 30 * 45, 64)
    return x, e, cos, sin


@pytest.mark.parametrize("B,k,ts", [(1, 2, [999]), (2, 1, [499, 499])])
def test_dit_k_layers_at_full_sequence_length_match_fp32_oracle(B, k, ts):
    cfg, oracle, model = _dit_pair(k)
    x, e, cos, sin = _dit_inputs(cfg, B)
    t = torch.tensor(ts, dtype=torch.int64)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = oracle(x.float(), e.float(), t, image_rotary_emb=(cos, sin))[0]
    cpu_s = time.perf_counter() - t0
    out = model(hidden_states=x.to(DEV), encoder_hidden_states=e.to(DEV), timestep=t.to(DEV), ofs=None,
                image_rotary_emb=(cos.to(DEV), sin.to(DEV)), attention_kwargs=None, return_dict=False)[0]
    assert out.shape == ref.shape == (B, F_LAT, cfg.out_channels, H_LAT, W_LAT) and out.dtype == torch.bfloat16
    assert torch.isfinite(out.float()).all()
    rel, mx = _rel(out, ref)
    print(f"full-size DiT B={B} layers={k}: rel-rms {rel:.3e}, max-abs {mx:.3e} (max|ref| {ref.abs().max():.3f}); "
          f"oracle {cpu_s:.1f} s on {torch.get_num_threads()} threads")
    assert rel <= 1.5e-2 and mx <= 0.15 * ref.abs().max().item(), (rel, mx)
    # per-batch-item and per-region error must be uniform: a mistake at the text/video boundary or in the ragged tile
    # shows up as a localised excess (last latent frame = last query rows; first rows of item 1 = row 15076 of the GEMM)
    for b in range(B):
        rb, _ = _rel(out[b], ref[b])
        r_last, _ = _rel(out[b, -1], ref[b, -1])
        r_first, _ = _rel(out[b, 0], ref[b, 0])
        assert max(rb, r_last, r_first) <= 2e-2, (b, rb, r_last, r_first)


@pytest.mark.parametrize("split_tail", [False, True])
def test_full_depth_cfg_batch_equals_two_single_forwards(split_tail):
    """42 layers, B = 2 (prediction / planning, reference :832-875) vs the same two items run one at a time.
    Without the attention split tail every kernel reduces over K / keys in an order that does not depend on the row's
    position in the batch: bit-equal.  With it (product default) the rows of the last, partially filled attention wave are
    merged from key-range partials, and WHICH rows those are depends on the item count (2832 vs 5664 items): their fp32
    merge order differs, and 42 layers of bf16 rounding spread that like any other rounding difference (measured rel-RMS
    1.0e-2, the size of the bf16 noise of the forward itself: a 2-layer forward is 4e-3 from the fp32 oracle)."""
    from aether_b200.transformer import AetherTransformer3D
    model = AetherTransformer3D(device=torch.device(DEV)).init_synthetic_(0)
    model.attention_split_tail = split_tail
    model.pack(release_unpacked=True)
    c = model.config
    from aether_b200.rope import prepare_rotary_positional_embeddings
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(2, F_LAT, c.in_channels, H_LAT, W_LAT, device=DEV, generator=g).bfloat16()
    e = (torch.randn(2, ST, c.text_embed_dim, device=DEV, generator=g) * 0.2).bfloat16()
    cos, sin = prepare_rotary_positional_embeddings(480, 720, F_LAT, patch_size=2, vae_scale_factor_spatial=8,
                                                    sample_height=60, sample_width=90, attention_head_dim=64,
                                                    base_fps=12, fps=12, device=torch.device(DEV))
    t = torch.tensor([739, 739], device=DEV)
    both = model(x, e, t, image_rotary_emb=(cos, sin))[0]
    singles = torch.cat([model(x[i:i + 1], e[i:i + 1], t[i:i + 1], image_rotary_emb=(cos, sin))[0] for i in range(2)])
    assert torch.isfinite(both.float()).all()
    rel, mx = _rel(both, singles)
    print(f"full depth B=2 vs 2 x B=1 (split_tail={split_tail}): rel-rms {rel:.3e}, max-abs {mx:.3e}, "
          f"bit-equal {torch.equal(both, singles)}")
    if split_tail:
        assert rel <= 2.5e-2, rel
    else:
        assert torch.equal(both, singles)
    assert _rel(both[0], both[1])[0] > 0.1              # the two items really differ (the check is not vacuous)
    model.release()


# ---------------------------------------------------------------------------------------------------- VAE, real geometry
def _vae_pair(seed=1):
    from oracle.vae import OracleVAE, VAEConfig, seeded_vae_init_
    from aether_b200.vae import AetherVAE
    cfg = VAEConfig()
    oracle = seeded_vae_init_(OracleVAE(cfg), seed=seed).eval()
    with torch.no_grad():
        for p in oracle.parameters():
            p.copy_(p.bfloat16().float())
    vae = AetherVAE(**cfg.to_dict())
    vae.load_state_dict(oracle.state_dict(), strict=True)
    for m in (oracle, vae):
        m.enable_slicing()
        m.enable_tiling()
    return cfg, oracle, vae.to(DEV).pack()


def test_vae_encode_tile_row_matches_fp32_oracle():
    cfg, oracle, vae = _vae_pair()
    g = torch.Generator().manual_seed(3)
    yy = torch.linspace(0, 1, 160)[None, None, :, None]
    xx = torch.linspace(0, 1, 432)[None, None, None, :]
    tt = torch.linspace(0, 1, 17)[None, :, None, None]
    ph = torch.rand(3, 1, 1, 1, generator=g) * 6.28
    x = (0.6 * torch.sin(6.28 * (xx * 3 + tt) + ph) * torch.cos(6.28 * (yy * 2 - tt) + ph)
         + 0.1 * torch.randn(3, 17, 160, 432, generator=g)).clamp(-1, 1).bfloat16()[None]
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = oracle.encode(x.float()).latent_dist
    cpu_s = time.perf_counter() - t0
    post = vae.encode(x.to(DEV)).latent_dist
    got_mean = post.mode()
    assert got_mean.shape == ref.mean.shape == (1, 16, 5, 20, 54)
    rel, mx = _rel(got_mean, ref.mean)
    print(f"full-geometry VAE encode 17x160x432 (2 tiles, 2 frame batches): mean rel-rms {rel:.3e}, max-abs {mx:.3e}; "
          f"oracle {cpu_s:.1f} s")
    assert rel <= 2.5e-2, (rel, mx)
    # sample = mean + std * noise with the caller's noise stream: same generator state -> same draw as the oracle
    gen = torch.Generator().manual_seed(5)
    z_ref = ref.mean + ref.std * torch.randn(ref.mean.shape, generator=gen, dtype=torch.bfloat16).float()
    z = post.sample(torch.Generator().manual_seed(5))
    rel_z, _ = _rel(z, z_ref)
    assert rel_z <= 2.5e-2, rel_z


def test_vae_decode_tile_row_matches_fp32_oracle():
    cfg, oracle, vae = _vae_pair()
    g = torch.Generator().manual_seed(4)
    z = torch.randn(1, 16, 4, 20, 54, generator=g).bfloat16()
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = oracle.decode(z.float()).sample
    cpu_s = time.perf_counter() - t0
    got = vae.decode(z.to(DEV)).sample
    assert got.shape == ref.shape == (1, 3, 16, 160, 432) and got.dtype == torch.bfloat16
    rel, mx = _rel(got, ref)
    print(f"full-geometry VAE decode 4x20x54 -> 16x160x432 (2 tiles, 2 frame batches): rel-rms {rel:.3e}, "
          f"max-abs {mx:.3e} (rms ref {ref.pow(2).mean().sqrt():.3f}); oracle {cpu_s:.1f} s")
    assert rel <= 2.5e-2, (rel, mx)
    # the horizontal tile seam (latent column 36 -> pixel 288, blended over 72 px) is not worse than the interior
    assert _rel(got[..., 288:360], ref[..., 288:360])[0] <= 3e-2


# ---------------------------------------------------------------------------------------------------- configs 2/3/4 at full size
@pytest.fixture(scope="module")
def full_pipeline():
    from aether_b200.pipeline import AetherV1PipelineCogVideoX
    from aether_b200.scheduler import AetherDPMScheduler
    from aether_b200.transformer import AetherTransformer3D
    from aether_b200.vae import AetherVAE
    dev = torch.device(DEV)
    tr = AetherTransformer3D(device=dev).init_synthetic_(0).pack(release_unpacked=True)
    vae = AetherVAE(device=dev).init_synthetic_(1)
    vae.enable_slicing()
    vae.enable_tiling()
    emb = torch.randn(1, ST, 4096, generator=torch.Generator().manual_seed(3)) * 0.2
    pipe = AetherV1PipelineCogVideoX(vae=vae, scheduler=AetherDPMScheduler(), transformer=tr,
                                     empty_prompt_embeds=emb).to(dev)
    yield pipe
    tr.release()


@pytest.mark.parametrize("task", ["reconstruction", "prediction", "planning"])
def test_pipeline_tasks_at_full_size(full_pipeline, task):
    rng = np.random.default_rng(0)
    video = rng.random((41, 480, 720, 3), dtype=np.float32)
    raymap = rng.standard_normal((41, 6, 60, 90)).astype(np.float32)
    kw = {"reconstruction": dict(video=video), "prediction": dict(image=video[0], raymap=raymap),
          "planning": dict(image=video[0], goal=video[-1])}[task]

    def run():
        return full_pipeline(task=task, height=480, width=720, num_frames=41, fps=12, num_inference_steps=2,
                             generator=torch.Generator(device=DEV).manual_seed(42), **kw)
    a = run()
    assert a.rgb.shape == (41, 480, 720, 3) and a.disparity.shape == (41, 480, 720) and a.raymap.shape == (41, 6, 60, 90)
    assert a.rgb.dtype == a.disparity.dtype == a.raymap.dtype == np.float32
    assert np.isfinite(a.rgb).all() and np.isfinite(a.disparity).all() and np.isfinite(a.raymap).all()
    assert a.rgb.min() >= 0.0 and a.rgb.max() <= 1.0 and a.disparity.min() >= 0.0     # clamp / square (:932-939)
    if task == "reconstruction":
        assert full_pipeline.guidance_scale == 1.0                                     # no CFG (:777)
    else:
        # dynamic CFG on the last timestep t = 499 of a 2-step trailing schedule (:879-893): python-float formula
        import math
        want = 1 + 3.0 * ((1 - math.cos(math.pi * ((2 - 499) / 2) ** 5.0)) / 2)
        assert full_pipeline.guidance_scale == want
    b = run()
    assert np.array_equal(a.disparity, b.disparity) and np.array_equal(a.raymap, b.raymap)   # same seed, same result
    lat = full_pipeline(task=task, height=480, width=720, num_frames=41, fps=12, num_inference_steps=2,
                        generator=torch.Generator(device=DEV).manual_seed(42), output_latents=True, **kw)
    assert lat.shape == (1, 11, 56, 60, 90) and lat.dtype == torch.bfloat16
    # raymap output = un-folded camera latents, last 41 of the 44 frames (:942-949)
    cam = lat[:, :, 32:].float().cpu().numpy()
    unfold = cam.reshape(1, 11, 4, 6, 60, 90).transpose(0, 2, 1, 3, 4, 5).reshape(1, 44, 6, 60, 90)[0, -41:]
    assert np.array_equal(unfold, a.raymap)
