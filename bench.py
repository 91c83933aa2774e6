#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: latent-frames/sec for a 41x480x720 50-step generation.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload = BASELINE.json configs[1] (reconstruction: 41 frames 480x720 -> 11 latent frames of 60x90,
S = 226 + 14850 = 15076 tokens, 50 DPM steps, guidance 1.0 => batch 1, bf16), synthetic seeded weights of the
full AetherV1 / CogVideoX-5b geometry (42 layers, 48 x 64, 5.55 B parameters) and synthetic latents.

A "step" is one pass of the denoise-loop body (reference aetherv1_pipeline_cogvideox.py:827-921): 96-channel
concat -> DiT forward (aether_dit_forward) -> fused CFG/DPM-Solver++ step (aether_cfg_dpm_step).  One 50-step
generation yields 11 latent frames, so  value = N_gpus * 11 * (K / 50) / seconds.
  value : inputs resident in HBM, CUDA-event time over exactly K steps, max over ranks.
  e2e   : the public API end to end: AetherV1PipelineCogVideoX.__call__(task="reconstruction", 41x480x720, 50
          steps) with a HOST numpy video in and HOST numpy rgb/disparity/raymap out -- preprocessing, H2D, VAE
          encode, the loop, 2x VAE decode and D2H inside the timed region (value = N * 11 / seconds per call);
          `denoise_step_pinned_host` additionally reports the bare step with pinned host buffers.
  roofline : the attention kernel (dominant), duration from CUDA events recorded around every attention launch
          on the launching stream inside the timed region; algorithmic flop = 4*B*H*S^2*64 per launch.
  cpu_baseline / --impl reference : the reference's CPU path for this loop = the fp32 torch restatement of the
          diffusers modules (oracle/), timed on the host cores on a bounded sample (one transformer block of one
          forward at full S plus embed/tail) and extrapolated x42 layers x50 steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LATENT_FRAMES, LAT_H, LAT_W = 11, 60, 90
TEXT_LEN, TEXT_DIM = 226, 4096
HEADS, LAYERS = 48, 42
S_TOKENS = TEXT_LEN + LATENT_FRAMES * (LAT_H // 2) * (LAT_W // 2)
DENOISE_STEPS = 50
WORKLOAD = "configs[1]: reconstruction, 41 frames 480x720 -> 11x60x90 latents, S=15076, 50 steps, batch 1 (no CFG)"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained 1.4 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if "Active" in v and "Not" not in v:
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_sample(threads: int, n_samples: int = 1):
    """One transformer block (+ embed and tail) of one forward at full S on the host cores, fp32 oracle."""
    import torch
    from oracle.dit import DiTConfig, OracleDiT
    from oracle.rope import prepare_rotary_positional_embeddings
    torch.set_num_threads(threads)
    cfg = DiTConfig(num_layers=1)
    torch.manual_seed(0)
    model = OracleDiT(cfg).eval()
    x = torch.randn(1, LATENT_FRAMES, 96, LAT_H, LAT_W)
    e = torch.randn(1, TEXT_LEN, TEXT_DIM) * 0.2
    cos, sin = prepare_rotary_positional_embeddings(480, 720, LATENT_FRAMES)
    ts = torch.tensor([999])
    times_full, times_nolayer = [], []
    with torch.no_grad():
        for _ in range(n_samples):
            t0 = time.perf_counter()
            model(x, e, ts, image_rotary_emb=(cos, sin))
            times_full.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            model(x, e, ts, image_rotary_emb=(cos, sin), n_layers=0)
            times_nolayer.append(time.perf_counter() - t0)
    t_full, t_rest = min(times_full), min(times_nolayer)
    t_layer = max(t_full - t_rest, 1e-9)
    t_forward = LAYERS * t_layer + t_rest
    value = LATENT_FRAMES / (DENOISE_STEPS * t_forward)
    return dict(value=value, t_layer_s=t_layer, t_embed_tail_s=t_rest, t_forward_extrapolated_s=t_forward,
                cpu_work_s=sum(times_full) + sum(times_nolayer))


def cpu_baseline_dict(r, threads, kind="port"):
    return {"value": r["value"], "unit": "latent-frames/s", "cores": threads, "kind": kind,
            "sample": (f"1 of {LAYERS} transformer blocks + embed/tail of ONE forward at full S={S_TOKENS} "
                       f"(fp32 torch restatement of the diffusers modules, {r['t_layer_s']:.2f} s/block), "
                       f"extrapolated x{LAYERS} layers x{DENOISE_STEPS} steps; diffusers itself is not installable here")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = max(1, min(args.steps, 3))
    r = cpu_reference_sample(threads, n_samples=n)
    ms_per_step = r["t_forward_extrapolated_s"] * 1000.0
    line = {"impl": "reference", "metric": "latent-frames/sec, 41x480x720 50-step generation (denoise loop)",
            "value": r["value"], "unit": "latent-frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "device": "host CPU", "samples_timed": n},
            "cpu_baseline": cpu_baseline_dict(r, threads),
            "e2e": {"value": r["value"], "unit": "latent-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ product arm
def run_product(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    from aether_b200 import _lib
    if not _lib.lib_path().exists():
        ge.build()
    from aether_b200.rope import prepare_rotary_positional_embeddings
    from aether_b200.scheduler import AetherDPMScheduler
    from aether_b200.transformer import AetherTransformer3D

    model = AetherTransformer3D(device=dev)
    model.init_synthetic_(seed=0)
    model.pack(release_unpacked=True)
    sched = AetherDPMScheduler()
    sched.set_timesteps(DENOISE_STEPS, device=dev)
    t_host = [int(v) for v in sched.timesteps.tolist()]
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    latents = torch.randn(1, LATENT_FRAMES, 56, LAT_H, LAT_W, device=dev, generator=g, dtype=torch.bfloat16)
    cond = (torch.randn(1, LATENT_FRAMES, 40, LAT_H, LAT_W, device=dev, generator=g) * 0.7).to(torch.bfloat16)
    text = (torch.randn(1, TEXT_LEN, TEXT_DIM, device=dev, generator=g) * 0.2).to(torch.bfloat16)
    rope = prepare_rotary_positional_embeddings(480, 720, LATENT_FRAMES, patch_size=2, vae_scale_factor_spatial=8,
                                                sample_height=60, sample_width=90, attention_head_dim=64,
                                                base_fps=12, fps=12, device=dev)
    noise_gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    state = {"latents": latents, "old": None}

    def step(i, lat_in=None):
        k = i % DENOISE_STEPS
        if k == 0:
            state["old"] = None
        lat = state["latents"] if lat_in is None else lat_in
        # the 96-channel concat of reference :857 happens inside the patch-gather kernel (aether_dit_forward_split)
        out = model.forward_split(lat, cond, text, sched.timesteps[k].reshape(1), rope)
        new, state["old"] = sched.step_fused(out, 1.0, state["old"], t_host[k], t_host[k - 1] if k > 0 else None, lat,
                                             generator=noise_gen)
        state["latents"] = new
        return new

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    model.enable_timing(True)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    attn_ms, attn_n = model.read_attention_timing()
    model.enable_timing(False)
    tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax.item())
    value = world * LATENT_FRAMES * (args.steps / DENOISE_STEPS) / (elapsed_ms / 1000.0)

    # ---- e2e: same step, pinned host buffers in the timed region
    n_e2e = max(1, min(args.steps, 10))
    host_in = torch.empty(1, LATENT_FRAMES, 56, LAT_H, LAT_W, dtype=torch.bfloat16).pin_memory()
    host_cond = cond.cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    host_in.copy_(state["latents"].cpu())
    dev_in = torch.empty_like(latents)
    h2d = host_in.numel() * 2 + host_cond.numel() * 2
    d2h = host_out.numel() * 2
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_e2e):
        dev_in.copy_(host_in, non_blocking=True)
        cond.copy_(host_cond, non_blocking=True)
        new = step(i, dev_in)
        host_out.copy_(new, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller consumes the step result on the host
        host_in.copy_(host_out)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_step_value = world * LATENT_FRAMES * (n_e2e / DENOISE_STEPS) / (float(e2e_ms.item()) / 1000.0)
    e2e = {"value": e2e_step_value, "unit": "latent-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": n_e2e, "scope": "denoise step via the public modules with pinned host buffers; VAE not included"}

    # ---- e2e proper: the public API.  AetherV1PipelineCogVideoX.__call__ (reconstruction, 41 x 480 x 720, 50 steps)
    # with a HOST numpy video in and HOST numpy rgb / disparity / raymap out: input preprocessing, H2D, VAE encode,
    # the 50-step loop, 2 x VAE decode and D2H are all inside the timed region.
    if not args.no_full_e2e:
        import numpy as np
        from aether_b200.pipeline import AetherV1PipelineCogVideoX
        from aether_b200.vae import AetherVAE
        vae = AetherVAE(device=dev)
        vae.init_synthetic_(seed=1)
        vae.enable_slicing()
        vae.enable_tiling()
        pipe = AetherV1PipelineCogVideoX(vae=vae, scheduler=AetherDPMScheduler(), transformer=model,
                                         empty_prompt_embeds=text.cpu()).to(dev)
        rng = np.random.default_rng(7 + rank)
        video = rng.random((41, 480, 720, 3), dtype=np.float32)

        def call(n):
            return pipe(task="reconstruction", video=video, height=480, width=720, num_frames=41, fps=12,
                        num_inference_steps=n, generator=torch.Generator(device=dev).manual_seed(42 + rank))

        call(1)                                                     # warm-up call (VAE kernels, allocator)
        barrier()
        t0 = time.perf_counter()
        out = call(DENOISE_STEPS)
        torch.cuda.synchronize()
        tc = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        t_call = float(tc.item())
        e2e = {"value": world * LATENT_FRAMES / t_call, "unit": "latent-frames/s",
               "h2d_bytes_per_step": int(video.size * 2),
               "d2h_bytes_per_step": int(out.rgb.nbytes + out.disparity.nbytes + out.raymap.nbytes),
               "seconds_per_call": t_call, "calls": 1,
               "scope": "AetherV1PipelineCogVideoX.__call__(reconstruction, 41x480x720, 50 steps): host numpy video -> "
                        "preprocess -> H2D -> VAE encode -> 50 x (DiT + DPM step) -> 2 x VAE decode -> D2H numpy outputs "
                        "(bytes are per call)",
               "denoise_step_pinned_host": {"value": e2e_step_value, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                            "steps": n_e2e}}

    if rank == 0:
        peak, peak_src = _peaks()
        flop = 4.0 * 1 * HEADS * float(S_TOKENS) ** 2 * 64
        avg_ms = attn_ms / max(attn_n, 1)
        achieved = flop / (avg_ms / 1000.0) / 1e12 if attn_n else None
        traffic = None
        prof = ROOT / "profiles" / "attention_traffic.json"
        if prof.exists():
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        line = {
            "metric": "latent-frames/sec, 41x480x720 50-step generation (denoise loop)",
            "value": value, "unit": "latent-frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "model": "AetherV1 DiT geometry (42 layers, 48x64, in 96 / out 56), seeded synthetic weights",
                       "denoise_steps": DENOISE_STEPS, "tokens": S_TOKENS, "attention_mode": int(model.attention_fp16_pv),
                       "parallelism": f"replicas x{world} (independent tiles per GPU, no data-path collective)",
                       "l2": "per-step working set (11.1 GB weights + ~1.4 GB activations) >> 126 MB L2; no explicit flush"},
            "e2e": e2e,
            "gpu_launches": (model.launches_per_forward(1) + 1) * args.steps,
            "clocks": clocks,
            "roofline": {"kernel": f"aether_attention_bf16 mode {int(model.attention_fp16_pv)} (tcgen05, attention_v3_kernel)",
                         "bound": "tensor", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "peak_source": peak_src, "launches_timed": attn_n, "avg_launch_ms": avg_ms,
                         "share_of_step": attn_ms / (elapsed_ms / args.steps) if attn_n else None},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_baseline_dict(cpu_reference_sample(threads, 1), threads)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-e2e", action="store_true", help="skip the full pipeline __call__ e2e measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
