"""Where does one sliding-window tile (41 x 480 x 720, `--steps` denoising steps) spend its time OUTSIDE the DiT and the
VAE?  Phase wall-clock (device-synchronised) of AetherV1PipelineCogVideoX.__call__ plus a torch.profiler table of every
CUDA kernel and the CPU-side gaps.  usage: python tools/pipeline_breakdown.py [--steps 4] [--host-input] [--json out]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from aether_b200.pipeline import AetherV1PipelineCogVideoX  # noqa: E402
from aether_b200.scheduler import AetherDPMScheduler  # noqa: E402
from aether_b200.transformer import AetherTransformer3D  # noqa: E402
from aether_b200.vae import AetherVAE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--host-input", action="store_true", help="host uint8 frames (the e2e path) instead of device uint8")
    ap.add_argument("--no-rgb", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = AetherTransformer3D(device=dev)
    model.init_synthetic_(seed=0)
    model.pack(release_unpacked=True)
    vae = AetherVAE(device=dev)
    vae.init_synthetic_(seed=1)
    vae.enable_slicing()
    vae.enable_tiling()
    text = torch.randn(1, 226, 4096, generator=torch.Generator().manual_seed(3)) * 0.2
    pipe = AetherV1PipelineCogVideoX(vae=vae, scheduler=AetherDPMScheduler(), transformer=model,
                                     empty_prompt_embeds=text).to(dev)
    frames = np.random.default_rng(0).integers(0, 256, (41, 480, 720, 3), dtype=np.uint8)
    video = frames if args.host_input else torch.from_numpy(frames).to(dev)

    def call():
        return pipe(video=video, num_inference_steps=args.steps, num_frames=41, fps=12, return_dict=False,
                    generator=torch.Generator(device=dev).manual_seed(3407), output_type="pt",
                    decode_rgb=not args.no_rgb)

    for _ in range(2):
        call()
    torch.cuda.synchronize()

    # ---- phase timing: wrap the methods the call goes through
    phases = {}

    def timed(obj, name, label):
        fn = getattr(obj, name)

        def wrapper(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            phases[label] = phases.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
            return r
        setattr(obj, name, wrapper)
        return fn

    saved = [(pipe, "preprocess_inputs", timed(pipe, "preprocess_inputs", "preprocess_inputs")),
             (pipe, "prepare_latents", timed(pipe, "prepare_latents", "prepare_latents (VAE encode + noise)")),
             (pipe, "_prepare_rotary_positional_embeddings",
              timed(pipe, "_prepare_rotary_positional_embeddings", "rotary table")),
             (pipe.scheduler, "set_timesteps", timed(pipe.scheduler, "set_timesteps", "set_timesteps")),
             (pipe.scheduler, "step_fused", timed(pipe.scheduler, "step_fused", "scheduler.step_fused (all steps)")),
             (pipe.transformer, "forward_split", timed(pipe.transformer, "forward_split", "DiT forward (all steps)")),
             (pipe, "decode_latents", timed(pipe, "decode_latents", "decode_latents (VAE decode, all)")),
             (pipe.vae, "encode", timed(pipe.vae, "encode", "  of which vae.encode"))]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
    for obj, name, fn in saved:
        setattr(obj, name, fn)
    top = sum(v for k, v in phases.items() if not k.startswith("  "))
    phases["(everything else in __call__)"] = total - top
    phases["total (with per-phase synchronisation)"] = total

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    free_running = (time.perf_counter() - t0) * 1e3

    # ---- kernels
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        call()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        dt = getattr(e, "device_time_total", 0.0) or getattr(e, "cuda_time_total", 0.0)
        if dt > 0 and e.device_type == torch.autograd.DeviceType.CUDA:
            rows.append({"name": e.key[:100], "count": e.count, "device_ms": dt / 1e3})
    rows.sort(key=lambda r: -r["device_ms"])
    kernel_sum = sum(r["device_ms"] for r in rows)
    out = {"steps": args.steps, "host_input": args.host_input, "decode_rgb": not args.no_rgb,
           "free_running_ms": free_running, "phases_ms": phases, "sum_kernel_ms": kernel_sum, "kernels": rows[:45]}
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}, indent=1))
    for r in rows[:45]:
        print(f"{r['device_ms']:10.3f} ms  x{r['count']:<6d} {r['name']}")
    if args.json:
        Path(args.json).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
