"""Full-size AetherVAE (CogVideoX-5b geometry, synthetic weights): encode a 41x480x720 clip and decode 11x60x90
latents with tiling + slicing, print timings.  Run with CUDA_LAUNCH_BLOCKING=1 to localise a failing launch."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200.vae import AetherVAE  # noqa: E402

dev = torch.device("cuda", 0)
vae = AetherVAE(device=dev).init_synthetic_(seed=1)
vae.enable_slicing()
vae.enable_tiling()
vae.pack()
print("params", sum(p.numel() for p in vae.parameters()) / 1e6, "M", flush=True)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 41
x = (torch.rand(1, 3, frames, 480, 720, device=dev) * 2 - 1).to(torch.bfloat16)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    post = vae.encode(x).latent_dist
    z = post.sample(torch.Generator(device=dev).manual_seed(0))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"encode {frames}f: {t1 - t0:.3f} s  z {tuple(z.shape)} finite={bool(torch.isfinite(z.float()).all())} "
          f"std={z.float().std().item():.3f}", flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y = vae.decode(z).sample
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"decode: {t1 - t0:.3f} s  y {tuple(y.shape)} finite={bool(torch.isfinite(y.float()).all())} "
          f"std={y.float().std().item():.3f}", flush=True)
print("mem GB", torch.cuda.max_memory_allocated() / 1e9)

if "--breakdown" in sys.argv:
    # CUDA-event time per kernel family (wrapping the module's own kernel-call helpers), one decode + one encode
    import collections
    rec = []

    def wrap(name):
        orig = getattr(vae, name)

        def f(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(*a, **k)
            e.record()
            tag = name
            if name == "_conv":
                p = vae._packed[a[1]]
                tag = f"_conv k{p['k']} cin{p['cin']} cout{p['cout']}"
            rec.append((tag, s, e))
            return r
        setattr(vae, name, f)

    for n in ("_conv", "_norm", "_upsample", "_downsample", "_fill_time_pad", "_to_cl", "_blend"):
        wrap(n)
    for what, fn in (("decode", lambda: vae.decode(z)), ("encode", lambda: vae.encode(x).latent_dist.mode())):
        rec.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
        agg = collections.defaultdict(float); cnt = collections.Counter()
        for tag, s, e in rec:
            agg[tag] += s.elapsed_time(e); cnt[tag] += 1
        print(f"--- {what}: wall {wall * 1e3:.0f} ms (nested helpers double count: _norm includes its GEMMs, "
              f"_upsample/_downsample include their conv)")
        for tag, ms in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
            print(f"  {ms:8.1f} ms  x{cnt[tag]:5d}  {tag}")
