"""Per-kernel device time of one full-size VAE encode (41 x 480 x 720) + decode (11 x 60 x 90 latents) through the
native schedule, aggregated with torch.profiler (CUPTI sees the kernels this library launches as well).
Prints a JSON table: kernel family -> {count, total_ms, share}."""
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200.vae import AetherVAE  # noqa: E402

dev = torch.device("cuda", 0)
vae = AetherVAE(device=dev).init_synthetic_(seed=1)
vae.enable_slicing()
vae.enable_tiling()
vae.pack()
x = (torch.rand(1, 3, 41, 480, 720, device=dev) * 2 - 1).to(torch.bfloat16)
z = torch.randn(1, 16, 11, 60, 90, device=dev).to(torch.bfloat16)
for what, fn in (("encode", lambda: vae.encode(x).latent_dist.mode()), ("decode", lambda: vae.decode(z).sample)):
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
    wall = a.elapsed_time(b)
    fam = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type.name != "CUDA":
            continue
        name = ev.name
        m = re.search(r"(conv_kernel<[^>]*>|gemm2?_kernel<[^>]*>|gn_apply_kernel<[^>]*>|gn_partial_kernel|gn_finalize_kernel|"
                      r"upsample_nearest_kernel|avgpool_time_kernel|crop_ncthw_to_thwc_kernel|copy_region_kernel|"
                      r"tile_blend_kernel|thwc_to_ncthw_kernel|Memcpy DtoD|Memset)", name)
        key = m.group(1) if m else name[:60]
        fam[key][0] += 1
        fam[key][1] += ev.device_time / 1000.0 if hasattr(ev, "device_time") else ev.cuda_time / 1000.0
    total = sum(v[1] for v in fam.values())
    table = {k: {"count": v[0], "total_ms": round(v[1], 3), "share": round(v[1] / total, 4)}
             for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])}
    print(json.dumps({"op": what, "wall_ms": round(wall, 2), "sum_kernel_ms": round(total, 2), "kernels": table}), flush=True)
