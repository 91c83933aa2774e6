"""Full-size AetherVAE (CogVideoX-5b geometry, synthetic weights): encode of a 41 x 480 x 720 clip and decode of
11 x 60 x 90 latents, per-op Python orchestration vs the native schedule (aether_vae_encode / aether_vae_decode),
CUDA-event times (median of 3 after a warm-up call) and the number of launches of the native schedule."""
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200.vae import AetherVAE  # noqa: E402

dev = torch.device("cuda", 0)
vae = AetherVAE(device=dev).init_synthetic_(seed=1)
vae.enable_slicing()
vae.enable_tiling()
vae.pack()
x = (torch.rand(1, 3, 41, 480, 720, device=dev) * 2 - 1).to(torch.bfloat16)
z = torch.randn(1, 16, 11, 60, 90, device=dev).to(torch.bfloat16)


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


res = {}
outs = {}
for per_op in (True, False):
    vae.per_op = per_op
    key = "per_op" if per_op else "native"
    res[key] = {"encode_ms": timed(lambda: vae.encode(x).latent_dist.mode()), "decode_ms": timed(lambda: vae.decode(z).sample)}
    outs[key] = (vae.encode(x).latent_dist.mode().clone(), vae.decode(z).sample.clone())
res["bit_identical"] = bool(torch.equal(outs["per_op"][0], outs["native"][0]) and torch.equal(outs["per_op"][1], outs["native"][1]))
res["native_launches"] = {"encode": vae.launches(0, 41, 480, 720), "decode": vae.launches(1, 11, 60, 90)}
res["workspace_gb"] = vae._ws.numel() / 1e9
res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(res), flush=True)
