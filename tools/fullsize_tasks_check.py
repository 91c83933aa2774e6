"""All three tasks of the pipeline at FULL size (41 x 480 x 720, AetherV1 DiT + CogVideoX-5b VAE geometry, synthetic
weights): reconstruction (B=1), prediction with a raymap and planning (CFG, B=2, dynamic guidance).  Prints the
seconds per call for `steps` denoise steps, output shapes and finiteness.  usage: fullsize_tasks_check.py [steps]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from aether_b200.pipeline import AetherV1PipelineCogVideoX  # noqa: E402
from aether_b200.scheduler import AetherDPMScheduler  # noqa: E402
from aether_b200.transformer import AetherTransformer3D  # noqa: E402
from aether_b200.vae import AetherVAE  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
tr = AetherTransformer3D(device=dev).init_synthetic_(0).pack(release_unpacked=True)
vae = AetherVAE(device=dev).init_synthetic_(1)
vae.enable_slicing(); vae.enable_tiling()
emb = (torch.randn(1, 226, 4096, generator=torch.Generator().manual_seed(3)) * 0.2)
pipe = AetherV1PipelineCogVideoX(vae=vae, scheduler=AetherDPMScheduler(), transformer=tr, empty_prompt_embeds=emb).to(dev)
rng = np.random.default_rng(0)
video = rng.random((41, 480, 720, 3), dtype=np.float32)
raymap = rng.standard_normal((41, 6, 60, 90)).astype(np.float32)
cases = {
    "reconstruction": dict(task="reconstruction", video=video),
    "prediction": dict(task="prediction", image=video[0], raymap=raymap),
    "planning": dict(task="planning", image=video[0], goal=video[-1]),
}
for name, kw in cases.items():
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = pipe(height=480, width=720, num_frames=41, fps=12, num_inference_steps=steps,
                   generator=torch.Generator(device=dev).manual_seed(42), **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ok = all(np.isfinite(a).all() for a in (out.rgb, out.disparity, out.raymap))
    print(f"{name}: {dt:.2f} s for {steps} steps (2nd call); rgb {out.rgb.shape} disparity {out.disparity.shape} "
          f"raymap {out.raymap.shape}; finite={ok}; guidance now {pipe.guidance_scale}", flush=True)
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
