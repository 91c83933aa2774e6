"""CPU oracle for the Aether denoising hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain fp32 PyTorch/numpy restatement of the algorithms on the
hot path named by BASELINE.json `north_star`.  It is the
*checker*: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  Nothing under
`aether_b200/` imports `oracle`; the product path fails loudly when the CUDA
extension is missing instead of falling back to this code.

Pinning status (see DESIGN.md "Oracle"):

* `oracle.rope`, `oracle.blend` restate code that lives in /root/reference and
  are PINNED against the reference itself (the reference functions are imported
  in this container by tests/golden/make_golden.py and their outputs committed
  under tests/golden/).  The pipeline glue, the sliding-window launcher, the
  pose / point-map post-processing and the depth evaluation are pinned the same
  way: the reference's OWN code is executed on top of the oracle modules
  (tests/golden/_reference_shim.py) and the product is compared with those outputs.
* `oracle.dit`, `oracle.scheduler`, `oracle.vae`, `oracle.video_processor` restate
  third-party `diffusers` (>=0.32.2, unpinned by the reference: requirements.txt:4)
  modules and `oracle.kalman` restates `filterpy.kalman.KalmanFilter`; none of them
  is present in /root/reference nor installable here (no network).
  PARITY UNPINNED for those: they follow the published diffusers v0.32
  algorithm and are anchored on the reference's call
  sites (aetherv1_pipeline_cogvideox.py:865-875, :907-915, :557, :931) and on
  self-consistency known-answer tests (scheduler end points, RoPE norm
  preservation, unpatchify/patchify inverse).
"""
