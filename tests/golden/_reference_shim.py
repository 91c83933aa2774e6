"""Import the UNMODIFIED reference sources from /root/reference in this container.

The reference's Python files import third-party packages that are not installed here (diffusers,
accelerate, imageio, rootutils, matplotlib, plyfile).  This shim registers minimal stand-ins in
`sys.modules` -- only the names the reference files touch -- so that the reference's OWN code
(aether/pipelines/aetherv1_pipeline_cogvideox.py, aether/utils/postprocess_utils.py,
evaluation/video_depth/launch_aether.py) can be imported and executed to produce golden vectors.

The stand-in for diffusers' `CogVideoXImageToVideoPipeline` base class restates the inherited members the
reference relies on; the three module objects injected into the reference pipeline are the
fp32 oracle modules (oracle/dit.py, oracle/vae.py, oracle/scheduler.py).  Everything the reference itself
owns (input checks, preprocessing, 96-channel latent assembly, RoPE table, CFG/denoise loop, output split,
raymap fold/unfold, tiling plan, compute_scale, blend chain) therefore runs from the reference's files.

Only used by tests/golden/make_golden.py (here, where /root/reference exists) -- never on the GPU box.
"""
from __future__ import annotations

import inspect
import sys
import types
from pathlib import Path

import numpy as np
import torch

REFERENCE_ROOT = Path("/root/reference")
REPO_ROOT = Path(__file__).resolve().parent.parent.parent


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class _AutoStub(types.ModuleType):
    """A module whose every attribute is a stub object (callable, subclassable, attribute-chainable)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, v)
        return v


def _install_autostubs(roots):
    import importlib.abc
    import importlib.machinery

    class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path=None, target=None):
            if fullname.split(".")[0] in roots and fullname not in sys.modules:
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            m = _AutoStub(spec.name)
            m.__path__ = []
            return m

        def exec_module(self, module):
            pass

    for r in roots:             # drop the minimal stand-ins registered earlier for these roots' submodules
        for k in [k for k in sys.modules if k == r or k.startswith(r + ".")]:
            if not isinstance(sys.modules[k], _AutoStub) and k not in ("matplotlib", "imageio", "imageio.v3"):
                pass
    sys.meta_path.append(Finder())


def install():
    if "aether_reference_shim_installed" in sys.modules:
        return
    if not REFERENCE_ROOT.exists():
        raise RuntimeError("/root/reference is not available (the shim only works in the build container)")
    if str(REPO_ROOT) not in sys.path:
        sys.path.insert(0, str(REPO_ROOT))
    import transformers  # noqa: F401  (real package; must be imported before the stand-ins exist: it probes
    from transformers import AutoTokenizer, T5EncoderModel  # noqa: F401   `accelerate` with importlib at import)
    from oracle.rope import get_1d_rotary_pos_embed as _oracle_rope1d
    from oracle.scheduler import OracleDPMScheduler
    from oracle.video_processor import OracleVideoProcessor as VideoProcessor   # independent of the product's copy

    # ---------------------------------------------------------------- diffusers
    d = _mod("diffusers")

    class AutoencoderKLCogVideoX:  # noqa: D401 - type placeholder
        pass

    class CogVideoXTransformer3DModel:
        pass

    class CogVideoXImageToVideoPipeline:
        """Inherited surface used by the reference."""

        def __init__(self, tokenizer, text_encoder, vae, scheduler, transformer):
            self.tokenizer, self.text_encoder = tokenizer, text_encoder
            self.vae, self.scheduler, self.transformer = vae, scheduler, transformer
            self.vae_scale_factor_spatial = 2 ** (len(vae.config.block_out_channels) - 1)
            self.vae_scale_factor_temporal = vae.config.temporal_compression_ratio
            self.vae_scaling_factor_image = vae.config.scaling_factor
            self.video_processor = VideoProcessor(vae_scale_factor=self.vae_scale_factor_spatial)
            self._interrupt = False
            self._guidance_scale = None

        def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance=False,
                          num_videos_per_prompt=1, prompt_embeds=None, **kw):
            return self.text_encoder(prompt), None      # test double: callable returning [1, St, text_dim]

        @property
        def _execution_device(self):
            return torch.device("cpu")

        @property
        def guidance_scale(self):
            return self._guidance_scale

        @property
        def interrupt(self):
            return self._interrupt

        def prepare_extra_step_kwargs(self, generator, eta):
            params = set(inspect.signature(self.scheduler.step).parameters.keys())
            kw = {}
            if "eta" in params:
                kw["eta"] = eta
            if "generator" in params:
                kw["generator"] = generator
            return kw

        def decode_latents(self, latents):
            latents = latents.permute(0, 2, 1, 3, 4)
            latents = 1 / self.vae_scaling_factor_image * latents
            return self.vae.decode(latents).sample

        def progress_bar(self, total=None):
            class _PB:
                def __enter__(s):
                    return s

                def __exit__(s, *a):
                    return False

                def update(s):
                    pass
            return _PB()

        def maybe_free_model_hooks(self):
            pass

    d.AutoencoderKLCogVideoX = AutoencoderKLCogVideoX
    d.CogVideoXTransformer3DModel = CogVideoXTransformer3DModel
    d.CogVideoXImageToVideoPipeline = CogVideoXImageToVideoPipeline
    d.CogVideoXDPMScheduler = OracleDPMScheduler          # isinstance(...) at reference :902
    ip = _mod("diffusers.image_processor")
    ip.PipelineImageInput = object
    _mod("diffusers.models")
    emb = _mod("diffusers.models.embeddings")

    def get_1d_rotary_pos_embed(dim, pos, theta=10000.0, use_real=False, **kw):
        assert use_real
        return _oracle_rope1d(dim, pos, theta)

    emb.get_1d_rotary_pos_embed = get_1d_rotary_pos_embed
    u = _mod("diffusers.utils")

    class BaseOutput:
        pass

    u.BaseOutput = BaseOutput
    tu = _mod("diffusers.utils.torch_utils")

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        gdev = generator.device if generator is not None else (device or torch.device("cpu"))
        return torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype).to(device)

    tu.randn_tensor = randn_tensor

    # ---------------------------------------------------------------- other absent packages
    acc = _mod("accelerate")
    acc.Accelerator = type("Accelerator", (), {"__init__": lambda self, **kw: None})
    acc.PartialState = type("PartialState", (), {})
    im = _mod("imageio")
    im.__path__ = []                     # a package: imageio.v2 etc. resolve through the auto-stub finder below
    im.v3 = _mod("imageio.v3")
    ru = _mod("rootutils")
    ru.setup_root = lambda *a, **k: None
    mpl = _mod("matplotlib")
    mpl.__path__ = []
    mpl.colormaps = {}
    pf = _mod("plyfile")
    pf.PlyData = pf.PlyElement = object
    # filterpy: a real (restated) Kalman filter, the smoothing code computes with it (oracle/kalman.py)
    from oracle.kalman import KalmanFilter as _OracleKalman
    fp = _mod("filterpy")
    fpk = _mod("filterpy.kalman")
    fpk.KalmanFilter = _OracleKalman
    fp.kalman = fpk
    # everything else the launcher / demo scripts import at module level but never touch on the paths we execute
    # (evo.*, trimesh, gradio, matplotlib.*, imageio.v2, tqdm is real): import-anything stand-ins
    _install_autostubs(("evo", "trimesh", "gradio", "matplotlib", "imageio", "viser", "moviepy"))
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    sys.modules["aether_reference_shim_installed"] = types.ModuleType("aether_reference_shim_installed")


def reference_pipeline_module():
    install()
    import aether.pipelines.aetherv1_pipeline_cogvideox as P
    return P


def reference_postprocess_module():
    install()
    import aether.utils.postprocess_utils as POST
    return POST


def reference_sliding_window_module():
    install()
    import evaluation.video_depth.launch_aether as EVD
    return EVD


def reference_depth_tools_module():
    """evaluation/video_depth/tools.py imports only numpy / torch / scipy: loaded as a plain file module."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("aether_reference_depth_tools",
                                                  str(REFERENCE_ROOT / "evaluation" / "video_depth" / "tools.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_rel_pose_module():
    install()
    import evaluation.rel_pose.launch_aether as EVP
    return EVP


def reference_demo_module():
    install()
    import scripts.demo as DEMO
    return DEMO
