"""ctypes binding of libaether_b200.so (the C ABI declared in include/aether_b200.h).

The product path has NO fallback: if the library is missing or a call returns a non-zero status a
RuntimeError is raised.  PyTorch is used by callers only for device memory and streams; this module
passes raw device pointers (`tensor.data_ptr()`) and the current CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libaether_b200.so"
_lib = None

STATUS = {0: "AETHER_OK", 1: "AETHER_ERR_INVALID", 2: "AETHER_ERR_CUDA", 3: "AETHER_ERR_WORKSPACE"}

c_void_p, c_int32, c_int64, c_float, c_double = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class DitConfig(C.Structure):
    _fields_ = [("num_heads", c_int32), ("head_dim", c_int32), ("num_layers", c_int32),
                ("in_channels", c_int32), ("out_channels", c_int32), ("patch_size", c_int32),
                ("time_embed_dim", c_int32), ("text_embed_dim", c_int32), ("flip_sin_to_cos", c_int32),
                ("freq_shift", c_float), ("norm_eps", c_float), ("ff_mult", c_int32),
                ("attention_fp16_pv", c_int32), ("fused_qkv_epilogue", c_int32), ("attention_split_tail", c_int32)]


class DitLayerWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "w_qkv", "b_qkv", "w_out", "b_out", "w_ff1", "b_ff1", "w_ff2", "b_ff2",
        "norm1_g", "norm1_b", "norm2_g", "norm2_b", "qn_g", "qn_b", "kn_g", "kn_b")]


class DitWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "w_time1", "b_time1", "w_time2", "b_time2", "w_text", "b_text", "w_patch", "b_patch",
        "w_adaln", "b_adaln", "normf_g", "normf_b", "normo_g", "normo_b", "w_proj", "b_proj",
        "pos_embedding")] + [("layers", C.POINTER(DitLayerWeights))]


class VaeConfig(C.Structure):
    _fields_ = [("in_channels", c_int32), ("out_channels", c_int32), ("latent_channels", c_int32),
                ("num_blocks", c_int32), ("layers_per_block", c_int32), ("norm_num_groups", c_int32),
                ("norm_eps", c_float), ("temporal_compression_ratio", c_int32), ("use_tiling", c_int32),
                ("num_latent_frames_batch_size", c_int32), ("num_sample_frames_batch_size", c_int32),
                ("tile_sample_min_height", c_int32), ("tile_sample_min_width", c_int32),
                ("tile_latent_min_height", c_int32), ("tile_latent_min_width", c_int32)] + [
        (f"{s}_{k}_{a}", c_int32) for s in ("enc", "dec") for k in ("overlap", "blend", "limit") for a in ("h", "w")]


class VaeParam(C.Structure):
    _fields_ = [("name", C.c_char_p), ("kind", c_int32), ("data", c_void_p), ("bias", c_void_p), ("kt", c_int32),
                ("kh", c_int32), ("kw", c_int32), ("cin", c_int32), ("cout", c_int32)]


class DpmCoeffs(C.Structure):
    _fields_ = [("sqrt_alpha", c_float), ("sqrt_one_minus_alpha", c_float), ("m1", c_float), ("m2", c_float),
                ("m3", c_float), ("m4", c_float), ("m_noise", c_float), ("second_order", c_int32),
                ("prediction_type", c_int32)]


# name -> (restype, argtypes); every symbol include/aether_b200.h declares.
SIGNATURES = {
    "aether_abi_version": (c_int32, []),
    "aether_device_ok": (c_int32, []),
    "aether_gemm_bf16": (C.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                   c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                   c_void_p]),
    "aether_gemm_qkv_norm_rope_bf16": (C.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int32,
                                                 c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_float, c_void_p, c_void_p, c_int32, c_void_p]),
    "aether_attention_bf16": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_int32, c_void_p]),
    "aether_attention_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "aether_attention_bf16_ws": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_int32, c_void_p,
                                           c_int64, c_void_p]),
    "aether_ln_modulate": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                     c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                     c_void_p]),
    "aether_qk_norm_rope": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "aether_small_m_linear": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p]),
    "aether_timestep_sinusoid": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "aether_patchify": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "aether_unpatchify": (C.c_int, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                    c_void_p]),
    "aether_add_pos_embed": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "aether_dit_create": (C.c_int, [C.POINTER(DitConfig), C.POINTER(DitWeights), C.POINTER(c_void_p)]),
    "aether_dit_destroy": (None, [c_void_p]),
    "aether_dit_set_pos_embedding": (C.c_int, [c_void_p, c_void_p]),
    "aether_dit_enable_timing": (C.c_int, [c_void_p, c_int32]),
    "aether_dit_read_timing": (C.c_int, [c_void_p, C.POINTER(c_float), C.POINTER(c_int32)]),
    "aether_dit_workspace_bytes": (c_int64, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "aether_dit_forward": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                     c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_int32, c_void_p]),
    "aether_dit_forward_split": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                           c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                           c_int32, c_void_p, c_int64, c_int32, c_void_p]),
    "aether_conv3d_bf16": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_void_p]),
    "aether_conv3d_bf16_1cta": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          c_int32, c_int32, c_int32, c_void_p]),
    "aether_gn_workspace_floats": (c_int64, [c_int32]),
    "aether_gn_stats": (C.c_int, [c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    "aether_gn_apply": (C.c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                  c_void_p]),
    "aether_upsample_nearest": (C.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          c_int32, c_int32, c_int32, c_void_p]),
    "aether_avgpool_time": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p]),
    "aether_ncthw_to_thwc": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p]),
    "aether_thwc_to_ncthw": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p]),
    "aether_posterior_sample": (C.c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "aether_tile_blend": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                    c_int32, c_void_p]),
    "aether_vae_create": (C.c_int, [C.POINTER(VaeConfig), C.POINTER(VaeParam), c_int32, C.POINTER(c_void_p)]),
    "aether_vae_destroy": (None, [c_void_p]),
    "aether_vae_workspace_bytes": (c_int64, [c_void_p, c_int32, c_int32, c_int32, c_int32]),
    "aether_vae_launch_count": (c_int64, [c_void_p, c_int32, c_int32, c_int32, c_int32]),
    "aether_vae_output_shape": (C.c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, C.POINTER(c_int32)]),
    "aether_vae_encode": (C.c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                    c_void_p, c_int64, c_void_p]),
    "aether_vae_decode": (C.c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                    c_void_p, c_int64, c_void_p]),
    "aether_project_points": (C.c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "aether_resize_bilinear_u8": (C.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "aether_u8_frames_to_model_input": (C.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int32, c_int32,
                                                  c_void_p]),
    "aether_cfg_dpm_step": (C.c_int, [c_void_p, c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                      C.POINTER(DpmCoeffs), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "aether_scale_reduce_work_bytes": (c_int64, []),
    "aether_scale_reduce": (C.c_int, [c_void_p, c_int32, c_int64, c_int64, c_void_p, c_int32, c_int64, c_int64,
                                      c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "aether_blend_crossfade": (C.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int64, c_int64, c_void_p,
                                         c_int32, c_int64, c_int64, c_double, c_void_p, c_int64, c_int64, c_int64,
                                         c_int32, c_void_p]),
    "aether_scale_copy": (C.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int64, c_int64, c_double,
                                    c_void_p, c_int32, c_int64, c_int64, c_int64, c_void_p]),
    "aether_blend_link": (C.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int64, c_int64, c_int64, c_int64,
                                    c_int64, c_int32, c_int64, c_int64, c_void_p, c_void_p]),
    "aether_disparity_to_depth": (C.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int64, c_int64, c_int64,
                                            c_int64, c_int64, c_void_p]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} not found: the aether_b200 CUDA extension is not built. "
            "Run `python -m aether_b200.build` (or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(str(_LIB_PATH), mode=getattr(os, "RTLD_NOW", 2))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


CALLS = [0]     # C-ABI calls that went through check() (bench.py derives its kernel-launch count from it)


def check(status: int, what: str):
    CALLS[0] += 1
    if status != 0:
        raise RuntimeError(f"aether_b200: {what} failed with status {status} ({STATUS.get(status, '?')})")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def _cuda_tensors(obj, found):
    import torch
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            found.append(obj)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _cuda_tensors(o, found)


def device_guard(fn):
    """Run `fn` with the CUDA device of its tensor arguments current.

    The C ABI takes raw device pointers plus a stream handle; `torch.cuda.current_stream()` and the kernels' launch
    context follow the CURRENT device, not the tensors' device.  This decorator makes every entry point safe for
    `module.to("cuda:1")` without `torch.cuda.set_device(1)`: it checks that all CUDA tensor arguments live on one
    device and switches to it for the duration of the call.  A bound method whose object exposes a CUDA `.device`
    (packed module) contributes that device too."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        import torch
        found = []
        _cuda_tensors(args, found)
        _cuda_tensors(tuple(kwargs.values()), found)
        devs = {t.device for t in found}
        if args and not isinstance(args[0], torch.Tensor):
            d = getattr(args[0], "device", None)
            if isinstance(d, torch.device) and d.type == "cuda":
                devs.add(d if d.index is not None else torch.device("cuda", torch.cuda.current_device()))
        if len(devs) > 1:
            raise ValueError(f"{fn.__qualname__}: tensors live on different CUDA devices: {sorted(map(str, devs))}")
        if not devs:
            return fn(*args, **kwargs)
        with torch.cuda.device(next(iter(devs))):
            return fn(*args, **kwargs)
    return wrapper


def require_device():
    lib = load()
    if not lib.aether_device_ok():
        raise RuntimeError("aether_b200 needs an sm_100 (B200) CUDA device; none is visible. No CPU fallback exists.")
    return lib
