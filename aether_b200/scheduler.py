"""`CogVideoXDPMScheduler`-shaped scheduler whose `step` is the CUDA kernel `aether_cfg_dpm_step`.

Mirrors the surface the reference pipeline touches
(/root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py):
  set_timesteps(n, device=) / .timesteps :227     init_noise_sigma :686      order :821,:919
  scale_model_input(x, t) :835 (identity)
  step(model_output, old_pred_original_sample, timestep, timestep_back, sample, eta=, generator=,
       return_dict=False) -> (prev_sample, pred_original_sample)  :907-915
plus `step_fused`, the one-kernel form of lines :876-916 (fp32 upcast, CFG combine, step, bf16 cast) used
by aether_b200.pipeline.

The schedule (betas -> alphas_cumprod -> SNR shift -> zero-terminal-SNR rescale) and the multipliers are
computed on the host in float64 exactly like diffusers' scheduling_dpm_cogvideox.py; the
Gaussian draws stay in torch (`torch.randn(..., generator=generator)` in the same order as diffusers'
`randn_tensor`) so that seeds reproduce the reference's noise stream draw for draw.

When `diffusers` is importable, `make_dropin_scheduler()` returns an instance of a subclass of the real
`CogVideoXDPMScheduler`, so the unmodified reference pipeline's `isinstance` branch (:902) takes the
DPM signature.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import ops


def _rescale_zero_terminal_snr(ac: np.ndarray) -> np.ndarray:
    s = np.sqrt(ac)
    s0, sT = s[0].copy(), s[-1].copy()
    s = (s - sT) * (s0 / (s0 - sT))
    return s ** 2


class AetherDPMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", snr_shift_scale: float = 1.0,
                 rescale_betas_zero_snr: bool = True, set_alpha_to_one: bool = True,
                 timestep_spacing: str = "trailing", prediction_type: str = "v_prediction", steps_offset: int = 0,
                 clip_sample: bool = False):
        if beta_schedule != "scaled_linear":
            raise NotImplementedError("CogVideoX uses beta_schedule='scaled_linear'")
        if prediction_type not in ("v_prediction", "epsilon"):
            raise NotImplementedError(prediction_type)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule,
                                      snr_shift_scale=snr_shift_scale, rescale_betas_zero_snr=rescale_betas_zero_snr,
                                      set_alpha_to_one=set_alpha_to_one, timestep_spacing=timestep_spacing,
                                      prediction_type=prediction_type, steps_offset=steps_offset,
                                      clip_sample=clip_sample)
        # torch.linspace(sqrt(b0), sqrt(b1), N, dtype=float64) ** 2
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64).numpy() ** 2
        ac = np.cumprod(1.0 - betas, axis=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
        if rescale_betas_zero_snr:
            ac = _rescale_zero_terminal_snr(ac)
        self.alphas_cumprod = torch.from_numpy(ac)
        self._ac = ac
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else float(ac[0])
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    # ------------------------------------------------------------------ schedule
    def set_timesteps(self, num_inference_steps: int, device=None):
        N = self.config.num_train_timesteps
        if num_inference_steps > N:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than {N}")
        self.num_inference_steps = num_inference_steps
        sp = self.config.timestep_spacing
        if sp == "trailing":
            ts = np.round(np.arange(N, 0, -N / num_inference_steps)).astype(np.int64) - 1
        elif sp == "leading":
            ts = (np.arange(0, num_inference_steps) * (N // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif sp == "linspace":
            ts = np.linspace(0, N - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(f"{sp} is not supported")
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, timestep: int, timestep_back: Optional[int]):
        """get_variables + get_mult of scheduling_dpm_cogvideox.py in float64."""
        prev_t = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a = self._ac[timestep]
        a_prev = self._ac[prev_t] if prev_t >= 0 else np.float64(self.final_alpha_cumprod)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            lamb = np.log((a / (1 - a)) ** 0.5)
            lamb_next = np.log((a_prev / (1 - a_prev)) ** 0.5)
            h = lamb_next - lamb
            m1 = ((1 - a_prev) / (1 - a)) ** 0.5 * np.exp(-h)
            m2 = np.expm1(-2 * h) * a_prev ** 0.5
            m3 = m4 = 0.0
            if timestep_back is not None:
                a_back = self._ac[timestep_back]
                r = (lamb - np.log((a_back / (1 - a_back)) ** 0.5)) / h
                m3 = 1 + 1 / (2 * r)
                m4 = 1 / (2 * r)
            mn = (1 - a_prev) ** 0.5 * (1 - np.exp(-2 * h)) ** 0.5
        return dict(sqrt_a=float(a ** 0.5), sqrt_1ma=float((1 - a) ** 0.5), m1=float(m1), m2=float(m2),
                    m3=float(m3), m4=float(m4), m_noise=float(mn), prev_t=int(prev_t))

    def _c_struct(self, timestep, timestep_back, have_old):
        c = self.coefficients(int(timestep), None if timestep_back is None else int(timestep_back))
        second = bool(have_old and c["prev_t"] >= 0)
        co = ops.dpm_coeffs(c["sqrt_a"], c["sqrt_1ma"], c["m1"], c["m2"], c["m3"] if second else 0.0,
                            c["m4"] if second else 0.0, c["m_noise"], second,
                            0 if self.config.prediction_type == "v_prediction" else 1)
        return co, second

    @staticmethod
    def _draw(sample, generator):
        # diffusers.utils.torch_utils.randn_tensor(shape, generator, device=sample.device, dtype=sample.dtype):
        # a CPU generator draws on the CPU and the result is moved to the sample's device
        gdev = generator.device if generator is not None else sample.device
        return torch.randn(sample.shape, generator=generator, device=gdev, dtype=sample.dtype).to(sample.device)

    # ------------------------------------------------------------------ drop-in step (reference :907-915)
    @torch.no_grad()
    def step(self, model_output, old_pred_original_sample, timestep, timestep_back, sample, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = False):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        co, second = self._c_struct(timestep, timestep_back, old_pred_original_sample is not None)
        sample_c = sample.to(torch.bfloat16).contiguous()
        noise1 = self._draw(sample_c, generator)
        noise2 = self._draw(sample_c, generator) if second else None
        mo = model_output.contiguous()
        if mo.dtype not in (torch.float32, torch.bfloat16):
            mo = mo.float()
        old = old_pred_original_sample.float().contiguous() if second else None
        _, prev32, x0 = ops.cfg_dpm_step(mo, sample_c, co, noise1, noise2, old, 1.0, want_prev_f32=True)
        if return_dict:
            return SimpleNamespace(prev_sample=prev32, pred_original_sample=x0)
        return prev32, x0

    # ------------------------------------------------------------------ fused form of reference :876-916
    @torch.no_grad()
    def step_fused(self, model_output_bf16, guidance_scale: float, old_pred_original_sample, timestep, timestep_back,
                   sample_bf16, generator=None):
        """model_output_bf16: [n_cfg, ...] straight from the transformer (uncond first).  Returns
        (latents_bf16, pred_original_sample_f32): `.float()`, CFG, step and `.to(bf16)` in one kernel."""
        co, second = self._c_struct(timestep, timestep_back, old_pred_original_sample is not None)
        noise1 = self._draw(sample_bf16, generator)
        noise2 = self._draw(sample_bf16, generator) if second else None
        prev, _, x0 = ops.cfg_dpm_step(model_output_bf16.contiguous(), sample_bf16.contiguous(), co, noise1, noise2,
                                       old_pred_original_sample if second else None, float(guidance_scale))
        return prev, x0


def make_dropin_scheduler(**kwargs):
    """An AetherDPMScheduler that also IS-A diffusers CogVideoXDPMScheduler (for the unmodified reference
    pipeline's isinstance check, :902).  Requires diffusers to be importable."""
    from diffusers import CogVideoXDPMScheduler  # noqa: WPS433 (optional dependency)

    class AetherDPMSchedulerDropIn(AetherDPMScheduler, CogVideoXDPMScheduler):  # type: ignore[misc]
        def __init__(self, **kw):
            AetherDPMScheduler.__init__(self, **kw)

    return AetherDPMSchedulerDropIn(**kwargs)
