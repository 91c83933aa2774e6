"""Oracle (test infrastructure): `CogVideoXDPMScheduler` restated in torch.

Reference call sites (all in /root/reference/aether/pipelines/aetherv1_pipeline_cogvideox.py):
  set_timesteps  via retrieve_timesteps :227, :780
  scale_model_input :835 (identity)      init_noise_sigma :686      order :821,:919
  step(noise_pred, old_pred_original_sample, t, timesteps[i-1] | None, latents,
       **extra_step_kwargs, return_dict=False) -> (prev_sample, pred_original_sample)  :907-915

The class lives in third-party diffusers (schedulers/scheduling_dpm_cogvideox.py),
absent here: PARITY UNPINNED, restated from the published v0.32 algorithm
 and anchored by known-answer end points (tests/test_oracle_scheduler.py):
first step alpha_t = 0 => mult1 = 0; last step alpha_prev = 1 => prev_sample == x0.

Type-promotion note (matters for bit-level parity): the coefficients are 0-dim
float64 tensors; `coef * sample` with a bf16 `sample` therefore rounds to bf16
before the fp32 subtraction of `coef * model_output` (fp32, reference :876).
The expressions below keep the exact diffusers operand order so torch applies the
same promotions.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch


def rescale_zero_terminal_snr(alphas_cumprod: torch.Tensor) -> torch.Tensor:
    s = alphas_cumprod.sqrt()
    s0 = s[0].clone()
    sT = s[-1].clone()
    s = s - sT
    s = s * (s0 / (s0 - sT))
    return s ** 2


class OracleDPMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 snr_shift_scale=1.0, rescale_betas_zero_snr=True, set_alpha_to_one=True,
                 timestep_spacing="trailing", prediction_type="v_prediction"):
        self.num_train_timesteps = num_train_timesteps
        self.timestep_spacing = timestep_spacing
        self.prediction_type = prediction_type
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
        ac = torch.cumprod(1.0 - betas, dim=0)
        ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
        if rescale_betas_zero_snr:
            ac = rescale_zero_terminal_snr(ac)
        self.alphas_cumprod = ac
        self.final_alpha_cumprod = torch.tensor(1.0, dtype=torch.float64) if set_alpha_to_one else ac[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        N = self.num_train_timesteps
        if self.timestep_spacing == "trailing":
            ts = np.round(np.arange(N, 0, -N / num_inference_steps)).astype(np.int64) - 1
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (N // num_inference_steps)).round()[::-1].copy().astype(np.int64)
        elif self.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    @staticmethod
    def _lamb(a):
        return ((a / (1 - a)) ** 0.5).log()

    def coefficients(self, timestep: int, timestep_back: Optional[int]):
        """(sqrt_a, sqrt_1ma, m1, m2, m3, m4, m_noise, second_order) as python floats (fp64)."""
        prev_t = timestep - self.num_train_timesteps // self.num_inference_steps
        a = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        a_back = self.alphas_cumprod[timestep_back] if timestep_back is not None else None
        lamb, lamb_next = self._lamb(a), self._lamb(a_prev)
        h = lamb_next - lamb
        m1 = ((1 - a_prev) / (1 - a)) ** 0.5 * (-h).exp()
        m2 = (-2 * h).expm1() * a_prev ** 0.5
        m3 = m4 = None
        if a_back is not None:
            r = (lamb - self._lamb(a_back)) / h
            m3 = 1 + 1 / (2 * r)
            m4 = 1 / (2 * r)
        mn = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        # The coefficients are 0-dim fp32 CPU tensors in diffusers.  On CUDA (where the reference runs) a
        # `coef * bf16_tensor` product is computed in fp32 and rounded once; torch's CPU kernel instead rounds a
        # 0-dim *first* operand to bf16 before multiplying.  Python floats take the fp32 path on both devices, so the
        # oracle reproduces the CUDA semantics wherever it runs.
        f = lambda v: None if v is None else float(v)  # noqa: E731
        return dict(sqrt_a=f(a ** 0.5), sqrt_1ma=f((1 - a) ** 0.5), m1=f(m1), m2=f(m2), m3=f(m3), m4=f(m4),
                    m_noise=f(mn), prev_t=prev_t)

    def step(self, model_output, old_pred_original_sample, timestep, timestep_back, sample,
             eta: float = 0.0, generator=None, return_dict: bool = False, noises=None):
        """`noises`: optional list of pre-drawn noise tensors (test hook so the CUDA
        path and the oracle consume identical draws); otherwise drawn here in the
        diffusers order (1 draw, or 2 on second-order steps)."""
        timestep = int(timestep)
        timestep_back = None if timestep_back is None else int(timestep_back)
        c = self.coefficients(timestep, timestep_back)

        def draw(i):
            if noises is not None:
                return noises[i]
            return torch.randn(sample.shape, generator=generator, dtype=sample.dtype, device=sample.device)

        if self.prediction_type == "v_prediction":
            x0 = c["sqrt_a"] * sample - c["sqrt_1ma"] * model_output
        elif self.prediction_type == "epsilon":
            x0 = (sample - c["sqrt_1ma"] * model_output) / c["sqrt_a"]
        else:
            raise ValueError(self.prediction_type)
        noise = draw(0)
        prev = c["m1"] * sample - c["m2"] * x0 + c["m_noise"] * noise
        if old_pred_original_sample is None or c["prev_t"] < 0:
            return prev, x0
        d = c["m3"] * x0 - c["m4"] * old_pred_original_sample
        noise = draw(1)
        prev = c["m1"] * sample - c["m2"] * d + c["m_noise"] * noise
        return prev, x0
