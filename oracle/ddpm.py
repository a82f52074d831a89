"""Restatement of the third-party DDPM scheduler the reference calls (TEST INFRASTRUCTURE).

The reference builds ``diffusers.schedulers.scheduling_ddpm.DDPMScheduler`` (diffusers ^0.18.2,
``pyproject.toml:16``) at ``ddiffpg/models/diffusion_mlp.py:167-173`` with
``num_train_timesteps=T, beta_schedule='squaredcos_cap_v2', clip_sample=True,
prediction_type='epsilon'`` (defaults: ``variance_type='fixed_small'``, ``clip_sample_range=1.0``,
``thresholding=False``) and uses ``set_timesteps`` (:225), ``.timesteps`` (:227), ``.step``
(:243-247), ``.config.num_train_timesteps`` (:303) and ``.add_noise`` (:309-310).

diffusers is not installed in this image and is not vendored by the reference, so its published
algorithm (Ho et al. 2020, eq. 7/15; Nichol & Dhariwal cosine schedule) is restated here with the
same operation order and the same fp32 0-dim-tensor scalar arithmetic.  It is pinned against the
reference's in-tree DDPM (``ddiffpg/models/baseline_models.py:97-133``) by
``tests/test_oracle_cpu.py`` and ``oracle/make_golden.py``.

Noise injection: the real scheduler draws ``randn`` inside ``step`` for every t > 0.  Here an
optional ``noise_queue`` (list of tensors, popped front-first) replaces those draws so that the
oracle and the CUDA kernels consume identical pre-drawn noise.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


def squaredcos_cap_v2_betas(num_steps, max_beta=0.999):
    """betas_for_alpha_bar: beta_i = min(1 - abar((i+1)/T)/abar(i/T), max_beta), float64 -> fp32."""
    def abar(u):
        return math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = []
    for i in range(num_steps):
        u1 = i / num_steps
        u2 = (i + 1) / num_steps
        betas.append(min(1 - abar(u2) / abar(u1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


class DDPMSchedulerRestated:
    """Drop-in for the calls the reference makes on ``DDPMScheduler`` (see module docstring)."""

    def __init__(self, num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2",
                 clip_sample=True, prediction_type="epsilon", variance_type="fixed_small",
                 clip_sample_range=1.0, noise_queue=None):
        if beta_schedule != "squaredcos_cap_v2" or prediction_type != "epsilon" \
                or variance_type != "fixed_small":
            raise NotImplementedError("only the configuration the reference uses is restated")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps,
                                      clip_sample=clip_sample,
                                      clip_sample_range=clip_sample_range)
        self.betas = squaredcos_cap_v2_betas(num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self.noise_queue = noise_queue

    def set_timesteps(self, num_inference_steps):
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError("num_inference_steps > num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = n_train // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def _prev(self, t):
        n = self.num_inference_steps or self.config.num_train_timesteps
        return t - self.config.num_train_timesteps // n

    def step(self, model_output, timestep, sample, generator=None):
        t = int(timestep)
        prev_t = self._prev(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        # epsilon prediction -> x0, clipped
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            r = self.config.clip_sample_range
            x0 = x0.clamp(-r, r)
        c_x0 = (a_prev ** 0.5 * cur_beta) / b_t
        c_xt = cur_alpha ** 0.5 * b_prev / b_t
        prev_sample = c_x0 * x0 + c_xt * sample
        if t > 0:
            if self.noise_queue is not None:
                z = self.noise_queue.pop(0).to(model_output.device, model_output.dtype)
            else:
                z = torch.randn(model_output.shape, generator=generator,
                                dtype=model_output.dtype).to(model_output.device)
            var = torch.clamp(b_prev / b_t * cur_beta, min=1e-20)
            prev_sample = prev_sample + (var ** 0.5) * z
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=x0)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        s_a = ac[timesteps] ** 0.5
        s_a = s_a.flatten()
        while s_a.dim() < original_samples.dim():
            s_a = s_a.unsqueeze(-1)
        s_b = (1 - ac[timesteps]) ** 0.5
        s_b = s_b.flatten()
        while s_b.dim() < original_samples.dim():
            s_b = s_b.unsqueeze(-1)
        return s_a * original_samples + s_b * noise


def ddpm_step_constants(T):
    """Per-timestep fp32 constants of ``step``: rows t=0..T-1 of (c_eps, c_inv, c_x0, c_xt, sigma).

    c_eps = sqrt(1-abar_t), c_inv = 1/sqrt(abar_t), c_x0/c_xt the posterior-mean coefficients and
    sigma = sqrt(clamp(var, 1e-20)) (0 at t == 0 where the reference adds no noise).  Computed with
    the same fp32 0-dim tensor arithmetic as ``DDPMSchedulerRestated.step``.
    """
    s = DDPMSchedulerRestated(num_train_timesteps=T)
    out = torch.zeros(T, 5, dtype=torch.float32)
    for t in range(T):
        a_t = s.alphas_cumprod[t]
        a_prev = s.alphas_cumprod[t - 1] if t >= 1 else s.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        out[t, 0] = b_t ** 0.5
        out[t, 1] = 1.0 / (a_t ** 0.5)
        out[t, 2] = (a_prev ** 0.5 * cur_beta) / b_t
        out[t, 3] = cur_alpha ** 0.5 * b_prev / b_t
        out[t, 4] = torch.clamp(b_prev / b_t * cur_beta, min=1e-20) ** 0.5 if t > 0 else 0.0
    return out
