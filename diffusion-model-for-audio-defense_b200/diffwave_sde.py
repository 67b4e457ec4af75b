"""Drop-in for the reverse-VP-SDE purifier (diffusion_models/diffwave_sde.py: RevVPSDE :34-133, RevDiffWave :136-217).

The reference integrates dx = f dt + g dW with torchsde's fixed-step Euler-Maruyama (``sdeint_adjoint(..., method='euler',
dt=1/T)``, diffwave_sde.py:200-203).  Here the host builds, with the reference's own float32 arithmetic, one coefficient
row per Euler step (discrete index d, beta(tau), sqrt(1-alpha_bar[d]), g, step length) and the fused CUDA update kernel
(``ap_sde_step``) consumes it; the network evaluation is ``DiffWave.compute_eps_t`` as in RevVPSDE.rvpsde_fn (:94).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .diffwave import DiffWave, _check_wave, create_diffwave_model

__all__ = ["RevVPSDE", "RevDiffWave", "euler_schedule"]


def euler_schedule(t_star: int, T: int = 200):
    """[(s, ds)] float32 tensors of the fixed-step Euler loop over ts = linspace(1 - t*/T, 1 - 1e-5, 2), dt = 1/T
    (diffwave_sde.py:193-196): ``while s < t1: s' = min(s + dt, t1)``, time kept as a float32 tensor."""
    ts = torch.linspace(1 - t_star / T, 1 - 1e-5, 2)
    cur, t1 = ts[0], ts[1]
    dt = 1.0 / T
    out = []
    while bool(cur < t1):
        nxt = torch.minimum(cur + dt, t1)
        out.append((cur.clone(), (nxt - cur).clone()))
        cur = nxt
    return out


class RevVPSDE(torch.nn.Module):
    """Coefficient provider with the reference's constructor signature (diffwave_sde.py:35-60)."""

    def __init__(self, model: DiffWave, score_type="ddpm", beta_min=0.02, beta_max=4, N=200, audio_shape=(1, 16000),
                 model_kwargs=None):
        super().__init__()
        self.model = model
        self.score_type = score_type
        self.model_kwargs = model_kwargs
        self.audio_shape = audio_shape
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.noise_type = "diagonal"
        self.sde_type = "ito"

    def _scale_timesteps(self, t):
        assert torch.all(t <= 1) and torch.all(t >= 0), f"t has to be in [0, 1], but get {t} with shape {t.shape}"
        return (t.float() * self.N).long()

    def step_coefficients(self, s: torch.Tensor, ds: torch.Tensor):
        """Discrete index and the coefficient row of one Euler step taken at solver time s (f/g are evaluated at
        tau = 1 - s, diffwave_sde.py:117-133)."""
        if self.score_type != "guided_diffusion":
            raise NotImplementedError(f"Unknown score type in RevVPSDE: {self.score_type}!")
        tau = 1 - s
        d = int(self._scale_timesteps(tau))
        beta = self.beta_0 + (tau * self.N - 1) / (self.N - 1) * (self.beta_1 - self.beta_0)      # :75
        diffusion = torch.sqrt(beta)
        if d > 0:                                                                                  # :108-113
            g = torch.sqrt(1 - self.alphas_cumprod[d - 1]) / torch.sqrt(1 - self.alphas_cumprod[d]) * diffusion
        else:
            g = torch.zeros(())
        coef = _lib.SdeCoef(float(beta), float(diffusion ** 2), float(self.sqrt_1m_alphas_cumprod[d]), float(ds), float(g),
                            float(torch.sqrt(ds)))
        return d, coef


class RevDiffWave(torch.nn.Module):
    """``RevDiffWave(args)``: args.{ddpm_path, ddpm_config, t, score_type, rand_t, t_delta, use_bm, sample_step}
    (diffwave_sde.py:136-217).  ``state_dict`` / ``noise`` / ``seed`` / ``mode`` are extensions for synthetic weights and
    parity tests; ``noise='torch'`` draws e with ``torch.randn_like`` and one Brownian increment per Euler step with
    ``torch.randn`` (also for a step whose diffusion coefficient is 0, as torchsde's Brownian interval does).

    Gradient (an input that requires grad): the reference evaluates the network inside the drift under ``torch.no_grad()``
    (``DiffWave.compute_eps_t`` is decorated, diffwave_ddpm.py:166, and RevVPSDE.rvpsde_fn calls it, diffwave_sde.py:94), so what
    ``sdeint_adjoint`` differentiates is the affine part of the drift with eps held constant.  That is the default here
    (``grad_through_eps=False``); ``grad_through_eps=True`` is the opt-in exact gradient of the computed chain, with the
    network's backward pass on the CUDA kernels.

    ``args.use_bm=True`` raises: the explicit ``torchsde.BrownianInterval`` object of diffwave_sde.py:199-201 is not reproduced
    (with ``use_bm=False`` torchsde builds the same kind of interval itself, so the increments have the same law)."""

    def __init__(self, args, device=None, state_dict=None, noise: str = "philox", seed: int | None = None, mode=None,
                 grad_through_eps: bool = False, diffwave: DiffWave | None = None):
        super().__init__()
        self.args = args
        if getattr(args, "use_bm", False):
            raise NotImplementedError("RevDiffWave: args.use_bm=True (an explicit torchsde.BrownianInterval, "
                                      "diffwave_sde.py:199-201) is not supported; use_bm=False draws Brownian increments of "
                                      "the same law")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
        self.device = torch.device(device)
        self.grad_through_eps = bool(grad_through_eps)
        audio_shape = (1, 16000)
        if diffwave is not None:      # share an existing DiffWave (one network handle / workspace for several purifiers)
            model, noise = diffwave, diffwave.noise
            model.reverse_timestep = args.t
        else:
            model = create_diffwave_model(model_path=getattr(args, "ddpm_path", None), config_path=args.ddpm_config,
                                          reverse_timestep=args.t, state_dict=state_dict, noise=noise, seed=seed, mode=mode,
                                          device=self.device)
        self.T = 200
        self.model = model
        self.rev_vpsde = RevVPSDE(model=model, score_type=args.score_type, beta_min=0.0001 * self.T,
                                  beta_max=0.02 * self.T, N=self.T, audio_shape=audio_shape, model_kwargs=None)
        self.betas = self.rev_vpsde.discrete_betas.float()
        self.noise = noise
        self._lib = _lib.load()

    def audio_editing_sample(self, audio):
        assert isinstance(audio, torch.Tensor)
        assert audio.ndim == 3, audio.ndim
        if audio.requires_grad and torch.is_grad_enabled():
            return self._audio_editing_sample_autograd(audio.to(self.device))
        x0 = _check_wave(audio.to(self.device), "RevDiffWave")
        B, L = x0.shape[0], int(np.prod(x0.shape[1:]))
        dw = self.model
        xs = []
        for _ in range(self.args.sample_step):
            total_noise_levels = self.args.t
            if self.args.rand_t:
                total_noise_levels = self.args.t + np.random.randint(-self.args.t_delta, self.args.t_delta)
            a = (1 - self.betas).cumprod(dim=0)
            sa, sb = float(a[total_noise_levels - 1].sqrt()), float((1.0 - a[total_noise_levels - 1]).sqrt())
            x = torch.empty_like(x0)
            if self.noise == "torch":
                e = torch.randn_like(x0)
                zp, seed, off = e.data_ptr(), 0, 0
            else:
                e, (_, zp, seed, off) = None, dw._noise_args(x0.shape, x0.device)
            with torch.cuda.device(x0.device):   # x = x0 * sqrt(a) + e * sqrt(1 - a)           (:190)
                _lib.check(self._lib.ap_diffuse(x0.data_ptr(), sa, sb, zp, seed, off, x.data_ptr(), B, L, _lib.stream_ptr()),
                           "ap_diffuse")
            eps = torch.empty_like(x)
            for s, ds in euler_schedule(self.args.t, self.T):
                d, coef = self.rev_vpsde.step_coefficients(s, ds)
                dw.model.eps(x, float(d), out=eps)
                if self.noise == "torch":
                    z = torch.randn(x.shape, device=x.device)
                    zp, seed, off = z.data_ptr(), 0, 0
                else:
                    z, zp, seed, off = dw._noise_args(x.shape, x.device)
                with torch.cuda.device(x.device):
                    _lib.check(self._lib.ap_sde_step(x.data_ptr(), eps.data_ptr(), coef, zp, seed, off, B, L,
                                                     _lib.stream_ptr()), "ap_sde_step")
            x0 = x
            xs.append(x0)
        return torch.cat(xs, dim=0)

    def _audio_editing_sample_autograd(self, x0):
        """The same Euler-Maruyama chain for an input that requires grad (the reference differentiates it with
        torchsde.sdeint_adjoint, diffwave_sde.py:200-203): the affine step in torch ops (discretise-then-differentiate), the
        network on the CUDA kernels -- as a constant of the differentiation by default, like the reference's no-grad
        ``compute_eps_t``; with ``grad_through_eps`` through ``_EpsVJP`` (the exact gradient of the computed output)."""
        dw = self.model
        xs = []
        x0 = x0.to(torch.float32)
        for _ in range(self.args.sample_step):
            total_noise_levels = self.args.t
            if self.args.rand_t:
                total_noise_levels = self.args.t + np.random.randint(-self.args.t_delta, self.args.t_delta)
            a = (1 - self.betas).cumprod(dim=0)
            sa, sb = float(a[total_noise_levels - 1].sqrt()), float((1.0 - a[total_noise_levels - 1]).sqrt())
            e = torch.randn_like(x0) if self.noise == "torch" else dw._randn(x0.shape, x0.device)
            x = sa * x0 + sb * e
            sched = list(euler_schedule(self.args.t, self.T))
            for i, (s, ds) in enumerate(sched):
                d, c = self.rev_vpsde.step_coefficients(s, ds)
                if self.grad_through_eps:
                    eps = dw.model.eps(x, float(d), keep_for_backward=(i == len(sched) - 1))
                else:
                    with torch.no_grad():
                        eps = dw.model.eps(x.detach(), float(d))
                f = 0.5 * c.beta * x - c.diff2 * eps / c.sqrt_1mab        # -(drift - g^2 score), score = -eps / sqrt(1 - abar)
                x = x + f * c.dt
                z = torch.randn(x.shape, device=x.device) if self.noise == "torch" else dw._randn(x.shape, x.device)
                if c.g != 0.0:
                    x = x + c.g * c.sqrt_dt * z
            x0 = x
            xs.append(x0)
        return torch.cat(xs, dim=0)

    def forward(self, x):
        return self.audio_editing_sample(x)
