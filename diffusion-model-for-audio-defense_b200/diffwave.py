"""Drop-in for the reference's DDPM purifier surface (diffusion_models/diffwave_ddpm.py) on top of the C ABI.

Mirrors, with the same names, argument meaning and error behaviour:
  * ``calc_diffusion_hyperparams``            DiffWave_Unconditional/util.py:96-123
  * ``WaveNet`` (``model((audio, steps))``)   DiffWave_Unconditional/WaveNet.py:138-172
  * ``DiffWave``                              diffwave_ddpm.py:16-249
  * ``create_diffwave_model``                 diffwave_ddpm.py:395-411
All tensors are CUDA fp32 ``(B, 1, L)`` owned by the caller; every kernel is enqueued on torch's current stream.
Gradients wrt the input flow through ``WaveNet.eps`` / ``WaveNet(...)`` / ``DiffWave.forward`` (``_EpsVJP`` -> ``ap_diffwave_eps_vjp``,
the product autograd forms when a white-box attack back-propagates through the purifier, robustness_eval/white_box_attack.py:438) in
the tensor-core modes; the remaining entry points (``compute_eps_t`` -- ``@torch.no_grad`` in the reference -- ``one_shot_denoise``,
the raw update helpers) are inference-only and raise on an input that requires grad instead of silently dropping the gradient.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Union

import numpy as np
import torch

from . import _lib
from ._lib import AP_MODE_BF16, AP_MODE_FP32, AudioPureError

__all__ = ["calc_diffusion_hyperparams", "WaveNet", "DiffWave", "ReffWave", "create_diffwave_model", "wavenet_weight_list"]


def calc_diffusion_hyperparams(T: int, beta_0: float, beta_T: float) -> dict:
    """Linear-beta DDPM tables as CPU float32 tensors, evaluated in the same order as util.py:107-117 so that the
    tables are bit-identical (sequential products, not cumprod)."""
    beta = torch.linspace(beta_0, beta_T, T)
    alpha = 1 - beta
    alpha_bar = alpha.clone()
    beta_tilde = beta.clone()
    for t in range(1, T):
        alpha_bar[t] = alpha_bar[t] * alpha_bar[t - 1]
        beta_tilde[t] = beta_tilde[t] * ((1 - alpha_bar[t - 1]) / (1 - alpha_bar[t]))
    return {"T": T, "Beta": beta, "Alpha": alpha, "Alpha_bar": alpha_bar, "Sigma": torch.sqrt(beta_tilde)}


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def _fold(lib, sd: dict, prefix: str) -> np.ndarray:
    """w = g * v / ||v|| (weight-norm keys of the reference checkpoints); plain ``.weight`` is accepted too."""
    if prefix + ".weight" in sd:
        return _np32(sd[prefix + ".weight"])
    g, v = _np32(sd[prefix + ".weight_g"]), _np32(sd[prefix + ".weight_v"])
    w = np.empty_like(v)
    _lib.check(lib.ap_fold_weight_norm(g.ctypes.data, v.ctypes.data, w.ctypes.data, v.shape[0], int(v[0].size)),
               "ap_fold_weight_norm")
    return w


def wavenet_weight_list(state_dict: dict, cfg: dict) -> list:
    """Reference state dict (408 tensors, WaveNet.py naming) -> the ordered fp32 host arrays ap_diffwave_create takes."""
    lib = _lib.load()
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()}
    out = [_fold(lib, sd, "init_conv.0.conv").reshape(-1), _np32(sd["init_conv.0.conv.bias"])]
    for name in ("fc_t1", "fc_t2"):
        out += [_np32(sd[f"residual_layer.{name}.weight"]), _np32(sd[f"residual_layer.{name}.bias"])]
    for n in range(cfg["num_res_layers"]):
        p = f"residual_layer.residual_blocks.{n}"
        out += [_np32(sd[p + ".fc_t.weight"]), _np32(sd[p + ".fc_t.bias"]),
                _fold(lib, sd, p + ".dilated_conv_layer.conv"), _np32(sd[p + ".dilated_conv_layer.conv.bias"]),
                _fold(lib, sd, p + ".res_conv").reshape(cfg["res_channels"], -1), _np32(sd[p + ".res_conv.bias"]),
                _fold(lib, sd, p + ".skip_conv").reshape(cfg["skip_channels"], -1), _np32(sd[p + ".skip_conv.bias"])]
    out += [_fold(lib, sd, "final_conv.0.conv").reshape(cfg["skip_channels"], -1), _np32(sd["final_conv.0.conv.bias"]),
            _np32(sd["final_conv.2.conv.weight"]).reshape(-1), _np32(sd["final_conv.2.conv.bias"])]
    return [np.ascontiguousarray(a) for a in out]


def _wants_grad(x) -> bool:
    return isinstance(x, torch.Tensor) and x.requires_grad and torch.is_grad_enabled()


class _EpsVJP(torch.autograd.Function):
    """eps_theta(x, t) with a backward pass through the CUDA kernels (``ap_diffwave_eps_vjp``): the reference's WaveNet is
    an ordinary autograd module, and white-box attacks differentiate through it (robustness_eval/white_box_attack.py:438).
    Only the input is kept; the backward call recomputes the forward with the activations it needs."""

    @staticmethod
    def forward(ctx, x, net, t, keep):
        xd = x.detach().to(torch.float32).contiguous()
        ctx.net, ctx.t = net, float(t)
        ctx.save_for_backward(xd)
        # keep: run the forward that keeps the backward's inputs (token 0: nothing kept -- batch too large, or not asked)
        eps, ctx.token = net._eps_save(xd, float(t)) if keep else (net.eps(xd, float(t)), 0)
        return eps

    @staticmethod
    def backward(ctx, g):
        (xd,) = ctx.saved_tensors
        if ctx.token:     # the last saving forward of the handle is the first one the backward pass reaches
            gx = ctx.net._eps_vjp_saved(ctx.token, xd, g)
            if gx is not None:
                return gx, None, None, None
        return ctx.net.eps_vjp(xd, ctx.t, g), None, None, None


def _check_wave(x: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor, got {type(x)}")
    if x.requires_grad and torch.is_grad_enabled():
        raise AudioPureError(f"{what}: this entry point is inference-only; got an input that requires grad (gradients "
                             "flow through WaveNet.eps / WaveNet(...) / DiffWave.forward / RevDiffWave.forward in the "
                             "bf16 and bf16x3 modes; otherwise wrap the call in torch.no_grad() or detach the input)")
    if not x.is_cuda:
        raise AudioPureError(f"{what}: input must be a CUDA tensor (there is no CPU path)")
    return x.detach().to(torch.float32).contiguous()


class WaveNet(torch.nn.Module):
    """``WaveNet_Speech_Commands`` look-alike: ``model((audio (B,1,L), diffusion_steps (B,1))) -> eps (B,1,L)``."""

    def __init__(self, state_dict: dict, device: Union[int, torch.device, None] = None, mode: str | None = None,
                 **wavenet_config):
        super().__init__()
        cfg = dict(in_channels=1, res_channels=256, skip_channels=256, out_channels=1, num_res_layers=36,
                   dilation_cycle=12, diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                   diffusion_step_embed_dim_out=512)
        cfg.update(wavenet_config)
        self.config = cfg
        self._lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device_index = torch.device(device).index if not isinstance(device, int) else device
        if self.device_index is None:
            self.device_index = 0
        c = _lib.WavenetCfg(cfg["in_channels"], cfg["res_channels"], cfg["skip_channels"], cfg["out_channels"],
                            cfg["num_res_layers"], cfg["dilation_cycle"], cfg["diffusion_step_embed_dim_in"],
                            cfg["diffusion_step_embed_dim_mid"], cfg["diffusion_step_embed_dim_out"])
        weights = wavenet_weight_list(state_dict, cfg)
        self._handle = C.c_void_p()
        _lib.check(self._lib.ap_diffwave_create(C.byref(self._handle), C.byref(c), _lib.ptr_array(weights), len(weights),
                                                self.device_index), "ap_diffwave_create")
        if mode is not None:
            self.set_mode(mode)

    # -- arithmetic mode ------------------------------------------------------------------------------------------
    def set_mode(self, mode: str) -> "WaveNet":
        m = {"bf16": AP_MODE_BF16, "fp32": AP_MODE_FP32, "fp16": _lib.AP_MODE_FP16, "bf16x3": _lib.AP_MODE_BF16X3}[mode]
        _lib.check(self._lib.ap_diffwave_set_mode(self._handle, m), "ap_diffwave_set_mode")
        return self

    @property
    def mode(self) -> str:
        return {AP_MODE_BF16: "bf16", AP_MODE_FP32: "fp32", _lib.AP_MODE_FP16: "fp16",
                _lib.AP_MODE_BF16X3: "bf16x3"}[self._lib.ap_diffwave_get_mode(self._handle)]

    def reserve(self, chunk: int, length: int) -> None:
        _lib.check(self._lib.ap_diffwave_reserve(self._handle, int(chunk), int(length)), "ap_diffwave_reserve")

    # -- forward --------------------------------------------------------------------------------------------------
    def eps(self, x: torch.Tensor, t: float, out: torch.Tensor | None = None, keep_for_backward: bool = True) -> torch.Tensor:
        """eps_theta(x, t) with the same diffusion step t for every row.  Differentiable wrt x (bf16 / bf16x3 modes).
        ``keep_for_backward``: for an input that requires grad, keep the backward's inputs during this forward (the handle
        holds ONE such state, so a chain of evaluations passes True only for its last one -- the first to be differentiated)."""
        if out is None and _wants_grad(x):
            if not x.is_cuda:
                raise AudioPureError("WaveNet: input must be a CUDA tensor (there is no CPU path)")
            _lib.check_device(x, self.device_index, "WaveNet")
            return _EpsVJP.apply(x, self, float(t), bool(keep_for_backward))
        x = _check_wave(x, "WaveNet")
        _lib.check_device(x, self.device_index, "WaveNet")
        assert x.ndim == 3 and x.shape[1] == 1, x.shape
        B, _, L = x.shape
        if out is None:
            out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_diffwave_eps(self._handle, x.data_ptr(), float(t), out.data_ptr(), B, L,
                                                 _lib.stream_ptr()), "ap_diffwave_eps")
        return out

    def _eps_save(self, x: torch.Tensor, t: float):
        """(eps, token): forward that keeps the backward's inputs when the batch fits (token != 0), else the plain forward."""
        B, _, L = x.shape
        out = torch.empty_like(x)
        token = C.c_ulonglong(0)
        with torch.cuda.device(x.device):
            rc = self._lib.ap_diffwave_eps_save(self._handle, x.data_ptr(), float(t), out.data_ptr(), B, L, _lib.stream_ptr(),
                                                C.byref(token))
        if rc != 0:
            return self.eps(x, t), 0
        return out, int(token.value)

    def _eps_vjp_saved(self, token: int, x: torch.Tensor, g_eps: torch.Tensor):
        g = g_eps.detach().to(torch.float32).contiguous()
        B, _, L = x.shape
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            rc = self._lib.ap_diffwave_eps_vjp_saved(self._handle, C.c_ulonglong(token), x.data_ptr(), g.data_ptr(), gx.data_ptr(),
                                                     B, L, _lib.stream_ptr())
        return gx if rc == 0 else None

    def eps_vjp(self, x: torch.Tensor, t: float, g_eps: torch.Tensor) -> torch.Tensor:
        """g_x = (d eps_theta(x, t) / d x)^T g_eps (the backward of ``eps``)."""
        x = _check_wave(x, "WaveNet.eps_vjp")
        _lib.check_device(x, self.device_index, "WaveNet.eps_vjp")
        g = g_eps.detach().to(torch.float32).contiguous()
        assert x.ndim == 3 and x.shape[1] == 1 and g.shape == x.shape, (x.shape, g.shape)
        B, _, L = x.shape
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_diffwave_eps_vjp(self._handle, x.data_ptr(), float(t), g.data_ptr(), gx.data_ptr(), None, B, L,
                                                     _lib.stream_ptr()), "ap_diffwave_eps_vjp")
        return gx

    def forward(self, input_data):
        audio, diffusion_steps = input_data
        steps = torch.as_tensor(diffusion_steps).reshape(-1).to(torch.float32).cpu()
        if steps.numel() == 1:
            return self.eps(audio, float(steps[0]))
        assert steps.numel() == audio.shape[0], "diffusion_steps must have one entry per waveform"
        uniq = torch.unique(steps)
        if uniq.numel() == 1:
            return self.eps(audio, float(uniq[0]))
        audio = _check_wave(audio, "WaveNet")
        out = torch.empty_like(audio)
        for tv in uniq.tolist():   # rows are grouped by step value (all reference callers use a batch-constant step)
            idx = torch.nonzero(steps == tv).reshape(-1).to(audio.device)
            out[idx] = self.eps(audio[idx], tv)
        return out

    def debug_layer(self, x: torch.Tensor, t: float, layer: int):
        """(u_{layer+1}, gate_layer) as (B, L, C) fp32 tensors -- test hook."""
        x = _check_wave(x, "WaveNet.debug_layer")
        B, _, L = x.shape
        Cc = self.config["res_channels"]
        u = torch.empty(B, L, Cc, device=x.device, dtype=torch.float32)
        g = torch.empty_like(u)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_diffwave_debug_layer(self._handle, x.data_ptr(), float(t), int(layer), u.data_ptr(),
                                                         g.data_ptr(), B, L, _lib.stream_ptr()), "ap_diffwave_debug_layer")
        return u, g

    def __del__(self):
        try:   # may run during interpreter shutdown, when torch's Module.__setattr__ no longer works
            h = self.__dict__.pop("_handle", None)
            if h:
                self._lib.ap_diffwave_destroy(h)
        except Exception:
            pass


class DiffWave(torch.nn.Module):
    """Reference ``DiffWave`` surface (diffwave_ddpm.py:16-249).

    ``noise='philox'`` (default) draws the Gaussian noise inside the update kernels (counter-based Philox4x32-10; the key is
    derived from ``seed`` by ``_lib.philox_key`` -- ``None`` gives every object, and every rank, its own stream; successive
    draws advance an internal offset).  ``noise='torch'`` draws it exactly as the reference does --
    ``torch.normal(0, 1, size=...)`` on the CPU generator, copied to the device -- which is what the parity tests use.
    """

    def __init__(self, model: WaveNet, diffusion_hyperparams: dict, reverse_timestep: int = 200, grad_enable=True,
                 noise: str = "philox", seed: int | None = None):
        super().__init__()
        self.model = model
        self.diffusion_hyperparams = diffusion_hyperparams
        self.reverse_timestep = reverse_timestep
        self.freeze = False
        self.grad_enable = grad_enable
        assert noise in ("philox", "torch")
        self.noise = noise
        self.seed = _lib.philox_key("diffwave", seed)   # per-consumer Philox key (see _lib.philox_key)
        self._offset = 0
        self._lib = _lib.load()

    # -- helpers --------------------------------------------------------------------------------------------------
    def _tables(self):
        hp = self.diffusion_hyperparams
        T, Alpha, Alpha_bar, Sigma = hp["T"], hp["Alpha"], hp["Alpha_bar"], hp["Sigma"]
        assert len(Alpha) == T
        assert len(Alpha_bar) == T
        assert len(Sigma) == T
        return T, Alpha, Alpha_bar, Sigma

    def _noise_args(self, shape, device):
        """(z tensor or None, device pointer or None, seed, offset)"""
        if self.noise == "torch":
            z = torch.normal(0, 1, size=tuple(shape)).to(device)
            return z, z.data_ptr(), 0, 0
        n = int(np.prod(shape))
        off = self._offset
        self._offset += (n + 3) // 4
        return None, None, self.seed, off

    @staticmethod
    def _as_tensor(x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        return x

    # -- reference API --------------------------------------------------------------------------------------------
    def forward(self, waveforms: Union[torch.Tensor, np.ndarray]):
        waveforms = self._as_tensor(waveforms)
        if _wants_grad(waveforms):
            return self._forward_autograd(waveforms)
        output = self._diffusion(waveforms)
        output = self._reverse(output)
        return output

    def _randn(self, shape, device) -> torch.Tensor:
        """A standard-normal tensor from the same stream the fused update kernels would consume (noise='philox': the
        Philox counters of this draw; noise='torch': torch.normal on the CPU generator like the reference)."""
        z, zp, seed, off = self._noise_args(shape, device)
        if z is not None:
            return z
        zero = torch.zeros(tuple(shape), device=device, dtype=torch.float32)
        out = torch.empty_like(zero)
        n = int(np.prod(shape))
        with torch.cuda.device(device):   # 0 * 0 + 1 * z
            _lib.check(self._lib.ap_diffuse(zero.data_ptr(), 0.0, 1.0, None, seed, off, out.data_ptr(), shape[0], n // shape[0],
                                            _lib.stream_ptr()), "ap_diffuse")
        return out

    def _forward_autograd(self, x_0: torch.Tensor) -> torch.Tensor:
        """``forward`` for an input that requires grad (diffwave_ddpm.py:36-104 is differentiable in the reference).  The
        network and its backward run in the CUDA kernels (``_EpsVJP``); the per-step affine updates are torch ops so that
        autograd chains them.  Same noise order as the inference path: z_diffuse, z_{t*-1}, ..., z_1."""
        _, _, Alpha_bar, _ = self._tables()
        assert x_0.ndim == 3
        if not x_0.is_cuda:
            raise AudioPureError("DiffWave: input must be a CUDA tensor (there is no CPU path)")
        t_star = self.reverse_timestep
        a, b = float(torch.sqrt(Alpha_bar[t_star - 1])), float(torch.sqrt(1 - Alpha_bar[t_star - 1]))
        x = a * x_0.to(torch.float32) + b * self._randn(x_0.shape, x_0.device)
        for t in range(t_star - 1, -1, -1):
            eps = self.model.eps(x, float(t), keep_for_backward=(t == 0))
            c_eps, sqrt_alpha, sigma = self._ddpm_coefficients(t)
            x = (x - c_eps * eps) / sqrt_alpha
            if t > 0:
                x = x + sigma * self._randn(x.shape, x.device)
        return x

    def _diffusion(self, x_0):
        x_0 = self._as_tensor(x_0)
        _, _, Alpha_bar, _ = self._tables()
        assert x_0.ndim == 3
        if not x_0.is_cuda:
            x_0 = x_0.cuda()
        x_0 = _check_wave(x_0, "DiffWave._diffusion")
        t = self.reverse_timestep
        a = float(torch.sqrt(Alpha_bar[t - 1]))
        b = float(torch.sqrt(1 - Alpha_bar[t - 1]))
        x_t = torch.empty_like(x_0)
        z, zp, seed, off = self._noise_args(x_0.shape, x_0.device)
        with torch.cuda.device(x_0.device):
            _lib.check(self._lib.ap_diffuse(x_0.data_ptr(), a, b, zp, seed, off, x_t.data_ptr(), x_0.shape[0],
                                            x_0.shape[1] * x_0.shape[2], _lib.stream_ptr()), "ap_diffuse")
        return x_t

    def _ddpm_coefficients(self, t: int):
        _, Alpha, Alpha_bar, Sigma = self._tables()
        c_eps = float((1 - Alpha[t]) / torch.sqrt(1 - Alpha_bar[t]))
        return c_eps, float(torch.sqrt(Alpha[t])), float(Sigma[t])

    def _reverse(self, x_t):
        x_t = self._as_tensor(x_t)
        self._tables()
        assert x_t.ndim == 3
        x = _check_wave(x_t if x_t.is_cuda else x_t.cuda(), "DiffWave._reverse").clone()
        eps = torch.empty_like(x)
        B, L = x.shape[0], x.shape[1] * x.shape[2]
        for t in range(self.reverse_timestep - 1, -1, -1):
            self.model.eps(x, float(t), out=eps)
            c_eps, sqrt_alpha, sigma = self._ddpm_coefficients(t)
            z, zp, seed, off = (None, None, 0, 0)
            if t > 0:
                z, zp, seed, off = self._noise_args(x.shape, x.device)
            else:
                sigma = 0.0
            with torch.cuda.device(x.device):
                _lib.check(self._lib.ap_ddpm_step(x.data_ptr(), eps.data_ptr(), c_eps, sqrt_alpha, sigma, zp, seed, off, B, L,
                                                  _lib.stream_ptr()), "ap_ddpm_step")
        return x

    def compute_coefficients(self, x_t, t: int):
        """(eps_theta, mu_theta, sigma_theta) of one reverse step (diffwave_ddpm.py:143-164)."""
        x_t = _check_wave(self._as_tensor(x_t), "DiffWave.compute_coefficients")
        _, _, _, Sigma = self._tables()
        eps = self.model.eps(x_t, float(t))
        c_eps, sqrt_alpha, _ = self._ddpm_coefficients(t)
        mu = x_t.clone()
        with torch.cuda.device(x_t.device):
            _lib.check(self._lib.ap_ddpm_step(mu.data_ptr(), eps.data_ptr(), c_eps, sqrt_alpha, 0.0, None, 0, 0, x_t.shape[0],
                                              x_t.shape[1] * x_t.shape[2], _lib.stream_ptr()), "ap_ddpm_step")
        return eps, mu, Sigma[t]

    @torch.no_grad()
    def compute_eps_t(self, x_t, t):
        return self.model.eps(self._as_tensor(x_t), float(t))

    def _predict_x0_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        Alpha_bar = self.diffusion_hyperparams["Alpha_bar"]
        a = float((1 / Alpha_bar).sqrt()[t])
        b = float((1 / Alpha_bar - 1).sqrt()[t])
        x_t = _check_wave(x_t, "DiffWave._predict_x0_from_eps")
        eps = _check_wave(eps, "DiffWave._predict_x0_from_eps")
        out = torch.empty_like(x_t)
        with torch.cuda.device(x_t.device):
            _lib.check(self._lib.ap_predict_x0(x_t.data_ptr(), eps.data_ptr(), a, b, out.data_ptr(), x_t.shape[0],
                                               int(np.prod(x_t.shape[1:])), _lib.stream_ptr()), "ap_predict_x0")
        return out

    def one_shot_denoise(self, x_t):
        x_t = _check_wave(self._as_tensor(x_t), "DiffWave.one_shot_denoise")
        t = self.reverse_timestep - 1
        return self._predict_x0_from_eps(x_t, t, self.model.eps(x_t, float(t)))

    def _predict_x1_from_eps(self, x_t, t, eps):
        hp = self.diffusion_hyperparams
        Alpha, Alpha_bar, Beta = hp["Alpha"], hp["Alpha_bar"], hp["Beta"]
        mu = float((Alpha_bar[t] / Alpha[0]).sqrt())
        sigma = float((1 - Alpha_bar[t] - (Alpha_bar[t] / Alpha[0]) * Beta[0] ** 2).sqrt())
        out = x_t.clone()   # (x_t - sigma * eps) / mu  == ddpm_step with c_eps = sigma, sqrt_alpha = mu, no noise
        with torch.cuda.device(x_t.device):
            _lib.check(self._lib.ap_ddpm_step(out.data_ptr(), eps.data_ptr(), sigma, mu, 0.0, None, 0, 0, x_t.shape[0],
                                              int(np.prod(x_t.shape[1:])), _lib.stream_ptr()), "ap_ddpm_step")
        return out

    def _predict_x0_from_x1(self, x_1):
        _, mu_0, _ = self.compute_coefficients(x_1, 0)
        return mu_0

    def two_shot_denoise(self, x_t):
        x_t = _check_wave(self._as_tensor(x_t), "DiffWave.two_shot_denoise")
        t = self.reverse_timestep - 1
        eps = self.model.eps(x_t, float(t))
        return self._predict_x0_from_x1(self._predict_x1_from_eps(x_t, t, eps))

    def fast_reverse(self, x_t, K: int = 3):
        """Respaced K-step sampler (diffwave_ddpm.py:106-141; sigma = Beta_tilde_new[t], noise also at the last step)."""
        x = _check_wave(self._as_tensor(x_t), "DiffWave.fast_reverse").clone()
        Alpha_bar = self.diffusion_hyperparams["Alpha_bar"]
        S = torch.round(torch.linspace(1, self.reverse_timestep, K)).int() - 1
        beta_new, beta_tilde_new = torch.zeros(K), torch.zeros(K)
        for i in range(K):
            if i > 0:
                beta_new[i] = 1 - Alpha_bar[S[i]] / Alpha_bar[S[i - 1]]
                beta_tilde_new[i] = (1 - Alpha_bar[S[i - 1]]) / (1 - Alpha_bar[S[i]]) * beta_new[i]
            else:
                beta_new[i] = 1 - Alpha_bar[S[i]]
        alpha_new = 1 - beta_new
        alpha_bar_new = torch.cumprod(alpha_new, dim=0)
        eps = torch.empty_like(x)
        B, L = x.shape[0], x.shape[1] * x.shape[2]
        for t in range(K - 1, -1, -1):
            self.model.eps(x, float(S[t]), out=eps)
            c_eps = float((1 - alpha_new[t]) / torch.sqrt(1 - alpha_bar_new[t]))
            sigma = float(beta_tilde_new[t])
            z, zp, seed, off = self._noise_args(x.shape, x.device)   # the reference draws noise even when sigma == 0
            with torch.cuda.device(x.device):
                _lib.check(self._lib.ap_ddpm_step(x.data_ptr(), eps.data_ptr(), c_eps, float(torch.sqrt(alpha_new[t])), sigma,
                                                  zp, seed, off, B, L, _lib.stream_ptr()), "ap_ddpm_step")
        return x

    def purify(self, waveforms: torch.Tensor) -> torch.Tensor:
        """Whole DDPM purifier in one C-ABI call (``forward`` without per-step Python); Philox noise only."""
        x0 = _check_wave(self._as_tensor(waveforms), "DiffWave.purify")
        t_star = self.reverse_timestep
        _, _, Alpha_bar, _ = self._tables()
        coef = np.zeros((t_star + 1, 4), dtype=np.float32)
        coef[0, 0], coef[0, 1] = float(torch.sqrt(Alpha_bar[t_star - 1])), float(torch.sqrt(1 - Alpha_bar[t_star - 1]))
        for i in range(t_star):
            t = t_star - 1 - i
            coef[1 + i, :3] = self._ddpm_coefficients(t)
            coef[1 + i, 3] = t
        B, L = x0.shape[0], x0.shape[1] * x0.shape[2]
        out = torch.empty_like(x0)
        off = self._offset
        self._offset += ((B * L + 3) // 4) * t_star
        with torch.cuda.device(x0.device):
            _lib.check(self._lib.ap_diffwave_purify_ddpm(self.model._handle, x0.data_ptr(), out.data_ptr(), t_star,
                                                         coef.ctypes.data, None, self.seed, off, B, L, _lib.stream_ptr()),
                       "ap_diffwave_purify_ddpm")
        return out


class ReffWave(DiffWave):
    """Repeated diffuse + one-shot denoise (diffwave_ddpm.py:251-348): ``num_re`` rounds of
    ``x <- one_shot_denoise(diffusion(x))`` at the fixed ``reverse_timestep``."""

    def __init__(self, model: WaveNet, diffusion_hyperparams: dict, reverse_timestep: int = 200, num_re: int = 5,
                 noise: str = "philox", seed: int | None = None):
        super().__init__(model, diffusion_hyperparams, reverse_timestep=reverse_timestep, noise=noise, seed=seed)
        self.num_re = num_re

    def diffusion(self, x_0):
        return self._diffusion(x_0)

    def forward(self, waveforms):
        output = self._as_tensor(waveforms)
        for _ in range(self.num_re):
            output = self.diffusion(output)
            output = self.one_shot_denoise(output)
        return output


def create_diffwave_model(model_path, config_path, reverse_timestep=25, state_dict: dict | None = None,
                          noise: str = "philox", seed: int | None = None, mode: str | None = None, device=None) -> DiffWave:
    """diffwave_ddpm.py:395-411: JSON config + ``torch.load(model_path)['model_state_dict']`` -> DiffWave.
    ``state_dict`` may be given directly (synthetic weights; the reference checkpoints are not in its tree)."""
    with open(config_path) as f:
        config = json.loads(f.read())
    wavenet_config = config["wavenet_config"]
    diffusion_hyperparams = calc_diffusion_hyperparams(**config["diffusion_config"])
    if state_dict is None:
        checkpoint = torch.load(model_path, map_location="cpu", weights_only=False)
        state_dict = checkpoint["model_state_dict"]
    net = WaveNet(state_dict, device=device, mode=mode, **wavenet_config)
    return DiffWave(model=net, diffusion_hyperparams=diffusion_hyperparams, reverse_timestep=reverse_timestep,
                    noise=noise, seed=seed)
