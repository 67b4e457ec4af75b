"""Drop-in for the spectrogram-domain purifier ("Diffusion-Spec", diffusion_models/improved_diffusion_sde.py: RevVPSDE :47-139,
RevImprovedDiffusion :142-226; selected at adaptive_attack_eval.py:134-137 with ``defense_type='spec'``).

``RevImprovedDiffusion(args)`` standardises the dB mel-spectrogram to [-1, 1], diffuses it to level ``args.t`` and integrates the
reverse VP-SDE with the UNet's eps (torchsde's fixed-step Euler-Maruyama, default dt = 1e-3: one coefficient row per step, built
on the host with the reference's float32 arithmetic and consumed by the fused ``ap_sde_step`` kernel), then maps back to dB.
The UNet (``UNet``: UNetModel of improved_diffusion/unet.py) runs on the CUDA kernels of csrc/ap_unet.cu.  The reference
differentiates THROUGH this UNet (no ``no_grad`` on the spectrogram path, unlike the waveform purifier's compute_eps_t), so an input
that requires grad takes the autograd route: the affine Euler steps in torch ops, the network through ``_UNetEps`` whose backward is
``ap_unet_eps_vjp`` (the input gradient; the weights are constants of a white-box attack).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .synthetic import DEFAULT_UNET_CONFIG, unet_structure

__all__ = ["UNet", "RevImprovedDiffusion", "melspec_standardize", "melspec_inv_standardize", "spec_euler_schedule"]

MEL_UPPER_BOUND, MEL_LOWER_BOUND = 38.22, -100.0          # sc09_spectrogram_dataset.py:61-62
_KINDS = {"conv_in": 0, "res": 1, "attn": 2, "push": 3, "pop": 4, "down": 5, "up": 6, "out": 7}


def melspec_standardize(x):
    """sc09_spectrogram_dataset.py:65-72"""
    return 2 * (x - MEL_LOWER_BOUND) / (MEL_UPPER_BOUND - MEL_LOWER_BOUND) - 1


def melspec_inv_standardize(x):
    """sc09_spectrogram_dataset.py:74-81"""
    return (x + 1) * (MEL_UPPER_BOUND - MEL_LOWER_BOUND) / 2 + MEL_LOWER_BOUND


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


class _UNetEps(torch.autograd.Function):
    """eps = UNet(x, t) with the input gradient on the CUDA kernels (``ap_unet_eps_vjp``).  Only x is kept; the backward call
    recomputes the forward with its operations recorded."""

    @staticmethod
    def forward(ctx, x, net, t):
        xd = x.detach().to(torch.float32).contiguous()
        ctx.net, ctx.t = net, float(t)
        ctx.save_for_backward(xd)
        return net.eps(xd, float(t))

    @staticmethod
    def backward(ctx, g):
        (xd,) = ctx.saved_tensors
        return ctx.net.eps_vjp(xd, ctx.t, g), None, None


class UNet(torch.nn.Module):
    """``UNetModel`` look-alike: ``model(x (B,1,32,32), timesteps (B,)) -> eps (B,1,32,32)`` (unet.py:462-497)."""

    def __init__(self, state_dict: dict, device=None, **config):
        super().__init__()
        ops, cfg = unet_structure(config or None)
        self.config = cfg
        self._lib = _lib.load()
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        names = ["time_embed.0.weight", "time_embed.0.bias", "time_embed.2.weight", "time_embed.2.bias"]
        for p, kind, cin, cout in ops:
            if kind == "conv_in":
                names += [p + ".weight", p + ".bias"]
            elif kind == "res":
                names += [p + s for s in (".in_layers.0.weight", ".in_layers.0.bias", ".in_layers.2.weight", ".in_layers.2.bias",
                                          ".emb_layers.1.weight", ".emb_layers.1.bias", ".out_layers.0.weight", ".out_layers.0.bias",
                                          ".out_layers.3.weight", ".out_layers.3.bias")]
                if cin != cout:
                    names += [p + ".skip_connection.weight", p + ".skip_connection.bias"]
            elif kind == "attn":
                names += [p + s for s in (".norm.weight", ".norm.bias", ".qkv.weight", ".qkv.bias", ".proj_out.weight", ".proj_out.bias")]
            elif kind == "down":
                names += [p + ".op.weight", p + ".op.bias"]
            elif kind == "up":
                names += [p + ".conv.weight", p + ".conv.bias"]
            elif kind == "out":
                names += ["out.0.weight", "out.0.bias", "out.2.weight", "out.2.bias"]
        missing = [n for n in names if n not in sd]
        if missing:
            raise KeyError(f"UNet: state dict lacks {missing[:4]} ...")
        self._weights = [_np32(sd[n]) for n in names]
        op_arr = np.ascontiguousarray(np.array([[_KINDS[k], ci, co] for _, k, ci, co in ops], dtype=np.int32))
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device_index = device if isinstance(device, int) else (torch.device(device).index or 0)
        c = _lib.UNetCfg(cfg["image_size"], cfg["in_channels"], cfg["model_channels"], cfg["out_channels"], cfg["num_res_blocks"],
                         cfg["num_heads"], int(cfg["use_scale_shift_norm"]))
        self._handle = C.c_void_p()
        _lib.check(self._lib.ap_unet_create(C.byref(self._handle), C.byref(c), op_arr.ctypes.data, len(ops),
                                            _lib.ptr_array(self._weights), len(self._weights), self.device_index), "ap_unet_create")

    def set_mode(self, mode: str) -> "UNet":
        """'tf32' (default: tensor-core convolutions, the precision of the reference's cuDNN path) or 'fp32' (FFMA, parity mode)"""
        _lib.check(self._lib.ap_unet_set_mode(self._handle, {"tf32": _lib.AP_MODE_TF32, "fp32": _lib.AP_MODE_FP32}[mode]),
                   "ap_unet_set_mode")
        return self

    def eps(self, x: torch.Tensor, t: float, out: torch.Tensor | None = None) -> torch.Tensor:
        if not x.is_cuda:
            raise _lib.AudioPureError("UNet: input must be a CUDA tensor (there is no CPU path)")
        if x.requires_grad and torch.is_grad_enabled():
            if out is not None:
                raise _lib.AudioPureError("UNet.eps: out= cannot be combined with an input that requires grad")
            return _UNetEps.apply(x, self, float(t))
        _lib.check_device(x, self.device_index, "UNet")
        x = x.detach().to(torch.float32).contiguous()
        S = self.config["image_size"]
        assert x.ndim == 4 and tuple(x.shape[1:]) == (1, S, S), f"expected (B,1,{S},{S}), got {tuple(x.shape)}"
        if out is None:
            out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_unet_eps(self._handle, x.data_ptr(), float(t), out.data_ptr(), x.shape[0], _lib.stream_ptr()),
                       "ap_unet_eps")
        return out

    def eps_vjp(self, x: torch.Tensor, t: float, g_eps: torch.Tensor, return_eps: bool = False):
        """g_x = (d eps / d x)^T g_eps at (x, t) -- torch.autograd.grad(UNetModel(x, t), x, g_eps) of the reference."""
        if not (x.is_cuda and g_eps.is_cuda):
            raise _lib.AudioPureError("UNet: inputs must be CUDA tensors (there is no CPU path)")
        _lib.check_device(x, self.device_index, "UNet")
        _lib.check_device(g_eps, self.device_index, "UNet")
        x = x.detach().to(torch.float32).contiguous()
        g = g_eps.detach().to(torch.float32).contiguous()
        S = self.config["image_size"]
        assert x.ndim == 4 and tuple(x.shape[1:]) == (1, S, S) and g.shape == x.shape, (tuple(x.shape), tuple(g.shape))
        gx = torch.empty_like(x)
        eps = torch.empty_like(x) if return_eps else None
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_unet_eps_vjp(self._handle, x.data_ptr(), float(t), g.data_ptr(), gx.data_ptr(),
                                                 eps.data_ptr() if return_eps else None, x.shape[0], _lib.stream_ptr()),
                       "ap_unet_eps_vjp")
        return (gx, eps) if return_eps else gx

    def forward(self, x, timesteps, y=None):
        assert y is None, "the spectrogram UNet is unconditional"
        steps = torch.as_tensor(timesteps).reshape(-1).to(torch.float32).cpu()
        uniq = torch.unique(steps)
        if uniq.numel() == 1:
            return self.eps(x, float(uniq[0]))
        parts, order = [], []
        for tv in uniq.tolist():
            idx = torch.nonzero(steps == tv).reshape(-1).to(x.device)
            parts.append(self.eps(x[idx], tv))
            order.append(idx)
        inv = torch.argsort(torch.cat(order))
        return torch.cat(parts, dim=0)[inv]

    def __del__(self):
        try:
            h = self.__dict__.pop("_handle", None)
            if h:
                self._lib.ap_unet_destroy(h)
        except Exception:
            pass


def spec_euler_schedule(t_star: int, dt: float = 1e-3):
    """[(s, ds)] of the fixed-step Euler loop over ts = linspace(1 - t*/1000, 1 - 1e-5, 2) (improved_diffusion_sde.py:193-196);
    sdeint_adjoint is called without dt there, i.e. with torchsde's default 1e-3."""
    ts = torch.linspace(1 - t_star * 1.0 / 1000, 1 - 1e-5, 2)
    cur, t1, out = ts[0], ts[1], []
    while bool(cur < t1):
        nxt = torch.minimum(cur + dt, t1)
        out.append((cur.clone(), (nxt - cur).clone()))
        cur = nxt
    return out


class RevImprovedDiffusion(torch.nn.Module):
    """``RevImprovedDiffusion(args)``: args.{ddpm_path, t, score_type, rand_t, t_delta, use_bm, sample_step}
    (improved_diffusion_sde.py:142-226); ``forward(spec (B,1,32,32) in dB) -> (B * sample_step, 1, 32, 32)``.
    ``state_dict`` / ``noise`` / ``seed`` as for RevDiffWave."""

    def __init__(self, args, config=None, device=None, state_dict=None, noise: str = "philox", seed: int | None = None):
        super().__init__()
        self.args = args
        if getattr(args, "use_bm", False):
            raise NotImplementedError("RevImprovedDiffusion: args.use_bm=True (an explicit torchsde.BrownianInterval) is not supported")
        if getattr(args, "score_type", "guided_diffusion") != "guided_diffusion":
            raise NotImplementedError(f"Unknown score type in RevVPSDE: {args.score_type}!")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
        self.device = torch.device(device)
        if state_dict is None:
            state_dict = torch.load(args.ddpm_path, map_location="cpu", weights_only=False)
        self.model = UNet(state_dict, device=self.device, **dict(DEFAULT_UNET_CONFIG))
        # RevVPSDE.__init__ defaults: beta_min = 0.1, beta_max = 20, N = 1000 (improved_diffusion_sde.py:48-70)
        self.beta_0, self.beta_1, self.N = 0.1, 20.0, 1000
        self.betas = torch.linspace(self.beta_0 / self.N, self.beta_1 / self.N, self.N).float()
        assert noise in ("philox", "torch")
        self.noise = noise
        self.seed = _lib.philox_key("diffwave", seed)
        self._offset = 0
        self._lib = _lib.load()

    def step_coefficients(self, s: torch.Tensor, ds: torch.Tensor):
        """(discrete step passed to the UNet, coefficient row) of one Euler step at solver time s (f / g at tau = 1 - s,
        improved_diffusion_sde.py:82-139): beta(tau) = beta_0 + tau (beta_1 - beta_0); score = eps * (-1 / sqrt(1 - abar(tau))) with the
        CONTINUOUS abar(tau) = exp(-0.5 (beta_1 - beta_0) tau^2 - beta_0 tau); g = sqrt(beta) (no discrete scale factor)."""
        tau = 1 - s
        d = int((tau.float() * self.N).long())
        beta = self.beta_0 + tau * (self.beta_1 - self.beta_0)
        diffusion = torch.sqrt(beta)
        ab = torch.exp(-0.5 * (self.beta_1 - self.beta_0) * tau ** 2 - self.beta_0 * tau)
        neg_recip = -1.0 / torch.sqrt(1.0 - ab)
        # ap_sde_step computes score = -eps / sqrt_1mab: pass sqrt_1mab = -1 / neg_recip so that the product is the reference's
        coef = _lib.SdeCoef(float(beta), float(diffusion ** 2), float(-1.0 / neg_recip), float(ds), float(diffusion), float(torch.sqrt(ds)))
        return d, coef

    def _noise_args(self, shape, device):
        if self.noise == "torch":
            z = torch.randn(tuple(shape), device=device)
            return z, z.data_ptr(), 0, 0
        n = int(np.prod(shape))
        off = self._offset
        self._offset += (n + 3) // 4
        return None, None, self.seed, off

    def image_editing_sample(self, img):
        assert isinstance(img, torch.Tensor)
        assert img.ndim == 4, img.ndim
        if img.requires_grad and torch.is_grad_enabled():
            return self._image_editing_sample_autograd(img.to(self.device))
        img = img.to(self.device).detach().to(torch.float32).contiguous()
        B, n = img.shape[0], int(np.prod(img.shape[1:]))
        x0 = melspec_standardize(img)
        xs = []
        for _ in range(self.args.sample_step):
            level = self.args.t
            if self.args.rand_t:
                level = self.args.t + np.random.randint(-self.args.t_delta, self.args.t_delta)
            a = (1 - self.betas).cumprod(dim=0)
            sa, sb = float(a[level - 1].sqrt()), float((1.0 - a[level - 1]).sqrt())
            x = torch.empty_like(x0)
            if self.noise == "torch":
                e = torch.randn_like(x0)
                zp, seed, off = e.data_ptr(), 0, 0
            else:
                e, zp, seed, off = self._noise_args(x0.shape, x0.device)
            with torch.cuda.device(x0.device):
                _lib.check(self._lib.ap_diffuse(x0.data_ptr(), sa, sb, zp, seed, off, x.data_ptr(), B, n, _lib.stream_ptr()), "ap_diffuse")
            eps = torch.empty_like(x)
            for s, ds in spec_euler_schedule(self.args.t):
                d, coef = self.step_coefficients(s, ds)
                self.model.eps(x, float(d), out=eps)
                z, zp, seed, off = self._noise_args(x.shape, x.device)
                with torch.cuda.device(x.device):
                    _lib.check(self._lib.ap_sde_step(x.data_ptr(), eps.data_ptr(), coef, zp, seed, off, B, n, _lib.stream_ptr()),
                               "ap_sde_step")
            x0 = x
            xs.append(melspec_inv_standardize(x0))
            # the reference feeds the de-standardised spectrogram of round k back into round k + 1 (improved_diffusion_sde.py:204-205)
            x0 = xs[-1]
        return torch.cat(xs, dim=0)

    def _randn(self, shape, device) -> torch.Tensor:
        z, zp, seed, off = self._noise_args(shape, device)
        if z is not None:
            return z
        zero = torch.zeros(tuple(shape), device=device, dtype=torch.float32)
        out = torch.empty_like(zero)
        n = int(np.prod(shape))
        with torch.cuda.device(device):   # 0 * 0 + 1 * z: the Philox counters the fused kernels would consume for this draw
            _lib.check(self._lib.ap_diffuse(zero.data_ptr(), 0.0, 1.0, None, seed, off, out.data_ptr(), shape[0], n // shape[0],
                                            _lib.stream_ptr()), "ap_diffuse")
        return out

    def _image_editing_sample_autograd(self, img):
        """The same chain for an input that requires grad (the reference differentiates it with torchsde.sdeint_adjoint,
        improved_diffusion_sde.py:198-202, THROUGH the UNet): discretise-then-differentiate, the affine step in torch ops and the
        network through ``_UNetEps``.  Same noise draws, in the same order, as the inference route."""
        xs = []
        x0 = melspec_standardize(img.to(torch.float32))
        for _ in range(self.args.sample_step):
            level = self.args.t
            if self.args.rand_t:
                level = self.args.t + np.random.randint(-self.args.t_delta, self.args.t_delta)
            a = (1 - self.betas).cumprod(dim=0)
            sa, sb = float(a[level - 1].sqrt()), float((1.0 - a[level - 1]).sqrt())
            e = torch.randn_like(x0) if self.noise == "torch" else self._randn(x0.shape, x0.device)
            x = sa * x0 + sb * e
            for s, ds in spec_euler_schedule(self.args.t):
                d, c = self.step_coefficients(s, ds)
                eps = self.model.eps(x, float(d))
                x = x + (0.5 * c.beta * x - c.diff2 * eps / c.sqrt_1mab) * c.dt
                z = self._randn(x.shape, x.device)
                if c.g != 0.0:
                    x = x + c.g * c.sqrt_dt * z
            xs.append(melspec_inv_standardize(x))
            x0 = xs[-1]
        return torch.cat(xs, dim=0)

    def forward(self, x):
        return self.image_editing_sample(x)
