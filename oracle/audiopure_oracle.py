"""CPU ORACLE for the AudioPure purification-and-classify hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported, linked or
executed by the product (``diffusion-model-for-audio-defense_b200/``); only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may use it, and only as the checker / reported baseline.

It is an independent functional restatement (plain torch CPU tensor ops on
explicit weight dicts -- no nn.Module, no code copied) of the reference
algorithm; every function cites the reference file:line it follows (paths are
relative to the reference repo root).

PARITY PIN: the reference ships no tests, golden vectors or fixtures for this
path, so the oracle is pinned against *outputs of the
reference itself*, produced in the build container by
``tests/golden/make_golden.py`` (imports the unmodified reference through a
4-line shim) and committed under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks the oracle against them.  The one piece
that cannot be pinned that way is the third-party Euler-Maruyama stepping of
``torchsde==0.2.5`` (requirements.txt:15; not installed, no network): it is
restated from its published algorithm in ``sde_euler_schedule`` and marked
"parity unpinned (torchsde stepping)".

Weights are dicts name -> array with the reference's state-dict keys.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _t(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


# --------------------------------------------------------------------------------------
# a1  calc_diffusion_hyperparams                     DiffWave_Unconditional/util.py:96-123
# --------------------------------------------------------------------------------------
def diffusion_hyperparams(T: int = 200, beta_0: float = 1e-4, beta_T: float = 0.02):
    """float32, sequential in-place cumprod exactly as util.py:111-117 does it."""
    beta = torch.linspace(beta_0, beta_T, T, dtype=torch.float32)
    alpha = 1 - beta
    alpha_bar = alpha.clone()
    beta_tilde = beta.clone()
    for t in range(1, T):
        alpha_bar[t] = alpha_bar[t] * alpha_bar[t - 1]
        beta_tilde[t] = beta_tilde[t] * ((1 - alpha_bar[t - 1]) / (1 - alpha_bar[t]))
    sigma = torch.sqrt(beta_tilde)
    return {"T": T, "Beta": beta, "Alpha": alpha, "Alpha_bar": alpha_bar, "Sigma": sigma}


# --------------------------------------------------------------------------------------
# a2  calc_diffusion_step_embedding                  DiffWave_Unconditional/util.py:68-93
# --------------------------------------------------------------------------------------
def step_embedding(steps: torch.Tensor, dim_in: int = 128) -> torch.Tensor:
    """steps (B,1) -> (B, dim_in) = [sin(t*w_j), cos(t*w_j)], w_j = exp(-j ln(1e4)/(half-1))."""
    half = dim_in // 2
    scale = np.log(10000) / (half - 1)                       # util.py:86 (python float64)
    w = torch.exp(torch.arange(half) * -scale)               # util.py:87 (float32 result)
    arg = steps.to(torch.float32) * w                        # util.py:88
    return torch.cat((torch.sin(arg), torch.cos(arg)), 1)    # util.py:89-90


# --------------------------------------------------------------------------------------
# a5  weight-norm fold                               WaveNet.py:27-28,66-73 (nn.utils.weight_norm, dim=0)
# --------------------------------------------------------------------------------------
def fold_weight_norm(g, v, dtype=torch.float32) -> torch.Tensor:
    """w = g * v / ||v||_2, norm over all dims but 0."""
    g = _t(g, dtype)
    v = _t(v, dtype)
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g.reshape(norm.shape) / norm)


def _swish(x):
    return x * torch.sigmoid(x)                              # WaveNet.py:10-11


# --------------------------------------------------------------------------------------
# a3,a4,a6  WaveNet_Speech_Commands.forward          WaveNet.py:75-97,120-135,164-172
# --------------------------------------------------------------------------------------
def wavenet_forward(sd: dict, audio, steps, num_res_layers: int = 36, dilation_cycle: int = 12,
                    embed_dim_in: int = 128, dtype=torch.float32, return_internals: bool = False):
    """audio (B,1,L), steps (B,1) -> eps (B,1,L).

    Reproduces the in-place alias at WaveNet.py:77,84: ``h = x; h += part_t`` mutates x, so the residual
    output is ((x + part_t) + res) * sqrt(.5).
    """
    w = lambda k: _t(sd[k], dtype)
    wn = lambda p: fold_weight_norm(sd[p + ".weight_g"], sd[p + ".weight_v"], dtype)
    x = _t(audio, dtype)
    steps = _t(steps, dtype)
    # init conv 1x1 + custom ReLU                            WaveNet.py:147,13-19
    h = F.conv1d(x, wn("init_conv.0.conv"), w("init_conv.0.conv.bias"))
    h = torch.maximum(h, torch.zeros_like(h))
    # step embedding                                          WaveNet.py:124-126
    emb = step_embedding(steps, embed_dim_in).to(dtype)
    emb = _swish(F.linear(emb, w("residual_layer.fc_t1.weight"), w("residual_layer.fc_t1.bias")))
    emb = _swish(F.linear(emb, w("residual_layer.fc_t2.weight"), w("residual_layer.fc_t2.bias")))
    skip_total = torch.zeros_like(h)
    internals = {}
    for n in range(num_res_layers):                          # WaveNet.py:131-133
        p = f"residual_layer.residual_blocks.{n}"
        d = 2 ** (n % dilation_cycle)                        # WaveNet.py:117
        C = h.shape[1]
        part_t = F.linear(emb, w(p + ".fc_t.weight"), w(p + ".fc_t.bias")).reshape(-1, C, 1)
        u = h + part_t                                       # WaveNet.py:82-84 (aliasing: x becomes u too)
        a = F.conv1d(u, wn(p + ".dilated_conv_layer.conv"), w(p + ".dilated_conv_layer.conv.bias"),
                     dilation=d, padding=d)                  # WaveNet.py:26,87
        o = torch.tanh(a[:, :C]) * torch.sigmoid(a[:, C:])   # WaveNet.py:90
        res = F.conv1d(o, wn(p + ".res_conv"), w(p + ".res_conv.bias"))
        skip = F.conv1d(o, wn(p + ".skip_conv"), w(p + ".skip_conv.bias"))
        if return_internals and n in (0, 1, num_res_layers - 1):
            internals[f"u{n}"] = u
            internals[f"o{n}"] = o
        h = (u + res) * math.sqrt(0.5)                       # WaveNet.py:97 with x == u
        skip_total = skip_total + skip                       # WaveNet.py:133
    s = skip_total * math.sqrt(1.0 / num_res_layers)         # WaveNet.py:135
    y = F.relu(F.conv1d(s, wn("final_conv.0.conv"), w("final_conv.0.conv.bias")))   # WaveNet.py:160-161
    eps = F.conv1d(y, w("final_conv.2.conv.weight"), w("final_conv.2.conv.bias"))   # WaveNet.py:162
    if return_internals:
        internals["skip_scaled"] = s
        return eps, internals
    return eps


def wavenet_forward_bf16_dataflow(sd: dict, audio, steps, num_res_layers: int = 36, dilation_cycle: int = 12,
                                  embed_dim_in: int = 128):
    """``wavenet_forward`` with the ROUNDING POINTS of the product's bf16 tensor-core mode made explicit (fp32 accumulation
    everywhere): the residual stream u is stored as bf16 after every block, every GEMM operand is bf16 -- u and the folded
    weights of the dilated convolution, the gate output o (A operand of the res 1x1 and of the deferred skip GEMM) with the
    res / skip weights, the scaled skip sum s with the head's 1x1 weights.  Not a reference function: it is the error budget
    of bf16 OPERANDS for this network (DESIGN.md section 3 'Numerics'), used to tell rounding that is inherent to the mode
    from kernel defects -- the CUDA path must agree with THIS function far more closely than either agrees with fp32."""
    q = lambda x: x.to(torch.bfloat16).to(torch.float32)
    w = lambda k: _t(sd[k], torch.float32)
    wn = lambda p: fold_weight_norm(sd[p + ".weight_g"], sd[p + ".weight_v"], torch.float32)
    x = _t(audio, torch.float32)
    steps = _t(steps, torch.float32)
    h = F.conv1d(x, wn("init_conv.0.conv"), w("init_conv.0.conv.bias"))
    h = torch.maximum(h, torch.zeros_like(h))
    emb = step_embedding(steps, embed_dim_in)
    emb = _swish(F.linear(emb, w("residual_layer.fc_t1.weight"), w("residual_layer.fc_t1.bias")))
    emb = _swish(F.linear(emb, w("residual_layer.fc_t2.weight"), w("residual_layer.fc_t2.bias")))
    C = h.shape[1]
    part = lambda n: F.linear(emb, w(f"residual_layer.residual_blocks.{n}.fc_t.weight"),
                              w(f"residual_layer.residual_blocks.{n}.fc_t.bias")).reshape(-1, C, 1)
    u = q(h + part(0))
    skip_total = torch.zeros_like(h)
    for n in range(num_res_layers):
        p = f"residual_layer.residual_blocks.{n}"
        d = 2 ** (n % dilation_cycle)
        a = F.conv1d(u, q(wn(p + ".dilated_conv_layer.conv")), w(p + ".dilated_conv_layer.conv.bias"), dilation=d, padding=d)
        o = q(torch.tanh(a[:, :C]) * torch.sigmoid(a[:, C:]))
        res = F.conv1d(o, q(wn(p + ".res_conv")), w(p + ".res_conv.bias"))
        skip_total = skip_total + F.conv1d(o, q(wn(p + ".skip_conv")), w(p + ".skip_conv.bias"))
        if n + 1 < num_res_layers:
            u = q((u + res) * math.sqrt(0.5) + part(n + 1))
    s = q(skip_total * math.sqrt(1.0 / num_res_layers))
    y = F.relu(F.conv1d(s, q(wn("final_conv.0.conv")), w("final_conv.0.conv.bias")))
    return F.conv1d(y, w("final_conv.2.conv.weight"), w("final_conv.2.conv.bias"))


# --------------------------------------------------------------------------------------
# a7-a10  DiffWave DDPM purifier                     diffusion_models/diffwave_ddpm.py
# --------------------------------------------------------------------------------------
class NoiseSource:
    """Pops pre-generated host noise tensors in call order: z_diffuse, z_{t*-1}, ..., z_1."""

    def __init__(self, tensors):
        self.tensors = [_t(z, torch.float32) for z in tensors]
        self.i = 0

    def __call__(self, shape):
        z = self.tensors[self.i]
        self.i += 1
        assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        return z


def ddpm_diffuse(x0, hp, reverse_timestep: int, noise):
    """diffwave_ddpm.py:49-73: x_t = sqrt(ab[t*-1]) x0 + sqrt(1-ab[t*-1]) z."""
    x0 = _t(x0, torch.float32)
    ab = hp["Alpha_bar"][reverse_timestep - 1]
    z = noise(x0.shape)
    return torch.sqrt(ab) * x0 + torch.sqrt(1 - ab) * z


def ddpm_coefficients(hp, t: int):
    """diffwave_ddpm.py:159-160: (c_eps, 1/sqrt(alpha) as applied, sigma) in float32."""
    alpha, ab, sigma = hp["Alpha"], hp["Alpha_bar"], hp["Sigma"]
    return (1 - alpha[t]) / torch.sqrt(1 - ab[t]), torch.sqrt(alpha[t]), sigma[t]


def ddpm_reverse(sd, x_t, hp, reverse_timestep: int, noise, eps_fn=None, **wn_kw):
    """diffwave_ddpm.py:75-104,143-164: ancestral sampling t = t*-1 .. 0."""
    x = _t(x_t, torch.float32).clone()
    eps_fn = eps_fn or (lambda xx, tt: wavenet_forward(sd, xx, tt * torch.ones(xx.shape[0], 1), **wn_kw))
    for t in range(reverse_timestep - 1, -1, -1):
        eps = eps_fn(x, t).to(torch.float32)
        c_eps, sqrt_alpha, sigma = ddpm_coefficients(hp, t)
        mu = (x - c_eps * eps) / sqrt_alpha
        x = mu + sigma * noise(x.shape) if t > 0 else mu
    return x


def ddpm_forward(sd, x0, hp, reverse_timestep: int, noise, **kw):
    """DiffWave.forward, diffwave_ddpm.py:36-47."""
    return ddpm_reverse(sd, ddpm_diffuse(x0, hp, reverse_timestep, noise), hp, reverse_timestep, noise, **kw)


def predict_x0_from_eps(hp, x_t, t: int, eps):
    """diffwave_ddpm.py:195-205."""
    ab = hp["Alpha_bar"]
    return (1 / ab).sqrt()[t] * _t(x_t, torch.float32) - (1 / ab - 1).sqrt()[t] * _t(eps, torch.float32)


def one_shot_denoise(sd, x_t, hp, reverse_timestep: int, eps_fn=None, **wn_kw):
    """diffwave_ddpm.py:174-182."""
    x_t = _t(x_t, torch.float32)
    t = reverse_timestep - 1
    eps_fn = eps_fn or (lambda xx, tt: wavenet_forward(sd, xx, tt * torch.ones(xx.shape[0], 1), **wn_kw))
    return predict_x0_from_eps(hp, x_t, t, eps_fn(x_t, t))


def two_shot_denoise(sd, x_t, hp, reverse_timestep: int, eps_fn=None, **wn_kw):
    """diffwave_ddpm.py:184-193,207-226."""
    x_t = _t(x_t, torch.float32)
    t = reverse_timestep - 1
    eps_fn = eps_fn or (lambda xx, tt: wavenet_forward(sd, xx, tt * torch.ones(xx.shape[0], 1), **wn_kw))
    alpha, ab, beta = hp["Alpha"], hp["Alpha_bar"], hp["Beta"]
    eps = eps_fn(x_t, t)
    mu = (ab[t] / alpha[0]).sqrt()
    sig = (1 - ab[t] - (ab[t] / alpha[0]) * beta[0] ** 2).sqrt()
    x1 = (x_t - sig * eps) / mu
    c_eps, sqrt_alpha, _ = ddpm_coefficients(hp, 0)
    return (x1 - c_eps * eps_fn(x1, 0)) / sqrt_alpha


def reffwave_forward(sd, x0, hp, reverse_timestep: int, num_re: int, noise, eps_fn=None, **wn_kw):
    """ReffWave.forward, diffwave_ddpm.py:272-283: num_re x (diffuse to t*, one-shot denoise)."""
    x = _t(x0, torch.float32)
    for _ in range(num_re):
        x = one_shot_denoise(sd, ddpm_diffuse(x, hp, reverse_timestep, noise), hp, reverse_timestep, eps_fn=eps_fn, **wn_kw)
    return x


def fast_reverse_tables(hp, reverse_timestep: int, K: int = 3):
    """diffwave_ddpm.py:118-131: respaced K-step schedule (S, Alpha_new, Alpha_bar_new, Beta_tilde_new)."""
    ab = hp["Alpha_bar"]
    S = torch.round(torch.linspace(1, reverse_timestep, K)).int() - 1
    beta_new, beta_tilde_new = torch.zeros(K), torch.zeros(K)
    for i in range(K):
        if i > 0:
            beta_new[i] = 1 - ab[S[i]] / ab[S[i - 1]]
            beta_tilde_new[i] = (1 - ab[S[i - 1]]) / (1 - ab[S[i]]) * beta_new[i]
        else:
            beta_new[i] = 1 - ab[S[i]]
            beta_tilde_new[i] = 0
    alpha_new = 1 - beta_new
    return S, alpha_new, torch.cumprod(alpha_new, dim=0), beta_tilde_new


def fast_reverse(sd, x_t, hp, reverse_timestep: int, noise, eps_fn=None, K: int = 3, **wn_kw):
    """diffwave_ddpm.py:106-141 (note: sigma = Beta_tilde_new[t], not its sqrt, and noise is added at t=0 too)."""
    x = _t(x_t, torch.float32)
    eps_fn = eps_fn or (lambda xx, tt: wavenet_forward(sd, xx, tt * torch.ones(xx.shape[0], 1), **wn_kw))
    S, alpha_new, ab_new, bt_new = fast_reverse_tables(hp, reverse_timestep, K)
    for t in range(K - 1, -1, -1):
        eps = eps_fn(x, int(S[t]))
        mu = (x - (1 - alpha_new[t]) / torch.sqrt(1 - ab_new[t]) * eps) / torch.sqrt(alpha_new[t])
        x = mu + bt_new[t] * noise(x.shape)
    return x


# --------------------------------------------------------------------------------------
# a12-a13  reverse VP-SDE purifier                   diffusion_models/diffwave_sde.py
# --------------------------------------------------------------------------------------
def sde_tables(N: int = 200, beta_min: float = 0.02, beta_max: float = 4.0):
    """RevVPSDE.__init__, diffwave_sde.py:53-60 (RevDiffWave passes beta_min=1e-4*T, beta_max=0.02*T, :155-158)."""
    betas = torch.linspace(beta_min / N, beta_max / N, N)
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    return {"N": N, "beta_0": beta_min, "beta_1": beta_max, "discrete_betas": betas,
            "alphas_cumprod": alphas_cumprod, "sqrt_1m_alphas_cumprod": torch.sqrt(1.0 - alphas_cumprod)}


def sde_euler_schedule(t_star: int, T: int = 200):
    """Solver-time grid of torchsde's fixed-step Euler for ts = linspace(1 - t*/T, 1 - 1e-5, 2), dt = 1/T.

    PARITY UNPINNED (torchsde==0.2.5 is absent): restated from its published fixed-step loop
    ``while curr_t < t1: next_t = min(curr_t + dt, t1); step(curr_t, next_t)`` with float32 tensor time
    (ts is a float32 tensor, diffwave_sde.py:196; dt is a python float).
    Returns a list of (s, ds) float32 pairs: f and g are evaluated at solver time s, step length ds.
    """
    ts = torch.linspace(1 - t_star / T, 1 - 1e-5, 2)          # diffwave_sde.py:194-196 (float32)
    curr, t1 = ts[0], ts[1]
    dt = 1.0 / T
    out = []
    while bool(curr < t1):
        nxt = torch.minimum(curr + dt, t1)
        out.append((curr.clone(), (nxt - curr).clone()))
        curr = nxt
    return out


def sde_step_coefficients(tab, s: torch.Tensor):
    """RevVPSDE.f / .g at solver time s (diffwave_sde.py:69-133).  Returns dict with the discrete index d,
    beta(tau), 1/sqrt(1-ab[d]) and the diffusion coefficient g."""
    N = tab["N"]
    tau = 1 - s                                                # f(), g(): t' = 1 - t       :121,130
    d = int((tau.float() * N).long())                          # _scale_timesteps          :69-71
    beta = tab["beta_0"] + (tau * N - 1) / (N - 1) * (tab["beta_1"] - tab["beta_0"])   # :75
    if d > 0:                                                  # :108-113
        scale = torch.sqrt(1 - tab["alphas_cumprod"][d - 1]) / torch.sqrt(1 - tab["alphas_cumprod"][d])
    else:
        scale = torch.tensor(0.0)
    return {"d": d, "beta": beta, "sqrt_1m_ab": tab["sqrt_1m_alphas_cumprod"][d], "g": scale * torch.sqrt(beta)}


def sde_purify(sd, x0, t_star: int, noise, T: int = 200, eps_fn=None, sample_step: int = 1, noise_level=None,
               grad_through_eps: bool = False, **wn_kw):
    """RevDiffWave.audio_editing_sample (diffwave_sde.py:166-211): ``sample_step`` chained rounds, outputs concatenated on the
    batch dimension (:182,:211); ``noise_level`` = the (rand_t-jittered) diffusion level of :185-190, default t_star -- the
    solver span always uses t_star (:193-196).

    noise order per round: e (diffusion), then one N(0,1) tensor per Euler step (dW = sqrt(ds) * z).
    The network evaluation inside the drift is a constant for autograd, as in the reference (``compute_eps_t`` is
    ``@torch.no_grad()``, diffwave_ddpm.py:166, called at diffwave_sde.py:94); ``grad_through_eps=True`` differentiates it.
    """
    x0 = _t(x0, torch.float32)
    tab = sde_tables(T, 0.0001 * T, 0.02 * T)
    eps_fn = eps_fn or (lambda xx, tt: wavenet_forward(sd, xx, tt * torch.ones(xx.shape[0], 1), **wn_kw))
    a = (1 - tab["discrete_betas"]).cumprod(dim=0)             # :189
    level = t_star if noise_level is None else noise_level
    xs = []
    for _ in range(sample_step):
        e = noise(x0.shape)
        x = x0 * a[level - 1].sqrt() + e * (1.0 - a[level - 1]).sqrt()   # :190
        for s, ds in sde_euler_schedule(t_star, T):
            c = sde_step_coefficients(tab, s)
            if grad_through_eps:
                eps = eps_fn(x, c["d"]).to(torch.float32)
            else:
                with torch.no_grad():
                    eps = eps_fn(x.detach(), c["d"]).to(torch.float32)      # compute_eps_t(x, disc_steps[0])   :94
            drift = -0.5 * c["beta"] * x                            # vpsde_fn                          :80
            score = -eps / c["sqrt_1m_ab"]                          #                                   :98
            rdrift = drift - torch.sqrt(c["beta"]) ** 2 * score     # diffusion[:, None] ** 2 * score   :103
            f = -rdrift                                             #                                   :124
            x = x + f * ds + c["g"] * torch.sqrt(ds) * noise(x.shape)
        x0 = x
        xs.append(x0)
    return torch.cat(xs, dim=0)


# --------------------------------------------------------------------------------------
# a14  mel front end (torchaudio MelSpectrogram + AmplitudeToDB restated)     SURVEY.md Appendix C
# --------------------------------------------------------------------------------------
def _hz_to_mel(f, scale):
    f = np.asarray(f, dtype=np.float64)
    if scale == "htk":
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m, scale):
    m = np.asarray(m, dtype=np.float64)
    if scale == "htk":
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                   norm: str | None, mel_scale: str) -> np.ndarray:
    """Triangular filterbank (n_freqs, n_mels), torchaudio.functional.melscale_fbanks semantics."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale), n_mels + 2)
    f_pts = _mel_to_hz(m_pts, mel_scale)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    if norm == "slaney":
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return fb.astype(np.float32)


MEL_SC09 = dict(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32, norm="slaney",
                pad_mode="constant", mel_scale="slaney")      # certified_robustness_eval.py:85-86
MEL_KWS = dict(sample_rate=16000, n_fft=400, hop_length=200, n_mels=32, norm=None,
               pad_mode="reflect", mel_scale="htk")           # kws_adaptive_attack_eval.py:74-75 (torchaudio defaults)


def mel_db(wave, sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32, norm="slaney",
           pad_mode="constant", mel_scale="slaney", dtype=torch.float64):
    """(B,1,L) -> (B,1,n_mels,1+L//hop): centre-padded STFT (periodic hann, win=n_fft) as an explicit DFT,
    power, mel filterbank, 10*log10(clamp(.,1e-10)) (AmplitudeToDB('power'), ref 1, no top_db)."""
    x = _t(wave, dtype)
    B, _, L = x.shape
    pad = n_fft // 2
    xp = F.pad(x, (pad, pad), mode=pad_mode)[:, 0]                                  # (B, L+n_fft)
    n_frames = 1 + L // hop_length
    frames = xp.unfold(1, n_fft, hop_length)[:, :n_frames]                          # (B, F, n_fft)
    n = torch.arange(n_fft, dtype=torch.float64)
    win = (0.5 - 0.5 * torch.cos(2 * math.pi * n / n_fft))                           # periodic hann
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(n, k) / n_fft
    fw = frames * win.to(dtype)
    re = fw @ torch.cos(ang).to(dtype)
    im = fw @ (-torch.sin(ang)).to(dtype)
    power = re * re + im * im                                                       # (B, F, n_freqs)
    fb = _t(mel_filterbank(n_fft // 2 + 1, 0.0, sample_rate / 2, n_mels, sample_rate, norm, mel_scale), dtype)
    mel = power @ fb                                                                 # (B, F, n_mels)
    db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    return db.transpose(1, 2).unsqueeze(1).to(torch.float32)                        # (B,1,n_mels,F)


# --------------------------------------------------------------------------------------
# a15  CifarResNeXt                                  audio_models/ConvNets_SpeechCommands/models/resnext.py
# --------------------------------------------------------------------------------------
def _bn_eval(sd, p, x, dtype, eps=1e-5):
    w, b = _t(sd[p + ".weight"], dtype), _t(sd[p + ".bias"], dtype)
    m, v = _t(sd[p + ".running_mean"], dtype), _t(sd[p + ".running_var"], dtype)
    shape = (1, -1) + (1,) * (x.dim() - 2)
    return (x - m.reshape(shape)) / torch.sqrt(v.reshape(shape) + eps) * w.reshape(shape) + b.reshape(shape)


def resnext_forward(sd, spec, cardinality=8, depth=29, widen_factor=4, dtype=torch.float32):
    """(B,1,32,32) -> (B,nlabels) logits; resnext.py:56-64 (bottleneck), :134-142 (net)."""
    w = lambda k: _t(sd[k], dtype)
    x = _t(spec, dtype)
    x = F.relu(_bn_eval(sd, "bn_1", F.conv2d(x, w("conv_1_3x3.weight"), padding=1), dtype))
    block_depth = (depth - 2) // 9
    stages = [64, 64 * widen_factor, 128 * widen_factor, 256 * widen_factor]
    for s in range(3):
        for b in range(block_depth):
            p = f"stage_{s + 1}.stage_{s + 1}_bottleneck_{b}"
            stride = 2 if (b == 0 and s > 0) else 1
            y = F.relu(_bn_eval(sd, p + ".bn_reduce", F.conv2d(x, w(p + ".conv_reduce.weight")), dtype))
            y = F.relu(_bn_eval(sd, p + ".bn", F.conv2d(y, w(p + ".conv_conv.weight"), stride=stride, padding=1,
                                                        groups=cardinality), dtype))
            y = _bn_eval(sd, p + ".bn_expand", F.conv2d(y, w(p + ".conv_expand.weight")), dtype)
            if (p + ".shortcut.shortcut_conv.weight") in sd:
                r = _bn_eval(sd, p + ".shortcut.shortcut_bn",
                             F.conv2d(x, w(p + ".shortcut.shortcut_conv.weight"), stride=stride), dtype)
            else:
                r = x
            x = F.relu(r + y)
    x = F.avg_pool2d(x, 8, 1).reshape(-1, stages[3])
    return F.linear(x, w("classifier.weight"), w("classifier.bias"))


# --------------------------------------------------------------------------------------
# a16  ResNet family                                 audio_models/ConvNets_SpeechCommands/models/resnet.py:32-160
# --------------------------------------------------------------------------------------
RESNET_LAYERS = {18: (False, (2, 2, 2, 2)), 34: (False, (3, 4, 6, 3)), 50: (True, (3, 4, 6, 3)), 101: (True, (3, 4, 23, 3)),
                 152: (True, (3, 8, 36, 3))}


def resnet_forward(sd, spec, depth=34, dtype=torch.float32):
    """(B,1,32,32) -> (B,num_classes): conv7x7 s2 + BN + ReLU, maxpool 3x3 s2, 4 stages of Basic/Bottleneck blocks,
    AvgPool2d(1) (identity), flatten, fc   (resnet.py:145-160; blocks :45-62, :80-101)."""
    w = lambda k: _t(sd[k], dtype)
    bottleneck, counts = RESNET_LAYERS[depth]
    x = _t(spec, dtype)
    x = F.relu(_bn_eval(sd, "bn1", F.conv2d(x, w("conv1.weight"), stride=2, padding=3), dtype))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for l, n in enumerate(counts):
        for b in range(n):
            p = f"layer{l + 1}.{b}"
            stride = 2 if (b == 0 and l > 0) else 1
            if bottleneck:
                y = F.relu(_bn_eval(sd, p + ".bn1", F.conv2d(x, w(p + ".conv1.weight")), dtype))
                y = F.relu(_bn_eval(sd, p + ".bn2", F.conv2d(y, w(p + ".conv2.weight"), stride=stride, padding=1), dtype))
                y = _bn_eval(sd, p + ".bn3", F.conv2d(y, w(p + ".conv3.weight")), dtype)
            else:
                y = F.relu(_bn_eval(sd, p + ".bn1", F.conv2d(x, w(p + ".conv1.weight"), stride=stride, padding=1), dtype))
                y = _bn_eval(sd, p + ".bn2", F.conv2d(y, w(p + ".conv2.weight"), padding=1), dtype)
            r = x
            if (p + ".downsample.0.weight") in sd:
                r = _bn_eval(sd, p + ".downsample.1", F.conv2d(x, w(p + ".downsample.0.weight"), stride=stride), dtype)
            x = F.relu(y + r)
    x = x.reshape(x.shape[0], -1)
    return F.linear(x, w("fc.weight"), w("fc.bias"))


# --------------------------------------------------------------------------------------
# a17  M5                                            audio_models/M5/M5Net.py:21-38
# --------------------------------------------------------------------------------------
VGG_CFG = {11: [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
           13: [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
           16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
           19: [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]}


def vgg_forward(sd, spec, depth=19, dtype=torch.float32):
    """VGG with batch norm in eval mode (models/vgg.py:49-53, make_layers :69-82): [conv3x3 + BN + ReLU | maxpool 2x2]*,
    flatten (1x1 map for a 32x32 input), Linear-ReLU-(Dropout)-Linear-ReLU-(Dropout)-Linear."""
    x = _t(spec, dtype)
    i = 0
    for v in VGG_CFG[depth]:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            i += 1
            continue
        x = F.conv2d(x, _t(sd[f"features.{i}.weight"], dtype), _t(sd[f"features.{i}.bias"], dtype), padding=1)
        x = F.relu(_bn_eval(sd, f"features.{i + 1}", x, dtype))
        i += 3
    x = x.reshape(x.shape[0], -1)
    for j in (0, 3, 6):
        x = F.linear(x, _t(sd[f"classifier.{j}.weight"], dtype), _t(sd[f"classifier.{j}.bias"], dtype))
        if j < 6:
            x = F.relu(x)
    return x


def wideresnet_forward(sd, spec, depth=28, widen_factor=10, dtype=torch.float32):
    """WideResNet in eval mode (models/wideresnet.py:30-39 block, :82-90 net).  A block whose width changes replaces x by
    relu(bn1(x)) for both the convolutions and the 1x1 shortcut; an equal-width block keeps x for the identity shortcut.
    BatchNorm goes through F.batch_norm like the reference's nn.BatchNorm2d: with these random weights a single ReLU whose
    pre-activation changes sign under a differently rounded normalisation moves the input gradient by ~1e-3 (measured:
    3 such flips among 6e6 units with the (x - m) / sqrt(v + eps) * w + b form, none with F.batch_norm, vs float64)."""
    w = lambda k: _t(sd[k], dtype)
    bn = lambda p, x: F.batch_norm(x, w(p + ".running_mean"), w(p + ".running_var"), w(p + ".weight"), w(p + ".bias"), False, 0.0, 1e-5)
    x = F.conv2d(_t(spec, dtype), w("conv1.weight"), padding=1)
    for s in range(3):
        for b in range((depth - 4) // 6):
            p = f"block{s + 1}.layer.{b}"
            stride = 2 if (b == 0 and s > 0) else 1
            a = F.relu(bn(p + ".bn1", x))
            h = F.relu(bn(p + ".bn2", F.conv2d(a, w(p + ".conv1.weight"), stride=stride, padding=1)))
            h = F.conv2d(h, w(p + ".conv2.weight"), padding=1)
            short = F.conv2d(a, w(p + ".convShortcut.weight"), stride=stride) if p + ".convShortcut.weight" in sd else x
            x = short + h
    x = F.relu(bn("bn1", x))
    x = F.avg_pool2d(x, 8).reshape(x.shape[0], -1)
    return F.linear(x, w("fc.weight"), w("fc.bias"))


def densenet_forward(sd, spec, depth=100, dtype=torch.float32):
    """DenseNet-BC in eval mode (models/densenet.py:26-36 bottleneck, :65-71 transition, :133-146 net).  F.batch_norm as in
    wideresnet_forward (same sensitivity of the input gradient to single ReLU sign changes)."""
    w = lambda k: _t(sd[k], dtype)
    bn = lambda p, x: F.batch_norm(x, w(p + ".running_mean"), w(p + ".running_var"), w(p + ".weight"), w(p + ".bias"), False, 0.0, 1e-5)
    x = F.conv2d(_t(spec, dtype), w("conv1.weight"), padding=1)
    for s in (1, 2, 3):
        for l in range((depth - 4) // 6):
            p = f"dense{s}.{l}"
            out = F.conv2d(F.relu(bn(p + ".bn1", x)), w(p + ".conv1.weight"))
            out = F.conv2d(F.relu(bn(p + ".bn2", out)), w(p + ".conv2.weight"), padding=1)
            x = torch.cat((x, out), 1)
        if s < 3:
            x = F.avg_pool2d(F.conv2d(F.relu(bn(f"trans{s}.bn1", x)), w(f"trans{s}.conv1.weight")), 2)
    x = F.avg_pool2d(F.relu(bn("bn", x)), 8).reshape(x.shape[0], -1)
    return F.linear(x, w("fc.weight"), w("fc.bias"))


def m5_forward(sd, wave, stride=16, dtype=torch.float32):
    w = lambda k: _t(sd[k], dtype)
    x = _t(wave, dtype)
    for i in range(1, 5):
        x = F.conv1d(x, w(f"conv{i}.weight"), w(f"conv{i}.bias"), stride=stride if i == 1 else 1)
        x = F.max_pool1d(F.relu(_bn_eval(sd, f"bn{i}", x, dtype)), 4)
    x = F.avg_pool1d(x, x.shape[-1]).reshape(x.shape[0], -1)
    return F.log_softmax(F.linear(x, w("fc1.weight"), w("fc1.bias")), dim=1)


# --------------------------------------------------------------------------------------
# a18  KWSModel (RCNN_KWS)                           audio_models/RCNN_KWS/model.py:5-113
# --------------------------------------------------------------------------------------
def _gru_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """x (T,B,I) -> (T,B,H); torch.nn.GRU gate order r,z,n."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = torch.zeros(B, H, dtype=x.dtype)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gi = F.linear(x[t], w_ih, b_ih)
        gh = F.linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, 0)


def kws_forward(sd, spec, kernel_size=(20, 5), stride=(8, 2), gru_num_layers=2, dtype=torch.float32):
    """(B,1,32,W) -> (B,4) log-probs."""
    w = lambda k: _t(sd[k], dtype)
    x = _t(spec, dtype)
    x = x.squeeze(1) if x.dim() == 4 else x                                         # model.py:94
    in_size = x.shape[1]
    x = F.conv1d(x, w("CRNN_model.sepconv.0.weight"), w("CRNN_model.sepconv.0.bias"), stride=stride[1],
                 groups=in_size)                                                     # model.py:7-9
    x = F.conv1d(x, w("CRNN_model.sepconv.1.weight"), w("CRNN_model.sepconv.1.bias"), stride=stride[0],
                 groups=int(in_size / kernel_size[0]))                               # model.py:10-11
    x = x.permute(2, 0, 1)                                                           # (seq,B,H)  model.py:28
    for layer in range(gru_num_layers):
        g = "CRNN_model.gru."
        fwd = _gru_dir(x, w(f"{g}weight_ih_l{layer}"), w(f"{g}weight_hh_l{layer}"),
                       w(f"{g}bias_ih_l{layer}"), w(f"{g}bias_hh_l{layer}"), False)
        bwd = _gru_dir(x, w(f"{g}weight_ih_l{layer}_reverse"), w(f"{g}weight_hh_l{layer}_reverse"),
                       w(f"{g}bias_ih_l{layer}_reverse"), w(f"{g}bias_hh_l{layer}_reverse"), True)
        x = torch.cat((fwd, bwd), dim=2)
    e = F.linear(torch.tanh(F.linear(x, w("attn_layer.Wx_b.weight"), w("attn_layer.Wx_b.bias"))),
                 w("attn_layer.Vt.weight"))[..., 0].transpose(0, 1)                  # (B,seq)  model.py:105-108
    a = F.softmax(e, dim=-1).unsqueeze(1)                                            # model.py:59
    c = torch.bmm(a, x.transpose(0, 1))[:, 0]                                        # model.py:58,60
    return F.log_softmax(F.linear(c, w("apply_attn.U.weight")), dim=-1)             # model.py:61-62


# --------------------------------------------------------------------------------------
# a20  AcousticSystem.forward                        acoustic_system.py:27-51
# --------------------------------------------------------------------------------------
def acoustic_rescale(x):
    x = _t(x, torch.float32)
    if 0.9 * x.max() > 1 and 0.9 * x.min() < -1:             # acoustic_system.py:29-30
        x = x / (2 ** 15)
    return x


# --------------------------------------------------------------------------------------
# a21  RobustCertificate                             robustness_eval/certified_robust.py
# --------------------------------------------------------------------------------------
def compute_t_star(hp, sigma: float) -> int:
    """certified_robust.py:50-51,102-110."""
    alpha_bar_star = 1 / (1 + sigma ** 2)
    return int(torch.abs(hp["Alpha_bar"] - alpha_bar_star).min(0, keepdim=True)[1].item()) + 1


def lower_conf_bound(k: int, n: int, alpha: float = 0.001) -> float:
    """certified_robust.py:113-117: statsmodels proportion_confint(k, n, 2*alpha, 'beta')[0]
    == Clopper-Pearson lower bound beta.ppf(alpha, k, n-k+1) (0 when k == 0)."""
    from scipy.stats import beta
    if k == 0:
        return 0.0
    return float(beta.ppf(alpha, k, n - k + 1))


def certify_from_counts(counts0, counts, n: int, sigma: float, alpha: float = 0.001):
    """certified_robust.py:84-96: (y_pred, radius)."""
    from scipy.stats import norm
    c_a = int(np.argmax(np.asarray(counts0)))
    pa = lower_conf_bound(int(counts[c_a]), n, alpha)
    if pa > 0.5:
        return c_a, float(sigma * norm.ppf(pa))
    return -1, 0.0


def smooth_counts(logits_fn, x, n: int, sigma: float, batch_size: int, noise, num_classes: int, hp):
    """certified_robust.py:33-67 with the diffusion denoiser: x (1,1,L); logits_fn(x_in, t_star) -> (b,K)."""
    x = _t(x, torch.float32)
    batches = [batch_size] * (n // batch_size) + ([n % batch_size] if n % batch_size else [])
    counts = np.zeros(num_classes, dtype=np.int64)
    alpha_bar_star = 1 / (1 + sigma ** 2)
    t_star = compute_t_star(hp, sigma)
    for b in batches:
        x_in = x.repeat(b, 1, 1) + sigma * noise((b,) + tuple(x.shape[1:]))
        x_in = alpha_bar_star ** 0.5 * x_in
        pred = logits_fn(x_in, t_star).max(1)[1]
        counts += np.bincount(pred.numpy(), minlength=num_classes)
    return counts


# --------------------------------------------------------------------------------------
# §8(f)4  spectrogram-domain purifier ("Diffusion-Spec")
#   UNet:  diffusion_models/Improved_Diffusion_Unconditional/improved_diffusion/unet.py:107-276,301-497, nn.py:12-21,93-121
#   SDE :  diffusion_models/improved_diffusion_sde.py:47-226
# --------------------------------------------------------------------------------------
def unet_timestep_embedding(timesteps, dim: int, max_period: float = 10000.0):
    """nn.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def unet_forward(sd: dict, x, timesteps, ops, cfg):
    """UNetModel.forward (unet.py:462-497) as plain torch ops over a weight dict; ``ops, cfg`` = synthetic.unet_structure()
    (the module walk of UNetModel.__init__).  x (B, C, H, W), timesteps (B,) -> (B, out_channels, H, W)."""
    w = lambda k: _t(sd[k], torch.float32)
    silu = lambda v: v * torch.sigmoid(v)                                   # nn.py:12-14
    gn = lambda v, p: F.group_norm(v.float(), 32, w(p + ".weight"), w(p + ".bias"), 1e-5)      # GroupNorm32, nn.py:17-19,100
    heads = cfg["num_heads"]
    emb = unet_timestep_embedding(_t(timesteps, torch.float32), cfg["model_channels"])
    emb = F.linear(silu(F.linear(emb, w("time_embed.0.weight"), w("time_embed.0.bias"))), w("time_embed.2.weight"),
                   w("time_embed.2.bias"))                                  # unet.py:335-339,476

    def res(p, h):                                                          # ResBlock._forward, unet.py:183-197
        y = F.conv2d(silu(gn(h, p + ".in_layers.0")), w(p + ".in_layers.2.weight"), w(p + ".in_layers.2.bias"), padding=1)
        e = F.linear(silu(emb), w(p + ".emb_layers.1.weight"), w(p + ".emb_layers.1.bias"))[..., None, None]
        if cfg["use_scale_shift_norm"]:
            scale, shift = torch.chunk(e, 2, dim=1)
            y = gn(y, p + ".out_layers.0") * (1 + scale) + shift
        else:
            y = gn(y + e, p + ".out_layers.0")
        y = F.conv2d(silu(y), w(p + ".out_layers.3.weight"), w(p + ".out_layers.3.bias"), padding=1)
        if p + ".skip_connection.weight" in sd:
            h = F.conv2d(h, w(p + ".skip_connection.weight"), w(p + ".skip_connection.bias"))
        return h + y

    def attn(p, h):                                                         # AttentionBlock._forward + QKVAttention, :219-253
        b, c, hh, ww = h.shape
        xf = h.reshape(b, c, -1)
        qkv = F.conv1d(gn(xf, p + ".norm"), w(p + ".qkv.weight"), w(p + ".qkv.bias"))
        qkv = qkv.reshape(b * heads, -1, qkv.shape[2])
        ch = qkv.shape[1] // 3
        q, k, v = torch.split(qkv, ch, dim=1)
        scale = 1 / math.sqrt(math.sqrt(ch))
        wgt = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale).float(), dim=-1)
        a = torch.einsum("bts,bcs->bct", wgt, v).reshape(b, -1, xf.shape[2])
        a = F.conv1d(a, w(p + ".proj_out.weight"), w(p + ".proj_out.bias"))
        return (xf + a).reshape(b, c, hh, ww)

    h = _t(x, torch.float32)
    hs = []
    for p, kind, cin, cout in ops:
        if kind == "conv_in":
            h = F.conv2d(h, w(p + ".weight"), w(p + ".bias"), padding=1)
            hs.append(h)
        elif kind == "res":
            h = res(p, h)
        elif kind == "attn":
            h = attn(p, h)
        elif kind == "push":
            hs.append(h)
        elif kind == "pop":
            h = torch.cat([h, hs.pop()], dim=1)                             # unet.py:493
        elif kind == "down":
            h = F.conv2d(h, w(p + ".op.weight"), w(p + ".op.bias"), stride=2, padding=1)      # Downsample, :93-104
        elif kind == "up":
            h = F.conv2d(F.interpolate(h, scale_factor=2, mode="nearest"), w(p + ".conv.weight"), w(p + ".conv.bias"),
                         padding=1)                                         # Upsample, :68-78
        elif kind == "out":
            h = F.conv2d(silu(gn(h, "out.0")), w("out.2.weight"), w("out.2.bias"), padding=1)  # unet.py:438-442
    return h


MEL_UPPER_BOUND, MEL_LOWER_BOUND = 38.22, -100.0                            # sc09_spectrogram_dataset.py:61-62


def spec_sde_tables(N: int = 1000, beta_min: float = 0.1, beta_max: float = 20.0):
    """RevVPSDE.__init__ of improved_diffusion_sde.py:48-76."""
    betas = torch.linspace(beta_min / N, beta_max / N, N)
    return {"N": N, "beta_0": beta_min, "beta_1": beta_max, "discrete_betas": betas}


def spec_sde_schedule(t_star: int, dt: float = 1e-3):
    """The fixed-step Euler grid of ``sdeint_adjoint(..., method='euler')`` over ts = linspace(1 - t*/1000, 1 - 1e-5, 2)
    (improved_diffusion_sde.py:193-203); no dt is passed there, so torchsde's default dt = 1e-3 applies.
    PARITY UNPINNED (torchsde absent): same restated loop as ``sde_euler_schedule``."""
    ts = torch.linspace(1 - t_star * 1.0 / 1000, 1 - 1e-5, 2)
    curr, t1, out = ts[0], ts[1], []
    while bool(curr < t1):
        nxt = torch.minimum(curr + dt, t1)
        out.append((curr.clone(), (nxt - curr).clone()))
        curr = nxt
    return out


def spec_sde_step_coefficients(tab, s: torch.Tensor):
    """RevVPSDE.f / .g at solver time s (improved_diffusion_sde.py:82-139): continuous beta, continuous alpha_bar, no
    discrete scale factor on the diffusion."""
    tau = 1 - s
    beta = tab["beta_0"] + tau * (tab["beta_1"] - tab["beta_0"])                                        # :84
    ab = torch.exp(-0.5 * (tab["beta_1"] - tab["beta_0"]) * tau ** 2 - tab["beta_0"] * tau)             # :72
    return {"d": int((tau.float() * tab["N"]).long()), "beta": beta, "neg_recip": -1.0 / torch.sqrt(1.0 - ab),
            "g": torch.sqrt(beta)}


def spec_sde_purify(sd, img, t_star: int, noise, ops, cfg, eps_fn=None, sample_step: int = 1):
    """RevImprovedDiffusion.image_editing_sample, rand_t = False (improved_diffusion_sde.py:175-219): standardise the dB
    mel-spectrogram to [-1, 1], diffuse to level t*, integrate the reverse VP-SDE with the UNet's eps, map back.  noise order per
    round: e, then one N(0,1) tensor per Euler step.  sample_step > 1: round k + 1 starts from the DE-standardised output of round k
    (:204-205 assign it to x0 without standardising it again); the rounds' outputs are concatenated on the batch axis (:217)."""
    img = _t(img, torch.float32)
    tab = spec_sde_tables()
    eps_fn = eps_fn or (lambda xx, d: unet_forward(sd, xx, d * torch.ones(xx.shape[0]), ops, cfg))
    x0 = 2 * (img - MEL_LOWER_BOUND) / (MEL_UPPER_BOUND - MEL_LOWER_BOUND) - 1                           # melspec_standardize
    a = (1 - tab["discrete_betas"]).cumprod(dim=0)
    xs = []
    for _ in range(sample_step):
        x = x0 * a[t_star - 1].sqrt() + noise(x0.shape) * (1.0 - a[t_star - 1]).sqrt()                   # :190
        for s, ds in spec_sde_schedule(t_star):
            c = spec_sde_step_coefficients(tab, s)
            eps = eps_fn(x, c["d"]).to(torch.float32)                                                    # :105
            drift = -0.5 * c["beta"] * x                                                                 # :85
            score = c["neg_recip"] * eps                                                                 # :110
            rdrift = drift - torch.sqrt(c["beta"]) ** 2 * score                                          # :115
            x = x + (-rdrift) * ds + c["g"] * torch.sqrt(ds) * noise(x.shape)
        x0 = (x + 1) * (MEL_UPPER_BOUND - MEL_LOWER_BOUND) / 2 + MEL_LOWER_BOUND                         # melspec_inv_standardize
        xs.append(x0)
    return torch.cat(xs, dim=0)


# ---------------------------------------------------------------------------------------------- black-box queries (§8f-3)
def query_loss(scores, labels, kind: str = "Entropy", targeted: bool = False, confidence: float = 0.0,
               clip_max: bool = True):
    """Per-query loss.  'Entropy': nn.CrossEntropyLoss(reduction='none'), what resolve_loss returns for task 'SCR'
    (robustness_eval/_utils.py:113-117).  'Margin': the closed-set branch of SEC4SR_MarginLoss
    (_utils.py:66-84: real + confidence - other, other + confidence - real when targeted; clip at 0, :98-99)."""
    s = _t(scores, torch.float32)
    y = torch.as_tensor(np.asarray(labels), dtype=torch.int64)
    if kind == "Entropy":
        lse = torch.logsumexp(s, dim=1)
        return lse - s.gather(1, y[:, None])[:, 0]
    onehot = F.one_hot(y, s.shape[1]).to(s.dtype)
    real = (onehot * s).sum(1)
    other = ((1 - onehot) * s - onehot * 10000).max(1)[0]
    loss = other + confidence - real if targeted else real + confidence - other
    return loss.clamp_min(0) if clip_max else loss


def eot_scores(model_fn, loss_fn, x, y, eot_size: int = 1, eot_batch: int = 1):
    """_EOT.py:18-69 without gradients: average scores / loss over eot_size evaluations of a stochastic model, plus the
    per-query list of argmax decisions.  model_fn: (n,1,L) -> (n,K)."""
    n = x.shape[0]
    nb = eot_size // eot_batch
    scores = loss = None
    decisions = [[] for _ in range(n)]
    for _ in range(nb):
        s = model_fn(x.repeat(eot_batch, 1, 1))
        l = loss_fn(s, y.repeat(eot_batch))
        sm, lm = s.view(eot_batch, n, -1).mean(0), l.view(eot_batch, n).mean(0)
        scores, loss = (sm, lm) if scores is None else (scores + sm, loss + lm)
        d = s.max(1)[1].view(eot_batch, n).numpy()
        for i in range(n):
            decisions[i] += list(d[:, i])
    return scores / nb, loss / nb, decisions


def nes_gradient(model_fn, loss_fn, x, y, samples_per_draw: int, batch: int, sigma: float, noise, eot_size: int = 1,
                 eot_batch: int = 1):
    """_NES.py:14-56.  x (A,1,L); noise(shape) -> (A, batch/2, 1, L) standard normals for each of the
    samples_per_draw // batch draws.  Returns (mean_loss (A,), grad (A,1,L), adver_loss (A,), adver_score (A,K),
    predict (A,)) plus the list of query batches (for the perturbation kernel's parity test)."""
    from collections import Counter
    x = _t(x, torch.float32)
    y = torch.as_tensor(np.asarray(y), dtype=torch.int64)
    A, C, N = x.shape
    nb = samples_per_draw // batch
    eot_nb = eot_size // eot_batch
    queries = []
    grad = mean_loss = adver_loss = adver_score = predict = None
    for i in range(nb):
        z = _t(noise((A, batch // 2, C, N)), torch.float32)
        z = torch.cat((z, -z), 1)                                           # antithetic pairs, :20
        if i == 0:
            z = torch.cat((torch.zeros_like(x).unsqueeze(1), z), 1)         # the un-noised query rides along, :21-22
        q = (z * sigma + x.unsqueeze(1)).view(-1, C, N)                     # :23-24
        queries.append(q)
        R = z.shape[1]
        scores, loss, dec = eot_scores(model_fn, loss_fn, q, y.repeat_interleave(R), eot_size, eot_batch)
        loss = (loss / eot_nb).view(A, R)                                   # second division by the EOT batch count, :35
        scores = (scores / eot_nb).view(A, R, -1)
        if i == 0:
            adver_loss, adver_score = loss[:, 0], scores[:, 0]
            loss, z = loss[:, 1:], z[:, 1:]
            predict = np.array([Counter(d).most_common(1)[0][0] for d in dec]).reshape(A, -1)[:, 0]
            grad = (loss[:, :, None, None] * z).mean(1)
            mean_loss = loss.mean(1)
        else:
            grad = grad + (loss[:, :, None, None] * z).mean(1)
            mean_loss = mean_loss + loss.mean(1)
    return mean_loss / nb, grad / sigma / nb, adver_loss, adver_score, predict, queries
