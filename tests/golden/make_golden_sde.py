"""Golden vectors of the reverse-SDE purifier: the reference's OWN ``RevDiffWave.audio_editing_sample`` / ``RevVPSDE.f`` /
``RevVPSDE.g`` (diffusion_models/diffwave_sde.py:34-217) run unmodified on CPU, values AND gradients.

    python tests/golden/make_golden_sde.py        # build container only; writes reference_golden_sde.npz (~1-2 min)

The ONE piece that is not the reference's code is ``torchsde.sdeint_adjoint`` (torchsde==0.2.5, requirements.txt:15, absent and
not installable offline): the stub below restates its fixed-step Euler-Maruyama loop
    while t < t1:  t' = min(t + dt, t1);  y += f(t, y) (t' - t) + g(t, y) dW,  dW ~ N(0, t' - t)
with the time kept as the float32 tensor ``ts`` the reference passes (diffwave_sde.py:196).  PARITY UNPINNED for that loop only;
everything it calls -- drift, diffusion, the discrete-index arithmetic, the rand_t draw, the sample_step chaining, the
no-grad network evaluation inside the drift (diffwave_ddpm.py:166 ``@torch.no_grad()`` on compute_eps_t, called at
diffwave_sde.py:94) -- is the reference's.

Because eps is computed under no_grad, autograd through this chain differentiates the affine drift only.  That is the gradient
the reference's white-box attacks see for the SDE purifier (adaptive_attack_eval.py:130-133), and what ``RevDiffWave`` of
audiopure_b200 reproduces by default (``grad_through_eps=False``).
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden import REF, synthetic, to_torch_sd  # noqa: E402

L = 16000        # RevDiffWave hard-codes audio_shape = (1, 16000)  (diffwave_sde.py:145)


class RandnInjector:
    """torch.randn_like / torch.randn -> host noise popped in call order (e of the diffusion, then one z per Euler step)."""

    def __init__(self, seed):
        self.seed, self.i = seed, 0

    def __enter__(self):
        self._like, self._randn = torch.randn_like, torch.randn

        def nxt(shape):
            z = torch.from_numpy(synthetic.host_noise(tuple(shape), self.seed, self.i))
            self.i += 1
            return z
        torch.randn_like = lambda t, **kw: nxt(t.shape)
        torch.randn = lambda *size, **kw: nxt(size[0] if len(size) == 1 and not isinstance(size[0], int) else size)
        return self

    def __exit__(self, *a):
        torch.randn_like, torch.randn = self._like, self._randn


def sdeint_euler(sde, y0, ts, method="euler", dt=None, bm=None, **kw):
    """Stand-in for torchsde.sdeint_adjoint(sde, y0, ts, method='euler', dt=dt[, bm=bm]): fixed-step Euler-Maruyama for a
    diagonal-noise Ito SDE.  Differentiable torch ops (discretise-then-differentiate)."""
    assert method == "euler" and sde.noise_type == "diagonal" and sde.sde_type == "ito"
    curr, t1, y = ts[0], ts[-1], y0
    while bool(curr < t1):
        nxt = torch.minimum(curr + dt, t1)
        h = nxt - curr
        g = sde.g(curr, y)
        y = y + sde.f(curr, y) * h + g * (torch.sqrt(h) * torch.randn(y.shape))
        curr = nxt
    return torch.stack([y0, y])


def main():
    mg.install_shim()
    sys.modules["torchsde"].sdeint_adjoint = sdeint_euler
    sys.modules["torchsde"].BrownianInterval = lambda **kw: None
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from diffusion_models.diffwave_sde import RevDiffWave

    tmp = tempfile.mkdtemp()
    ckpt = os.path.join(tmp, "synthetic_diffwave.pkl")
    torch.save({"model_state_dict": to_torch_sd(synthetic.wavenet_state_dict(seed=0))}, ckpt)
    cfg = os.path.join(REF, "configs", "config.json")

    def make(t, **kw):
        a = dict(ddpm_path=ckpt, ddpm_config=cfg, t=t, score_type="guided_diffusion", rand_t=False, t_delta=15, use_bm=False,
                 sample_step=1)
        a.update(kw)
        m = RevDiffWave(argparse.Namespace(**a)).eval()
        for p in m.parameters():
            p.requires_grad_(False)
        return m

    out = {}
    x = torch.from_numpy(synthetic.synthetic_waveforms(1, L, seed=21))
    w = torch.from_numpy(synthetic.host_noise((2, 1, L), 78, 0))
    out["w"] = w.numpy()

    # ---- t* = 3, sample_step = 1: value and gradient of <w, purified>
    m = make(3)
    xr = x.clone().requires_grad_(True)
    with RandnInjector(3103) as inj:
        y = m(xr)
        out["t3_noise_draws"] = np.array(inj.i)
    (g,) = torch.autograd.grad((y * w[:1]).sum(), xr)
    out["t3_out"], out["t3_grad"] = y.detach().numpy(), g.numpy()

    # ---- t* = 7 (first t* whose float32 time grid lands on indices [7..1] rather than [6..0])
    m = make(7)
    with torch.no_grad(), RandnInjector(3107):
        out["t7_out"] = m(x.clone()).numpy()

    # ---- sample_step = 2 at t* = 2 (output (2B,1,L): [x^(1); x^(2)], x^(2) purifies x^(1), diffwave_sde.py:182-211)
    m = make(2, sample_step=2)
    xr = x.clone().requires_grad_(True)
    with RandnInjector(3202) as inj:
        y = m(xr)
        out["ss2_noise_draws"] = np.array(inj.i)
    (g,) = torch.autograd.grad((y * w).sum(), xr)
    out["ss2_out"], out["ss2_grad"] = y.detach().numpy(), g.numpy()

    # ---- rand_t (the diffusion level is jittered by np.random.randint(-t_delta, t_delta), the solver span is not, :186-196)
    m = make(4, rand_t=True, t_delta=3)
    np.random.seed(11)
    lvl = 4 + np.random.randint(-3, 3)
    np.random.seed(11)
    with torch.no_grad(), RandnInjector(3304):
        out["randt_out"] = m(x.clone()).numpy()
    out["randt_level"] = np.array(lvl)

    np.savez_compressed(os.path.join(HERE, "reference_golden_sde.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
