"""Golden vectors of the spectrogram-domain purifier ("Diffusion-Spec"): the reference's UNetModel and RevImprovedDiffusion
(diffusion_models/improved_diffusion_sde.py, Improved_Diffusion_Unconditional/improved_diffusion/unet.py) run unmodified on CPU.

    python tests/golden/make_golden_unet.py        # build container only; writes reference_golden_unet.npz

Shim as in make_golden.py plus a ``librosa`` stub (sc09_spectrogram_dataset.py imports it at module level; no code on this path
uses it).  ``torchsde.sdeint_adjoint`` is the restated fixed-step Euler loop of make_golden_sde.py -- here WITHOUT a dt argument,
so torchsde's default dt = 1e-3 applies (improved_diffusion_sde.py:200-203).  Weights: synthetic.unet_state_dict (the reference's
446 keys, zero-initialised output convolutions re-randomised).
"""
import argparse
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden import synthetic, to_torch_sd  # noqa: E402
from make_golden_sde import RandnInjector, sdeint_euler  # noqa: E402


def main():
    mg.install_shim()
    sys.modules["librosa"] = types.ModuleType("librosa")
    sys.modules["torchsde"].sdeint_adjoint = lambda sde, y0, ts, method="euler", dt=1e-3, bm=None, **kw: sdeint_euler(
        sde, y0, ts, method=method, dt=dt, bm=bm)
    sys.modules["torchsde"].BrownianInterval = lambda **kw: None
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from diffusion_models.improved_diffusion_sde import RevImprovedDiffusion
    from diffusion_models.Improved_Diffusion_Unconditional.improved_diffusion.script_util import (create_model_and_diffusion,
                                                                                                  model_and_diffusion_defaults)
    out = {}
    sd = synthetic.unet_state_dict(seed=0)
    model, _ = create_model_and_diffusion(**model_and_diffusion_defaults())
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(to_torch_sd(sd))
    model.eval()
    # ---- one UNet evaluation: standardised-spectrogram-like input, B = 3, two timestep values
    x = torch.from_numpy(synthetic.host_noise((3, 1, 32, 32), 5100, 0)) * 0.5
    out["unet_x"] = x.numpy()
    with torch.no_grad():
        out["unet_eps_t37"] = model(x, torch.tensor([37, 37, 37])).numpy()
        out["unet_eps_t1"] = model(x[:1], torch.tensor([1])).numpy()

    # ---- RevImprovedDiffusion.image_editing_sample on dB mel-spectrograms (t = 2: two Euler steps + a 1e-5-short tail)
    tmp = tempfile.mkdtemp()
    ckpt = os.path.join(tmp, "synthetic_unet.pt")
    torch.save(to_torch_sd(sd), ckpt)
    args = argparse.Namespace(ddpm_path=ckpt, t=2, score_type="guided_diffusion", rand_t=False, t_delta=15, use_bm=False, sample_step=1)
    rid = RevImprovedDiffusion(args).eval()
    spec = torch.from_numpy(synthetic.host_noise((2, 1, 32, 32), 5200, 0)) * 15.0 - 30.0        # dB-scale values
    out["spec_in"] = spec.numpy()
    with torch.no_grad(), RandnInjector(5300) as inj:
        out["spec_purified_t2"] = rid(spec.clone()).numpy()
        out["spec_noise_draws"] = np.array(inj.i)
    # ---- sample_step = 2: the de-standardised output of round 1 is fed to round 2 WITHOUT standardising it again (:182,:204-205)
    args2 = argparse.Namespace(ddpm_path=ckpt, t=1, score_type="guided_diffusion", rand_t=False, t_delta=15, use_bm=False, sample_step=2)
    rid2 = RevImprovedDiffusion(args2).eval()
    with torch.no_grad(), RandnInjector(5310) as inj:
        out["spec_purified_t1_s2"] = rid2(spec.clone()).numpy()
        out["spec_noise_draws_s2"] = np.array(inj.i)
    # ---- gradients (the reference back-propagates through this UNet: no no_grad on the spectrogram path)
    for prm in model.parameters():
        prm.requires_grad_(False)
    g_eps = torch.from_numpy(synthetic.host_noise((3, 1, 32, 32), 5400, 0))
    out["unet_g_eps"] = g_eps.numpy()
    xr = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(model(xr, torch.tensor([37, 37, 37])), xr, g_eps)
    out["unet_vjp_t37"] = gx.numpy()
    for prm in rid.parameters():
        prm.requires_grad_(False)
    w = torch.from_numpy(synthetic.host_noise((2, 1, 32, 32), 5500, 0))
    out["spec_grad_w"] = w.numpy()
    sr = spec.clone().requires_grad_(True)
    with RandnInjector(5300):
        y = rid(sr)
    (gs,) = torch.autograd.grad((y * w).sum(), sr)
    out["spec_purified_grad_t2"] = gs.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_golden_unet.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
