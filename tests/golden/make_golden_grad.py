"""Golden GRADIENTS from the unmodified reference (autograd through WaveNet_Speech_Commands / DiffWave.forward), for the
backward pass of the CUDA path (ap_diffwave_eps_vjp).  Same shim and synthetic weights as make_golden.py.

    python tests/golden/make_golden_grad.py        # in the build container; writes reference_golden_grad.npz
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden import REF, NoiseInjector, synthetic, to_torch_sd  # noqa: E402


def main():
    mg.install_shim()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from diffusion_models.diffwave_ddpm import DiffWave
    from diffusion_models.DiffWave_Unconditional.WaveNet import WaveNet_Speech_Commands
    from diffusion_models.DiffWave_Unconditional.util import calc_diffusion_hyperparams

    cfg = json.load(open(os.path.join(REF, "configs", "config.json")))
    hp = calc_diffusion_hyperparams(**cfg["diffusion_config"])
    net = WaveNet_Speech_Commands(**cfg["wavenet_config"]).eval()
    net.load_state_dict(to_torch_sd(synthetic.wavenet_state_dict(seed=0)))
    for p in net.parameters():
        p.requires_grad_(False)
    out = {}

    # ---- VJP of the network: g_x = (d eps / d x)^T g_eps at t = 7 and t = 65
    x = torch.from_numpy(synthetic.synthetic_waveforms(2, 1024, seed=1234))
    g_eps = torch.from_numpy(synthetic.host_noise((2, 1, 1024), 4242, 0))
    out["vjp_g_eps"] = g_eps.numpy()
    for t in (7.0, 65.0):
        xr = x.clone().requires_grad_(True)
        eps = net((xr, t * torch.ones(2, 1)))
        (gx,) = torch.autograd.grad(eps, xr, g_eps)
        out[f"vjp_gx_L1024_t{int(t)}"] = gx.numpy()
    # ragged length: zero padding at both ends with every dilation
    x3 = torch.from_numpy(synthetic.synthetic_waveforms(1, 3001, seed=77))
    g3 = torch.from_numpy(synthetic.host_noise((1, 1, 3001), 4243, 0))
    out["vjp_g_eps_L3001"] = g3.numpy()
    xr = x3.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(net((xr, 7.0 * torch.ones(1, 1))), xr, g3)
    out["vjp_gx_L3001_t7"] = gx.numpy()

    # ---- gradient through the whole purifier (DiffWave.forward, t* = 2, injected noise): d <w, purified> / d x
    dw = DiffWave(model=net, diffusion_hyperparams=hp, reverse_timestep=2).eval()
    w = torch.from_numpy(synthetic.host_noise((2, 1, 1024), 4244, 0))
    out["ddpm_grad_w"] = w.numpy()
    xr = x.clone().requires_grad_(True)
    with NoiseInjector(2024) as inj:
        y = dw(xr)
        assert inj.i == 2
    (gx,) = torch.autograd.grad((y * w).sum(), xr)
    out["ddpm_t2_purified"] = y.detach().numpy()
    out["ddpm_t2_grad_L1024"] = gx.numpy()
    # ---- mel front ends (torchaudio autograd): g_wav for a seeded g_spec
    import torchaudio
    xm = torch.from_numpy(synthetic.synthetic_waveforms(2, 16000, seed=99))
    mel_sc = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32,
                                                  norm="slaney", pad_mode="constant", mel_scale="slaney")
    mel_kws = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_mels=32)
    todb = torchaudio.transforms.AmplitudeToDB(stype="power")
    for name, mel in (("sc09", mel_sc), ("kws", mel_kws)):
        xr = xm.clone().requires_grad_(True)
        spec = todb(mel(xr))
        g = torch.from_numpy(synthetic.host_noise(tuple(spec.shape), 4250, 0))
        (gx,) = torch.autograd.grad(spec, xr, g)
        out[f"mel_{name}_g_spec"] = g.numpy()
        out[f"mel_{name}_grad"] = gx.numpy()

    # ---- ResNeXt-29 8x64 (eval mode): g_spec for a seeded g_logits, on the reference's own mel features
    from models.resnext import CifarResNeXt
    rx = CifarResNeXt(nlabels=10, in_channels=1).eval()
    rx.load_state_dict(to_torch_sd(synthetic.resnext_state_dict(seed=0)))
    for p in rx.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        spec0 = todb(mel_sc(xm))
    sr = spec0.clone().requires_grad_(True)
    gl = torch.from_numpy(synthetic.host_noise((2, 10), 4251, 0))
    (gs,) = torch.autograd.grad(rx(sr), sr, gl)
    out["resnext_in_spec"] = spec0.numpy()
    out["resnext_g_logits"] = gl.numpy()
    out["resnext_grad"] = gs.numpy()

    # ---- ResNet family (BasicBlock and Bottleneck variants, eval mode) on the same mel features
    from models.resnet import resnet34, resnet50
    for depth, ctor in ((34, resnet34), (50, resnet50)):
        rn = ctor(num_classes=10, in_channels=1).eval()
        rn.load_state_dict(to_torch_sd(synthetic.resnet_state_dict(depth=depth, seed=0)))
        sr = spec0.clone().requires_grad_(True)
        (gr,) = torch.autograd.grad(rn(sr), sr, gl)
        out[f"resnet{depth}_grad"] = gr.numpy()

    # ---- M5 raw-waveform classifier (eval mode): g_wav for a seeded g_logp
    sys.path.insert(0, os.path.join(REF, "audio_models", "M5"))
    from M5Net import M5
    m5 = M5(n_input=1, first_kernel_size=160, n_output=10).eval()
    m5.load_state_dict(to_torch_sd(synthetic.m5_state_dict(seed=0)))
    xr = xm.clone().requires_grad_(True)
    gl5 = torch.from_numpy(synthetic.host_noise((2, 10), 4252, 0))
    (g5,) = torch.autograd.grad(m5(xr), xr, gl5)
    out["m5_g_logp"] = gl5.numpy()
    out["m5_grad"] = g5.numpy()

    # ---- RCNN_KWS (eval mode) on the reference's KWS mel features: g_spec for a seeded g_logp
    import importlib.util
    kspec = importlib.util.spec_from_file_location("rcnn_kws_model", os.path.join(REF, "audio_models", "RCNN_KWS", "model.py"))
    kmod = importlib.util.module_from_spec(kspec)
    kspec.loader.exec_module(kmod)
    kws = kmod.KWSModel(in_size=32).eval()
    kws.load_state_dict(to_torch_sd(synthetic.kws_state_dict(seed=0)))
    with torch.no_grad():
        spec_k = todb(mel_kws(xm))
    skr = spec_k.clone().requires_grad_(True)
    glk = torch.from_numpy(synthetic.host_noise((2, 4), 4253, 0))
    (gk,) = torch.autograd.grad(kws(skr), skr, glk)
    out["kws_in_spec"] = spec_k.numpy()
    out["kws_g_logp"] = glk.numpy()
    out["kws_grad"] = gk.numpy()

    # ---- end to end: d CrossEntropy(AcousticSystem(x), y) / d x through DDPM t*=2 -> mel -> ResNeXt, one 1 s clip
    from acoustic_system import AcousticSystem
    transform = lambda w: todb(mel_sc(w))
    dw.reverse_timestep = 2
    system = AcousticSystem(classifier=rx, transform=transform, defender=dw, defense_type="wave")
    x1 = torch.from_numpy(synthetic.synthetic_waveforms(1, 16000, seed=1234)).requires_grad_(True)
    with NoiseInjector(2027) as inj:
        logits = system(x1)
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor([3]))
    (gx,) = torch.autograd.grad(loss, x1)
    out["system_logits"] = logits.detach().numpy()
    out["system_loss_grad"] = gx.numpy()

    np.savez_compressed(os.path.join(HERE, "reference_golden_grad.npz"), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).max()))


if __name__ == "__main__":
    main()
