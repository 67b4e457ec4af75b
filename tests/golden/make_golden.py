"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference is imported through the shim described in SURVEY.md §8c:
  (1) stub ``diffusion_models.DiffWave_Unconditional.dataset`` (imports librosa / removed torchaudio API),
  (2) put ``diffusion_models/DiffWave_Unconditional`` on sys.path (WaveNet.py:7 does ``from util import``),
  (3) make ``.cuda()`` the identity (hard-coded in util.py:65,88 / diffwave_ddpm.py:66,100,157),
  (4) stub ``torchsde`` and ``statsmodels`` (absent; only RevVPSDE.f/g and smooth_predict are exercised).
Weights are the seeded synthetic state dicts of ``audiopure_b200.synthetic`` loaded into the reference
modules with ``load_state_dict``; noise is host-generated (``synthetic.host_noise``) and injected by
patching ``torch.normal``, so the oracle / CUDA path can be fed the identical tensors.
"""
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AUDIOPURE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

import audiopure_b200  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402


def install_shim():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "diffusion_models", "DiffWave_Unconditional"))
    sys.path.insert(0, os.path.join(REF, "audio_models", "ConvNets_SpeechCommands"))
    stub = types.ModuleType("diffusion_models.DiffWave_Unconditional.dataset")
    stub.load_Qualcomm_keyword = lambda *a, **k: None
    sys.modules["diffusion_models.DiffWave_Unconditional.dataset"] = stub
    sys.modules["torchsde"] = types.ModuleType("torchsde")
    sm = types.ModuleType("statsmodels")
    sms = types.ModuleType("statsmodels.stats")
    smp = types.ModuleType("statsmodels.stats.proportion")
    smp.proportion_confint = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("statsmodels absent"))
    sys.modules.update({"statsmodels": sm, "statsmodels.stats": sms, "statsmodels.stats.proportion": smp})
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


class NoiseInjector:
    """Replaces torch.normal(mean, std, size=...) by host noise popped in call order."""

    def __init__(self, seed=2024):
        self.seed, self.i, self.log = seed, 0, []
        self._orig = torch.normal

    def __enter__(self):
        def fake(mean, std, size=None, **kw):
            z = synthetic.host_noise(tuple(size), self.seed, self.i)
            self.i += 1
            self.log.append(z)
            return torch.from_numpy(z) * std + mean
        torch.normal = fake
        return self

    def __exit__(self, *a):
        torch.normal = self._orig


def to_torch_sd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def main():
    install_shim()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from diffusion_models.diffwave_ddpm import DiffWave
    from diffusion_models.diffwave_sde import RevVPSDE
    from diffusion_models.DiffWave_Unconditional.WaveNet import WaveNet_Speech_Commands
    from diffusion_models.DiffWave_Unconditional.util import (calc_diffusion_hyperparams,
                                                              calc_diffusion_step_embedding)
    from robustness_eval.certified_robust import RobustCertificate
    from acoustic_system import AcousticSystem
    import torchaudio

    cfg = json.load(open(os.path.join(REF, "configs", "config.json")))
    out = {}

    # ---- a1 / a2: schedule tables and step embedding
    hp = calc_diffusion_hyperparams(**cfg["diffusion_config"])
    for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"):
        out["hp_" + k] = hp[k].numpy()
    steps = torch.tensor([[0.0], [1.0], [65.0], [199.0]])
    out["emb_steps"] = steps.numpy()
    out["emb"] = calc_diffusion_step_embedding(steps, 128).numpy()

    # ---- a3-a6: WaveNet eps, full architecture (36 x 256), short clip
    net = WaveNet_Speech_Commands(**cfg["wavenet_config"]).eval()
    net.load_state_dict(to_torch_sd(synthetic.wavenet_state_dict(seed=0)))
    x = torch.from_numpy(synthetic.synthetic_waveforms(2, 1024, seed=1234))
    with torch.no_grad():
        out["eps_full_L1024_t1"] = net((x.clone(), 1.0 * torch.ones(2, 1))).numpy()
        out["eps_full_L1024_t65"] = net((x.clone(), 65.0 * torch.ones(2, 1))).numpy()
    # odd length (not a multiple of any tile), exercises zero padding at both ends with dilation up to 2048
    x3 = torch.from_numpy(synthetic.synthetic_waveforms(1, 3001, seed=77))
    with torch.no_grad():
        out["eps_full_L3001_t7"] = net((x3.clone(), 7.0 * torch.ones(1, 1))).numpy()

    # ---- small architecture (different channel/layer counts; checks the config plumbing)
    small_cfg = dict(cfg["wavenet_config"], res_channels=64, skip_channels=64, num_res_layers=5, dilation_cycle=3)
    net_s = WaveNet_Speech_Commands(**small_cfg).eval()
    net_s.load_state_dict(to_torch_sd(synthetic.wavenet_state_dict(seed=3, config=small_cfg)))
    xs = torch.from_numpy(synthetic.synthetic_waveforms(3, 500, seed=5))
    with torch.no_grad():
        out["eps_small_L500_t3"] = net_s((xs.clone(), 3.0 * torch.ones(3, 1))).numpy()

    # ---- a7-a10: DDPM purifier with injected noise
    dw = DiffWave(model=net, diffusion_hyperparams=hp, reverse_timestep=2).eval()
    with torch.no_grad(), NoiseInjector(2024) as inj:
        out["ddpm_t2_L1024"] = dw(x.clone()).numpy()
        assert inj.i == 2
    dw.reverse_timestep = 3
    with torch.no_grad(), NoiseInjector(2025) as inj:
        out["ddpm_t3_L1024"] = dw(x.clone()).numpy()
        assert inj.i == 3
    dw.reverse_timestep = 66
    with torch.no_grad():
        out["oneshot_t66_L1024"] = dw.one_shot_denoise(x.clone()).numpy()
        out["twoshot_t66_L1024"] = dw.two_shot_denoise(x.clone()).numpy()
    dw.reverse_timestep = 9
    with torch.no_grad(), NoiseInjector(2026) as inj:
        out["fastrev_t9_L1024"] = dw.fast_reverse(x.clone()).numpy()
        assert inj.i == 3

    # ---- a11: ReffWave (repeated diffuse + one-shot denoise)
    from diffusion_models.diffwave_ddpm import ReffWave
    rw = ReffWave(model=net, diffusion_hyperparams=hp, reverse_timestep=5, num_re=2).eval()
    with torch.no_grad(), NoiseInjector(2029) as inj:
        out["reffwave_t5_re2_L1024"] = rw(x.clone()).numpy()
        assert inj.i == 2

    # ---- a12: RevVPSDE drift / diffusion at solver times (sdeint itself is torchsde: absent)
    dw.reverse_timestep = 5
    sde = RevVPSDE(model=dw, score_type="guided_diffusion", beta_min=0.0001 * 200, beta_max=0.02 * 200, N=200,
                   audio_shape=(1, 1024))
    xf = x.reshape(2, -1).clone()
    svals = [0.975, 0.98, 0.985, 0.99, 0.995, 0.99999]
    out["sde_s"] = np.array(svals, dtype=np.float32)
    fs, gs = [], []
    with torch.no_grad():
        for s in svals:
            st = torch.tensor(s, dtype=torch.float32)
            fs.append(sde.f(st, xf.clone()).numpy())
            gs.append(sde.g(st, xf.clone()).numpy()[:, 0])
    out["sde_f"] = np.stack(fs)
    out["sde_g"] = np.stack(gs)

    # ---- a14: mel front ends (torchaudio)
    xm = torch.from_numpy(synthetic.synthetic_waveforms(2, 16000, seed=99))
    mel_sc = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32,
                                                  norm="slaney", pad_mode="constant", mel_scale="slaney")
    todb = torchaudio.transforms.AmplitudeToDB(stype="power")
    out["mel_sc09"] = todb(mel_sc(xm)).numpy()
    mel_kws = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_mels=32)
    out["mel_kws"] = todb(mel_kws(xm)).numpy()
    out["mel_sc09_fb"] = mel_sc.mel_scale.fb.numpy()
    out["mel_kws_fb"] = mel_kws.mel_scale.fb.numpy()

    # ---- a15: ResNeXt-29 8x64 logits on the reference's own mel features
    from models.resnext import CifarResNeXt
    rx = CifarResNeXt(nlabels=10, in_channels=1).eval()
    rx.load_state_dict(to_torch_sd(synthetic.resnext_state_dict(seed=0)))
    with torch.no_grad():
        out["resnext_logits"] = rx(torch.from_numpy(out["mel_sc09"])).numpy()

    # ---- a16: ResNet family (BasicBlock and Bottleneck variants) on the same mel features
    from models.resnet import resnet34, resnet50
    for depth, ctor in ((34, resnet34), (50, resnet50)):
        rn = ctor(num_classes=10, in_channels=1).eval()
        rn.load_state_dict(to_torch_sd(synthetic.resnet_state_dict(depth=depth, seed=0)))
        with torch.no_grad():
            out[f"resnet{depth}_logits"] = rn(torch.from_numpy(out["mel_sc09"])).numpy()

    # ---- a17 / a18: M5 and RCNN_KWS
    sys.path.insert(0, os.path.join(REF, "audio_models", "M5"))
    from M5Net import M5
    m5 = M5(n_input=1, first_kernel_size=160, n_output=10).eval()
    m5.load_state_dict(to_torch_sd(synthetic.m5_state_dict(seed=0)))
    with torch.no_grad():
        out["m5_logprobs"] = m5(xm).numpy()
    sys.path.insert(0, os.path.join(REF, "audio_models", "RCNN_KWS"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("rcnn_kws_model", os.path.join(REF, "audio_models", "RCNN_KWS", "model.py"))
    kmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kmod)
    kws = kmod.KWSModel(in_size=32).eval()
    kws.load_state_dict(to_torch_sd(synthetic.kws_state_dict(seed=0)))
    with torch.no_grad():
        out["kws_logprobs"] = kws(torch.from_numpy(out["mel_kws"])).numpy()

    # ---- a20: full AcousticSystem (DDPM t*=2 -> mel -> ResNeXt), one 1 s clip
    import torchvision  # noqa: F401  (the drivers compose with torchvision; a lambda is equivalent)
    transform = lambda w: todb(mel_sc(w))
    dw.reverse_timestep = 2
    system = AcousticSystem(classifier=rx, transform=transform, defender=dw, defense_type="wave")
    x1 = torch.from_numpy(synthetic.synthetic_waveforms(1, 16000, seed=1234))
    with torch.no_grad(), NoiseInjector(2027) as inj:
        out["system_purified"] = dw(x1.clone()).numpy()
    with torch.no_grad(), NoiseInjector(2027) as inj:
        out["system_logits"] = system(x1.clone()).numpy()
    with torch.no_grad():   # int16-range input triggers the /2**15 branch (acoustic_system.py:29-30)
        out["system_logits_int16_nodefense"] = system(x1.clone() * 2 ** 15, defend=False).numpy()

    # ---- a21: smooth_predict counts on a short draw budget (n=12, batch 5 -> ragged last batch), sigma=0.5
    rc = RobustCertificate(classifier=rx, transform=transform, denoiser=dw, num_classes=10)
    with torch.no_grad(), NoiseInjector(2028) as inj:
        counts = rc.smooth_predict(x1[0:1].clone(), num_sampling=12, sigma=0.5, batch_size=5)
        assert inj.i == 3
    out["smooth_counts_n12"] = counts.numpy()
    out["smooth_t_star"] = np.array([rc.compute_t_star(1 / (1 + s ** 2)) for s in (0.25, 0.5, 1.0)])

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    for k, v in out.items():
        print(f"{k:34s} {str(v.shape):18s} {v.dtype}  absmax={np.abs(v).max():.5g}")


if __name__ == "__main__":
    main()
