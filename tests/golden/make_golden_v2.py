"""Golden pack v2: outputs of the UNMODIFIED reference (/root/reference, CPU) at the BASELINE shapes.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_v2.py            # ~10-15 min of host work (about 100 network evaluations at L = 16 000)

Writes
  tests/golden/reference_golden_v2.npz       values listed below
  tests/golden/reference_checkpoints.npz     the state dicts of the two TRAINED classifiers the reference ships
                                             (audio_models/M5/checkpoints/kernel_size=160/vanilla-best-acc.pth,
                                              audio_models/RCNN_KWS/checkpoints/vanilla-best-acc-kws-attn_rcnn-n_mels=32.pth)
  tests/golden/m5_k160_vanilla_best_acc.pth  byte copy of the M5 pickle (a whole pickled module: what create_model() loads)

Contents of reference_golden_v2.npz (same shim, synthetic weights and noise injection as make_golden.py):
  eps_L16000_t{1,65,116}            WaveNet_Speech_Commands eps, B = 2, L = 16 000          (WaveNet.py:164-172)
  smooth_x0_sigma{0.25,0.5,1.0}     one_shot_denoise of the smoothing-level inputs sqrt(abar*) (x + sigma z), B = 2,
                                    t* = 34 / 66 / 117                                       (certified_robust.py:44-54)
  smooth_logits_sigma{...}          ... -> mel -> ResNeXt logits of the same 2 draws + 6 more (8 draws per sigma)
  smooth_counts_sigma{...}          RobustCertificate.smooth_predict counts over those 8 draws (batch 4)
  top1_purified_first2              DDPM t* = 2 purified waveforms of the first 2 of 32 clips
  top1_logits, top1_m5_logprobs     DDPM t* = 2 -> mel -> ResNeXt logits / -> trained M5 log-probs for 32 clips
  resnext_centred_bias              classifier.bias of the ResNeXt used above (random init whose logits are centred over
                                    the 32 clips, so that top-1 varies from clip to clip instead of being unanimous)
  kws2s_mel, kws2s_logprobs         KWS mel (400/200/32) and trained RCNN_KWS log-probs of 16 clips of 2 s (W = 161)
  kws2s_purified_logprobs           DDPM t* = 2 at L = 32 000 -> KWS mel -> trained RCNN_KWS, 2 clips (BASELINE configs[4])
  m5_trained_logprobs               trained M5 on 16 raw 1 s clips
"""
import json
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, ROOT, NoiseInjector, install_shim, synthetic, to_torch_sd  # noqa: E402

M5_PICKLE = os.path.join(REF, "audio_models", "M5", "checkpoints", "kernel_size=160", "vanilla-best-acc.pth")
KWS_CKPT = os.path.join(REF, "audio_models", "RCNN_KWS", "checkpoints", "vanilla-best-acc-kws-attn_rcnn-n_mels=32.pth")
SIGMAS = ((0.25, 34), (0.5, 66), (1.0, 117))


def main():
    install_shim()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    from diffusion_models.diffwave_ddpm import DiffWave
    from diffusion_models.DiffWave_Unconditional.WaveNet import WaveNet_Speech_Commands
    from diffusion_models.DiffWave_Unconditional.util import calc_diffusion_hyperparams
    from robustness_eval.certified_robust import RobustCertificate
    from models.resnext import CifarResNeXt
    import torchaudio

    cfg = json.load(open(os.path.join(REF, "configs", "config.json")))
    hp = calc_diffusion_hyperparams(**cfg["diffusion_config"])
    net = WaveNet_Speech_Commands(**cfg["wavenet_config"]).eval()
    net.load_state_dict(to_torch_sd(synthetic.wavenet_state_dict(seed=0)))
    dw = DiffWave(model=net, diffusion_hyperparams=hp, reverse_timestep=2).eval()
    mel_sc = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32,
                                                  norm="slaney", pad_mode="constant", mel_scale="slaney")
    mel_kws = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_mels=32)
    todb = torchaudio.transforms.AmplitudeToDB(stype="power")
    transform = lambda w: todb(mel_sc(w))
    out = {}

    # ---- trained classifiers shipped with the reference
    sys.path.insert(0, os.path.join(REF, "audio_models", "M5"))
    m5 = torch.load(M5_PICKLE, map_location="cpu", weights_only=False).float().eval()
    import importlib.util
    spec = importlib.util.spec_from_file_location("rcnn_kws_model", os.path.join(REF, "audio_models", "RCNN_KWS", "model.py"))
    kmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kmod)
    kws = kmod.KWSModel(in_size=32).eval()
    kws_sd = torch.load(KWS_CKPT, map_location="cpu", weights_only=False)
    kws.load_state_dict(kws_sd)
    ck = {"m5." + k: v.numpy() for k, v in m5.state_dict().items()}
    ck.update({"kws." + k: v.numpy() for k, v in kws_sd.items()})
    np.savez_compressed(os.path.join(HERE, "reference_checkpoints.npz"), **ck)
    shutil.copyfile(M5_PICKLE, os.path.join(HERE, "m5_k160_vanilla_best_acc.pth"))

    # ---- (1) eps at the benchmark length
    x2 = torch.from_numpy(synthetic.synthetic_waveforms(2, 16000, seed=1234))
    with torch.no_grad():
        for t in (1, 65, 116):
            out[f"eps_L16000_t{t}"] = net((x2.clone(), float(t) * torch.ones(2, 1))).numpy()
            print("eps", t, flush=True)

    # ---- (3) DDPM t* = 2 -> mel -> ResNeXt on 32 clips; classifier bias centred over them so that top-1 varies
    x32 = torch.from_numpy(synthetic.synthetic_waveforms(32, 16000, seed=4321))
    dw.reverse_timestep = 2
    pur = []
    with torch.no_grad(), NoiseInjector(2040) as inj:
        for i in range(0, 32, 8):          # one DiffWave.forward per 8 clips: noise indices 2i, 2i+1 of seed 2040, shape (8,1,L)
            pur.append(dw(x32[i:i + 8].clone()))
            print("ddpm", i, flush=True)
        assert inj.i == 8
    pur = torch.cat(pur)
    out["top1_purified_first2"] = pur[:2].numpy()
    rx_sd = synthetic.resnext_state_dict(seed=0)
    rx = CifarResNeXt(nlabels=10, in_channels=1).eval()
    rx.load_state_dict(to_torch_sd(rx_sd))
    with torch.no_grad():
        raw = rx(transform(pur))
    bias = (torch.from_numpy(rx_sd["classifier.bias"]) - raw.mean(0)).float()
    out["resnext_centred_bias"] = bias.numpy()
    rx.classifier.bias.data.copy_(bias)
    with torch.no_grad():
        out["top1_logits"] = rx(transform(pur)).numpy()
        out["top1_m5_logprobs"] = m5(pur).numpy()
    print("top-1:", out["top1_logits"].argmax(1).tolist(), "m5:", out["top1_m5_logprobs"].argmax(1).tolist(), flush=True)

    # ---- (2) one-shot denoising of smoothing-level inputs, sigma in {0.25, 0.5, 1.0}
    x1 = x2[0:1]
    rc = RobustCertificate(classifier=rx, transform=transform, denoiser=dw, num_classes=10)
    for sigma, t_star in SIGMAS:
        assert rc.compute_t_star(1 / (1 + sigma ** 2)) == t_star
        seed = 3000 + int(sigma * 100)
        with torch.no_grad(), NoiseInjector(seed) as inj:
            counts = rc.smooth_predict(x1.clone(), num_sampling=8, sigma=sigma, batch_size=4)
            assert inj.i == 2
        out[f"smooth_counts_sigma{sigma}"] = counts.numpy()
        # the same two micro-batches step by step (certified_robust.py:46-54), keeping x0_hat and the logits
        lg = []
        for b in range(2):
            x_in = x1.repeat(4, 1, 1)
            delta = torch.from_numpy(synthetic.host_noise((4, 1, 16000), seed, b)) * sigma + 0
            x_in = x_in + delta
            alpha_bar_star = 1 / (1 + sigma ** 2)
            dw.reverse_timestep = t_star
            x_in = alpha_bar_star ** 0.5 * x_in
            with torch.no_grad():
                x0 = dw.one_shot_denoise(x_in)
                lg.append(rx(transform(x0)).numpy())
            if b == 0:
                out[f"smooth_x0_sigma{sigma}"] = x0[:2].numpy()
        lg = np.concatenate(lg)
        out[f"smooth_logits_sigma{sigma}"] = lg
        assert np.array_equal(np.bincount(lg.argmax(1), minlength=10), counts.numpy())
        print("smooth", sigma, counts.tolist(), flush=True)

    # ---- (4) KWS at 2 s (W = 161) and trained M5 / RCNN_KWS
    xk = torch.from_numpy(synthetic.synthetic_waveforms(16, 32000, seed=555))
    with torch.no_grad():
        out["kws2s_mel"] = todb(mel_kws(xk[:2])).numpy()
        out["kws2s_logprobs"] = kws(todb(mel_kws(xk))).numpy()
        out["m5_trained_logprobs"] = m5(torch.from_numpy(synthetic.synthetic_waveforms(16, 16000, seed=556))).numpy()
    dw.reverse_timestep = 2
    with torch.no_grad(), NoiseInjector(2041) as inj:
        pk = dw(xk[:2].clone())
    with torch.no_grad():
        out["kws2s_purified_first1"] = pk[:1].numpy()
        out["kws2s_purified_logprobs"] = kws(todb(mel_kws(pk))).numpy()

    np.savez_compressed(os.path.join(HERE, "reference_golden_v2.npz"), **out)
    for k, v in out.items():
        print(f"{k:34s} {str(v.shape):18s} {v.dtype}  absmax={np.abs(v).max():.5g}")


if __name__ == "__main__":
    main()
