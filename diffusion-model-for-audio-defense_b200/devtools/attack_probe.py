"""Development probe: time of one white-box attack iteration (forward + loss.backward() through DDPM purifier -> log-mel ->
ResNeXt) and of its parts.  Usage: python attack_probe.py [B] [t_star] [mode]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402

CONFIG_JSON = os.path.join(ROOT, "diffusion-model-for-audio-defense_b200", "configs", "config.json")


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    t_star = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=t_star, state_dict=synthetic.wavenet_state_dict(seed=0),
                                  noise="philox", seed=1, mode=mode)
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    tr = ap.sc09_transform()
    system = ap.AcousticSystem(classifier=rx, transform=tr, defender=dw, defense_type="wave", check_int16_range=False)
    x = torch.from_numpy(synthetic.synthetic_waveforms(B, 16000, seed=5)).cuda()
    y = torch.zeros(B, dtype=torch.long, device="cuda")

    def step():
        xr = x.clone().requires_grad_(True)
        loss = torch.nn.functional.cross_entropy(system(xr), y)
        (g,) = torch.autograd.grad(loss, xr)
        return g

    def fwd():
        with torch.no_grad():
            system(x)

    def cls_only():
        s = tr(x).detach().requires_grad_(True)
        (g,) = torch.autograd.grad(rx(s).sum(), s)

    def mel_only():
        xr = x.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(tr(xr).sum(), xr)

    print(f"B={B} t*={t_star} mode={mode}: inference forward {timed(fwd):.1f} ms; attack iteration (forward + backward) {timed(step):.1f} ms; "
          f"of which ResNeXt forward+backward {timed(cls_only):.1f} ms, mel forward+backward {timed(mel_only):.1f} ms", flush=True)


if __name__ == "__main__":
    main()
