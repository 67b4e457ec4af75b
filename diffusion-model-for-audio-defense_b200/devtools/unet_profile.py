"""Development aid: one forward (B = 512) and one input-gradient call (B = 128) of the spectrogram UNet between
cudaProfilerStart / Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum`; prints CUDA-event times first."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    net = ap.UNet(synthetic.unet_state_dict(seed=0))
    mode = os.environ.get("AP_UNET_MODE", "tf32")
    net.set_mode(mode)
    B, Bg = int(os.environ.get("AP_UNET_B", 512)), int(os.environ.get("AP_UNET_BG", 128))
    x = torch.randn(B, 1, 32, 32, device="cuda")
    g = torch.randn(Bg, 1, 32, 32, device="cuda")
    print(f"UNet {mode}: eps B={B} {timed(lambda: net.eps(x, 37.0)):.2f} ms; eps_vjp B={Bg} {timed(lambda: net.eps_vjp(x[:Bg], 37.0, g)):.2f} ms")
    torch.cuda.profiler.start()
    net.eps(x, 37.0)
    net.eps_vjp(x[:Bg], 37.0, g)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
