"""Development aid: time ap_diffwave_eps_vjp (forward recompute + backward) and the backward alone, and dump g_x so that the fused
(default) and two-launch (AP_BWD_UNFUSED=1) backward passes can be compared across processes.
usage: bwd_profile.py out.npy [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402


def main():
    out, B, L = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 32, 16000
    net = ap.WaveNet(synthetic.wavenet_state_dict(seed=0), mode=os.environ.get("AP_AB_MODE", "bf16"), **synthetic.DEFAULT_WAVENET_CONFIG)
    x = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=3)).cuda()
    g = torch.from_numpy(synthetic.host_noise((B, 1, L), 77, 0)).cuda()
    gx = net.eps_vjp(x, 3.0, g)
    torch.cuda.synchronize()
    np.save(out, gx.cpu().numpy())

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

    t_vjp = timed(lambda: net.eps_vjp(x, 3.0, g))
    t_fwd = timed(lambda: net.eps(x, 3.0))
    print(f"B={B}: eps_vjp {t_vjp:.2f} ms, plain forward {t_fwd:.2f} ms, |g_x| {float(gx.abs().mean()):.4e}, finite {bool(torch.isfinite(gx).all())}")


if __name__ == "__main__":
    main()
