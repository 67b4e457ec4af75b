"""Development probe (not the benchmark): times the backward pass of the DiffWave network (ap_diffwave_eps_vjp =
recomputed forward with saved gate derivatives + 1 + 2 x 36 backward GEMM launches) on 1 s waveforms.
Usage: python vjp_probe.py [B] [mode]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    L = 16000
    net = ap.WaveNet(synthetic.wavenet_state_dict(seed=0), mode=mode, **synthetic.DEFAULT_WAVENET_CONFIG)
    x = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=1)).cuda()
    g = torch.randn_like(x)
    net.eps_vjp(x, 7.0, g)
    net.eps(x, 7.0)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 3
    e0.record()
    for _ in range(reps):
        net.eps(x, 7.0)
    e1.record()
    for _ in range(reps):
        net.eps_vjp(x, 7.0, g)
    e2.record()
    torch.cuda.synchronize()
    fwd, vjp = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
    # VJP = forward recompute (606.1 GFLOP) + backward GEMMs (603.98 GFLOP: the same contraction sizes transposed)
    print(f"mode={mode} B={B}: forward {fwd:.2f} ms ({606.1 * B / fwd:.0f} TFLOP/s); vjp (forward recompute + backward) {vjp:.2f} ms "
          f"({(606.1 + 603.98) * B / vjp:.0f} TFLOP/s of bf16-equivalent algorithmic work), backward alone ~{vjp - fwd:.2f} ms "
          f"({603.98 * B / (vjp - fwd):.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
