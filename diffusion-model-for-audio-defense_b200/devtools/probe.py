"""Quick GPU probe (development aid, not the benchmark): times the tensor-core WaveNet at several chunk sizes."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import _lib, synthetic  # noqa: E402


def main():
    sd = synthetic.wavenet_state_dict(seed=0)
    net = ap.WaveNet(sd, mode="bf16", **synthetic.DEFAULT_WAVENET_CONFIG)
    lib = _lib.load()
    L = 16000
    for B, chunk in [(8, 8), (32, 16), (32, 32), (64, 64)]:
        x = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=1)).cuda()
        out = torch.empty_like(x)
        net.reserve(chunk, L)
        net.eps(x, 1.0, out=out)
        torch.cuda.synchronize()
        lib.ap_diffwave_profile(net._handle, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 2
        for _ in range(reps):
            net.eps(x, 1.0, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        pm = (C.c_double * 2)()
        pc = (C.c_int * 2)()
        lib.ap_diffwave_profile_read(net._handle, pm, pc)
        lib.ap_diffwave_profile(net._handle, 0)
        fl = 606.1e9 * B
        k1_fl = 14.68e9 * chunk
        print(f"B={B} chunk={chunk}: {ms:.2f} ms/eps  -> {fl / ms / 1e9:.1f} TFLOP/s overall; "
              f"k1 avg {pm[0] / max(pc[0], 1):.3f} ms ({k1_fl / (pm[0] / max(pc[0], 1)) / 1e9:.1f} TFLOP/s, n={pc[0]}); "
              f"k2 avg {pm[1] / max(pc[1], 1):.3f} ms (n={pc[1]})", flush=True)


if __name__ == "__main__":
    main()
