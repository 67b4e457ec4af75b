"""Development probe (not the benchmark): times the tensor-core WaveNet and prints the k1_layer wait-cycle counters.
Usage: AP_TC_DEBUG=1 [AP_TC_PAIR=0|1] python probe.py [B] [chunk]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import _lib, synthetic  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else B
    sd = synthetic.wavenet_state_dict(seed=0)
    net = ap.WaveNet(sd, mode=os.environ.get("AP_PROBE_MODE", "bf16"), **synthetic.DEFAULT_WAVENET_CONFIG)
    lib = _lib.load()
    L = 16000
    x = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=1)).cuda()
    out = torch.empty_like(x)
    net.reserve(chunk, L)
    net.eps(x, 1.0, out=out)
    torch.cuda.synchronize()
    lib.ap_diffwave_profile(net._handle, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        net.eps(x, 1.0, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pm, pc = (C.c_double * 2)(), (C.c_int * 2)()
    lib.ap_diffwave_profile_read(net._handle, pm, pc)
    lib.ap_diffwave_profile(net._handle, 0)
    k1 = pm[0] / max(pc[0], 1)
    print(f"pair={os.environ.get('AP_TC_PAIR', '1')} B={B} chunk={chunk}: {ms:.2f} ms/eps ({606.1 * B / ms:.1f} TFLOP/s); "
          f"k1 avg {k1:.3f} ms ({14.68 * min(B, chunk) / k1:.1f} TFLOP/s, n={pc[0]}); k2 avg {pm[1] / max(pc[1], 1):.3f} ms", flush=True)
    if os.environ.get("AP_TC_DEBUG") == "1":
        buf = np.zeros((256, 16), dtype=np.int64)
        _lib.check(lib.ap_diffwave_debug_counters(net._handle, buf.ctypes.data))
        names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_accempty", "mma_wait_outready", "mma_total",
                 "tiles", "epi_wait_accfull", "epi_wait_g2", "epi_total", "epi_wait_bulk"]
        cnt = buf[:200]
        act = cnt[cnt[:, 1] > 0]
        lead = act[act[:, 5] > 0]
        print(f"active CTAs {len(act)}, MMA-issuing CTAs {len(lead)} (last k1 launch; cycles, mean over CTAs)")
        for i, nm in enumerate(names):
            src = lead if nm.startswith("mma") or nm == "tiles" else act
            print(f"  {nm:20s} {src[:, i].mean():14.0f}")
        tr = buf[200:].reshape(-1)[: 3 * 32].reshape(3, 32)
        if tr.any():
            names_t = {0: "G1c0 wait acc_empty", 1: "G1c0 issue start", 2: "G1c0 issued", 28: "G2 wait acc_empty", 29: "G2 issue start(after out_ready)",
                       30: "G2 issued", 4: "G1c1 wait acc_empty", 5: "G1c1 issue start", 6: "G1c1 issued", 8: "gate0 acc_full", 9: "gate0 tmem loaded",
                       10: "gate0 math done", 11: "gate0 staged", 24: "res begin(u loads issued)", 25: "res acc_full", 26: "res tmem loaded",
                       27: "res done", 31: "res u loads issued", 12: "gate0 barrier1 passed", 13: "gate0 barrier2 passed", 14: "gate0 tma stores issued",
                       20: "gate1 barrier1 passed", 21: "gate1 barrier2 passed", 22: "gate1 tma stores issued", 16: "gate1 acc_full", 17: "gate1 tmem loaded", 18: "gate1 math done", 19: "gate1 staged"}
            t0 = tr[0][tr[0] > 0].min()
            ev = sorted((int(tr[k, s] - t0), k, names_t[s]) for k in range(3) for s in names_t if tr[k, s] > 0)
            print("timeline of CTA 0 (clk since the first event; tile index relative to tile 60):")
            for clk, k, nm in ev:
                print(f"  {clk:8d}  tile+{k}  {'MMA ' if nm.startswith('G') else 'EPI '} {nm}")
        t = lead[:, 6].mean()
        print(f"  per tile: mma_total {lead[:, 5].mean() / t:.0f} clk = ideal MMA 14336 (12288 for the last layer) + wait_full {lead[:, 2].mean() / t:.0f}"
              f" + wait_accempty {lead[:, 3].mean() / t:.0f} + wait_outready {lead[:, 4].mean() / t:.0f} + issue/other")


if __name__ == "__main__":
    main()
