"""Development aid (not the benchmark): A/B two builds of libaudiopure_b200.so on the SAME GPU box, alternating, with
sustained (power-capped) timing of the tensor-core WaveNet.  Box-to-box and boost-vs-sustained differences are larger
than the few-percent kernel changes this is used to judge.

Usage: python ab_compare.py libA.so libB.so [libC.so ...] [rounds] [seconds_per_measurement]
       python ab_compare.py --one lib.so seconds     (internal: one measurement, prints a JSON line)"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(seconds):
    sys.path.insert(0, ROOT)
    import torch
    import audiopure_b200 as ap
    from audiopure_b200 import _lib, synthetic
    B, L = 512, 16000
    sd = synthetic.wavenet_state_dict(seed=0)
    net = ap.WaveNet(sd, mode=os.environ.get("AP_AB_MODE", "bf16"), **synthetic.DEFAULT_WAVENET_CONFIG)
    lib = _lib.load()
    x = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=1)).cuda()
    out = torch.empty_like(x)
    net.eps(x, 1.0, out=out)
    torch.cuda.synchronize()
    t_end = time.time() + seconds * 0.4          # warm into the power-capped regime
    while time.time() < t_end:
        net.eps(x, 1.0, out=out)
        torch.cuda.synchronize()
    lib.ap_diffwave_profile(net._handle, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 0
    e0.record()
    t_end = time.time() + seconds * 0.6
    while time.time() < t_end:
        net.eps(x, 1.0, out=out)
        torch.cuda.synchronize()
        reps += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pm, pc = (C.c_double * 2)(), (C.c_int * 2)()
    lib.ap_diffwave_profile_read(net._handle, pm, pc)
    print(json.dumps({"ms_per_eps512": ms, "tflops": 606.1 * B / ms, "k1_ms": pm[0] / max(pc[0], 1),
                      "k2_ms": pm[1] / max(pc[1], 1), "reps": reps}))


def main():
    if sys.argv[1] == "--one":
        return one(float(sys.argv[3]))
    libs = [os.path.abspath(a) for a in sys.argv[1:] if a.endswith(".so")]
    nums = [a for a in sys.argv[1:] if not a.endswith(".so")]
    rounds = int(nums[0]) if nums else 2
    secs = nums[1] if len(nums) > 1 else "12"
    for r in range(rounds):
        for name, lib in zip("ABCDEFGH", libs):
            env = dict(os.environ, AP_LIB_PATH=lib)
            o = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", lib, secs], env=env,
                               capture_output=True, text=True)
            line = o.stdout.strip().splitlines()[-1] if o.stdout.strip() else o.stderr[-400:]
            print(f"round {r} {name} {os.path.basename(lib)}: {line}", flush=True)


if __name__ == "__main__":
    main()
