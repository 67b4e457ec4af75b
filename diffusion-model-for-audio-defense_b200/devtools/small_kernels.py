"""Development aid: launches each of the small hot kernels once at the benchmark shape, in a fixed order, so that one
`ncu --set full -k regex:"ew_kernel|k_mel|mel_prep|k1_split|k_conv|init_smooth" -c 30` capture covers them all:
  4 update kernels (B = 512 x 16000), the log-mel pair (512 waveforms), two k1_split launches (bf16x3, 74 waveforms),
  then the ResNeXt-29 forward (64 spectrograms; the capture count cuts it off)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import _lib, synthetic  # noqa: E402


def main():
    lib = _lib.load()
    B, L = 512, 16000
    st = _lib.stream_ptr()
    x = torch.randn(B, L, device="cuda")
    e = torch.randn(B, L, device="cuda")
    out = torch.empty(B, L, device="cuda")
    sd = synthetic.wavenet_state_dict(seed=0)
    net = ap.WaveNet(sd, mode="bf16x3", **synthetic.DEFAULT_WAVENET_CONFIG)
    tr = ap.sc09_transform()
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    wav = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=5)).cuda()
    spec = tr(wav)                      # warm-up: allocations, tensor maps
    rx(spec[:64])
    net.debug_layer(wav[:74], 1.0, 1)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()          # ncu --profile-from-start off: only what follows is captured
    _lib.check(lib.ap_ddpm_step(x.data_ptr(), e.data_ptr(), 0.0115, 0.9999, 0.0082, None, 7, 0, B, L, st))
    _lib.check(lib.ap_diffuse(x.data_ptr(), 0.9997, 0.0245, None, 7, 0, out.data_ptr(), B, L, st))
    _lib.check(lib.ap_smooth_inputs(x.data_ptr(), 0.5, 0.8944, None, 7, 0, out.data_ptr(), B, L, st))
    _lib.check(lib.ap_predict_x0(x.data_ptr(), e.data_ptr(), 1.118, 0.5, out.data_ptr(), B, L, st))
    tr(wav)
    net.debug_layer(wav[:74], 1.0, 1)
    rx(spec[:64])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
