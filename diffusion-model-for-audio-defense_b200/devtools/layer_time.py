"""Development aid: time of each of the first 12 residual layers (one dilation cycle) of a 148-waveform launch, as differences of
net.debug_layer(x, t, layer) run times; AP_LIB_PATH selects the build."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import audiopure_b200 as ap
from audiopure_b200 import synthetic
net = ap.WaveNet(synthetic.wavenet_state_dict(seed=0), mode="bf16", **synthetic.DEFAULT_WAVENET_CONFIG)
x = torch.from_numpy(synthetic.synthetic_waveforms(148, 16000, seed=1)).cuda()
def t_upto(layer, reps=6):
    net.debug_layer(x, 1.0, layer); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): net.debug_layer(x, 1.0, layer)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ts = [t_upto(l) for l in range(0, 13)]
print(os.environ.get("AP_LIB_PATH", "default")[-16:], " ".join(f"{ts[i+1]-ts[i]:.3f}" for i in range(12)))
