"""Development probe: ResNeXt-29 8x64 classifier throughput at batch 512 (AP_CLS_CHUNK = images per pass)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402

rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
x = torch.randn(512, 1, 32, 32, device="cuda")
rx(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    rx(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"AP_CLS_CHUNK={os.environ.get('AP_CLS_CHUNK', '64')}: {ms:.2f} ms per 512 images ({10.77 * 512 / ms:.0f} TFLOP/s tf32)")
