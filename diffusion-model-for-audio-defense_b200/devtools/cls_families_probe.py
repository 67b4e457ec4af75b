"""Development probe: forward and forward+backward time of every classifier family at batch 512 (one JSON line)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402

B = 512
FAMILIES = {   # name: (constructor, GFLOP per sample forward, input)
    "resnext29_8_64": (lambda: ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0)), 10.77, "spec"),
    "resnet34": (lambda: ap.ResNetClassifier(synthetic.resnet_state_dict(depth=34, seed=0), depth=34), 0.146, "spec"),
    "vgg19_bn": (lambda: ap.VGGClassifier(synthetic.vgg_state_dict(depth=19, seed=0), depth=19), 0.80, "spec"),
    "wideresnet28_10": (lambda: ap.WideResNetClassifier(synthetic.wideresnet_state_dict(28, 10, seed=0), 28, 10), 10.5, "spec"),
    "densenet_bc_100_12": (lambda: ap.DenseNetClassifier(synthetic.densenet_state_dict(100, 12, seed=0), 100, 12), 0.59, "spec"),
    "m5": (lambda: ap.M5Classifier(synthetic.m5_state_dict(seed=0)), 0.013, "wave"),
}


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"batch": B}
for name, (ctor, gflop, kind) in FAMILIES.items():
    clf = ctor()
    if name in ("vgg19_bn", "wideresnet28_10", "resnet34") and os.environ.get("AP_PROBE_TF32", "1") == "1":
        clf.set_mode("tf32")
    x = torch.randn(B, 1, 32, 32, device="cuda") if kind == "spec" else torch.randn(B, 1, 16000, device="cuda") * 0.1

    def fwd():
        with torch.no_grad():
            clf(x)

    def fwd_bwd():
        xr = x.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(clf(xr).sum(), xr)

    f, fb = timed(fwd), timed(fwd_bwd)
    out[name] = {"forward_ms": round(f, 2), "forward_backward_ms": round(fb, 2), "images_per_s": round(1e3 * B / f),
                 "forward_tflops": round(gflop * B / f, 1)}
    del clf
    torch.cuda.empty_cache()
print(json.dumps(out), flush=True)
