"""Shared helpers of the -m gpu parity tests (everything goes through the C ABI via the audiopure_b200 host modules)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import audiopure_b200  # noqa: E402,F401
from audiopure_b200 import synthetic  # noqa: E402

CONFIG_JSON = os.path.join(ROOT, "diffusion-model-for-audio-defense_b200", "configs", "config.json")


def rel_l2(a, b):
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class TorchNormalInjector:
    """Replaces torch.normal(mean, std, size=...) by synthetic.host_noise tensors popped in call order -- the same
    injection tests/golden/make_golden.py applied to the reference."""

    def __init__(self, seed):
        self.seed, self.i = seed, 0
        self._orig = torch.normal

    def __enter__(self):
        def fake(mean, std, size=None, **kw):
            z = synthetic.host_noise(tuple(size), self.seed, self.i)
            self.i += 1
            return torch.from_numpy(z) * std + mean
        torch.normal = fake
        return self

    def __exit__(self, *a):
        torch.normal = self._orig


class ToyDefendedModel:
    """The stochastic stand-in model of tests/golden/make_golden_blackbox.py: scores = tanh((x + 0.05 e) W) * 4 with e
    popped from a fixed host-noise list.  Works on whatever device x lives on (torch ops: test scaffolding only)."""

    def __init__(self, L, K=10, seed=900):
        self.W = torch.from_numpy(synthetic.host_noise((L, K), 777, 0) * np.float32(4.0 / np.sqrt(L)))
        self.seed, self.i = seed, 0

    def __call__(self, x):
        e = torch.from_numpy(synthetic.host_noise(tuple(x.shape), self.seed, self.i)).to(x.device)
        self.i += 1
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False          # the stand-in must be an fp32 function on every device
        try:
            return torch.tanh((x + 0.05 * e)[:, 0, :] @ self.W.to(x.device)) * 4
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev


# ---- stand-ins for the reference's pickled classifier modules (create_model.py:8-16 unpickles WHOLE modules; the GPU box has
# no reference checkout, so these carry the class names, attributes and state_dict keys that loader dispatches on)
from types import SimpleNamespace as _NS  # noqa: E402


class _PickledStandIn(torch.nn.Module):
    def __init__(self, sd, **attrs):
        super().__init__()
        self._sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
        for k, v in attrs.items():
            object.__setattr__(self, k, v)

    def state_dict(self, *a, **k):
        return dict(self._sd)


class CifarResNeXt(_PickledStandIn):
    pass


class ResNet(_PickledStandIn):
    pass


class Bottleneck:          # type(model.layer1[0]).__name__ decides BasicBlock vs Bottleneck
    pass


class BasicBlock:
    pass


class VGG(_PickledStandIn):
    pass


class Conv2d:
    def __init__(self, in_channels=1):
        self.in_channels = in_channels


class WideResNet(_PickledStandIn):
    pass


class DenseNet(_PickledStandIn):
    pass


class M5(_PickledStandIn):
    pass


class DataParallelStandIn(torch.nn.Module):      # `.module` is unwrapped by the loader (create_model.py:12-13)
    def __init__(self, module):
        super().__init__()
        object.__setattr__(self, "module", module)


def pickled_classifier(kind):
    """(stand-in module, directly constructed classifier factory) for one create_model branch."""
    import audiopure_b200 as ap
    if kind == "resnext":
        sd = synthetic.resnext_state_dict(seed=0)
        return (CifarResNeXt(sd, nlabels=10, cardinality=8, depth=29, base_width=64, widen_factor=4, conv_1_3x3=_NS(in_channels=1)),
                lambda: ap.ResNeXtClassifier(sd))
    if kind == "resnet50":
        sd = synthetic.resnet_state_dict(depth=50, seed=0)
        layers = {f"layer{i + 1}": [Bottleneck()] * n for i, n in enumerate((3, 4, 6, 3))}
        return (ResNet(sd, fc=_NS(out_features=10), conv1=_NS(in_channels=1), **layers), lambda: ap.ResNetClassifier(sd, depth=50))
    if kind == "vgg19":
        sd = synthetic.vgg_state_dict(depth=19, seed=0)
        feats = [Conv2d(1) if v != "M" else None for v in synthetic.VGG_CFG[19]]
        return (VGG(sd, features=feats, classifier={6: _NS(out_features=10)}), lambda: ap.VGGClassifier(sd, depth=19))
    if kind == "wrn16_4":
        sd = synthetic.wideresnet_state_dict(depth=16, widen_factor=4, seed=0)
        return (WideResNet(sd, block1=_NS(layer=[0, 0]), nChannels=256, fc=_NS(out_features=10), conv1=_NS(in_channels=1)),
                lambda: ap.WideResNetClassifier(sd, depth=16, widen_factor=4))
    if kind == "densenet22":
        sd = synthetic.densenet_state_dict(depth=22, growth_rate=12, seed=0)
        return (DenseNet(sd, dense1=[0, 0, 0], growthRate=12, trans1=_NS(conv1=_NS(in_channels=60, out_channels=30)),
                         fc=_NS(out_features=10), conv1=_NS(in_channels=1)),
                lambda: ap.DenseNetClassifier(sd, depth=22, growth_rate=12))
    if kind == "m5":
        sd = synthetic.m5_state_dict(seed=0)
        return (M5(sd, conv1=_NS(kernel_size=(160,), stride=(16,), out_channels=32), fc1=_NS(out_features=10)),
                lambda: ap.M5Classifier(sd))
    raise KeyError(kind)
