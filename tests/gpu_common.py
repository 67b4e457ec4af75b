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
        return torch.tanh((x + 0.05 * e)[:, 0, :] @ self.W.to(x.device)) * 4
