"""Golden vectors for the black-box query path from the UNMODIFIED reference:
``robustness_eval._NES.NES``, ``robustness_eval._EOT.EOT``, ``robustness_eval._utils.{resolve_loss, SEC4SR_MarginLoss}``
run on CPU around a small stochastic stand-in model (the estimator is model-agnostic; the defended system itself is pinned
by make_golden.py).  ``torch.randn`` is patched so the NES noise is host noise the CUDA path can be fed.

    python tests/golden/make_golden_blackbox.py      # in the build container; writes reference_golden_blackbox.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden import synthetic  # noqa: E402

K = 10


def toy_weights(L):
    return synthetic.host_noise((L, K), 777, 0) * np.float32(4.0 / np.sqrt(L))


class ToyModel(torch.nn.Module):
    """scores = tanh((x + 0.05 * e) W) * 4 with e popped from a fixed list: a randomised defence in miniature."""

    def __init__(self, L, seed=900):
        super().__init__()
        self.W = torch.from_numpy(toy_weights(L))
        self.seed, self.i, self.inputs = seed, 0, []

    def forward(self, x):
        self.inputs.append(x.detach().clone().numpy())
        e = torch.from_numpy(synthetic.host_noise(tuple(x.shape), self.seed, self.i))
        self.i += 1
        return torch.tanh((x + 0.05 * e)[:, 0, :] @ self.W) * 4


class RandnInjector:
    def __init__(self, seed):
        self.seed, self.i, self._orig = seed, 0, torch.randn

    def __enter__(self):
        def fake(size, device=None, **kw):
            z = synthetic.host_noise(tuple(size), self.seed, self.i)
            self.i += 1
            return torch.from_numpy(z)
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self._orig


def main():
    mg.install_shim()
    from robustness_eval._EOT import EOT
    from robustness_eval._NES import NES
    from robustness_eval._utils import SEC4SR_MarginLoss, resolve_loss
    out = {}

    # ---- per-query losses and their gradients
    scores = torch.from_numpy(synthetic.host_noise((64, K), 31, 0) * 3)
    scores[5, 3] = scores[5, 7] = scores[5].max() + 1          # a tie for the decision
    labels = torch.from_numpy(np.random.default_rng(5).integers(0, K, 64))
    out["loss_scores"], out["loss_labels"] = scores.numpy(), labels.numpy()
    ce, sign = resolve_loss("Margin", False, 0.5, "SCR", None, False)
    assert sign == 1 and resolve_loss("Entropy", True, 0.0, "SCR")[1] == -1
    s = scores.clone().requires_grad_(True)
    l = ce(s, labels)
    l.backward(torch.ones_like(l))
    out["loss_entropy"], out["loss_entropy_grad"] = l.detach().numpy(), s.grad.numpy()
    for targeted in (False, True):
        for clip in (False, True):
            m = SEC4SR_MarginLoss(targeted=targeted, confidence=0.5, task="CSI", clip_max=clip)
            s = scores.clone().requires_grad_(True)
            l = m(s, labels)
            l.backward(torch.ones_like(l))
            out[f"loss_margin_t{int(targeted)}_c{int(clip)}"] = l.detach().numpy()
            out[f"loss_margin_t{int(targeted)}_c{int(clip)}_grad"] = s.grad.numpy()
    out["loss_decision"] = scores.max(1, keepdim=True)[1][:, 0].numpy()

    # ---- NES around the toy model: (name, A, L, samples_per_draw, batch, sigma, EOT_size, EOT_batch)
    for name, A, L, spd, bs, sigma, es, eb in (("a", 2, 1024, 16, 8, 0.001, 1, 1), ("b", 3, 1001, 12, 12, 0.01, 4, 2)):
        x = torch.from_numpy(synthetic.synthetic_waveforms(A, L, seed=55))
        y = torch.tensor([3, 7, 1][:A])
        model = ToyModel(L)
        eot = EOT(model, ce, es, eb, False)
        with RandnInjector(4000) as inj, torch.no_grad():
            mean_loss, grad, adver_loss, adver_score, predict = NES(spd, bs, sigma, eot)(x, y)
            assert inj.i == spd // bs
        out[f"nes_{name}_cfg"] = np.array([A, L, spd, bs, es, eb], dtype=np.int64)
        out[f"nes_{name}_sigma"] = np.float32(sigma)
        out[f"nes_{name}_x"], out[f"nes_{name}_y"] = x.numpy(), y.numpy()
        out[f"nes_{name}_mean_loss"], out[f"nes_{name}_grad"] = mean_loss.numpy(), grad.numpy()
        out[f"nes_{name}_adver_loss"], out[f"nes_{name}_adver_score"] = adver_loss.numpy(), adver_score.numpy()
        out[f"nes_{name}_predict"] = np.asarray(predict)
        # the first query batch of every draw (the EOT copies repeat it) pins the perturbation kernel bit for bit
        per_draw = (es // eb)
        for i in range(spd // bs):
            q = model.inputs[i * per_draw]
            out[f"nes_{name}_queries{i}"] = q[: q.shape[0] // eb]

    path = os.path.join(HERE, "reference_golden_blackbox.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("nes_a")})


if __name__ == "__main__":
    main()
