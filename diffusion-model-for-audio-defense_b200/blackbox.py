"""Black-box query serving: the pieces FAKEBOB runs around the defended system every iteration.

``EOT`` (robustness_eval/_EOT.py:4-69), ``NES`` (robustness_eval/_NES.py:5-56), ``resolve_loss`` /
``resolve_prediction`` (robustness_eval/_utils.py:103-136) with the reference's constructor arguments and return values,
so ``black_box_attack.py:186-190`` can import them from here.  The model they query is any ``(B,1,L) -> (B,K)`` callable --
normally ``AcousticSystem`` over the CUDA purifier and classifier of this package.

What runs on the GPU: the query batch of a NES draw is written by ``ap_nes_perturb`` (antithetic Philox noise generated in
the kernel), the per-query loss and decision by ``ap_query_loss``, and the gradient estimate by ``ap_nes_gradient``, which
regenerates the noise from its counters instead of reading a stored copy.  ``noise='torch'`` draws ``torch.randn`` on the
device like the reference (the parity tests inject a fixed tensor through it).
"""
from __future__ import annotations

from collections import Counter

import numpy as np
import torch

from . import _lib
from ._lib import AudioPureError

__all__ = ["QueryLoss", "resolve_loss", "resolve_prediction", "EOT", "NES"]

_KIND = {"Entropy": 0, "Margin": 1}


class _QueryLossVJP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, mod, label):
        s = scores.detach().to(torch.float32).contiguous()
        ctx.mod, ctx.label = mod, label
        ctx.save_for_backward(s)
        return mod._run(s, label)[0]

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        mod = ctx.mod
        g = g.detach().to(torch.float32).contiguous()
        out = torch.empty_like(s)
        with torch.cuda.device(s.device):
            _lib.check(mod._lib.ap_query_loss_vjp(s.data_ptr(), ctx.label.data_ptr(), g.data_ptr(), s.shape[0], s.shape[1],
                                                  mod.kind, int(mod.targeted), mod.confidence, int(mod.clip_max),
                                                  out.data_ptr(), _lib.stream_ptr()), "ap_query_loss_vjp")
        return out, None, None


class QueryLoss(torch.nn.Module):
    """Per-query loss, ``(B,K) scores, (B,) labels -> (B,)``: ``nn.CrossEntropyLoss(reduction='none')`` (``'Entropy'``,
    what ``resolve_loss`` returns for the speech-command task, _utils.py:116-117) or the closed-set margin loss
    ``score_real + confidence - max_other`` (``'Margin'``, _utils.py:73-84,98-99)."""

    def __init__(self, loss_name: str = "Entropy", targeted: bool = False, confidence: float = 0.0, clip_max: bool = True):
        super().__init__()
        if loss_name not in _KIND:
            raise AssertionError(f"loss_name must be one of {sorted(_KIND)}")
        self.kind = _KIND[loss_name]
        self.targeted, self.confidence, self.clip_max = bool(targeted), float(confidence), bool(clip_max)
        self._lib = _lib.load()

    def _run(self, scores: torch.Tensor, label: torch.Tensor, want_pred: bool = False):
        B, K = scores.shape
        loss = torch.empty(B, device=scores.device, dtype=torch.float32)
        pred = torch.empty(B, device=scores.device, dtype=torch.int32) if want_pred else None
        with torch.cuda.device(scores.device):
            _lib.check(self._lib.ap_query_loss(scores.data_ptr(), label.data_ptr(), B, K, self.kind, int(self.targeted),
                                               self.confidence, int(self.clip_max), loss.data_ptr(),
                                               pred.data_ptr() if want_pred else None, _lib.stream_ptr()), "ap_query_loss")
        return loss, pred

    @staticmethod
    def _prepare(scores, label):
        if not scores.is_cuda:
            raise AudioPureError("QueryLoss: scores must be a CUDA tensor (there is no CPU path)")
        if scores.ndim != 2 or label.shape != scores.shape[:1]:
            raise AssertionError(f"QueryLoss: expected (B,K) scores and (B,) labels, got {tuple(scores.shape)}, {tuple(label.shape)}")
        return label.to(device=scores.device, dtype=torch.int64).contiguous()

    def forward(self, scores: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        label = self._prepare(scores, label)
        if scores.requires_grad and torch.is_grad_enabled():
            return _QueryLossVJP.apply(scores, self, label)
        return self._run(scores.detach().to(torch.float32).contiguous(), label)[0]

    def loss_and_decision(self, scores: torch.Tensor, label: torch.Tensor):
        """(loss (B,), argmax decision (B,) int32) in one launch; no autograd."""
        label = self._prepare(scores, label)
        return self._run(scores.detach().to(torch.float32).contiguous(), label, want_pred=True)


def resolve_loss(loss_name="Entropy", targeted=False, confidence=0.0, task="SCR", threshold=None, clip_max=True):
    """_utils.py:103-125: for the speech-command task the reference returns cross entropy whatever ``loss_name`` says;
    ``grad_sign`` is -1 for a targeted attack.  ('SV' raises NotImplementedError there as well.)"""
    assert loss_name in ["Entropy", "Margin"]
    assert task in ["SCR", "SV"]
    if task != "SCR":
        raise NotImplementedError(f"unsupported task yet: {task}!")
    return QueryLoss("Entropy"), (-1 if targeted else 1)


def resolve_prediction(decisions):
    """_utils.py:127-136: majority decision per query over the EOT draws (first seen wins a tie)."""
    return np.array([Counter(d).most_common(1)[0][0] for d in decisions])


class EOT(torch.nn.Module):
    """Expectation over the defence's randomness: ``EOT_size`` evaluations in batches of ``EOT_batch_size`` copies
    (_EOT.py:4-69).  Returns ``(scores, loss, grad, decisions)`` averaged over the copies; ``grad`` is None unless
    ``use_grad``, in which case ``x_batch`` must be part of an autograd graph (the reference calls ``retain_grad`` on the
    repeated batch, :35-36)."""

    def __init__(self, model, loss, EOT_size=1, EOT_batch_size=1, use_grad=True):
        super().__init__()
        self.model = model
        self.loss = loss
        self.EOT_size = EOT_size
        self.EOT_batch_size = EOT_batch_size
        self.EOT_num_batches = self.EOT_size // self.EOT_batch_size
        self.use_grad = use_grad

    def _run(self, x_batch, y_batch, EOT_size=None, EOT_batch_size=None, use_grad=None):
        """forward() with the decisions left on the device: (scores, loss, grad, decisions (EOT draws, n) int32)."""
        EOT_size = EOT_size if EOT_size else self.EOT_size
        EOT_batch_size = EOT_batch_size if EOT_batch_size else self.EOT_batch_size
        num_batches = EOT_size // EOT_batch_size
        use_grad = use_grad if use_grad else self.use_grad      # _EOT.py:22 (a False argument defers to the attribute)
        n, n_channels, max_len = x_batch.size()
        fused = isinstance(self.loss, QueryLoss) and not use_grad
        scores = loss = grad = None
        decisions = []
        for _ in range(num_batches):
            x_rep = x_batch.repeat(EOT_batch_size, 1, 1) if EOT_batch_size > 1 else x_batch
            y_rep = y_batch.repeat(EOT_batch_size) if EOT_batch_size > 1 else y_batch
            if use_grad:
                if x_rep is x_batch:
                    x_rep = x_batch.repeat(1, 1, 1)
                x_rep.retain_grad()
            s = self.model(x_rep)
            if fused:
                l, d = self.loss.loss_and_decision(s, y_rep)
            else:
                l = self.loss(s, y_rep)
                d = s.max(1)[1].to(torch.int32)
            if use_grad:
                l.backward(torch.ones_like(l))
                g = x_rep.grad.view(EOT_batch_size, -1, n_channels, max_len).mean(0)
                grad = g if grad is None else grad + g
            s_mean = s.detach().view(EOT_batch_size, -1, s.shape[1]).mean(0)
            l_mean = l.detach().view(EOT_batch_size, -1).mean(0)
            scores = s_mean if scores is None else scores + s_mean
            loss = l_mean if loss is None else loss + l_mean
            decisions.append(d.view(EOT_batch_size, n))
        scores = scores / num_batches
        loss = loss / num_batches
        if grad is not None:
            grad = grad / num_batches
        return scores, loss, grad, torch.cat(decisions, 0)

    def forward(self, x_batch, y_batch, EOT_size=None, EOT_batch_size=None, use_grad=None):
        scores, loss, grad, dec = self._run(x_batch, y_batch, EOT_size, EOT_batch_size, use_grad)
        dec = dec.cpu().numpy()                                  # one read-back instead of one per EOT batch
        return scores, loss, grad, [list(dec[:, i]) for i in range(dec.shape[1])]


class _PhiloxStream:
    """Process-wide counter for the in-kernel noise: FAKEBOB builds a fresh ``NES`` per iteration
    (black_box_attack.py:187), so the offset cannot live on the NES object."""

    def __init__(self, seed: int | None = None):
        self.seed, self.offset = _lib.philox_key("nes", seed), 0

    def take(self, blocks: int) -> int:
        off = self.offset
        self.offset += int(blocks)
        return off


PHILOX = _PhiloxStream()


class NES(torch.nn.Module):
    """Natural-evolution-strategy gradient estimate from ``samples_per_draw`` antithetic queries (_NES.py:5-56).

    ``forward(x (A,1,L), y (A,)) -> (mean_loss (A,), grad (A,1,L), adver_loss (A,), adver_score (A,K), predict (A,))``
    """

    def __init__(self, samples_per_draw, samples_per_draw_batch, sigma, EOT_wrapper, noise: str = "philox",
                 stream: _PhiloxStream | None = None):
        super().__init__()
        self.samples_per_draw = samples_per_draw
        self.samples_per_draw_batch_size = samples_per_draw_batch
        self.sigma = sigma
        self.EOT_wrapper = EOT_wrapper
        assert noise in ("philox", "torch")
        self.noise = noise
        self.stream = stream if stream is not None else PHILOX
        self._lib = _lib.load()

    def _draw(self, A, H, C, N, device):
        """(z tensor or None, seed, offset) for one batch of H antithetic pairs per audio."""
        if self.noise == "torch":
            return torch.randn([A, H, C, N], device=device), 0, 0          # _NES.py:18
        return None, self.stream.seed, self.stream.take(self._lib.ap_nes_noise_blocks(A, 2 * H, N))

    @torch.no_grad()
    def forward(self, x, y):
        if not x.is_cuda:
            raise AudioPureError("NES: input must be a CUDA tensor (there is no CPU path)")
        A, C, N = x.shape
        assert C == 1, "Only Support Mono Audio"
        S = int(self.samples_per_draw_batch_size)
        if S < 2 or S % 2:
            raise AssertionError("samples_per_draw_batch_size must be even (antithetic pairs, _NES.py:18-20)")
        num_batches = self.samples_per_draw // S
        x32 = x.detach().to(torch.float32).contiguous()
        y = y.to(x.device)
        eot = self.EOT_wrapper
        eot_batches = int(eot.EOT_size // eot.EOT_batch_size)
        grad = torch.empty_like(x32)
        scale = 1.0 / (S * float(self.sigma) * num_batches)               # mean over S, / sigma / num_batches (:47,53)
        st = _lib.stream_ptr
        mean_loss = adver_loss = adver_score = predict = None
        for i in range(num_batches):
            first = int(i == 0)
            R = S + first
            z, seed, off = self._draw(A, S // 2, C, N, x.device)
            zp = z.data_ptr() if z is not None else None
            eval_input = torch.empty(A * R, C, N, device=x.device, dtype=torch.float32)
            with torch.cuda.device(x.device):
                _lib.check(self._lib.ap_nes_perturb(x32.data_ptr(), float(self.sigma), zp, seed, off, first,
                                                    eval_input.data_ptr(), A, S, N, st()), "ap_nes_perturb")
            eval_y = y.repeat_interleave(R)
            scores, loss, _, dec = eot._run(eval_input, eval_y)
            loss = (loss / eot_batches).view(A, R).contiguous()            # the reference divides a second time (:35-36)
            if first:
                scores = (scores / eot_batches).view(A, R, -1)
                adver_loss, adver_score = loss[:, 0], scores[:, 0, :]
                clean = dec.view(dec.shape[0], A, R)[:, :, 0].cpu().numpy()   # EOT decisions of the un-noised queries
                predict = resolve_prediction([list(clean[:, a]) for a in range(A)])
            with torch.cuda.device(x.device):
                _lib.check(self._lib.ap_nes_gradient(loss.data_ptr(), zp, seed, off, first, scale, int(i > 0),
                                                     grad.data_ptr(), A, S, N, st()), "ap_nes_gradient")
            m = loss[:, first:].mean(1)
            mean_loss = m if mean_loss is None else mean_loss + m
        mean_loss = mean_loss / num_batches
        return mean_loss, grad.view(A, C, N), adver_loss, adver_score, predict
