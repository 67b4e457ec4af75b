"""Randomized-smoothing certification (robustness_eval/certified_robust.py:6-127) on the B200 kernels.

``RobustCertificate`` keeps the reference's methods (``forward``, ``smooth_predict``, ``certify``, ``compute_t_star``,
``lower_conf_bound``) and return values.  Differences, all on the device side:
  * the N noisy copies of one input are built by one fused kernel (``ap_smooth_inputs``: repeat + N(0, sigma) noise +
    sqrt(alpha_bar*) rescale, in-kernel Philox unless ``noise='torch'``);
  * argmax + per-class counting is one kernel accumulating into a device int64[K] (``ap_vote_counts``) instead of
    ``cat`` + ``max`` + K ``.item()`` syncs (certified_robust.py:59-67);
  * with ``torch.distributed`` initialised, the draws of ONE input are sharded over the ranks (rank r takes the r-th
    slice of every smooth_predict call, with its own Philox offset range) and the count vectors are combined by a single
    all-reduce per smooth_predict call.
"""
from __future__ import annotations

import math

import torch

from . import _lib

__all__ = ["RobustCertificate", "shard_draws", "reduce_counts", "certify_dataset"]


def shard_draws(n: int, world_size: int, rank: int):
    """[start, stop) of the draws rank ``rank`` takes out of ``n`` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank vote vectors in place: ONE all-reduce of int64[K] (NCCL over NVLink on the GPU box; the same call
    runs over gloo in the CPU tests).  This is the only collective on the certification path."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(counts, op=torch.distributed.ReduceOp.SUM, group=group)
    return counts


class RobustCertificate:
    def __init__(self, classifier, transform=None, denoiser=None, one_shot_rev: bool = False, num_classes=10,
                 noise: str = "philox", seed: int = 0, process_group=None, distributed: bool | None = None) -> None:
        self.classifier = classifier
        self.transform = transform
        self.denoiser = denoiser
        self.num_classes = num_classes
        self.one_shot_rev = one_shot_rev
        assert noise in ("philox", "torch")
        self.noise = noise
        self.seed = int(seed)
        self._offset = 0
        self.process_group = process_group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = distributed
        self._lib = _lib.load()

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        x_in = x
        if self.denoiser is not None:
            x_in = self.denoiser.one_shot_denoise(x_in)
        if self.transform is not None:
            x_in = self.transform(x_in)
        return self.classifier(x_in)

    def _world(self):
        if not self.distributed:
            return 1, 0
        return (torch.distributed.get_world_size(self.process_group), torch.distributed.get_rank(self.process_group))

    @torch.no_grad()
    def smooth_predict(self, x: torch.Tensor, num_sampling: int = 100, sigma=0.25, batch_size=64):
        """Class counts (CPU int64[K]) over ``num_sampling`` noisy copies of one input x (1, 1, L)."""
        assert (x.shape[0] == 1)
        if not x.is_cuda:
            x = x.cuda()
        x = x.detach().to(torch.float32).contiguous()
        L = x.shape[-1]
        world, rank = self._world()
        start, stop = shard_draws(num_sampling, world, rank)
        n_local = stop - start
        batches = [batch_size for _ in range(n_local // batch_size)]
        if n_local % batch_size:
            batches.append(n_local % batch_size)
        scale = 1.0
        if self.denoiser is not None:
            alpha_bar_star = 1 / (1 + sigma ** 2)
            self.denoiser.reverse_timestep = self.compute_t_star(alpha_bar_star)
            scale = alpha_bar_star ** 0.5
        counts = torch.zeros(self.num_classes, dtype=torch.int64, device=x.device)
        # Philox stream layout: this call owns offsets [base, base + ceil(num_sampling*L/4) + world); rank r starts at its
        # first draw, so the union over ranks is the same stream as a single-GPU run of the same seed.
        base = self._offset
        self._offset += (num_sampling * L + 3) // 4 + world
        done = start
        for batch in batches:
            x_in = torch.empty(batch, 1, L, device=x.device, dtype=torch.float32)
            if self.noise == "torch":
                delta = torch.normal(0, sigma, size=(batch, 1, L)).to(x.device)     # reference: CPU RNG, then H2D
                zp, sig = delta.data_ptr(), 1.0
            else:
                delta, zp, sig = None, None, float(sigma)
            assert L % 4 == 0 or self.noise == "torch", "in-kernel noise sharding needs L % 4 == 0"
            with torch.cuda.device(x.device):
                _lib.check(self._lib.ap_smooth_inputs(x.data_ptr(), sig, float(scale), zp, self.seed,
                                                      base + (done * L) // 4, x_in.data_ptr(), batch, L, _lib.stream_ptr()),
                           "ap_smooth_inputs")
            logits = self.forward(x_in)
            with torch.cuda.device(x.device):
                _lib.check(self._lib.ap_vote_counts(logits.data_ptr(), batch, logits.shape[-1], counts.data_ptr(),
                                                    _lib.stream_ptr()), "ap_vote_counts")
            done += batch
        if self.distributed and world > 1:
            reduce_counts(counts, self.process_group)
        return counts.cpu()

    @torch.no_grad()
    def certify(self, x: torch.Tensor, y: torch.Tensor, sigma: float = 0.25, n_0: int = 100, n: int = 100000,
                alpha: float = 0.001, batch_size: int = 64):
        from scipy.stats import norm
        y_pred, radius = -torch.ones_like(y), torch.zeros_like(y, dtype=torch.float32)
        for i in range(x.shape[0]):
            x_in = x[i]
            if x_in.ndim == 2:
                x_in = x_in.unsqueeze(0)
            counts_0 = self.smooth_predict(x_in, num_sampling=n_0, sigma=sigma, batch_size=batch_size)
            c_A = counts_0.max(0, keepdim=True)[1].item()
            counts = self.smooth_predict(x_in, num_sampling=n, sigma=sigma, batch_size=batch_size)
            pa = self.lower_conf_bound(k=int(counts[c_A]), n=n, alpha=alpha)
            if pa > 0.5:
                y_pred[i] = c_A
                radius[i] = sigma * norm.ppf(pa)
            else:
                y_pred[i] = -1
                radius[i] = 0
        return y_pred, radius

    def compute_t_star(self, alpha_bar_star):
        Alpha_bar = self.denoiser.diffusion_hyperparams["Alpha_bar"]
        return torch.abs(Alpha_bar - alpha_bar_star).min(0, keepdim=True)[1].item() + 1

    def lower_conf_bound(self, k, n, alpha=0.001):
        """Clopper-Pearson lower bound == statsmodels proportion_confint(k, n, alpha=2*alpha, method='beta')[0]."""
        from scipy.stats import beta
        if k <= 0:
            return 0.0
        p = float(beta.ppf(alpha, k, n - k + 1))
        return 0.0 if math.isnan(p) else p

    def certified_robust_correct(self, y_pred, y_target, r_c, r: float = 1.0):
        return sum(1 for i in range(len(y_pred)) if y_pred[i] == y_target[i] and r_c[i] >= r)


def certify_dataset(rc: RobustCertificate, batches, sigma: float, num_sampling: int = 100000, n_0: int = 100,
                    alpha: float = 0.001, batch_size: int = 512, save_path: str | None = None):
    """The dataset loop of certified_robustness_eval.py:110-146: certify every input of every batch and keep the
    reference's record format ``{'id', 'y_true', 'y_pred', 'certified_radius'}``; the JSON file
    ``<save_path>/sigma=<s>/sigma=<s>_N=<n>.json`` is rewritten after every batch (the reference's resume-by-inspection
    behaviour).  ``batches`` yields ``(waveforms (B,1,L) or (B,L), targets (B,))`` or dicts with 'samples' / 'target'.
    Under torch.distributed every rank takes part in every decision (sharded draws); rank 0 writes the file."""
    import json
    import os
    records, total = [], 0
    is_rank0 = not (torch.distributed.is_available() and torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0
    for batch in batches:
        waveforms, targets = (batch["samples"], batch["target"]) if isinstance(batch, dict) else batch
        if waveforms.ndim == 2:
            waveforms = torch.unsqueeze(waveforms, 1)
        waveforms, targets = waveforms.cuda(), targets.cuda()
        y_cert, r_cert = rc.certify(x=waveforms, y=targets, sigma=sigma, n_0=n_0, n=num_sampling, alpha=alpha,
                                    batch_size=batch_size)
        for i in range(waveforms.shape[0]):
            records.append({"id": i + total, "y_true": targets[i].item(), "y_pred": y_cert[i].item(),
                            "certified_radius": r_cert[i].item()})
        total += waveforms.shape[0]
        if save_path is not None and is_rank0:
            d = os.path.join(save_path, "sigma={}".format(sigma))
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "sigma={}_N={}.json".format(sigma, num_sampling)), "w") as f:
                json.dump(records, f, indent=4)
    return records
