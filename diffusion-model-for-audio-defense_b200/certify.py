"""Randomized-smoothing certification (robustness_eval/certified_robust.py:6-127) on the B200 kernels.

``RobustCertificate`` keeps the reference's methods (``forward``, ``smooth_predict``, ``certify``, ``compute_t_star``,
``lower_conf_bound``) and return values.  What differs is how the device is driven:

  * one micro-batch of noisy copies is ONE fused pipeline -- ``ap_diffwave_smooth_denoise`` (the copies
    ``sqrt(abar*) (x + sigma z)`` are built inside the network's first kernel from in-kernel Philox noise, and
    ``x0_hat = a x_in - b eps`` is formed in the epilogue of its last kernel) -> log-mel -> classifier -> ``ap_vote_counts``
    (argmax + per-class counting into a device int64[K], instead of ``cat`` + ``max`` + K ``.item()`` syncs,
    certified_robust.py:59-67);
  * that pipeline is captured once per (sigma, batch, L, mode) in a CUDA graph and replayed for every full micro-batch; the
    Philox offset lives in a device word the graph itself advances, every buffer is allocated once;
  * ``certify`` runs the inputs as a software pipeline: the n_0 and n passes of input i+1 are enqueued before the host looks at
    input i, whose vote vectors are combined by ONE all-reduce of int64[2K] and copied to pinned memory on a side stream, so the
    Clopper-Pearson statistics of input i (scipy, host) overlap the draws of input i+1 and the compute stream never waits;
  * with ``torch.distributed`` initialised, the draws of ONE input are sharded over the ranks (rank r takes the r-th slice of
    every smooth_predict call, with its own Philox offset range): the union over ranks is the single-GPU stream.
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


class _Pipeline:
    """Device-side state of the fused micro-batch: static buffers and the captured CUDA graph for one key."""

    def __init__(self):
        self.key = None
        self.graph = None
        self.x1 = None            # (1, 1, L) static copy of the input being certified
        self.x0 = None            # (batch, 1, L) noisy copies -> denoised copies
        self.offset_dev = None    # int64[1]: Philox block offset of the next micro-batch (advanced by the graph)
        self.votes = None         # int64[K] accumulated by the graph's vote kernel
        self.generation = None    # ap_alloc_generation() right after the capture: the graph holds raw workspace pointers


class RobustCertificate:
    def __init__(self, classifier, transform=None, denoiser=None, one_shot_rev: bool = False, num_classes=None,
                 noise: str = "philox", seed: int | None = None, process_group=None, distributed: bool | None = None,
                 use_graph: bool = True) -> None:
        """``num_classes`` is accepted for signature compatibility; like the reference (certified_robust.py:60-63) the count
        vector is sized from the classifier's output.  ``use_graph``: replay full micro-batches from a CUDA graph (Philox
        noise + audiopure_b200 modules only; anything else runs the same kernels eagerly)."""
        self.classifier = classifier
        self.transform = transform
        self.denoiser = denoiser
        self.num_classes = num_classes
        self.one_shot_rev = one_shot_rev
        assert noise in ("philox", "torch")
        self.noise = noise
        self.seed = _lib.philox_key("certify", seed)
        self._offset = 0
        self.process_group = process_group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.distributed = distributed
        self.use_graph = use_graph
        self._lib = _lib.load()
        self._pipe = _Pipeline()
        self._side = None           # side stream: all-reduce + device->host copy of finished vote vectors
        self._slots = None          # ring of 2 (device int64[2, K], pinned int64[2, K], event)

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

    # -- one micro-batch ------------------------------------------------------------------------------------------
    def _fusable(self) -> bool:
        """the fused front end needs our DiffWave as the denoiser and in-kernel (or injected) noise handled by the C side"""
        from .diffwave import DiffWave
        return isinstance(self.denoiser, DiffWave) and type(self).forward is RobustCertificate.forward

    def _x0_coefficients(self):
        ab = self.denoiser.diffusion_hyperparams["Alpha_bar"]
        t = self.denoiser.reverse_timestep - 1
        return float(t), float((1 / ab).sqrt()[t]), float((1 / ab - 1).sqrt()[t])

    def _logits_of_draws(self, x1, x0_buf, batch, sigma, scale, zp, seed, off, off_dev):
        """logits of `batch` noisy copies of x1 (1,1,L); fused front end when the denoiser is ours."""
        L = x1.shape[-1]
        if self._fusable():
            t, a, b = self._x0_coefficients()
            with torch.cuda.device(x1.device):
                _lib.check(self._lib.ap_diffwave_smooth_denoise(self.denoiser.model._handle, x1.data_ptr(), float(sigma),
                                                                float(scale), zp, seed, off, off_dev, t, a, b,
                                                                x0_buf.data_ptr(), batch, L, _lib.stream_ptr()),
                           "ap_diffwave_smooth_denoise")
            x_in = x0_buf[:batch]
            if self.transform is not None:
                x_in = self.transform(x_in)
            return self.classifier(x_in)
        with torch.cuda.device(x1.device):
            _lib.check(self._lib.ap_smooth_inputs(x1.data_ptr(), float(sigma), float(scale), zp, seed, off, x0_buf.data_ptr(),
                                                  batch, L, _lib.stream_ptr()), "ap_smooth_inputs")
        return self.forward(x0_buf[:batch])

    def _vote(self, logits, p):
        """accumulate the argmax histogram of `logits` into p.votes (sized from the classifier's output on first use, like
        ``counts = torch.zeros(output.shape[-1])`` at certified_robust.py:60-63)"""
        if p.votes is None or p.votes.numel() != logits.shape[-1]:
            p.votes = torch.zeros(logits.shape[-1], dtype=torch.int64, device=logits.device)
        votes = p.votes
        with torch.cuda.device(logits.device):
            _lib.check(self._lib.ap_vote_counts(logits.data_ptr(), logits.shape[0], logits.shape[-1], votes.data_ptr(),
                                                votes.numel(), _lib.stream_ptr()), "ap_vote_counts")

    def _prepare(self, x, batch, sigma, scale):
        """(re)build the static buffers -- and, when enabled, the CUDA graph -- for this (input shape, batch, sigma, t*, mode)."""
        p = self._pipe
        L = x.shape[-1]
        mode = getattr(getattr(self.denoiser, "model", None), "mode", None)
        t_star = getattr(self.denoiser, "reverse_timestep", None)
        key = (x.device, L, batch, float(sigma), float(scale), t_star, mode, self.noise)
        if p.key == key and (p.graph is None or p.generation == self._lib.ap_alloc_generation()):
            return p            # (a library workspace that moved since the capture -- another batch size, length or mode through
                                #  the same handles -- invalidates the graph: fall through and capture again)
        p.key, p.graph, p.votes = key, None, None
        p.x1 = torch.empty(1, 1, L, device=x.device, dtype=torch.float32)
        p.x0 = torch.empty(batch, 1, L, device=x.device, dtype=torch.float32)
        p.offset_dev = torch.zeros(1, dtype=torch.int64, device=x.device)
        p.x1.copy_(x.reshape(1, 1, L))
        if self.use_graph and self.noise == "philox" and self._fusable() and L % 4 == 0 and mode != "fp32":
            # warm-up micro-batch (votes discarded): sizes every library workspace before the capture, and fixes K
            self._vote(self._logits_of_draws(p.x1, p.x0, batch, sigma, scale, None, self.seed, 0, None), p)
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                lg = self._logits_of_draws(p.x1, p.x0, batch, sigma, scale, None, self.seed, 0, p.offset_dev.data_ptr())
                self._vote(lg, p)
                _lib.check(self._lib.ap_u64_add(p.offset_dev.data_ptr(), (batch * L) // 4, _lib.stream_ptr()), "ap_u64_add")
            p.graph = g
            p.generation = self._lib.ap_alloc_generation()
        return p

    def _enqueue_counts(self, x, num_sampling, sigma, batch_size):
        """Enqueue the local share of ``num_sampling`` draws of x (1,1,L); returns the pipeline's vote vector (device int64[K],
        reused by the next call: copy it).  No host synchronisation."""
        L = x.shape[-1]
        world, rank = self._world()
        start, stop = shard_draws(num_sampling, world, rank)
        n_local = stop - start
        scale = 1.0
        if self.denoiser is not None:
            alpha_bar_star = 1 / (1 + sigma ** 2)
            self.denoiser.reverse_timestep = self.compute_t_star(alpha_bar_star)
            scale = alpha_bar_star ** 0.5
        # Philox stream layout: this call owns offsets [base, base + ceil(num_sampling*L/4) + world); rank r starts at its
        # first draw, so the union over ranks is the same stream as a single-GPU run of the same seed.
        base = self._offset
        self._offset += (num_sampling * L + 3) // 4 + world
        assert L % 4 == 0 or self.noise == "torch", "in-kernel noise sharding needs L % 4 == 0"
        p = self._prepare(x, batch_size, sigma, scale)
        p.x1.copy_(x.reshape(1, 1, L))
        if p.votes is not None:
            p.votes.zero_()
        done = start
        n_full = n_local // batch_size
        if p.graph is not None and n_full:
            p.offset_dev.fill_(base + (done * L) // 4)
            for _ in range(n_full):
                p.graph.replay()
            done += n_full * batch_size
        else:
            for _ in range(n_full):
                self._eager_batch(p, batch_size, sigma, scale, base + (done * L) // 4)
                done += batch_size
        if n_local % batch_size:
            self._eager_batch(p, n_local % batch_size, sigma, scale, base + (done * L) // 4)
        if p.votes is None:      # this rank drew nothing (more ranks than draws): K from the classifier, or one probe draw
            K = getattr(self.classifier, "num_classes", None) or self.num_classes
            if K is None:
                self._eager_batch(p, 1, sigma, scale, base)
                p.votes.zero_()
            else:
                p.votes = torch.zeros(int(K), dtype=torch.int64, device=x.device)
        return p.votes

    def _eager_batch(self, p, batch, sigma, scale, off):
        L = p.x1.shape[-1]
        if self.noise == "torch":
            delta = torch.normal(0, sigma, size=(batch, 1, L)).to(p.x1.device)     # reference: CPU RNG, then H2D
            zp, sig = delta.data_ptr(), 1.0
        else:
            delta, zp, sig = None, None, float(sigma)
        logits = self._logits_of_draws(p.x1, p.x0, batch, sig, scale, zp, self.seed, off, None)
        self._vote(logits, p)

    @staticmethod
    def _as_input(x):
        if not x.is_cuda:
            x = x.cuda()
        return x.detach().to(torch.float32).contiguous()

    @torch.no_grad()
    def smooth_predict(self, x: torch.Tensor, num_sampling: int = 100, sigma=0.25, batch_size=64):
        """Class counts (CPU int64[K]) over ``num_sampling`` noisy copies of one input x (1, 1, L)."""
        assert (x.shape[0] == 1)
        x = self._as_input(x)
        world, _ = self._world()
        counts = self._enqueue_counts(x, num_sampling, sigma, batch_size).clone()
        if self.distributed and world > 1:
            reduce_counts(counts, self.process_group)
        return counts.cpu()

    @torch.no_grad()
    def certify(self, x: torch.Tensor, y: torch.Tensor, sigma: float = 0.25, n_0: int = 100, n: int = 100000,
                alpha: float = 0.001, batch_size: int = 64):
        """certified_robust.py:69-100, pipelined: input i+1's draws are enqueued before input i's statistics are read."""
        from scipy.stats import norm
        world, _ = self._world()
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        preds, radii = [], []

        def finalize(slot):
            _, host_counts, ev = slot
            ev.synchronize()                                    # the side stream's copy of this input has landed
            counts_0, counts = host_counts[0], host_counts[1]
            c_A = counts_0.max(0, keepdim=True)[1].item()
            pa = self.lower_conf_bound(k=int(counts[c_A]), n=n, alpha=alpha)
            if pa > 0.5:
                preds.append(c_A), radii.append(sigma * norm.ppf(pa))
            else:
                preds.append(-1), radii.append(0.0)

        pending = None
        for i in range(x.shape[0]):
            x_in = x[i]
            if x_in.ndim == 2:
                x_in = x_in.unsqueeze(0)
            x_in = self._as_input(x_in)
            v0 = self._enqueue_counts(x_in, n_0, sigma, batch_size)
            K = v0.numel()
            if self._slots is None or self._slots[0][0].shape[1] != K or self._slots[0][0].device != v0.device:
                self._slots = [(torch.zeros(2, K, dtype=torch.int64, device=v0.device),
                                torch.zeros(2, K, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(2)]
            slot = self._slots[i % 2]      # free: finalize(i - 2) has returned
            slot[0][0].copy_(v0)
            slot[0][1].copy_(self._enqueue_counts(x_in, n, sigma, batch_size))
            # side stream: ONE all-reduce of both vote vectors + copy to pinned memory; the compute stream moves on to input i+1
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                if self.distributed and world > 1:
                    reduce_counts(slot[0], self.process_group)
                slot[1].copy_(slot[0], non_blocking=True)
                slot[2].record(self._side)
            if pending is not None:
                finalize(pending)          # host statistics of input i-1 while the GPU draws for input i
            pending = slot
        if pending is not None:
            finalize(pending)
        y_pred = torch.tensor(preds, dtype=y.dtype, device=y.device) if preds else -torch.ones_like(y)
        radius = torch.tensor(radii, dtype=torch.float32, device=y.device) if radii else torch.zeros_like(y, dtype=torch.float32)
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
        y_host, r_host, t_host = y_cert.tolist(), r_cert.tolist(), targets.tolist()     # one read-back per tensor, not per item
        for i in range(waveforms.shape[0]):
            records.append({"id": i + total, "y_true": t_host[i], "y_pred": y_host[i], "certified_radius": r_host[i]})
        total += waveforms.shape[0]
        if save_path is not None and is_rank0:
            d = os.path.join(save_path, "sigma={}".format(sigma))
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "sigma={}_N={}.json".format(sigma, num_sampling)), "w") as f:
                json.dump(records, f, indent=4)
    return records
