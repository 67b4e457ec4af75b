"""BASELINE LEG, NOT PRODUCT: the reference ALGORITHM dispatched to PyTorch's own CUDA libraries on the same GPU.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (like the rest of ``oracle/``): used by ``tests/test_gpu_eager_baseline.py`` and by
``bench.py``'s ``eager_gpu`` baseline leg.  It runs the oracle's plain torch ops (``audiopure_oracle``) on CUDA tensors, i.e. the
reference's formulation -- weight norm re-folded on every call like the reference's pre-forward hook, 3 cuDNN convolutions and
~13 kernels per residual block, fp32 activations, ``skip += `` read-modify-write per layer, torchaudio mel -- through
cuDNN / cuBLAS / cuFFT.  This is "the existing implementation on Blackwell" that the hand-written path is to beat
(BASELINE.md section 3).  It reads nothing under /root/reference.
"""
from __future__ import annotations

import contextlib
import time

import numpy as np
import torch

import audiopure_oracle as orc


@contextlib.contextmanager
def oracle_on_cuda(allow_tf32: bool):
    """Within the block the oracle's ``_t`` uploads (and caches) every weight to the GPU; TF32 as requested."""
    cache = {}
    orig_t, orig_se = orc._t, orc.step_embedding

    def t_cuda(a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to("cuda", dtype) if a.numel() > 1 else a.to(dtype)
        key = (id(a), dtype)
        if key not in cache:
            cache[key] = torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)
        return cache[key]

    flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    orc._t = t_cuda
    orc.step_embedding = lambda steps, d=128: orig_se(steps.cpu(), d).cuda()
    try:
        yield
    finally:
        orc._t, orc.step_embedding = orig_t, orig_se
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = flags


def make_pipeline(sd, rx_sd, t_star: int = 2):
    """fn(x (B,1,L) numpy, noise list of numpy) -> (purified, logits): DDPM t* -> torchaudio SC09 mel -> ResNeXt, all eager CUDA."""
    import torchaudio
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32, norm="slaney",
                                               pad_mode="constant", mel_scale="slaney").cuda()      # the reference's transform
    todb = torchaudio.transforms.AmplitudeToDB(stype="power").cuda()
    hp = orc.diffusion_hyperparams()

    def run(x, zs):
        with torch.no_grad():
            noise = orc.NoiseSource([z if isinstance(z, torch.Tensor) else torch.from_numpy(z).cuda() for z in zs])
            eps_fn = lambda xx, tt: orc.wavenet_forward(sd, xx, tt * torch.ones(xx.shape[0], 1, device="cuda"))
            xd = x if isinstance(x, torch.Tensor) else torch.from_numpy(x).cuda()
            y = orc.ddpm_forward(sd, xd, dict(hp), t_star, noise, eps_fn=eps_fn)
            return y, orc.resnext_forward(rx_sd, todb(mel(y)))
    return run


def time_pipeline(sd, rx_sd, x, zs, allow_tf32: bool, t_star: int = 2, warmup: int = 1, reps: int = 3):
    """median seconds per pass over `reps` timed passes after `warmup` (cuDNN autotune, weight upload); also returns outputs"""
    with oracle_on_cuda(allow_tf32):
        run = make_pipeline(sd, rx_sd, t_star)
        xd = torch.from_numpy(x).cuda()
        zd = [torch.from_numpy(z).cuda() for z in zs]
        for _ in range(warmup):
            out = run(xd, zd)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = run(xd, zd)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], out
