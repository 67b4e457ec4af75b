"""Context measurement (not a parity test): the reference ALGORITHM dispatched to PyTorch's own CUDA libraries (cuDNN / cuBLAS /
cuFFT-free matmul mel) on the same B200 -- "the existing Blackwell implementation" of this path -- next to
the hand-written path, on the same batch.  The oracle's torch ops are simply run on CUDA tensors (TF32 allowed, as PyTorch
does by default for convolutions).  Asserts only that the hand-written path is faster and that both agree."""
import time

import numpy as np
import pytest
import torch

from gpu_common import CONFIG_JSON, cuda, rel_l2, synthetic

pytestmark = pytest.mark.gpu


def test_vs_pytorch_eager_on_the_same_gpu():
    import audiopure_b200 as ap
    import audiopure_oracle as orc

    B, L, t_star = 32, 16000, 2
    sd = synthetic.wavenet_state_dict(seed=0)
    rx_sd = synthetic.resnext_state_dict(seed=0)
    x = synthetic.synthetic_waveforms(B, L, seed=1234)
    zs = [synthetic.host_noise(x.shape, 2024, i) for i in range(t_star)]
    hp = orc.diffusion_hyperparams()

    # ---- oracle on CUDA: cache every weight on the device, keep the schedule scalars where the reference keeps them (CPU)
    cache = {}
    orig_t = orc._t

    def t_cuda(a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to("cuda", dtype) if a.numel() > 1 else a.to(dtype)
        key = (id(a), dtype)
        if key not in cache:
            cache[key] = torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)
        return cache[key]

    tf32_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    import torchaudio
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32, norm="slaney",
                                               pad_mode="constant", mel_scale="slaney").cuda()      # the reference's transform
    todb = torchaudio.transforms.AmplitudeToDB(stype="power").cuda()
    orig_se = orc.step_embedding
    orc._t = t_cuda
    orc.step_embedding = lambda steps, d=128: orig_se(steps.cpu(), d).cuda()
    try:
        def eager():
            with torch.no_grad():
                noise = orc.NoiseSource([torch.from_numpy(z).cuda() for z in zs])
                eps_fn = lambda xx, tt: orc.wavenet_forward(sd, xx, tt * torch.ones(xx.shape[0], 1, device="cuda"))
                y = orc.ddpm_forward(sd, torch.from_numpy(x).cuda(), {k: v for k, v in hp.items()}, t_star, noise, eps_fn=eps_fn)
                return y, orc.resnext_forward(rx_sd, todb(mel(y)))
        y_ref, logits_ref = eager()          # warm-up (cuDNN autotune, weight upload)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y_ref, logits_ref = eager()
        torch.cuda.synchronize()
        t_eager = time.perf_counter() - t0
    finally:
        orc._t, orc.step_embedding = orig_t, orig_se
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_flags     # later tests use fp32 torch ops

    # ---- hand-written path, same inputs and the same host noise
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=t_star, state_dict=sd, noise="torch")
    rx = ap.ResNeXtClassifier(rx_sd)
    tr = ap.sc09_transform()
    it = iter(zs)
    orig = torch.normal
    torch.normal = lambda mean, std, size=None, **kw: torch.from_numpy(next(it))
    try:
        def ours():
            nonlocal it
            it = iter(zs)
            y = dw(cuda(x))
            return y, rx(tr(y))
        y, logits = ours()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y, logits = ours()
        torch.cuda.synchronize()
        t_ours = time.perf_counter() - t0
    finally:
        torch.normal = orig
    err = rel_l2(y, y_ref)
    print(f"B={B}: PyTorch eager (cuDNN/cuBLAS, TF32) {B / t_eager:.1f} waveforms/s ({t_eager * 1e3:.0f} ms); hand-written bf16 path "
          f"{B / t_ours:.1f} waveforms/s ({t_ours * 1e3:.0f} ms); speed-up {t_eager / t_ours:.1f}x; purified rel-L2 {err:.2e}")
    assert err < 1e-2
    assert t_ours < t_eager
