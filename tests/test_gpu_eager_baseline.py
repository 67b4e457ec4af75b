"""Context measurement (not a parity test): the reference ALGORITHM dispatched to PyTorch's own CUDA libraries (cuDNN / cuBLAS /
torchaudio mel) on the same B200 -- "the existing Blackwell implementation" of this path -- next to the
hand-written path, on the same batch (oracle/eager_gpu.py; bench.py reports the same leg as ``eager_gpu``).  Asserts only that the
hand-written path is faster and that both agree."""
import time

import pytest
import torch

from gpu_common import CONFIG_JSON, cuda, rel_l2, synthetic

pytestmark = pytest.mark.gpu


def test_vs_pytorch_eager_on_the_same_gpu():
    import audiopure_b200 as ap
    import eager_gpu

    B, L, t_star = 32, 16000, 2
    sd = synthetic.wavenet_state_dict(seed=0)
    rx_sd = synthetic.resnext_state_dict(seed=0)
    x = synthetic.synthetic_waveforms(B, L, seed=1234)
    zs = [synthetic.host_noise(x.shape, 2024, i) for i in range(t_star)]
    t_eager, (y_ref, logits_ref) = eager_gpu.time_pipeline(sd, rx_sd, x, zs, allow_tf32=True, t_star=t_star, warmup=1, reps=1)

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
