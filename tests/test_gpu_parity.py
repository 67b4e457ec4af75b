"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the C ABI, against
  * the committed outputs of the unmodified reference (tests/golden/reference_golden.npz), and
  * the CPU oracle (oracle/audiopure_oracle.py) on the same seeded inputs.
Tolerances (BASELINE.json north_star): purified waveform rel-L2 <= 1e-5 in fp32 mode, <= 1e-2 in bf16 mode; top-1 exact.
"""
import numpy as np
import pytest
import torch

from gpu_common import (CONFIG_JSON, DataParallelStandIn, TorchNormalInjector, ToyDefendedModel, cuda, pickled_classifier, rel_l2,
                        synthetic)

pytestmark = pytest.mark.gpu

# north_star: purified waveform within 1e-2 (bf16) / 1e-5 (fp32 mode) relative L2 of the reference.  bf16x3 is the
# fp32-class tensor-core mode (hi/lo bf16 planes, 3 MMAs per product): it is held to the fp32 bar on the waveform.
TOL = {"fp32": 1e-5, "bf16": 1e-2, "fp16": 1e-3, "bf16x3": 1e-5}
TOL_EPS = {"fp32": 2e-5, "bf16": 1e-2, "fp16": 2e-3, "bf16x3": 1e-4}
# one-/two-shot x0 at t* = 66 weight eps by 0.5 (a DDPM step by ~0.012), so the 16-significand-bit eps error of bf16x3
# (2.3e-5..3e-5) shows as ~1.1e-5 there and as 5e-7..9e-7 on the purified waveform
TOL_X0 = dict(TOL, bf16x3=2e-5)


@pytest.fixture(scope="module")
def ap():
    import audiopure_b200 as m
    return m


@pytest.fixture(scope="module")
def orc():
    import audiopure_oracle
    return audiopure_oracle


@pytest.fixture(scope="module")
def sd_full():
    return synthetic.wavenet_state_dict(seed=0)


@pytest.fixture(scope="module")
def diffwave(ap, sd_full):
    return ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=2, state_dict=sd_full, noise="torch")


# ---------------------------------------------------------------------------------------------------- building blocks
@pytest.mark.parametrize("K", [64, 256, 768])
def test_umma_selftest(ap, K):
    """one 128x256xK tile through TMA (SWIZZLE_128B) -> tcgen05.mma -> TMEM -> tcgen05.ld"""
    from audiopure_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(K)
    a = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(256, K, generator=g).to(torch.bfloat16).cuda()
    d = torch.zeros(128, 256, device="cuda")
    _lib.check(lib.ap_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), K, _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    assert rel_l2(d, ref) < 1e-5


def test_update_kernels_match_reference_arithmetic(ap):
    from audiopure_b200 import _lib
    lib = _lib.load()
    for B, L in [(3, 1000), (2, 1023), (1, 5)]:       # vectorised, ragged, tiny
        g = torch.Generator().manual_seed(B * L)
        x = torch.randn(B, L, generator=g).cuda()
        e = torch.randn(B, L, generator=g).cuda()
        z = torch.randn(B, L, generator=g).cuda()
        st = _lib.stream_ptr()
        # expected values on the CPU: torch's CUDA true-division by a CPU scalar multiplies by the reciprocal (1 ulp off
        # IEEE division); the kernels follow the IEEE arithmetic the CPU reference (and the golden vectors) use.
        xc, ec, zc = x.cpu(), e.cpu(), z.cpu()
        a, b = float(np.float32(0.9997)), float(np.float32(0.0245))
        out = torch.empty_like(x)
        _lib.check(lib.ap_diffuse(x.data_ptr(), a, b, z.data_ptr(), 0, 0, out.data_ptr(), B, L, st))
        assert torch.equal(out.cpu(), torch.tensor(a) * xc + torch.tensor(b) * zc)
        c, sa, sg = float(np.float32(0.0115)), float(np.float32(0.9999)), float(np.float32(0.0082))
        y = x.clone()
        _lib.check(lib.ap_ddpm_step(y.data_ptr(), e.data_ptr(), c, sa, sg, z.data_ptr(), 0, 0, B, L, st))
        assert torch.equal(y.cpu(), (xc - torch.tensor(c) * ec) / torch.tensor(sa) + torch.tensor(sg) * zc)
        y = x.clone()
        _lib.check(lib.ap_ddpm_step(y.data_ptr(), e.data_ptr(), c, sa, 0.0, None, 0, 0, B, L, st))
        assert torch.equal(y.cpu(), (xc - torch.tensor(c) * ec) / torch.tensor(sa))
        _lib.check(lib.ap_predict_x0(x.data_ptr(), e.data_ptr(), float(np.float32(1.118)), 0.5, out.data_ptr(), B, L, st))
        assert torch.equal(out.cpu(), torch.tensor(np.float32(1.118)) * xc - torch.tensor(np.float32(0.5)) * ec)
        x1 = torch.randn(L, generator=g).cuda()
        so = torch.empty(B, L, device="cuda")
        _lib.check(lib.ap_smooth_inputs(x1.data_ptr(), 1.0, float(np.float32(0.8944)), z.data_ptr(), 0, 0, so.data_ptr(), B, L, st))
        assert torch.equal(so.cpu(), torch.tensor(np.float32(0.8944)) * (x1.cpu()[None] + zc))


def test_philox_noise_statistics_and_determinism(ap):
    from audiopure_b200 import _lib
    lib = _lib.load()
    n = 1 << 22
    a = torch.empty(n, device="cuda")
    b = torch.empty(n, device="cuda")
    st = _lib.stream_ptr()
    _lib.check(lib.ap_randn(a.data_ptr(), n, 1234, 0, st))
    _lib.check(lib.ap_randn(b.data_ptr(), n, 1234, 0, st))
    assert torch.equal(a, b)                                    # counter-based: reproducible
    _lib.check(lib.ap_randn(b.data_ptr(), n // 2, 1234, n // 8, st))
    assert torch.equal(b[: n // 2], a[n // 2:])                 # offset o == element 4*o of the same stream
    _lib.check(lib.ap_randn(b.data_ptr(), n, 1235, 0, st))
    assert not torch.equal(a, b)
    assert abs(a.mean().item()) < 3e-3 and abs(a.var().item() - 1) < 5e-3
    assert abs((a ** 4).mean().item() - 3) < 0.05 and a.abs().max().item() < 6.5
    assert abs(torch.corrcoef(torch.stack([a[:-1], a[1:]]))[0, 1].item()) < 3e-3


def test_vote_counts(ap):
    from audiopure_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(1001, 10, generator=g)
    logits[5] = 0.0                       # all-tie row -> class 0 like torch.max
    logits[7, 3] = logits[7, 8] = 9.0     # tie -> lowest index
    counts = torch.zeros(10, dtype=torch.int64, device="cuda")
    lc = logits.cuda()
    _lib.check(lib.ap_vote_counts(lc.data_ptr(), 1001, 10, counts.data_ptr(), 10, _lib.stream_ptr()))
    _lib.check(lib.ap_vote_counts(lc.data_ptr(), 1001, 10, counts.data_ptr(), 10, _lib.stream_ptr()))   # accumulates
    assert lib.ap_vote_counts(lc.data_ptr(), 1001, 10, counts.data_ptr(), 9, _lib.stream_ptr()) == -1   # K > len(counts)
    ref = torch.bincount(logits.max(1)[1], minlength=10)
    assert torch.equal(counts.cpu(), 2 * ref) and int(counts.sum()) == 2002


# ---------------------------------------------------------------------------------------------------- WaveNet
def _eps(ap, sd, cfg, x, t, mode):
    net = ap.WaveNet(sd, mode=mode, **cfg)
    B = x.shape[0]
    return net((cuda(x), t * torch.ones(B, 1))).cpu().numpy()


@pytest.mark.parametrize("key,L,B,t,seed", [("eps_full_L1024_t1", 1024, 2, 1.0, 1234), ("eps_full_L1024_t65", 1024, 2, 65.0, 1234),
                                            ("eps_full_L3001_t7", 3001, 1, 7.0, 77)])
@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16", "bf16x3"])
def test_wavenet_eps_vs_reference_golden(ap, golden, sd_full, key, L, B, t, seed, mode):
    x = synthetic.synthetic_waveforms(B, L, seed=seed)
    eps = _eps(ap, sd_full, synthetic.DEFAULT_WAVENET_CONFIG, x, t, mode)
    err = rel_l2(eps, golden[key])
    print(f"{key} {mode}: rel-L2 {err:.3e}")
    assert err < TOL_EPS[mode]


def test_wavenet_small_config_fp32(ap, golden):
    cfg = dict(synthetic.DEFAULT_WAVENET_CONFIG, res_channels=64, skip_channels=64, num_res_layers=5, dilation_cycle=3)
    sd = synthetic.wavenet_state_dict(seed=3, config=cfg)
    net = ap.WaveNet(sd, **cfg)
    assert net.mode == "fp32"                       # tensor-core kernels need 256 channels
    with pytest.raises(ap.AudioPureError):
        net.set_mode("bf16")
    x = synthetic.synthetic_waveforms(3, 500, seed=5)
    eps = net((cuda(x), 3.0 * torch.ones(3, 1))).cpu().numpy()
    assert rel_l2(eps, golden["eps_small_L500_t3"]) < 2e-5


def test_wavenet_layers_bf16_vs_fp32_and_oracle(ap, orc, sd_full):
    """Per-layer localisation: u_{n+1} and the gate output of selected layers, tensor-core path vs fp32 path vs oracle."""
    B, L, t = 2, 1024, 7.0
    x = synthetic.synthetic_waveforms(B, L, seed=11)
    _, internals = orc.wavenet_forward(sd_full, x, t * torch.ones(B, 1), return_internals=True)
    nets = {m: ap.WaveNet(sd_full, mode=m, **synthetic.DEFAULT_WAVENET_CONFIG) for m in ("fp32", "bf16")}
    xc = cuda(x)
    report = []
    for layer in (0, 1, 5, 11, 12, 23, 35):
        u32, g32 = nets["fp32"].debug_layer(xc, t, layer)
        u16, g16 = nets["bf16"].debug_layer(xc, t, layer)
        report.append((layer, rel_l2(g16, g32), rel_l2(u16, u32)))
        if f"o{layer}" in internals:     # oracle keeps layers 0, 1, 35: (B, C, L) -> (B, L, C)
            o = internals[f"o{layer}"].permute(0, 2, 1).numpy()
            assert rel_l2(g32, o) < 2e-5, f"fp32 gate of layer {layer}"
        if layer == 0:
            u1 = internals["u1"].permute(0, 2, 1).numpy()
            assert rel_l2(u32, u1) < 2e-5
    print("layer, gate rel-L2 (bf16 vs fp32), u_next rel-L2:", report)
    for layer, eg, eu in report:
        assert eg < 2e-2, report
        if layer != 35:                  # the last block's residual output is unused (WaveNet.py:131-135)
            assert eu < 2e-2, report


def test_wavenet_batch_chunking_and_mixed_steps(ap, sd_full):
    net = ap.WaveNet(sd_full, mode="bf16", **synthetic.DEFAULT_WAVENET_CONFIG)
    x = cuda(synthetic.synthetic_waveforms(5, 640, seed=3))
    net.reserve(2, 640)                              # 5 waveforms through a 2-waveform workspace: 3 chunks
    a = net.eps(x, 4.0)
    net.reserve(5, 640)
    b = net.eps(x, 4.0)
    assert torch.equal(a, b)
    steps = torch.tensor([[4.0], [9.0], [4.0], [9.0], [4.0]])
    c = net((x, steps))
    assert torch.equal(c[0::2], a[0::2]) and not torch.equal(c[1], a[1])


@pytest.mark.parametrize("B,L", [(1, 100), (3, 129), (2, 4097)])
def test_wavenet_edge_shapes_bf16_vs_fp32(ap, sd_full, B, L):
    """single partial tile, one position past a tile boundary, odd tile counts (ragged CTA pairs) -- tensor-core path vs the
    fp32 path (itself pinned to the reference at other shapes)"""
    nets = {m: ap.WaveNet(sd_full, mode=m, **synthetic.DEFAULT_WAVENET_CONFIG) for m in ("fp32", "bf16", "bf16x3")}
    x = cuda(synthetic.synthetic_waveforms(B, L, seed=L))
    e32, e16, ex3 = nets["fp32"].eps(x, 33.0), nets["bf16"].eps(x, 33.0), nets["bf16x3"].eps(x, 33.0)
    err, err3 = rel_l2(e16, e32), rel_l2(ex3, e32)
    print(f"eps bf16 vs fp32 at B={B} L={L}: rel-L2 {err:.3e}; bf16x3 vs fp32: {err3:.3e}")
    assert torch.isfinite(ex3).all() and err3 < 2e-4
    # clips much shorter than the receptive field have a small eps norm (most taps read zero padding), so the same
    # absolute bf16 noise is a larger relative error than at L >= 1024 (7e-3..9.5e-3): measured 1.5e-2..1.7e-2
    assert torch.isfinite(e16).all() and err < (2.5e-2 if L < 1024 else 1.2e-2)


def test_error_paths(ap, diffwave):
    from audiopure_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(2, 1, 256, device="cuda")
    assert lib.ap_diffwave_eps(diffwave.model._handle, x.data_ptr(), 1.0, x.data_ptr(), 0, 256, None) == -1      # B == 0
    assert b"positive" in lib.ap_last_error()
    assert lib.ap_diffwave_eps(None, x.data_ptr(), 1.0, x.data_ptr(), 2, 256, None) == -1                          # null handle
    assert lib.ap_vote_counts(x.data_ptr(), 4, 0, x.data_ptr(), 10, None) == -1                                       # K == 0
    with pytest.raises(_lib.AudioPureError):
        ap.MelSpectrogramDB(n_fft=2048, hop_length=512, n_mels=32, pad_mode="reflect")(torch.zeros(1, 1, 512, device="cuda"))
    with pytest.raises(AssertionError):
        ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))(torch.zeros(1, 1, 16, 16, device="cuda"))


def test_full_size_batch_properties(ap, sd_full):
    """BASELINE configs[1] size (512 x 1 s, t* = 2, bf16, in-kernel noise): batch independence (rows of the big batch equal
    the same rows purified alone with the same Philox offsets), determinism, and a bounded purification distance."""
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=2, state_dict=sd_full, noise="philox", seed=11)
    B, L = 512, 16000
    x = cuda(synthetic.synthetic_waveforms(B, L, seed=4321))
    y = dw.purify(x)
    dw._offset = 0
    y2 = dw.purify(x)
    assert torch.equal(y, y2)                                            # deterministic
    assert torch.isfinite(y).all()
    d = (y - x).flatten(1).norm(dim=1) / x.flatten(1).norm(dim=1)
    assert d.max().item() < 0.6 and d.min().item() > 1e-3                # every waveform was diffused and denoised
    # row r of a (B, L) Philox call uses counters offset + r*L/4 ..., so a single-row call at that offset reproduces it
    for r in (0, 255, 511):
        eps_full = dw.model.eps(x, 1.0)[r]
        eps_one = dw.model.eps(x[r:r + 1], 1.0)[0]
        assert torch.equal(eps_full, eps_one)                            # network output is batch independent, bit for bit


def test_inference_only_entry_points_and_cpu_inputs_raise(ap, diffwave):
    """gradients flow through WaveNet / DiffWave.forward (backward-pass tests below); the other entry points refuse an
    input that requires grad instead of silently dropping the gradient, and there is no CPU path"""
    x = torch.zeros(1, 1, 256, device="cuda", requires_grad=True)
    with pytest.raises(ap.AudioPureError):
        diffwave.one_shot_denoise(x)
    with pytest.raises(ap.AudioPureError):
        diffwave.model.eps(torch.zeros(1, 1, 256), 1.0)


# ---------------------------------------------------------------------------------------------------- backward pass
# Gradient tolerances.  The backward GEMMs use bf16 operands (their own rounding costs ~1 %).  The head's ReLU makes the
# gradient a discontinuous function of the forward values: a bf16 forward (skip sum off by ~0.8 %) flips ~1 % of the
# ReLU masks relative to the fp32 reference, which alone moves the gradient by 7-8 % (reproduced on the CPU by perturbing
# the oracle's skip sum by 0.8 %).  With the bf16x3 forward the masks agree and only the backward's rounding remains.
TOL_GRAD = {"bf16x3": 2e-2, "bf16": 1.5e-1}


@pytest.mark.parametrize("key,L,B,t,seed,gkey", [("vjp_gx_L1024_t7", 1024, 2, 7.0, 1234, "vjp_g_eps"),
                                                 ("vjp_gx_L1024_t65", 1024, 2, 65.0, 1234, "vjp_g_eps"),
                                                 ("vjp_gx_L3001_t7", 3001, 1, 7.0, 77, "vjp_g_eps_L3001")])
@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_wavenet_vjp_vs_reference_autograd(ap, golden_grad, sd_full, key, L, B, t, seed, gkey, mode):
    """g_x = (d eps / d x)^T g_eps from the CUDA backward kernels vs autograd through the unmodified reference WaveNet
    (tests/golden/make_golden_grad.py)."""
    net = ap.WaveNet(sd_full, mode=mode, **synthetic.DEFAULT_WAVENET_CONFIG)
    x = cuda(synthetic.synthetic_waveforms(B, L, seed=seed)).requires_grad_(True)
    g = cuda(golden_grad[gkey])
    eps = net((x, t * torch.ones(B, 1)))
    assert eps.requires_grad
    (gx,) = torch.autograd.grad(eps, x, g)
    err = rel_l2(gx, golden_grad[key])
    print(f"{key} {mode}: gradient rel-L2 {err:.3e}")
    assert err < TOL_GRAD[mode]
    # direct C-ABI call: linear in g_eps
    gx2 = net.eps_vjp(x.detach(), t, 2.0 * g)
    assert rel_l2(gx2, 2.0 * gx) < 1e-2


def test_vjp_modes_without_a_backward_raise(ap, sd_full):
    x = cuda(synthetic.synthetic_waveforms(1, 256, seed=1))
    for mode in ("fp32", "fp16"):
        with pytest.raises(ap.AudioPureError):
            ap.WaveNet(sd_full, mode=mode, **synthetic.DEFAULT_WAVENET_CONFIG).eps_vjp(x, 3.0, torch.ones_like(x))


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_ddpm_purifier_gradient_vs_reference_autograd(diffwave, golden_grad, mode):
    """d <w, DiffWave.forward(x)> / d x through two network evaluations, with the reference's injected noise."""
    diffwave.model.set_mode(mode)
    diffwave.reverse_timestep = 2
    x = cuda(synthetic.synthetic_waveforms(2, 1024, seed=1234)).requires_grad_(True)
    with TorchNormalInjector(2024) as inj:
        y = diffwave(x)
        assert inj.i == 2
    assert rel_l2(y.detach(), golden_grad["ddpm_t2_purified"]) < TOL[mode]
    (gx,) = torch.autograd.grad((y * cuda(golden_grad["ddpm_grad_w"])).sum(), x)
    err = rel_l2(gx, golden_grad["ddpm_t2_grad_L1024"])
    print(f"ddpm_t2 gradient {mode}: rel-L2 {err:.3e}")
    # the purifier's Jacobian is close to a scaled identity: d eps / d x enters with c_eps ~ 0.012 per step
    assert err < 0.1 * TOL_GRAD[mode]
    diffwave.model.set_mode("bf16")


def test_vjp_sub_batches_and_mode_switches(ap, sd_full):
    """48 x 1 s waveforms exceed the 24 GB bound on saved activations (0.62 GB each): the backward runs in sub-batches
    (38 + 10) and must equal the per-row results; switching bf16 <-> bf16x3 re-lays the workspace (one / two planes)."""
    net = ap.WaveNet(sd_full, mode="bf16", **synthetic.DEFAULT_WAVENET_CONFIG)
    x = cuda(synthetic.synthetic_waveforms(48, 16000, seed=5))
    g = torch.randn(x.shape, generator=torch.Generator().manual_seed(1)).cuda()
    e0 = net.eps(x[40:], 9.0)
    gx_all = net.eps_vjp(x, 9.0, g)
    gx_tail = net.eps_vjp(x[40:].contiguous(), 9.0, g[40:].contiguous())
    assert torch.isfinite(gx_all).all() and torch.equal(gx_all[40:], gx_tail)
    net.set_mode("bf16x3")
    e3 = net.eps(x[40:], 9.0)
    assert rel_l2(e0, e3) < 2.5e-2 and not torch.equal(e0, e3)      # bf16 vs fp32-class eps at full length: 1.6e-2
    gx3 = net.eps_vjp(x[40:].contiguous(), 9.0, g[40:].contiguous())
    assert rel_l2(gx_tail, gx3) < 0.2
    net.set_mode("bf16")
    assert torch.equal(net.eps(x[40:], 9.0), e0)


# ---------------------------------------------------------------------------------------------------- purifier
@pytest.mark.parametrize("mode", ["fp32", "bf16", "fp16", "bf16x3"])
def test_ddpm_purifier_vs_reference_golden(diffwave, golden, mode):
    diffwave.model.set_mode(mode)
    x = cuda(synthetic.synthetic_waveforms(2, 1024, seed=1234))
    for t_star, seed, key in [(2, 2024, "ddpm_t2_L1024"), (3, 2025, "ddpm_t3_L1024")]:
        diffwave.reverse_timestep = t_star
        with TorchNormalInjector(seed) as inj:
            y = diffwave(x)
            assert inj.i == t_star
        err = rel_l2(y, golden[key])
        print(f"{key} {mode}: rel-L2 {err:.3e}")
        assert err < TOL[mode]
    diffwave.reverse_timestep = 66
    assert rel_l2(diffwave.one_shot_denoise(x), golden["oneshot_t66_L1024"]) < TOL_X0[mode]
    assert rel_l2(diffwave.two_shot_denoise(x), golden["twoshot_t66_L1024"]) < TOL_X0[mode]
    diffwave.reverse_timestep = 9
    with TorchNormalInjector(2026):
        assert rel_l2(diffwave.fast_reverse(x), golden["fastrev_t9_L1024"]) < TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_reffwave_vs_reference_golden(ap, golden, diffwave, mode):
    diffwave.model.set_mode(mode)
    rw = ap.ReffWave(diffwave.model, diffwave.diffusion_hyperparams, reverse_timestep=5, num_re=2, noise="torch")
    x = cuda(synthetic.synthetic_waveforms(2, 1024, seed=1234))
    with TorchNormalInjector(2029) as inj:
        y = rw(x)
        assert inj.i == 2
    assert rel_l2(y, golden["reffwave_t5_re2_L1024"]) < TOL[mode]


def test_purify_entry_point_philox(ap, sd_full):
    """ap_diffwave_purify_ddpm (whole purifier in one call) == the step-by-step API on the same Philox stream."""
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=3, state_dict=sd_full, noise="philox", seed=7)
    x = cuda(synthetic.synthetic_waveforms(2, 768, seed=5))
    a = dw.purify(x)
    dw._offset = 0
    b = dw(x)
    assert torch.equal(a, b)
    assert torch.isfinite(a).all() and rel_l2(a, x) < 0.5


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("t_star", [1, 3])
def test_sde_purifier_vs_oracle(ap, orc, sd_full, mode, t_star):
    import argparse
    args = argparse.Namespace(ddpm_path=None, ddpm_config=CONFIG_JSON, t=t_star, score_type="guided_diffusion", rand_t=False,
                              t_delta=0, use_bm=False, sample_step=1)
    rdw = ap.RevDiffWave(args, state_dict=sd_full, noise="torch", mode=mode)
    x = synthetic.synthetic_waveforms(2, 1024, seed=21)
    n_steps = len(orc.sde_euler_schedule(t_star))
    zs = [synthetic.host_noise(x.shape, 3000 + t_star, i) for i in range(1 + n_steps)]
    want = orc.sde_purify(sd_full, x, t_star, orc.NoiseSource(zs)).numpy()
    it = iter(zs)
    orig_like, orig_randn = torch.randn_like, torch.randn
    torch.randn_like = lambda t, **kw: cuda(next(it))
    torch.randn = lambda *a, **kw: cuda(next(it))
    try:
        got = rdw(cuda(x))
    finally:
        torch.randn_like, torch.randn = orig_like, orig_randn
    err = rel_l2(got, want)
    print(f"sde t*={t_star} {mode}: rel-L2 {err:.3e}")
    assert err < (2e-5 if mode == "fp32" else 1e-2)


def test_sde_purifier_gradient_vs_oracle_autograd(ap, orc, sd_full):
    """The OPT-IN exact gradient of the Euler-Maruyama chain (``grad_through_eps=True``: back-propagation through the network,
    which the reference does not do -- its compute_eps_t is no_grad; the default, reference-faithful gradient is pinned to the
    reference in test_gpu_golden_v2.py): gradient of <w, purified> (t* = 2) vs autograd over the oracle, same injected noise."""
    import argparse
    t_star = 2
    args = argparse.Namespace(ddpm_path=None, ddpm_config=CONFIG_JSON, t=t_star, score_type="guided_diffusion", rand_t=False,
                              t_delta=0, use_bm=False, sample_step=1)
    rdw = ap.RevDiffWave(args, state_dict=sd_full, noise="torch", mode="bf16x3", grad_through_eps=True)
    x = synthetic.synthetic_waveforms(2, 1024, seed=21)
    w = synthetic.host_noise(x.shape, 77, 0)
    n_steps = len(orc.sde_euler_schedule(t_star))
    zs = [synthetic.host_noise(x.shape, 3000 + t_star, i) for i in range(1 + n_steps)]
    xo = torch.from_numpy(x).requires_grad_(True)
    yo = orc.sde_purify(sd_full, xo, t_star, orc.NoiseSource(zs), grad_through_eps=True)
    (g_want,) = torch.autograd.grad((yo * torch.from_numpy(w)).sum(), xo)
    it = iter(zs)
    orig_like, orig_randn = torch.randn_like, torch.randn
    torch.randn_like = lambda t, **kw: cuda(next(it))
    torch.randn = lambda *a, **kw: cuda(next(it))
    try:
        xg = cuda(x).requires_grad_(True)
        got = rdw(xg)
    finally:
        torch.randn_like, torch.randn = orig_like, orig_randn
    assert rel_l2(got.detach(), yo.detach().numpy()) < 1e-5
    (g_got,) = torch.autograd.grad((got * cuda(w)).sum(), xg)
    err = rel_l2(g_got, g_want.numpy())
    print(f"sde t*={t_star} gradient (bf16x3 forward): rel-L2 {err:.3e}")
    assert err < 2e-3


# ---------------------------------------------------------------------------------------------------- mel + classifiers
def test_mel_vs_torchaudio_golden(ap, golden):
    xm = cuda(synthetic.synthetic_waveforms(2, 16000, seed=99))
    sc = ap.sc09_transform()(xm).cpu().numpy()
    assert sc.shape == golden["mel_sc09"].shape == (2, 1, 32, 32)
    assert np.abs(sc - golden["mel_sc09"]).max() < 2e-3
    kw = ap.kws_transform()(xm).cpu().numpy()
    assert kw.shape == golden["mel_kws"].shape == (2, 1, 32, 81)
    assert np.abs(kw - golden["mel_kws"]).max() < 2e-3


def test_classifiers_vs_reference_golden(ap, golden):
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    assert rx.mode == "tf32"                                    # tensor-core convolutions by default
    spec = cuda(golden["mel_sc09"])
    want = golden["resnext_logits"]
    l_tf = rx(spec).cpu().numpy()
    rx.set_mode("fp32")
    l_32 = rx(spec).cpu().numpy()
    assert np.abs(l_32 - want).max() < 2e-3 * np.abs(want).max()
    assert (l_32.argmax(1) == want.argmax(1)).all()             # fp32 mode: top-1 exact
    err = np.abs(l_tf - want).max()
    print(f"resnext tf32 max|dlogit| {err:.3e} (spread {want.max() - want.min():.2f})")
    assert err < 2e-2                                           # tf32 operands (the reference's own cuDNN precision class)
    srt = np.sort(want, 1)
    clear = (srt[:, -1] - srt[:, -2]) > 4 * err
    assert (l_tf.argmax(1)[clear] == want.argmax(1)[clear]).all()
    xm = cuda(synthetic.synthetic_waveforms(2, 16000, seed=99))
    m5 = ap.M5Classifier(synthetic.m5_state_dict(seed=0))
    np.testing.assert_allclose(m5(xm).cpu().numpy(), golden["m5_logprobs"], atol=2e-4, rtol=0)
    kws = ap.KWSClassifier(synthetic.kws_state_dict(seed=0))
    np.testing.assert_allclose(kws(cuda(golden["mel_kws"])).cpu().numpy(), golden["kws_logprobs"], atol=2e-4, rtol=0)


@pytest.mark.parametrize("depth", [34, 50])
def test_resnet_family_vs_reference_golden(ap, golden, depth):
    rn = ap.ResNetClassifier(synthetic.resnet_state_dict(depth=depth, seed=0), depth=depth)
    assert rn.mode == "tf32"                                  # default, like cuDNN for the reference on this GPU
    want = golden[f"resnet{depth}_logits"]
    lt = rn(cuda(golden["mel_sc09"])).cpu().numpy()
    logits = rn.set_mode("fp32")(cuda(golden["mel_sc09"])).cpu().numpy()
    print(f"ResNet-{depth} logits max err: fp32 {np.abs(logits - want).max():.2e}, tf32 {np.abs(lt - want).max():.2e}")
    assert np.abs(logits - want).max() < 1e-3 * max(1.0, np.abs(want).max())
    assert (logits.argmax(1) == want.argmax(1)).all()
    assert np.abs(lt - want).max() < 2e-2 * max(1.0, np.abs(want).max()) and (lt.argmax(1) == want.argmax(1)).all()
    x70 = cuda(golden["mel_sc09"]).repeat(35, 1, 1, 1)        # many images per tensor-core tile at the 2x2 / 1x1 stages, ragged tail
    assert torch.allclose(rn.set_mode("tf32")(x70)[-2:], torch.from_numpy(lt).cuda(), atol=1e-5)


@pytest.mark.parametrize("mask", [1, 2, 4, 7])
def test_resnext_tensor_core_convs_by_kind(ap, mask, monkeypatch):
    """tf32 tensor-core convolutions enabled per kind (1: 1x1, 2: 3x3 stride 1, 4: stride 2) against the fp32 FFMA path,
    on a batch that spans two workspace chunks (64 + 6) including an odd image count for the 2-images-per-tile stage."""
    monkeypatch.setenv("AP_CLS_TC_MASK", str(mask))
    sd = synthetic.resnext_state_dict(seed=0)
    rx = ap.ResNeXtClassifier(sd)
    g = torch.Generator().manual_seed(mask)
    spec = (torch.randn(70, 1, 32, 32, generator=g) * 20 - 30).cuda()
    l_tf = rx(spec)
    rx.set_mode("fp32")
    l_32 = rx(spec)
    err = (l_tf - l_32).abs().max().item()
    print(f"mask {mask}: max|dlogit| {err:.3e}, spread {(l_32.max() - l_32.min()).item():.2f}")
    assert err < 3e-2
    sub = rx(spec[:7])                                            # per-sample results do not depend on the batch
    assert torch.equal(sub, l_32[:7])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_acoustic_system_vs_reference_golden(ap, golden, diffwave, mode):
    diffwave.model.set_mode(mode)
    diffwave.reverse_timestep = 2
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    rx.set_mode("fp32" if mode == "fp32" else "tf32")
    system = ap.AcousticSystem(classifier=rx, transform=ap.sc09_transform(), defender=diffwave, defense_type="wave")
    x1 = cuda(synthetic.synthetic_waveforms(1, 16000, seed=1234))
    with TorchNormalInjector(2027):
        purified = diffwave(x1)
    assert rel_l2(purified, golden["system_purified"]) < TOL[mode]
    with TorchNormalInjector(2027):
        logits = system(x1).cpu().numpy()
    srt = np.sort(golden["system_logits"], 1)
    if mode == "fp32" or srt[0, -1] - srt[0, -2] > 0.05:
        assert logits.argmax(1)[0] == golden["system_logits"].argmax(1)[0]                   # top-1 exact
    assert np.abs(logits - golden["system_logits"]).max() < (5e-3 if mode == "fp32" else 5e-2) * np.abs(golden["system_logits"]).max()
    raw = system(x1 * 2 ** 15, defend=False).cpu().numpy()                                   # int16-range branch
    assert np.abs(raw - golden["system_logits_int16_nodefense"]).max() < 5e-3 * np.abs(raw).max()


# ---------------------------------------------------------------------------------------------------- certification
def test_smooth_predict_counts_vs_reference_golden(ap, golden, diffwave):
    diffwave.model.set_mode("fp32")
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0)).set_mode("fp32")
    rc = ap.RobustCertificate(classifier=rx, transform=ap.sc09_transform(), denoiser=diffwave, num_classes=10, noise="torch")
    x1 = cuda(synthetic.synthetic_waveforms(1, 16000, seed=1234))
    with TorchNormalInjector(2028) as inj:
        counts = rc.smooth_predict(x1, num_sampling=12, sigma=0.5, batch_size=5)     # ragged last batch
        assert inj.i == 3
    assert counts.dtype == torch.int64 and int(counts.sum()) == 12
    assert np.array_equal(counts.numpy(), golden["smooth_counts_n12"])
    assert [rc.compute_t_star(1 / (1 + s ** 2)) for s in (0.25, 0.5, 1.0)] == list(golden["smooth_t_star"])


def test_sharded_draws_equal_single_rank(ap, diffwave):
    """Rank r of W draws slice r of the SAME Philox stream: the per-rank vote vectors of an emulated 3-rank run (ranks
    executed one after the other on this GPU, no process group) add up to the single-rank counts exactly."""
    diffwave.model.set_mode("bf16")
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    tr = ap.sc09_transform()
    x1 = cuda(synthetic.synthetic_waveforms(1, 16000, seed=31))

    class Emulated(ap.RobustCertificate):
        world, rank = 1, 0

        def _world(self):
            return self.world, self.rank

        def forward(self, x_in):                                   # checksum of every noisy copy, in draw order
            self.seen.append(x_in.double().sum(dim=(1, 2)).cpu())
            return super().forward(x_in)

    def run(world, rank, batch):
        rc = Emulated(classifier=rx, transform=tr, denoiser=diffwave, num_classes=10, seed=17, distributed=False)
        rc.world, rc.rank, rc.seen = world, rank, []
        c = rc.smooth_predict(x1, num_sampling=50, sigma=1.0, batch_size=batch)
        return c, torch.cat(rc.seen)

    whole, sums = run(1, 0, 32)
    parts = [run(3, r, 7) for r in range(3)]                       # 17 + 17 + 16 draws, ragged batches of 7
    assert int(whole.sum()) == 50 and torch.equal(whole, sum(c for c, _ in parts))
    assert torch.equal(sums, torch.cat([s_ for _, s_ in parts]))   # draw i sees the same noise whichever rank takes it
    assert sums.unique().numel() == 50                             # and every draw is different


def test_philox_counts_agree_with_reference_noise_within_binomial_tolerance(ap):
    """Vote counts drawn with the in-kernel Philox generator vs the reference's torch.normal CPU noise (pure randomized
    smoothing, denoiser=None -- 'randsmooth' in certified_robustness_eval.py:96-97) on an input / classifier pair whose votes
    split between two classes: the two count vectors are independent samples of the same multinomial."""
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=1))
    tr = ap.sc09_transform()
    x0 = cuda(synthetic.synthetic_waveforms(1, 16000, seed=31))
    # find an (input scale, sigma) on this classifier's decision boundary: random-init networks vote unanimously almost everywhere
    torch.manual_seed(1)
    pick = None
    for scale, sigma in [(0.8944, 0.4472), (0.8944, 0.40), (0.8944, 0.50), (1.0, 0.40), (1.0, 0.45), (0.8, 0.45), (0.95, 0.45)]:
        probe = ap.RobustCertificate(rx, tr, denoiser=None, noise="torch").smooth_predict(x0 * scale, 100, sigma=sigma, batch_size=100)
        if int(probe.max()) <= 85:
            pick = (scale, sigma)
            break
    if pick is None:
        pytest.skip("no split vote found for the synthetic classifier on this build")
    x1, sigma = x0 * pick[0], pick[1]
    n = 600
    torch.manual_seed(0)
    c_ref = ap.RobustCertificate(rx, tr, denoiser=None, noise="torch").smooth_predict(x1, n, sigma=sigma, batch_size=200)
    c_phx = ap.RobustCertificate(rx, tr, denoiser=None, noise="philox", seed=99).smooth_predict(x1, n, sigma=sigma, batch_size=150)
    print(f"scale {pick[0]} sigma {sigma}: votes torch.normal:", c_ref.tolist(), " votes philox:", c_phx.tolist())
    assert int(c_ref.sum()) == n and int(c_phx.sum()) == n
    assert int(c_ref.max()) < 0.95 * n                                   # a real split, not a degenerate vote
    p = c_ref.double() / n
    tol = 4.0 * torch.sqrt(2 * n * p * (1 - p)) + 3                      # 4 sigma of the difference of two binomials
    assert bool(((c_ref - c_phx).abs().double() <= tol).all())


def test_certify_dataset_records(ap, diffwave, tmp_path):
    import json
    diffwave.model.set_mode("bf16")
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    rc = ap.RobustCertificate(classifier=rx, transform=ap.sc09_transform(), denoiser=diffwave, num_classes=10, seed=2)
    xs = torch.from_numpy(synthetic.synthetic_waveforms(3, 16000, seed=5))
    batches = [{"samples": xs[:2, 0], "target": torch.tensor([6, 1])}, (xs[2:], torch.tensor([6]))]
    recs = ap.certify_dataset(rc, batches, sigma=0.5, num_sampling=64, n_0=16, batch_size=32, save_path=str(tmp_path))
    assert [r["id"] for r in recs] == [0, 1, 2] and set(recs[0]) == {"id", "y_true", "y_pred", "certified_radius"}
    on_disk = json.load(open(tmp_path / "sigma=0.5" / "sigma=0.5_N=64.json"))
    assert on_disk == recs
    for r in recs:
        assert (r["y_pred"] == -1 and r["certified_radius"] == 0) or (0 <= r["y_pred"] < 10 and r["certified_radius"] > 0)


def test_certify_philox_end_to_end(ap, diffwave):
    diffwave.model.set_mode("bf16")
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    rc = ap.RobustCertificate(classifier=rx, transform=ap.sc09_transform(), denoiser=diffwave, num_classes=10, seed=5)
    x = cuda(synthetic.synthetic_waveforms(2, 16000, seed=8))
    y = torch.tensor([1, 2])
    y_pred, radius = rc.certify(x, y, sigma=0.5, n_0=16, n=96, batch_size=32)
    assert y_pred.shape == (2,) and radius.shape == (2,) and radius.dtype == torch.float32
    rc2 = ap.RobustCertificate(classifier=rx, transform=ap.sc09_transform(), denoiser=diffwave, num_classes=10, seed=5)
    y_pred2, radius2 = rc2.certify(x, y, sigma=0.5, n_0=16, n=96, batch_size=48)     # other batching, same Philox stream
    assert torch.equal(y_pred, y_pred2) and torch.equal(radius, radius2)
    for i in range(2):
        assert (y_pred[i] == -1 and radius[i] == 0) or (0 <= y_pred[i] < 10 and radius[i] > 0)
    # Clopper-Pearson KATs
    assert abs(rc.lower_conf_bound(99000, 100000) - 0.988989) < 1e-5
    assert abs(rc.lower_conf_bound(100000, 100000) - 0.999931) < 1e-5
    assert rc.lower_conf_bound(50200, 100000) < 0.5 + 1e-3


@pytest.mark.parametrize("B", [1, 8])
def test_cuda_graph_capture_and_replay(ap, sd_full, B):
    """The whole purify -> mel -> classify pipeline is free of host synchronisation and hidden allocation after warm-up, so
    it can be captured in a CUDA graph and replayed (small-batch query serving)."""
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=2, state_dict=sd_full, noise="philox", seed=3)
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    system = ap.AcousticSystem(classifier=rx, transform=ap.sc09_transform(), defender=dw.purify, defense_type="wave",
                               check_int16_range=False)
    x = cuda(synthetic.synthetic_waveforms(B, 16000, seed=77))
    dw._offset = 0
    want = system(x)                       # warm-up: workspaces, tensor maps, kernel attributes
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    dw._offset = 0                         # the Philox offsets are baked into the captured launches
    with torch.cuda.graph(g):
        got = system(x)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    x.copy_(cuda(synthetic.synthetic_waveforms(B, 16000, seed=78)))     # new input, same graph
    g.replay()
    torch.cuda.synchronize()
    dw._offset = 0
    assert torch.equal(got, system(x))
    def timeit(fn, n=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    def eager():
        dw._offset = 0
        system(x)
    print(f"B={B}: eager {timeit(eager):.2f} ms, graph replay {timeit(g.replay):.2f} ms per query batch")


# ---------------------------------------------------------------------------------------------------- end-to-end gradients
@pytest.mark.parametrize("name", ["sc09", "kws"])
def test_mel_vjp_vs_torchaudio_autograd(ap, golden_grad, name):
    tr = ap.sc09_transform() if name == "sc09" else ap.kws_transform()
    x = cuda(synthetic.synthetic_waveforms(2, 16000, seed=99)).requires_grad_(True)
    spec = tr(x)
    assert spec.requires_grad
    (gx,) = torch.autograd.grad(spec, x, cuda(golden_grad[f"mel_{name}_g_spec"]))
    err = rel_l2(gx, golden_grad[f"mel_{name}_grad"])
    print(f"mel {name} gradient: rel-L2 {err:.3e}")
    assert err < 1e-4


@pytest.mark.parametrize("mode", ["fp32", "tf32-bwd", "tf32"])
def test_resnext_vjp_vs_reference_autograd(ap, golden_grad, mode, monkeypatch):
    """fp32: FFMA forward recompute and data-gradient convolutions.  tf32 (the default mode): both on the tcgen05 kernel --
    the gradient of the tf32 forward, whose ReLU masks differ from the fp32 forward's in ~0.1 % of the units per layer: a ReLU
    network's gradient is discontinuous in the forward values, and that alone moves it by ~6 %.  tf32-bwd keeps the recomputed
    forward in fp32 (AP_CLS_VJP_FWD_FP32=1) and so checks the tensor-core data-gradient kernels alone."""
    if mode == "tf32-bwd":
        monkeypatch.setenv("AP_CLS_VJP_FWD_FP32", "1")
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0)).set_mode("fp32" if mode == "fp32" else "tf32")
    spec = cuda(golden_grad["resnext_in_spec"]).requires_grad_(True)
    logits = rx(spec)
    (gs,) = torch.autograd.grad(logits, spec, cuda(golden_grad["resnext_g_logits"]))
    err = rel_l2(gs, golden_grad["resnext_grad"])
    print(f"ResNeXt gradient ({mode}): rel-L2 {err:.3e}")
    assert err < {"fp32": 1e-5, "tf32-bwd": 5e-3, "tf32": 1.5e-1}[mode]


@pytest.mark.parametrize("depth", [34, 50])
def test_resnet_vjp_vs_reference_autograd(ap, golden_grad, depth, monkeypatch):
    rn = ap.ResNetClassifier(synthetic.resnet_state_dict(depth=depth, seed=0), depth=depth).set_mode("fp32")
    spec = cuda(golden_grad["resnext_in_spec"]).requires_grad_(True)
    g_logits = cuda(golden_grad["resnext_g_logits"])
    (gs,) = torch.autograd.grad(rn(spec), spec, g_logits)
    err = rel_l2(gs, golden_grad[f"resnet{depth}_grad"])
    rn.set_mode("tf32")                                       # tensor-core forward tape and data-gradient convolutions
    (gt,) = torch.autograd.grad(rn(spec), spec, g_logits)
    monkeypatch.setenv("AP_CLS_VJP_FWD_FP32", "1")            # fp32 forward tape, tensor-core data gradients only
    (gtb,) = torch.autograd.grad(rn(spec), spec, g_logits)
    err_t, err_tb = rel_l2(gt, golden_grad[f"resnet{depth}_grad"]), rel_l2(gtb, golden_grad[f"resnet{depth}_grad"])
    print(f"ResNet-{depth} gradient rel-L2: fp32 {err:.3e}, tf32 {err_t:.3e}, tf32 backward on an fp32 forward {err_tb:.3e}")
    assert err < 1e-4 and err_t < 0.3 and err_tb < 2e-2


def test_kws_vjp_vs_reference_autograd(ap, golden_grad):
    """separable conv -> 2-layer bidirectional GRU -> additive attention -> log_softmax: back-propagation through time in one
    CTA per sample vs autograd through the reference's KWSModel"""
    kws = ap.KWSClassifier(synthetic.kws_state_dict(seed=0))
    spec = cuda(golden_grad["kws_in_spec"]).requires_grad_(True)
    (gs,) = torch.autograd.grad(kws(spec), spec, cuda(golden_grad["kws_g_logp"]))
    err = rel_l2(gs, golden_grad["kws_grad"])
    print(f"RCNN_KWS gradient: rel-L2 {err:.3e}")
    assert err < 1e-4


def test_m5_vjp_vs_reference_autograd(ap, golden_grad):
    m5 = ap.M5Classifier(synthetic.m5_state_dict(seed=0))
    x = cuda(synthetic.synthetic_waveforms(2, 16000, seed=99)).requires_grad_(True)
    (gx,) = torch.autograd.grad(m5(x), x, cuda(golden_grad["m5_g_logp"]))
    err = rel_l2(gx, golden_grad["m5_grad"])
    print(f"M5 gradient: rel-L2 {err:.3e}")
    assert err < 1e-4


def test_acoustic_system_loss_gradient_vs_reference_autograd(ap, golden_grad, sd_full):
    """d CrossEntropy(AcousticSystem(x), y) / d x through DDPM t* = 2 -> log-mel -> ResNeXt: the white-box attack gradient
    (robustness_eval/white_box_attack.py:430-438), every stage on the CUDA backward kernels."""
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=2, state_dict=sd_full, noise="torch", mode="bf16x3")
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0)).set_mode("fp32")   # tf32 forward: ReLU-mask flips (see above)
    system = ap.AcousticSystem(classifier=rx, transform=ap.sc09_transform(), defender=dw, defense_type="wave")
    x = cuda(synthetic.synthetic_waveforms(1, 16000, seed=1234)).requires_grad_(True)
    with TorchNormalInjector(2027) as inj:
        logits = system(x)
        assert inj.i == 2
    assert np.abs(logits.detach().cpu().numpy() - golden_grad["system_logits"]).max() < 5e-3
    loss = torch.nn.functional.cross_entropy(logits, torch.tensor([3], device="cuda"))
    (gx,) = torch.autograd.grad(loss, x)
    err = rel_l2(gx, golden_grad["system_loss_grad"])
    print(f"AcousticSystem loss gradient (bf16x3 purifier, fp32 classifier): rel-L2 {err:.3e}")
    assert err < 2e-2


def test_pgd_steps_through_the_defended_system_raise_the_loss(ap, sd_full):
    """The reference's white-box attack loop (robustness_eval/white_box_attack.py:430-447: loss.backward(), delta += lr * sign(grad),
    clamp to eps) run against the defended system on the CUDA path: fixed Philox noise (seed reset every step, as an attacker with
    a fixed defender draw), L_inf eps = 0.002, 4 steps -- the loss on the true label must go up."""
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=1, state_dict=sd_full, noise="philox", seed=11, mode="bf16")
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    system = ap.AcousticSystem(classifier=rx, transform=ap.sc09_transform(), defender=dw, defense_type="wave")
    x = cuda(synthetic.synthetic_waveforms(2, 16000, seed=321))
    with torch.no_grad():
        dw._offset = 0
        y = system(x).argmax(1)
    delta = torch.zeros_like(x, requires_grad=True)
    eps, lr, losses = 2e-3, 5e-4, []
    for _ in range(5):
        dw._offset = 0
        loss = torch.nn.functional.cross_entropy(system(x + delta), y)
        losses.append(float(loss.detach()))
        (grad,) = torch.autograd.grad(loss, delta)
        assert torch.isfinite(grad).all() and float(grad.abs().max()) > 0
        delta.data = (delta.data + lr * grad.sign()).clamp_(-eps, eps)
    print("PGD losses:", [f"{v:.4f}" for v in losses])
    assert losses[-1] > losses[0]


def test_backward_forms_give_the_same_gradient(tmp_path):
    """The three forms of the DiffWave backward pass -- one fused launch per layer boundary (default), two launches per layer
    (AP_BWD_UNFUSED) and the fused launch with TMA-staged epilogue streams (AP_BWD_STAGED) -- are selected per process; run each in
    its own interpreter on the same input and compare g_x bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "diffusion-model-for-audio-defense_b200", "devtools", "bwd_profile.py")
    outs = {}
    for name, extra in (("fused", {}), ("unfused", {"AP_BWD_UNFUSED": "1"}), ("staged", {"AP_BWD_STAGED": "1"})):
        env = {k: v for k, v in os.environ.items() if k not in ("AP_BWD_UNFUSED", "AP_BWD_STAGED")}
        env.update(extra)
        out = str(tmp_path / f"gx_{name}.npy")
        r = subprocess.run([sys.executable, script, out, "3"], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = np.load(out)
    assert np.isfinite(outs["fused"]).all() and np.abs(outs["fused"]).max() > 0
    assert np.array_equal(outs["fused"], outs["unfused"]) and np.array_equal(outs["fused"], outs["staged"])


def test_pgd_through_the_spectrogram_purifier_raises_the_loss(ap):
    """The same attack loop against the 'spec' defense (adaptive_attack_eval.py:134-137): waveform -> log-mel -> Diffusion-Spec
    reverse-SDE t* = 1 -> ResNeXt; the gradient runs through the mel, UNet and classifier backward kernels."""
    import argparse
    args = argparse.Namespace(ddpm_path=None, t=1, score_type="guided_diffusion", rand_t=False, t_delta=15, use_bm=False, sample_step=1)
    rid = ap.RevImprovedDiffusion(args, state_dict=synthetic.unet_state_dict(seed=0), noise="philox", seed=11)
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    system = ap.AcousticSystem(classifier=rx, transform=ap.sc09_transform(), defender=rid, defense_type="spec")
    x = cuda(synthetic.synthetic_waveforms(2, 16000, seed=321))
    with torch.no_grad():
        rid._offset = 0
        logits0 = system(x)
        y = logits0.argmax(1)
    delta = torch.zeros_like(x, requires_grad=True)
    eps, lr, losses = 2e-3, 5e-4, []
    for i in range(5):
        rid._offset = 0
        logits = system(x + delta)
        if i == 0:     # the autograd route draws the same noise and computes the same chain as the inference route
            assert float((logits.detach() - logits0).abs().max()) < 1e-2
        loss = torch.nn.functional.cross_entropy(logits, y)
        losses.append(float(loss.detach()))
        (grad,) = torch.autograd.grad(loss, delta)
        assert torch.isfinite(grad).all() and float(grad.abs().max()) > 0
        delta.data = (delta.data + lr * grad.sign()).clamp_(-eps, eps)
    print("PGD through Diffusion-Spec, losses:", [f"{v:.4f}" for v in losses])
    assert losses[-1] > losses[0]


# ------------------------------------------------------------------------------------ black-box query serving (section 8f-3)
def test_query_loss_kernels_vs_reference_golden(ap, golden_blackbox):
    """ap_query_loss / ap_query_loss_vjp vs nn.CrossEntropyLoss(reduction='none') and SEC4SR_MarginLoss of the reference
    (robustness_eval/_utils.py:30-125), values and gradients (EOT backpropagates ones, _EOT.py:43-44)."""
    from audiopure_b200.blackbox import QueryLoss, resolve_loss
    g = golden_blackbox
    y = cuda(g["loss_labels"])
    loss_fn, sign = resolve_loss("Margin", False, 0.5, "SCR", None, False)
    assert sign == 1 and resolve_loss("Entropy", True)[1] == -1
    s = cuda(g["loss_scores"]).requires_grad_(True)
    l = loss_fn(s, y)
    l.backward(torch.ones_like(l))
    np.testing.assert_allclose(l.detach().cpu().numpy(), g["loss_entropy"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(s.grad.cpu().numpy(), g["loss_entropy_grad"], rtol=0, atol=1e-6)
    _, dec = loss_fn.loss_and_decision(s.detach(), y)
    assert np.array_equal(dec.cpu().numpy(), g["loss_decision"])          # includes a tied row: first index wins
    for targeted in (0, 1):
        for clip in (0, 1):
            m = QueryLoss("Margin", bool(targeted), 0.5, bool(clip))
            s = cuda(g["loss_scores"]).requires_grad_(True)
            l = m(s, y)
            l.backward(torch.ones_like(l))
            assert np.array_equal(l.detach().cpu().numpy(), g[f"loss_margin_t{targeted}_c{clip}"])
            assert np.array_equal(s.grad.cpu().numpy(), g[f"loss_margin_t{targeted}_c{clip}_grad"])
    bad = loss_fn(cuda(g["loss_scores"][:2]), torch.tensor([-1, 10], device="cuda"))
    assert torch.isnan(bad).all()                                         # out-of-range label: NaN, never an OOB read


class _RandnInjector:
    """torch.randn(size, device=...) -> host noise popped in call order (what make_golden_blackbox.py fed the reference)."""

    def __init__(self, seed):
        self.seed, self.i, self._orig = seed, 0, torch.randn

    def __enter__(self):
        def fake(size, device=None, **kw):
            z = synthetic.host_noise(tuple(size), self.seed, self.i)
            self.i += 1
            return torch.from_numpy(z).to(device)
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self._orig


@pytest.mark.parametrize("name", ["a", "b"])
def test_nes_vs_reference_golden(ap, golden_blackbox, name):
    """NES + EOT on the CUDA kernels vs robustness_eval._NES.NES / _EOT.EOT of the reference around the same stand-in
    model and the same injected noise: the query batches bit for bit, scores / losses to fp32 rounding, the estimate to
    the conditioning of the estimator (loss differences of ~sigma divided by sigma).  'b' has a ragged length (1001:
    scalar path), EOT 4 x 2 and a single draw."""
    from audiopure_b200.blackbox import EOT, NES, QueryLoss
    g = golden_blackbox
    A, L, spd, bs, es, eb = (int(v) for v in g[f"nes_{name}_cfg"])
    sigma = float(g[f"nes_{name}_sigma"])
    toy, seen = ToyDefendedModel(L), []

    def model(x):
        seen.append(x.detach().clone())
        return toy(x)

    nes = NES(spd, bs, sigma, EOT(model, QueryLoss("Entropy"), es, eb, False), noise="torch")
    with _RandnInjector(4000) as inj:
        mean_loss, grad, adver_loss, adver_score, predict = nes(cuda(g[f"nes_{name}_x"]), cuda(g[f"nes_{name}_y"]))
        assert inj.i == spd // bs
    per_draw = es // eb
    for i in range(spd // bs):
        q = seen[i * per_draw]
        assert np.array_equal(q[: q.shape[0] // eb].cpu().numpy(), g[f"nes_{name}_queries{i}"])
    np.testing.assert_allclose(adver_score.cpu().numpy(), g[f"nes_{name}_adver_score"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(adver_loss.cpu().numpy(), g[f"nes_{name}_adver_loss"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(mean_loss.cpu().numpy(), g[f"nes_{name}_mean_loss"], rtol=0, atol=5e-6)
    assert np.array_equal(np.asarray(predict), g[f"nes_{name}_predict"])
    err = rel_l2(grad, g[f"nes_{name}_grad"])
    print(f"NES[{name}] gradient estimate vs reference: rel-L2 {err:.2e}")
    assert grad.shape == (A, 1, L) and err < 2e-4       # measured 2e-6


@pytest.mark.parametrize("A,S,L,first", [(2, 8, 1024, 1), (3, 6, 1001, 0), (1, 200, 16000, 1)])
def test_nes_kernels_exact_arithmetic_and_philox_regeneration(ap, A, S, L, first):
    """ap_nes_gradient against float64 numpy on given losses and noise; and the Philox path: the noise the perturbation
    kernel adds is the stream ap_randn produces at the same (seed, offset), and the gradient kernel regenerates exactly it."""
    from audiopure_b200 import _lib
    lib = _lib.load()
    H, R = S // 2, S + first
    x = cuda(synthetic.synthetic_waveforms(A, L, seed=9)[:, 0])
    z = cuda(synthetic.host_noise((A, H, L), 123, 0))
    loss = cuda(synthetic.host_noise((A, R), 124, 0))
    st = _lib.stream_ptr()
    ref = (1.0 / S) * np.einsum("aj,ajn->an", (loss[:, first:first + H] - loss[:, first + H:]).cpu().numpy().astype(np.float64),
                                z.cpu().numpy().astype(np.float64))
    grad = torch.full((A, L), 7.0, device="cuda")
    _lib.check(lib.ap_nes_gradient(loss.data_ptr(), z.data_ptr(), 0, 0, first, 1.0 / S, 0, grad.data_ptr(), A, S, L, st))
    assert rel_l2(grad, ref) < 1e-6
    _lib.check(lib.ap_nes_gradient(loss.data_ptr(), z.data_ptr(), 0, 0, first, 1.0 / S, 1, grad.data_ptr(), A, S, L, st))
    assert rel_l2(grad, 2 * ref) < 1e-6                                    # accumulate = 1 adds to the previous draw
    # in-kernel noise
    seed, off, sigma = 77, 12345, 0.05
    out = torch.empty(A, R, L, device="cuda")
    _lib.check(lib.ap_nes_perturb(x.data_ptr(), sigma, None, seed, off, first, out.data_ptr(), A, S, L, st))
    zr = torch.empty(A, H, L, device="cuda")
    _lib.check(lib.ap_randn(zr.data_ptr(), zr.numel(), seed, off, st))
    assert int(lib.ap_nes_noise_blocks(A, S, L)) == (A * H * L + 3) // 4
    assert torch.equal(out[:, first:first + H], zr * sigma + x[:, None])
    assert torch.equal(out[:, first + H:], (-zr) * sigma + x[:, None])
    if first:
        assert torch.equal(out[:, 0], x)
    g_philox, g_given = torch.empty(A, L, device="cuda"), torch.empty(A, L, device="cuda")
    _lib.check(lib.ap_nes_gradient(loss.data_ptr(), None, seed, off, first, 0.5, 0, g_philox.data_ptr(), A, S, L, st))
    _lib.check(lib.ap_nes_gradient(loss.data_ptr(), zr.data_ptr(), 0, 0, first, 0.5, 0, g_given.data_ptr(), A, S, L, st))
    assert torch.equal(g_philox, g_given)
    with pytest.raises(ap.AudioPureError):
        _lib.check(lib.ap_nes_perturb(x.data_ptr(), sigma, None, seed, off, first, out.data_ptr(), A, 7, L, st))


def test_nes_estimate_aligns_with_the_autograd_gradient(ap):
    """Property at full size: around the (deterministic) log-mel + ResNeXt system, the NES estimate from 2000 Philox
    queries of one 16000-sample clip correlates with the exact gradient of the same loss from the CUDA backward kernels --
    two independent paths to the same quantity.  For a linear loss the expected cosine is sqrt(S / (S + L)) = 0.33."""
    from audiopure_b200.blackbox import EOT, NES, QueryLoss, _PhiloxStream
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0)).set_mode("fp32")
    system = ap.AcousticSystem(classifier=rx, transform=ap.sc09_transform(), defender=None)
    x = cuda(synthetic.synthetic_waveforms(1, 16000, seed=4321))
    y = torch.tensor([3], device="cuda")
    loss_fn = QueryLoss("Entropy")
    xr = x.clone().requires_grad_(True)
    loss = loss_fn(system(xr), y)
    (g_true,) = torch.autograd.grad(loss.sum(), xr)
    loss = loss.detach()
    nes = NES(2000, 200, 0.001, EOT(system, loss_fn, 1, 1, False), stream=_PhiloxStream(seed=5))
    mean_loss, g_nes, adver_loss, adver_score, predict = nes(x, y)
    cos = float((g_nes * g_true).sum() / (g_nes.norm() * g_true.norm()))
    print(f"NES (2000 queries) vs autograd gradient: cosine {cos:.3f}, |g_nes| {float(g_nes.norm()):.3e}, |g| {float(g_true.norm()):.3e}")
    assert abs(float(adver_loss[0]) - float(loss[0])) < 1e-4 and abs(float(mean_loss[0]) - float(loss[0])) < 1e-2
    assert int(predict[0]) == int(adver_score.argmax(1)[0])
    assert cos > 0.15                                   # measured 0.25; unrelated directions give |cos| ~ 1 / sqrt(L) = 0.008


# ------------------------------------------------------------------------------------ VGG classifiers (section 8f-4)
@pytest.mark.parametrize("depth", [11, 19])
def test_vgg_vs_reference_golden(ap, golden, golden_grad, golden_vgg, depth, monkeypatch):
    """vgg11_bn / vgg19_bn (models/vgg.py:32-95; `--classifier_model vgg19_bn`, adaptive_attack_eval.py:21) on the CUDA
    path: logits, top-1 and the input gradient against the unmodified reference; a batch that spans two chunks of the
    backward (64 + 6) must reproduce the single-image results."""
    vg = ap.VGGClassifier(synthetic.vgg_state_dict(depth=depth, seed=0), depth=depth).set_mode("fp32")
    logits = vg(cuda(golden["mel_sc09"])).cpu().numpy()
    want = golden_vgg[f"vgg{depth}_logits"]
    assert np.abs(logits - want).max() < 1e-3 * max(1.0, np.abs(want).max())
    assert (logits.argmax(1) == want.argmax(1)).all()
    spec = cuda(golden_grad["resnext_in_spec"]).requires_grad_(True)
    g_logits = cuda(golden_grad["resnext_g_logits"])
    (gs,) = torch.autograd.grad(vg(spec), spec, g_logits)
    err = rel_l2(gs, golden_vgg[f"vgg{depth}_grad"])
    print(f"VGG-{depth} logits max err {np.abs(logits - want).max():.2e}, gradient rel-L2 {err:.3e}")
    assert err < 1e-4
    big = spec.detach().repeat(35, 1, 1, 1).requires_grad_(True)                # 70 images
    out = vg(big)
    (gb,) = torch.autograd.grad(out, big, g_logits.repeat(35, 1))
    assert torch.allclose(out[-2:], vg(spec.detach()), atol=1e-5) and rel_l2(gb[-2:], gs) < 1e-5
    # tf32 tensor-core convolutions (what cuDNN runs for the reference on this GPU by default): 10-bit mantissa operands
    lt = vg.set_mode("tf32")(cuda(golden["mel_sc09"])).cpu().numpy()
    print(f"VGG-{depth} tf32 logits max err {np.abs(lt - want).max():.2e}")
    assert np.abs(lt - want).max() < 2e-2 * max(1.0, np.abs(want).max()) and (lt.argmax(1) == want.argmax(1)).all()
    big_t = vg(big.detach())                                                     # 70 images: several images per tile + a ragged tail
    assert torch.allclose(big_t[-2:], torch.from_numpy(lt).cuda(), atol=1e-5) and torch.allclose(big_t[:2], big_t[-2:], atol=1e-5)
    # the tf32-mode gradient is the gradient of the tf32 network (tensor-core forward tape and data-gradient convolutions)
    (gt,) = torch.autograd.grad(vg(spec), spec, g_logits)
    (gbt,) = torch.autograd.grad(vg(big), big, g_logits.repeat(35, 1))
    err_t = rel_l2(gt, golden_vgg[f"vgg{depth}_grad"])
    monkeypatch.setenv("AP_CLS_VJP_FWD_FP32", "1")          # fp32 forward tape, tensor-core data-gradient convolutions only
    (gtb,) = torch.autograd.grad(vg(spec), spec, g_logits)
    err_tb = rel_l2(gtb, golden_vgg[f"vgg{depth}_grad"])
    print(f"VGG-{depth} gradient rel-L2: tf32 {err_t:.3e}, tf32 backward on an fp32 forward {err_tb:.3e}")
    assert err_t < 0.3 and err_tb < 2e-2 and rel_l2(gbt[-2:], gt) < 1e-5       # measured 0.13 / 6e-3 (VGG-19)


@pytest.mark.parametrize("key,depth,k", [("wrn28_10", 28, 10), ("wrn16_1", 16, 1)])
def test_wideresnet_vs_reference_golden(ap, golden, golden_grad, golden_vgg, key, depth, k, monkeypatch):
    """WideResNet-28-10 (`--classifier_model wideresnet28_10`, adaptive_attack_eval.py:21; models/wideresnet.py:15-92) and a
    narrow WRN-16-1 (equal-width first block) on the CUDA path: logits, top-1 and the input gradient against the unmodified
    reference; a batch that spans two chunks of the backward (32 + 4) must reproduce the two-image results."""
    wr = ap.WideResNetClassifier(synthetic.wideresnet_state_dict(depth=depth, widen_factor=k, seed=0), depth=depth, widen_factor=k)
    wr.set_mode("fp32")
    logits = wr(cuda(golden["mel_sc09"])).cpu().numpy()
    want = golden_vgg[f"{key}_logits"]
    assert np.abs(logits - want).max() < 1e-3 * max(1.0, np.abs(want).max())
    assert (logits.argmax(1) == want.argmax(1)).all()
    spec = cuda(golden_grad["resnext_in_spec"]).requires_grad_(True)
    g_logits = cuda(golden_grad["resnext_g_logits"])
    (gs,) = torch.autograd.grad(wr(spec), spec, g_logits)
    err = rel_l2(gs, golden_vgg[f"{key}_grad"])
    print(f"{key} logits max err {np.abs(logits - want).max():.2e} (max |logit| {np.abs(want).max():.1f}), gradient rel-L2 {err:.3e}")
    # WRN-28-10 with these random weights: one ReLU whose pre-activation changes sign under a differently rounded fp32
    # normalisation moves the gradient by ~1e-3 (oracle/audiopure_oracle.py::wideresnet_forward measured 3 flips = 2.4e-3)
    assert err < (5e-3 if key == "wrn28_10" else 1e-4)      # measured 3.0e-4 / 7.4e-7
    big = spec.detach().repeat(18, 1, 1, 1).requires_grad_(True)                # 36 images
    out = wr(big)
    (gb,) = torch.autograd.grad(out, big, g_logits.repeat(18, 1))
    assert torch.allclose(out[-2:], wr(spec.detach()), rtol=1e-5, atol=1e-4) and rel_l2(gb[-2:], gs) < 1e-5
    # tf32 tensor-core convolutions (N = 160 tiles for the 160 / 320 / 640-channel layers of WRN-28-10; WRN-16-1 has none)
    lt = wr.set_mode("tf32")(cuda(golden["mel_sc09"])).cpu().numpy()
    print(f"{key} tf32 logits max err {np.abs(lt - want).max():.2e} (max |logit| {np.abs(want).max():.1f})")
    assert np.abs(lt - want).max() < 2e-2 * max(1.0, np.abs(want).max()) and (lt.argmax(1) == want.argmax(1)).all()
    big_t = wr(big.detach())
    assert torch.allclose(big_t[-2:], torch.from_numpy(lt).cuda(), rtol=1e-5, atol=1e-4)
    # the tf32-mode gradient is the gradient of the tf32 network (tensor-core forward tape and data-gradient convolutions);
    # ReLU units whose sign differs from the fp32 forward move it by per cents (cf. the ResNeXt tf32 gradient test)
    (gt,) = torch.autograd.grad(wr(spec), spec, g_logits)
    (gbt,) = torch.autograd.grad(wr(big), big, g_logits.repeat(18, 1))
    err_t = rel_l2(gt, golden_vgg[f"{key}_grad"])
    monkeypatch.setenv("AP_CLS_VJP_FWD_FP32", "1")          # fp32 forward tape, tensor-core data-gradient convolutions only
    (gtb,) = torch.autograd.grad(wr(spec), spec, g_logits)
    err_tb = rel_l2(gtb, golden_vgg[f"{key}_grad"])
    print(f"{key} gradient rel-L2: tf32 {err_t:.3e}, tf32 backward on an fp32 forward {err_tb:.3e}")
    assert err_t < 0.5 and err_tb < 2e-2 and rel_l2(gbt[-2:], gt) < 1e-5


@pytest.mark.parametrize("key,depth", [("densenet100_12", 100), ("densenet22_12", 22)])
def test_densenet_vs_reference_golden(ap, golden, golden_grad, golden_vgg, key, depth):
    """DenseNet-BC-100-12 (`--classifier_model densenet_bc_100_12`, adaptive_attack_eval.py:21; models/densenet.py:15-147) and
    a shallow BC-22-12 whose third dense block starts at 33 channels (layout padding inside the concatenated tensor) on the
    CUDA path: logits, top-1 and the input gradient against the unmodified reference; a batch that spans two chunks of the
    backward (32 + 4) must reproduce the two-image results."""
    dn = ap.DenseNetClassifier(synthetic.densenet_state_dict(depth=depth, growth_rate=12, seed=0), depth=depth, growth_rate=12)
    logits = dn(cuda(golden["mel_sc09"])).cpu().numpy()
    want = golden_vgg[f"{key}_logits"]
    assert np.abs(logits - want).max() < 1e-3 * max(1.0, np.abs(want).max())
    assert (logits.argmax(1) == want.argmax(1)).all()
    spec = cuda(golden_grad["resnext_in_spec"]).requires_grad_(True)
    g_logits = cuda(golden_grad["resnext_g_logits"])
    (gs,) = torch.autograd.grad(dn(spec), spec, g_logits)
    err = rel_l2(gs, golden_vgg[f"{key}_grad"])
    print(f"{key} logits max err {np.abs(logits - want).max():.2e} (max |logit| {np.abs(want).max():.1f}), gradient rel-L2 {err:.3e}")
    assert err < 5e-3                      # single ReLU sign changes under different fp32 rounding, see the WideResNet test
    big = spec.detach().repeat(18, 1, 1, 1).requires_grad_(True)                # 36 images
    out = dn(big)
    (gb,) = torch.autograd.grad(out, big, g_logits.repeat(18, 1))
    assert torch.allclose(out[-2:], dn(spec.detach()), rtol=1e-5, atol=1e-4) and rel_l2(gb[-2:], gs) < 1e-5
    with pytest.raises(ap.AudioPureError):
        dn.set_mode("tf32")


@pytest.mark.parametrize("kind", ["resnext", "resnet50", "vgg19", "wrn16_4", "densenet22", "m5"])
def test_create_model_rebuilds_pickled_modules(ap, golden, tmp_path, kind):
    """create_model (audio_models/ConvNets_SpeechCommands/create_model.py:8-16): torch.load of a WHOLE pickled module,
    `.module` unwrapped when it was saved from DataParallel, rebuilt on the CUDA kernels from its state dict.  Stand-in classes
    carry the reference's class names and the attributes the loader dispatches on."""
    module, direct = pickled_classifier(kind)
    path = str(tmp_path / f"{kind}-best-acc.pth")
    torch.save(DataParallelStandIn(module) if kind in ("resnext", "vgg19") else module, path)
    clf = ap.create_model(path)
    x = cuda(synthetic.synthetic_waveforms(2, 16000, seed=3)) if kind == "m5" else cuda(golden["mel_sc09"])
    assert type(clf) is type(direct()) and torch.equal(clf(x), direct()(x))
    with pytest.raises(NotImplementedError):
        torch.save(torch.nn.Linear(2, 2), path)
        ap.create_model(path)
