"""Pins the CPU oracle (oracle/audiopure_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/reference_golden.npz, produced by tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

import audiopure_b200  # noqa: F401
from audiopure_b200 import synthetic

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import audiopure_oracle as orc  # noqa: E402


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="module")
def sd_full():
    return synthetic.wavenet_state_dict(seed=0)


@pytest.fixture(scope="module")
def hp():
    return orc.diffusion_hyperparams(200, 1e-4, 0.02)


def noise_list(seed, n, shape):
    return orc.NoiseSource([synthetic.host_noise(shape, seed, i) for i in range(n)])


def test_hyperparams_bit_exact(golden, hp):
    for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"):
        assert np.array_equal(hp[k].numpy(), golden["hp_" + k]), k
    # KATs quoted in SURVEY.md §8 a1
    np.testing.assert_allclose(hp["Alpha_bar"][:3].numpy(), [0.99989998, 0.99970001, 0.99940014], rtol=0, atol=1e-8)
    np.testing.assert_allclose(hp["Sigma"][:3].numpy(), [0.01, 0.00816578, 0.01224867], rtol=0, atol=1e-8)


def test_step_embedding(golden):
    e = orc.step_embedding(torch.from_numpy(golden["emb_steps"]), 128).numpy()
    np.testing.assert_allclose(e, golden["emb"], rtol=0, atol=1e-6)
    assert abs(e[1, 0] - 0.84147) < 1e-5 and abs(e[1, 64] - 0.54030) < 1e-5


@pytest.mark.parametrize("key,L,B,t,seed", [("eps_full_L1024_t1", 1024, 2, 1.0, 1234),
                                            ("eps_full_L1024_t65", 1024, 2, 65.0, 1234),
                                            ("eps_full_L3001_t7", 3001, 1, 7.0, 77)])
def test_wavenet_full(golden, sd_full, key, L, B, t, seed):
    x = synthetic.synthetic_waveforms(B, L, seed=seed)
    eps = orc.wavenet_forward(sd_full, x, t * torch.ones(B, 1)).numpy()
    assert rel_l2(eps, golden[key]) < 2e-5


def test_wavenet_small_config(golden):
    cfg = dict(synthetic.DEFAULT_WAVENET_CONFIG, res_channels=64, skip_channels=64, num_res_layers=5, dilation_cycle=3)
    sd = synthetic.wavenet_state_dict(seed=3, config=cfg)
    x = synthetic.synthetic_waveforms(3, 500, seed=5)
    eps = orc.wavenet_forward(sd, x, 3.0 * torch.ones(3, 1), num_res_layers=5, dilation_cycle=3).numpy()
    assert rel_l2(eps, golden["eps_small_L500_t3"]) < 2e-5


def test_ddpm_purifier(golden, sd_full, hp):
    x = synthetic.synthetic_waveforms(2, 1024, seed=1234)
    y2 = orc.ddpm_forward(sd_full, x, hp, 2, noise_list(2024, 2, x.shape)).numpy()
    assert rel_l2(y2, golden["ddpm_t2_L1024"]) < 1e-5
    y3 = orc.ddpm_forward(sd_full, x, hp, 3, noise_list(2025, 3, x.shape)).numpy()
    assert rel_l2(y3, golden["ddpm_t3_L1024"]) < 1e-5


def test_one_two_shot_and_fast_reverse(golden, sd_full, hp):
    x = synthetic.synthetic_waveforms(2, 1024, seed=1234)
    assert rel_l2(orc.one_shot_denoise(sd_full, x, hp, 66).numpy(), golden["oneshot_t66_L1024"]) < 1e-5
    assert rel_l2(orc.two_shot_denoise(sd_full, x, hp, 66).numpy(), golden["twoshot_t66_L1024"]) < 1e-5
    y = orc.fast_reverse(sd_full, x, hp, 9, noise_list(2026, 3, x.shape)).numpy()
    assert rel_l2(y, golden["fastrev_t9_L1024"]) < 1e-5


def test_sde_drift_and_diffusion(golden, sd_full):
    """RevVPSDE.f / .g (diffwave_sde.py:117-133) at fixed solver times; x is (B, L) flattened audio."""
    x = torch.from_numpy(synthetic.synthetic_waveforms(2, 1024, seed=1234))
    tab = orc.sde_tables(200, 0.0001 * 200, 0.02 * 200)
    for i, s in enumerate(golden["sde_s"]):
        c = orc.sde_step_coefficients(tab, torch.tensor(s, dtype=torch.float32))
        eps = orc.wavenet_forward(sd_full, x, c["d"] * torch.ones(2, 1))
        f = -(-0.5 * c["beta"] * x - c["beta"] * (-eps / c["sqrt_1m_ab"]))
        assert rel_l2(f.reshape(2, -1).numpy(), golden["sde_f"][i]) < 1e-5, s
        np.testing.assert_allclose(np.full(2, float(c["g"])), golden["sde_g"][i], rtol=1e-6, atol=1e-8)


def test_sde_schedule_indices():
    """Float32 step-index hazard, SURVEY.md §8 a13: d = t*-1..0 for t* <= 6, t*..1 for t* >= 7."""
    tab = orc.sde_tables(200, 0.0001 * 200, 0.02 * 200)
    for t_star in range(1, 11):
        sched = orc.sde_euler_schedule(t_star, 200)
        ds = [orc.sde_step_coefficients(tab, s)["d"] for s, _ in sched]
        assert len(sched) == t_star
        expect = list(range(t_star - 1, -1, -1)) if t_star <= 6 else list(range(t_star, 0, -1))
        assert ds == expect, (t_star, ds)
        assert abs(float(sched[-1][1]) - 0.00499) < 2e-5


def test_sde_equals_ddpm_for_small_t(sd_full, hp):
    """With shared noise the Euler chain matches the DDPM chain to first order for t* <= 6."""
    cfg = dict(synthetic.DEFAULT_WAVENET_CONFIG, res_channels=64, skip_channels=64, num_res_layers=5, dilation_cycle=3)
    sd = synthetic.wavenet_state_dict(seed=3, config=cfg)
    kw = dict(num_res_layers=5, dilation_cycle=3)
    x = synthetic.synthetic_waveforms(2, 400, seed=5)
    a = orc.ddpm_forward(sd, x, hp, 3, noise_list(7, 3, x.shape), **kw).numpy()
    b = orc.sde_purify(sd, x, 3, noise_list(7, 4, x.shape), **kw).numpy()
    assert rel_l2(b, a) < 2e-3


def test_mel_front_ends(golden):
    x = synthetic.synthetic_waveforms(2, 16000, seed=99)
    sc = orc.mel_db(x, **orc.MEL_SC09).numpy()
    assert sc.shape == (2, 1, 32, 32)
    assert np.abs(sc - golden["mel_sc09"]).max() < 2e-3     # dB; torchaudio runs an fp32 FFT
    kw = orc.mel_db(x, **orc.MEL_KWS).numpy()
    assert kw.shape == (2, 1, 32, 81)
    assert np.abs(kw - golden["mel_kws"]).max() < 2e-3
    fb = orc.mel_filterbank(1025, 0.0, 8000.0, 32, 16000, "slaney", "slaney")
    np.testing.assert_allclose(fb, golden["mel_sc09_fb"], rtol=1e-4, atol=1e-7)
    fbk = orc.mel_filterbank(201, 0.0, 8000.0, 32, 16000, None, "htk")
    np.testing.assert_allclose(fbk, golden["mel_kws_fb"], rtol=1e-4, atol=5e-6)


def test_classifiers(golden):
    rx = orc.resnext_forward(synthetic.resnext_state_dict(seed=0), golden["mel_sc09"]).numpy()
    np.testing.assert_allclose(rx, golden["resnext_logits"], rtol=0, atol=2e-4)
    assert np.array_equal(rx.argmax(1), golden["resnext_logits"].argmax(1))
    x = synthetic.synthetic_waveforms(2, 16000, seed=99)
    m5 = orc.m5_forward(synthetic.m5_state_dict(seed=0), x).numpy()
    np.testing.assert_allclose(m5, golden["m5_logprobs"], rtol=0, atol=1e-4)
    kws = orc.kws_forward(synthetic.kws_state_dict(seed=0), golden["mel_kws"]).numpy()
    np.testing.assert_allclose(kws, golden["kws_logprobs"], rtol=0, atol=1e-4)


def test_acoustic_system_end_to_end(golden, sd_full, hp):
    x = synthetic.synthetic_waveforms(1, 16000, seed=1234)
    pur = orc.ddpm_forward(sd_full, orc.acoustic_rescale(x), hp, 2, noise_list(2027, 2, x.shape))
    assert rel_l2(pur.numpy(), golden["system_purified"]) < 1e-5
    rsd = synthetic.resnext_state_dict(seed=0)
    logits = orc.resnext_forward(rsd, orc.mel_db(pur, **orc.MEL_SC09)).numpy()
    np.testing.assert_allclose(logits, golden["system_logits"], rtol=0, atol=2e-3)
    assert logits.argmax(1)[0] == golden["system_logits"].argmax(1)[0]
    raw = orc.resnext_forward(rsd, orc.mel_db(orc.acoustic_rescale(x * 2 ** 15), **orc.MEL_SC09)).numpy()
    np.testing.assert_allclose(raw, golden["system_logits_int16_nodefense"], rtol=0, atol=2e-3)


def test_certification(golden, sd_full, hp):
    assert [orc.compute_t_star(hp, s) for s in (0.25, 0.5, 1.0)] == list(golden["smooth_t_star"]) == [34, 66, 117]
    x = synthetic.synthetic_waveforms(1, 16000, seed=1234)
    rsd = synthetic.resnext_state_dict(seed=0)

    def logits_fn(x_in, t_star):
        x0 = orc.one_shot_denoise(sd_full, x_in, hp, t_star)
        return orc.resnext_forward(rsd, orc.mel_db(x0, **orc.MEL_SC09))

    shapes = [(5, 1, 16000), (5, 1, 16000), (2, 1, 16000)]
    noise = orc.NoiseSource([synthetic.host_noise(s, 2028, i) for i, s in enumerate(shapes)])
    counts = orc.smooth_counts(logits_fn, x, 12, 0.5, 5, noise, 10, hp)
    assert np.array_equal(counts, golden["smooth_counts_n12"])
    # Clopper-Pearson KATs (scipy), SURVEY.md §8 a21
    assert abs(orc.lower_conf_bound(99000, 100000) - 0.988989) < 1e-6
    y, r = orc.certify_from_counts([0, 5, 95], [0, 1000, 99000], 100000, 0.5)
    assert y == 2 and abs(r - 1.1450) < 1e-3
    y, r = orc.certify_from_counts([0, 100], [0, 100000], 100000, 0.5)
    assert y == 1 and abs(r - 1.9057) < 1e-3
    assert orc.certify_from_counts([60, 40], [50200, 49800], 100000, 0.5) == (-1, 0.0)


@pytest.mark.parametrize("depth", [34, 50])
def test_resnet_family(golden, depth):
    sd = synthetic.resnet_state_dict(depth=depth, seed=0)
    logits = orc.resnet_forward(sd, golden["mel_sc09"], depth=depth).numpy()
    want = golden[f"resnet{depth}_logits"]
    assert np.abs(logits - want).max() < 1e-4 * max(1.0, np.abs(want).max())


def test_reffwave(golden, sd_full, hp):
    x = synthetic.synthetic_waveforms(2, 1024, seed=1234)
    y = orc.reffwave_forward(sd_full, x, hp, 5, 2, noise_list(2029, 2, x.shape)).numpy()
    assert rel_l2(y, golden["reffwave_t5_re2_L1024"]) < 1e-5


# ---------------------------------------------------------------------------------------------------- gradients
def test_oracle_autograd_matches_reference_gradients(golden_grad, sd_full, hp):
    """The oracle is differentiable (torch ops); its autograd gradients are the checker of the CUDA backward pass
    (ap_diffwave_eps_vjp).  Pinned here to gradients of the unmodified reference (tests/golden/make_golden_grad.py)."""
    x = torch.from_numpy(synthetic.synthetic_waveforms(2, 1024, seed=1234))
    g = torch.from_numpy(golden_grad["vjp_g_eps"])
    xr = x.clone().requires_grad_(True)
    eps = orc.wavenet_forward(sd_full, xr, 7.0 * torch.ones(2, 1))
    (gx,) = torch.autograd.grad(eps, xr, g)
    assert rel_l2(gx.numpy(), golden_grad["vjp_gx_L1024_t7"]) < 2e-5
    # whole purifier: d <w, DiffWave.forward(x)> / d x with the reference's noise order
    xr = x.clone().requires_grad_(True)
    y = orc.ddpm_forward(sd_full, xr, hp, 2, noise_list(2024, 2, (2, 1, 1024)))
    assert rel_l2(y.detach().numpy(), golden_grad["ddpm_t2_purified"]) < 1e-5
    (gx,) = torch.autograd.grad((y * torch.from_numpy(golden_grad["ddpm_grad_w"])).sum(), xr)
    assert rel_l2(gx.numpy(), golden_grad["ddpm_t2_grad_L1024"]) < 2e-5


def test_oracle_mel_and_resnext_gradients_match_reference(golden_grad):
    """torchaudio / CifarResNeXt autograd gradients (golden) vs autograd over the oracle's restatements."""
    xm = torch.from_numpy(synthetic.synthetic_waveforms(2, 16000, seed=99))
    for name, kw in (("sc09", orc.MEL_SC09), ("kws", orc.MEL_KWS)):
        xr = xm.clone().requires_grad_(True)
        (gx,) = torch.autograd.grad(orc.mel_db(xr, **kw), xr, torch.from_numpy(golden_grad[f"mel_{name}_g_spec"]))
        assert rel_l2(gx.numpy(), golden_grad[f"mel_{name}_grad"]) < 1e-4, name
    sr = torch.from_numpy(golden_grad["resnext_in_spec"]).requires_grad_(True)
    logits = orc.resnext_forward(synthetic.resnext_state_dict(seed=0), sr)
    (gs,) = torch.autograd.grad(logits, sr, torch.from_numpy(golden_grad["resnext_g_logits"]))
    assert rel_l2(gs.numpy(), golden_grad["resnext_grad"]) < 1e-4


def test_query_losses_match_reference(golden_blackbox):
    g = golden_blackbox
    s, y = g["loss_scores"], g["loss_labels"]
    np.testing.assert_allclose(orc.query_loss(s, y, "Entropy").numpy(), g["loss_entropy"], rtol=0, atol=2e-6)
    for targeted in (0, 1):
        for clip in (0, 1):
            out = orc.query_loss(s, y, "Margin", bool(targeted), 0.5, bool(clip)).numpy()
            assert np.array_equal(out, g[f"loss_margin_t{targeted}_c{clip}"])


@pytest.mark.parametrize("name", ["a", "b"])
def test_nes_estimator_matches_reference(golden_blackbox, name):
    """oracle nes_gradient / eot_scores vs robustness_eval._NES.NES + _EOT.EOT around the same stand-in model."""
    from gpu_common import ToyDefendedModel
    g = golden_blackbox
    A, L, spd, bs, es, eb = (int(v) for v in g[f"nes_{name}_cfg"])
    sigma = float(g[f"nes_{name}_sigma"])
    draws = iter(range(1000))
    mean_loss, grad, adver_loss, adver_score, predict, queries = orc.nes_gradient(
        ToyDefendedModel(L), lambda s, y: orc.query_loss(s, y, "Entropy"), g[f"nes_{name}_x"], g[f"nes_{name}_y"], spd, bs,
        sigma, lambda shape: synthetic.host_noise(shape, 4000, next(draws)), es, eb)
    for i, q in enumerate(queries):
        assert np.array_equal(q.numpy(), g[f"nes_{name}_queries{i}"])
    np.testing.assert_allclose(adver_score.numpy(), g[f"nes_{name}_adver_score"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(adver_loss.numpy(), g[f"nes_{name}_adver_loss"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(mean_loss.numpy(), g[f"nes_{name}_mean_loss"], rtol=0, atol=2e-6)
    assert np.array_equal(predict, g[f"nes_{name}_predict"])
    # the estimate divides loss differences of ~sigma by sigma: 1e-7 of loss rounding is ~1e-4 of the gradient
    assert rel_l2(grad.numpy(), g[f"nes_{name}_grad"]) < 2e-3


@pytest.mark.parametrize("depth", [11, 19])
def test_vgg_family(golden, golden_grad, golden_vgg, depth):
    """oracle vgg_forward (logits and autograd input gradient) vs the reference's vgg11_bn / vgg19_bn (models/vgg.py)."""
    sd = synthetic.vgg_state_dict(depth=depth, seed=0)
    logits = orc.vgg_forward(sd, golden["mel_sc09"], depth=depth).numpy()
    want = golden_vgg[f"vgg{depth}_logits"]
    assert np.abs(logits - want).max() < 1e-4 * max(1.0, np.abs(want).max())
    spec = torch.from_numpy(golden_grad["resnext_in_spec"]).clone().requires_grad_(True)
    (gs,) = torch.autograd.grad(orc.vgg_forward(sd, spec, depth=depth), spec, torch.from_numpy(golden_grad["resnext_g_logits"]))
    assert rel_l2(gs.numpy(), golden_vgg[f"vgg{depth}_grad"]) < 1e-4


@pytest.mark.parametrize("key,depth,k", [("wrn28_10", 28, 10), ("wrn16_1", 16, 1)])
def test_wideresnet_family(golden, golden_grad, golden_vgg, key, depth, k):
    """oracle wideresnet_forward (logits and autograd input gradient) vs the reference's WideResNet (models/wideresnet.py)."""
    sd = synthetic.wideresnet_state_dict(depth=depth, widen_factor=k, seed=0)
    logits = orc.wideresnet_forward(sd, golden["mel_sc09"], depth=depth, widen_factor=k).numpy()
    want = golden_vgg[f"{key}_logits"]
    assert np.abs(logits - want).max() < 1e-4 * max(1.0, np.abs(want).max())
    spec = torch.from_numpy(golden_grad["resnext_in_spec"]).clone().requires_grad_(True)
    (gs,) = torch.autograd.grad(orc.wideresnet_forward(sd, spec, depth=depth, widen_factor=k), spec,
                                torch.from_numpy(golden_grad["resnext_g_logits"]))
    assert rel_l2(gs.numpy(), golden_vgg[f"{key}_grad"]) < 1e-4


@pytest.mark.parametrize("key,depth", [("densenet100_12", 100), ("densenet22_12", 22)])
def test_densenet_family(golden, golden_grad, golden_vgg, key, depth):
    """oracle densenet_forward (logits and autograd input gradient) vs the reference's DenseNet-BC (models/densenet.py)."""
    sd = synthetic.densenet_state_dict(depth=depth, growth_rate=12, seed=0)
    logits = orc.densenet_forward(sd, golden["mel_sc09"], depth=depth).numpy()
    want = golden_vgg[f"{key}_logits"]
    assert np.abs(logits - want).max() < 1e-4 * max(1.0, np.abs(want).max())
    spec = torch.from_numpy(golden_grad["resnext_in_spec"]).clone().requires_grad_(True)
    (gs,) = torch.autograd.grad(orc.densenet_forward(sd, spec, depth=depth), spec, torch.from_numpy(golden_grad["resnext_g_logits"]))
    assert rel_l2(gs.numpy(), golden_vgg[f"{key}_grad"]) < 1e-4


# ================================================================================================ golden pack v2 (BASELINE shapes)
def test_v2_wavenet_at_benchmark_length(golden_v2, sd_full):
    """eps at L = 16 000 (every dilation up to 2048 reads real samples, not only padding), t = 65."""
    x = synthetic.synthetic_waveforms(2, 16000, seed=1234)
    eps = orc.wavenet_forward(sd_full, x, 65.0 * torch.ones(2, 1)).numpy()
    assert rel_l2(eps, golden_v2["eps_L16000_t65"]) < 2e-5


def test_v2_one_shot_on_smoothing_level_input(golden_v2, sd_full, hp):
    """certified_robust.py:44-54 at sigma = 1.0 (t* = 117): x0_hat of sqrt(abar*) (x + sigma z)."""
    sigma, t_star = 1.0, 117
    assert orc.compute_t_star(hp, sigma) == t_star
    x1 = torch.from_numpy(synthetic.synthetic_waveforms(2, 16000, seed=1234))[0:1]
    z = torch.from_numpy(synthetic.host_noise((4, 1, 16000), 3000 + int(sigma * 100), 0))
    x_in = (1 / (1 + sigma ** 2)) ** 0.5 * (x1.repeat(4, 1, 1) + z * sigma)
    x0 = orc.one_shot_denoise(sd_full, x_in[:2], hp, t_star).numpy()
    assert rel_l2(x0, golden_v2["smooth_x0_sigma1.0"]) < 1e-5


def test_v2_classifiers_on_purified_clips(golden_v2, trained_checkpoints):
    """mel -> ResNeXt (centred bias) and the TRAINED M5 on the reference's purified clips: logits, top-1."""
    pur = golden_v2["top1_purified_first2"]
    rx = synthetic.resnext_state_dict(seed=0)
    rx["classifier.bias"] = golden_v2["resnext_centred_bias"]
    lg = orc.resnext_forward(rx, orc.mel_db(pur, **orc.MEL_SC09)).numpy()
    assert np.abs(lg - golden_v2["top1_logits"][:2]).max() < 2e-3
    assert (lg.argmax(1) == golden_v2["top1_logits"][:2].argmax(1)).all()
    lp = orc.m5_forward(trained_checkpoints["m5"], pur).numpy()
    assert np.abs(lp - golden_v2["top1_m5_logprobs"][:2]).max() < 1e-3
    assert len(set(golden_v2["top1_logits"].argmax(1).tolist())) >= 5        # the 32-clip top-1 golden is not unanimous
    assert len(set(golden_v2["top1_m5_logprobs"].argmax(1).tolist())) >= 5


def test_v2_trained_m5_and_kws_at_two_seconds(golden_v2, trained_checkpoints):
    xk = synthetic.synthetic_waveforms(16, 32000, seed=555)
    mel = orc.mel_db(xk, **orc.MEL_KWS).numpy()
    assert mel.shape == (16, 1, 32, 161)
    assert np.abs(mel[:2] - golden_v2["kws2s_mel"]).max() < 2e-3
    lp = orc.kws_forward(trained_checkpoints["kws"], mel).numpy()
    assert np.abs(lp - golden_v2["kws2s_logprobs"]).max() < 2e-3
    assert (lp.argmax(1) == golden_v2["kws2s_logprobs"].argmax(1)).all()
    x5 = synthetic.synthetic_waveforms(16, 16000, seed=556)
    l5 = orc.m5_forward(trained_checkpoints["m5"], x5).numpy()
    assert np.abs(l5 - golden_v2["m5_trained_logprobs"]).max() < 1e-3
    assert (l5.argmax(1) == golden_v2["m5_trained_logprobs"].argmax(1)).all()


def test_create_model_unpickles_the_reference_m5_checkpoint(golden_v2):
    """create_model.py:8-16 loads a WHOLE pickled module; the vendored file is the reference's own M5 checkpoint.  Here (CPU) only
    the unpickle + state-dict extraction is exercised; the -m gpu twin runs it through classifiers.create_model."""
    import types
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "m5_k160_vanilla_best_acc.pth")
    mod = types.ModuleType("M5Net")
    mod.M5 = type("M5", (torch.nn.Module,), {})
    sys.modules["M5Net"] = mod
    try:
        model = torch.load(path, map_location="cpu", weights_only=False)
    finally:
        del sys.modules["M5Net"]
    assert type(model).__name__ == "M5" and model.conv1.kernel_size == (160,) and model.conv1.stride == (16,)
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    x5 = synthetic.synthetic_waveforms(16, 16000, seed=556)
    assert np.abs(orc.m5_forward(sd, x5).numpy() - golden_v2["m5_trained_logprobs"]).max() < 1e-3


# ================================================================================================ SDE purifier vs the reference's RevDiffWave
def _noise(seed, n, shape):
    return orc.NoiseSource([synthetic.host_noise(shape, seed, i) for i in range(n)])


def test_sde_oracle_vs_reference_revdiffwave(golden_sde, sd_full):
    """value and gradient of the reference's RevDiffWave.audio_editing_sample (its f / g / level / chaining code; only the Euler
    stepping loop is a restatement): t* = 3, sample_step = 1.  eps is a constant of the differentiation (no_grad in the reference)."""
    x = torch.from_numpy(synthetic.synthetic_waveforms(1, 16000, seed=21)).requires_grad_(True)
    w = torch.from_numpy(golden_sde["w"])
    y = orc.sde_purify(sd_full, x, 3, _noise(3103, int(golden_sde["t3_noise_draws"]), (1, 1, 16000)))
    assert rel_l2(y.detach().numpy(), golden_sde["t3_out"]) < 1e-5
    (g,) = torch.autograd.grad((y * w[:1]).sum(), x)
    assert rel_l2(g.numpy(), golden_sde["t3_grad"]) < 1e-6


def test_sde_oracle_sample_step_and_rand_t(golden_sde, sd_full):
    x = torch.from_numpy(synthetic.synthetic_waveforms(1, 16000, seed=21)).requires_grad_(True)
    w = torch.from_numpy(golden_sde["w"])
    y = orc.sde_purify(sd_full, x, 2, _noise(3202, int(golden_sde["ss2_noise_draws"]), (1, 1, 16000)), sample_step=2)
    assert y.shape == (2, 1, 16000) and rel_l2(y.detach().numpy(), golden_sde["ss2_out"]) < 1e-5
    (g,) = torch.autograd.grad((y * w).sum(), x)
    assert rel_l2(g.numpy(), golden_sde["ss2_grad"]) < 1e-6
    with torch.no_grad():
        yr = orc.sde_purify(sd_full, x.detach(), 4, _noise(3304, 5, (1, 1, 16000)), noise_level=int(golden_sde["randt_level"]))
    assert rel_l2(yr.numpy(), golden_sde["randt_out"]) < 1e-5


# ================================================================================================ spectrogram-domain purifier (Diffusion-Spec)
def test_unet_oracle_vs_reference(golden_unet):
    ops, cfg = synthetic.unet_structure()
    sd = synthetic.unet_state_dict(seed=0)
    x = golden_unet["unet_x"]
    with torch.no_grad():
        e37 = orc.unet_forward(sd, x, 37 * torch.ones(3), ops, cfg).numpy()
        e1 = orc.unet_forward(sd, x[:1], torch.ones(1), ops, cfg).numpy()
    assert rel_l2(e37, golden_unet["unet_eps_t37"]) < 1e-5
    assert rel_l2(e1, golden_unet["unet_eps_t1"]) < 1e-5


def test_spec_sde_oracle_vs_reference_revimproveddiffusion(golden_unet):
    ops, cfg = synthetic.unet_structure()
    sd = synthetic.unet_state_dict(seed=0)
    n = int(golden_unet["spec_noise_draws"])
    with torch.no_grad():
        y = orc.spec_sde_purify(sd, golden_unet["spec_in"], 2, _noise(5300, n, (2, 1, 32, 32)), ops, cfg).numpy()
    assert rel_l2(y, golden_unet["spec_purified_t2"]) < 1e-5
    n2 = int(golden_unet["spec_noise_draws_s2"])        # sample_step = 2: round 2 starts from the de-standardised output of round 1
    with torch.no_grad():
        y2 = orc.spec_sde_purify(sd, golden_unet["spec_in"], 1, _noise(5310, n2, (2, 1, 32, 32)), ops, cfg, sample_step=2).numpy()
    assert y2.shape == (4, 1, 32, 32) and rel_l2(y2, golden_unet["spec_purified_t1_s2"]) < 1e-5


def test_unet_oracle_gradients_vs_reference(golden_unet):
    """the oracle is differentiable: its autograd gradients through the UNet and through the spectrogram purifier (which, unlike
    the waveform purifier, the reference does NOT wrap in no_grad) match the reference's"""
    ops, cfg = synthetic.unet_structure()
    sd = synthetic.unet_state_dict(seed=0)
    x = torch.from_numpy(golden_unet["unet_x"]).requires_grad_(True)
    (gx,) = torch.autograd.grad(orc.unet_forward(sd, x, 37 * torch.ones(3), ops, cfg), x, torch.from_numpy(golden_unet["unet_g_eps"]))
    assert rel_l2(gx.numpy(), golden_unet["unet_vjp_t37"]) < 1e-4
    s = torch.from_numpy(golden_unet["spec_in"]).requires_grad_(True)
    y = orc.spec_sde_purify(sd, s, 2, _noise(5300, int(golden_unet["spec_noise_draws"]), (2, 1, 32, 32)), ops, cfg)
    (gs,) = torch.autograd.grad((y * torch.from_numpy(golden_unet["spec_grad_w"])).sum(), s)
    assert rel_l2(gs.numpy(), golden_unet["spec_purified_grad_t2"]) < 1e-4
