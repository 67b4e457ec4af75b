"""CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/audiopure.h declares, the
host-only helpers work, the host logic mirrors the reference, compute entry points fail loudly without a GPU, and the
N > 1 certification plumbing (draw sharding + the single vote all-reduce) runs over gloo with world_size 2."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import audiopure_b200  # noqa: E402,F401
from audiopure_b200 import synthetic  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    from audiopure_b200 import build, _lib
    build.build_library()            # no-op when the in-tree .so is current; cross-compiles for sm_100a otherwise
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from audiopure_b200 import _lib
    header = open(os.path.join(ROOT, "include", "audiopure.h")).read()
    declared = set(re.findall(r"\b(ap_[a-z0-9_]+)\s*\(", header))
    declared -= {"ap_diffwave_s", "ap_mel_s", "ap_classifier_s"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"libaudiopure_b200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)   # the ctypes table covers the whole ABI
    assert lib.ap_version() == 100


def test_library_is_sm100a_only_and_self_contained():
    from audiopure_b200 import _lib
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda" not in out and "libcudart" not in out and "libtorch" not in out     # plain C ABI, static cudart
    if os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        arch = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
        assert "sm_100a" in arch and "sm_90" not in arch and "sm_80" not in arch


def test_fold_weight_norm_host_helper(lib):
    import audiopure_oracle as orc
    rng = np.random.default_rng(0)
    v = rng.standard_normal((7, 5, 3)).astype(np.float32)
    g = rng.uniform(0.5, 2, (7, 1, 1)).astype(np.float32)
    w = np.empty_like(v)
    assert lib.ap_fold_weight_norm(g.ctypes.data, v.ctypes.data, w.ctypes.data, 7, 15) == 0
    np.testing.assert_allclose(w, orc.fold_weight_norm(g, v).numpy(), rtol=2e-7, atol=0)
    assert lib.ap_fold_weight_norm(None, v.ctypes.data, w.ctypes.data, 7, 15) == -1
    assert b"bad arguments" in lib.ap_last_error()
    assert lib.ap_noise_offset_stride(3, 1001) == (3 * 1001 + 3) // 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_compute_entry_points_fail_loudly_without_gpu(lib):
    from audiopure_b200 import _lib
    import audiopure_b200 as ap
    cfg = _lib.MelCfg(16000, 2048, 512, 32, 1, 1, 0)
    h = C.c_void_p()
    assert lib.ap_mel_create(C.byref(h), C.byref(cfg), 0) == -2           # AP_ERR_CUDA: there is no CPU fallback
    assert b"no CPU fallback" in lib.ap_last_error() or b"CUDA" in lib.ap_last_error()
    with pytest.raises(ap.AudioPureError):
        ap.sc09_transform()
    with pytest.raises(ap.AudioPureError):
        ap.WaveNet(synthetic.wavenet_state_dict(seed=3, config=dict(res_channels=64, skip_channels=64, num_res_layers=2)),
                   res_channels=64, skip_channels=64, num_res_layers=2)


def test_blackbox_host_logic_and_argument_checks(lib):
    """Host side of the black-box query path (no compute): argument validation of the NES / loss entry points, the noise-block
    accounting, resolve_loss / resolve_prediction semantics (robustness_eval/_utils.py:103-136), and CPU tensors refused."""
    import audiopure_b200 as ap
    from audiopure_b200.blackbox import EOT, NES, QueryLoss, resolve_loss, resolve_prediction
    assert lib.ap_nes_noise_blocks(3, 8, 1001) == (3 * 4 * 1001 + 3) // 4
    one = C.c_void_p(16)                                                   # never dereferenced: validation comes first
    assert lib.ap_nes_perturb(one, 0.1, None, 0, 0, 1, one, 2, 7, 100, None) == -1      # odd samples_per_draw_batch
    assert b"even" in lib.ap_last_error()
    assert lib.ap_nes_perturb(None, 0.1, None, 0, 0, 1, one, 2, 8, 100, None) == -1
    assert lib.ap_nes_gradient(one, None, 0, 0, 2, 1.0, 0, one, 2, 8, 100, None) == -1  # first must be 0 or 1
    assert lib.ap_query_loss(one, one, 4, 10, 7, 0, 0.0, 0, one, None, None) == -1      # unknown loss kind
    assert lib.ap_query_loss(one, one, 0, 10, 0, 0, 0.0, 0, one, None, None) == 0       # empty batch is a no-op
    loss, sign = resolve_loss("Margin", targeted=True, task="SCR")
    assert isinstance(loss, QueryLoss) and loss.kind == 0 and sign == -1   # the reference returns cross entropy for 'SCR'
    with pytest.raises(NotImplementedError):
        resolve_loss("Entropy", task="SV")
    assert list(resolve_prediction([[3, 1, 3], [2, 5, 5, 2], [7]])) == [3, 2, 7]        # majority, first seen wins a tie
    with pytest.raises(ap.AudioPureError):
        QueryLoss("Entropy")(torch.zeros(2, 10), torch.zeros(2, dtype=torch.long))      # CPU tensors: no CPU path
    with pytest.raises(ap.AudioPureError):
        NES(8, 8, 0.001, EOT(lambda x: x, loss, 1, 1, False))(torch.zeros(1, 1, 16), torch.zeros(1, dtype=torch.long))


def test_hyperparams_and_sde_schedule_match_oracle(golden):
    import audiopure_oracle as orc
    from audiopure_b200.diffwave import calc_diffusion_hyperparams
    from audiopure_b200.diffwave_sde import euler_schedule
    hp = calc_diffusion_hyperparams(200, 1e-4, 0.02)
    for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"):
        assert np.array_equal(hp[k].numpy(), golden["hp_" + k]), k           # bit-exact vs the reference's tables
    for t_star in (1, 2, 6, 7, 10):
        a, b = euler_schedule(t_star), orc.sde_euler_schedule(t_star)
        assert len(a) == len(b) == t_star
        for (s1, d1), (s2, d2) in zip(a, b):
            assert float(s1) == float(s2) and float(d1) == float(d2)


def test_wavenet_weight_list_layout(lib):
    from audiopure_b200.diffwave import wavenet_weight_list
    cfg = dict(synthetic.DEFAULT_WAVENET_CONFIG, res_channels=64, skip_channels=64, num_res_layers=3, dilation_cycle=2)
    sd = synthetic.wavenet_state_dict(seed=1, config=cfg)
    ws = wavenet_weight_list(sd, cfg)
    assert len(ws) == 6 + 8 * 3 + 4
    assert ws[0].shape == (64,) and ws[2].shape == (512, 128) and ws[6 + 2].shape == (128, 64, 3) and ws[6 + 4].shape == (64, 64)
    assert ws[-2].shape == (64,) and ws[-1].shape == (1,)
    v, g = sd["residual_layer.residual_blocks.0.res_conv.weight_v"], sd["residual_layer.residual_blocks.0.res_conv.weight_g"]
    want = v * (g / np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))).astype(np.float32)
    np.testing.assert_allclose(ws[6 + 4], want.reshape(64, 64), rtol=3e-7)


def test_certification_statistics_and_sharding():
    from audiopure_b200.certify import RobustCertificate, shard_draws
    rc = RobustCertificate.__new__(RobustCertificate)
    assert abs(rc.lower_conf_bound(99000, 100000) - 0.988989) < 1e-5          # KATs, SURVEY.md section 8 a21
    assert abs(rc.lower_conf_bound(100000, 100000) - 0.999931) < 1e-5
    assert rc.lower_conf_bound(50200, 100000) < 0.5 and rc.lower_conf_bound(0, 100) == 0.0
    for n in (0, 1, 7, 100, 100000, 100003):
        for world in (1, 2, 3, 8):
            parts = [shard_draws(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, n, K, out_q):
    import torch.distributed as dist
    from audiopure_b200.certify import reduce_counts, shard_draws
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # every rank votes on its own shard of the SAME deterministic prediction stream
        preds = torch.from_numpy(np.random.default_rng(123).integers(0, K, size=n))
        a, b = shard_draws(n, world, rank)
        counts = torch.bincount(preds[a:b], minlength=K).to(torch.int64)
        reduce_counts(counts)                     # the single collective of the certification path
        out_q.put((rank, counts.tolist()))
    finally:
        dist.destroy_process_group()


def test_vote_allreduce_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, K, world, port = 1001, 10, 2, _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, K, q)) for r in range(world)]
    [p.start() for p in procs]
    results = dict(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    want = np.bincount(np.random.default_rng(123).integers(0, K, size=n), minlength=K).tolist()
    assert results[0] == want and results[1] == want


def test_philox_keys_are_distinct_per_consumer_and_per_object():
    """a defender's diffusion noise, the smoothing noise and an attacker's NES probes must not walk the same Philox
    blocks for equal user seeds; default-seeded objects get their own streams."""
    from audiopure_b200 import _lib
    keys = {c: _lib.philox_key(c, 0) for c in ("diffwave", "certify", "nes")}
    assert len(set(keys.values())) == 3 and all(0 <= k < 2 ** 64 for k in keys.values())
    assert _lib.philox_key("diffwave", 7) == _lib.philox_key("diffwave", 7) != _lib.philox_key("diffwave", 8)
    a, b = _lib.philox_key("diffwave", None), _lib.philox_key("diffwave", None)
    assert a != b


def test_unet_structure_matches_the_reference_walk():
    """synthetic.unet_structure reproduces UNetModel.__init__'s module walk: 446 state-dict tensors for the default configuration
    (checked key by key against the reference in tests/golden/make_golden_unet.py), a balanced skip-connection stack, 30 ResBlocks
    and 15 attention blocks."""
    from audiopure_b200 import synthetic
    ops, cfg = synthetic.unet_structure()
    kinds = [k for _, k, _, _ in ops]
    assert kinds.count("res") == 30 and kinds.count("attn") == 15 and kinds.count("down") == 3 and kinds.count("up") == 3
    assert kinds.count("push") + 1 == kinds.count("pop") == 16          # conv_in pushes implicitly
    assert len(synthetic.unet_state_dict(seed=0)) == 446
    depth = 1
    for k in kinds:
        depth += (k == "push") - (k == "pop")
        assert depth >= 0
    assert depth == 0


def test_spec_and_wave_euler_schedules():
    from audiopure_b200.diffwave_sde import euler_schedule
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import audiopure_oracle as orc
    for t in (1, 3, 7, 10):
        a, b = euler_schedule(t), orc.sde_euler_schedule(t)
        assert len(a) == len(b) == t and all(float(x[0]) == float(y[0]) and float(x[1]) == float(y[1]) for x, y in zip(a, b))
    s = orc.spec_sde_schedule(2)
    assert len(s) == 2 and abs(float(s[0][1]) - 1e-3) < 1e-7
