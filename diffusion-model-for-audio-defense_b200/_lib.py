"""ctypes binding of libaudiopure_b200.so (C ABI declared in include/audiopure.h).

There is no CPU fallback: if the shared library is missing the import of any compute module fails loudly, and on a
machine without an sm_100 GPU every compute entry point returns AP_ERR_CUDA, which is raised as ``AudioPureError``.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# AP_LIB_PATH: development aid (devtools/ab_compare.py loads two builds of the library back to back on one GPU box)
LIB_PATH = os.environ.get("AP_LIB_PATH") or os.path.join(_PKG_DIR, "libaudiopure_b200.so")

AP_MODE_BF16, AP_MODE_FP32, AP_MODE_TF32, AP_MODE_FP16, AP_MODE_BF16X3 = 0, 1, 2, 3, 4
AP_CLS_RESNEXT, AP_CLS_M5, AP_CLS_KWS, AP_CLS_RESNET, AP_CLS_VGG, AP_CLS_WRN, AP_CLS_DENSENET = 0, 1, 2, 3, 4, 5, 6


class AudioPureError(RuntimeError):
    pass


class WavenetCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("in_channels", "res_channels", "skip_channels", "out_channels", "num_res_layers",
                                      "dilation_cycle", "embed_dim_in", "embed_dim_mid", "embed_dim_out")]


class SdeCoef(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("beta", "diff2", "sqrt_1mab", "dt", "g", "sqrt_dt")]


class MelCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("sample_rate", "n_fft", "hop_length", "n_mels", "slaney_norm", "slaney_scale",
                                      "reflect_pad")]


class ClassifierCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("kind", "num_classes", "cardinality", "depth", "base_width", "widen_factor",
                                      "in_channels", "m5_first_kernel", "m5_stride", "m5_channels", "kws_in_size",
                                      "kws_hidden")]


class UNetCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("image_size", "in_channels", "model_channels", "out_channels", "num_res_blocks", "num_heads",
                                      "use_scale_shift_norm")]


_vp, _fp, _i, _f, _u64 = C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint64   # device pointers travel as void*
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes): every symbol include/audiopure.h declares
SIGNATURES = {
    "ap_last_error": (C.c_char_p, []),
    "ap_version": (_i, []),
    "ap_launch_count": (C.c_ulonglong, []),
    "ap_alloc_generation": (C.c_ulonglong, []),
    "ap_fold_weight_norm": (_i, [_vp, _vp, _vp, _i, _i]),
    "ap_diffwave_create": (_i, [_PP, C.POINTER(WavenetCfg), _PP, _i, _i]),
    "ap_diffwave_destroy": (None, [_vp]),
    "ap_diffwave_set_mode": (_i, [_vp, _i]),
    "ap_diffwave_get_mode": (_i, [_vp]),
    "ap_diffwave_reserve": (_i, [_vp, _i, _i]),
    "ap_diffwave_eps": (_i, [_vp, _fp, _f, _fp, _i, _i, _vp]),
    "ap_diffwave_eps_vjp": (_i, [_vp, _fp, _f, _fp, _fp, _fp, _i, _i, _vp]),
    "ap_diffwave_eps_save": (_i, [_vp, _fp, _f, _fp, _i, _i, _vp, C.POINTER(C.c_ulonglong)]),
    "ap_diffwave_eps_vjp_saved": (_i, [_vp, C.c_ulonglong, _fp, _fp, _fp, _i, _i, _vp]),
    "ap_noise_offset_stride": (_u64, [_i, _i]),
    "ap_diffuse": (_i, [_fp, _f, _f, _fp, _u64, _u64, _fp, _i, _i, _vp]),
    "ap_ddpm_step": (_i, [_fp, _fp, _f, _f, _f, _fp, _u64, _u64, _i, _i, _vp]),
    "ap_sde_step": (_i, [_fp, _fp, C.POINTER(SdeCoef), _fp, _u64, _u64, _i, _i, _vp]),
    "ap_predict_x0": (_i, [_fp, _fp, _f, _f, _fp, _i, _i, _vp]),
    "ap_smooth_inputs": (_i, [_fp, _f, _f, _fp, _u64, _u64, _fp, _i, _i, _vp]),
    "ap_randn": (_i, [_fp, _u64, _u64, _u64, _vp]),
    "ap_diffwave_smooth_denoise": (_i, [_vp, _fp, _f, _f, _fp, _u64, _u64, _vp, _f, _f, _f, _fp, _i, _i, _vp]),
    "ap_u64_add": (_i, [_vp, _u64, _vp]),
    "ap_diffwave_purify_ddpm": (_i, [_vp, _fp, _fp, _i, _vp, _fp, _u64, _u64, _i, _i, _vp]),
    "ap_mel_create": (_i, [_PP, C.POINTER(MelCfg), _i]),
    "ap_mel_destroy": (None, [_vp]),
    "ap_mel_frames": (_i, [_vp, _i]),
    "ap_mel_db": (_i, [_vp, _fp, _fp, _i, _i, _vp]),
    "ap_mel_vjp": (_i, [_vp, _fp, _fp, _fp, _i, _i, _vp]),
    "ap_classifier_create": (_i, [_PP, C.POINTER(ClassifierCfg), _PP, _i, _i]),
    "ap_classifier_destroy": (None, [_vp]),
    "ap_classifier_forward": (_i, [_vp, _fp, _fp, _i, _i, _vp]),
    "ap_classifier_vjp": (_i, [_vp, _fp, _fp, _fp, _i, _i, _vp]),
    "ap_classifier_set_mode": (_i, [_vp, _i]),
    "ap_classifier_get_mode": (_i, [_vp]),
    "ap_unet_create": (_i, [_PP, C.POINTER(UNetCfg), _vp, _i, _PP, _i, _i]),
    "ap_unet_destroy": (None, [_vp]),
    "ap_unet_set_mode": (_i, [_vp, _i]),
    "ap_unet_eps": (_i, [_vp, _fp, _f, _fp, _i, _vp]),
    "ap_unet_eps_vjp": (_i, [_vp, _fp, _f, _fp, _fp, _fp, _i, _vp]),
    "ap_vote_counts": (_i, [_fp, _i, _i, _vp, _i, _vp]),
    "ap_argmax": (_i, [_fp, _i, _i, _vp, _vp]),
    "ap_nes_noise_blocks": (_u64, [_i, _i, _i]),
    "ap_nes_perturb": (_i, [_fp, _f, _fp, _u64, _u64, _i, _fp, _i, _i, _i, _vp]),
    "ap_nes_gradient": (_i, [_fp, _fp, _u64, _u64, _i, _f, _i, _fp, _i, _i, _i, _vp]),
    "ap_query_loss": (_i, [_fp, _vp, _i, _i, _i, _i, _f, _i, _fp, _vp, _vp]),
    "ap_query_loss_vjp": (_i, [_fp, _vp, _fp, _i, _i, _i, _i, _f, _i, _fp, _vp]),
    "ap_selftest_umma": (_i, [_vp, _vp, _fp, _i, _vp]),
    "ap_diffwave_debug_layer": (_i, [_vp, _fp, _f, _i, _fp, _fp, _i, _i, _vp]),
    "ap_diffwave_profile": (_i, [_vp, _i]),
    "ap_diffwave_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "ap_diffwave_debug_counters": (_i, [_vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AudioPureError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `python {os.path.join(_PKG_DIR, 'build.py')}`). audiopure_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ap_last_error()
        raise AudioPureError(f"{what or 'audiopure call'} failed (code {rc}): {msg.decode() if msg else '?'}")


_STREAM_IDS = {"diffwave": 0x4449464657415645, "certify": 0x4345525449465921, "nes": 0x4E45535155455259}
_instances = 0


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def philox_key(consumer: str, seed) -> int:
    """Philox key of one noise consumer ('diffwave', 'certify', 'nes').

    The reference draws the defender's diffusion noise, the smoothing noise and an attacker's NES probes from independent
    RNG streams.  Here every consumer owns a counter-based Philox4x32-10 stream; its 64-bit key is
    ``splitmix64(seed) ^ consumer_id``, so the same user seed never makes two KINDS of consumer walk the same blocks.
    ``seed=None`` (the default of the host classes) derives a fresh key per object:
    ``splitmix64(torch.initial_seed() + 2^32 * rank + instance_index)`` -- reproducible under ``torch.manual_seed``, distinct
    across the objects of one process and across the ranks of a torch.distributed job.  Pass an explicit seed (bench.py uses
    ``2024 + rank``) to pin a stream; certification shards ONE stream across ranks by offset, not by key (certify.py)."""
    global _instances
    if seed is None:
        import torch
        rank = 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank()
        seed = (int(torch.initial_seed()) + (rank << 32) + _instances) & 0xFFFFFFFFFFFFFFFF
        _instances += 1
    return _splitmix64(int(seed) & 0xFFFFFFFFFFFFFFFF) ^ _STREAM_IDS[consumer]


def check_device(x, device_index: int, what: str) -> None:
    """The C side launches on the handle's device with the caller's stream: a tensor on another GPU is an error here, not an
    'invalid resource handle' from the driver later."""
    if x.is_cuda and x.device.index != device_index:
        raise AudioPureError(f"{what}: input is on cuda:{x.device.index} but the handle was created on cuda:{device_index}")


def launch_count() -> int:
    return int(load().ap_launch_count())


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr_array(arrays) -> "C.Array":
    """Array of host pointers to contiguous float32 numpy arrays (which the caller keeps alive)."""
    arr = (C.c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data
    return arr
