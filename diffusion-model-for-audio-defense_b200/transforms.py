"""Log-mel transform: callable ``(B,1,L) -> (B,1,n_mels,frames)`` replacing the
``Compose([MelSpectrogram(...), AmplitudeToDB(stype='power')])`` the reference drivers build
(certified_robustness_eval.py:85-87; kws_adaptive_attack_eval.py:74-76)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

__all__ = ["MelSpectrogramDB", "sc09_transform", "kws_transform"]


class _MelVJP(torch.autograd.Function):
    """log-mel with a backward pass through ``ap_mel_vjp`` (torchaudio's transform is an autograd module in the reference)."""

    @staticmethod
    def forward(ctx, wav, mod):
        x = wav.detach().to(torch.float32).contiguous()
        ctx.mod = mod
        ctx.save_for_backward(x)
        return mod._db(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ctx.mod.vjp(x, g).reshape(x.shape), None


class MelSpectrogramDB(torch.nn.Module):
    def __init__(self, sample_rate=16000, n_fft=400, hop_length=None, n_mels=128, norm=None, pad_mode="reflect",
                 mel_scale="htk", device=None):
        super().__init__()
        if hop_length is None:
            hop_length = n_fft // 2          # torchaudio default: win_length // 2
        assert pad_mode in ("reflect", "constant") and mel_scale in ("htk", "slaney") and norm in (None, "slaney")
        self.n_mels, self.hop_length = n_mels, hop_length
        self._lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        dev = device if isinstance(device, int) else (torch.device(device).index or 0)
        cfg = _lib.MelCfg(sample_rate, n_fft, hop_length, n_mels, int(norm == "slaney"), int(mel_scale == "slaney"),
                          int(pad_mode == "reflect"))
        self._handle = C.c_void_p()
        self.device_index = dev
        _lib.check(self._lib.ap_mel_create(C.byref(self._handle), C.byref(cfg), dev), "ap_mel_create")

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        if not wav.is_cuda:
            raise _lib.AudioPureError("MelSpectrogramDB: input must be a CUDA tensor (there is no CPU path)")
        _lib.check_device(wav, self.device_index, "MelSpectrogramDB")
        if wav.requires_grad and torch.is_grad_enabled():
            return _MelVJP.apply(wav, self)
        return self._db(wav.detach().to(torch.float32).contiguous())

    def _db(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-1]
        L = x.shape[-1]
        B = int(x.numel() // L)
        frames = 1 + L // self.hop_length
        out = torch.empty(*lead, self.n_mels, frames, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_mel_db(self._handle, x.data_ptr(), out.data_ptr(), B, L, _lib.stream_ptr()), "ap_mel_db")
        return out

    def vjp(self, wav: torch.Tensor, g_spec: torch.Tensor) -> torch.Tensor:
        """g_wav = (d spec / d wav)^T g_spec."""
        x = wav.detach().to(torch.float32).contiguous()
        g = g_spec.detach().to(torch.float32).contiguous()
        L = x.shape[-1]
        B = int(x.numel() // L)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_mel_vjp(self._handle, x.data_ptr(), g.data_ptr(), out.data_ptr(), B, L, _lib.stream_ptr()),
                       "ap_mel_vjp")
        return out

    def cuda(self, device=None):   # the drivers call .cuda() on transforms (acoustic_system.py:39-40)
        return self

    def __del__(self):
        try:   # may run during interpreter shutdown, when torch's Module.__setattr__ no longer works
            h = self.__dict__.pop("_handle", None)
            if h:
                self._lib.ap_mel_destroy(h)
        except Exception:
            pass


def sc09_transform(device=None) -> MelSpectrogramDB:
    """certified_robustness_eval.py:85-86"""
    return MelSpectrogramDB(sample_rate=16000, n_fft=2048, hop_length=512, n_mels=32, norm="slaney", pad_mode="constant",
                            mel_scale="slaney", device=device)


def kws_transform(device=None) -> MelSpectrogramDB:
    """kws_adaptive_attack_eval.py:74-75 (torchaudio defaults: n_fft 400, hop 200, reflect, htk, no norm)"""
    return MelSpectrogramDB(sample_rate=16000, n_mels=32, device=device)
