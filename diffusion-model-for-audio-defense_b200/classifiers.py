"""Classifier drop-ins: ``Classifier(x) -> logits / log-probs`` (acoustic_system.py:49, certified_robust.py:30).

  * ``ResNeXtClassifier``  CifarResNeXt, audio_models/ConvNets_SpeechCommands/models/resnext.py:67-142
  * ``M5Classifier``       M5, audio_models/M5/M5Net.py:4-38
  * ``KWSClassifier``      KWSModel, audio_models/RCNN_KWS/model.py:66-113
Each takes the reference module's ``state_dict()`` (tensors or numpy arrays).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["ResNeXtClassifier", "ResNetClassifier", "VGGClassifier", "WideResNetClassifier", "DenseNetClassifier", "M5Classifier", "KWSClassifier", "create_model"]


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


class _ClassifierVJP(torch.autograd.Function):
    """classifier forward with a backward pass through ``ap_classifier_vjp`` (ResNeXt)."""

    @staticmethod
    def forward(ctx, x, mod, B, in_len):
        xd = x.detach().to(torch.float32).contiguous()
        ctx.mod, ctx.B, ctx.in_len = mod, B, in_len
        ctx.save_for_backward(xd)
        return mod._run(xd, B, in_len)

    @staticmethod
    def backward(ctx, g):
        (xd,) = ctx.saved_tensors
        gx = torch.empty_like(xd)
        gl = g.detach().to(torch.float32).contiguous()
        with torch.cuda.device(xd.device):
            _lib.check(ctx.mod._lib.ap_classifier_vjp(ctx.mod._handle, xd.data_ptr(), gl.data_ptr(), gx.data_ptr(), ctx.B,
                                                      ctx.in_len, _lib.stream_ptr()), "ap_classifier_vjp")
        return gx, None, None, None


class _Classifier(torch.nn.Module):
    kind = -1
    differentiable = False      # ResNeXt has a backward pass (ap_classifier_vjp)

    def _create(self, cfg: "_lib.ClassifierCfg", weights, device):
        self._lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        dev = device if isinstance(device, int) else (torch.device(device).index or 0)
        self._weights = [_np32(w) for w in weights]
        self._handle = C.c_void_p()
        self.device_index = dev
        _lib.check(self._lib.ap_classifier_create(C.byref(self._handle), C.byref(cfg), _lib.ptr_array(self._weights),
                                                  len(self._weights), dev), "ap_classifier_create")
        self.num_classes = cfg.num_classes

    def _run(self, x: torch.Tensor, B: int, in_len: int) -> torch.Tensor:
        if not x.is_cuda:
            raise _lib.AudioPureError(f"{type(self).__name__}: input must be a CUDA tensor (there is no CPU path)")
        _lib.check_device(x, self.device_index, type(self).__name__)
        if x.requires_grad and torch.is_grad_enabled():
            if not self.differentiable:
                raise _lib.AudioPureError(f"{type(self).__name__}: inference-only (input requires grad)")
            return _ClassifierVJP.apply(x, self, B, in_len)
        x = x.detach().to(torch.float32).contiguous()
        out = torch.empty(B, self.num_classes, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(self._lib.ap_classifier_forward(self._handle, x.data_ptr(), out.data_ptr(), B, in_len,
                                                       _lib.stream_ptr()), "ap_classifier_forward")
        return out

    def cuda(self, device=None):
        return self

    def set_mode(self, mode: str):
        """'tf32' (tensor-core convolutions, default for ResNeXt, ResNet, VGG and WideResNet) or 'fp32' (FFMA everywhere)."""
        m = {"fp32": _lib.AP_MODE_FP32, "tf32": _lib.AP_MODE_TF32}[mode]
        _lib.check(self._lib.ap_classifier_set_mode(self._handle, m), "ap_classifier_set_mode")
        return self

    @property
    def mode(self) -> str:
        return "tf32" if self._lib.ap_classifier_get_mode(self._handle) == _lib.AP_MODE_TF32 else "fp32"

    def __del__(self):
        try:   # may run during interpreter shutdown, when torch's Module.__setattr__ no longer works
            h = self.__dict__.pop("_handle", None)
            if h:
                self._lib.ap_classifier_destroy(h)
        except Exception:
            pass


def _strip(sd: dict) -> dict:
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items() if not k.endswith("num_batches_tracked")}


class ResNeXtClassifier(_Classifier):
    differentiable = True

    def __init__(self, state_dict: dict, nlabels=10, cardinality=8, depth=29, base_width=64, widen_factor=4,
                 in_channels=1, device=None):
        super().__init__()
        sd = _strip(state_dict)
        bn = lambda p: [sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"]]
        w = [sd["conv_1_3x3.weight"], *bn("bn_1")]
        for s in range(3):
            for b in range((depth - 2) // 9):
                p = f"stage_{s + 1}.stage_{s + 1}_bottleneck_{b}"
                w += [sd[p + ".conv_reduce.weight"], *bn(p + ".bn_reduce"), sd[p + ".conv_conv.weight"], *bn(p + ".bn"),
                      sd[p + ".conv_expand.weight"], *bn(p + ".bn_expand")]
                if p + ".shortcut.shortcut_conv.weight" in sd:
                    w += [sd[p + ".shortcut.shortcut_conv.weight"], *bn(p + ".shortcut.shortcut_bn")]
        w += [sd["classifier.weight"], sd["classifier.bias"]]
        cfg = _lib.ClassifierCfg(_lib.AP_CLS_RESNEXT, nlabels, cardinality, depth, base_width, widen_factor, in_channels,
                                 0, 0, 0, 0, 0)
        self._create(cfg, w, device)

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        assert spec.ndim == 4 and tuple(spec.shape[1:]) == (1, 32, 32), f"expected (B,1,32,32), got {tuple(spec.shape)}"
        return self._run(spec, spec.shape[0], 32)


class ResNetClassifier(_Classifier):
    """torchvision-style ResNet-18/34/50/101/152 with ``in_channels`` (models/resnet.py:103-220) on (B,1,32,32) input."""
    differentiable = True
    LAYERS = {18: (False, (2, 2, 2, 2)), 34: (False, (3, 4, 6, 3)), 50: (True, (3, 4, 6, 3)), 101: (True, (3, 4, 23, 3)),
              152: (True, (3, 8, 36, 3))}

    def __init__(self, state_dict: dict, depth: int = 34, num_classes=10, in_channels=1, device=None):
        super().__init__()
        sd = _strip(state_dict)
        bn = lambda p: [sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"]]
        bottleneck, counts = self.LAYERS[depth]
        w = [sd["conv1.weight"], *bn("bn1")]
        for l, n in enumerate(counts):
            for b in range(n):
                p = f"layer{l + 1}.{b}"
                for j in range(3 if bottleneck else 2):
                    w += [sd[f"{p}.conv{j + 1}.weight"], *bn(f"{p}.bn{j + 1}")]
                if f"{p}.downsample.0.weight" in sd:
                    w += [sd[f"{p}.downsample.0.weight"], *bn(f"{p}.downsample.1")]
        w += [sd["fc.weight"], sd["fc.bias"]]
        cfg = _lib.ClassifierCfg(_lib.AP_CLS_RESNET, num_classes, 0, depth, 0, 0, in_channels, 0, 0, 0, 0, 0)
        self._create(cfg, w, device)

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        assert spec.ndim == 4 and tuple(spec.shape[1:]) == (1, 32, 32), f"expected (B,1,32,32), got {tuple(spec.shape)}"
        return self._run(spec, spec.shape[0], 32)


class VGGClassifier(_Classifier):
    """VGG-11/13/16/19 with batch norm and ``in_channels`` (models/vgg.py:32-95; ``vgg19_bn`` is the SC09 factory's fallback
    and one of ``--classifier_model``'s choices, adaptive_attack_eval.py:21) on (B,1,32,32) input."""
    differentiable = True
    CFG = {11: [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
           13: [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
           16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
           19: [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]}

    def __init__(self, state_dict: dict, depth: int = 19, num_classes=10, in_channels=1, device=None):
        super().__init__()
        sd = _strip(state_dict)
        w, i = [], 0
        for v in self.CFG[depth]:                       # make_layers (vgg.py:69-82): conv, BatchNorm, ReLU | MaxPool
            if v == "M":
                i += 1
                continue
            c, b = f"features.{i}", f"features.{i + 1}"
            if b + ".running_mean" not in sd:
                raise NotImplementedError("VGGClassifier: only the batch-norm variants (vgg*_bn) have a kernel path")
            w += [sd[c + ".weight"], sd[c + ".bias"], sd[b + ".weight"], sd[b + ".bias"], sd[b + ".running_mean"],
                  sd[b + ".running_var"]]
            i += 3
        for j in (0, 3, 6):
            w += [sd[f"classifier.{j}.weight"], sd[f"classifier.{j}.bias"]]
        cfg = _lib.ClassifierCfg(_lib.AP_CLS_VGG, num_classes, 0, depth, 0, 0, in_channels, 0, 0, 0, 0, 0)
        self._create(cfg, w, device)

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        assert spec.ndim == 4 and tuple(spec.shape[1:]) == (1, 32, 32), f"expected (B,1,32,32), got {tuple(spec.shape)}"
        return self._run(spec, spec.shape[0], 32)


class WideResNetClassifier(_Classifier):
    """WideResNet-depth-widen_factor with ``in_channels`` (models/wideresnet.py:15-92; ``wideresnet28_10`` is one of
    ``--classifier_model``'s choices, adaptive_attack_eval.py:21) on (B,1,32,32) input.  Dropout (``wideresnet28_10D``) is the
    identity in eval mode."""
    differentiable = True

    def __init__(self, state_dict: dict, depth: int = 28, widen_factor: int = 10, num_classes=10, in_channels=1, device=None):
        super().__init__()
        sd = _strip(state_dict)
        bn = lambda p: [sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"]]
        w = [sd["conv1.weight"]]
        for s in (1, 2, 3):
            for b in range((depth - 4) // 6):
                p = f"block{s}.layer.{b}"
                w += [*bn(p + ".bn1"), sd[p + ".conv1.weight"], *bn(p + ".bn2"), sd[p + ".conv2.weight"]]
                if p + ".convShortcut.weight" in sd:
                    w.append(sd[p + ".convShortcut.weight"])
        w += [*bn("bn1"), sd["fc.weight"], sd["fc.bias"]]
        cfg = _lib.ClassifierCfg(_lib.AP_CLS_WRN, num_classes, 0, depth, 0, widen_factor, in_channels, 0, 0, 0, 0, 0)
        self._create(cfg, w, device)

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        assert spec.ndim == 4 and tuple(spec.shape[1:]) == (1, 32, 32), f"expected (B,1,32,32), got {tuple(spec.shape)}"
        return self._run(spec, spec.shape[0], 32)


class DenseNetClassifier(_Classifier):
    """DenseNet-BC-depth-growthRate with ``in_channels`` (models/densenet.py:15-147; ``densenet_bc_100_12`` is one of
    ``--classifier_model``'s choices, adaptive_attack_eval.py:21) on (B,1,32,32) input.  Bottleneck blocks only (the factory
    builds nothing else, models/__init__.py:37-42)."""
    differentiable = True

    def __init__(self, state_dict: dict, depth: int = 100, growth_rate: int = 12, compression_rate: int = 2, num_classes=10,
                 in_channels=1, device=None):
        super().__init__()
        sd = _strip(state_dict)
        bn = lambda p: [sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"]]
        n = (depth - 4) // 6
        w = [sd["conv1.weight"]]
        for s in (1, 2, 3):
            for l in range(n):
                p = f"dense{s}.{l}"
                w += [*bn(p + ".bn1"), sd[p + ".conv1.weight"], *bn(p + ".bn2"), sd[p + ".conv2.weight"]]
            if s < 3:
                w += [*bn(f"trans{s}.bn1"), sd[f"trans{s}.conv1.weight"]]
        w += [*bn("bn"), sd["fc.weight"], sd["fc.bias"]]
        cfg = _lib.ClassifierCfg(_lib.AP_CLS_DENSENET, num_classes, 0, depth, growth_rate, compression_rate, in_channels, 0, 0, 0, 0, 0)
        self._create(cfg, w, device)

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        assert spec.ndim == 4 and tuple(spec.shape[1:]) == (1, 32, 32), f"expected (B,1,32,32), got {tuple(spec.shape)}"
        return self._run(spec, spec.shape[0], 32)


class M5Classifier(_Classifier):
    differentiable = True
    def __init__(self, state_dict: dict, n_input=1, first_kernel_size=160, n_output=10, stride=16, n_channel=32,
                 device=None):
        super().__init__()
        assert n_input == 1
        sd = _strip(state_dict)
        w = []
        for i in range(1, 5):
            w += [sd[f"conv{i}.weight"], sd[f"conv{i}.bias"], sd[f"bn{i}.weight"], sd[f"bn{i}.bias"],
                  sd[f"bn{i}.running_mean"], sd[f"bn{i}.running_var"]]
        w += [sd["fc1.weight"], sd["fc1.bias"]]
        cfg = _lib.ClassifierCfg(_lib.AP_CLS_M5, n_output, 0, 0, 0, 0, 1, first_kernel_size, stride, n_channel, 0, 0)
        self._create(cfg, w, device)

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        assert wav.ndim == 3 and wav.shape[1] == 1, f"expected (B,1,L), got {tuple(wav.shape)}"
        return self._run(wav, wav.shape[0], wav.shape[2])


class KWSClassifier(_Classifier):
    differentiable = True
    def __init__(self, state_dict: dict, in_size=32, hidden_size=64, num_classes=4, device=None):
        super().__init__()
        sd = _strip(state_dict)
        w = [sd["CRNN_model.sepconv.0.weight"], sd["CRNN_model.sepconv.0.bias"], sd["CRNN_model.sepconv.1.weight"],
             sd["CRNN_model.sepconv.1.bias"]]
        for layer in range(2):
            for suffix in ("", "_reverse"):
                w += [sd[f"CRNN_model.gru.{n}_l{layer}{suffix}"] for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        w += [sd["attn_layer.Wx_b.weight"], sd["attn_layer.Wx_b.bias"], sd["attn_layer.Vt.weight"], sd["apply_attn.U.weight"]]
        cfg = _lib.ClassifierCfg(_lib.AP_CLS_KWS, num_classes, 0, 0, 0, 0, 1, 0, 0, 0, in_size, hidden_size)
        self._create(cfg, w, device)
        self.in_size = in_size

    def forward(self, spec: torch.Tensor) -> torch.Tensor:
        x = spec.squeeze(1) if spec.ndim == 4 else spec          # model.py:94
        assert x.ndim == 3 and x.shape[1] == self.in_size, f"expected (B,1,{self.in_size},W), got {tuple(spec.shape)}"
        return self._run(x, x.shape[0], x.shape[2])


def create_model(path: str, device=None):
    """audio_models/ConvNets_SpeechCommands/create_model.py:8-16: unpickle a whole (DataParallel) module and rebuild
    it on the B200 kernels from its state dict.  Needs the reference's model classes importable for the unpickle."""
    model = torch.load(path, map_location="cpu", weights_only=False)
    if hasattr(model, "module"):
        model = model.module
    name = type(model).__name__
    sd = model.state_dict()
    if name == "CifarResNeXt":
        return ResNeXtClassifier(sd, nlabels=model.nlabels, cardinality=model.cardinality, depth=model.depth,
                                 base_width=model.base_width, widen_factor=model.widen_factor,
                                 in_channels=model.conv_1_3x3.in_channels, device=device)
    if name == "ResNet":
        bottleneck = type(model.layer1[0]).__name__ == "Bottleneck"
        counts = tuple(len(getattr(model, f"layer{i}")) for i in range(1, 5))
        depth = {v: k for k, v in ResNetClassifier.LAYERS.items()}[(bottleneck, counts)]
        return ResNetClassifier(sd, depth=depth, num_classes=model.fc.out_features, in_channels=model.conv1.in_channels,
                                device=device)
    if name == "VGG":
        convs = sum(1 for m in model.features if type(m).__name__ == "Conv2d")
        return VGGClassifier(sd, depth=convs + 3, num_classes=model.classifier[6].out_features,
                             in_channels=model.features[0].in_channels, device=device)
    if name == "WideResNet":
        n = len(model.block1.layer)
        return WideResNetClassifier(sd, depth=6 * n + 4, widen_factor=model.nChannels // 64, num_classes=model.fc.out_features,
                                    in_channels=model.conv1.in_channels, device=device)
    if name == "DenseNet":
        n = len(model.dense1)
        growth = model.growthRate
        comp = round(model.trans1.conv1.in_channels / model.trans1.conv1.out_channels)
        return DenseNetClassifier(sd, depth=6 * n + 4, growth_rate=growth, compression_rate=comp,
                                  num_classes=model.fc.out_features, in_channels=model.conv1.in_channels, device=device)
    if name == "M5":
        return M5Classifier(sd, first_kernel_size=model.conv1.kernel_size[0], n_output=model.fc1.out_features,
                            stride=model.conv1.stride[0], n_channel=model.conv1.out_channels, device=device)
    if name == "KWSModel":
        return KWSClassifier(sd, in_size=model.in_size, hidden_size=model.hidden_size, num_classes=model.num_classes,
                             device=device)
    raise NotImplementedError(f"create_model: classifier {name} has no B200 kernel path (SURVEY.md section 2, row 8a)")
