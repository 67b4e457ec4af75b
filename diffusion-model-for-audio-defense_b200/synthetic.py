"""Deterministic synthetic weights and waveforms (SURVEY.md §8c "Weights", §8d "Synthetic inputs").

The DiffWave and ResNeXt checkpoints of the reference are not in its tree
(README.md:3 points at Google Drive), so throughput and parity are measured
on seeded random-init weights of the named architectures.  Every tensor is
drawn from its own PCG64 stream keyed by (seed, crc32(name)), so the result
does not depend on creation order, torch's RNG, or the torch version.

State-dict names and shapes are exactly the reference's:
  * WaveNet_Speech_Commands  -- diffusion_models/DiffWave_Unconditional/WaveNet.py:138-172
    (legacy weight-norm keys ``conv.weight_g`` / ``conv.weight_v``)
  * CifarResNeXt             -- audio_models/ConvNets_SpeechCommands/models/resnext.py:67-142
  * M5                       -- audio_models/M5/M5Net.py:4-38
  * KWSModel                 -- audio_models/RCNN_KWS/model.py:66-113

``final_conv.2`` is zero-initialised in the reference (WaveNet.py:43-44), which
makes eps == 0 and every parity check vacuous; it is re-randomised here.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np

DEFAULT_WAVENET_CONFIG = dict(  # configs/config.json:7-17
    in_channels=1, res_channels=256, skip_channels=256, out_channels=1,
    num_res_layers=36, dilation_cycle=12,
    diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
    diffusion_step_embed_dim_out=512)
DEFAULT_DIFFUSION_CONFIG = dict(T=200, beta_0=0.0001, beta_T=0.02)  # configs/config.json:2-6


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def _normal(seed, name, shape, std):
    return (_rng(seed, name).standard_normal(shape) * std).astype(np.float32)


def _uniform(seed, name, shape, lo, hi):
    return _rng(seed, name).uniform(lo, hi, shape).astype(np.float32)


def _wn_conv(sd, seed, prefix, cout, cin, k):
    """Weight-normed Conv1d: v ~ kaiming normal, g = ||v|| * U(0.8, 1.2), bias ~ U(+-1/sqrt(fan_in))."""
    fan_in = cin * k
    v = _normal(seed, prefix + ".weight_v", (cout, cin, k), np.sqrt(2.0 / fan_in))
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2))).astype(np.float32)
    g = norm * _uniform(seed, prefix + ".weight_g", (cout,), 0.8, 1.2)
    b = _uniform(seed, prefix + ".bias", (cout,), -1, 1) / np.float32(np.sqrt(fan_in))
    sd[prefix + ".bias"] = b.astype(np.float32)
    sd[prefix + ".weight_g"] = g.reshape(cout, 1, 1).astype(np.float32)
    sd[prefix + ".weight_v"] = v


def _linear(sd, seed, prefix, cout, cin, bias=True, gain=1.0):
    bound = gain / np.sqrt(cin)
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (cout, cin), -bound, bound)
    if bias:
        sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (cout,), -bound, bound)


def wavenet_state_dict(seed: int = 0, config: dict | None = None) -> "OrderedDict[str, np.ndarray]":
    """Seeded state dict with the reference's 408 keys (for the default config)."""
    c = dict(DEFAULT_WAVENET_CONFIG)
    if config:
        c.update(config)
    C, S = c["res_channels"], c["skip_channels"]
    e_in, e_mid, e_out = (c["diffusion_step_embed_dim_in"], c["diffusion_step_embed_dim_mid"],
                          c["diffusion_step_embed_dim_out"])
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    _wn_conv(sd, seed, "init_conv.0.conv", C, c["in_channels"], 1)
    _linear(sd, seed, "residual_layer.fc_t1", e_mid, e_in)
    _linear(sd, seed, "residual_layer.fc_t2", e_out, e_mid)
    for n in range(c["num_res_layers"]):
        p = f"residual_layer.residual_blocks.{n}"
        _linear(sd, seed, p + ".fc_t", C, e_out)
        _wn_conv(sd, seed, p + ".dilated_conv_layer.conv", 2 * C, C, 3)
        _wn_conv(sd, seed, p + ".res_conv", C, C, 1)
        _wn_conv(sd, seed, p + ".skip_conv", S, C, 1)
    _wn_conv(sd, seed, "final_conv.0.conv", S, S, 1)
    sd["final_conv.2.conv.weight"] = _normal(seed, "final_conv.2.conv.weight",
                                             (c["out_channels"], S, 1), np.sqrt(2.0 / S))
    sd["final_conv.2.conv.bias"] = _uniform(seed, "final_conv.2.conv.bias", (c["out_channels"],), -0.05, 0.05)
    return sd


def _bn(sd, seed, prefix, ch):
    sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (ch,), 0.5, 1.5)
    sd[prefix + ".bias"] = _normal(seed, prefix + ".bias", (ch,), 0.1)
    sd[prefix + ".running_mean"] = _normal(seed, prefix + ".running_mean", (ch,), 0.1)
    sd[prefix + ".running_var"] = _uniform(seed, prefix + ".running_var", (ch,), 0.5, 1.5)
    sd[prefix + ".num_batches_tracked"] = np.array(1, dtype=np.int64)


def _conv2d(sd, seed, name, cout, cin_per_group, kh, kw):
    fan_out = cout * kh * kw
    sd[name + ".weight"] = _normal(seed, name + ".weight", (cout, cin_per_group, kh, kw), np.sqrt(2.0 / fan_out))


def resnext_state_dict(seed: int = 0, nlabels: int = 10, cardinality: int = 8, depth: int = 29,
                       base_width: int = 64, widen_factor: int = 4, in_channels: int = 1):
    """CifarResNeXt state dict (resnext.py:67-142); BN gets non-trivial running stats."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    block_depth = (depth - 2) // 9
    stages = [64, 64 * widen_factor, 128 * widen_factor, 256 * widen_factor]
    _conv2d(sd, seed, "conv_1_3x3", 64, in_channels, 3, 3)
    _bn(sd, seed, "bn_1", 64)
    for s in range(3):
        cin, cout = stages[s], stages[s + 1]
        for b in range(block_depth):
            p = f"stage_{s + 1}.stage_{s + 1}_bottleneck_{b}"
            bin_ = cin if b == 0 else cout
            width_ratio = cout / (widen_factor * 64.0)
            D = cardinality * int(base_width * width_ratio)
            _conv2d(sd, seed, p + ".conv_reduce", D, bin_, 1, 1)
            _bn(sd, seed, p + ".bn_reduce", D)
            _conv2d(sd, seed, p + ".conv_conv", D, D // cardinality, 3, 3)
            _bn(sd, seed, p + ".bn", D)
            _conv2d(sd, seed, p + ".conv_expand", cout, D, 1, 1)
            _bn(sd, seed, p + ".bn_expand", cout)
            if bin_ != cout:
                _conv2d(sd, seed, p + ".shortcut.shortcut_conv", cout, bin_, 1, 1)
                _bn(sd, seed, p + ".shortcut.shortcut_bn", cout)
    _linear(sd, seed, "classifier", nlabels, stages[3], gain=4.0)
    return sd


RESNET_LAYERS = {18: (False, (2, 2, 2, 2)), 34: (False, (3, 4, 6, 3)), 50: (True, (3, 4, 6, 3)), 101: (True, (3, 4, 23, 3)),
                 152: (True, (3, 8, 36, 3))}


def resnet_state_dict(depth: int = 34, seed: int = 0, num_classes: int = 10, in_channels: int = 1):
    """torchvision-style ResNet state dict (models/resnet.py:103-160), BN with non-trivial running stats."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    bottleneck, counts = RESNET_LAYERS[depth]
    exp = 4 if bottleneck else 1
    _conv2d(sd, seed, "conv1", 64, in_channels, 7, 7)
    _bn(sd, seed, "bn1", 64)
    inpl = 64
    for l, n in enumerate(counts):
        planes = 64 << l
        for b in range(n):
            p = f"layer{l + 1}.{b}"
            stride = 2 if (b == 0 and l > 0) else 1
            if bottleneck:
                _conv2d(sd, seed, p + ".conv1", planes, inpl, 1, 1); _bn(sd, seed, p + ".bn1", planes)
                _conv2d(sd, seed, p + ".conv2", planes, planes, 3, 3); _bn(sd, seed, p + ".bn2", planes)
                _conv2d(sd, seed, p + ".conv3", planes * 4, planes, 1, 1); _bn(sd, seed, p + ".bn3", planes * 4)
            else:
                _conv2d(sd, seed, p + ".conv1", planes, inpl, 3, 3); _bn(sd, seed, p + ".bn1", planes)
                _conv2d(sd, seed, p + ".conv2", planes, planes, 3, 3); _bn(sd, seed, p + ".bn2", planes)
            if b == 0 and (stride != 1 or inpl != planes * exp):
                _conv2d(sd, seed, p + ".downsample.0", planes * exp, inpl, 1, 1); _bn(sd, seed, p + ".downsample.1", planes * exp)
            inpl = planes * exp
    _linear(sd, seed, "fc", num_classes, 512 * exp, gain=4.0)
    return sd


VGG_CFG = {11: [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
           13: [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
           16: [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
           19: [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]}


def vgg_state_dict(depth: int = 19, seed: int = 0, num_classes: int = 10, in_channels: int = 1):
    """vgg*_bn state dict (models/vgg.py:32-95): conv weight + bias, BN with non-trivial running stats, three Linear layers."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    i, cin = 0, in_channels
    for v in VGG_CFG[depth]:
        if v == "M":
            i += 1
            continue
        _conv2d(sd, seed, f"features.{i}", v, cin, 3, 3)
        sd[f"features.{i}.bias"] = _normal(seed, f"features.{i}.bias", (v,), 0.1)
        _bn(sd, seed, f"features.{i + 1}", v)
        i, cin = i + 3, v
    _linear(sd, seed, "classifier.0", 4096, 512, gain=2.0)
    _linear(sd, seed, "classifier.3", 4096, 4096, gain=2.0)
    _linear(sd, seed, "classifier.6", num_classes, 4096, gain=4.0)
    return sd


def wideresnet_state_dict(depth: int = 28, widen_factor: int = 10, seed: int = 0, num_classes: int = 10, in_channels: int = 1):
    """WideResNet state dict (models/wideresnet.py:53-80), BN with non-trivial running stats."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    ch = [16, 16 * widen_factor, 32 * widen_factor, 64 * widen_factor]
    _conv2d(sd, seed, "conv1", ch[0], in_channels, 3, 3)
    for s in range(3):
        for b in range((depth - 4) // 6):
            p = f"block{s + 1}.layer.{b}"
            cin = ch[s] if b == 0 else ch[s + 1]
            _bn(sd, seed, p + ".bn1", cin)
            _conv2d(sd, seed, p + ".conv1", ch[s + 1], cin, 3, 3)
            _bn(sd, seed, p + ".bn2", ch[s + 1])
            _conv2d(sd, seed, p + ".conv2", ch[s + 1], ch[s + 1], 3, 3)
            if cin != ch[s + 1]:
                _conv2d(sd, seed, p + ".convShortcut", ch[s + 1], cin, 1, 1)
    _bn(sd, seed, "bn1", ch[3])
    _linear(sd, seed, "fc", num_classes, ch[3], gain=4.0)
    return sd


def densenet_state_dict(depth: int = 100, growth_rate: int = 12, compression_rate: int = 2, seed: int = 0, num_classes: int = 10,
                        in_channels: int = 1):
    """DenseNet-BC state dict (models/densenet.py:74-123, Bottleneck blocks), BN with non-trivial running stats.  Convolutions
    are drawn with std sqrt(1 / fan_in): the reference's fan-out rule amplifies every 48 -> 12 convolution ~3x without trained
    BatchNorm statistics, which over 48 concatenated layers overflows any useful dynamic range (logits ~ 4e7)."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    n, g = (depth - 4) // 6, growth_rate

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = _normal(seed, name + ".weight", (cout, cin, k, k), np.sqrt(1.0 / (cin * k * k)))

    C = 2 * g
    conv("conv1", C, in_channels, 3)
    for s in (1, 2, 3):
        for l in range(n):
            p = f"dense{s}.{l}"
            _bn(sd, seed, p + ".bn1", C)
            conv(p + ".conv1", 4 * g, C, 1)
            _bn(sd, seed, p + ".bn2", 4 * g)
            conv(p + ".conv2", g, 4 * g, 3)
            C += g
        if s < 3:
            _bn(sd, seed, f"trans{s}.bn1", C)
            conv(f"trans{s}.conv1", C // compression_rate, C, 1)
            C = C // compression_rate
    _bn(sd, seed, "bn", C)
    _linear(sd, seed, "fc", num_classes, C, gain=4.0)
    return sd


def m5_state_dict(seed: int = 0, n_input=1, first_kernel_size=160, n_output=10, n_channel=32):
    """M5 state dict (M5Net.py:4-20)."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    shapes = [("conv1", n_channel, n_input, first_kernel_size), ("conv2", n_channel, n_channel, 3),
              ("conv3", 2 * n_channel, n_channel, 3), ("conv4", 2 * n_channel, 2 * n_channel, 3)]
    for i, (name, co, ci, k) in enumerate(shapes):
        bound = 1.0 / np.sqrt(ci * k)
        sd[name + ".weight"] = _uniform(seed, name + ".weight", (co, ci, k), -bound, bound) * np.float32(2.0)
        sd[name + ".bias"] = _uniform(seed, name + ".bias", (co,), -bound, bound)
        _bn(sd, seed, f"bn{i + 1}", co)
    _linear(sd, seed, "fc1", n_output, 2 * n_channel, gain=4.0)
    return sd


def kws_state_dict(seed: int = 0, in_size=32, hidden_size=64, kernel_size=(20, 5), gru_num_layers=2, num_classes=4):
    """KWSModel state dict (RCNN_KWS/model.py:5-113): 24 tensors."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    p = "CRNN_model.sepconv"
    b0 = 1.0 / np.sqrt(kernel_size[1])
    sd[p + ".0.weight"] = _uniform(seed, p + ".0.weight", (in_size, 1, kernel_size[1]), -b0, b0)
    sd[p + ".0.bias"] = _uniform(seed, p + ".0.bias", (in_size,), -b0, b0)
    groups = int(in_size / kernel_size[0])
    cin_g = in_size // groups
    b1 = 1.0 / np.sqrt(cin_g)
    sd[p + ".1.weight"] = _uniform(seed, p + ".1.weight", (hidden_size, cin_g, 1), -b1, b1)
    sd[p + ".1.bias"] = _uniform(seed, p + ".1.bias", (hidden_size,), -b1, b1)
    H = hidden_size
    bg = 1.0 / np.sqrt(H)
    for layer in range(gru_num_layers):
        in_l = H if layer == 0 else 2 * H
        for suffix in ("", "_reverse"):
            g = "CRNN_model.gru."
            sd[f"{g}weight_ih_l{layer}{suffix}"] = _uniform(seed, f"{g}weight_ih_l{layer}{suffix}", (3 * H, in_l), -bg, bg)
            sd[f"{g}weight_hh_l{layer}{suffix}"] = _uniform(seed, f"{g}weight_hh_l{layer}{suffix}", (3 * H, H), -bg, bg)
            sd[f"{g}bias_ih_l{layer}{suffix}"] = _uniform(seed, f"{g}bias_ih_l{layer}{suffix}", (3 * H,), -bg, bg)
            sd[f"{g}bias_hh_l{layer}{suffix}"] = _uniform(seed, f"{g}bias_hh_l{layer}{suffix}", (3 * H,), -bg, bg)
    _linear(sd, seed, "attn_layer.Wx_b", 2 * H, 2 * H)
    _linear(sd, seed, "attn_layer.Vt", 1, 2 * H, bias=False)
    _linear(sd, seed, "apply_attn.U", num_classes, 2 * H, bias=False, gain=4.0)
    return sd


def synthetic_waveforms(batch: int, length: int = 16000, seed: int = 1234, sample_rate: int = 16000) -> np.ndarray:
    """x = clamp(0.1*randn + 0.3*sin(2*pi*f*n/sr), -1, 1), f ~ U(100, 4000) per sample."""
    rng = np.random.Generator(np.random.PCG64([seed, 0x5C09]))
    f = rng.uniform(100.0, 4000.0, size=(batch, 1, 1))
    n = np.arange(length, dtype=np.float64).reshape(1, 1, length)
    x = 0.1 * rng.standard_normal((batch, 1, length)) + 0.3 * np.sin(2 * np.pi * f * n / sample_rate)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def host_noise(shape, seed: int = 2024, index: int = 0) -> np.ndarray:
    """Host-generated standard-normal tensors handed identically to the oracle and the CUDA path."""
    rng = np.random.Generator(np.random.PCG64([seed, index]))
    return rng.standard_normal(shape).astype(np.float32)


# ------------------------------------------------------------------------------------------------ spectrogram UNet (Diffusion-Spec)
DEFAULT_UNET_CONFIG = dict(  # improved_diffusion/script_util.py:12-33,85-127 (model_and_diffusion_defaults + create_model)
    image_size=32, in_channels=1, model_channels=128, out_channels=1, num_res_blocks=3, attention_resolutions=(2, 4),
    channel_mult=(1, 2, 2, 2), num_heads=4, num_heads_upsample=4, use_scale_shift_norm=True)


def unet_structure(config: dict | None = None):
    """The module walk of UNetModel.__init__ (improved_diffusion/unet.py:301-443) as a flat list of
    ``(state-dict prefix, kind, cin, cout)`` with kind in {'conv_in', 'res', 'attn', 'down', 'up', 'out'}; 'res' of the
    output path has cin = channels of the concatenated input.  Shared by the synthetic weights, the oracle and the
    host module, so the three agree on names and order by construction."""
    c = dict(DEFAULT_UNET_CONFIG)
    if config:
        c.update(config)
    mc, nrb, mult, att = c["model_channels"], c["num_res_blocks"], c["channel_mult"], tuple(c["attention_resolutions"])
    ops = [("input_blocks.0.0", "conv_in", c["in_channels"], mc)]
    chans, ch, ds, idx = [mc], mc, 1, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            ops.append((f"input_blocks.{idx}.0", "res", ch, m * mc))
            ch = m * mc
            if ds in att:
                ops.append((f"input_blocks.{idx}.1", "attn", ch, ch))
            ops.append((f"input_blocks.{idx}", "push", ch, ch))
            chans.append(ch)
            idx += 1
        if level != len(mult) - 1:
            ops.append((f"input_blocks.{idx}.0", "down", ch, ch))
            ops.append((f"input_blocks.{idx}", "push", ch, ch))
            chans.append(ch)
            idx += 1
            ds *= 2
    ops += [("middle_block.0", "res", ch, ch), ("middle_block.1", "attn", ch, ch), ("middle_block.2", "res", ch, ch)]
    idx = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            skip = chans.pop()
            ops.append((f"output_blocks.{idx}", "pop", skip, ch + skip))
            ops.append((f"output_blocks.{idx}.0", "res", ch + skip, mc * m))
            ch = mc * m
            sub = 1
            if ds in att:
                ops.append((f"output_blocks.{idx}.{sub}", "attn", ch, ch))
                sub += 1
            if level and i == nrb:
                ops.append((f"output_blocks.{idx}.{sub}", "up", ch, ch))
                ds //= 2
            idx += 1
    ops.append(("out", "out", ch, c["out_channels"]))
    return ops, c


def unet_state_dict(seed: int = 0, config: dict | None = None) -> "OrderedDict[str, np.ndarray]":
    """Seeded UNetModel state dict with the reference's keys.  The reference zero-initialises the second convolution of every
    ResBlock, every attention output projection and the final convolution (``zero_module``), which makes eps == 0 for a
    fresh model; they are re-randomised here (scaled down, as a trained model would have them small but non-zero)."""
    ops, c = unet_structure(config)
    mc = c["model_channels"]
    ted = 4 * mc
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def conv(prefix, cout, cin, k, gain=1.0):
        fan_in = cin * k * k
        sd[prefix + ".weight"] = _normal(seed, prefix + ".weight", (cout, cin, k, k), gain * np.sqrt(1.0 / fan_in))
        sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (cout,), -1, 1) * np.float32(gain / np.sqrt(fan_in))

    def gn(prefix, ch):
        sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (ch,), 0.5, 1.5)
        sd[prefix + ".bias"] = _normal(seed, prefix + ".bias", (ch,), 0.1)

    _linear(sd, seed, "time_embed.0", ted, mc)
    _linear(sd, seed, "time_embed.2", ted, ted)
    for prefix, kind, cin, cout in ops:
        if kind == "conv_in":
            conv(prefix, cout, cin, 3)
        elif kind == "res":
            gn(prefix + ".in_layers.0", cin)
            conv(prefix + ".in_layers.2", cout, cin, 3)
            _linear(sd, seed, prefix + ".emb_layers.1", 2 * cout if c["use_scale_shift_norm"] else cout, ted)
            gn(prefix + ".out_layers.0", cout)
            conv(prefix + ".out_layers.3", cout, cout, 3, gain=0.5)
            if cin != cout:
                conv(prefix + ".skip_connection", cout, cin, 1)
        elif kind == "attn":
            gn(prefix + ".norm", cin)
            w = _normal(seed, prefix + ".qkv.weight", (3 * cin, cin, 1), np.sqrt(1.0 / cin))
            sd[prefix + ".qkv.weight"] = w
            sd[prefix + ".qkv.bias"] = _normal(seed, prefix + ".qkv.bias", (3 * cin,), 0.02)
            sd[prefix + ".proj_out.weight"] = _normal(seed, prefix + ".proj_out.weight", (cin, cin, 1), 0.5 * np.sqrt(1.0 / cin))
            sd[prefix + ".proj_out.bias"] = _normal(seed, prefix + ".proj_out.bias", (cin,), 0.02)
        elif kind == "down":
            conv(prefix + ".op", cout, cin, 3)
        elif kind == "up":
            conv(prefix + ".conv", cout, cin, 3)
        elif kind == "out":
            gn("out.0", cin)
            conv("out.2", cout, cin, 3, gain=0.5)
    return sd
