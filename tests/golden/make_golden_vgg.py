"""Golden logits and input gradients of the reference's VGG, WideResNet and DenseNet classifiers
(audio_models/ConvNets_SpeechCommands/models/{vgg,wideresnet,densenet}.py, built through the reference factory models.create_model) on the mel features already pinned in reference_golden*.npz.

    python tests/golden/make_golden_vgg.py        # in the build container; writes reference_golden_vgg.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from make_golden import synthetic, to_torch_sd  # noqa: E402


def main():
    mg.install_shim()
    torch.set_num_threads(os.cpu_count())
    import models
    from models.vgg import vgg11_bn
    g = dict(np.load(os.path.join(HERE, "reference_golden.npz")))
    gg = dict(np.load(os.path.join(HERE, "reference_golden_grad.npz")))
    out = {}
    # the factory's fallback IS vgg19_bn (models/__init__.py:44-45) -- and 'resnet18' falls through to it as well (:18-21)
    for name, key, depth in (("vgg19_bn", "vgg19", 19), ("resnet18", None, 19), ("vgg11_bn", "vgg11", 11)):
        net = (vgg11_bn(num_classes=10, in_channels=1) if depth == 11 else models.create_model(name, 10, 1)).eval()
        assert type(net).__name__ == "VGG"
        if key is None:
            continue
        net.load_state_dict(to_torch_sd(synthetic.vgg_state_dict(depth=depth, seed=0)))
        for p in net.parameters():
            p.requires_grad_(False)
        with torch.no_grad():
            out[f"{key}_logits"] = net(torch.from_numpy(g["mel_sc09"])).numpy()
        sr = torch.from_numpy(gg["resnext_in_spec"]).clone().requires_grad_(True)
        (gs,) = torch.autograd.grad(net(sr), sr, torch.from_numpy(gg["resnext_g_logits"]))
        out[f"{key}_grad"] = gs.numpy()
    # WideResNet-28-10 through the factory (models/__init__.py:29-30) and a narrow WRN-16-1 (equal-width first block)
    from models.wideresnet import WideResNet
    for key, depth, k in (("wrn28_10", 28, 10), ("wrn16_1", 16, 1)):
        net = (models.create_model("wideresnet28_10", 10, 1) if depth == 28 else
               WideResNet(depth=depth, widen_factor=k, dropRate=0, num_classes=10, in_channels=1)).eval()
        net.load_state_dict(to_torch_sd(synthetic.wideresnet_state_dict(depth=depth, widen_factor=k, seed=0)))
        for p in net.parameters():
            p.requires_grad_(False)
        with torch.no_grad():
            out[f"{key}_logits"] = net(torch.from_numpy(g["mel_sc09"])).numpy()
        sr = torch.from_numpy(gg["resnext_in_spec"]).clone().requires_grad_(True)
        (gs,) = torch.autograd.grad(net(sr), sr, torch.from_numpy(gg["resnext_g_logits"]))
        out[f"{key}_grad"] = gs.numpy()
    # DenseNet-BC-100-12 through the factory (models/__init__.py:37-38) and a shallow BC-22-12 whose third block starts at 33
    # channels (an input width that is not a multiple of 4)
    from models.densenet import DenseNet
    for key, depth in (("densenet100_12", 100), ("densenet22_12", 22)):
        net = (models.create_model("densenet_bc_100_12", 10, 1) if depth == 100 else
               DenseNet(depth=depth, growthRate=12, compressionRate=2, num_classes=10, in_channels=1)).eval()
        net.load_state_dict(to_torch_sd(synthetic.densenet_state_dict(depth=depth, growth_rate=12, seed=0)))
        for p in net.parameters():
            p.requires_grad_(False)
        with torch.no_grad():
            out[f"{key}_logits"] = net(torch.from_numpy(g["mel_sc09"])).numpy()
        sr = torch.from_numpy(gg["resnext_in_spec"]).clone().requires_grad_(True)
        (gs,) = torch.autograd.grad(net(sr), sr, torch.from_numpy(gg["resnext_g_logits"]))
        out[f"{key}_grad"] = gs.numpy()
    path = os.path.join(HERE, "reference_golden_vgg.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
