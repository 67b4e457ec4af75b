"""Study (not a test, not collected by pytest): what would FP8 (e4m3) operands for the dilated convolution of the DiffWave
residual block (GEMM-1, 75 % of the network's FLOPs) cost in accuracy?  DESIGN.md "What comes next", item 1: the north-star
certification time needs more than the bf16 roofline allows, and whether an `fp8` mode is admissible is an accuracy question that
can be answered on the CPU oracle before any kernel work.

Runs the oracle's WaveNet restatement (oracle/audiopure_oracle.py::wavenet_forward, same arithmetic) with the A operand (u, the
residual stream) and the B operand (the folded dilated-conv weights) of GEMM-1 rounded to
  bf16            : what the product's default mode does,
  fp8_tensor      : e4m3, one scale per activation tensor and one per output channel of the weights,
  fp8_channel     : e4m3, one activation scale per input channel (foldable into the weights) and per-output-channel weight scales,
  fp8_w_only      : e4m3 weights, bf16 activations,
  fp8_skip        : GEMM-1 in bf16, but the gate output o and the skip weights of the K = 9216 skip GEMM (k2_head) in e4m3 -- o is
                    in (-1, 1) and is the tensor k1 writes and k2 re-reads (38 GB per 128 waveforms), so storing it in 8 bits would
                    halve that traffic,
and reports, against the fp32 forward: rel-L2 of eps at t = 65 and t = 1, and of the one-shot x0 estimate at t* = 66.

    python tests/studies/operand_precision_study.py            # ~1 min on CPU; prints one JSON line
Random-init weights (no checkpoint in this environment): the numbers are indicative of rounding behaviour, not of a trained model.
"""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import audiopure_b200  # noqa: E402,F401
import audiopure_oracle as orc  # noqa: E402
from audiopure_b200 import synthetic  # noqa: E402

E4M3_MAX = 448.0


def q_e4m3(x, scale):
    return (x / scale).to(torch.float8_e4m3fn).to(torch.float32) * scale


def q_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def quantise(u, wd, mode):
    """(A operand u (B,C,L), B operand wd (2C,C,3)) as GEMM-1 would see them."""
    if mode == "fp32":
        return u, wd
    if mode in ("bf16", "fp8_skip"):
        return q_bf16(u), q_bf16(wd)
    w_scale = wd.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-30) / E4M3_MAX          # per output channel
    if mode == "fp8_w_only":
        return q_bf16(u), q_e4m3(wd, w_scale)
    if mode == "fp8_tensor":
        return q_e4m3(u, u.abs().amax().clamp_min(1e-30) / E4M3_MAX), q_e4m3(wd, w_scale)
    if mode == "fp8_channel":
        a_scale = u.abs().amax(dim=(0, 2), keepdim=True).clamp_min(1e-30) / E4M3_MAX          # per input channel
        return q_e4m3(u, a_scale), q_e4m3(wd, w_scale)
    raise KeyError(mode)


def wavenet(sd, audio, t, mode):
    """oracle wavenet_forward with the GEMM-1 operands passed through `quantise` (everything else fp32)."""
    w = lambda k: torch.as_tensor(sd[k])
    wn = lambda p: orc.fold_weight_norm(sd[p + ".weight_g"], sd[p + ".weight_v"])
    x = torch.as_tensor(audio)
    steps = t * torch.ones(x.shape[0], 1)
    h = torch.relu(F.conv1d(x, wn("init_conv.0.conv"), w("init_conv.0.conv.bias")))
    emb = orc.step_embedding(steps, 128)
    emb = orc._swish(F.linear(emb, w("residual_layer.fc_t1.weight"), w("residual_layer.fc_t1.bias")))
    emb = orc._swish(F.linear(emb, w("residual_layer.fc_t2.weight"), w("residual_layer.fc_t2.bias")))
    skip_total = torch.zeros_like(h)
    for n in range(36):
        p = f"residual_layer.residual_blocks.{n}"
        d = 2 ** (n % 12)
        C = h.shape[1]
        u = h + F.linear(emb, w(p + ".fc_t.weight"), w(p + ".fc_t.bias")).reshape(-1, C, 1)
        uq, wq = quantise(u, wn(p + ".dilated_conv_layer.conv"), mode)
        a = F.conv1d(uq, wq, w(p + ".dilated_conv_layer.conv.bias"), dilation=d, padding=d)
        o = torch.tanh(a[:, :C]) * torch.sigmoid(a[:, C:])
        h = (u + F.conv1d(o, wn(p + ".res_conv"), w(p + ".res_conv.bias"))) * math.sqrt(0.5)
        ws, os_ = wn(p + ".skip_conv"), o
        if mode == "fp8_skip":
            ws = q_e4m3(ws, ws.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-30) / E4M3_MAX)
            os_ = q_e4m3(o, torch.tensor(1.0 / E4M3_MAX * 1.0))          # |o| < 1: a fixed scale, no amax pass
        skip_total = skip_total + F.conv1d(os_, ws, w(p + ".skip_conv.bias"))
    s = skip_total * math.sqrt(1.0 / 36)
    y = torch.relu(F.conv1d(s, wn("final_conv.0.conv"), w("final_conv.0.conv.bias")))
    return F.conv1d(y, w("final_conv.2.conv.weight"), w("final_conv.2.conv.bias"))


def rel(a, b):
    return float((a - b).norm() / b.norm())


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    sd = synthetic.wavenet_state_dict(seed=0)
    hp = orc.diffusion_hyperparams()
    B, L = 4, 4000
    x0 = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=1234))
    out = {"B": B, "L": L, "weights": "random init (synthetic.wavenet_state_dict(seed=0))"}
    with torch.no_grad():
        ab = hp["Alpha_bar"][65]
        xt = torch.sqrt(ab) * (x0 + 0.5 * torch.from_numpy(synthetic.host_noise((B, 1, L), 7, 0)))    # smoothing input at sigma = 0.5
        x1 = x0 + 0.01 * torch.from_numpy(synthetic.host_noise((B, 1, L), 7, 1))
        ref65, ref1 = wavenet(sd, xt, 65.0, "fp32"), wavenet(sd, x1, 1.0, "fp32")
        assert rel(orc.wavenet_forward(sd, xt, 65.0 * torch.ones(B, 1)), ref65) < 1e-6               # same function as the oracle
        x0_ref = orc.predict_x0_from_eps(hp, xt, 65, ref65)
        for mode in ("bf16", "fp8_skip", "fp8_w_only", "fp8_channel", "fp8_tensor"):
            e65, e1 = wavenet(sd, xt, 65.0, mode), wavenet(sd, x1, 1.0, mode)
            out[mode] = {"eps_rel_l2_t65": rel(e65, ref65), "eps_rel_l2_t1": rel(e1, ref1),
                         "one_shot_x0_rel_l2_t66": rel(orc.predict_x0_from_eps(hp, xt, 65, e65), x0_ref)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
