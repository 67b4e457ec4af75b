"""Development probe: black-box query serving.  One FAKEBOB iteration = one NES draw of
`samples_per_draw` antithetic queries (+ the clean query) of ONE clip through the defended system (adaptive_attack_eval.py:209-216:
samples_per_draw = samples_per_draw_batch_size = 200), plus the small-batch latency of a single query batch.
Prints one JSON line.  Usage: python nes_probe.py [samples_per_draw] [t_star] [mode]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import _lib, synthetic  # noqa: E402
from audiopure_b200.blackbox import EOT, NES, QueryLoss  # noqa: E402

CONFIG_JSON = os.path.join(ROOT, "diffusion-model-for-audio-defense_b200", "configs", "config.json")


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    t_star = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    L = 16000
    dw = ap.create_diffwave_model(None, CONFIG_JSON, reverse_timestep=t_star, state_dict=synthetic.wavenet_state_dict(seed=0),
                                  noise="philox", seed=1, mode=mode)
    rx = ap.ResNeXtClassifier(synthetic.resnext_state_dict(seed=0))
    system = ap.AcousticSystem(classifier=rx, transform=ap.sc09_transform(), defender=dw, defense_type="wave",
                               check_int16_range=False)
    x = torch.from_numpy(synthetic.synthetic_waveforms(1, L, seed=5)).cuda()
    y = torch.tensor([3], device="cuda")
    loss_fn = QueryLoss("Entropy")
    nes = NES(S, S, 0.001, EOT(system, loss_fn, 1, 1, False))
    lib = _lib.load()
    st = _lib.stream_ptr()
    R = S + 1
    q = torch.empty(R, 1, L, device="cuda")
    loss = torch.randn(1, R, device="cuda")
    grad = torch.empty(1, L, device="cuda")
    scores = torch.randn(R, 10, device="cuda")
    yy = y.repeat(R)

    def forward_only():
        with torch.no_grad():
            system(q)

    out = {"samples_per_draw": S, "t_star": t_star, "mode": mode, "length": L}
    _lib.check(lib.ap_nes_perturb(x.data_ptr(), 0.001, None, 1, 0, 1, q.data_ptr(), 1, S, L, st))
    out["nes_iteration_ms"] = timed(lambda: nes(x, y))
    out["forward_of_the_same_batch_ms"] = timed(forward_only)
    out["ap_nes_perturb_us"] = 1e3 * timed(lambda: lib.ap_nes_perturb(x.data_ptr(), 0.001, None, 1, 0, 1, q.data_ptr(), 1, S, L, st), 50)
    out["ap_nes_gradient_us"] = 1e3 * timed(lambda: lib.ap_nes_gradient(loss.data_ptr(), None, 1, 0, 1, 1.0, 0, grad.data_ptr(), 1, S, L, st), 50)
    out["ap_query_loss_us"] = 1e3 * timed(lambda: loss_fn.loss_and_decision(scores, yy), 50)
    out["queries_per_s"] = 1e3 * R / out["nes_iteration_ms"]
    out["fakebob_200_iterations_s"] = 0.2 * out["nes_iteration_ms"]
    # what the reference's NES does around the model on the same GPU (torch eager ops, _NES.py:18-24,47)
    def eager_nes_ops():
        noise = torch.randn([1, S // 2, 1, L], device="cuda")
        noise = torch.cat((noise, -noise), 1)
        noise = torch.cat((torch.zeros_like(x).unsqueeze(1), noise), 1)
        ev = (noise * 0.001 + x.unsqueeze(1)).view(-1, 1, L)
        return torch.mean(loss.view(1, R)[..., 1:].unsqueeze(2).unsqueeze(3) * noise[:, 1:], 1), ev
    out["torch_eager_nes_ops_us"] = 1e3 * timed(eager_nes_ops, 50)
    # small-batch latency through a CUDA graph
    for B in (1, 8):
        xb = torch.from_numpy(synthetic.synthetic_waveforms(B, L, seed=6)).cuda()
        with torch.no_grad():
            system(xb)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            dw._offset = 0
            with torch.cuda.graph(g):
                yb = system(xb)
        out[f"graph_query_B{B}_ms"] = timed(g.replay, 20)
        with torch.no_grad():
            out[f"eager_query_B{B}_ms"] = timed(lambda: system(xb), 20)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
