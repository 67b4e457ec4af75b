#!/usr/bin/env python
"""Benchmark of the AudioPure purify-and-classify hot path (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host CPU cores

A step = one pass of the hot path over one batch of synthetic SC09-shaped waveforms:
  DiffWave DDPM purification (t* = 2, bf16 tensor-core mode, in-kernel Philox noise) -> log-mel -> ResNeXt-29 8x64 -> argmax.
Workload of the headline line at every N: BASELINE.json configs[1] (batch 512 x 1 s @ 16 kHz per GPU; replicas only --
independent waveforms need no collective).  Rank 0 prints ONE JSON line.  `value` is timed with the inputs resident in HBM;
`e2e` goes through the public API from pinned host buffers (H2D copy of the batch and D2H read of the predictions inside the
timed region).

Besides the headline the same line carries (key `configs`) every other BASELINE config measured in the same process --
configs[1] in the fp32-class modes (bf16x3, fp32 FFMA), configs[3] (reverse-SDE purifier, t* in {1, 5, 10} x input noise
sigma in {0.25, 1.0}), configs[4] (M5; KWS mel + RCNN_KWS on 2 s clips), configs[2] (certification, N = 100 100 draws when 8 GPUs
are present, a bounded leg otherwise) -- a `kernels` table (achieved GB/s or TFLOP/s of each hot kernel against the measured
peak), and two reported baselines: `cpu_baseline` (the CPU port of the reference algorithm on the host cores) and `eager_gpu`
(the same algorithm through PyTorch eager / cuDNN / cuBLAS on this GPU, TF32 on and off): the real bar.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "purified+classified waveforms/sec"
UNIT = "waveforms/s"
T_STAR = 2
LENGTH = 16000
WAVENET_GFLOP = 606.10            # per waveform per network evaluation
K1_GFLOP_PER_WAVEFORM = 14.680    # k1_layer, per layer per 1 s waveform: dilated conv 12.583 + res 1x1 2.097 (skip 1x1 is in k2_head)
K2_GFLOP_PER_WAVEFORM = 77.6      # k2_head: skip path of all 36 layers (75.5) + head (2.1)
K2_BYTES_PER_WAVEFORM = 36 * 16000 * 256 * 2 + 16000 * 4     # O of every layer read once (bf16) + eps written
RESNEXT_GFLOP = 10.77
REF_SAMPLE = 2                    # waveforms per step of the CPU arms: a bounded sample of the 512-waveform step
CERT_BATCH = 592                  # certification micro-batch: 4 workspace chunks of 148 waveforms, no ragged chunk


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=512, help="waveforms per GPU per step")
    p.add_argument("--mode", default="bf16", choices=["bf16", "fp16", "fp32", "bf16x3"])
    p.add_argument("--chunk", type=int, default=0, help="waveforms per workspace chunk (0 = library default)")
    p.add_argument("--certify-draws", type=int, default=-1,
                   help="draws of the certification leg (-1 = N = 100100 at 8 GPUs, 8192 per GPU otherwise; 0 = skip)")
    p.add_argument("--cpu-sample", type=int, default=REF_SAMPLE,
                   help="waveforms per pass of the cpu_baseline leg (0 = skip); 1 warm-up + 3 timed passes, median")
    p.add_argument("--workload", default="sc09", choices=["sc09", "sde", "m5", "kws"],
                   help="headline workload: sc09 = BASELINE configs[1]; sde = configs[3] (reverse-SDE purifier, --t-star 1..10); "
                        "m5 / kws = configs[4] (DDPM purifier + raw-waveform M5 / mel(400,200,32) + RCNN_KWS)")
    p.add_argument("--t-star", type=int, default=T_STAR)
    p.add_argument("--length", type=int, default=LENGTH)
    p.add_argument("--extras", default="auto", choices=["auto", "all", "scaling", "none"],
                   help="the `configs` / `kernels` / `eager_gpu` legs: auto = all at 1 GPU, the multi-GPU subset (sde t*=10, m5, "
                        "kws 2 s, certification) at N > 1")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tflops_sustained": d.get("bf16_tflops_sustained"), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def workload_name(workload, t_star, length, batch, sigma_in=None):
    if workload == "spec":
        return (f"SURVEY 8(f)4 Diffusion-Spec: SC09 log-mel -> spectrogram UNet (128 ch, 30 ResBlocks, 15 attention blocks) reverse-SDE "
                f"t*={t_star} -> ResNeXt-29 8x64, batch {batch} x {length / 16000:g} s @ 16 kHz per GPU, random-init weights, tf32 convolutions")
    purifier = {"sc09": f"DDPM t*={t_star}", "sde": f"reverse-SDE (Euler-Maruyama) t*={t_star}",
                "m5": f"DDPM t*={t_star}", "kws": f"DDPM t*={t_star}"}[workload]
    head = {"sc09": "SC09 log-mel + ResNeXt-29 8x64", "sde": "SC09 log-mel + ResNeXt-29 8x64", "m5": "M5 raw-waveform classifier",
            "kws": "KWS log-mel (n_fft 400, hop 200, 32 mels) + RCNN_KWS"}[workload]
    name = {"sc09": "BASELINE configs[1]", "sde": "BASELINE configs[3]", "m5": "BASELINE configs[4] (M5)",
            "kws": "BASELINE configs[4] (RCNN_KWS)"}[workload]
    noise = f", input noise sigma={sigma_in}" if sigma_in else ""
    return (f"{name}: DiffWave(36 layers, C=256) {purifier} + {head}, batch {batch} x {length / 16000:g} s @ 16 kHz per GPU"
            f"{noise}, random-init weights")


def workload_config(args, world):
    return {"workload": workload_name(args.workload, args.t_star, args.length, args.batch),
            "batch_per_gpu": args.batch, "length": args.length, "t_star": args.t_star, "mode": args.mode,
            "parallelism": f"replicas x{world} (no data-path collective)",
            "l2": "256 MiB buffer rewritten between timed iterations; activations per chunk >> L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def oracle_pipeline(n_waveforms: int, seed: int = 1234):
    """The reference algorithm (CPU restatement under oracle/, pinned to the reference's outputs by tests/): DDPM t*=2 ->
    mel -> ResNeXt -> argmax, on `n_waveforms` synthetic 1 s clips.  Returns seconds."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import audiopure_oracle as orc
    import audiopure_b200  # noqa: F401
    from audiopure_b200 import synthetic
    st = oracle_pipeline.__dict__.setdefault("state", {})
    if not st:
        st["sd"] = synthetic.wavenet_state_dict(seed=0)
        st["rx"] = synthetic.resnext_state_dict(seed=0)
        st["hp"] = orc.diffusion_hyperparams()
    x = synthetic.synthetic_waveforms(n_waveforms, LENGTH, seed=seed)
    zs = [synthetic.host_noise(x.shape, 2024, i) for i in range(T_STAR)]
    t0 = time.perf_counter()
    with torch.no_grad():
        y = orc.ddpm_forward(st["sd"], x, st["hp"], T_STAR, orc.NoiseSource(zs))
        spec = orc.mel_db(y, **orc.MEL_SC09)
        pred = orc.resnext_forward(st["rx"], spec).argmax(1)
    _ = pred.tolist()
    return time.perf_counter() - t0


def cpu_baseline(sample: int):
    """1 warm-up + 3 timed passes of the CPU port on `sample` waveforms, median."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle_pipeline(sample)
    ts = sorted(oracle_pipeline(sample, seed=1234 + i) for i in range(3))
    return {"value": sample / ts[1], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} of the {512} waveforms of one step per pass; 1 warm-up + 3 timed passes, median "
                      f"({ts[1]:.2f} s per pass; oracle port = the reference algorithm in torch CPU ops, {cores} threads)"}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = REF_SAMPLE
    for _ in range(args.warmup):
        oracle_pipeline(sample)
    t = 0.0
    for i in range(args.steps):
        t += oracle_pipeline(sample, seed=1234 + i)
    value = sample * args.steps / t
    cfg = workload_config(args, args.gpus)
    cfg["batch_per_gpu"] = sample
    cfg["parallelism"] = f"host CPU, {cores} threads, rank 0 only"
    cfg["l2"] = "n/a (CPU)"
    cfg["sample"] = (f"each step is a {sample}-waveform sample of the {args.batch}-waveform step of the GPU arm (same per-waveform work: "
                     "DDPM t*=2 + log-mel + ResNeXt-29)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} waveforms per step x {args.steps} steps after {args.warmup} warm-up steps "
                                       f"(torch CPU ops, {cores} threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class Harness:
    """Shared state of one rank: the DiffWave network (one handle, reused by every workload), classifiers, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import audiopure_b200 as ap
        from audiopure_b200 import _lib, synthetic
        self.torch, self.dist, self.ap, self._lib, self.synthetic = torch, dist, ap, _lib, synthetic
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.dev = torch.device("cuda", self.local_rank)
        self.lib = _lib.load()
        self.cfg_json = os.path.join(ROOT, "diffusion-model-for-audio-defense_b200", "configs", "config.json")
        self.sd = synthetic.wavenet_state_dict(seed=0)
        self.dw = ap.create_diffwave_model(None, self.cfg_json, reverse_timestep=args.t_star, state_dict=self.sd, noise="philox",
                                           seed=2024 + self.rank, mode=args.mode)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > L2 (126 MB)
        self._cls = {}

    # -- modules --------------------------------------------------------------------------------------------------
    def classifier(self, kind):
        ap, syn = self.ap, self.synthetic
        if kind not in self._cls:
            if kind == "m5":
                self._cls[kind] = (None, ap.M5Classifier(syn.m5_state_dict(seed=0)))
            elif kind == "kws":
                self._cls[kind] = (ap.kws_transform(), ap.KWSClassifier(syn.kws_state_dict(seed=0)))
            else:
                self._cls[kind] = (ap.sc09_transform(), ap.ResNeXtClassifier(syn.resnext_state_dict(seed=0)))
        return self._cls[kind]

    def system(self, workload, t_star):
        ap = self.ap
        self.dw.reverse_timestep = t_star
        ns = argparse.Namespace(ddpm_path=None, ddpm_config=self.cfg_json, t=t_star, score_type="guided_diffusion", rand_t=False,
                                t_delta=0, use_bm=False, sample_step=1)
        transform, classifier = self.classifier(workload if workload in ("m5", "kws") else "sc09")
        if workload == "spec":      # Diffusion-Spec: the UNet reverse-SDE purifier acts on the log-mel spectrogram (defense_type 'spec')
            if "spec" not in self._cls:
                self._cls["spec"] = ap.RevImprovedDiffusion(ns, state_dict=self.synthetic.unet_state_dict(seed=0), seed=3024 + self.rank)
            self._cls["spec"].args.t = t_star
            return ap.AcousticSystem(classifier=classifier, transform=transform, defender=self._cls["spec"], defense_type="spec")
        defender = ap.RevDiffWave(ns, diffwave=self.dw) if workload == "sde" else self.dw
        return ap.AcousticSystem(classifier=classifier, transform=transform, defender=defender, defense_type="wave")

    # -- timing ---------------------------------------------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world > 1:
            t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def timed(self, fn, steps, warmup=0, profile=False):
        """(ms for `steps` steps: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks; launches)"""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            self.lib.ap_diffwave_profile(self.dw.model._handle, 1)
        l0 = self._lib.launch_count()
        e0.record()
        for _ in range(steps):
            self.flush.zero_()
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        return self.max_over_ranks(ms), self._lib.launch_count() - l0

    def kernel_time(self, fns, reps=7):
        """median ms per launch: `fns` is one callable or a list of callables doing the same work on DIFFERENT buffers; the
        list is enqueued back to back between one pair of CUDA events (amortises the ~5 us event overhead of a 15 us kernel) after
        an L2 flush, so every launch streams from HBM."""
        torch = self.torch
        if callable(fns):
            fns = [fns]
        for fn in fns:
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for fn in fns:
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / len(fns))
        return statistics.median(ts)

    def inputs(self, batch, length, sigma_in=None):
        torch = self.torch
        x = torch.from_numpy(self.synthetic.synthetic_waveforms(batch, length, seed=1234 + self.rank))
        if sigma_in:   # smoothing-level inputs x + sigma * randn (no clamp): value ranges, not cost
            x = x + sigma_in * torch.from_numpy(self.synthetic.host_noise(tuple(x.shape), 77 + self.rank, 0))
        return x.pin_memory()

    def throughput(self, workload, t_star, batch, length, steps, warmup, mode=None, sigma_in=None, e2e=False, profile=False):
        """waveforms/s of one workload at this rank count.  Returns a dict (value; e2e when asked)."""
        torch = self.torch
        if mode is not None:
            self.dw.model.set_mode(mode)
        system = self.system(workload, t_star)
        x_host = self.inputs(batch, length, sigma_in)
        x_dev = x_host.to(self.dev)
        pred_host = torch.empty(batch, dtype=torch.int64).pin_memory()

        def step_resident():
            return system(x_dev).argmax(1)

        def step_e2e():
            xd = x_host.to(self.dev, non_blocking=True)
            pred_host.copy_(system(xd).argmax(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return pred_host

        ms, launches = self.timed(step_resident, steps, warmup, profile=profile)
        out = {"value": self.world * batch * steps / (ms / 1e3), "ms_per_step": ms / steps, "launches": int(launches),
               "steps": steps, "warmup": warmup, "h2d": int(x_host.numel() * 4), "d2h": int(pred_host.numel() * 8)}
        if e2e:
            step_e2e()
            ms2, _ = self.timed(step_e2e, steps)
            out["e2e"] = self.world * batch * steps / (ms2 / 1e3)
            out["e2e_ms_per_step"] = ms2 / steps
        return out


def certification_leg(h: Harness, draws: int):
    """BASELINE configs[2]: RobustCertificate.certify of ONE input = n_0 = 100 + N draws, sigma = 0.5 (t* = 66, one-shot denoise),
    draws sharded over the ranks, one all-reduce per decision.  Two decisions are timed back to back (the second input's draws
    overlap the first input's host statistics) after one small warm-up decision (builds the CUDA graph); wall clock, max over
    ranks."""
    torch, ap = h.torch, h.ap
    h.dw.model.set_mode(h.args.mode)
    transform, classifier = h.classifier("sc09")
    rc = ap.RobustCertificate(classifier=classifier, transform=transform, denoiser=h.dw, seed=99)
    x = torch.from_numpy(h.synthetic.synthetic_waveforms(2, LENGTH, seed=4242)).to(h.dev)
    y = torch.zeros(2, dtype=torch.int64, device=h.dev)
    rc.certify(x[:1], y[:1], sigma=0.5, n_0=100, n=CERT_BATCH * h.world, batch_size=CERT_BATCH)
    h.barrier()
    l0 = h._lib.launch_count()
    t0 = time.perf_counter()
    y_pred, radius = rc.certify(x, y, sigma=0.5, n_0=100, n=draws, batch_size=CERT_BATCH)
    _ = y_pred.tolist(), radius.tolist()
    h.barrier()
    dt = h.max_over_ranks(time.perf_counter() - t0) / 2
    counts = rc.smooth_predict(x[:1], num_sampling=CERT_BATCH * h.world, sigma=0.5, batch_size=CERT_BATCH)
    h.dw.reverse_timestep = h.args.t_star
    full = draws >= 100000
    return {"workload": "BASELINE configs[2]: randomized-smoothing certification, sigma=0.5 (t*=66, one-shot denoise), "
                        f"n_0=100 + N={draws} draws per decision sharded over {h.world} GPU(s), micro-batch {CERT_BATCH}, one all-reduce "
                        "of the vote vectors per decision", "draws": draws + 100, "sec_per_decision": dt,
            "draws_per_s": (draws + 100) / dt, "full_N100000_measured": full,
            "projected_sec_per_N100000": dt if full else 100100 / ((draws + 100) / dt), "sigma": 0.5, "t_star": 66,
            "pipeline": "CUDA-graph micro-batch (fused smoothing-input init + x0 epilogue), side-stream all-reduce + D2H, host "
                        "Clopper-Pearson of decision i overlapped with the draws of decision i+1",
            "decisions_timed": 2, "y_pred": y_pred.tolist(), "radius": [round(r, 4) for r in radius.tolist()],
            "votes_sample": counts.tolist(), "gpu_launches": int(h._lib.launch_count() - l0),
            "bf16_floor_sec": "5.5 s per decision at 100 % of the sustained bf16 peak on 8 GPUs: 'a few seconds' is below the bf16 "
                              "roofline (DESIGN.md section 5)"}


def unet_gflop(synthetic) -> float:
    """Algorithmic GFLOP of one UNet evaluation on a 32 x 32 spectrogram (2 x MACs of every convolution + QK^T and PV), walked
    over the same op list the kernels execute."""
    ops, cfg = synthetic.unet_structure(None)
    H, fl = cfg["image_size"], 0
    for _, kind, cin, cout in ops:
        px = H * H
        if kind in ("conv_in", "out"):
            fl += 2 * 9 * cin * cout * px
        elif kind == "res":
            fl += 2 * 9 * cin * cout * px + 2 * 9 * cout * cout * px + (2 * cin * cout * px if cin != cout else 0)
        elif kind == "attn":
            fl += 2 * cin * 3 * cin * px + 2 * cin * cin * px + 4 * px * px * cin
        elif kind == "down":
            H //= 2
            fl += 2 * 9 * cin * cout * H * H
        elif kind == "up":
            H *= 2
            fl += 2 * 9 * cin * cout * H * H
    return fl / 1e9


def kernels_table(h: Harness, peaks, prof_main):
    """Achieved rate of each hot kernel against the measured peak, timed alone with CUDA events (L2 flushed), B = 512 x 1 s."""
    torch, lib, _lib = h.torch, h.lib, h._lib
    B, L = 512, LENGTH
    n = B * L
    st = _lib.stream_ptr()
    NSET = 6                                     # 6 x 3 x 32.8 MB: consecutive launches never find their operands in the 126 MB L2
    xs = [torch.randn(B, L, device=h.dev) for _ in range(NSET)]
    es = [torch.randn(B, L, device=h.dev) for _ in range(NSET)]
    outs = [torch.empty(B, L, device=h.dev) for _ in range(NSET)]
    z = torch.randn(B, L, device=h.dev)
    hbm = peaks["hbm_gbs"]
    rows = []

    def row(name, what, ms, nbytes=None, gflop=None, peak_tf=None, note=None):
        r = {"kernel": name, "what": what, "ms": ms}
        if nbytes is not None:
            r.update({"bound": "hbm", "algorithmic_bytes": nbytes, "achieved": nbytes / ms / 1e6, "peak": hbm, "unit": "GB/s",
                      "frac": nbytes / ms / 1e6 / hbm})
        if gflop is not None:
            r.update({"tflops": gflop / ms})
            if peak_tf:
                r.update({"bound": "tensor", "achieved": gflop / ms, "peak": peak_tf, "unit": "TFLOP/s", "frac": gflop / ms / peak_tf})
        if note:
            r["note"] = note
        rows.append(r)

    ck = lambda rc: _lib.check(rc, "kernels_table")
    each = lambda f: [(lambda i=i: ck(f(i))) for i in range(NSET)]
    ms = h.kernel_time(each(lambda i: lib.ap_ddpm_step(xs[i].data_ptr(), es[i].data_ptr(), 0.0115, 0.9999, 0.0082, None, 7, 0, B, L, st)))
    row("ew_kernel<DdpmStepOp> (Philox)", "x = (x - c eps)/sqrt(alpha) + sigma z, in-kernel noise: x rd, eps rd, x wr", ms, 12 * n)
    ms = h.kernel_time(each(lambda i: lib.ap_ddpm_step(xs[i].data_ptr(), es[i].data_ptr(), 0.0115, 0.9999, 0.0082, outs[i].data_ptr(), 0, 0,
                                                        B, L, st)))
    row("ew_kernel<DdpmStepOp> (host noise)", "same with z read from HBM", ms, 16 * n)
    ms = h.kernel_time(each(lambda i: lib.ap_diffuse(xs[i].data_ptr(), 0.9997, 0.0245, None, 7, 0, outs[i].data_ptr(), B, L, st)))
    row("ew_kernel<DiffuseOp> (Philox)", "x_t = a x0 + b z: x0 rd, x_t wr", ms, 8 * n)
    ms = h.kernel_time(each(lambda i: lib.ap_smooth_inputs(xs[i].data_ptr(), 0.5, 0.8944, None, 7, 0, outs[i].data_ptr(), B, L, st)))
    row("ew_kernel<SmoothOp> (Philox)", "x_in[b] = scale (x + sigma z[b]): one 64 KB input broadcast, x_in wr", ms, 4 * n,
        note="write-only stream: one Philox block + Box-Muller (8 MUFU) per 16 B written")
    ms = h.kernel_time(each(lambda i: lib.ap_predict_x0(xs[i].data_ptr(), es[i].data_ptr(), 1.118, 0.5, outs[i].data_ptr(), B, L, st)))
    row("ew_kernel<PredictX0Op>", "x0 = a x_t - b eps: 2 rd, 1 wr", ms, 12 * n)
    del xs, es, outs, z
    tr, rx = h.classifier("sc09")
    wav = torch.from_numpy(h.synthetic.synthetic_waveforms(B, L, seed=5)).to(h.dev)
    ms = h.kernel_time(lambda: tr(wav))
    rows.append({"kernel": "log-mel (SC09): mel_prep_kernel + k_mel (tcgen05 DFT GEMM, bf16 hi/lo split operands, power -> filterbank -> dB "
                           "fused into the epilogue)", "what": "frames[32B x 2048] . basis[2048 x 2048] -> power -> 32 mels -> dB, 512 waveforms",
                 "ms": ms, "bound": "tensor", "achieved": B * 0.2687 / ms, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
                 "frac": B * 0.2687 / ms / peaks["tflops_burst"], "mma_flops_per_algorithmic_flop": 3,
                 "mma_tflops": 3 * B * 0.2687 / ms, "algorithmic_bytes": B * (L * 4 + 32 * 32 * 4),
                 "note": "achieved = the DFT-as-GEMM flop count (268.7 MFLOP per waveform) / time; the split issues 3 MMAs per product (fp32-class "
                         "accuracy: 2e-3 dB needs more than tf32), so the tensor pipe runs at mma_tflops; peak = burst bf16 (kernel timed alone); "
                         "ncu: tensor pipe 92.7 % active, 54 MB of DRAM traffic per 512 waveforms (profiles/r02_small_kernels_ncu_summary.txt)"})
    spec = tr(wav)
    ms = h.kernel_time(lambda: rx(spec))
    tf32_peak = measure_tf32_peak(h)
    row("ResNeXt-29 8x64 forward (convtc::k_conv tf32 tcgen05 + stem/pool)", "whole classifier, 512 spectrograms", ms,
        gflop=B * RESNEXT_GFLOP, peak_tf=tf32_peak,
        note="peak = torch.matmul 8192^3 with TF32 allowed, measured in this run (burst, best of 10)")
    del rx
    unet = h.ap.UNet(h.synthetic.unet_state_dict(seed=0), device=h.dev)
    xs_ = torch.randn(B, 1, 32, 32, device=h.dev)
    gs_ = torch.randn(128, 1, 32, 32, device=h.dev)
    ug = unet_gflop(h.synthetic)
    ms = h.kernel_time(lambda: unet.eps(xs_, 37.0))
    row("spectrogram UNet eps (Diffusion-Spec): k_conv tf32 tcgen05 + gn_tile_kernel + unet_attn_mma_kernel",
        f"one evaluation, 512 spectrograms ({ug:.2f} GFLOP each)", ms, gflop=B * ug, peak_tf=tf32_peak,
        note="113 convolutions, 61 GroupNorms and 15 attention blocks per evaluation in sub-batches of 128")
    ms = h.kernel_time(lambda: unet.eps_vjp(xs_[:128], 37.0, gs_))
    row("spectrogram UNet input gradient (ap_unet_eps_vjp: recomputed forward + reverse walk)",
        "g_x = (d eps / d x)^T g_eps, 128 spectrograms; flops counted as 2 x forward (recompute + data gradients)", ms,
        gflop=128 * 2 * ug, peak_tf=tf32_peak)
    del unet, xs_, gs_
    if prof_main and prof_main.get("k2_n", 0) > 0:
        ms2 = prof_main["k2_ms"] / prof_main["k2_n"]
        wf = prof_main["wf_per_k2"]
        r = {"kernel": "k2_head (skip path of all layers + head: K = 9216 tcgen05 GEMM, fused ReLU / 256->1 epilogue)",
             "what": f"{wf:g} waveforms per launch, timed inside the headline step", "ms": ms2,
             "bound": "tensor+hbm", "achieved": K2_GFLOP_PER_WAVEFORM * wf / ms2, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
             "frac": K2_GFLOP_PER_WAVEFORM * wf / ms2 / peaks["tflops_sustained"],
             "hbm_gbs": K2_BYTES_PER_WAVEFORM * wf / ms2 / 1e6, "hbm_frac": K2_BYTES_PER_WAVEFORM * wf / ms2 / 1e6 / hbm}
        rows.append(r)
    return rows


def measure_tf32_peak(h: Harness):
    torch = h.torch
    flag = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device=h.dev)
        b = torch.randn(8192, 8192, device=h.dev)
        torch.matmul(a, b)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = flag
    return 2 * 8192 ** 3 / best / 1e9


def eager_gpu_leg(h: Harness, ours_ms_per_waveform_e2e):
    """The reference algorithm through PyTorch eager (cuDNN / cuBLAS / torchaudio) on this GPU: TF32 on and off, 1 warm-up + 3
    timed passes, median.  A reported baseline (oracle/eager_gpu.py), the real 'kernel to beat'."""
    torch = h.torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import eager_gpu
    B = 64
    x = h.synthetic.synthetic_waveforms(B, LENGTH, seed=1234)
    zs = [h.synthetic.host_noise(x.shape, 2024, i) for i in range(T_STAR)]
    rx_sd = h.synthetic.resnext_state_dict(seed=0)
    out = {"what": "the reference algorithm (oracle torch ops: weight norm re-folded per call, 3 cuDNN convolutions per residual "
                   "block, fp32 activations, torchaudio mel, ResNeXt) executed by PyTorch eager on the same GPU",
           "batch": B, "protocol": "1 warm-up + 3 timed passes, median; inputs and noise resident on the GPU"}
    for name, tf32 in (("tf32_on", True), ("tf32_off", False)):
        try:
            sec, _ = eager_gpu.time_pipeline(h.sd, rx_sd, x, zs, allow_tf32=tf32, t_star=T_STAR, warmup=1, reps=3)
            out[name] = {"value": B / sec, "unit": UNIT, "sec_per_pass": sec}
        except Exception as exc:   # e.g. out of memory next to our own workspace
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    h = Harness(args)
    torch, lib, _lib, dw = h.torch, h.lib, h._lib, h.dw
    world, rank = h.world, h.rank
    if args.chunk:
        dw.model.reserve(args.chunk, args.length)
    B = args.batch
    extras = args.extras
    if extras == "auto":
        extras = "all" if world == 1 else "scaling"

    # ---- headline: args.workload in args.mode
    sampler = ClockSampler(h.local_rank)
    warm = max(args.warmup, 3)
    system = h.system(args.workload, args.t_star)
    x_host = h.inputs(B, args.length)
    x_dev = x_host.to(h.dev)
    pred_host = torch.empty(B, dtype=torch.int64).pin_memory()

    def step_resident():
        return system(x_dev).argmax(1)

    def step_e2e():
        xd = x_host.to(h.dev, non_blocking=True)
        pred_host.copy_(system(xd).argmax(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return pred_host

    for _ in range(warm):
        step_resident()
    if rank == 0:
        sampler.start()
    tc_mode = args.mode != "fp32"
    ms, launches = h.timed(step_resident, args.steps, profile=tc_mode)
    clocks = sampler.stop() if rank == 0 else None
    prof_ms, prof_n = (C.c_double * 2)(), (C.c_int * 2)()
    if tc_mode:
        _lib.check(lib.ap_diffwave_profile_read(dw.model._handle, prof_ms, prof_n), "profile_read")
        lib.ap_diffwave_profile(dw.model._handle, 0)
    value = world * B * args.steps / (ms / 1e3)
    step_e2e()
    ms_e2e, _ = h.timed(step_e2e, args.steps)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---- every other BASELINE config, in this process
    configs = {}
    if extras != "none":
        st_, wu_ = 2, 1

        def leg(key, **kw):
            r = h.throughput(**kw)
            configs[key] = {"workload": workload_name(kw["workload"], kw["t_star"], kw["length"], kw["batch"], kw.get("sigma_in")),
                            "mode": kw.get("mode") or args.mode, "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                            "steps": r["steps"], "warmup": r["warmup"], "n_gpus": world, "gpu_launches": r["launches"]}
            if "e2e" in r:
                configs[key]["e2e"] = {"value": r["e2e"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]}

        if args.workload == "sc09" and args.mode == "bf16":
            if extras == "all":
                leg("sc09_bf16x3", workload="sc09", t_star=2, batch=512, length=LENGTH, steps=st_, warmup=wu_, mode="bf16x3", e2e=True)
                leg("sc09_fp32_ffma", workload="sc09", t_star=2, batch=64, length=LENGTH, steps=1, warmup=1, mode="fp32")
                dw.model.set_mode("bf16")
                for t_star in (1, 5, 10):
                    for sigma_in in (0.25, 1.0):
                        leg(f"sde_t{t_star}_sigma{sigma_in}", workload="sde", t_star=t_star, batch=512, length=LENGTH, steps=st_,
                            warmup=wu_ if (t_star, sigma_in) == (1, 0.25) else 0, mode="bf16", sigma_in=sigma_in)
            else:
                leg("sde_t10_sigma1.0", workload="sde", t_star=10, batch=512, length=LENGTH, steps=st_, warmup=wu_, mode="bf16",
                    sigma_in=1.0)
            if extras == "all":
                leg("spec_sde_t2", workload="spec", t_star=2, batch=512, length=LENGTH, steps=st_, warmup=wu_, mode="bf16")
            leg("m5", workload="m5", t_star=2, batch=512, length=LENGTH, steps=3, warmup=wu_, mode="bf16", e2e=True)
            leg("kws_2s", workload="kws", t_star=2, batch=512, length=32000, steps=st_, warmup=wu_, mode="bf16", e2e=True)
            dw.model.set_mode(args.mode)
            dw.reverse_timestep = args.t_star

    # ---- certification leg (BASELINE configs[2])
    cert = None
    draws = args.certify_draws
    if draws < 0:
        draws = 100000 if world >= 8 else 8192 * world
    if draws > 0 and args.workload == "sc09" and extras != "none":
        cert = certification_leg(h, draws)

    peaks = measured_peaks()
    kernels = None
    if extras == "all" and args.workload == "sc09" and tc_mode and rank == 0:
        n_k2 = int(prof_n[1])
        kernels = kernels_table(h, peaks, {"k2_ms": prof_ms[1], "k2_n": n_k2,
                                           "wf_per_k2": B * args.t_star * args.steps / max(n_k2, 1) * args.length / LENGTH})
    eager = None
    if extras == "all" and args.workload == "sc09" and rank == 0:
        eager = eager_gpu_leg(h, ms_e2e / args.steps / B)
        eager["ours_e2e_over_eager_tf32_on"] = (e2e / eager["tf32_on"]["value"]) if "value" in eager.get("tf32_on", {}) else None
        eager["ours_e2e_over_eager_tf32_off"] = (e2e / eager["tf32_off"]["value"]) if "value" in eager.get("tf32_off", {}) else None

    if rank == 0:
        roofline = None
        if tc_mode and prof_n[0] > 0:
            k1_ms = prof_ms[0] / prof_n[0]
            # launches over the last (ragged) chunk process fewer waveforms: use the exact average per launch
            n_layers, evals = 36, args.t_star * args.steps
            avg_wf = B * n_layers * evals / prof_n[0]
            achieved = K1_GFLOP_PER_WAVEFORM * (args.length / 16000) * avg_wf / k1_ms        # GFLOP / ms == TFLOP/s
            roofline = {"kernel": "k1_layer (DiffWave residual block: tcgen05 implicit GEMM K=768/N=512 + K=256/N=256, fused "
                                  "gate / residual epilogues)", "bound": "tensor", "achieved": achieved,
                        "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                        "traffic": 24.430e6 * avg_wf if args.mode != "bf16x3" else None,
                        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel "
                                          "(profiles/r02_k1_staged_ncu_full_summary.txt: 3.127e9 B for a 128-waveform launch; algorithmic "
                                          "3.146e9 B), scaled to this run's waveforms per launch" if args.mode != "bf16x3" else "see profiles/ for the split kernel",
                        "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                        "mma_flops_per_algorithmic_flop": 3 if args.mode == "bf16x3" else 1,
                        "avg_launch_ms": k1_ms, "launches": int(prof_n[0]), "waveforms_per_launch": avg_wf,
                        "share_of_step": prof_ms[0] / ms,
                        "k2_head": {"avg_launch_ms": prof_ms[1] / max(prof_n[1], 1), "launches": int(prof_n[1]),
                                    "share_of_step": prof_ms[1] / ms}}
        cpu = cpu_baseline(args.cpu_sample) if (args.cpu_sample > 0 and world == 1) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
                "config": workload_config(args, world),
                "tflops_per_gpu": value / world * (args.t_star * WAVENET_GFLOP * args.length / 16000 +
                                                   (RESNEXT_GFLOP + 0.27 if args.workload in ("sc09", "sde") else 0.02)) / 1e3,
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4),
                        "d2h_bytes_per_step": int(pred_host.numel() * 8), "ms_per_step": ms_e2e / args.steps},
                "roofline": roofline, "cpu_baseline": cpu, "certification": cert, "configs": configs or None,
                "kernels": kernels, "eager_gpu": eager,
                "parity_unpinned": "torchsde==0.2.5 Euler-Maruyama stepping of the configs[3] legs (restated; the reference's own "
                                   "RevDiffWave / f / g are pinned, tests/golden/make_golden_sde.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        h.dist.barrier()
        h.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
