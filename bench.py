#!/usr/bin/env python
"""Benchmark of the tfkaldi hot path (BASELINE.json): training frames/sec of
CrossEnthropyTrainer.update on the C2 network (440 -> 6x2048 -> 1936 pdf-ids, 8192 frames per GPU,
bf16 operands / fp32 accumulate + fp32 master weights, Adam), data-parallel over N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4] [--precision bf16|bf16x3]

One "step" = one full optimizer step: forward (training mode), summed softmax-CE, backward, gradient
all-reduce over ranks, mean -> clip -> Adam (reference: neuralNetworks/trainer.py:260-354).
`--impl reference` times the reference's own CPU implementation of the same step: the reference is
Python-2 / TensorFlow-0.1x and cannot run here, so it is the CPU restatement in
oracle/ (fp32, torch-CPU GEMMs on every host core) — the one other place allowed to execute oracle/.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1] / [2]: 6x2048 DNN, 440-in, 1936 tri pdf-ids, batch 8192 per GPU
    "c2": dict(num_layers=6, input_dim=440, hidden_dim=2048, output_dim=1936, frames=8192, batch_norm=False, keep_prob=1.0),
    # BASELINE.json configs[3]: + Batchnorm + Dropout 0.5, 3401 lda_mllt pdf-ids, batch 4096
    "c4": dict(num_layers=6, input_dim=440, hidden_dim=2048, output_dim=3401, frames=4096, batch_norm=True, keep_prob=0.5),
}


def flops_per_frame(c):
    """3*2*sum(K_l*N_l) - 2*I*H (fwd + dgrad + wgrad, no dgrad for layer 0)"""
    dims = [c["input_dim"]] + [c["hidden_dim"]] * c["num_layers"] + [c["output_dim"]]
    fwd = 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    return 3 * fwd - 2 * c["input_dim"] * c["hidden_dim"], fwd


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops_sustained": p.get("bf16_tflops_sustained"), "tflops_burst": p.get("bf16_tflops"), "hbm_gbs": p.get("hbm_gbs"), "source": "measured"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.lines, self.proc, self.idx = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark(self):
        """host time stamp: samples between two marks are the ones taken during a timed region"""
        return time.monotonic()

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi's start-up (NVML initialisation on every GPU of the box) perturbs running work for tens of
        milliseconds: it must be over before anything is timed"""
        t0 = time.monotonic()
        while self.proc is not None and not self.lines and time.monotonic() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, regions=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        lines = self.lines
        if regions:  # a sample taken up to one period after a region's end still describes it
            inside = [l for t, l in lines if any(a <= t <= b + 0.03 for a, b in regions)]
            lines = [(0, l) for l in inside] if inside else lines
        for _, line in lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


def make_trainer(c, precision, distributed, seed=1234):
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer

    chain = act.Batchnorm(None) if c["batch_norm"] else None
    chain = act.TfActivation(chain, act.relu)
    if c["keep_prob"] < 1:
        chain = act.Dropout(chain, c["keep_prob"])
    dnn = DNN(c["output_dim"], c["num_layers"], c["hidden_dim"], chain, False)
    tr = CrossEnthropyTrainer(dnn, c["input_dim"], c["frames"], c["frames"], 1e-3, 1.0, 1000000, 1,
                              precision=precision, seed=seed, max_frames=c["frames"], distributed=distributed)
    tr.initialize()
    # output layer N(0, 1/sqrt(H)) for timing so the softmax is non-degenerate
    rng = np.random.default_rng(seed + 1)
    from tfkaldi_b200 import _lib as L

    tr.engine.set_tensor(L.T_WEIGHTS, c["num_layers"], (rng.standard_normal((c["hidden_dim"], c["output_dim"])) / math.sqrt(c["hidden_dim"])).astype(np.float32))
    return tr


def bench_config(args, world):
    """the `config` object both arms print (identical for `--impl ours` and `--impl reference`)"""
    c = CONFIGS[args.config]
    B = c["frames"]
    return {"workload": "C2: 440-6x2048-1936 DNN, ReLU, softmax-CE, Adam, %d frames/GPU/step" % B if args.config == "c2"
            else "C4: 440-6x2048-3401 DNN, BN+ReLU+dropout(keep 0.5), softmax-CE, Adam, %d frames/GPU/step" % B,
            "frames_per_gpu": B, "global_frames": B * world, "parallelism": "dp%d" % world}


NUM_WINDOWS = 5


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # clocks are sampled from BEFORE the warm-up (nvidia-smi's own start-up disturbs the GPUs and, on rank 0 only,
    # the host thread: inside the timed region every other rank would wait for it in the first collective)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    c = CONFIGS[args.config]
    B, I, O = c["frames"], c["input_dim"], c["output_dim"]
    # synthetic data of the reference's shape: post-CMVN spliced frames ~ N(0,1), uniform pdf labels;
    # a pool of distinct batches, per-rank data seed, identical weights on every rank
    pool = 4
    g = torch.Generator().manual_seed(1000 + rank)
    host_x = [torch.randn((B, I), generator=g, dtype=torch.float32).pin_memory() for _ in range(pool)]
    host_y = [torch.randint(0, O, (B,), generator=g, dtype=torch.int32).pin_memory() for _ in range(pool)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]
    lr = 1e-3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    regions = []

    def measure(tr, windows, per_step_events=False):
        """device-resident and end-to-end timing of one trainer: `windows` windows of exactly args.steps steps each,
        every window bracketed by barrier + synchronize and timed with CUDA events on the launching stream, max over
        ranks per window; the reported number is the MEDIAN window (all windows are printed)."""
        eng = tr.engine

        def step_resident(i):
            # tfk_train_step == tfk_accumulate + tfk_apply (tests/test_gpu_parity.py checks bit-equality); the loss
            # is copied to pinned memory asynchronously and not waited for
            eng.train_step(dev_x[i % pool], dev_y[i % pool], lr, want_loss=False)

        for i in range(args.warmup):
            step_resident(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res, launches = [], 0
        for w in range(windows):
            barrier()
            t_a = sampler.mark()
            launches0 = eng.kernel_launches()
            e0.record()
            for i in range(args.steps):
                step_resident(i)
            e1.record()
            barrier()
            regions.append((t_a, sampler.mark()))
            res.append(max_over_ranks(e0.elapsed_time(e1)) / args.steps)
            launches = (eng.kernel_launches() - launches0) // args.steps
        per_step = None
        if per_step_events:  # one more window with an event after every step: exposes a stall inside a window
            barrier()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
            evs[0].record()
            for i in range(args.steps):
                step_resident(i)
                evs[i + 1].record()
            barrier()
            ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
            per_step = {"min": min(ms), "median": float(np.median(ms)), "max": max(ms)}
        # ---- end to end: host buffers in, loss out, through the Trainer API
        for i in range(3):
            tr.update_packed(host_x[i % pool], host_y[i % pool])
        e2e, last_loss = [], None
        for w in range(min(windows, 3)):
            barrier()
            t_a = sampler.mark()
            e0.record()
            tr.prefetch(host_x[0], host_y[0])
            for i in range(args.steps):
                # update_packed = wait for this batch's H2D, full step, loss D2H + host sync.  The NEXT batch's
                # H2D is queued on the copy stream so it overlaps the step (north_star: "overlapped with the
                # next batch's H2D copy")
                nxt = (host_x[(i + 1) % pool], host_y[(i + 1) % pool]) if i + 1 < args.steps else None
                last_loss = tr.update_packed(host_x[i % pool], host_y[i % pool], prefetch=nxt)
            e1.record()
            barrier()
            regions.append((t_a, sampler.mark()))
            e2e.append(max_over_ranks(e0.elapsed_time(e1)) / args.steps)
        return {"ms": float(np.median(res)), "windows_ms": [round(v, 4) for v in res], "launches": int(launches),
                "e2e_ms": float(np.median(e2e)), "e2e_windows_ms": [round(v, 4) for v in e2e], "last_loss": last_loss,
                "per_step_ms": per_step}

    tr = make_trainer(c, args.precision, world > 1)
    eng = tr.engine
    m = measure(tr, NUM_WINDOWS, per_step_events=True)
    ms_resident, ms_e2e, launches, last_loss = m["ms"], m["e2e_ms"], m["launches"], m["last_loss"]
    timed_seconds = (sum(m["windows_ms"]) + sum(m["e2e_windows_ms"])) * args.steps * 1e-3

    # ---------------- per-kernel CUDA-event timing over the same steps (roofline numerator)
    eng.enable_timers(True)
    for i in range(args.steps):
        eng.train_step(dev_x[i % pool], dev_y[i % pool], lr, want_loss=False)
    timers = eng.timers()
    eng.enable_timers(False)
    clocks = sampler.stop(regions) if rank == 0 else None
    train_fl, fwd_fl = flops_per_frame(c)
    gemm_ms = (timers["gemm_fwd"][0] + timers["gemm_bwd"][0]) / args.steps
    gemm_launches = (timers["gemm_fwd"][1] + timers["gemm_bwd"][1]) // args.steps
    peaks = measured_peaks()
    achieved = train_fl * B / (gemm_ms * 1e-3) / 1e12
    breakdown = {k: round(v[0] / args.steps * 1e3, 1) for k, v in timers.items() if v[1]}  # us per step
    # which measured cuBLAS figure is the fair denominator: the BURST one when the timed stretch is short and the SM
    # clock stayed at its maximum (no power cap), the SUSTAINED one for seconds-long runs under the cap
    sm, sm_max = (clocks or {}).get("sm_mhz"), (clocks or {}).get("sm_max_mhz")
    capped = "sw_power_cap" in ((clocks or {}).get("reasons") or []) or (sm and sm_max and sm < 0.95 * sm_max)
    use_burst = (timed_seconds < 2.0) and not capped
    peak = peaks["tflops_burst"] if use_burst else peaks["tflops_sustained"]

    # ---------------- the fp32-equivalent (parity) mode on the same line
    parity = None
    if args.precision == "bf16" and not args.no_parity_mode:
        del tr, eng
        torch.cuda.empty_cache()
        tr3 = make_trainer(c, "bf16x3", world > 1)
        m3 = measure(tr3, 3)
        parity = {"dtype": "bf16x3 (every operand as bf16 hi+lo, 3 tensor-core passes per product: fp32-equivalent; the mode every <=1e-3 parity test runs in)",
                  "value": B * world / (m3["ms"] * 1e-3), "unit": "frames/s", "ms_per_step": m3["ms"], "windows_ms_per_step": m3["windows_ms"],
                  "e2e": {"value": B * world / (m3["e2e_ms"] * 1e-3), "unit": "frames/s", "ms_per_step": m3["e2e_ms"], "last_loss": m3["last_loss"]},
                  "gpu_launches_per_step": m3["launches"]}
        del tr3

    # HBM-bound kernels: algorithmic bytes / in-situ CUDA-event time (SURVEY.md 8d byte counts, adjusted to
    # what this engine actually stores: bf16 (or bf16 hi+lo) gradients / shadows)
    nparams = sum(a * b for a, b in zip([c["input_dim"]] + [c["hidden_dim"]] * c["num_layers"], [c["hidden_dim"]] * c["num_layers"] + [O]))
    sh = 2 if args.precision == "bf16" else 4
    hbm_bytes = {
        "adam": (28 + sh) * nparams / max(world, 1) if world > 1 else (28 + sh) * nparams,
        "softmax_ce": B * O * (4 + sh),
        "bn": (c["num_layers"] * B * c["hidden_dim"] * (2 * sh + 3 * sh)) if c["batch_norm"] else 0,  # fwd: read z, write y; bwd: read dy + z, write dz
    }
    hbm = {}
    for k, nbytes in hbm_bytes.items():
        if nbytes and timers[k][1]:
            us = timers[k][0] / args.steps * 1e3
            hbm[k] = {"bytes_per_step": int(nbytes), "us_per_step": round(us, 1), "gbs": round(nbytes / us / 1e3, 1),
                      "frac_of_measured_hbm_peak": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3)}

    out = None
    if rank == 0:
        frames_total = B * world
        out = {
            "metric": "training frames/sec (spliced-fbank->pdf CE), full optimizer step",
            "value": frames_total / (ms_resident * 1e-3),
            "unit": "frames/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_resident,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3(fp32-equivalent)",
            "data": "synthetic",
            "config": bench_config(args, world),
            "timing": {"windows": NUM_WINDOWS, "steps_per_window": args.steps, "value_is": "median window",
                       "windows_ms_per_step": m["windows_ms"], "per_step_ms_in_an_extra_window": m["per_step_ms"],
                       "timed_seconds": round(timed_seconds, 3),
                       "l2": "per-step working set (weights+Adam state+grads+activations ~0.9 GB) exceeds the 126 MB L2; %d rotating input batches; no explicit flush" % pool},
            "e2e": {"value": frames_total / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e, "windows_ms_per_step": m["e2e_windows_ms"],
                    "h2d_bytes_per_step": B * I * 4 + B * 4, "d2h_bytes_per_step": 16,
                    "api": "CrossEnthropyTrainer.update_packed(pinned x, pinned labels) -> loss", "last_loss": last_loss},
            "parity_mode": parity,
            "gpu_launches": int(launches * args.steps),
            "gpu_launches_per_step": int(launches),
            "roofline": {"bound": "tensor", "kernel": "tfk_gemm2_kernel (cta_group::2 CTA pairs; fused FFLayer fwd + fused wgrad/dgrad, %d launches/step)" % gemm_launches,
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_kind": "burst" if use_burst else "sustained",
                         "frac_of_burst_peak": achieved / peaks["tflops_burst"], "frac_of_sustained_peak": achieved / peaks["tflops_sustained"],
                         "peak_source": peaks["source"] + " (cuBLAS bf16, MEASURED_PEAKS.json: burst for a sub-2-s timed stretch at max SM clock, else sustained)",
                         "algorithmic_flops_per_step": train_fl * B, "kernel_ms_per_step": gemm_ms, **_ncu_traffic(args, B),
                         "step_share": gemm_ms / ms_resident, "per_step_us_by_kernel_class": breakdown},
            "hbm_kernels": hbm,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(c, steps=2)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def _ncu_traffic(args, frames):
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture of this exact
    workload (profiles/r2_ncu_traffic.json from the last build's capture, tools/ncu_traffic.py; else the round-1 file);
    null for workloads that were not captured."""
    t = None
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", name)
        try:
            t = json.load(open(path)).get("%s:%s:%d" % (args.config, args.precision, frames))
        except (OSError, ValueError):
            t = None
        if t:
            break
    if not t:
        return {"traffic": None}
    return {"traffic": t["dram_bytes_per_launch"], "traffic_unit": "bytes of DRAM read+write per launch (mean of the %d launches of a step)" % t["launches_per_step"],
            "traffic_source": t["source"]}


_THREADS = {}


def _best_thread_count(c):
    """the baseline gets the thread count that serves it best on this host (more is not always faster
    on many-core boxes): probe a few settings on one hidden-layer-sized GEMM pair and keep the fastest"""
    import torch

    key = os.cpu_count()
    if key in _THREADS:
        return _THREADS[key]
    n = os.cpu_count() or 1
    cands = sorted({n, max(1, n // 2), min(n, 64), min(n, 32), min(n, 16)}, reverse=True)
    a = torch.randn(4096, c["hidden_dim"])
    w = torch.randn(c["hidden_dim"], c["hidden_dim"])
    best, best_t = n, float("inf")
    for t in cands:
        torch.set_num_threads(t)
        torch.mm(a, w)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.mm(a, w)
            torch.mm(a.t(), a)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
    _THREADS[key] = best
    return best


def cpu_baseline(c, steps, frames=None, warmup=1):
    """the oracle restatement of the same step on the host cores (reported, not optimised against)"""
    import torch

    from oracle.dnn_oracle import OracleConfig, OracleDNN, reference_init, set_matmul_backend

    set_matmul_backend("torch")
    frames = frames or c["frames"]
    threads = _best_thread_count(c)
    torch.set_num_threads(threads)
    cfg = OracleConfig(c["num_layers"], c["input_dim"], c["hidden_dim"], c["output_dim"], batch_norm=c["batch_norm"], keep_prob=c["keep_prob"])
    rng = np.random.default_rng(1234)
    params = reference_init(cfg, rng)
    params["W%d" % c["num_layers"]] = (rng.standard_normal((c["hidden_dim"], c["output_dim"])) / math.sqrt(c["hidden_dim"])).astype(np.float32)
    orc = OracleDNN(cfg, params)
    x = rng.standard_normal((frames, c["input_dim"])).astype(np.float32)
    y = rng.integers(0, c["output_dim"], frames)
    for _ in range(warmup):
        orc.accumulate(x, y, dropout_seed=1)
        orc.apply(1e-3)
    t0 = time.perf_counter()
    for s in range(steps):
        orc.accumulate(x, y, dropout_seed=s)
        orc.apply(1e-3)
    dt = time.perf_counter() - t0
    return {"value": frames * steps / dt, "unit": "frames/s", "cores": threads, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": "%d full optimizer steps of %d frames after %d warm-up (CPU restatement of the reference path; TensorFlow 0.1x cannot run here)" % (steps, frames, warmup),
            "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's CPU implementation (oracle port), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    c = CONFIGS[args.config]
    # bound the run to ~2 minutes whatever K is: the first (warm-up) steps measure the per-step cost
    frames = c["frames"]
    budget_s = 120.0
    probe = cpu_baseline(c, steps=1, frames=frames, warmup=1)
    per_step = probe["seconds"]
    total_steps = args.steps + args.warmup
    if per_step * total_steps > budget_s:
        frames = max(256, int(frames * budget_s / (per_step * total_steps)) // 256 * 256)
    res = cpu_baseline(c, steps=args.steps, frames=frames, warmup=min(args.warmup, 3) if frames == c["frames"] else args.warmup)
    train_fl, _ = flops_per_frame(c)
    out = {
        "impl": "reference",
        "metric": "training frames/sec (spliced-fbank->pdf CE), full optimizer step",
        "value": res["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, args.gpus),
        "reference_note": "reference = vrenkens/tfkaldi CrossEnthropyTrainer.update semantics restated on CPU (oracle/dnn_oracle.py); the TF-0.1x/Python-2 "
                          "original cannot execute in this image; rank 0 alone runs it, %d frames per step" % frames,
        "cpu_baseline": {"value": res["value"], "unit": "frames/s", "cores": res["cores"], "kind": "port",
                         "sample": "%d steps of %d frames each (of the %d-frame workload)" % (args.steps, frames, c["frames"])},
        "e2e": {"value": res["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "achieved_tflops": train_fl * res["value"] / 1e12,
    }
    print(json.dumps(out))


def run_decode(args):
    """BASELINE.json configs[4]: trained-shape 6x2048 net, forward-only log-likelihood emission of
    100k-frame utterances through Decoder + ArkWriter (reference: nnet.py:246-289)."""
    import tempfile

    import torch

    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.decoder import Decoder
    from tfkaldi_b200.processing import ark

    torch.cuda.set_device(0)
    c = CONFIGS["c2"]
    T, I, O = 100000, c["input_dim"], c["output_dim"]
    dnn = DNN(O, c["num_layers"], c["hidden_dim"], act.TfActivation(None, act.relu), False)
    dec = Decoder(dnn, I, T, precision=args.precision, max_frames=16384)
    rng = np.random.default_rng(7)
    dec.engine.load_params(dnn.initial_parameters(I, rng))
    from tfkaldi_b200 import _lib as L

    dec.engine.set_tensor(L.T_WEIGHTS, c["num_layers"], (rng.standard_normal((c["hidden_dim"], O)) / math.sqrt(c["hidden_dim"])).astype(np.float32))
    prior = torch.full((O,), 1.0 / O, dtype=torch.float32, device="cuda")
    x_host = torch.randn((T, I), dtype=torch.float32).pin_memory()
    x_dev = x_host.cuda()
    out_dev = torch.empty((T, O), dtype=torch.float32, device="cuda")
    steps, warm = max(3, args.steps // 20), 3
    for _ in range(warm):
        dec.loglik(x_dev, prior, out=out_dev)
    torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start()
    l0 = dec.engine.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        dec.loglik(x_dev, prior, out=out_dev)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (dec.engine.kernel_launches() - l0) // steps
    clocks = sampler.stop()
    # end to end, the way Nnet.decode runs it (neuralNetworks/nnet.py decode -> decoder.LoglikStreamer): RAW 40-dim
    # utterances in host memory -> H2D -> CMVN + splice + network + log(softmax/prior) on the device, tile by tile ->
    # D2H into pinned slots on a copy stream -> positional writes into the archive (tmpfs if available) by writer
    # threads, all overlapped; the archive bytes are the ones ArkWriter.write_next_utt produces
    from tfkaldi_b200.neuralNetworks.decoder import LoglikStreamer

    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    D, ctx = 40, 5
    raw_utts = [np.random.default_rng(20 + i).standard_normal((T, D)).astype(np.float32) for i in range(2)]
    stats = np.zeros((2, D + 1), np.float32)  # accumulated CMVN statistics of a zero-mean unit-variance speaker
    stats[0, -1] = 1000.0
    stats[1, :-1] = 1000.0
    n_utts, t_e2e = 4, []
    for rep in range(2):  # first pass warms the page cache / pinned allocations; the second is reported
        for f in ("feats.scp", "likelihoods.ark"):
            if os.path.exists(os.path.join(tmp, f)):
                os.remove(os.path.join(tmp, f))
        writer = ark.ArkWriter(os.path.join(tmp, "feats.scp"), os.path.join(tmp, "likelihoods.ark"))
        stream = LoglikStreamer(dec, writer, prior.cpu().numpy())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n_utts):
            stream.decode_raw("utt%d" % i, raw_utts[i % 2], stats, ctx)
        stream.close()
        writer.close()
        t_e2e.append((time.perf_counter() - t0) / n_utts)
    check = ark.ArkReader(os.path.join(tmp, "feats.scp"))
    got = check.read_utt("utt1")[:512]
    want = dec.engine.loglik_raw(raw_utts[1][:600], np.array([0, 600], np.int32), np.stack([np.zeros(D, np.float32), np.ones(D, np.float32)])[None],
                                 D, ctx, prior)[:512].cpu().numpy()
    e2e_check = float(np.abs(got - want).max())
    import shutil

    shutil.rmtree(tmp, ignore_errors=True)
    _, fwd_fl = flops_per_frame(c)
    dec.engine.enable_timers(True)
    for _ in range(steps):
        dec.loglik(x_dev, prior, out=out_dev)
    tm = dec.engine.timers()
    peaks = measured_peaks()
    gemm_ms = tm["gemm_fwd"][0] / steps
    print(json.dumps({
        "metric": "decoder frames/sec (forward-only log-likelihood emission)", "value": T / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "C5: 440-6x2048-1936 DNN, one 100000-frame utterance per step, log(softmax/prior) -> [T,1936] fp32", "frames": T},
        "e2e": {"value": T / t_e2e[-1], "unit": "frames/s", "seconds_per_utt": t_e2e[-1], "h2d_bytes_per_step": T * 40 * 4, "d2h_bytes_per_step": T * O * 4,
                "api": "LoglikStreamer.decode_raw (what Nnet.decode runs): raw utterance -> device CMVN/splice/network -> pinned D2H -> ArkWriter positional writes (774 MB archive entry per utterance), %d utterances back to back" % n_utts,
                "first_pass_seconds_per_utt": t_e2e[0], "archive_vs_direct_max_abs_diff": e2e_check},
        "gpu_launches": int(launches * steps), "gpu_launches_per_step": int(launches),
        "roofline": {"bound": "tensor", "kernel": "tfk_gemm2_kernel (forward)", "achieved": fwd_fl * T / (gemm_ms * 1e-3) / 1e12, "peak": peaks["tflops_sustained"],
                     "unit": "TFLOP/s", "frac": fwd_fl * T / (gemm_ms * 1e-3) / 1e12 / peaks["tflops_sustained"], "traffic": None,
                     "per_step_us_by_kernel_class": {k: round(v[0] / steps * 1e3, 1) for k, v in tm.items() if v[1]}},
        "hbm_kernels": {"decode_out": {"bytes_per_step": T * O * 8, "us_per_step": round(tm["decode_out"][0] / steps * 1e3, 1),
                                       "gbs": round(T * O * 8 / (tm["decode_out"][0] / steps * 1e3) / 1e3, 1)}},
        "clocks": clocks}))


def run_feed(args):
    """The data plane in front of the hot path: a synthetic Kaldi corpus (feats.ark/scp, per-speaker CMVN archives,
    gzip'ed pdf alignments) on tmpfs -> AlignmentBatchDispenser -> RawBatchFeeder thread (pinned buffers, copy stream)
    -> tfk_train_step_raw (device CMVN + splice + the whole C2 step), against the reference's own loop shape
    (dispenser.get_batch() then trainer.update(), neuralNetworks/nnet.py:157-160) on the same files."""
    import shutil
    import tempfile

    import torch

    from tfkaldi_b200 import synth
    from tfkaldi_b200.neuralNetworks.classifiers import activation as act
    from tfkaldi_b200.neuralNetworks.classifiers.dnn import DNN
    from tfkaldi_b200.neuralNetworks.trainer import CrossEnthropyTrainer
    from tfkaldi_b200.processing import batchdispenser, feature_reader, target_coder
    from tfkaldi_b200.processing.feeder import RawBatchFeeder

    torch.cuda.set_device(0)
    c = CONFIGS["c2"]
    size = 16  # utterances per step x ~512 frames = ~8192 frames
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        info = synth.make_corpus(tmp, num_utts=768, min_len=462, max_len=562, feat_dim=40, num_speakers=24, num_pdfs=c["output_dim"], seed=0)

        def dispenser():
            reader = feature_reader.FeatureReader(tmp + "/feats_shuffled.scp", tmp + "/cmvn.scp", tmp + "/utt2spk", 5, info["max_length"])
            return batchdispenser.AlignmentBatchDispenser(reader, target_coder.AlignmentCoder(lambda x, y: x, c["output_dim"]), size, info["alifile"])

        dnn = DNN(c["output_dim"], c["num_layers"], c["hidden_dim"], act.TfActivation(None, act.relu), False)
        tr = CrossEnthropyTrainer(dnn, c["input_dim"], info["max_length"], info["max_length"], 1e-3, 1.0, 1000000, size,
                                  precision=args.precision, seed=1234)
        tr.initialize()
        steps = min(args.steps, 100)
        feeder = RawBatchFeeder(dispenser(), size)
        # warm-up = one whole epoch: every batch length of the corpus has been seen once, so the timed region measures
        # the steady state of a long run (tensor maps and work lists of every launch shape are cached by then)
        warm = max(args.warmup, feeder.num_batches)
        for _ in range(warm):
            tr.update_prefetched(feeder)
        torch.cuda.synchronize()
        sampler = ClockSampler(0); sampler.start()
        f0, l0, t0 = feeder.frames_out, tr.engine.kernel_launches(), time.perf_counter()
        for _ in range(steps):
            loss = tr.update_prefetched(feeder)  # returns the step's loss: one host sync per step
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        frames, launches = feeder.frames_out - f0, tr.engine.kernel_launches() - l0
        clocks = sampler.stop()
        feeder.close()
        # host pipeline alone (no GPU work): how fast the thread can read + pack
        feeder = RawBatchFeeder(dispenser(), size)
        b = feeder.get(); feeder.release(b)
        th, nh = time.perf_counter(), 0
        for _ in range(200):
            b = feeder.get(); nh += b.frames; feeder.release(b)
        host_rate = nh / (time.perf_counter() - th)
        feeder.close()
        # the reference's loop shape on the same files: host CMVN + splice, then the step
        d = dispenser()
        for _ in range(2):
            tr.update(*d.get_batch())
        torch.cuda.synchronize()
        ts, ns = time.perf_counter(), 0
        for _ in range(max(5, steps // 10)):
            x, y = d.get_batch()
            ns += sum(m.shape[0] for m in x)
            tr.update(x, y)
        torch.cuda.synchronize()
        sync_rate = ns / (time.perf_counter() - ts)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps({
        "metric": "training frames/sec from Kaldi archives (ark/scp + CMVN + alignments -> full optimizer step)", "value": frames / dt,
        "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "C2 net fed from a synthetic Kaldi corpus on tmpfs: 768 utterances of 462..562 x 40-dim frames, 16 utterances (~8192 frames) per step",
                   "mean_frames_per_step": frames / steps},
        "e2e": {"value": frames / dt, "unit": "frames/s", "h2d_bytes_per_step": int(frames / steps * (40 * 4 + 4)), "d2h_bytes_per_step": 16,
                "api": "RawBatchFeeder(AlignmentBatchDispenser) -> CrossEnthropyTrainer.update_prefetched -> loss", "last_loss": loss},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps,
        "host_pipeline_frames_per_s": host_rate,
        "reference_loop_shape": {"value": sync_rate, "unit": "frames/s", "what": "dispenser.get_batch() (host CMVN + splice) then trainer.update(), same engine"},
        "clocks": clocks}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS) + ["c5", "feed"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-mode", action="store_true", help="skip the bf16x3 (fp32-equivalent) measurement on the same line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.config == "c5":
        run_decode(args)
    elif args.config == "feed":
        run_feed(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
