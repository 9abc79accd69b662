#!/usr/bin/env python
"""Benchmark of the CTC hot path on B200 (see DESIGN.md section 7 for the definitions).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload loss|beam|greedy]

Headline (default workload "loss") = BASELINE.json configs[1]: CTC loss fwd+bwd, batch 256 x T=1000
x V=48, labels U[50,200], one fused kernel launch per step.  One JSON line is printed by rank 0:

  value         frames/s over all ranks, inputs resident in HBM (rotating buffer sets > L2)
  e2e           the same metric through the reference-facing module (DistillationCTCLoss ->
                loss.backward()) with HOST inputs: pinned H2D copy of the step's tensors and a D2H
                read of the loss inside the timed region
  roofline      algorithmic bytes (384 B/frame) / measured kernel time vs the measured HBM peak
  cpu_baseline  the reference's own CPU implementation of the step (torch ATen ctc_loss via the
                reference's call sequence) timed on this box's host cores (rank 0, N=1 only)
  beam50        secondary figures for the decode half of BASELINE.json's metric (utterances/s)

`--impl reference` times the reference CPU path alone (rank 0; other ranks exit).
Multi-GPU: one process per GPU (torchrun); every rank runs the per-GPU workload (weak scaling),
the only collective is the 2-scalar all-reduce of the batch loss.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG2 = dict(B=256, T=1000, V=48, S=200, seed=1002)
CFG3 = dict(N=10000, T=500, V=48, seed=1003, width=50, prune=-10.0)
LOSS_BYTES_PER_FRAME = 2 * 48 * 4   # read log-probs once + write gradient once
DECODE_BYTES_PER_FRAME = 48 * 4


# ------------------------------------------------------------------ helpers
def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic(kernel: str, frames: int):
    """dram bytes per launch (read+write) from the committed ncu capture, scaled to this launch's
    frame count (profiles/traffic.json holds bytes per frame), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f).get(kernel)
        return None if rec is None else rec["bytes_per_frame"] * frames
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run."""

    FIELDS = ("clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        busy = [r for r in self.rows if len(r) >= 7 and r[2].isdigit() and int(r[2]) > 0] or [r for r in self.rows if len(r) >= 7]
        for r in busy:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower() == "active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cfg2_inputs():
    from silent_speech_decoder_b200 import synth

    c = CFG2
    lp = synth.flat(c["B"], c["T"], c["seed"], c["V"])
    g = torch.Generator().manual_seed(c["seed"])
    tl = torch.randint(c["S"] // 4, c["S"] + 1, (c["B"],), generator=g, dtype=torch.int64)   # U[50,200]
    tg = torch.randint(2, c["V"], (c["B"], c["S"]), generator=g, dtype=torch.int64)
    tg = tg * (torch.arange(c["S"])[None, :] < tl[:, None])
    il = torch.full((c["B"],), c["T"], dtype=torch.int64)
    return lp, il, tg, tl


def cfg3_inputs(n: int, kind: str):
    from silent_speech_decoder_b200 import synth

    c = CFG3
    if kind == "flat":
        return synth.flat(n, c["T"], c["seed"], c["V"]), torch.full((n,), c["T"], dtype=torch.int64)
    base_n = min(n, 500)
    base = synth.peaked(base_n, c["T"], c["seed"], c["V"])
    reps = (n + base_n - 1) // base_n
    return base.repeat(reps, 1, 1)[:n].contiguous(), torch.full((n,), c["T"], dtype=torch.int64)


def dist_max(x: float, world: int, dev) -> float:
    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


# ------------------------------------------------------------------ reference arm (CPU)
def cpu_loss_reference(steps: int, warmup: int):
    """The reference's CTC step on the host: losses.py:52-56 call sequence -> ATen ctc_loss_cpu + backward."""
    from oracle import ctc_loss as ol

    lp, il, tg, tl = cfg2_inputs()
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(max(warmup, 1)):
        ol.reference_module_ctc(lp, il, tg, tl, 1)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ol.reference_module_ctc(lp, il, tg, tl, 1)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return CFG2["B"] * CFG2["T"] / (ms * 1e-3), ms, torch.get_num_threads()


def _beam_worker(args):
    from oracle import beam as ob

    return ob.decode_one(args)


def cpu_beam_reference(n_utts: int, kind: str, budget_s: float = 25.0):
    """oracle/beam.py (restatement of pyctcdecode; the library itself is unavailable) over all host
    cores with multiprocessing, on a bounded sample of cfg3."""
    import multiprocessing as mp

    from oracle import beam as ob
    from silent_speech_decoder_b200.data.vocab import Vocab

    v = Vocab.default()
    dec = ob.BeamDecoder(v.tokens, v.blank_id, v.pad_id, CFG3["width"], CFG3["prune"])
    lp, ln = cfg3_inputs(n_utts, kind)
    mat = ob.wrapper_matrix(lp.numpy(), v.blank_id, v.pad_id, 0.0, dec.keep)
    jobs = [(mat[i, : int(ln[i])], dec.labels, CFG3["width"], CFG3["prune"], -5.0) for i in range(n_utts)]
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    ob.decode_one(jobs[0])
    one = time.perf_counter() - t0
    n = max(cores, min(n_utts, int(budget_s * cores / max(one, 1e-3))))
    n = min(n, n_utts)
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(_beam_worker, jobs[:n], chunksize=1)
        dt = time.perf_counter() - t0
    return n / dt, cores, n, one


def run_reference(args, rank):
    if rank != 0:
        return
    if args.workload == "loss":
        val, ms, cores = cpu_loss_reference(args.steps, args.warmup)
        line = {"metric": "ctc_loss_fwd_bwd_frames_per_sec", "value": val, "unit": "frames/s", "ms_per_step": ms,
                "sample": "full config 2 batch per step", "cores": cores, "kind": "reference",
                "workload": "cfg2: CTC loss fwd+bwd B=256 T=1000 V=48 S~U[50,200] per GPU (configs[1])"}
    else:
        val, cores, n, one = cpu_beam_reference(256, args.beam_input)
        line = {"metric": "ctc_beam50_decode_utterances_per_sec", "value": val, "unit": "utterances/s", "ms_per_step": 1e3 * n / val,
                "sample": f"{n} utterances of cfg3 ({args.beam_input}) over {cores} processes", "cores": cores, "kind": "port",
                "workload": f"cfg3: prefix beam search width 50, T=500, {args.beam_input} posteriors"}
    out = {
        "impl": "reference", "metric": line["metric"], "value": line["value"], "unit": line["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": line["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.workload == "loss" else "f64", "data": "synthetic",
        "config": {"workload": line["workload"]},
        "cpu_baseline": {"value": line["value"], "unit": line["unit"], "cores": line["cores"], "kind": line["kind"], "sample": line["sample"]},
        "e2e": {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ native arm
def bench_loss(args, rank, world, dev):
    from silent_speech_decoder_b200 import dist as sdist
    from silent_speech_decoder_b200 import ops
    from silent_speech_decoder_b200.models.losses import DistillationCTCLoss, LossWeights

    c = CFG2
    lp, il, tg, tl = cfg2_inputs()
    B, T, V = lp.shape
    nsets = 4  # 4 x (49 MB in + 49 MB out) = 393 MB > 126 MB L2
    lps = [lp.to(dev).clone() for _ in range(nsets)]
    il_d, tg_d, tl_d = il.to(dev), tg.to(dev), tl.to(dev)
    global_b = B * world

    def step(i):
        # gradients are scaled by the global batch in-kernel; the 2-scalar all-reduce of the loss value runs on
        # NCCL's stream and is waited for where the value is read (end of the timed region)
        nll, grad = ops.ctc_loss_fwd_bwd(lps[i % nsets], il_d, tg_d, tl_d, blank=1, zero_infinity=True, mean_denominator=global_b)
        return sdist.global_mean_loss_async(nll, tl_d)

    for i in range(args.warmup):
        step(i)()
    barrier(world)
    l0 = ops.LAUNCHES["loss"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pending = [step(i) for i in range(args.steps)]
    for wait in pending:
        wait()
    e1.record()
    barrier(world)
    ms_total = dist_max(e0.elapsed_time(e1), world, dev)
    launches = ops.LAUNCHES["loss"] - l0
    ms_step = ms_total / args.steps

    # kernel-only duration for the roofline (events bracket each launch; no other work between)
    kt = []
    for i in range(min(args.steps, 50)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.ctc_loss_fwd_bwd(lps[i % nsets], il_d, tg_d, tl_d, blank=1, zero_infinity=True, mean_denominator=global_b)
        b.record()
        b.synchronize()
        kt.append(a.elapsed_time(b))
    k_ms = statistics.median(kt)

    # end to end through the reference-facing module, host buffers
    fn = DistillationCTCLoss(vocab_size=V, blank_id=1, weights=LossWeights(0.0, 1.0))
    h_lp, h_il, h_tg, h_tl = lp.pin_memory(), il.pin_memory(), tg.pin_memory(), tl.pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (h_lp, h_il, h_tg, h_tl))

    def e2e_step():
        x = h_lp.to(dev, non_blocking=True).requires_grad_(True)
        out = fn(x, h_il.to(dev, non_blocking=True), h_tg.to(dev, non_blocking=True), h_tl.to(dev, non_blocking=True), x, None)
        out["total"].backward()
        return float(out["ctc"].item())  # D2H read of the step's result

    e2e_steps = max(3, min(args.steps, 30))
    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = dist_max(1e3 * (time.perf_counter() - t0) / e2e_steps, world, dev)

    peak, peak_src = measured_peak_gbs()
    achieved = LOSS_BYTES_PER_FRAME * B * T / (k_ms * 1e-3) / 1e9
    # the entry point runs the fp32 fast-path kernel, then the fp64 kernel over whatever it flagged
    n_fallback = ops.loss_fallback_count(B, T, V, tg_d.shape[1], dev)
    fast = n_fallback is not None
    line = {
        "metric": "ctc_loss_fwd_bwd_frames_per_sec", "value": world * B * T / (ms_step * 1e-3), "unit": "frames/s",
        "ms_per_step": ms_step, "dtype": "f32" if fast else "f64",
        "config": {"workload": "cfg2: CTC loss fwd+bwd B=256 T=1000 V=48 S~U[50,200] per GPU (configs[1])",
                   "l2": f"{nsets} rotating input/gradient sets = {nsets * 2 * B * T * V * 4 / 1e6:.0f} MB > 126 MB L2",
                   "parallelism": f"dp{world} by utterance, 2-scalar all-reduce of the loss"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic("loss_fast_kernel" if fast else "loss_kernel", B * T), "peak_source": peak_src,
                     "kernel": "ssd::loss_fast_kernel<14> (+ ssd::loss_kernel over flagged utterances)" if fast else "ssd::loss_kernel",
                     "kernel_ms": k_ms, "algorithmic_bytes": LOSS_BYTES_PER_FRAME * B * T,
                     "exact_path_utterances": n_fallback if fast else B,
                     "note": "kernel_ms spans both launches of one ssd_ctc_loss call; the loss is bound by its serial "
                             "recursion and instruction issue, not HBM (DESIGN.md section 4)"},
        "e2e": {"value": world * B * T / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "api": "DistillationCTCLoss(...)(log_probs, ...)['total'].backward()"},
        "gpu_launches": launches,
    }
    return line


def bench_decode(args, rank, world, dev, method: str):
    from silent_speech_decoder_b200 import dist as sdist
    from silent_speech_decoder_b200 import ops
    from silent_speech_decoder_b200.data.vocab import Vocab
    from silent_speech_decoder_b200.decoding import ctc

    v = Vocab.default()
    c = CFG3
    n_local = c["N"] if args.scaling == "weak" else c["N"] // world
    lp, ln = cfg3_inputs(n_local, args.beam_input)
    x, ln_d = lp.to(dev), ln.to(dev)
    _, space = ctc._beam_label_strings(v)

    def kernel():
        if method == "beam":
            return ops.beam_ids(x, ln_d, v.blank_id, v.pad_id, space, c["width"], c["prune"])
        return ops.greedy_ids(x, ln_d, v.blank_id)

    steps = max(1, min(args.steps, 5 if method == "beam" else 50))
    for _ in range(max(1, min(args.warmup, 3))):
        kernel()
    barrier(world)
    key = "beam" if method == "beam" else "greedy"
    l0 = ops.LAUNCHES[key]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        kernel()
    e1.record()
    barrier(world)
    ms = dist_max(e0.elapsed_time(e1), world, dev) / steps
    launches = ops.LAUNCHES[key] - l0

    decoder = ctc.build_decoder(method, v, beam_width=c["width"], beam_prune_logp=c["prune"])
    h_lp, h_ln = lp.pin_memory(), ln.pin_memory()

    def e2e_step():
        return decoder(h_lp.to(dev, non_blocking=True), h_ln)  # H2D of the posteriors, D2H of ids, strings built

    e2e_step()
    barrier(world)
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        texts = e2e_step()
    e2e_ms = dist_max(1e3 * (time.perf_counter() - t0) / reps, world, dev)
    d2h = n_local * 4 + sum(len(t) for t in texts) * 4

    peak, peak_src = measured_peak_gbs()
    achieved = DECODE_BYTES_PER_FRAME * n_local * c["T"] / (ms * 1e-3) / 1e9
    kname = "beam_kernel" if method == "beam" else "greedy_kernel"
    return {
        "metric": f"ctc_{'beam50' if method == 'beam' else 'greedy'}_decode_utterances_per_sec",
        "value": world * n_local / (ms * 1e-3), "unit": "utterances/s", "ms_per_step": ms, "dtype": "f64" if method == "beam" else "f32",
        "config": {"workload": f"cfg3: {method} decode of {n_local} utterances per GPU, T=500, V=48, {args.beam_input} posteriors"
                               + (", width 50, prune -10" if method == "beam" else ""),
                   "l2": f"input {n_local * c['T'] * 48 * 4 / 1e6:.0f} MB > 126 MB L2",
                   "parallelism": f"dp{world} by utterance, no data-path collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(kname + (("_" + args.beam_input) if method == "beam" else ""), n_local * c["T"]), "peak_source": peak_src, "kernel": f"ssd::{kname}", "kernel_ms": ms,
                     "algorithmic_bytes": DECODE_BYTES_PER_FRAME * n_local * c["T"],
                     "note": "beam search is latency/issue bound, not HBM bound; see profiles/ for stall breakdown" if method == "beam" else ""},
        "e2e": {"value": world * n_local / (e2e_ms * 1e-3), "unit": "utterances/s", "h2d_bytes_per_step": h_lp.numel() * 4 + h_ln.numel() * 8,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "api": f"build_decoder('{method}', vocab)(log_probs, lengths)"},
        "gpu_launches": launches,
    }


def run_native(args, rank, world, local):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native path has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    with ClockSampler(local) as clk:
        if args.workload == "loss":
            line = bench_loss(args, rank, world, dev)
        else:
            line = bench_decode(args, rank, world, dev, args.workload)
        secondary = {}
        if args.workload == "loss" and not args.no_secondary:
            for kind in ("peaked", "flat"):
                sub = argparse.Namespace(**vars(args))
                sub.beam_input, sub.workload = kind, "beam"
                r = bench_decode(sub, rank, world, dev, "beam")
                secondary[f"beam50_{kind}"] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "roofline", "e2e", "config", "gpu_launches")}
        time.sleep(0.12)
    if rank != 0:
        return
    if world == 1 and not args.no_cpu:
        if args.workload == "loss":
            val, ms, cores = cpu_loss_reference(3, 1)
            line["cpu_baseline"] = {"value": val, "unit": "frames/s", "cores": cores, "kind": "reference",
                                    "sample": "3 full config-2 steps (the reference's call sequence into ATen ctc_loss_cpu, fp32)", "ms_per_step": ms}
            if secondary:
                for kind in ("peaked", "flat"):
                    v, cores, n, one = cpu_beam_reference(256, kind, budget_s=12.0)
                    secondary[f"beam50_{kind}"]["cpu_baseline"] = {
                        "value": v, "unit": "utterances/s", "cores": cores, "kind": "port",
                        "sample": f"{n} cfg3 utterances ({kind}) over {cores} processes; oracle/beam.py restatement of pyctcdecode (library unavailable)",
                        "single_process_utt_per_s": 1.0 / one}
        else:
            v, cores, n, one = cpu_beam_reference(256, args.beam_input) if args.workload == "beam" else (None, None, None, None)
            if v:
                line["cpu_baseline"] = {"value": v, "unit": "utterances/s", "cores": cores, "kind": "port",
                                        "sample": f"{n} cfg3 utterances over {cores} processes; oracle/beam.py restatement of pyctcdecode"}
    out = {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": line["dtype"], "data": "synthetic", "config": line["config"], "clocks": clk.summary(),
           "e2e": line["e2e"], "gpu_launches": line["gpu_launches"], "roofline": line["roofline"]}
    if "cpu_baseline" in line:
        out["cpu_baseline"] = line["cpu_baseline"]
    out.update(secondary)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=["loss", "beam", "greedy"], default="loss")
    ap.add_argument("--beam-input", choices=["peaked", "flat"], default="peaked")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    from silent_speech_decoder_b200 import dist as sdist

    rank, world, local = sdist.init_from_env()
    if world > 1:
        sdist.bind_host_thread_to_gpu(local)  # pinned e2e buffers on the GPU's own NUMA node (N=1 keeps all cores for cpu_baseline)
    run_native(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
