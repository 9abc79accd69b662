#!/usr/bin/env python
"""Benchmark of the CTC hot path on B200 (see DESIGN.md section 7 for the definitions).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--workload loss|beam|greedy|eval] [--T 500 --width 50 --utts 10000 --beam-input peaked|flat]

Headline (default workload "loss") = BASELINE.json configs[1]: CTC loss fwd+bwd, batch 256 x T=1000
x V=48, labels U[50,200], one fused kernel launch per step.  One JSON line is printed by rank 0:

  value         frames/s over all ranks, inputs resident in HBM (rotating buffer sets > L2)
  e2e           the same metric through the reference-facing module (DistillationCTCLoss ->
                loss.backward()) with HOST inputs: every step's pinned H2D copy (prefetched one step ahead on
                a copy stream, data/prefetch.py) and a D2H read of the loss are inside the timed region
  roofline      algorithmic bytes (384 B/frame) / the loss launch's average duration over a back-to-back run of
                ssd_ctc_loss calls (CUDA events on the launching stream) vs the measured HBM peak
  cpu_baseline  the reference's CPU step (a port of its call sequence, losses.py:52-56, into the same ATen
                ctc_loss_cpu op it calls) timed on this box's host cores (rank 0, N=1 only)
  gpu_library_baseline   torch's own CUDA ctc_loss fwd+bwd on the same inputs and box (SURVEY 2.1's "kernel to beat")
  secondary objects      loss_cfg2_trained (confident posteriors on their own transcripts, English label
                frequencies), beam50_peaked / beam50_flat / greedy (config 3), cfg4_width_sweep (T=4000, widths
                1/10/50/200), cfg5_eval (PyTorch encoder forward + native beam-50 + WER/CER over the synthetic
                test set, sharded by utterance over the ranks)

`--impl reference` times the reference CPU path alone (rank 0; other ranks exit).
Multi-GPU: one process per GPU (torchrun); every rank runs the per-GPU workload (weak scaling; config 5's fixed
test set is split over the ranks), the only data-path collective is the 2-scalar all-reduce of the batch loss.
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
CFG4 = dict(N=256, T=4000, V=48, seed=1004, widths=(1, 10, 50, 200), prune=-10.0)
CFG5 = dict(N=159 * 64, seed=1005, t_lo=200, t_hi=1200, feat=640, batch=96)
LOSS_BYTES_PER_FRAME = 2 * 48 * 4   # read log-probs once + write gradient once
DECODE_BYTES_PER_FRAME = 48 * 4
LOSS_WORKLOAD = "cfg2: CTC loss fwd+bwd B=256 T=1000 V=48 S~U[50,200] per GPU (configs[1])"


def loss_config(world: int) -> dict:
    """The SAME dictionary in the native and the reference arm."""
    c = CFG2
    nsets = 4
    return {"workload": LOSS_WORKLOAD,
            "l2": f"{nsets} rotating input/gradient sets = {nsets * 2 * c['B'] * c['T'] * c['V'] * 4 / 1e6:.0f} MB > 126 MB L2",
            "parallelism": f"dp{world} by utterance, 2-scalar all-reduce of the loss"}


def decode_config(method: str, n: int, T: int, kind: str, width: int, world: int) -> dict:
    return {"workload": f"{'cfg3' if T == CFG3['T'] else 'cfg4'}: {method} decode of {n} utterances per GPU, T={T}, V=48, {kind} posteriors"
                        + (f", width {width}, prune -10" if method == "beam" else ""),
            "l2": f"input {n * T * 48 * 4 / 1e6:.0f} MB > 126 MB L2",
            "parallelism": f"dp{world} by utterance, no data-path collective"}


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


def cfg2_trained_inputs():
    """Config 2's shape in the regime a training run lives in: confident posteriors laid on each utterance's OWN
    transcript, label counts U[50,200], English label frequencies (space ~0.15, e ~0.10, ...)."""
    from silent_speech_decoder_b200 import synth
    from silent_speech_decoder_b200.data.vocab import Vocab

    c = CFG2
    w = synth.english_label_weights(Vocab.default().tokens)
    lp, tg, tl = synth.peaked_aligned(c["B"], c["T"], c["seed"] + 7, c["V"], n_labels=(c["S"] // 4, c["S"]), label_weights=w)
    if tg.size(1) < c["S"]:
        tg = torch.nn.functional.pad(tg, (0, c["S"] - tg.size(1)))
    return lp, torch.full((c["B"],), c["T"], dtype=torch.int64), tg, tl


def decode_inputs(n: int, T: int, kind: str, seed: int):
    """n DISTINCT utterances (peaked: each with its own transcript and alignment)."""
    from silent_speech_decoder_b200 import synth

    if kind == "flat":
        return synth.flat(n, T, seed, 48), torch.full((n,), T, dtype=torch.int64)
    return synth.peaked_aligned(n, T, seed, 48)[0], torch.full((n,), T, dtype=torch.int64)


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


CPU_LOSS_KIND = "port"
CPU_LOSS_SAMPLE = ("full config-2 steps; port of the reference's call sequence (losses.py:52-56: transpose, lengths to host, "
                   "nn.CTCLoss(blank, zero_infinity=True), backward) into the same ATen ctc_loss_cpu op the reference calls, fp32")


def _beam_worker(args):
    from oracle import beam as ob

    return ob.decode_one(args)


def cpu_beam_reference(n_utts: int, kind: str, budget_s: float = 25.0, T: int = CFG3["T"], width: int = CFG3["width"], seed: int = CFG3["seed"]):
    """oracle/beam.py (restatement of pyctcdecode; the library itself is unavailable) over all host
    cores with multiprocessing, on a bounded sample of the workload."""
    import multiprocessing as mp

    from oracle import beam as ob
    from silent_speech_decoder_b200.data.vocab import Vocab

    v = Vocab.default()
    dec = ob.BeamDecoder(v.tokens, v.blank_id, v.pad_id, width, CFG3["prune"])
    lp, ln = decode_inputs(n_utts, T, kind, seed)
    mat = ob.wrapper_matrix(lp.numpy(), v.blank_id, v.pad_id, 0.0, dec.keep)
    jobs = [(mat[i, : int(ln[i])], dec.labels, width, CFG3["prune"], -5.0) for i in range(n_utts)]
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


LM_CFG = {"words": 2000, "lines": 20000, "order": 5, "alpha": 0.6, "beta": 0.0, "seed": 4242}


def lm_workload(n: int, T: int, seed: int):
    """Scope row f-4 at config 3's shape: a 5-gram word LM (what build_char_lm.py:35-40 asks lmplz for) over a synthetic
    2000-word domain, and n utterances whose posteriors spell in-domain sentences with a rival letter on every emitting
    frame (no-LM beam search gets most words wrong on them; the LM decides).  -> (lp, lengths, arpa path)"""
    import tempfile

    from silent_speech_decoder_b200 import synth
    from silent_speech_decoder_b200.data.vocab import Vocab

    c = LM_CFG
    words = synth.word_list(c["words"], c["seed"])
    arpa = os.path.join(tempfile.gettempdir(), f"ssd_bench_lm_{os.getpid()}.arpa")
    synth.write_arpa(arpa, synth.word_corpus(words, c["lines"], c["seed"] + 1), c["order"])
    sents = synth.word_corpus(words, n, seed + 17)
    lp, ln = synth.spelled(sents, Vocab.default().token_to_id, T, seed)
    return lp, ln, arpa


_LM_DEC = None


def _beam_lm_init(arpa, width, prune, alpha, beta):
    global _LM_DEC
    from oracle import beam as ob
    from oracle import lm as olm
    from silent_speech_decoder_b200.data.vocab import Vocab

    v = Vocab.default()
    _LM_DEC = ob.BeamDecoder(v.tokens, v.blank_id, v.pad_id, width, prune, lm=olm.LanguageModel(olm.ArpaModel(arpa), alpha, beta))


def _beam_lm_worker(job):
    lp, n = job
    return _LM_DEC(lp[None], [n])[0]


def cpu_beam_lm_reference(n_utts: int, budget_s: float = 12.0, T: int = CFG3["T"], width: int = CFG3["width"], seed: int = CFG3["seed"]):
    """oracle/beam.py + oracle/lm.py (restatements of pyctcdecode / kenlm; both unavailable) over all host cores."""
    import multiprocessing as mp

    lp, ln, arpa = lm_workload(n_utts, T, seed)
    jobs = [(lp[i].numpy(), int(ln[i])) for i in range(n_utts)]
    cores = os.cpu_count() or 1
    c = LM_CFG
    _beam_lm_init(arpa, width, CFG3["prune"], c["alpha"], c["beta"])
    t0 = time.perf_counter()
    _beam_lm_worker(jobs[0])
    one = time.perf_counter() - t0
    n = min(n_utts, max(cores, int(budget_s * cores / max(one, 1e-3))))
    with mp.get_context("fork").Pool(cores) as pool:  # forked workers inherit the decoder built above
        t0 = time.perf_counter()
        pool.map(_beam_lm_worker, jobs[:n], chunksize=1)
        dt = time.perf_counter() - t0
    os.unlink(arpa)
    return n / dt, cores, n, one


def run_reference(args, rank):
    if rank != 0:
        return
    if args.workload == "loss":
        val, ms, cores = cpu_loss_reference(args.steps, args.warmup)
        line = {"metric": "ctc_loss_fwd_bwd_frames_per_sec", "value": val, "unit": "frames/s", "ms_per_step": ms,
                "sample": CPU_LOSS_SAMPLE, "cores": cores, "kind": CPU_LOSS_KIND, "config": loss_config(args.gpus)}
    else:
        n_utts = min(args.utts, 256)
        val, cores, n, one = cpu_beam_reference(n_utts, args.beam_input, T=args.T, width=args.width)
        line = {"metric": "ctc_beam50_decode_utterances_per_sec" if args.width == 50 else f"ctc_beam{args.width}_decode_utterances_per_sec",
                "value": val, "unit": "utterances/s", "ms_per_step": 1e3 * n / val,
                "sample": f"{n} utterances ({args.beam_input}, T={args.T}, width {args.width}) over {cores} processes; oracle/beam.py restatement of pyctcdecode",
                "cores": cores, "kind": "port", "config": decode_config("beam", args.utts, args.T, args.beam_input, args.width, args.gpus)}
    out = {
        "impl": "reference", "metric": line["metric"], "value": line["value"], "unit": line["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": line["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.workload == "loss" else "f64", "data": "synthetic",
        "config": line["config"],
        "cpu_baseline": {"value": line["value"], "unit": line["unit"], "cores": line["cores"], "kind": line["kind"], "sample": line["sample"]},
        "e2e": {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ native arm
def _loss_abi_call(lib, ops, x, il, tg, tl, nll, grad, ws, global_b, stream):
    """One ssd_ctc_loss call straight through the C ABI on pre-allocated buffers: nothing but the library's own work
    (two 256-byte memsets, the fp32 kernel, the fp64 kernel over flagged rows) lands on the stream."""
    B, T, V = x.shape
    lib.check(lib.ssd_ctc_loss(x.data_ptr(), B, T, V, T * V, V, tg.data_ptr(), tg.stride(0), tg.size(1), il.data_ptr(), tl.data_ptr(),
                               None, 1, 1, global_b, None, nll.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(), stream))


def time_loss_launch(inputs_dev, il_d, tg_d, tl_d, global_b, steps):
    """Average device time of one ssd_ctc_loss call over `steps` back-to-back calls (events on the launching stream)."""
    from silent_speech_decoder_b200 import _native, ops

    lib = _native.load()
    B, T, V = inputs_dev[0].shape
    dev = inputs_dev[0].device
    nll = torch.zeros(B, dtype=torch.float32, device=dev)
    grads = [torch.empty_like(x) for x in inputs_dev]
    ws = torch.empty(int(lib.ssd_ctc_loss_workspace_bytes(B, T, V, tg_d.size(1))), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for i in range(3):
        _loss_abi_call(lib, ops, inputs_dev[i % len(inputs_dev)], il_d, tg_d, tl_d, nll, grads[i % len(grads)], ws, global_b, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        _loss_abi_call(lib, ops, inputs_dev[i % len(inputs_dev)], il_d, tg_d, tl_d, nll, grads[i % len(grads)], ws, global_b, stream)
    e1.record()
    torch.cuda.synchronize()
    off = int(lib.ssd_ctc_loss_fallback_counter_offset(B, T, V, tg_d.size(1)))
    n_exact = None if off == 2 ** 64 - 1 else int(ws[off:off + 4].view(torch.int32).item())
    return e0.elapsed_time(e1) / steps, n_exact


def torch_cuda_ctc_ms(lp, il, tg, tl, dev, steps=5):
    """torch.nn.functional.ctc_loss on CUDA, forward + backward, same inputs (the library kernel SURVEY 2.1 names)."""
    x = lp.to(dev).transpose(0, 1).contiguous().requires_grad_(True)
    tg_d = tg.to(dev)

    def step():
        x.grad = None
        loss = torch.nn.functional.ctc_loss(x, tg_d, il, tl, blank=1, reduction="mean", zero_infinity=True)
        loss.backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_loss(args, rank, world, dev):
    from silent_speech_decoder_b200 import dist as sdist
    from silent_speech_decoder_b200 import ops
    from silent_speech_decoder_b200.data.prefetch import DevicePrefetcher
    from silent_speech_decoder_b200.models.losses import DistillationCTCLoss, LossWeights

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

    # duration of the loss launch itself for the roofline: back-to-back C-ABI calls, events on the launching stream
    k_ms, n_fallback = time_loss_launch(lps, il_d, tg_d, tl_d, global_b, min(args.steps, 100))
    fast = n_fallback is not None

    # end to end through the reference-facing module, host buffers, next step's copy in flight during this step
    fn = DistillationCTCLoss(vocab_size=V, blank_id=1, weights=LossWeights(0.0, 1.0))
    host = tuple(t.pin_memory() for t in (lp, il, tg, tl))
    h2d = sum(t.numel() * t.element_size() for t in host)
    e2e_steps = max(3, min(args.steps, 30))
    n_warm = max(1, min(args.warmup, 3))

    def host_batches(n):
        for _ in range(n):
            yield host

    def e2e_run(n):
        last = 0.0
        for x, il_, tg_, tl_ in DevicePrefetcher(host_batches(n), dev):
            x.requires_grad_(True)
            out = fn(x, il_, tg_, tl_, x, None)
            out["total"].backward()
            last = float(out["ctc"].item())  # D2H read of the step's result
        return last

    e2e_run(n_warm)
    barrier(world)
    t0 = time.perf_counter()
    e2e_run(e2e_steps)  # the first batch's copy is inside the timed region too
    torch.cuda.synchronize()
    e2e_ms = dist_max(1e3 * (time.perf_counter() - t0) / e2e_steps, world, dev)

    peak, peak_src = measured_peak_gbs()
    achieved = LOSS_BYTES_PER_FRAME * B * T / (k_ms * 1e-3) / 1e9
    line = {
        "metric": "ctc_loss_fwd_bwd_frames_per_sec", "value": world * B * T / (ms_step * 1e-3), "unit": "frames/s",
        "ms_per_step": ms_step, "dtype": "f32" if fast else "f64", "config": loss_config(world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic("loss_fast_kernel" if fast else "loss_kernel", B * T), "peak_source": peak_src,
                     "kernel": "ssd::loss_fast_kernel (+ ssd::loss_kernel over flagged utterances)" if fast else "ssd::loss_kernel",
                     "kernel_ms": k_ms, "algorithmic_bytes": LOSS_BYTES_PER_FRAME * B * T,
                     "exact_path_utterances": n_fallback if fast else B,
                     "note": "kernel_ms = average over back-to-back ssd_ctc_loss calls on pre-allocated buffers (both launches of a "
                             "call); the loss is bound by its serial recursion and instruction issue, not HBM (DESIGN.md section 4)"},
        "e2e": {"value": world * B * T / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "api": "DistillationCTCLoss(...)(log_probs, ...)['total'].backward(), batches from data.prefetch.DevicePrefetcher"},
        "gpu_launches": launches,
    }
    extras = {}
    if not args.no_secondary:
        # the regime a trained model puts the loss in
        tlp, til, ttg, ttl = cfg2_trained_inputs()
        tsets = [tlp.to(dev).clone() for _ in range(nsets)]
        t_ms, t_exact = time_loss_launch(tsets, til.to(dev), ttg.to(dev), ttl.to(dev), global_b, min(args.steps, 100))
        extras["loss_cfg2_trained"] = {
            "metric": "ctc_loss_fwd_bwd_frames_per_sec", "value": world * B * T / (t_ms * 1e-3), "unit": "frames/s", "ms_per_step": t_ms,
            "exact_path_utterances": t_exact, "vs_flat_input": t_ms / k_ms,
            "config": {"workload": "cfg2 shape, confident posteriors (+8 logit) on each utterance's own transcript, labels U[50,200] with English "
                                   "label frequencies (space ~0.15)"},
            "roofline": {"bound": "hbm", "achieved": LOSS_BYTES_PER_FRAME * B * T / (t_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": LOSS_BYTES_PER_FRAME * B * T / (t_ms * 1e-3) / 1e9 / peak}}
        del tsets
        # the same step from logits (scope row f-2)
        lg = [x + 3.0 for x in lps[:2]]
        for i in range(3):
            ops.ctc_loss_fwd_bwd(lg[i % 2], il_d, tg_d, tl_d, blank=1, mean_denominator=global_b, from_logits=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n_l = min(args.steps, 50)
        for i in range(n_l):
            ops.ctc_loss_fwd_bwd(lg[i % 2], il_d, tg_d, tl_d, blank=1, mean_denominator=global_b, from_logits=True)
        b.record()
        torch.cuda.synchronize()
        l_ms = a.elapsed_time(b) / n_l
        extras["loss_cfg2_from_logits"] = {"ms_per_step": l_ms, "value": world * B * T / (l_ms * 1e-3), "unit": "frames/s",
                                           "note": "ssd_ctc_loss_from_logits: row log-sum-exp pre-pass + the loss kernels; replaces log_softmax "
                                                   "forward, the loss and log_softmax backward of the unfused path"}
        del lg
        if rank == 0 and world == 1:
            g_ms = torch_cuda_ctc_ms(lp, il, tg, tl, dev)
            extras["gpu_library_baseline"] = {"kernel": "torch.nn.functional.ctc_loss (CUDA, fp32) forward + backward", "ms_per_step": g_ms,
                                              "value": B * T / (g_ms * 1e-3), "unit": "frames/s", "native_speedup": g_ms / ms_step}
    return line, extras


def bench_decode(args, rank, world, dev, method: str, kind: str, n: int, T: int, width: int, seed: int, e2e: bool = True):
    from silent_speech_decoder_b200 import ops
    from silent_speech_decoder_b200.data.vocab import Vocab
    from silent_speech_decoder_b200.decoding import ctc

    v = Vocab.default()
    prune = CFG3["prune"]
    n_local = n if args.scaling == "weak" else n // world
    lp, ln = decode_inputs(n_local, T, kind, seed + rank)
    x, ln_d = lp.to(dev), ln.to(dev)
    _, space = ctc._beam_label_strings(v)

    def kernel():
        if method == "beam":
            return ops.beam_ids(x, ln_d, v.blank_id, v.pad_id, space, width, prune)
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

    peak, peak_src = measured_peak_gbs()
    achieved = DECODE_BYTES_PER_FRAME * n_local * T / (ms * 1e-3) / 1e9
    kname = "beam_kernel" if method == "beam" else "greedy_kernel"  # traffic.json keys; the kernel is ssd::beam_kernel<false>
    out = {
        "metric": f"ctc_{('beam' + str(width)) if method == 'beam' else 'greedy'}_decode_utterances_per_sec",
        "value": world * n_local / (ms * 1e-3), "unit": "utterances/s", "ms_per_step": ms, "dtype": "f64" if method == "beam" else "f32",
        "config": decode_config(method, n_local, T, kind, width, world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(kname + (("_" + kind) if method == "beam" else ""), n_local * T), "peak_source": peak_src,
                     "kernel": f"ssd::{kname}" + ("<false>" if method == "beam" else ""), "kernel_ms": ms, "algorithmic_bytes": DECODE_BYTES_PER_FRAME * n_local * T,
                     "note": "beam search is latency/issue bound, not HBM bound; see profiles/ for stall breakdown" if method == "beam" else ""},
        "gpu_launches": launches,
    }
    if e2e:
        decoder = ctc.build_decoder(method, v, beam_width=width, beam_prune_logp=prune)
        h_lp, h_ln = lp.pin_memory(), ln.pin_memory()
        del x
        decoder(h_lp, h_ln)  # H2D of the posteriors (chunked, overlapped with decode), D2H of ids, strings built
        barrier(world)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            texts = decoder(h_lp, h_ln)
        e2e_ms = dist_max(1e3 * (time.perf_counter() - t0) / reps, world, dev)
        out["e2e"] = {"value": world * n_local / (e2e_ms * 1e-3), "unit": "utterances/s", "h2d_bytes_per_step": h_lp.numel() * 4 + h_ln.numel() * 8,
                      "d2h_bytes_per_step": n_local * 4 + sum(len(t) for t in texts) * 4, "ms_per_step": e2e_ms,
                      "api": f"build_decoder('{method}', vocab)(log_probs, lengths) on pinned host tensors"}
    return out


def bench_beam_lm(args, rank, world, dev, n: int, T: int, width: int, seed: int):
    """Row f-4: LM-fused beam search (ssd_ctc_beam_lm) on config 3's shape."""
    from silent_speech_decoder_b200 import ops
    from silent_speech_decoder_b200.data.vocab import Vocab
    from silent_speech_decoder_b200.decoding import ctc
    from silent_speech_decoder_b200.decoding import lm as plm

    v = Vocab.default()
    c = LM_CFG
    lp, ln, arpa = lm_workload(n, T, seed + rank)
    text, space = ctc._beam_label_strings(v)
    keep = [i for i in range(v.size) if i not in {v.blank_id, v.pad_id}]
    tables = plm.build_tables(arpa, [""] + [text[i] for i in keep], c["alpha"], c["beta"])
    x, ln_d = lp.to(dev), ln.to(dev)
    settings = [(width, CFG3["prune"], 0.0, -5.0)]

    def kernel():
        return ops.beam_lm_ids(x, ln_d, v.blank_id, v.pad_id, space, tables, settings)

    for _ in range(2):
        kernel()
    barrier(world)
    steps = 3
    l0 = ops.LAUNCHES["beam"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tok, cnt, _ = kernel()
    e1.record()
    barrier(world)
    ms = dist_max(e0.elapsed_time(e1), world, dev) / steps
    launches = ops.LAUNCHES["beam"] - l0
    hyp = ctc._rows_to_strings(tok[0], cnt[0], text, normalise_space=True)
    no_lm = ctc.build_decoder("beam", v, beam_width=width, beam_prune_logp=CFG3["prune"])(x[:512], ln_d[:512])
    from silent_speech_decoder_b200 import synth

    sents = synth.word_corpus(synth.word_list(c["words"], c["seed"]), n, seed + rank + 17)  # the sentences lm_workload spelled
    cut = [s[: max(1, (T - 2) // 2)] for s in sents]
    exact = sum(h == " ".join(r.split()) for h, r in zip(hyp, cut)) / len(hyp)
    exact0 = sum(h == " ".join(r.split()) for h, r in zip(no_lm, cut[:512])) / len(no_lm)
    peak, peak_src = measured_peak_gbs()
    achieved = DECODE_BYTES_PER_FRAME * n * T / (ms * 1e-3) / 1e9
    out = {
        "metric": f"ctc_beam{width}_lm_decode_utterances_per_sec", "value": world * n / (ms * 1e-3), "unit": "utterances/s",
        "ms_per_step": ms, "dtype": "f64",
        "config": {"workload": f"row f-4 at cfg3's shape: LM-fused beam decode of {n} utterances per GPU, T={T}, V=48, width {width}, prune -10, "
                               f"{c['order']}-gram word LM over {c['words']} words ({len(tables.ngram_prob)} n-grams), alpha {c['alpha']}, beta {c['beta']}",
                   "l2": f"input {n * T * 48 * 4 / 1e6:.0f} MB > 126 MB L2", "parallelism": f"dp{world} by utterance, no data-path collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": recorded_traffic("beam_kernel_lm", n * T),
                     "peak_source": peak_src, "kernel": "ssd::beam_kernel<true>", "kernel_ms": ms,
                     "algorithmic_bytes": DECODE_BYTES_PER_FRAME * n * T,
                     "note": "latency/issue bound like the no-LM search; every beam proposes every candidate label (no staircase bound under LM scores)"},
        "gpu_launches": launches,
        "sentences_exactly_right": {"with_lm": exact, "without_lm_first_512": exact0},
    }
    decoder = ctc.build_decoder("beam", v, lm_path=arpa, beam_width=width, alpha=c["alpha"], beta=c["beta"], beam_prune_logp=CFG3["prune"])
    h_lp, h_ln = lp.pin_memory(), ln.pin_memory()
    del x
    decoder(h_lp, h_ln)
    barrier(world)
    t0 = time.perf_counter()
    texts = decoder(h_lp, h_ln)
    e2e_ms = dist_max(1e3 * (time.perf_counter() - t0), world, dev)
    out["e2e"] = {"value": world * n / (e2e_ms * 1e-3), "unit": "utterances/s", "h2d_bytes_per_step": h_lp.numel() * 4 + h_ln.numel() * 8,
                  "d2h_bytes_per_step": n * 4 + sum(len(t) for t in texts) * 4, "ms_per_step": e2e_ms,
                  "api": "build_decoder('beam', vocab, lm_path=<arpa>, alpha, beta)(log_probs, lengths) on pinned host tensors"}
    os.unlink(arpa)
    return out


def bench_cfg4(args, rank, world, dev):
    """Config 4: beam width sweep 1/10/50/200 on long utterances (T=4000), peaked and flat, 256 utterances per GPU."""
    c = CFG4
    rows = {}
    for kind in ("peaked", "flat"):
        for w in c["widths"]:
            r = bench_decode(args, rank, world, dev, "beam", kind, c["N"], c["T"], w, c["seed"], e2e=False)
            rows[f"{kind}_w{w}"] = {"value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"], "hbm_frac": r["roofline"]["frac"]}
    return {"config": {"workload": f"cfg4: beam widths {list(c['widths'])} on {c['N']} utterances per GPU, T={c['T']}, peaked and flat posteriors"},
            "results": rows,
            "cpu_reference": "profiles/r01_cfg4_cpu.json (oracle/beam.py on 16 cores: peaked 205/97/33/13, flat 61/10/1.8/0.44 utterances/s); "
                             "`bench.py --impl reference --workload beam --T 4000 --width W` re-times one point on this box"}


def bench_cfg5(args, rank, world, dev):
    """Config 5: the evaluate pipeline (reference src/evaluation/evaluate.py:290-328) on a synthetic
    silent_parallel_data-shaped test set: PyTorch encoder + CTC head forward (random-init mps_fast_plus), native
    beam-50 on the log-probs where the model left them, transcripts gathered on rank 0, WER/CER on the GPU.
    The fixed set of 159 x 64 utterances is split over the ranks by cost (dist.shard_by_cost)."""
    import torch.distributed as dist

    from silent_speech_decoder_b200 import dist as sdist
    from silent_speech_decoder_b200.data.vocab import Vocab
    from silent_speech_decoder_b200.decoding import ctc
    from silent_speech_decoder_b200.evaluation.evaluate import Evaluator
    from silent_speech_decoder_b200.models.emg_encoder import MPS_FAST_PLUS, EMGConformerEncoder, EncoderConfig
    from silent_speech_decoder_b200.models.heads import CTCHead

    c = CFG5
    v = Vocab.default()
    g = torch.Generator().manual_seed(c["seed"])
    lengths = torch.randint(c["t_lo"], c["t_hi"] + 1, (c["N"],), generator=g)   # raw frames: 2-12 s at 100 frames/s
    n_words = torch.randint(3, 12, (c["N"],), generator=g).tolist()
    letters = [t for t in v.tokens if len(t) == 1 and t.isalpha()]
    pick = torch.randint(0, len(letters), (sum(n_words) * 5,), generator=g).tolist()
    refs, k = [], 0
    for nw in n_words:
        refs.append(" ".join("".join(letters[i] for i in pick[k + 5 * j: k + 5 * j + 5]) for j in range(nw)))
        k += 5 * nw
    mine = sdist.shard_by_cost(lengths.tolist(), world)[rank]   # longest first
    torch.manual_seed(c["seed"])
    enc = EMGConformerEncoder(EncoderConfig(**MPS_FAST_PLUS)).to(dev).eval()
    head = CTCHead(MPS_FAST_PLUS["d_model"], v.size, dropout=0.12).to(dev).eval()
    decoder = ctc.build_decoder("beam", v, beam_width=50, beam_prune_logp=-10.0)
    # a small pool of pinned host feature batches, copied every step (the reference's loader yields host tensors)
    t_max = int(lengths[mine[0]]) if mine else c["t_lo"]
    pool = [torch.randn(c["batch"], t_max, c["feat"], generator=g).pin_memory() for _ in range(2)]
    ev_enc = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def run():
        e = Evaluator(decoder, {"type": "beam", "beam_width": 50, "beam_prune_logp": -10.0, "blank_bias": 0.0, "lm_path": None})
        enc_ms = 0.0
        with torch.no_grad():
            for bi, lo in enumerate(range(0, len(mine), c["batch"])):
                idx = mine[lo: lo + c["batch"]]
                ln = lengths[idx]
                tb = int(ln.max())
                ev_enc[0].record()
                # the pool holds noise: reinterpret its leading bytes as this batch's (B, T, feat) block -- one contiguous copy
                x = pool[bi & 1].view(-1)[: len(idx) * tb * c["feat"]].view(len(idx), tb, c["feat"]).to(dev, non_blocking=True)
                y, out_len = enc(x, ln.to(dev))
                lp = head(y)
                ev_enc[1].record()
                e.add(idx, [refs[i] for i in idx], lp, out_len)   # native beam-50 on the device tensor; strings on the host
                enc_ms += ev_enc[0].elapsed_time(ev_enc[1])
        return e, enc_ms

    run()  # warm-up: cuDNN/cuBLAS autotune, allocator
    if world > 1:  # and the object collective's one-time set-up (communicator, pickling buffers)
        warm = [None] * world if rank == 0 else None
        dist.gather_object([{"utterance_id": -1, "ref": "", "hyp": ""}], warm, dst=0)
    barrier(world)
    from silent_speech_decoder_b200 import ops

    l0 = sum(ops.LAUNCHES.values())
    t0 = time.perf_counter()
    e, enc_ms = run()
    torch.cuda.synchronize()
    local_s = time.perf_counter() - t0
    native_launches = sum(ops.LAUNCHES.values()) - l0
    # transcripts to rank 0 (a few hundred KB of text)
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((e.records), gathered, dst=0)
        records = [r for part in gathered for r in part] if rank == 0 else []
    else:
        records = e.records
    total_s = dist_max(time.perf_counter() - t0, world, dev)
    out = {"config": {"workload": f"cfg5: evaluate pipeline, {c['N']} utterances (159 x 64), raw T~U[{c['t_lo']},{c['t_hi']}] x {c['feat']} features, "
                                  f"mps_fast_plus encoder (random init, fp32) + CTC head in PyTorch, native beam-50, WER/CER; batches of {c['batch']}",
                      "parallelism": f"{world} rank(s), utterances dealt by cost, transcripts gathered on rank 0"},
           "value": c["N"] / total_s, "unit": "utterances/s", "seconds": total_s, "scaling": "strong",
           "encoder_share": dist_max(enc_ms * 1e-3, world, dev) / total_s,
           "decode_and_host_share": 1.0 - dist_max(enc_ms * 1e-3, world, dev) / total_s,
           "rank0_local_seconds": local_s, "native_launches": native_launches}
    if rank == 0:
        from silent_speech_decoder_b200.evaluation import evaluate as evm

        t1 = time.perf_counter()
        records.sort(key=lambda r: r["utterance_id"])
        m = evm.compute_metrics([r["ref"] for r in records], [r["hyp"] for r in records])
        out["scoring_seconds"] = time.perf_counter() - t1
        out["num_samples"] = len(records)
        out["wer"] = m["wer"]
        out["cer"] = m["cer"]
    return out


def run_native(args, rank, world, local):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native path has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    secondary = {}
    with ClockSampler(local) as clk:
        if args.workload == "loss":
            line, secondary = bench_loss(args, rank, world, dev)
        elif args.workload == "eval":
            r = bench_cfg5(args, rank, world, dev)
            line = {"metric": "evaluate_pipeline_utterances_per_sec", "value": r["value"], "unit": "utterances/s", "ms_per_step": 1e3 * r["seconds"],
                    "dtype": "f32", "config": r["config"], "e2e": {"value": r["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": r["native_launches"], "roofline": None, "cfg5": r}
        else:
            line = bench_decode(args, rank, world, dev, args.workload, args.beam_input, args.utts, args.T, args.width, CFG3["seed"] if args.T == CFG3["T"] else CFG4["seed"])
        if args.workload == "loss" and not args.no_secondary:
            keep = ("metric", "value", "unit", "ms_per_step", "roofline", "e2e", "config", "gpu_launches")
            for kind in ("peaked", "flat"):
                r = bench_decode(args, rank, world, dev, "beam", kind, CFG3["N"], CFG3["T"], CFG3["width"], CFG3["seed"])
                secondary[f"beam50_{kind}"] = {k: r[k] for k in keep if k in r}
            r = bench_decode(args, rank, world, dev, "greedy", "flat", CFG3["N"], CFG3["T"], 0, CFG3["seed"])
            secondary["greedy"] = {k: r[k] for k in keep if k in r}
            secondary["beam50_lm"] = bench_beam_lm(args, rank, world, dev, CFG3["N"], CFG3["T"], CFG3["width"], CFG3["seed"])
            secondary["cfg4_width_sweep"] = bench_cfg4(args, rank, world, dev)
            secondary["cfg5_eval"] = bench_cfg5(args, rank, world, dev)
        time.sleep(0.12)
    if rank != 0:
        return
    if world == 1 and not args.no_cpu:
        if args.workload == "loss":
            val, ms, cores = cpu_loss_reference(3, 1)
            line["cpu_baseline"] = {"value": val, "unit": "frames/s", "cores": cores, "kind": CPU_LOSS_KIND, "sample": "3 " + CPU_LOSS_SAMPLE, "ms_per_step": ms}
            if secondary:
                for kind in ("peaked", "flat"):
                    v, cores, n, one = cpu_beam_reference(256, kind, budget_s=12.0)
                    secondary[f"beam50_{kind}"]["cpu_baseline"] = {
                        "value": v, "unit": "utterances/s", "cores": cores, "kind": "port",
                        "sample": f"{n} cfg3 utterances ({kind}) over {cores} processes; oracle/beam.py restatement of pyctcdecode (library unavailable)",
                        "single_process_utt_per_s": 1.0 / one}
                v, cores, n, one = cpu_beam_lm_reference(128, budget_s=10.0)
                secondary["beam50_lm"]["cpu_baseline"] = {
                    "value": v, "unit": "utterances/s", "cores": cores, "kind": "port",
                    "sample": f"{n} utterances of the same workload over {cores} processes; oracle/beam.py + oracle/lm.py restatements of "
                              "pyctcdecode / kenlm (both unavailable)", "single_process_utt_per_s": 1.0 / one}
        elif args.workload == "beam":
            v, cores, n, one = cpu_beam_reference(min(args.utts, 256), args.beam_input, T=args.T, width=args.width)
            line["cpu_baseline"] = {"value": v, "unit": "utterances/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} utterances (T={args.T}, width {args.width}) over {cores} processes; oracle/beam.py restatement of pyctcdecode"}
    out = {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
           "vs_baseline": None, "dtype": line["dtype"], "data": "synthetic", "config": line["config"], "clocks": clk.summary(),
           "e2e": line.get("e2e"), "gpu_launches": line["gpu_launches"], "roofline": line["roofline"]}
    for k in ("cpu_baseline", "cfg5"):
        if k in line:
            out[k] = line[k]
    out.update(secondary)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=["loss", "beam", "greedy", "eval"], default="loss")
    ap.add_argument("--beam-input", choices=["peaked", "flat"], default="peaked")
    ap.add_argument("--T", type=int, default=CFG3["T"], help="frames per utterance of the decode workloads (4000 = config 4)")
    ap.add_argument("--width", type=int, default=CFG3["width"], help="beam width of --workload beam")
    ap.add_argument("--utts", type=int, default=None, help="utterances per GPU of the decode workloads (default 10000; 256 at T=4000)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.utts is None:
        args.utts = CFG3["N"] if args.T <= 1000 else CFG4["N"]
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
