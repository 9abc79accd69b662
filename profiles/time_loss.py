"""Kernel-only timing of the loss path on the config-2 workload (CUDA events, rotating inputs > L2).
    python profiles/time_loss.py [reps]   -> fast path (fp32 + flagged fp64 fallbacks) and SSD_LOSS_EXACT=1 (fp64 only)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from silent_speech_decoder_b200 import _native, ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda", 0)
lp, il, tg, tl = bench.cfg2_inputs()
sets = [(lp.to(dev).clone(), il.to(dev), tg.to(dev), tl.to(dev)) for _ in range(4)]
lib = _native.load()
B, T, V = lp.shape
off = lib.ssd_ctc_loss_fallback_counter_offset(B, T, V, tg.shape[1])


def run(tag):
    for i in range(5):
        nll, g = ops.ctc_loss_fwd_bwd(*sets[i % 4], blank=1, mean_denominator=256)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        nll, g = ops.ctc_loss_fwd_bwd(*sets[i % 4], blank=1, mean_denominator=256)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag}: {ms * 1e3:.1f} us/step  {B * T / ms / 1e3:.1f} Mframes/s  roofline {B * T * V * 8 / ms / 1e6 / 6543.7:.3f}  nll mean {float(nll.mean()):.4f}")
    return nll, g


nll_f, g_f = run("fast+fallback")
print("fallback utterances:", ops.loss_fallback_count(B, T, V, tg.shape[1]))
os.environ["SSD_LOSS_EXACT"] = "1"
nll_e, g_e = run("exact fp64 only")
print("max |grad fast - exact| / max|grad|:", float((g_f - g_e).abs().max() / g_e.abs().max()), " nll rel:", float(((nll_f - nll_e) / nll_e).abs().max()))
