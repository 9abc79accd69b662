"""Secondary measurements SURVEY.md section 8(d) asks for next to the headline (device-timed, inputs resident):
  * loss fwd+bwd at B in {256, 1024, 4096} (cfg2 shape otherwise)           -> frames/s, fraction of the HBM roofline
  * greedy decode of cfg3 (10 000 x T=500)                                   -> utterances/s
  * beam width sweep of cfg4 (T=4000, widths 1/10/50/200, peaked and flat)    -> utterances/s
    python profiles/sweeps.py > profiles/r01_sweeps.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from silent_speech_decoder_b200 import ops, synth
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc

dev = torch.device("cuda", 0)
peak, _ = bench.measured_peak_gbs()
out = {"hbm_peak_gbs": peak}


def timed(fn, reps):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- loss, batch sweep (rotating input sets larger than L2)
loss = []
for B in (256, 1024, 4096):
    T, V, S = 1000, 48, 200
    g = torch.Generator().manual_seed(1002)
    tl = torch.randint(S // 4, S + 1, (B,), generator=g, dtype=torch.int64)
    tg = torch.randint(2, V, (B, S), generator=g, dtype=torch.int64) * (torch.arange(S)[None, :] < tl[:, None])
    il = torch.full((B,), T, dtype=torch.int64)
    nsets = 4 if B == 256 else 2
    lps = [synth.flat(B, T, 1002 + k, V).to(dev) for k in range(nsets)]
    il_d, tg_d, tl_d = il.to(dev), tg.to(dev), tl.to(dev)
    ms = timed(lambda i: ops.ctc_loss_fwd_bwd(lps[i % nsets], il_d, tg_d, tl_d, blank=1, mean_denominator=B), 30 if B == 256 else 10)
    loss.append({"B": B, "T": T, "ms": ms, "frames_per_s": B * T / ms * 1e3, "hbm_frac": 384 * B * T / (ms * 1e-3) / 1e9 / peak,
                 "exact_path_utterances": ops.loss_fallback_count(B, T, V, S, dev)})
    del lps
    torch.cuda.empty_cache()
out["loss_batch_sweep"] = loss

v = Vocab.default()
_, space = ctc._beam_label_strings(v)
# ---- greedy, cfg3
for kind in ("peaked", "flat"):
    lp, ln = bench.cfg3_inputs(10000, kind)
    x, ln_d = lp.to(dev), ln.to(dev)
    ms = timed(lambda i: ops.greedy_ids(x, ln_d, v.blank_id), 10)
    out[f"greedy_cfg3_{kind}"] = {"N": 10000, "T": 500, "ms": ms, "utt_per_s": 10000 / ms * 1e3, "hbm_frac": 192 * 10000 * 500 / (ms * 1e-3) / 1e9 / peak}
    del x
torch.cuda.empty_cache()

# ---- beam width sweep, cfg4
sweep = []
N, T = 256, 4000
for kind in ("peaked", "flat"):
    lp = (synth.peaked(32, T, 1004).repeat(N // 32, 1, 1) if kind == "peaked" else synth.flat(N, T, 1004)).contiguous()
    x, ln_d = lp.to(dev), torch.full((N,), T, dtype=torch.int64, device=dev)
    for w in (1, 10, 50, 200):
        ms = timed(lambda i: ops.beam_ids(x, ln_d, v.blank_id, v.pad_id, space, w, -10.0), 3)
        sweep.append({"input": kind, "N": N, "T": T, "width": w, "ms": ms, "utt_per_s": N / ms * 1e3})
    del x
out["beam_width_sweep_cfg4"] = sweep
print(json.dumps(out, indent=1))
