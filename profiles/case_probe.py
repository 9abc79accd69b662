"""One fuzz case three ways: fp32 fast path (GPU), fp64 kernel (GPU), torch fp64 (CPU).  python profiles/case_probe.py seed case"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from silent_speech_decoder_b200 import ops, synth
from oracle import ctc_loss as ol
seed, case = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
for c in range(case + 1):
    B = int(rng.choice([1, 2, 3, 7, 33, 150, 297, 600, 900]))
    T = int(rng.choice([1, 5, 8, 9, 40, 97, 160, 200, 333, 600]))
    S = int(min(rng.choice([1, 2, 10, 40, 63, 64, 100, 127, 128, 200, 255]), max(1, T)))
    kind = rng.choice(["flat", "peaked", "mild", "trained"])
    grid = int(rng.choice([0, 0, 1, 5, 100]))
    rng.random()
    if rng.random() < 0.5:
        rng.integers(0, 2)
print(c, B, T, S, kind)
if kind == "trained" and T >= 8:
    lp, tg, tl = synth.peaked_aligned(B, T, 7000 + c, n_labels=(1, max(1, min(S, T // 2))))
    if tg.size(1) < S:
        tg = torch.nn.functional.pad(tg, (0, S - tg.size(1)))
else:
    lp = synth.flat(B, T, 7000 + c) if kind == "flat" else synth.peaked(B, T, 7000 + c, boost=8.0 if kind == "peaked" else 2.0)
    tg, tl = synth.targets(B, S, 8000 + c, s_min=1)
il = synth.ragged_lengths(B, T, 9000 + c)
if B > 2:
    il[1] = 0
    tl[2] = 0
n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
x = lp.cuda()
nf, gf = ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=B)  # fast path + whatever it hands down
torch.cuda.synchronize()
print("rows on the exact path:", ops.loss_fallback_count(B, T, 48, tg.shape[1]))
os.environ["SSD_LOSS_EXACT"] = "1"
nx, gx = ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=B)
torch.cuda.synchronize()
gmax = g64.abs().max()
ef = (gf.cpu().double() - g64).abs().amax(dim=(1, 2)) / gmax
ex = (gx.cpu().double() - g64).abs().amax(dim=(1, 2)) / gmax
print("fast  vs torch fp64: worst", float(ef.nan_to_num(9).max()), "row", int(ef.nan_to_num(9).argmax()))
print("exact vs torch fp64: worst", float(ex.nan_to_num(9).max()), "row", int(ex.nan_to_num(9).argmax()))
w = int(ef.nan_to_num(9).argmax())
t = int((gf[w].cpu().double() - g64[w]).abs().amax(dim=1).argmax())
print(" worst frame", t, "of row", w, "grad max there: fast", float(gf[w, t].abs().max()), "exact", float(gx[w, t].abs().max()), "fp64", float(g64[w, t].abs().max()), "| row max |grad|", float(g64[w].abs().max()), "batch max", float(gmax))
print(" row", w, "il", int(il[w]), "tl", int(tl[w]), "nll64", float(n64[w]), "nll exact", float(nx[w]), "nll fast", float(nf[w]))
