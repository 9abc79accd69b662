"""Per-role mean block timings of ONE utterance alone on the GPU (B=1, T=1000, S=200), stamped build:
    profiles/build_variant.sh tl "-DSSD_LOSS_TIMELINE" ; SSD_LIB=profiles/scratch/tl.so python profiles/timeline_b1.py [grad]
mains: wait PFULL | sweep 8 frames + arrive | renorm + checkpoint ; loaders: wait PEMPTY | wait RAW | convert + arrive"""
import os, sys
os.environ["SSD_LOSS_DBG"] = "8"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathlib import Path
import numpy as np, torch
from silent_speech_decoder_b200 import _native, ops, synth
if os.environ.get("SSD_LIB"):
    _native._LIB = _native.Library(Path(os.environ["SSD_LIB"]).resolve())
want = len(sys.argv) > 1
B, T, S = 1, 1000, 200
lp = synth.flat(B, T, 1).cuda()
tg, tl = synth.targets(B, S, 2); tl[0] = S
il = torch.full((B,), T)
for _ in range(3):
    ops.ctc_loss_fwd_bwd(lp, il.cuda(), tg.cuda(), tl.cuda(), blank=1, mean_denominator=B, want_grad=want)
torch.cuda.synchronize()
ws = ops._WS[(0, "loss")]
NW, N, SL = 8, 256, 6
log = ws[-NW * N * SL * 8:].view(torch.int64).cpu().numpy().reshape(NW, N, SL)
names = ["MF", "MB", "LF", "LB", "W0", "W1", "W2", "W3"]
for w in range(NW):
    a = log[w]
    idx = np.where(a[:, 0] > 0)[0]
    if len(idx) < 3:
        continue
    st = a[idx][:, :5].astype(np.float64)
    nst = 5 if w >= 4 else 4
    d = [np.mean(st[1:-1, k + 1] - st[1:-1, k]) for k in range(nst - 1)]
    period = np.mean(np.diff(st[:, 0])[1:-1]) if len(idx) > 3 else float("nan")
    print(f"{names[w]}: {len(idx):3d} events  mean deltas {[round(x) for x in d]}  period {period:7.0f}  span {st[:, :nst].max() - st[:, 0].min():9.0f}")
