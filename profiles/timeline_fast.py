"""Per-role block timeline of CTA 0's first utterance in the fp32 fast loss kernel (SSD_LOSS_DBG=8).
Needs the stamped build: SSD_NVCC_FLAGS=-DSSD_LOSS_TIMELINE python -m silent_speech_decoder_b200.build --force"""
import os, sys
os.environ["SSD_LOSS_DBG"] = "8"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from silent_speech_decoder_b200 import ops
lp, il, tg, tl = bench.cfg2_inputs()
x, il, tg, tl = lp.cuda(), il.cuda(), tg.cuda(), tl.cuda()
for _ in range(2):
    ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=256)
torch.cuda.synchronize()
ws = ops._WS[(0, "loss")]
N, S = 256, 6
NW = 8  # warps per CTA of the default worker shape (4 workers)
log = ws[-NW * N * S * 8:].view(torch.int64).cpu().numpy().reshape(NW, N, S)
t0 = log[:, :, :5][log[:, :, :5] > 0].min()
print("utterance 0: S =", int(tl[0]), " cycles since first stamp")
names = ["MF", "MB", "LF", "LB", "W0", "W1", "W2", "W3"]  # role map of an unrotated CTA
for w in range(NW):
    a = log[w]
    idx = np.where(a[:, 0] > 0)[0]
    if len(idx) == 0:
        continue
    print(f"{names[w]}: {len(idx)} events, first {a[idx[0], 0] - t0}, last {a[idx[-1], :5].max() - t0}")
    sel = list(idx[:3]) + list(idx[len(idx) // 2: len(idx) // 2 + 3]) + list(idx[-3:])
    for i in sel:
        st = [int(v - t0) if v > 0 else -1 for v in a[i, :5]]
        d = [st[k + 1] - st[k] if st[k + 1] >= 0 and st[k] >= 0 else None for k in range(4)]
        print(f"   ev {i:3d} q={a[i, 5] if w >= 4 else '-'} stamps {st}  deltas {d}")
