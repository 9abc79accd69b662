"""Dump the per-role segment timeline of CTA 0 (SSD_LOSS_DBG=8) for the cfg2 loss launch."""
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
n = 8 * 512 * 2
log = ws[-n * 8:].view(torch.int64).cpu().numpy().reshape(4, 2, 512, 2)
t0 = log[log > 0].min()
names = ["M", "H", "L", "G"]
print("S of utterance 0:", int(tl[0]), "K from plan; all times in cycles since first stamp")
for h in range(2):
    for role in range(4):
        a = log[role, h]
        idx = np.where(a[:, 1] > 0)[0]
        if len(idx) == 0: continue
        start, end = a[idx, 0] - t0, a[idx, 1] - t0
        print(f"half {h} {names[role]}: {len(idx)} segments, first start {start[0]}, last end {end[-1]}")
        for i in list(idx[:4]) + list(idx[len(idx)//2-2:len(idx)//2+6]) + list(idx[-3:]):
            print(f"    idx {i:3d} start {a[i,0]-t0:8d} end {a[i,1]-t0:8d} dur {a[i,1]-a[i,0]:6d}")
