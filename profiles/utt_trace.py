"""Per-utterance start/end times of one config-2 loss launch (build with SSD_NVCC_FLAGS=-DSSD_LOSS_UTTRACE, run with SSD_LOSS_DBG=8).
Prints, per strip width, count / start / end / duration statistics and the SM sharing pattern."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SSD_LOSS_DBG"] = "8"
import torch, numpy as np
import bench
from silent_speech_decoder_b200 import _native, ops
dev = torch.device("cuda", 0)
lp, il, tg, tl = bench.cfg2_trained_inputs() if "trained" in sys.argv else bench.cfg2_inputs()
lib = _native.load()
B, T, V = lp.shape
x = lp.to(dev); il_d, tg_d, tl_d = il.to(dev), tg.to(dev), tl.to(dev)
nll = torch.zeros(B, device=dev); grad = torch.empty_like(x)
nbytes = int(lib.ssd_ctc_loss_workspace_bytes(B, T, V, tg.size(1)))
ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
for _ in range(3):
    bench._loss_abi_call(lib, ops, x, il_d, tg_d, tl_d, nll, grad, ws, B, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
# the fast workspace sits behind the exact one; the trace is its last 8 * 256 * 6 * 8 bytes
tlog_bytes = 8 * 256 * 6 * 8
t = ws[nbytes - tlog_bytes:].view(torch.int64)[: 4 * B].cpu().numpy().reshape(B, 4)
t0 = t[:, 0].min()
st, en, sm, nc = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2], t[:, 3]
print(f"launch span {en.max():.1f} us")
for w in sorted(set(nc.tolist())):
    m = nc == w
    print(f"width {w:2d}: n {m.sum():3d}  start {st[m].min():6.1f}..{st[m].max():6.1f}  end {en[m].min():6.1f}..{en[m].max():6.1f}  dur mean {np.mean(en[m]-st[m]):6.1f} max {np.max(en[m]-st[m]):6.1f}")
from collections import Counter
per_sm = Counter(sm.tolist())
print("SMs used", len(per_sm), "utterances per SM histogram", sorted(Counter(per_sm.values()).items()))
pairs = Counter()
for s_ in per_sm:
    ws_ = tuple(sorted(nc[sm == s_].tolist()))
    pairs[ws_] += 1
print("width sets per SM:", sorted(pairs.items(), key=lambda kv: -kv[1])[:20])
late = np.argsort(-en)[:8]
print("last to finish:", [(int(nc[i]), int(sm[i]), round(float(st[i]), 1), round(float(en[i]), 1)) for i in late])
