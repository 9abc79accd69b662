"""Which utterances does the fp32 fast path hand to the exact kernel, and why (reason bits of FastParams::flags)?
    python profiles/flag_probe.py S_max [SSD_LOSS_NC=..]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from silent_speech_decoder_b200 import _native, ops, synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
B, T, V = 256, 1000, 48
g = torch.Generator().manual_seed(5)
tl = torch.randint(S // 4, S + 1, (B,), generator=g, dtype=torch.int64)
tg = torch.randint(2, V, (B, S), generator=g, dtype=torch.int64) * (torch.arange(S)[None, :] < tl[:, None])
il = torch.full((B,), T, dtype=torch.int64)
lib = _native.load()
off = lib.ssd_ctc_loss_fallback_counter_offset(B, T, V, S)
for k in range(4):
    lp = synth.flat(B, T, 7 + k, V).to(dev)
    nll, grad = ops.ctc_loss_fwd_bwd(lp, il.to(dev), tg.to(dev), tl.to(dev), blank=1, mean_denominator=B)
    torch.cuda.synchronize()
    ws = ops._WS[(0, "loss")]
    flags = ws[off - 4 + 256: off - 4 + 256 + 4 * B].view(torch.int32).cpu()
    bad = flags.nonzero().flatten().tolist()
    print("set", k, "flagged", [(b, int(flags[b]), int(tl[b])) for b in bad])
