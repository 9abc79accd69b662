import os, sys
sys.path.insert(0, '/root/repo')
import torch
from silent_speech_decoder_b200 import ops, synth
dev = torch.device("cuda", 0)
B, T, V = 256, 1000, 48
for S in (60, 127, 191, 200, 255):
    g = torch.Generator().manual_seed(5)
    tl = torch.randint(S // 4, S + 1, (B,), generator=g, dtype=torch.int64)
    tg = torch.randint(2, V, (B, S), generator=g, dtype=torch.int64) * (torch.arange(S)[None, :] < tl[:, None])
    il = torch.full((B,), T, dtype=torch.int64)
    lps = [synth.flat(B, T, 7 + k, V).to(dev) for k in range(4)]
    il_d, tg_d, tl_d = il.to(dev), tg.to(dev), tl.to(dev)
    for i in range(5): ops.ctc_loss_fwd_bwd(lps[i % 4], il_d, tg_d, tl_d, blank=1, mean_denominator=B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): ops.ctc_loss_fwd_bwd(lps[i % 4], il_d, tg_d, tl_d, blank=1, mean_denominator=B)
    e1.record(); torch.cuda.synchronize()
    print(f"S_max {S:3d} (NC {2*((2*S+1+63)//64) if False else ''}): {e0.elapsed_time(e1)/40*1e3:.1f} us  fallbacks {ops.loss_fallback_count(B,T,V,S)}")
