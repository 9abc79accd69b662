"""Loss fast path, time per launch against the number of co-resident CTAs (B utterances, one CTA each):
forward-only (the two serial sweeps alone) and with the gradient.  T=1000, S_max=200 (NC=14)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from silent_speech_decoder_b200 import ops, synth
T, S = 1000, 200
for B in (1, 16, 148, 256, 296):
    lp = synth.flat(B, T, 1).cuda()
    tg, tl = synth.targets(B, S, 2)
    tl[0] = S
    il = torch.full((B,), T)
    tg, tl, il = tg.cuda(), tl.cuda(), il.cuda()
    for want in (False, True):
        for _ in range(3): ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B, want_grad=want)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B, want_grad=want)
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 10
        print(f"B={B:4d} grad={int(want)}  {us:7.1f} us  = {us * 1.965:6.0f} cycles/frame-step at 1965 MHz (per 1000 steps)")
