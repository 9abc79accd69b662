"""Forward-only (grad=NULL) time of the loss fast path at config 2: the two main sweeps over all T frames with the
workers doing a single block -- i.e. what the serial chains cost when nothing competes with them."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from silent_speech_decoder_b200 import ops
lp, il, tg, tl = bench.cfg2_inputs()
x, il, tg, tl = lp.cuda(), il.cuda(), tg.cuda(), tl.cuda()
for want in (False, True):
    for _ in range(3): ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=256, want_grad=want)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=256, want_grad=want)
    e1.record(); torch.cuda.synchronize()
    print("want_grad", want, "us", 1e3*e0.elapsed_time(e1)/20)
