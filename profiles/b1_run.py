"""One utterance (the longest target of config 2) alone on the GPU, a few launches: the subject of the single-CTA
ncu capture.   python profiles/b1_run.py [grad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from silent_speech_decoder_b200 import ops
LP, IL, TG, TL = bench.cfg2_inputs()
j = int(TL.argmax())
lp, il, tg, tl = LP[j:j + 1].cuda(), IL[j:j + 1].cuda(), TG[j:j + 1].cuda(), TL[j:j + 1].cuda()
want = len(sys.argv) > 1
for _ in range(4):
    nll, g = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=1, want_grad=want)
torch.cuda.synchronize()
print("nll", float(nll[0]))
