import sys
sys.path.insert(0, '/root/repo')
import torch
from silent_speech_decoder_b200 import ops, synth
lp = synth.flat(8, 200, 3).cuda()
tg, tl = synth.targets(8, 200, 4)
il = torch.full((8,), 200)
for _ in range(2):
    nll, g = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=8)
torch.cuda.synchronize()
print("ok", nll[:3].tolist(), ops.loss_fallback_count(8, 200, 48, 200))
