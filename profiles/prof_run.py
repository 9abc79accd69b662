"""Small driver used for ncu captures: a few launches of one kernel on its bench workload.
    python profiles/prof_run.py loss|beam_peaked|beam_flat|greedy"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from silent_speech_decoder_b200 import ops
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc

what = sys.argv[1] if len(sys.argv) > 1 else "loss"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
if what == "loss":
    lp, il, tg, tl = bench.cfg2_inputs()
    x, il, tg, tl = lp.to(dev), il.to(dev), tg.to(dev), tl.to(dev)
    for _ in range(reps):
        nll, g = ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=256)
    torch.cuda.synchronize()
    print("nll mean", float(nll.mean()), "grad abs max", float(g.abs().max()))
else:
    v = Vocab.default()
    kind = "flat" if what.endswith("flat") else "peaked"
    n = 2000
    lp, ln = bench.decode_inputs(n, 500, kind, 1003)
    x, ln = lp.to(dev), ln.to(dev)
    _, space = ctc._beam_label_strings(v)
    for _ in range(reps):
        if what.startswith("beam"):
            out = ops.beam_ids(x, ln, v.blank_id, v.pad_id, space, 50, -10.0)
        else:
            out = ops.greedy_ids(x, ln, v.blank_id)
    torch.cuda.synchronize()
    print(what, "counts", int(out[1].sum()))
