"""Small driver used for ncu captures: a few launches of one kernel on its bench workload.
    python profiles/prof_run.py loss|beam_peaked|beam_flat|beam_lm|greedy"""
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
elif what == "beam_lm":
    from silent_speech_decoder_b200.decoding import lm as plm

    v = Vocab.default()
    lp, ln, arpa = bench.lm_workload(2000, 500, 1003)
    text, space = ctc._beam_label_strings(v)
    keep = [i for i in range(v.size) if i not in {v.blank_id, v.pad_id}]
    tables = plm.build_tables(arpa, [""] + [text[i] for i in keep], bench.LM_CFG["alpha"], bench.LM_CFG["beta"])
    os.unlink(arpa)
    x, ln = lp.to(dev), ln.to(dev)
    for _ in range(reps):
        out = ops.beam_lm_ids(x, ln, v.blank_id, v.pad_id, space, tables, [(50, -10.0, 0.0, -5.0)])
    torch.cuda.synchronize()
    print(what, "counts", int(out[1].sum()))
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
