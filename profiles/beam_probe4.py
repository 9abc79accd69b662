"""Config 4 point (256 utterances, T=4000) for one library build:  SSD_LIB=... python profiles/beam_probe4.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathlib import Path
import torch
import bench
from silent_speech_decoder_b200 import _native, ops
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc
if os.environ.get("SSD_LIB"):
    _native._LIB = _native.Library(Path(os.environ["SSD_LIB"]).resolve())
v = Vocab.default()
_, space = ctc._beam_label_strings(v)
res = []
for kind in ("peaked", "flat"):
    lp, ln = bench.decode_inputs(256, 4000, kind, bench.CFG4["seed"])
    x, ln = lp.cuda(), ln.cuda()
    for width in (1, 50, 200):
        for _ in range(2):
            out = ops.beam_ids(x, ln, v.blank_id, v.pad_id, space, width, -10.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            out = ops.beam_ids(x, ln, v.blank_id, v.pad_id, space, width, -10.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        res.append(f"{kind} w{width}: {256 / ms:5.2f} k")
print(os.environ.get("SSD_LIB", "shipped"), " | ".join(res))
