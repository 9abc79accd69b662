"""Times the no-LM beam kernel of one library build at config 3 (10 000 utterances, T=500, width 50), confident and flat:
    SSD_LIB=profiles/scratch/NAME.so python profiles/beam_probe.py [n_utts]"""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
v = Vocab.default()
_, space = ctc._beam_label_strings(v)
res = []
for kind in ("peaked", "flat"):
    lp, ln = bench.decode_inputs(n, 500, kind, 1003)
    x, ln = lp.cuda(), ln.cuda()
    for width in (50,) if kind == "peaked" else (50, 1):
        for _ in range(3):  # clocks ramp up over the first launches after the host-side input generation
            out = ops.beam_ids(x, ln, v.blank_id, v.pad_id, space, width, -10.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = ops.beam_ids(x, ln, v.blank_id, v.pad_id, space, width, -10.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res.append(f"{kind} w{width}: {ms:6.1f} ms {n / ms:6.1f} k utt/s (checksum {int(out[1].sum())}, {float(out[2].sum()):.3f})")
    del x
print(os.environ.get("SSD_LIB", "shipped"), " | ".join(res))
