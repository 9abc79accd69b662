"""Times the LM-fused beam kernel of one library build on the bench's LM workload (2000 utterances, T=500, width 50):
    SSD_LIB=profiles/scratch/NAME.so python profiles/beam_lm_probe.py [n_utts]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathlib import Path
import torch
import bench
from silent_speech_decoder_b200 import _native, ops
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc, lm as plm
if os.environ.get("SSD_LIB"):
    _native._LIB = _native.Library(Path(os.environ["SSD_LIB"]).resolve())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
v = Vocab.default()
lp, ln, arpa = bench.lm_workload(n, 500, 1003)
text, space = ctc._beam_label_strings(v)
keep = [i for i in range(v.size) if i not in {v.blank_id, v.pad_id}]
tables = plm.build_tables(arpa, [""] + [text[i] for i in keep], 0.6, 0.0)
os.unlink(arpa)
x, ln = lp.cuda(), ln.cuda()
res = []
for width in (50, 20, 100):
    st = [(width, -10.0, 0.0, -5.0)]
    for _ in range(2):
        out = ops.beam_lm_ids(x, ln, v.blank_id, v.pad_id, space, tables, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        out = ops.beam_lm_ids(x, ln, v.blank_id, v.pad_id, space, tables, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    res.append(f"w{width}: {ms:7.1f} ms {n / ms:6.1f} k utt/s (sum of counts {int(out[1].sum())})")
print(os.environ.get("SSD_LIB", "shipped"), " | ".join(res))
