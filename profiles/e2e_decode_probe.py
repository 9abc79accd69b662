"""Where the host-buffer decode call spends its time (config 3): per-chunk kernel, D2H, string assembly."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from silent_speech_decoder_b200 import ops
from silent_speech_decoder_b200.data.vocab import Vocab
from silent_speech_decoder_b200.decoding import ctc
kind = sys.argv[1] if len(sys.argv) > 1 else "flat"
v = Vocab.default()
lp, ln = bench.decode_inputs(10000, 500, kind, 1003)
h_lp, h_ln = lp.pin_memory(), ln.pin_memory()
dec = ctc.build_decoder("beam", v, beam_width=50)
for chunk in (1776, 1184, 2368, 10**9):
    ctc.PIPELINE_CHUNK = chunk
    dec(h_lp, h_ln)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = dec(h_lp, h_ln)
    torch.cuda.synchronize()
    print(kind, "chunk", chunk, "e2e ms", round(1e3 * (time.perf_counter() - t0), 1), flush=True)
# pieces, unpipelined
x = h_lp.cuda(); torch.cuda.synchronize()
_, space = ctc._beam_label_strings(v)
text, _ = ctc._beam_label_strings(v)
t0 = time.perf_counter(); tok, cnt, _ = ops.beam_ids(x, h_ln, v.blank_id, v.pad_id, space, 50, -10.0); torch.cuda.synchronize(); t1 = time.perf_counter()
s = ctc._rows_to_strings(tok, cnt, text, True); t2 = time.perf_counter()
print("kernel ms", round(1e3 * (t1 - t0), 1), "strings+d2h ms", round(1e3 * (t2 - t1), 1), "chars", sum(map(len, s)), "max count", int(cnt.max()))
