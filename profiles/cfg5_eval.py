"""Config 5: evaluation of a silent_parallel_data-sized test subset, decoder + scoring only.

159 utterances x SCALE, post-subsampling lengths T ~ U[100,600] (2-12 s at 100 Hz, subsample 2), peaked
synthetic posteriors standing in for the encoder output (the encoder is PyTorch and outside this path);
beam-50 decode of device-resident log-probs, then WER/CER through ssd_edit_counts, in batches of BATCH
utterances (the reference: batch 4, `.cpu()` per batch, evaluate.py:158,299).  References are the greedy
transcripts of a noisier copy, so the error rates are non-trivial.  Prints one JSON line.
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from silent_speech_decoder_b200 import ops, synth  # noqa: E402
from silent_speech_decoder_b200.data.vocab import Vocab  # noqa: E402
from silent_speech_decoder_b200.decoding import ctc  # noqa: E402
from silent_speech_decoder_b200.evaluation import evaluate as ev  # noqa: E402

SCALE = int(sys.argv[1]) if len(sys.argv) > 1 else 64
BATCH = int(sys.argv[2]) if len(sys.argv) > 2 else 2544
N, T = 159 * SCALE, 600
V = Vocab.default()
g = torch.Generator().manual_seed(1005)
lengths = torch.randint(100, T + 1, (N,), generator=g)
# B distinct utterances tiled to N (generating 10k peaked rows on the host is the slow part, not the subject)
base = synth.peaked(159, T, 1005)
noisy = torch.log_softmax(base + 2.0 * torch.randn(base.shape, generator=g), -1)
ref159 = [t if t.strip() else "x" for t in ctc.build_decoder("greedy", V)(noisy, torch.full((159,), T))]
lp = base.repeat(SCALE, 1, 1).cuda()
refs = ref159 * SCALE
dec = ctc.build_decoder("beam", V, beam_width=50)


def run():
    e = ev.Evaluator(dec, {"type": "beam", "beam_width": 50})
    for lo in range(0, N, BATCH):
        hi = min(N, lo + BATCH)
        e.add(range(lo, hi), refs[lo:hi], lp[lo:hi], lengths[lo:hi].cuda())
    return e.metrics()


run()
torch.cuda.synchronize()
t0 = time.perf_counter()
before = dict(ops.LAUNCHES)
m = run()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
# scoring alone
t1 = time.perf_counter()
e2 = ev.compute_metrics(refs, [r for r in refs])
dt_score = time.perf_counter() - t1
print(json.dumps({"workload": f"cfg5: {N} utterances (159 x {SCALE}), T~U[100,600], peaked, beam-50 + WER/CER, batch {BATCH}",
                  "utterances_per_s": N / dt, "seconds": dt, "wer": m["wer"], "cer": m["cer"],
                  "launches": {k: ops.LAUNCHES[k] - before[k] for k in before}, "score_only_seconds": dt_score}))
