"""Quick A/B of the loss launch: config 2 flat, config 2 "trained", B=1024 and B=4096 flat, a lone longest utterance.
    python profiles/loss_ab.py [tag]     (env knobs of ctc_loss_fast.cu apply: SSD_LOSS_NODEAL, SSD_LOSS_NC, ...)
Prints one JSON line; device time per ssd_ctc_loss call over back-to-back calls (bench.time_loss_launch)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from silent_speech_decoder_b200 import _native, synth

if os.environ.get("SSD_LIB"):  # a variant build (profiles/build_variant.sh)
    from pathlib import Path
    _native._LIB = _native.Library(Path(os.environ["SSD_LIB"]).resolve())

dev = torch.device("cuda", 0)
out = {"tag": sys.argv[1] if len(sys.argv) > 1 else "", "env": {k: v for k, v in os.environ.items() if k.startswith("SSD_")}}
lp, il, tg, tl = bench.cfg2_inputs()
sets = [lp.to(dev).clone() for _ in range(4)]
ms, nx = bench.time_loss_launch(sets, il.to(dev), tg.to(dev), tl.to(dev), 256, 100)
out["cfg2_flat_us"], out["cfg2_flat_exact"] = round(ms * 1e3, 1), nx
tlp, til, ttg, ttl = bench.cfg2_trained_inputs()
tsets = [tlp.to(dev).clone() for _ in range(4)]
ms, nx = bench.time_loss_launch(tsets, til.to(dev), ttg.to(dev), ttl.to(dev), 256, 100)
out["cfg2_trained_us"], out["cfg2_trained_exact"] = round(ms * 1e3, 1), nx
del tsets
# the longest utterance alone (S = 200): the latency floor of the launch
j = int(tl.argmax())
one = [lp[j:j + 1].to(dev).clone() for _ in range(2)]
ms, _ = bench.time_loss_launch(one, il[j:j + 1].to(dev), tg[j:j + 1].to(dev), tl[j:j + 1].to(dev), 1, 50)
out["lone_S200_us"] = round(ms * 1e3, 1)
for B in (1024, 4096):
    x = synth.flat(B, 1000, 5, 48).to(dev)
    tgb, tlb = synth.targets(B, 200, 6)
    ilb = torch.full((B,), 1000)
    ms, nx = bench.time_loss_launch([x], ilb.to(dev), tgb.to(dev), tlb.to(dev), B, 20)
    out[f"B{B}_us"], out[f"B{B}_frac"], out[f"B{B}_exact"] = round(ms * 1e3, 1), round(384 * B * 1000 / (ms * 1e-3) / 1e9 / 6543.7, 4), nx
    del x
print(json.dumps(out))
