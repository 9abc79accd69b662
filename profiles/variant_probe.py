"""Times one loss-kernel build variant (a libssd_ctc.so built with extra -D flags) at config-2 shape:
    SSD_LIB=path/to/lib.so python profiles/variant_probe.py
B=1 (single-CTA latency, nothing else on the GPU) and B=256, forward-only and with the gradient; also checks the
gradient against the shipped exact kernel (SSD_LOSS_EXACT=1 is read per call by the library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathlib import Path
import torch
import bench
from silent_speech_decoder_b200 import _native, ops, synth
if os.environ.get("SSD_LIB"):
    _native._LIB = _native.Library(Path(os.environ["SSD_LIB"]).resolve())
T, S = 1000, 200
out = []
LP, IL, TG, TL = bench.cfg2_inputs()
for B in (1, 256):
    if B == 1:  # the longest target of config 2, alone on the GPU
        j = int(TL.argmax())
        lp, il, tg, tl = LP[j:j + 1].cuda(), IL[j:j + 1].cuda(), TG[j:j + 1].cuda(), TL[j:j + 1].cuda()
    else:
        lp, il, tg, tl = LP.cuda(), IL.cuda(), TG.cuda(), TL.cuda()
    for want in (False, True):
        for _ in range(3): r = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B, want_grad=want)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B, want_grad=want)
        e1.record(); torch.cuda.synchronize()
        out.append(f"B={B} grad={int(want)} {1e3 * e0.elapsed_time(e1) / 10:6.1f}us")
    nll, g = r
    nfb = ops.loss_fallback_count(B, T, 48, tg.shape[1])
    os.environ["SSD_LOSS_EXACT"] = "1"
    nll2, g2 = ops.ctc_loss_fwd_bwd(lp, il, tg, tl, blank=1, mean_denominator=B)
    del os.environ["SSD_LOSS_EXACT"]
    out.append(f"err={float((g - g2).abs().max() / g2.abs().max()):.1e} fb={nfb}")
print(os.environ.get("SSD_LIB", "shipped"), " | ".join(out))
