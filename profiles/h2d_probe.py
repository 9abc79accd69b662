"""Raw pinned host->device bandwidth for the config-2 log-prob batch (49 MB), alone and while the loss kernel runs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from silent_speech_decoder_b200 import ops
dev = torch.device("cuda", 0)
lp, il, tg, tl = bench.cfg2_inputs()
h = lp.pin_memory()
d = torch.empty_like(lp, device=dev)
cs = torch.cuda.Stream()
for label, busy in (("alone", False), ("with the loss kernel running", True)):
    x = lp.to(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record()
        for _ in range(20):
            d.copy_(h, non_blocking=True)
        e1.record()
    if busy:
        for _ in range(60):
            ops.ctc_loss_fwd_bwd(x, il.to(dev), tg.to(dev), tl.to(dev), blank=1, mean_denominator=256)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D {h.numel() * 4 / 1e6:.1f} MB {label}: {ms:.3f} ms = {h.numel() * 4 / ms / 1e6:.1f} GB/s")
