"""Random-shape check of ssd_ctc_loss on the GPU: the shipped path (fp32 fast kernel + fp64 fallback) against the exact
fp64 kernel alone (SSD_LOSS_EXACT=1) of the same library, with a per-case watchdog so that a hang is reported as such.
    timeout 600 python profiles/fuzz_loss.py [cases] [seed]"""
import os, signal, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from silent_speech_decoder_b200 import ops, synth


class Hang(Exception):
    pass


def on_alarm(*_):
    raise Hang()


signal.signal(signal.SIGALRM, on_alarm)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst_n = worst_g = 0.0
for c in range(cases):
    B = int(rng.choice([1, 2, 3, 7, 33, 150, 297, 600, 900]))
    T = int(rng.choice([1, 5, 8, 9, 40, 97, 160, 200, 333, 600]))
    S = int(min(rng.choice([1, 2, 10, 40, 63, 64, 100, 127, 128, 200, 255]), max(1, T)))
    kind = rng.choice(["flat", "peaked", "mild", "trained"])
    if kind == "trained" and T >= 8:  # confident posteriors on the utterances' own transcripts
        lp, tg, tl = synth.peaked_aligned(B, T, 7000 + c, n_labels=(1, max(1, min(S, T // 2))))
        if tg.size(1) < S:
            tg = torch.nn.functional.pad(tg, (0, S - tg.size(1)))
    else:
        lp = synth.flat(B, T, 7000 + c) if kind == "flat" else synth.peaked(B, T, 7000 + c, boost=8.0 if kind == "peaked" else 2.0)
        tg, tl = synth.targets(B, S, 8000 + c, s_min=1)
    il = synth.ragged_lengths(B, T, 9000 + c)
    if B > 2:
        il[1] = 0
        tl[2] = 0
    grid = int(rng.choice([0, 0, 1, 5, 100]))
    env = {"SSD_LOSS_GRID": str(grid)} if grid else {}
    if rng.random() < 0.3:
        env["SSD_LOSS_NOLATE"] = "1"
    if rng.random() < 0.5:
        env["SSD_LOSS_MULTI"] = str(int(rng.integers(0, 2)))  # one kernel per strip width / one launch-wide width
    os.environ.update(env)
    tag = f"case {c}: B={B} T={T} S={S} {kind} {env}"
    try:
        signal.alarm(20)
        x = lp.cuda()
        nll, g = ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=B)
        torch.cuda.synchronize()
        nfb = ops.loss_fallback_count(B, T, 48, tg.shape[1])
        for k in env:
            del os.environ[k]
        os.environ["SSD_LOSS_EXACT"] = "1"
        nll_x, g_x = ops.ctc_loss_fwd_bwd(x, il, tg, tl, blank=1, mean_denominator=B)
        torch.cuda.synchronize()
        del os.environ["SSD_LOSS_EXACT"]
        signal.alarm(0)
    except Hang:
        print("HANG", tag, flush=True)
        os._exit(3)
    fin = torch.isfinite(nll_x)
    assert torch.equal(torch.isfinite(nll), fin), tag
    # the exact kernel itself against the reference's ATen op in fp64 on the host (bounded sizes: seconds of CPU)
    et = 0.0
    if B * T <= 120000:
        from oracle import ctc_loss as ol

        n64, _, g64 = ol.torch_ctc(lp, il, tg, tl, 1, True, torch.float64)
        f64 = torch.isfinite(n64) & (n64 != 0)
        assert torch.isfinite(g_x).all(), tag + " (non-finite gradient from the exact path)"
        if f64.any():
            et = float(((nll_x.cpu().double() - n64)[f64].abs() / n64[f64].abs().clamp_min(1e-3)).max())
        et = max(et, float((g_x.cpu().double() - g64).abs().max() / g64.abs().max().clamp_min(1e-30)))
        if et > 1e-5:
            print("BAD " + tag, f"exact path vs torch fp64: {et:.1e}", flush=True)
            os._exit(2)
    en = float(((nll[fin] - nll_x[fin]).abs() / nll_x[fin].abs().clamp_min(1e-3)).max()) if fin.any() else 0.0
    eg = float((g - g_x).abs().max() / g_x.abs().max().clamp_min(1e-30))
    worst_n, worst_g = max(worst_n, en), max(worst_g, eg)
    ok = en <= 1e-5 and eg <= 1e-5  # the tolerance the scope contract states
    print(("ok  " if ok else "BAD ") + tag, f"fallbacks={nfb} nll_rel={en:.1e} grad_rel={eg:.1e}", flush=True)
    if not ok:
        os._exit(2)
print(f"all {cases} cases ok; worst nll_rel={worst_n:.1e} grad_rel={worst_g:.1e}")
