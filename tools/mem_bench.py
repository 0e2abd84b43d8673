#!/usr/bin/env python
"""Per-kernel time of the HBM-bound kernels of the 8-image step at the step's own shapes (A/B tool): us per launch and
algorithmic GB/s (bytes each kernel must read + write once) against the measured copy bandwidth.

    python tools/mem_bench.py [filter]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fots.pytorch_b200 import _cabi  # noqa: E402
if os.environ.get("SWEEP_LIB"):        # A/B against another build of the library on the same box (development only)
    _cabi.LIB_PATH = os.environ["SWEEP_LIB"]

from fots.pytorch_b200.pipeline import conv as TC, fused  # noqa: E402

dev = torch.device("cuda:0")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6553.0


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


def cl(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def report(name, us, nbytes):
    gbs = nbytes / us * 1e-3
    print("%-46s %8.2f us  %7.0f GB/s  %5.2f of copy peak" % (name, us, gbs, gbs / PEAK), flush=True)


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    cases = []
    # InstanceNorm: (name, B, C, H, W, residual, crelu)
    for name, B, C, H, W, res, crelu in [
            ("IN crelu0 apply 8x16x720x1280", 8, 16, 720, 1280, False, True), ("IN crelu1 8x32x360x640", 8, 32, 360, 640, False, True),
            ("IN layer1 8x64x180x320", 8, 64, 180, 320, False, False), ("IN layer1+res 8x64x180x320", 8, 64, 180, 320, True, False),
            ("IN layer2 8x128x90x160", 8, 128, 90, 160, False, False), ("IN layer2+res 8x128x90x160", 8, 128, 90, 160, True, False),
            ("IN layer3 8x256x45x80", 8, 256, 45, 80, False, False), ("IN layer3+res 8x256x45x80", 8, 256, 45, 80, True, False),
            ("IN layer4+res 8x512x23x40", 8, 512, 23, 40, True, False),
            ("IN batch5 512x128x8x64", 512, 128, 8, 64, False, False), ("IN batch7 512x256x4x64", 512, 256, 4, 64, False, False),
            ("IN batch10 512x256x1x64", 512, 256, 1, 64, False, False)]:
        if flt not in name:
            continue
        x = cl(B, C, H, W)
        r = cl(B, C, H, W) if res else None
        cout = 2 * C if crelu else C
        g, b = torch.randn(cout, device=dev), torch.randn(cout, device=dev)
        n = x.numel() * 2
        ws = fused.instnorm_stats(x)
        report(name + " [stats]", timed(lambda: fused.instnorm_stats(x)), n)
        report(name + " [apply]", timed(lambda: fused.instnorm_act(x, g, b, 1e-5, 0.01, r, crelu=crelu, stats=ws)),
               n * ((2 if res else 1) + (2 if crelu else 1)))
        if not crelu:
            report(name + " [whole]", timed(lambda: fused.instnorm_act(x, g, b, 1e-5, 0.01, r)), n * (4 if res else 3))
    # top-down merge
    for name, B, C, h, w, H, W in [("merge f4->f3 8x256 23x40->45x80", 8, 256, 23, 40, 45, 80),
                                   ("merge 8x256 45x80->90x160", 8, 256, 45, 80, 90, 160),
                                   ("merge 8x256 90x160->180x320", 8, 256, 90, 160, 180, 320)]:
        if flt not in name:
            continue
        lo, hi, c = cl(B, C, h, w), cl(B, C, H, W), cl(B, C, H, W)
        gate = torch.randn(B, 1, h, w, device=dev).to(torch.bfloat16)
        report(name + " [up(a)+b*gate]", timed(lambda: fused.fpn_merge(a_lo=lo, b_hi=hi, gate_logits_lo=gate)), lo.numel() * 2 + hi.numel() * 4)
        report(name + " [c+b*gate]", timed(lambda: fused.fpn_merge(c_hi=c, b_hi=hi, gate_logits_lo=gate)), hi.numel() * 6)
        conv = torch.nn.Conv2d(C, C, 3, 1, 1, groups=C, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        with torch.no_grad():
            report(name + " [dw(up(lo))]", timed(lambda: TC.dwconv_up(conv, lo, (H, W))), lo.numel() * 2 + hi.numel() * 2)
            report(name + " [dw(hi)]", timed(lambda: TC.dwconv(conv, hi)), hi.numel() * 4)
    # heads
    if flt in "heads":
        for H, W in ((90, 160), (180, 320)):
            x = cl(8, 256, H, W)
            act, rb, an = (torch.nn.Conv2d(256, k, 1).to(dev).to(torch.bfloat16) for k in (1, 4, 2))
            pk = TC.pack_heads(act, rb, an)
            p1 = TC.pack_to1(act)
            report("heads 8x256x%dx%d" % (H, W), timed(lambda: TC.heads(x, pk)), x.numel() * 2 + 8 * H * W * 7 * 4)
            report("to1   8x256x%dx%d" % (H, W), timed(lambda: TC.conv1x1_to1(x, p1)), x.numel() * 2 + 8 * H * W * 2)


if __name__ == "__main__":
    main()
