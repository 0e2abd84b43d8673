#!/usr/bin/env python
"""A/B of the persistent depthwise kernels at the step's shapes (SWEEP_LIB=<other build of the library> for the B side)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fots.pytorch_b200 import _cabi  # noqa: E402
if os.environ.get("SWEEP_LIB"):
    _cabi.LIB_PATH = os.environ["SWEEP_LIB"]
from fots.pytorch_b200.pipeline import conv as TC, fused  # noqa: E402

dev = torch.device("cuda:0")
cl = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def timed(fn, reps=20):
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


with torch.no_grad():
    for B in (8, 32):
        for (C, H, W) in ((256, 45, 80), (512, 23, 40), (256, 90, 160)):
            conv = torch.nn.Conv2d(C, C, 3, 1, 1, groups=C, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
            norm = torch.nn.InstanceNorm2d(C, affine=True).to(dev)
            x = cl(B, C, H, W)
            st = fused.instnorm_stats(x)
            lo = cl(B, C, (H + 1) // 2, (W + 1) // 2)
            a = timed(lambda: TC.dwconv(conv, x))
            b = timed(lambda: TC.dwconv_norm(conv, x, st, norm, 0.01, stats_out=True))
            c = timed(lambda: TC.dwconv_up(conv, lo, (H, W)))
            print("B=%-2d %3dx%3dx%3d  plain %7.2f us  norm+stats %7.2f us  up %7.2f us" % (B, C, H, W, a, b, c), flush=True)
