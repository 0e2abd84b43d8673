#!/usr/bin/env python
"""Time of the folded-heads forms at the step's shapes: dwconv_up + heads_merged vs GEMM (f2 -> T) + heads_gather."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fots.pytorch_b200.pipeline import conv as TC  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
h, w, H, W = 90, 160, 180, 320
cl = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
f2, s3 = cl(B, 256, h, w), cl(B, 64, H, W)
gate = torch.rand(B, 1, h, w, device=dev).to(torch.bfloat16)
act, rbox, angle = torch.nn.Conv2d(256, 1, 1).to(dev), torch.nn.Conv2d(256, 4, 1).to(dev), torch.nn.Conv2d(256, 2, 1).to(dev)
pw, lat = torch.nn.Conv2d(256, 256, 1, bias=False).to(dev), torch.nn.Conv2d(64, 256, 1, bias=False).to(dev)
dw = torch.nn.Conv2d(256, 256, 3, 1, 1, groups=256, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
mh = TC.pack_merged_heads(act, rbox, angle, pw, lat)
a72 = TC.pack_gather_heads(act, rbox, angle, pw, dw)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


d = TC.dwconv_up(dw, f2, (H, W))
print("B=%d  dwconv_up %.1f us, heads_merged %.1f us" % (B, timed(lambda: TC.dwconv_up(dw, f2, (H, W))), timed(lambda: TC.heads_merged(d, s3, gate, mh))))
print("      conv f2 -> T %.1f us, conv + heads_gather %.1f us" % (
    timed(lambda: TC.conv2d(f2, a72)), timed(lambda: TC.heads_gather(f2, s3, gate, a72, mh))))
