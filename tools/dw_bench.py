#!/usr/bin/env python
"""Depthwise 3x3 kernel and fused detection heads vs the library path, per shape of the 8-image step (A/B tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fots.pytorch_b200.pipeline import conv as TC

dev = torch.device("cuda:0")


def bench(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


with torch.no_grad():
    for name, N, C, H, W, st in [("sep3 first 128ch s2 90x160", 8, 128, 90, 160, 2), ("sep3 256ch 45x80", 8, 256, 45, 80, 1),
                                 ("sep4 first 256ch s2", 8, 256, 45, 80, 2), ("sep4 512ch 23x40", 8, 512, 23, 40, 1),
                                 ("upconv1 256ch 90x160", 8, 256, 90, 160, 1), ("upconv2 256ch 180x320", 8, 256, 180, 320, 1)]:
        conv = torch.nn.Conv2d(C, C, 3, st, 1, groups=C, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        x = torch.randn(N, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ours, lib = bench(lambda: TC.dwconv(conv, x)), bench(lambda: conv(x))
        mb = (x.numel() + x.numel() // (st * st)) * 2 / 1e6
        print("dw %-28s ours %7.2f us (%5.0f GB/s)   cuDNN %7.2f us" % (name, ours, mb / ours * 1e3, lib), flush=True)
    for H, W in ((180, 320), (90, 160)):
        mk = lambda co: torch.nn.Conv2d(256, co, 1, bias=True).to(dev).to(torch.bfloat16)
        act, rbox, angle = mk(1), mk(4), mk(2)
        x = torch.randn(8, 256, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        pk = TC.pack_heads(act, rbox, angle)

        def lib_heads():
            seg = torch.sigmoid(act(x).float())
            rb = torch.sigmoid(rbox(x).float()) * 128
            a = torch.sigmoid(angle(x).float()) * 2 - 1
            return seg, rb, a / torch.sqrt((a * a).sum(1, keepdim=True))

        ours, lib = bench(lambda: TC.heads(x, pk)), bench(lib_heads)
        print("heads 8x256x%dx%d   ours %7.2f us (%5.0f GB/s)   library path %7.2f us" % (H, W, ours, x.numel() * 2 / 1e6 / ours * 1e3, lib), flush=True)
