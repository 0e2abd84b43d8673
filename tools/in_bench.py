#!/usr/bin/env python
"""Single-pass cluster InstanceNorm vs the two-pass kernels, per shape of the step (A/B tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fots.pytorch_b200.pipeline import fused

dev = torch.device("cuda:0")
shapes = [("stage2 8x128x90x160", 8, 128, 90, 160, True), ("stage3 8x256x45x80", 8, 256, 45, 80, True),
          ("stage4 8x512x23x40", 8, 512, 23, 40, True), ("batch5 512x128x8x64", 512, 128, 8, 64, False),
          ("batch7 512x256x4x64", 512, 256, 4, 64, False), ("batch10 512x256x1x64", 512, 256, 1, 64, False),
          ("sep3 8x256x45x80 nores", 8, 256, 45, 80, False)]
for name, B, C, H, W, res in shapes:
    x = torch.randn(B, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = torch.randn_like(x) if res else None
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    out = []
    for single in (2, 0):
        fused.set_single_pass(single)
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fused.instnorm_act(x, g, b, 1e-5, 0.01, r)
        torch.cuda.synchronize()
        with torch.cuda.graph(gr):
            for _ in range(20):
                fused.instnorm_act(x, g, b, 1e-5, 0.01, r)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 200 * 1e3)
    fused.set_single_pass(1)
    mb = x.numel() * 2 * (3 if res else 2) / 1e6
    print("%-26s single-pass %7.2f us (%5.0f GB/s)   two-pass %7.2f us (%5.0f GB/s)" % (name, out[0], mb / out[0] * 1e3, out[1], mb / out[1] * 1e3), flush=True)
