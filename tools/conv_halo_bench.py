#!/usr/bin/env python
"""Halo-reuse conv_tc vs per-tap conv_tc vs bare cuDNN on the backbone's 3x3 stride-1 shapes (A/B tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fots.pytorch_b200.pipeline import conv as TC

dev = torch.device("cuda:0")


def t(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


with torch.no_grad():
    for name, N, H, W, cin, cout in [("layer0_1[0] 64->64 360x640", 8, 360, 640, 64, 64), ("layer1 64->64 180x320", 8, 180, 320, 64, 64),
                                     ("layer2 128->128 90x160", 8, 90, 160, 128, 128), ("conv5 64->128 512x8x64", 512, 8, 64, 64, 128),
                                     ("conv6 128->128 512x8x64", 512, 8, 64, 128, 128)]:
        x = torch.randn(N, cin, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gf = 2.0 * N * H * W * cout * cin * 9 / 1e6
        res = []
        for mode in (1, 0):
            TC.set_halo(mode)
            res.append(t(lambda: TC.conv2d(x, w, None, (1, 1), 1.0)))
        TC.set_halo(-1)
        lib = t(lambda: F.conv2d(x, w, None, 1, 1))
        print("%-28s halo %7.1f us %6.0f TF/s | per-tap %7.1f us %6.0f TF/s | cuDNN bare %7.1f us %6.0f TF/s" % (
            name, res[0], gf / res[0], res[1], gf / res[1], lib, gf / lib), flush=True)
