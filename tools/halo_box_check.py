#!/usr/bin/env python
"""Experiment: single-box halo (A descriptors at whole-pixel offsets inside the 128B-swizzled box) -- parity and speed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fots.pytorch_b200.pipeline import conv as TC

dev = torch.device("cuda:0")
with torch.no_grad():
    for (N, H, W) in ((1, 45, 80), (8, 180, 320)):
        x = torch.randn(N, 64, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = (torch.randn(64, 64, 3, 3, device=dev) / 24).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ref = F.conv2d(x.float(), w.float(), None, 1, 1)
        for mode in (1, 2, 0):
            TC.set_halo(mode)
            y = TC.conv2d(x, w, None, (1, 1), 1.0)
            torch.cuda.synchronize()
            err = float((y.float() - ref).abs().max())
            for _ in range(3):
                TC.conv2d(x, w, None, (1, 1), 1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                TC.conv2d(x, w, None, (1, 1), 1.0)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print("N=%d %dx%d halo mode %d: max err %.4f (ref max %.2f)  %.1f us  %.0f TF/s" % (
                N, H, W, mode, err, float(ref.abs().max()), us, 2.0 * N * H * W * 64 * 64 * 9 / us / 1e6), flush=True)
        TC.set_halo(-1)
