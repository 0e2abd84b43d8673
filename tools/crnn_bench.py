#!/usr/bin/env python
"""Consumer B (CRNN, tools/models.py:853-909) on 64 crops of 32x256: torch fp32 / torch bf16 autocast / to_b200()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fots.pytorch_b200.pipeline import CRNN


def t(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = CRNN(nclass=7500).to(dev).eval()
    x = torch.randn(64, 3, 32, 256, device=dev)
    xc = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        print("CRNN 64 x 3x32x256 (207 GFLOP): torch fp32 %.2f ms" % t(lambda: net(x)))
        netc = CRNN(nclass=7500).to(dev).eval().to(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            print("torch bf16 autocast channels-last %.2f ms" % t(lambda: netc(xc)))
        net.to_b200(dev)
        print("to_b200 (folded BN, tcgen05 conv + bias + ReLU) %.2f ms" % t(lambda: net(x)))
        print("  cnn part only %.2f ms" % t(lambda: net._cnn_b200(x)))
