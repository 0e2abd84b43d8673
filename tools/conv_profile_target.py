#!/usr/bin/env python
"""A few eager launches of the tcgen05 convolution on the recogniser's and the backbone's shapes, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 4 -o gpurun_out/prof_conv \
        python tools/conv_profile_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fots.pytorch_b200.pipeline import conv as TC  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    shapes = [(512, 4, 64, 256, 256), (512, 8, 64, 128, 128), (8, 180, 320, 64, 64), (8, 90, 160, 128, 128)]
    for rep in range(2):          # first pass = warm-up (-s 4 skips it)
        for (N, H, W, Cin, Cout) in shapes:
            x = torch.randn(N, Cin, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            TC.conv2d(x, w, None, (1, 1), 0.01)
    torch.cuda.synchronize()
    print("done")
