import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fots.pytorch_b200.pipeline import conv as TC
dev = torch.device("cuda:0")
xf = torch.randn(8, 3, 720, 1280, device=dev).contiguous(memory_format=torch.channels_last)
conv = torch.nn.Conv2d(3, 16, 3, 1, 1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
with torch.no_grad():
    for _ in range(4):
        TC.stem_conv_stats(xf, conv.weight)
torch.cuda.synchronize()
