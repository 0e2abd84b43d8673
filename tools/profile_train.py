#!/usr/bin/env python
"""Per-kernel time breakdown of the cfg3 training step (torch profiler, CUDA activities).  python tools/profile_train.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from fots.pytorch_b200.pipeline import FOTSNet  # noqa: E402
from fots.pytorch_b200.pipeline.train import TrainStep, synthetic_targets  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    torch.manual_seed(0)
    net = FOTSNet(attention=True, nclass=89).to(dev).to(memory_format=torch.channels_last)
    step = TrainStep(net, lr=1e-4)
    images = torch.randn(B, 3, 720, 1280, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    tgt = synthetic_targets(B, 64, 720, 1280, 89, dev, seed=0)
    for _ in range(2):
        step(images, tgt)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(images, tgt)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=100))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof2:
        step(images, tgt)
        torch.cuda.synchronize()
    rows = [e for e in prof2.key_averages(group_by_stack_n=8) if e.key in ("aten::item", "aten::_local_scalar_dense", "cudaMemcpyAsync", "aten::nonzero")]
    rows.sort(key=lambda e: -e.count)
    for e in rows[:6]:
        print(e.key, e.count, "\n    " + "\n    ".join(e.stack[:8]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step(images, tgt)
    e1.record()
    torch.cuda.synchronize()
    print("ms per training step of %d images: %.1f" % (B, e0.elapsed_time(e1) / 3))
