#!/usr/bin/env python
"""Per-kernel time breakdown of the end-to-end inference step (torch profiler, CUDA activities)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from fots.pytorch_b200.pipeline import FOTSNet, FOTSPipeline  # noqa: E402
from fots.pytorch_b200.pipeline.infer import planted_quads  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = FOTSNet(attention=True, nclass=89).to_b200(dev, inference=True)
    pipe = FOTSPipeline(net, 8, 64, 0.25, amp_dtype=torch.bfloat16)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    images = torch.randint(0, 256, (B, 720, 1280, 3), device=dev, dtype=torch.uint8).permute(0, 3, 1, 2)   # raw uint8, as bench.py feeds it
    quads = torch.from_numpy(planted_quads(B, 64)).to(dev)
    for _ in range(3):
        pipe.step_local(images, quads)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            pipe.step_local(images, quads)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pipe.step_local(images, quads)
    e1.record()
    torch.cuda.synchronize()
    print("ms per step of %d images: %.2f" % (B, e0.elapsed_time(e1) / 5))
    g = pipe.capture(images, quads, micro=8)
    g()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g()
    e1.record()
    torch.cuda.synchronize()
    print("ms per step of %d images, CUDA graph: %.2f" % (B, e0.elapsed_time(e1) / 5))
