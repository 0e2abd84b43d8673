#!/usr/bin/env python
"""Eager (no CUDA graph) launches of one RoIRotate configuration, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:rroi_ -s 40 -c 3 -o gpurun_out/prof \
        python tools/profile_target.py --layout nhwc --channels 64 --variant 1 --steps 60
"""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from fots.pytorch_b200 import _cabi  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="nhwc")
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--images", type=int, default=1)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--cg", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=0)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--backward", type=int, default=0)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--nchw-tma", type=int, default=0)
    ap.add_argument("--rois-ready", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--bwd-mode", type=int, default=0)
    a = ap.parse_args()
    args = types.SimpleNamespace(channels=a.channels, layout=a.layout, images=a.images, rois_per_image=64,
                                 sets=0, dtype=a.dtype)
    dev = torch.device("cuda:0")
    wl = bench.Workload(args, dev, torch)
    if a.backward:
        wl.enable_backward(torch)
    wl.set_opts(_cabi, nchw_cg=a.cg, variant=a.variant, pdl=bool(a.pdl), nchw_tma=a.nchw_tma, rois_ready=bool(a.rois_ready),
                zero_chunk_images=a.chunk, bwd_mode=a.bwd_mode)
    st = torch.cuda.current_stream().cuda_stream
    for i in range(a.steps):
        wl.launch(i % wl.sets, _cabi.lib(), _cabi, st)
    torch.cuda.synchronize()
    print("done", a)
