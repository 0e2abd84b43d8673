#!/usr/bin/env python
"""GPU tuning sweep for the RoIRotate forward/backward kernels (development tool, run under gpurun).

    python tools/sweep.py fwd|bwd|bf16|nchw [...] > gpurun_out/sweep.txt

Same timing harness as bench.py (rotating buffer sets > L2, one CUDA graph per step, CUDA events); every point passes
its kernel choice through the per-call rroi_b200_opts.  Prints one line per point and appends to gpurun_out/sweep.json.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from fots.pytorch_b200 import _cabi  # noqa: E402

RECS = []
if os.environ.get("SWEEP_LIB"):        # A/B against another build of the library on the same box (development only)
    _cabi.LIB_PATH = os.environ["SWEEP_LIB"]


def point(layout="nhwc", C=64, images=1, streams=1, backward=False, dtype="fp32", xform=False, target=2000, steps=8, **opts):
    args = types.SimpleNamespace(channels=C, layout=layout, images=images, rois_per_image=64, sets=0, dtype=dtype)
    dev = torch.device("cuda:0")
    wl = bench.Workload(args, dev, torch)
    wl.set_opts(_cabi, **opts)
    lib = _cabi.lib()
    if xform:
        wl.enable_xform(torch, lib, torch.cuda.current_stream().cuda_stream)
    if backward:
        wl.enable_backward(torch)
    torch.cuda.synchronize()
    per_step = bench.launches_per_step_for(wl, target=max(8, target // (images * (4 if backward else 1))))
    ms = bench.timed_steps(wl, steps, 3, per_step, torch, lib, _cabi, lambda: None, streams)
    us = ms / (steps * per_step) * 1e3
    alg = float(np.mean(wl.alg_bytes_bwd if backward else wl.alg_bytes))
    peak, _ = bench.measured_peak_gbs()
    gbs = alg / us / 1e3
    rec = dict(layout=layout, C=C, images=images, rois=wl.N, streams=streams, backward=backward, dtype=dtype, xform=xform,
               opts=opts, us_per_launch=us, alg_mb=alg / 1e6, gbs=gbs, frac=gbs / peak)
    RECS.append(rec)
    print("%s %-4s %-4s C=%-3d img=%-2d S=%d xf=%d %-60s | %8.2f us  %7.1f GB/s  frac %.3f" % (
        "BWD" if backward else "FWD", layout, dtype, C, images, streams, int(xform), json.dumps(opts, sort_keys=True), us, gbs, gbs / peak),
        flush=True)
    del wl
    torch.cuda.empty_cache()
    return rec


def sweep_fwd():
    for S in (1, 8):
        for v in (1, 6, 5, 4, 11, 17, 12, 13, 16, 14, 15):
            for early in (False, True):
                point(streams=S, variant=v, rois_ready=early)
        point(streams=S, variant=12, rois_ready=True, xform=True)
        point(streams=S, variant=13, rois_ready=True, xform=True)
        point(streams=S, variant=5, rois_ready=True, xform=True)
        point(streams=S, variant=12, pdl=False)
        point(streams=S)                                   # auto, no hints
        point(streams=S, concurrency=S, rois_ready=True)   # auto with hints
    for S in (2, 4):
        for v in (5, 12, 13, 14):
            point(streams=S, variant=v, rois_ready=True)
    for v in (5, 4, 14, 15, 12):                           # large launches
        point(images=8, variant=v)
        point(images=32, variant=v)
    point(images=32, variant=5, xform=True)
    for v in (1, 5, 12, 13, 14):
        point(C=256, streams=1, variant=v, rois_ready=True)
        point(C=256, streams=8, variant=v, rois_ready=True)


def sweep_bwd():
    for chunk in (-1, 0, 1, 2, 3, 4, 6, 8):
        point(images=32, backward=True, zero_chunk_images=chunk)
    point(images=32, backward=True, zero_chunk_images=-1, bwd_mode=2)
    point(images=32, backward=True, zero_chunk_images=2, bwd_mode=1)
    for chunk in (-1, 0, 2):
        point(images=32, C=256, backward=True, zero_chunk_images=chunk)
        point(images=32, layout="nchw", backward=True, zero_chunk_images=chunk)
        point(images=8, backward=True, zero_chunk_images=chunk)
    point(images=1, backward=True)
    point(images=1, layout="nchw", backward=True)


def sweep_bwd_nchw():
    for images in (32, 8, 1):
        for mode in (4, 1, 3):
            point(images=images, layout="nchw", backward=True, zero_chunk_images=-1, bwd_mode=mode)
    point(images=32, layout="nchw", C=256, backward=True, zero_chunk_images=-1, bwd_mode=4)
    point(images=32, layout="nchw", C=256, backward=True, zero_chunk_images=-1, bwd_mode=1)


def sweep_bf16():
    for images, S in ((1, 8), (32, 1)):
        for C in (64, 256):
            for v in (5, 7, 0):
                point(C=C, images=images, streams=S, dtype="bf16", variant=v, rois_ready=True, concurrency=S)


def sweep_nchw2():
    """Row-segment kernel: ring depth / stage size / L2 policy hints (variants 3-12 of launch_fwd_nchw)."""
    vs = [int(v) for v in os.environ.get("SWEEP_VARIANTS", "2,3,4").split(",")]
    for images, S in ((32, 1), (1, 8)):
        for v in vs:
            point(layout="nchw", images=images, streams=S, variant=v, rois_ready=True)


def set_l2_fetch_granularity(nbytes):
    """cudaLimitMaxL2FetchGranularity (a device-wide hint, experiment only): 32 / 64 / 128 bytes fetched from DRAM per L2 miss."""
    import ctypes
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    rt = ctypes.CDLL("libcudart.so.12") if not os.path.exists("/usr/local/cuda/lib64/libcudart.so") else ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    got = ctypes.c_size_t(0)
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(nbytes))
    rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
    print("# cudaLimitMaxL2FetchGranularity <- %d: rc %d, now %d" % (nbytes, rc, got.value), flush=True)


def sweep_nchw_ab():
    """Default NCHW forward at the bench's points (for alternating runs of two library builds, SWEEP_LIB)."""
    if os.environ.get("SWEEP_L2_FETCH"):
        set_l2_fetch_granularity(int(os.environ["SWEEP_L2_FETCH"]))
    point(layout="nhwc", images=32, streams=1)
    point(layout="nchw", images=32, streams=1)
    point(layout="nchw", images=32, streams=1, rois_ready=True)
    point(layout="nchw", images=1, streams=8, concurrency=8, rois_ready=True)


def sweep_cfg4():
    """cfg4's per-GPU batch (2 048 RoIs, one launch), channels-last: every forward variant on one box."""
    for v in (5, 4, 3, 2, 1, 6, 12, 14, 15, 16, 5):
        point(layout="nhwc", images=32, streams=1, variant=v)
    for v in (0, 1, 2, 3, 4, 6, 7):
        point(layout="nhwc", images=32, streams=1, dtype="bf16", variant=v)


def sweep_nchw():
    for images, S in ((1, 1), (1, 8), (32, 1)):
        for cg in (2, 4, 8):
            point(layout="nchw", images=images, streams=S, nchw_cg=cg, rois_ready=True)
        point(layout="nchw", images=images, streams=S, nchw_tma=1)
        point(layout="nchw", images=images, streams=S, variant=2, rois_ready=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fwd"]
    for w in which:
        {"fwd": sweep_fwd, "bwd": sweep_bwd, "bwd_nchw": sweep_bwd_nchw, "bf16": sweep_bf16, "nchw": sweep_nchw, "nchw2": sweep_nchw2, "nchw_ab": sweep_nchw_ab, "cfg4": sweep_cfg4}[w]()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep_%s.json" % "_".join(which)), "w") as f:
        json.dump(RECS, f, indent=1)
