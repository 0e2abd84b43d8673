#!/usr/bin/env python
"""GPU tuning sweep for the RoIRotate forward/backward kernels (development tool, run under gpurun).

    python tools/sweep.py [--quick] > gpurun_out/sweep.txt

Same timing harness as bench.py (rotating buffer sets > L2, CUDA graphs, CUDA events) over the grid
layout x channels x RoIs-per-step x tuning knobs.  Prints one line per point and writes
gpurun_out/sweep.json.
"""
import argparse
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


def point(layout, C, images, steps, cg=0, unroll=0, pdl=0, chunk=500, rois_per_image=64, streams=1, backward=False, dedupe=1):
    args = types.SimpleNamespace(channels=C, layout=layout, images=images, rois_per_image=rois_per_image,
                                 sets=0, graph_chunk=chunk, pdl=pdl)
    dev = torch.device("cuda:0")
    wl = bench.Workload(args, dev, torch)
    if backward:
        wl.enable_backward(torch)
    _cabi.set_tuning(_cabi.TUNE_BWD_DEDUPE, dedupe)
    _cabi.set_tuning(_cabi.TUNE_NCHW_CG, cg)
    _cabi.set_tuning(_cabi.TUNE_NHWC_UNROLL, unroll)
    _cabi.set_tuning(_cabi.TUNE_USE_PDL, pdl)
    ms = bench.timed_steps(wl, steps, 50, chunk, torch, _cabi.lib(), _cabi, lambda: None, streams)
    us = ms / steps * 1e3
    alg = float(np.mean(wl.alg_bytes))
    peak, _ = bench.measured_peak_gbs()
    gbs = alg / us / 1e3
    rec = dict(layout=layout, C=C, images=images, rois=wl.N, cg=cg, unroll=unroll, pdl=pdl, chunk=chunk, streams=streams,
               us_per_launch=us, alg_mb=alg / 1e6, gbs=gbs, frac=gbs / peak,
               mfeat_px_s=wl.feat_px_per_step / us, sets=wl.sets)
    if backward:
        import workloads as WL
        # SURVEY 8d backward bytes: read top_diff on valid elements + define the whole grad map + 24 N
        # (the RMW of the touched pixels is L2 traffic, not counted)
        alg = float(np.mean([4 * C * int(WL.valid_counts(r, 8, 64).sum()) + 24 * len(r) for r in wl.rois_np])) + 4.0 * images * C * 180 * 320
        gbs = alg / us / 1e3
        rec.update(alg_mb=alg / 1e6, gbs=gbs, frac=gbs / peak, backward=True, dedupe=dedupe)
        print("BWD  ", end="")
    print("%-5s C=%-3d img=%-3d N=%-5d cg=%-2d U=%d pdl=%d S=%d | %8.2f us  %7.1f GB/s  frac %.3f" % (
        layout, C, images, wl.N, cg, unroll, pdl, streams, us, gbs, gbs / peak), flush=True)
    del wl
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    recs = []
    steps = 5000 if a.quick else 20000
    for C in (64, 256):
        for pdl in (0, 1):
            for U in (1, 2):
                recs.append(point("nhwc", C, 1, steps, unroll=U, pdl=pdl))
            for cg in (2, 4, 8):
                recs.append(point("nchw", C, 1, steps, cg=cg, pdl=pdl))
    for C in (64, 256):
        for S in (2, 3, 4):
            recs.append(point("nhwc", C, 1, steps, unroll=1, pdl=1, streams=S))
            recs.append(point("nchw", C, 1, steps, cg=4, pdl=1, streams=S))
    # batched steps (cfg2: 8 images, cfg4 per GPU: 32 images)
    for images in (8, 32):
        for layout, kw in (("nhwc", dict(unroll=3)), ("nhwc", dict(unroll=5)), ("nchw", dict(cg=2)), ("nchw", dict(cg=4)), ("nchw", dict(cg=8))):
            recs.append(point(layout, 64, images, max(steps // images, 500), pdl=1, **kw))
    for images in (1, 8, 32):
        for layout, kws in (("nhwc", [dict()]), ("nchw", [dict(cg=8, dedupe=1), dict(cg=8, dedupe=0), dict(cg=4, dedupe=1), dict(cg=16, dedupe=1)])):
            for kw in kws:
                recs.append(point(layout, 64, images, max(steps // (4 * images), 200), pdl=1, backward=True, **kw))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
        json.dump(recs, f, indent=1)
