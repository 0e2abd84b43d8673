#!/usr/bin/env python
"""bf16 RoIRotate forward: kernel variants x (cfg1 on 8 streams, cfg4 per-GPU batch on 1 stream)."""
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

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    peak, _ = bench.measured_peak_gbs()
    for images, streams in ((1, 8), (32, 1)):
        for C in (64, 256):
            a = types.SimpleNamespace(channels=C, layout="nhwc", images=images, rois_per_image=64, sets=0, dtype="bf16")
            w = bench.Workload(a, dev, torch)
            for variant in (5, 7, 0):
                _cabi.set_tuning(_cabi.TUNE_NHWC_UNROLL, variant)
                steps = max(200, 20000 // images)
                ms = bench.timed_steps(w, steps, 20, 500, torch, _cabi.lib(), _cabi, lambda: None, streams)
                us = ms / steps * 1e3
                alg = float(np.mean(w.alg_bytes))
                print("bf16 C=%d images=%d streams=%d variant=%d: %.2f us  %.0f GB/s  %.1f%% of %.0f" % (
                    C, images, streams, variant, us, alg / us / 1e3, 100 * alg / us / 1e3 / peak, peak), flush=True)
            del w
            torch.cuda.empty_cache()
