#!/usr/bin/env python
"""Backward sweep: packed vs generic channels-last kernel, NCHW with/without warp dedupe (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sweep  # noqa: E402

if __name__ == "__main__":
    from fots.pytorch_b200 import _cabi
    for fused in (0, 1):           # memset + scatter, then the one-pass zero + scatter (maps >= 96 MB)
        _cabi.set_tuning(_cabi.TUNE_BWD_ZERO_FUSED, fused)
        print("zero-fill:", "fused (maps >= 96 MB)" if fused else "memset + scatter", flush=True)
        for images in (8, 32):
            sweep.point("nhwc", 64, images, max(200, 4000 // images), pdl=1, backward=True, dedupe=1)
            sweep.point("nhwc", 256, images, max(200, 4000 // images), pdl=1, backward=True, dedupe=1)
    _cabi.set_tuning(_cabi.TUNE_BWD_ZERO_FUSED, 0)
    for images in (1, 8, 32):
        steps = max(200, 4000 // images)
        for dd in (1, 2):
            sweep.point("nhwc", 64, images, steps, pdl=1, backward=True, dedupe=dd)
        sweep.point("nhwc", 256, images, steps, pdl=1, backward=True, dedupe=1)
        sweep.point("nchw", 64, images, steps, cg=4, pdl=1, backward=True, dedupe=1)
