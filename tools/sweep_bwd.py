#!/usr/bin/env python
"""Backward sweep: packed vs generic channels-last kernel, NCHW with/without warp dedupe (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sweep  # noqa: E402

if __name__ == "__main__":
    for images in (1, 8, 32):
        steps = max(200, 4000 // images)
        for dd in (1, 2):
            sweep.point("nhwc", 64, images, steps, pdl=1, backward=True, dedupe=dd)
        sweep.point("nhwc", 256, images, steps, pdl=1, backward=True, dedupe=1)
        sweep.point("nchw", 64, images, steps, cg=4, pdl=1, backward=True, dedupe=1)
