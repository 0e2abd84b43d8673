#!/usr/bin/env python
"""cfg1 forward: streams x PDL x NHWC variant sweep (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sweep  # noqa: E402

if __name__ == "__main__":
    for S in (8, 12, 16):
        for U in (1, 3, 5):
            sweep.point("nhwc", 64, 1, 20000, unroll=U, pdl=1, streams=S)
    for S in (8,):
        sweep.point("nhwc", 256, 1, 20000, unroll=1, pdl=1, streams=S)
        sweep.point("nchw", 64, 1, 20000, cg=4, pdl=1, streams=S)
        sweep.point("nchw", 64, 1, 20000, cg=2, pdl=1, streams=S)
