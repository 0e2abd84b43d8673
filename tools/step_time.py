#!/usr/bin/env python
"""CUDA-graph time of the 8-image end-to-end inference step under the current FOTS_B200_TC_* switches (A/B tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fots.pytorch_b200 import _cabi  # noqa: E402
if os.environ.get("SWEEP_LIB"):        # A/B against another build of the library on the same box (development only)
    _cabi.LIB_PATH = os.environ["SWEEP_LIB"]

from fots.pytorch_b200.pipeline import FOTSNet, FOTSPipeline  # noqa: E402
from fots.pytorch_b200.pipeline import conv as TC  # noqa: E402
from fots.pytorch_b200.pipeline.infer import planted_quads  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = FOTSNet(attention=True, nclass=89).to_b200(dev, inference=True)
    pipe = FOTSPipeline(net, 8, 64, 0.25, amp_dtype=torch.bfloat16)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    images = torch.randint(0, 256, (B, 720, 1280, 3), device=dev, dtype=torch.uint8).permute(0, 3, 1, 2)
    quads = torch.from_numpy(planted_quads(B, 64)).to(dev)
    lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    micro = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    g = pipe.capture(images, quads, micro=micro, lanes=lanes)
    for _ in range(3):
        g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g()
    e1.record()
    torch.cuda.synchronize()
    print("micro=%d lanes=%d LEVEL=%d STATS=%d ENABLED=%d: %.3f ms per %d-image step (CUDA graph), %.0f images/s" % (
        micro, lanes, TC.LEVEL, int(TC.FUSE_STATS), int(TC.ENABLED), e0.elapsed_time(e1) / 20, B, B / (e0.elapsed_time(e1) / 20e3)), flush=True)
