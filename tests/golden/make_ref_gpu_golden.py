"""Generate golden vectors from the reference CUDA kernel ITSELF, run on a B200.

    gpurun -- 'python tests/golden/make_ref_gpu_golden.py'      (writes gpurun_out/golden/)
    cp gpurun_out/golden/ref_sm100a_golden.npz tests/golden/

The kernel is /root/reference/rroi_align/src/rroi_align_kernel.cu compiled UNMODIFIED for sm_100a by
oracle/Makefile into oracle/_ref/libref_rroi_sm100a.so (that .so travels to the GPU box; the
reference sources do not).  Launches follow rroi_align/functions/rroi_align.py:17-28,35-38: zero-filled
outputs, RROIAlignForwardLaucher / RROIAlignBackwardLaucher on the current stream.

Inputs are the seeded generators of tests/workloads.py (a checksum of each input is stored so a
change of the generator is detected); outputs are stored as fp32 bit patterns.  To keep the fixture
small only channel 0 of the big cases' values is kept -- the sample centres do not depend on the
channel (checked here before dropping them).
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as WL  # noqa: E402
from oracle import rroi_oracle as O  # noqa: E402


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


def cases():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_test2_jpeg.npz"))
    img = g["timg"].astype(np.float32).transpose(2, 0, 1)[None]
    rois2, ph2, pw2 = WL.test2_rois()
    yield "cfg0", WL.cfg0(), None
    yield "test2", (img, rois2, ph2, pw2, 1.0), None
    yield "cfg1", WL.cfg1(64), 1
    yield "stress1", (WL.features(1, 2, 5, 45, 80), WL.stress_rois(1, 96, 2, 320, 180), 8, 64, 0.25), 1
    yield "stress2", (WL.features(2, 3, 8, 45, 80), WL.stress_rois(2, 64, 3, 320, 180), 11, 37, 0.25), 1


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    out = {}
    for name, (feats, rois, ph, pw, scale), keep_c in cases():
        f = torch.from_numpy(feats).to(dev)
        r = torch.from_numpy(rois).to(dev)
        y, ix, iy = O.ref_gpu_forward(f, r, ph, pw, scale)
        torch.cuda.synchronize()
        assert (ix == ix[:, :1]).all() and (iy == iy[:, :1]).all(), "centres depend on the channel?!"
        gen = torch.Generator(device="cpu").manual_seed(7)
        gtop = torch.randn(y.shape, generator=gen)
        gb = O.ref_gpu_backward(gtop.to(dev), r, ix, iy, tuple(f.shape), scale)
        torch.cuda.synchronize()
        y, ix, iy, gb = y.cpu().numpy(), ix.cpu().numpy(), iy.cpu().numpy(), gb.cpu().numpy()
        out[name + "_feat_sha"] = digest(feats)
        out[name + "_rois"] = rois
        out[name + "_meta"] = np.array([ph, pw, scale, feats.shape[0], feats.shape[1], feats.shape[2], feats.shape[3]], np.float64)
        out[name + "_idx_x"] = ix[:, 0]
        out[name + "_idx_y"] = iy[:, 0]
        out[name + "_out"] = y if keep_c is None else y[:, :keep_c]
        out[name + "_grad_sha_seed"] = np.array([7])
        out[name + "_bgrad"] = gb if keep_c is None else gb[:, :keep_c]
        print(name, y.shape, "valid frac %.3f" % float((ix != 0).mean()))
    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    path = os.path.join(dst, "ref_sm100a_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))
