"""Decode the reference's only committed golden artefacts into a small numpy fixture.

Run in the BUILD container (needs /root/reference):  python tests/golden/make_jpeg_golden.py

Inputs (read-only, /root/reference/rroi_align/data/):
  timg.jpeg            276x500 input image of rroi_align/test2.py:24-31
  res0.jpg..res2.jpg   44x349 forward outputs written by rroi_align/test2.py:79-86
  grad.jpg             uint8-wrapped gradient image written by rroi_align/test2.py:92-97
Output: tests/golden/ref_test2_jpeg.npz  (decoded uint8 arrays, BGR channel order as cv2.imread
returns them -- the same order test2.py feeds to the op).
"""
import os
import cv2
import numpy as np

SRC = "/root/reference/rroi_align/data"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_test2_jpeg.npz")

if __name__ == "__main__":
    arrs = {"timg": cv2.imread(os.path.join(SRC, "timg.jpeg"))}
    for i in range(3):
        arrs["res%d" % i] = cv2.imread(os.path.join(SRC, "res%d.jpg" % i))
    arrs["grad"] = cv2.imread(os.path.join(SRC, "grad.jpg"))
    for k, v in arrs.items():
        assert v is not None and v.dtype == np.uint8, k
        print(k, v.shape)
    np.savez_compressed(DST, **arrs)
    print("wrote", DST, os.path.getsize(DST), "bytes")
