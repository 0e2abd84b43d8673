"""Detection decode on the GPU: threshold + per-pixel quadrangle + raster-order
compaction in two launches, replacing the first half of nms/adaptor.cpp (:76-117) and the three full-map D2H
copies of test.py:86-96; then the reference's sequential merge (nms/nms.h) on the host over the compact candidate
rows, after one device-to-host copy (`merge_candidates`) -- together the replacement of nms.get_boxes
(nms/__init__.py:19-30, called at test.py:96)."""
import ctypes

import numpy as np
import torch

from .. import _cabi
from .rois import _lib

_bound = False


def _bind():
    global _bound
    L = _lib()
    if not _bound:
        import ctypes
        i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        L.fots_b200_decode_candidates.restype = i
        L.fots_b200_decode_candidates.argtypes = [vp, vp, vp, i, i, i, f, i, vp, vp, vp, vp]
        _bound = True
    return L


def decode_candidates(seg, rbox, angle, segm_threshold=0.5, max_per_image=4096):
    """seg [B,1,h,w], rbox [B,4,h,w], angle [B,2,h,w] (the network's first-scale outputs, fp32 CUDA) ->
    (counts int32 [B], cand int32 [B, max_per_image, 16]); row layout in include/fots_b200_pipeline.h."""
    if not (seg.is_cuda and rbox.is_cuda and angle.is_cuda):
        raise RuntimeError("decode_candidates: CUDA tensors required")
    seg, rbox, angle = seg.float().contiguous(), rbox.float().contiguous(), angle.float().contiguous()
    B, _, h, w = seg.shape
    if rbox.shape != (B, 4, h, w) or angle.shape != (B, 2, h, w):
        raise ValueError("decode_candidates: expected seg [B,1,h,w], rbox [B,4,h,w], angle [B,2,h,w]")
    counts = torch.empty((B,), dtype=torch.int32, device=seg.device)
    cand = torch.zeros((B, max_per_image, 16), dtype=torch.int32, device=seg.device)
    scratch = torch.empty((B * ((h * w + 255) // 256),), dtype=torch.int32, device=seg.device)
    with torch.cuda.device(seg.device):
        st = _bind().fots_b200_decode_candidates(seg.data_ptr(), rbox.data_ptr(), angle.data_ptr(), B, h, w,
                                                 float(segm_threshold), int(max_per_image), counts.data_ptr(),
                                                 cand.data_ptr(), scratch.data_ptr(),
                                                 torch.cuda.current_stream(seg.device).cuda_stream)
    _cabi.check(st, "fots_b200_decode_candidates")
    return counts, cand


def candidates_to_quads(cand_rows):
    """int32 rows [M,16] -> float32 [M,9] (x0..y3 in pixels, score), the row format nms/__init__.py:10-15 returns."""
    q = cand_rows[:, :8].to(torch.float32) / 10000.0
    return torch.cat((q, cand_rows[:, 8:9].contiguous().view(torch.float32)), 1)


def merge_candidates(counts, cand, w, h, iou_threshold1=0.4, iou_threshold2=0.2, max_boxes=1024):
    """Host stage of the detector post-processing (fots_b200_merge_candidates_host): counts int32 [B], cand int32
    [B, cap, 16] as returned by decode_candidates (CUDA or CPU tensors) -> list of B float32 arrays [k_b, 9]
    (x0,y0..x3,y3 in input-image pixels + accumulated score), what nms.get_boxes returns per image.  One D2H copy of
    the candidate rows that exist; thresholds default to the reference's (nms/__init__.py:29)."""
    L = _bind()
    if not getattr(L, "_merge_bound", False):
        i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        L.fots_b200_merge_candidates_host.restype = i
        L.fots_b200_merge_candidates_host.argtypes = [vp, i, i, i, f, f, vp, i, vp]
        L._merge_bound = True
    counts_h = counts.detach().cpu().numpy()
    cap = cand.size(1)
    top = int(min(int(counts_h.max()) if counts_h.size else 0, cap))
    rows = np.ascontiguousarray(cand[:, :top].detach().cpu().numpy()) if top > 0 else np.zeros((cand.size(0), 0, 16), np.int32)
    out = []
    buf = np.empty((max_boxes, 9), np.float32)
    nb = ctypes.c_int(0)
    for b in range(rows.shape[0]):
        n = int(min(counts_h[b], cap))
        img = np.ascontiguousarray(rows[b, :n])
        st = L.fots_b200_merge_candidates_host(img.ctypes.data if n else None, n, int(w), int(h), float(iou_threshold1),
                                               float(iou_threshold2), buf.ctypes.data, int(max_boxes), ctypes.byref(nb))
        _cabi.check(st, "fots_b200_merge_candidates_host")
        k = min(nb.value, max_boxes)
        boxes = buf[:k].copy()
        boxes[:, :8] /= 10000.0                     # nms/__init__.py:13-15
        out.append(boxes)
    return out
