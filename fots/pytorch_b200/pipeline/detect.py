"""Detection decode on the GPU: threshold + per-pixel quadrangle + raster-order
compaction in two launches, replacing the first half of nms/adaptor.cpp (:76-117) and the three full-map D2H
copies of test.py:86-96; then the reference's sequential merge (nms/nms.h) on the host over the compact candidate
rows, after one device-to-host copy (`merge_candidates`) -- together the replacement of nms.get_boxes
(nms/__init__.py:19-30, called at test.py:96)."""
import ctypes
import os
import warnings

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


def decode_candidates(seg, rbox, angle, segm_threshold=0.5, max_per_image=None, out=None):
    """seg [B,1,h,w], rbox [B,4,h,w], angle [B,2,h,w] (the network's first-scale outputs, fp32 CUDA) ->
    (counts int32 [B], cand int32 [B, max_per_image, 16]); row layout in include/fots_b200_pipeline.h."""
    if not (seg.is_cuda and rbox.is_cuda and angle.is_cuda):
        raise RuntimeError("decode_candidates: CUDA tensors required")
    seg, rbox, angle = seg.float().contiguous(), rbox.float().contiguous(), angle.float().contiguous()
    B, _, h, w = seg.shape
    if rbox.shape != (B, 4, h, w) or angle.shape != (B, 2, h, w):
        raise ValueError("decode_candidates: expected seg [B,1,h,w], rbox [B,4,h,w], angle [B,2,h,w]")
    if max_per_image is None:
        max_per_image = h * w                      # every positive pixel, like the reference
    if out is not None:                            # caller-owned (counts, cand, scratch): static buffers of a CUDA graph
        counts, cand, scratch = out
    else:
        counts = torch.empty((B,), dtype=torch.int32, device=seg.device)
        cand = torch.empty((B, max_per_image, 16), dtype=torch.int32, device=seg.device)
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


def host_threads():
    """Host cores this process may use for the merge worker pool."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def merge_rows_host(counts_h, rows_h, w, h, iou_threshold1=0.4, iou_threshold2=0.2, max_boxes=1024, threads=0):
    """The merge proper on HOST arrays: counts_h int32 [B], rows_h int32 [B, cap, 16] (C-contiguous) -> (boxes float32
    [B, max_boxes, 9] with coordinates already divided by 10000 (nms/__init__.py:13-15), num int32 [B]).  One native
    call for the whole micro-batch; images are dealt to `threads` host threads (0 = all this process may use)."""
    L = _bind()
    if not getattr(L, "_merge_bound", False):
        i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        L.fots_b200_merge_candidates_host.restype = i
        L.fots_b200_merge_candidates_host.argtypes = [vp, i, i, i, f, f, vp, i, vp]
        L.fots_b200_merge_candidates_host_batch.restype = i
        L.fots_b200_merge_candidates_host_batch.argtypes = [vp, vp, i, i, i, i, f, f, vp, i, vp, i]
        L._merge_bound = True
    B, cap = rows_h.shape[0], rows_h.shape[1]
    counts_h = np.ascontiguousarray(counts_h, np.int32)
    rows_h = np.ascontiguousarray(rows_h, np.int32)
    boxes = np.empty((B, max_boxes, 9), np.float32)
    num = np.zeros((B,), np.int32)
    st = L.fots_b200_merge_candidates_host_batch(rows_h.ctypes.data if rows_h.size else None, counts_h.ctypes.data, B, cap,
                                                 int(w), int(h), float(iou_threshold1), float(iou_threshold2),
                                                 boxes.ctypes.data, int(max_boxes), num.ctypes.data,
                                                 int(threads) if threads > 0 else host_threads())
    _cabi.check(st, "fots_b200_merge_candidates_host_batch")
    return boxes, num


def merge_candidates(counts, cand, w, h, iou_threshold1=0.4, iou_threshold2=0.2, max_boxes=1024, threads=0):
    """Host stage of the detector post-processing (fots_b200_merge_candidates_host_batch): counts int32 [B], cand int32
    [B, cap, 16] as returned by decode_candidates (CUDA or CPU tensors) -> list of B float32 arrays [k_b, 9]
    (x0,y0..x3,y3 in input-image pixels + accumulated score), what nms.get_boxes returns per image.  One D2H copy of
    the candidate rows that exist; thresholds default to the reference's (nms/__init__.py:29).  Nothing is dropped
    silently: more positive pixels than `cap` rows, or more merged boxes than `max_boxes`, raise a RuntimeWarning (the
    reference processes every pixel above the threshold; size `max_per_image` = h * w to guarantee the same)."""
    counts_h = counts.detach().cpu().numpy()
    cap = cand.size(1)
    if counts_h.size and int(counts_h.max()) > cap:
        warnings.warn("merge_candidates: %d positive pixels in one image but only %d candidate rows were kept (raster "
                      "order); pass max_per_image = h * w to decode_candidates" % (int(counts_h.max()), cap), RuntimeWarning)
    top = int(min(int(counts_h.max()) if counts_h.size else 0, cap))
    rows = np.ascontiguousarray(cand[:, :top].detach().cpu().numpy()) if top > 0 else np.zeros((cand.size(0), 0, 16), np.int32)
    boxes, num = merge_rows_host(np.minimum(counts_h, cap), rows, w, h, iou_threshold1, iou_threshold2, max_boxes, threads)
    if num.size and int(num.max()) > max_boxes:
        warnings.warn("merge_candidates: %d boxes in one image, only max_boxes = %d returned" % (int(num.max()), max_boxes),
                      RuntimeWarning)
    out = []
    for b in range(rows.shape[0]):
        k = min(int(num[b]), max_boxes)
        bx = boxes[b, :k].copy()
        bx[:, :8] /= 10000.0                     # nms/__init__.py:13-15
        out.append(bx)
    return out
