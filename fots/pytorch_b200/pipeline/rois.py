"""Batched RoI construction on the GPU: replaces the per-box Python of
tools/ocr_utils.py:133-152 (inference) and the width rule of src/ocr_process.py:260-263 (training)."""
import math

import torch

from .. import _cabi


def _lib():
    L = _cabi.lib()
    if not getattr(L, "_pipeline_bound", False):
        import ctypes
        i, vp = ctypes.c_int, ctypes.c_void_p
        L.fots_b200_boxes_to_rois.restype = i
        L.fots_b200_boxes_to_rois.argtypes = [vp, i, vp, i, vp, vp]
        L.fots_b200_ctc_greedy.restype = i
        L.fots_b200_ctc_greedy.argtypes = [vp, i, i, i, vp, vp, vp]
        L._pipeline_bound = True
    return L


def boxes_to_rois(quads, batch_idx=None):
    """quads [M, >=8] fp32 CUDA (x0,y0..x3,y3[, score]) -> RoI rows [M, 6] = [b, int(cx), int(cy), h, w, -angle_deg],
    the row tools/ocr_utils.py:143-145 builds per box.  batch_idx: int32 [M] or None (all image 0)."""
    if not quads.is_cuda or quads.dtype != torch.float32 or quads.dim() != 2 or quads.size(1) < 8:
        raise ValueError("boxes_to_rois: quads must be an fp32 CUDA tensor [M, >=8]")
    quads = quads.contiguous()
    M = quads.size(0)
    if batch_idx is not None:
        batch_idx = batch_idx.to(device=quads.device, dtype=torch.int32).contiguous()
        if batch_idx.numel() != M:
            raise ValueError("boxes_to_rois: batch_idx must have one entry per box")
    rois = torch.empty((M, 6), dtype=torch.float32, device=quads.device)
    with torch.cuda.device(quads.device):
        st = _lib().fots_b200_boxes_to_rois(quads.data_ptr(), quads.size(1),
                                            batch_idx.data_ptr() if batch_idx is not None else None, M,
                                            rois.data_ptr(), torch.cuda.current_stream(quads.device).cuda_stream)
    _cabi.check(st, "fots_b200_boxes_to_rois")
    return rois


def pooled_width_for(h, w, pooled_height=11, mode="infer"):
    """Pooled width rule of the reference's callers, from host scalars (no device sync).
    infer: tools/ocr_utils.py:147-150  max(2, (int(w * PH / max(1, h)) + PH) // 32) * 32   (one box)
    train: src/ocr_process.py:260-263  ceil(PH * max(w / h))                                (h, w iterables)"""
    if mode == "infer":
        scale = pooled_height / max(1.0, float(h))
        return max(2, (int(float(w) * scale) + pooled_height) // 32) * 32
    ratios = [float(wi) / float(hi) for hi, wi in zip(h, w)]
    return int(math.ceil(pooled_height * max(ratios)))
