"""Detection decode on the GPU: threshold + per-pixel quadrangle + raster-order
compaction in two launches, replacing the first half of nms/adaptor.cpp (:76-117) and the three full-map D2H
copies of test.py:86-96.  The merge itself (nms/nms.h + Clipper) is CPU code and out of scope; it can be fed the
compact candidate rows this returns."""
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
