"""ctypes front-end of oracle/librroi_oracle.so (the C restatement of the reference kernels,
/root/reference/rroi_align/src/rroi_align_kernel.cu:28-162 and :193-278) and of
oracle/_ref/libref_rroi_sm100a.so (the reference .cu itself, compiled unmodified for sm_100a).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CPU_SO = os.path.join(_HERE, "librroi_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_rroi_sm100a.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_cpu = None
_ref = None


def build(quiet=True):
    """(Re)build the oracle with oracle/Makefile (also builds oracle/_ref when the reference is present)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def cpu_lib():
    global _cpu
    if _cpu is None:
        if not os.path.exists(_CPU_SO):
            build()
        lib = ctypes.CDLL(_CPU_SO)
        lib.rroi_oracle_forward.restype = ctypes.c_int
        lib.rroi_oracle_forward.argtypes = [
            _f32p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, _f32p, ctypes.c_int]
        lib.rroi_oracle_backward.restype = ctypes.c_int
        lib.rroi_oracle_backward.argtypes = [
            _f32p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, _f32p, ctypes.c_int]
        lib.rroi_oracle_cosf.restype = ctypes.c_float
        lib.rroi_oracle_cosf.argtypes = [ctypes.c_float]
        lib.rroi_oracle_sinf.restype = ctypes.c_float
        lib.rroi_oracle_sinf.argtypes = [ctypes.c_float]
        lib.rroi_oracle_max_threads.restype = ctypes.c_int
        lib.detect_oracle_decode.restype = ctypes.c_int
        lib.detect_oracle_decode.argtypes = [_f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
        _cpu = lib
    return _cpu


def _p(a):
    return a.ctypes.data_as(_f32p)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def forward(features, rois, pooled_height, pooled_width, spatial_scale, threads=1):
    """features [B,C,H,W] fp32, rois [N,6] -> (out, idx_x, idx_y), each [N,C,PH,PW] fp32.

    Buffers are zero-filled here like rroi_align/functions/rroi_align.py:17-20 does.
    """
    features = _c32(features)
    rois = _c32(rois).reshape(-1, 6)
    B, C, H, W = features.shape
    N = rois.shape[0]
    shape = (N, C, int(pooled_height), int(pooled_width))
    out = np.zeros(shape, np.float32)
    ix = np.zeros(shape, np.float32)
    iy = np.zeros(shape, np.float32)
    rc = cpu_lib().rroi_oracle_forward(_p(features), float(spatial_scale), N, H, W, C,
                                       int(pooled_height), int(pooled_width), _p(rois),
                                       _p(out), _p(ix), _p(iy), int(threads))
    assert rc == 1
    return out, ix, iy


def backward(top_diff, rois, idx_x, idx_y, feature_size, spatial_scale, threads=1):
    """top_diff/idx_x/idx_y [N,C,PH,PW], feature_size (B,C,H,W) -> bottom_diff [B,C,H,W]."""
    top_diff = _c32(top_diff)
    idx_x = _c32(idx_x)
    idx_y = _c32(idx_y)
    rois = _c32(rois).reshape(-1, 6)
    B, C, H, W = feature_size
    N, C2, PH, PW = top_diff.shape
    assert C2 == C and idx_x.shape == top_diff.shape and idx_y.shape == top_diff.shape
    grad = np.zeros((B, C, H, W), np.float32)
    rc = cpu_lib().rroi_oracle_backward(_p(top_diff), float(spatial_scale), B, N, H, W, C, PH, PW,
                                        _p(rois), _p(grad), _p(idx_x), _p(idx_y), int(threads))
    assert rc == 1
    return grad


def decode_candidates(segm, rbox, angle, segm_threshold, max_rows):
    """nms/adaptor.cpp:76-117 restated: segm [h,w], rbox [4,h,w], angle [2,h,w] -> (count, cand int32 [max_rows,16])."""
    segm, rbox, angle = _c32(segm), _c32(rbox), _c32(angle)
    h, w = segm.shape
    cand = np.zeros((max_rows, 16), np.int32)
    n = cpu_lib().detect_oracle_decode(_p(segm), _p(rbox), _p(angle), h, w, float(segm_threshold), int(max_rows),
                                       cand.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return n, cand


def max_threads():
    return int(cpu_lib().rroi_oracle_max_threads())


def cosf(x):
    return float(cpu_lib().rroi_oracle_cosf(float(np.float32(x))))


def sinf(x):
    return float(cpu_lib().rroi_oracle_sinf(float(np.float32(x))))


# --------------------------------------------------------------------------------------
# The reference CUDA kernel itself (GPU box only): oracle/_ref/libref_rroi_sm100a.so exports
# RROIAlignForwardLaucher / RROIAlignBackwardLaucher (rroi_align/src/rroi_align_kernel.h:8-18).

def ref_gpu_available():
    return os.path.exists(_REF_SO)


def ref_gpu_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            raise RuntimeError("oracle/_ref/libref_rroi_sm100a.so missing: run `make -C oracle ref` "
                               "in the build container (needs /root/reference)")
        lib = ctypes.CDLL(_REF_SO)
        vp = ctypes.c_void_p
        lib.RROIAlignForwardLaucher.restype = ctypes.c_int
        lib.RROIAlignForwardLaucher.argtypes = [vp, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, vp, vp, vp, vp, vp]
        lib.RROIAlignBackwardLaucher.restype = ctypes.c_int
        lib.RROIAlignBackwardLaucher.argtypes = [vp, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
        _ref = lib
    return _ref


def ref_gpu_forward(features, rois, pooled_height, pooled_width, spatial_scale):
    """Run the unmodified reference kernel on cuda tensors, with the zero-fills of
    rroi_align/functions/rroi_align.py:17-20.  Returns (out, idx_x, idx_y) cuda tensors."""
    import torch
    assert features.is_cuda and rois.is_cuda
    features = features.contiguous().float()
    rois = rois.contiguous().float().view(-1, 6)
    B, C, H, W = features.shape
    N = rois.shape[0]
    out = torch.zeros(N, C, pooled_height, pooled_width, device=features.device)
    ix = torch.zeros_like(out)
    iy = torch.zeros_like(out)
    st = torch.cuda.current_stream(features.device).cuda_stream
    rc = ref_gpu_lib().RROIAlignForwardLaucher(features.data_ptr(), float(spatial_scale), N, H, W, C,
                                               int(pooled_height), int(pooled_width), rois.data_ptr(),
                                               out.data_ptr(), ix.data_ptr(), iy.data_ptr(), st)
    assert rc == 1
    return out, ix, iy


def ref_gpu_backward(top_diff, rois, idx_x, idx_y, feature_size, spatial_scale):
    import torch
    top_diff = top_diff.contiguous().float()
    rois = rois.contiguous().float().view(-1, 6)
    B, C, H, W = feature_size
    N, _, PH, PW = top_diff.shape
    grad = torch.zeros(B, C, H, W, device=top_diff.device)
    st = torch.cuda.current_stream(top_diff.device).cuda_stream
    rc = ref_gpu_lib().RROIAlignBackwardLaucher(top_diff.data_ptr(), float(spatial_scale), B, N, H, W, C,
                                                PH, PW, rois.data_ptr(), grad.data_ptr(),
                                                idx_x.contiguous().data_ptr(),
                                                idx_y.contiguous().data_ptr(), st)
    assert rc == 1
    return grad


# --------------------------------------------------------------------------------------
# The reference's own merge (nms/nms.h + vendored Clipper) behind oracle/nms_ref_shim.cpp:
# oracle/_ref/libref_nms.so, built by oracle/Makefile where /root/reference is present; CPU code, so it
# runs in the build container and on the GPU box alike.
_REF_NMS_SO = os.path.join(_HERE, "_ref", "libref_nms.so")
_ref_nms = None


def ref_nms_available():
    return os.path.exists(_REF_NMS_SO)


def ref_merge_candidates(cand, w, h, thr1=0.4, thr2=0.2, max_boxes=4096):
    """cand int32 [n,16] (one image, raster order) -> float32 [k,9] exactly as nms/adaptor.cpp:13-29 flattens the
    result of nms::merge_iou (coordinates in x10000 units)."""
    global _ref_nms
    if _ref_nms is None:
        if not os.path.exists(_REF_NMS_SO):
            raise RuntimeError("oracle/_ref/libref_nms.so missing: run `make -C oracle ref` in the build container")
        lib = ctypes.CDLL(_REF_NMS_SO)
        lib.ref_merge_candidates.restype = ctypes.c_int
        lib.ref_merge_candidates.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
        _ref_nms = lib
    cand = np.ascontiguousarray(cand, dtype=np.int32).reshape(-1, 16)
    out = np.zeros((max_boxes, 9), np.float32)
    n = _ref_nms.ref_merge_candidates(cand.ctypes.data, cand.shape[0], int(w), int(h), float(thr1), float(thr2),
                                      out.ctypes.data, int(max_boxes))
    return out[:min(n, max_boxes)].copy()
