"""ctypes binding of include/rroi_align_b200.h (fots/pytorch_b200/lib/librroi_b200.so).

The library is plain C ABI (device pointers + sizes + cudaStream_t); torch only supplies the device
memory and the current stream.  Loading fails loudly when the .so has not been built -- there is no
fallback implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librroi_b200.so")

OK = 0
ERR_INVALID_ARG = -1
ERR_TOO_LARGE = -2
ERR_CUDA = -3

LAYOUT_NCHW = 0
LAYOUT_NHWC = 1

FLAG_NO_PDL = 1
FLAG_ROIS_READY = 2

ABI_VERSION = 2


class Opts(ctypes.Structure):
    """rroi_b200_opts (include/rroi_align_b200.h): per-call launch options.  The library keeps no tuning state."""
    _fields_ = [("size", ctypes.c_uint), ("flags", ctypes.c_uint), ("concurrency", ctypes.c_int),
                ("variant", ctypes.c_int), ("nchw_cg", ctypes.c_int), ("bwd_mode", ctypes.c_int),
                ("nchw_tma", ctypes.c_int), ("zero_chunk_images", ctypes.c_int)]


def opts(pdl=True, rois_ready=False, concurrency=0, variant=0, nchw_cg=0, bwd_mode=0, nchw_tma=0, zero_chunk_images=0):
    """Build an rroi_b200_opts; pass the result (or None = defaults) as `opts` to the *_opt entry points."""
    o = Opts()
    o.size = ctypes.sizeof(Opts)
    o.flags = (0 if pdl else FLAG_NO_PDL) | (FLAG_ROIS_READY if rois_ready else 0)
    o.concurrency, o.variant, o.nchw_cg = int(concurrency), int(variant), int(nchw_cg)
    o.bwd_mode, o.nchw_tma, o.zero_chunk_images = int(bwd_mode), int(nchw_tma), int(zero_chunk_images)
    return o


def opts_ref(o):
    return ctypes.byref(o) if o is not None else None


# every symbol include/*.h declares
EXPORTS = (
    "RROIAlignForwardLaucher", "RROIAlignBackwardLaucher",
    "rroi_b200_forward", "rroi_b200_backward", "rroi_b200_expand_idx", "rroi_b200_forward_bf16",
    "rroi_b200_forward_opt", "rroi_b200_backward_opt", "rroi_b200_forward_bf16_opt", "rroi_b200_roi_xform",
    "rroi_b200_last_cuda_error",
    "rroi_b200_strerror", "rroi_b200_abi_version", "rroi_b200_build_info",
    # include/fots_b200_pipeline.h
    "fots_b200_boxes_to_rois", "fots_b200_ctc_greedy", "fots_b200_instnorm_nhwc_bf16",
    "fots_b200_fpn_merge_nhwc_bf16", "fots_b200_fpn_merge_prob_nhwc_bf16", "fots_b200_decode_candidates",
    "fots_b200_conv2d_nhwc_bf16", "fots_b200_conv2d_strided_nhwc_bf16", "fots_b200_conv_set_tile", "fots_b200_conv_set_halo", "fots_b200_conv2d_stats_nhwc_bf16",
    "fots_b200_instnorm_apply_nhwc_bf16", "fots_b200_instnorm_bwd_nhwc_bf16", "fots_b200_instnorm_crelu_bwd_nhwc_bf16", "fots_b200_upsample_bilinear_bwd_nhwc_bf16", "fots_b200_instnorm_set_single_pass", "fots_b200_merge_candidates_host", "fots_b200_merge_candidates_host_batch",
    "fots_b200_stem_conv3x3_c3_c16", "fots_b200_stem_conv3x3_c3_c16_u8", "fots_b200_maxpool_h2_nhwc_bf16",
    "fots_b200_gemm_bf16w", "fots_b200_bilstm_recurrent", "fots_b200_dwconv3x3_nhwc_bf16", "fots_b200_dwconv3x3_norm_nhwc_bf16", "fots_b200_dwconv3x3_up_nhwc_bf16", "fots_b200_instnorm_stats_nhwc_bf16",
    "fots_b200_heads_nhwc_bf16", "fots_b200_heads_merged_nhwc_bf16", "fots_b200_heads_gather_nhwc_bf16", "fots_b200_conv1x1_to1_nhwc_bf16", "fots_b200_conv3x3_c3_pool_nhwc_bf16", "fots_b200_maxpool_nhwc_bf16",
)

_lib = None


class RRoiAlignError(RuntimeError):
    """A non-success status from the C ABI."""


def lib():
    """Load (once) and return the ctypes handle with argtypes set."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "fots.pytorch_b200: %s is missing -- build it with `make -C fots/pytorch_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU/PyTorch "
            "fallback for RoIRotate." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
    L.RROIAlignForwardLaucher.restype = i
    L.RROIAlignForwardLaucher.argtypes = [vp, f, i, i, i, i, i, i, vp, vp, vp, vp, vp]
    L.RROIAlignBackwardLaucher.restype = i
    L.RROIAlignBackwardLaucher.argtypes = [vp, f, i, i, i, i, i, i, i, vp, vp, vp, vp, vp]
    L.rroi_b200_forward.restype = i
    L.rroi_b200_forward.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, f, i, vp]
    L.rroi_b200_forward_bf16.restype = i
    L.rroi_b200_forward_bf16.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, f, vp]
    L.rroi_b200_backward.restype = i
    L.rroi_b200_backward.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, f, i, i, vp]
    L.rroi_b200_expand_idx.restype = i
    L.rroi_b200_expand_idx.argtypes = [vp, vp, i, i, i, i, vp]
    op = ctypes.POINTER(Opts)
    L.rroi_b200_forward_opt.restype = i
    L.rroi_b200_forward_opt.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, f, i, op, vp]
    L.rroi_b200_forward_bf16_opt.restype = i
    L.rroi_b200_forward_bf16_opt.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, f, op, vp]
    L.rroi_b200_backward_opt.restype = i
    L.rroi_b200_backward_opt.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, f, i, i, op, vp]
    L.rroi_b200_roi_xform.restype = i
    L.rroi_b200_roi_xform.argtypes = [vp, vp, i, i, f, vp]
    L.rroi_b200_last_cuda_error.restype = i
    L.rroi_b200_strerror.restype = ctypes.c_char_p
    L.rroi_b200_strerror.argtypes = [i]
    L.rroi_b200_abi_version.restype = i
    L.rroi_b200_build_info.restype = ctypes.c_char_p
    if L.rroi_b200_abi_version() != ABI_VERSION:
        raise ImportError("librroi_b200.so ABI version %d != binding %d: rebuild the library"
                          % (L.rroi_b200_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(status, what):
    if status != OK:
        L = lib()
        msg = L.rroi_b200_strerror(status).decode()
        if status == ERR_CUDA:
            msg += " [cudaError_t=%d]" % L.rroi_b200_last_cuda_error()
        raise RRoiAlignError("%s failed: %s" % (what, msg))


def build_info():
    return lib().rroi_b200_build_info().decode()
