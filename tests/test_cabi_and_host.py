"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/rroi_align_b200.h declares; argument errors are reported before any CUDA call; the
Python host mirrors the reference's module/function surface and refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(RROIAlign\w+|rroi_b200_\w+|fots_b200_\w+)\s*\(", src))
    return sorted(names)


def test_header_symbols_are_exported():
    from fots.pytorch_b200 import _cabi
    names = _declared_functions()
    assert set(names) == set(_cabi.EXPORTS), (names, _cabi.EXPORTS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_reference_launcher_names_present():
    """The two symbols of rroi_align/src/rroi_align_kernel.h:8-18, spelling included."""
    from fots.pytorch_b200 import _cabi
    lib = _cabi.lib()
    assert lib.RROIAlignForwardLaucher and lib.RROIAlignBackwardLaucher


def test_argument_errors_without_gpu():
    from fots.pytorch_b200 import _cabi
    L = _cabi.lib()
    assert L.rroi_b200_abi_version() == _cabi.ABI_VERSION
    assert "sm_100a" in _cabi.build_info()
    assert L.rroi_b200_forward(None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1.0, 0, None) == _cabi.ERR_INVALID_ARG
    assert L.rroi_b200_backward(None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1.0, 0, 1, None) == _cabi.ERR_INVALID_ARG
    assert L.rroi_b200_expand_idx(None, None, 1, 1, 1, 1, None) == _cabi.ERR_INVALID_ARG
    assert L.RROIAlignForwardLaucher(None, 1.0, 1, 1, 1, 1, 1, 1, None, None, None, None, None) == 0
    assert L.RROIAlignBackwardLaucher(None, 1.0, 1, 1, 1, 1, 1, 1, 1, None, None, None, None, None) == 0
    fake = ctypes.c_void_p(16)
    assert L.rroi_b200_forward(fake, fake, fake, None, None, 1, 0, 1, 1, 1, 1, 1, 1.0, 0, None) == _cabi.ERR_INVALID_ARG  # batch 0
    assert L.rroi_b200_forward(fake, fake, fake, fake, None, 1, 1, 1, 1, 1, 1, 1, 1.0, 0, None) == _cabi.ERR_INVALID_ARG  # half an idx pair
    assert L.rroi_b200_forward(fake, fake, fake, None, None, 1, 1, 1, 1, 1, 1, 1, 1.0, 5, None) == _cabi.ERR_INVALID_ARG  # layout
    assert L.rroi_b200_forward(fake, fake, fake, None, None, 0, 1, 1, 1, 1, 1, 1, 1.0, 0, None) == _cabi.OK               # no RoIs: nothing to do
    # per-call options are validated before anything touches CUDA; there is no process-global tuning state
    assert not hasattr(L, "rroi_b200_set_tuning")
    ok = _cabi.opts(concurrency=8, variant=5)
    assert L.rroi_b200_forward_opt(fake, fake, None, fake, None, None, 0, 1, 1, 1, 1, 1, 1, 1.0, 0, ctypes.byref(ok), None) == _cabi.OK
    for bad in (_cabi.opts(nchw_cg=3), _cabi.opts(variant=99), _cabi.opts(bwd_mode=7), _cabi.opts(nchw_tma=6), _cabi.opts(concurrency=-1)):
        assert L.rroi_b200_forward_opt(fake, fake, None, fake, None, None, 0, 1, 1, 1, 1, 1, 1, 1.0, 0, ctypes.byref(bad), None) == _cabi.ERR_INVALID_ARG
        assert L.rroi_b200_backward_opt(fake, fake, None, None, fake, 0, 1, 1, 1, 1, 1, 1, 1.0, 0, 0, ctypes.byref(bad), None) == _cabi.ERR_INVALID_ARG
    flags = _cabi.opts()
    flags.flags = 0x80
    assert L.rroi_b200_forward_opt(fake, fake, None, fake, None, None, 0, 1, 1, 1, 1, 1, 1, 1.0, 0, ctypes.byref(flags), None) == _cabi.ERR_INVALID_ARG
    short = _cabi.opts(variant=99)
    short.size = 12                                    # an older caller whose struct ends before `variant`: reads as 0
    assert L.rroi_b200_forward_opt(fake, fake, None, fake, None, None, 0, 1, 1, 1, 1, 1, 1, 1.0, 0, ctypes.byref(short), None) == _cabi.OK
    assert L.rroi_b200_forward_opt(fake, fake, ctypes.c_void_p(8), fake, None, None, 1, 1, 1, 1, 1, 1, 1, 1.0, 0, None, None) == _cabi.ERR_INVALID_ARG  # misaligned xform
    assert L.rroi_b200_roi_xform(None, None, 1, 8, 0.25, None) == _cabi.ERR_INVALID_ARG
    assert b"invalid" in L.rroi_b200_strerror(_cabi.ERR_INVALID_ARG)


def test_folded_heads_entry_points_reject_bad_arguments_without_gpu():
    """fots_b200_heads_merged / _gather_nhwc_bf16: argument checks run before anything touches CUDA."""
    from fots.pytorch_b200 import _cabi
    L = _cabi.lib()
    vp, i = ctypes.c_void_p, ctypes.c_int
    L.fots_b200_heads_gather_nhwc_bf16.restype = i
    L.fots_b200_heads_gather_nhwc_bf16.argtypes = [vp, i] + [vp] * 7 + [i] * 6 + [vp]
    L.fots_b200_heads_merged_nhwc_bf16.restype = i
    L.fots_b200_heads_merged_nhwc_bf16.argtypes = [vp] * 9 + [i] * 7 + [vp]
    a = ctypes.c_void_p(64)                                                    # aligned, never dereferenced
    bad = _cabi.ERR_INVALID_ARG
    assert L.fots_b200_heads_gather_nhwc_bf16(None, 128, a, a, a, a, a, a, a, 1, 8, 8, 4, 4, 64, None) == bad      # no tap map
    assert L.fots_b200_heads_gather_nhwc_bf16(a, 64, a, a, a, a, a, a, a, 1, 8, 8, 4, 4, 64, None) == bad         # fewer than 72 tap channels
    assert L.fots_b200_heads_gather_nhwc_bf16(a, 76, a, a, a, a, a, a, a, 1, 8, 8, 4, 4, 64, None) == bad         # not whole 16-byte vectors
    assert L.fots_b200_heads_gather_nhwc_bf16(a, 128, a, a, a, a, a, a, a, 1, 8, 8, 4, 4, 32, None) == bad        # s must have 64 channels
    assert L.fots_b200_heads_gather_nhwc_bf16(a, 128, ctypes.c_void_p(16), a, a, a, a, a, a, 1, 8, 8, 4, 4, 64, None) == bad   # 32-byte loads of s
    assert L.fots_b200_heads_gather_nhwc_bf16(a, 128, a, a, a, a, a, a, a, 0, 8, 8, 4, 4, 64, None) == bad        # empty batch
    assert L.fots_b200_heads_merged_nhwc_bf16(a, a, a, a, a, a, a, a, a, 1, 8, 8, 128, 64, 4, 4, None) == bad     # d must have 256 channels
    assert L.fots_b200_heads_merged_nhwc_bf16(a, a, a, a, None, a, a, a, a, 1, 8, 8, 256, 64, 4, 4, None) == bad  # no gate


def test_missing_library_fails_loudly(monkeypatch):
    from fots.pytorch_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/librroi_b200.so")
    with pytest.raises(ImportError, match="no CPU/PyTorch fallback"):
        _cabi.lib()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fots/ or rroi_align/ may reference it."""
    for base in ("fots", "rroi_align"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                    txt = open(os.path.join(d, f)).read()
                    assert "oracle" not in txt.lower(), os.path.join(d, f)


def test_module_surface_matches_reference():
    from rroi_align.modules.rroi_align import _RRoiAlign
    from rroi_align.functions.rroi_align import RRoiAlignFunction
    import fots.pytorch_b200 as P
    assert P._RRoiAlign is _RRoiAlign and P.RRoiAlignFunction is RRoiAlignFunction
    m = _RRoiAlign("11", 352.0, "0.25")                       # reference casts: int(), int(), float()
    assert (m.pooled_height, m.pooled_width, m.spatial_scale) == (11, 352, 0.25)
    assert isinstance(m, torch.nn.Module) and len(list(m.parameters())) == 0
    fn = RRoiAlignFunction(8, 64, 0.25)
    assert (fn.pooled_height, fn.pooled_width, fn.spatial_scale, fn.feature_size) == (8, 64, 0.25, None)
    assert fn.idx_x is None and fn.idx_y is None and fn.rois is None


def test_no_cpu_fallback():
    from fots.pytorch_b200 import _RRoiAlign, rroi_align
    f, r = torch.zeros(1, 3, 8, 8), torch.zeros(2, 6)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _RRoiAlign(4, 8, 1.0)(f, r)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rroi_align(f, r, 4, 8, 1.0)


def test_layout_canonicalisation():
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.rroi_align import _layout
    x = torch.zeros(2, 8, 5, 7)
    assert _layout.canonical(x)[1] == _cabi.LAYOUT_NCHW
    xc = x.contiguous(memory_format=torch.channels_last)
    t, lay = _layout.canonical(xc)
    assert lay == _cabi.LAYOUT_NHWC and t.data_ptr() == xc.data_ptr()
    sl = x[:, ::2]                                            # neither: falls back to a contiguous copy
    t, lay = _layout.canonical(sl)
    assert lay == _cabi.LAYOUT_NCHW and t.is_contiguous()
    e = _layout.empty((3, 8, 2, 4), _cabi.LAYOUT_NHWC, x)
    assert e.is_contiguous(memory_format=torch.channels_last) and e.shape == (3, 8, 2, 4)
    assert _layout.as_layout(x, _cabi.LAYOUT_NHWC).is_contiguous(memory_format=torch.channels_last)


def test_workload_accounting():
    import workloads as WL
    feats, rois, ph, pw, scale = WL.cfg1(64)
    assert feats.shape == (1, 64, 180, 320) and rois.shape == (64, 6) and (ph, pw, scale) == (8, 64, 0.25)
    v = WL.valid_counts(rois, ph, pw)
    assert abs(v.sum() / (64 * 8 * 64) - 0.639) < 0.005                       # 63.9 % valid
    assert abs(WL.algorithmic_bytes_fwd(rois, 64, ph, pw) / 1e6 - 13.75) < 0.05  # 13.75 MB
    r = WL.batch_rois(3, 64)
    assert r.shape == (192, 6) and set(np.unique(r[:, 0])) == {0.0, 1.0, 2.0}
