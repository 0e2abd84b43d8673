"""Shared helpers for the parity tests: run the product path (through the Python host API, which
calls the C ABI) and compare bit patterns."""
import ctypes

import numpy as np


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    neq = (bits(a) != bits(b)) & ~both_nan
    if neq.any():
        idx = np.argwhere(neq)
        first = tuple(idx[0])
        raise AssertionError("%s: %d / %d elements differ bitwise; first at %s: %r vs %r"
                             % (what, len(idx), a.size, first, a[first], b[first]))


def assert_close_rel(a, b, rel, what=""):
    """|a-b| <= rel * max(|b|_inf-scale, |b|) elementwise-ish: normalised by the tensor's max magnitude
    for near-zero entries (sum-order noise of fp32 atomics), by |b| elsewhere."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, what
    scale = np.maximum(np.abs(b), np.abs(b).max() * 1e-3 + 1e-30)
    err = np.abs(a - b) / scale
    assert err.max() <= rel, "%s: max rel err %.3e > %.1e at %s" % (
        what, err.max(), rel, np.unravel_index(err.argmax(), err.shape))


def to_cuda(x, device, channels_last=False):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)
    if channels_last:
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def run_new_forward(feats, rois, ph, pw, scale, device, channels_last=False, opts=None, with_xform=False):
    """Product forward via fots.pytorch_b200 (-> rroi_b200_forward).  Returns numpy (out NCHW-logical,
    idx_x, idx_y compact [N,PH,PW])."""
    from fots.pytorch_b200.rroi_align.functions.rroi_align import forward_raw
    f = to_cuda(feats, device, channels_last)
    r = to_cuda(rois, device)
    xf = None
    if with_xform:
        from fots.pytorch_b200.rroi_align.functions.rroi_align import roi_xform
        xf = roi_xform(r, ph, scale)
    out, ix, iy, _ = forward_raw(f, r, ph, pw, scale, want_idx=True, opts=opts, xform=xf)
    return out.cpu().numpy(), ix.cpu().numpy(), iy.cpu().numpy()


def run_new_backward(top_diff, rois, idx, feature_size, scale, device, channels_last=False, opts=None):
    """Product backward via rroi_b200_backward.  idx = (idx_x, idx_y) compact numpy or None (recompute)."""
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.rroi_align.functions.rroi_align import backward_raw
    g = to_cuda(top_diff, device, channels_last)
    r = to_cuda(rois, device)
    ix = iy = None
    if idx is not None:
        ix, iy = to_cuda(idx[0], device), to_cuda(idx[1], device)
    layout = _cabi.LAYOUT_NHWC if channels_last else _cabi.LAYOUT_NCHW
    return backward_raw(g, r, ix, iy, tuple(feature_size), scale, layout, opts=opts).cpu().numpy()


def run_legacy_forward(feats, rois, ph, pw, scale, device, with_idx=True):
    """RROIAlignForwardLaucher (the reference's C symbol) exported by librroi_b200.so, called exactly
    like rroi_align_cuda.c:37-41 calls it -- but on NON-zeroed buffers, to prove every element is written."""
    import torch
    from fots.pytorch_b200 import _cabi
    f = to_cuda(feats, device)
    r = to_cuda(rois, device)
    B, C, H, W = f.shape
    N = r.shape[0]
    out = torch.full((N, C, ph, pw), 7.0, device=device)
    ix = torch.full_like(out, 7.0) if with_idx else None
    iy = torch.full_like(out, 7.0) if with_idx else None
    rc = _cabi.lib().RROIAlignForwardLaucher(
        f.data_ptr(), ctypes.c_float(scale), N, H, W, C, ph, pw, r.data_ptr(), out.data_ptr(),
        ix.data_ptr() if with_idx else None, iy.data_ptr() if with_idx else None,
        torch.cuda.current_stream(device).cuda_stream)
    assert rc == 1
    torch.cuda.synchronize(device)
    if with_idx:
        return out.cpu().numpy(), ix.cpu().numpy(), iy.cpu().numpy()
    return out.cpu().numpy(), None, None


def run_legacy_backward(top_diff, rois, idx_x_full, idx_y_full, feature_size, scale, device):
    import torch
    from fots.pytorch_b200 import _cabi
    g = to_cuda(top_diff, device)
    r = to_cuda(rois, device)
    ix, iy = to_cuda(idx_x_full, device), to_cuda(idx_y_full, device)
    B, C, H, W = feature_size
    N, _, ph, pw = g.shape
    grad = torch.zeros((B, C, H, W), device=device)
    rc = _cabi.lib().RROIAlignBackwardLaucher(
        g.data_ptr(), ctypes.c_float(scale), B, N, H, W, C, ph, pw, r.data_ptr(), grad.data_ptr(),
        ix.data_ptr(), iy.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
    assert rc == 1
    torch.cuda.synchronize(device)
    return grad.cpu().numpy()


def expand_idx(idx_compact, channels):
    """[N,PH,PW] -> the reference's [N,C,PH,PW]."""
    return np.repeat(idx_compact[:, None, :, :], channels, axis=1)
