"""Memory-layout plumbing between torch tensors and the C ABI's layout flag.

Logical shapes are always the reference's ([B,C,H,W] features, [N,C,PH,PW] pooled).  Physically a
tensor is either contiguous NCHW (RROI_B200_LAYOUT_NCHW) or channels_last, i.e. NHWC in memory
(RROI_B200_LAYOUT_NHWC); the pooled output and the feature gradient follow the features' layout.
"""
import torch

from .. import _cabi


def canonical(t):
    """Return (tensor usable as-is by the ABI, layout flag) for a 4-d tensor."""
    if t.is_contiguous():
        return t, _cabi.LAYOUT_NCHW
    if t.is_contiguous(memory_format=torch.channels_last):
        return t, _cabi.LAYOUT_NHWC
    return t.contiguous(), _cabi.LAYOUT_NCHW


def as_layout(t, layout):
    if layout == _cabi.LAYOUT_NHWC:
        # for C == 1 both formats coincide; contiguous(memory_format=...) is then a no-op
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()


def empty(shape, layout, like):
    fmt = torch.channels_last if layout == _cabi.LAYOUT_NHWC else torch.contiguous_format
    return torch.empty(shape, dtype=torch.float32, device=like.device, memory_format=fmt)
