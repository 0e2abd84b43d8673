"""autograd face of RoIRotate.  Mirrors rroi_align/functions/rroi_align.py:6-40 of the reference.

The reference class is a legacy instance-style torch.autograd.Function (`RRoiAlignFunction(ph, pw,
scale)(features, rois)`), which torch >= 1.3 rejects.  `RRoiAlignFunction` here keeps that call form
and the attribute names (pooled_width, pooled_height, spatial_scale, feature_size, rois, idx_x,
idx_y) on top of a static Function, and drops the C-fold redundant buffers: the sample centres are
kept compact ([N, PH, PW]) and only expanded to the reference's [N, C, PH, PW] when somebody reads
`.idx_x` / `.idx_y`.
"""
import torch

from .. import _layout
from ... import _cabi


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _check_inputs(features, rois):
    if not (isinstance(features, torch.Tensor) and isinstance(rois, torch.Tensor)):
        raise TypeError("rroi_align: features and rois must be tensors")
    if not features.is_cuda or not rois.is_cuda:
        # reference: the non-CUDA branch (functions/rroi_align.py:22-25) dies on an undefined name;
        # backward asserts is_cuda (:33).  No CPU fallback here either.
        raise RuntimeError("rroi_align: features and rois must be CUDA tensors (there is no CPU path)")
    if features.device != rois.device:
        raise RuntimeError("rroi_align: features and rois are on different devices")
    if features.dtype != torch.float32 or rois.dtype != torch.float32:
        raise TypeError("rroi_align: fp32 tensors required (got %s, %s)" % (features.dtype, rois.dtype))
    if features.dim() != 4:
        raise ValueError("rroi_align: features must be [B, C, H, W]")
    if rois.dim() != 2 or rois.size(1) != 6:
        # reference: rroi_align_cuda.c:23-26 returns 0 (ignored by Python -> silent zeros)
        raise ValueError("rroi_align: rois must be [N, 6] = [batch_idx, cx, cy, h, w, angle_deg]")


def roi_xform(rois, pooled_height, spatial_scale):
    """[N, 8] per-RoI transform table (rroi_b200_roi_xform) for `forward_raw(..., xform=...)`: a producer of RoI rows
    can compute it once so that the forward's CTAs start at the bin geometry."""
    rois = rois.contiguous()
    out = torch.empty((rois.size(0), 8), dtype=torch.float32, device=rois.device)
    with torch.cuda.device(rois.device):
        st = _cabi.lib().rroi_b200_roi_xform(rois.data_ptr(), out.data_ptr(), rois.size(0), int(pooled_height),
                                             float(spatial_scale), _stream(rois.device))
        _cabi.check(st, "rroi_b200_roi_xform")
    return out


def forward_raw(features, rois, pooled_height, pooled_width, spatial_scale, want_idx=True, opts=None, xform=None):
    """One forward launch.  Returns (pooled, idx_x, idx_y, layout); idx_* are compact [N, PH, PW] or None.
    opts: _cabi.opts(...) per-call launch options (None = defaults); xform: optional table from roi_xform()."""
    _check_inputs(features, rois)
    ph, pw = int(pooled_height), int(pooled_width)
    if ph <= 0 or pw <= 0:
        raise ValueError("rroi_align: pooled size must be positive")
    features, layout = _layout.canonical(features)
    rois = rois.contiguous()
    B, C, H, W = features.shape
    N = rois.size(0)
    with torch.cuda.device(features.device):
        pooled = _layout.empty((N, C, ph, pw), layout, features)
        if want_idx:
            idx_x = torch.empty((N, ph, pw), dtype=torch.float32, device=features.device)
            idx_y = torch.empty_like(idx_x)
        else:
            idx_x = idx_y = None
        if N > 0:
            st = _cabi.lib().rroi_b200_forward_opt(
                features.data_ptr(), rois.data_ptr(), xform.data_ptr() if xform is not None else None, pooled.data_ptr(),
                idx_x.data_ptr() if want_idx else None, idx_y.data_ptr() if want_idx else None,
                N, B, C, H, W, ph, pw, float(spatial_scale), layout, _cabi.opts_ref(opts), _stream(features.device))
            _cabi.check(st, "rroi_b200_forward_opt")
    return pooled, idx_x, idx_y, layout


def backward_raw(grad_output, rois, idx_x, idx_y, feature_size, spatial_scale, layout, opts=None):
    """One backward launch (+ the zero-fill of the gradient map).  Returns grad wrt features."""
    if not grad_output.is_cuda:
        raise RuntimeError("rroi_align backward: grad_output must be a CUDA tensor")  # reference :33
    B, C, H, W = feature_size
    N, C2, ph, pw = grad_output.shape
    if C2 != C or N != rois.size(0):
        raise ValueError("rroi_align backward: grad_output shape does not match the forward")
    grad_output = _layout.as_layout(grad_output.float(), layout)
    with torch.cuda.device(grad_output.device):
        grad_input = _layout.empty((B, C, H, W), layout, grad_output)
        st = _cabi.lib().rroi_b200_backward_opt(
            grad_output.data_ptr() if N > 0 else None, rois.data_ptr() if N > 0 else None,
            idx_x.data_ptr() if idx_x is not None else None,
            idx_y.data_ptr() if idx_y is not None else None,
            grad_input.data_ptr(), N, B, C, H, W, ph, pw, float(spatial_scale), layout, 1,
            _cabi.opts_ref(opts), _stream(grad_output.device))
        _cabi.check(st, "rroi_b200_backward_opt")
    return grad_input


class _RRoiAlignOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, pooled_height, pooled_width, spatial_scale, holder, opts=None):
        want_idx = ctx.needs_input_grad[0] or holder is not None
        pooled, idx_x, idx_y, layout = forward_raw(features, rois, pooled_height, pooled_width,
                                                   spatial_scale, want_idx=want_idx, opts=opts)
        ctx.feature_size = tuple(features.shape)
        ctx.spatial_scale = float(spatial_scale)
        ctx.layout = layout
        if want_idx:
            ctx.save_for_backward(rois.contiguous(), idx_x, idx_y)
        else:
            ctx.save_for_backward(rois.contiguous())
        if holder is not None:
            holder._remember(features, rois, idx_x, idx_y, layout)
        return pooled

    @staticmethod
    def backward(ctx, grad_output):
        saved = ctx.saved_tensors
        rois = saved[0]
        idx_x, idx_y = (saved[1], saved[2]) if len(saved) == 3 else (None, None)
        grad_input = backward_raw(grad_output, rois, idx_x, idx_y, ctx.feature_size,
                                  ctx.spatial_scale, ctx.layout)
        return grad_input, None, None, None, None, None, None   # reference: (grad_input, None)


def rroi_align(features, rois, pooled_height, pooled_width, spatial_scale, concurrency=0, rois_ready=False):
    """Functional form: RoIRotate of `rois` over `features` -> [N, C, PH, PW] (differentiable in features).
    concurrency / rois_ready: per-call launch hints (rroi_b200_opts); results never depend on them."""
    opts = _cabi.opts(concurrency=concurrency, rois_ready=rois_ready) if (concurrency or rois_ready) else None
    return _RRoiAlignOp.apply(features, rois, int(pooled_height), int(pooled_width), float(spatial_scale), None, opts)


def rroi_align_bf16(features, rois, pooled_height, pooled_width, spatial_scale, opts=None):
    """Inference-only bf16 RoIRotate (rroi_b200_forward_bf16): bf16 channels-last features [B, C, H, W] ->
    bf16 channels-last pooled [N, C, PH, PW], C in {32, 64, 128, 256}.  Equal to
    rroi_align(features.float(), ...) rounded to bf16; not differentiable (training uses the fp32 op)."""
    if not (features.is_cuda and rois.is_cuda and features.dtype == torch.bfloat16 and rois.dtype == torch.float32):
        raise TypeError("rroi_align_bf16: CUDA bf16 features and fp32 rois required")
    if features.dim() != 4 or rois.dim() != 2 or rois.size(1) != 6:
        raise ValueError("rroi_align_bf16: features [B, C, H, W], rois [N, 6]")
    ph, pw = int(pooled_height), int(pooled_width)
    features = features.contiguous(memory_format=torch.channels_last)
    rois = rois.contiguous()
    B, C, H, W = features.shape
    N = rois.size(0)
    with torch.cuda.device(features.device):
        pooled = torch.empty((N, C, ph, pw), dtype=torch.bfloat16, device=features.device,
                             memory_format=torch.channels_last)
        if N > 0:
            st = _cabi.lib().rroi_b200_forward_bf16_opt(features.data_ptr(), rois.data_ptr(), None, pooled.data_ptr(), None, None,
                                                        N, B, C, H, W, ph, pw, float(spatial_scale), _cabi.opts_ref(opts),
                                                        _stream(features.device))
            _cabi.check(st, "rroi_b200_forward_bf16_opt")
    return pooled


class RRoiAlignFunction(object):
    """Call-compatible stand-in for the reference's legacy Function instance.

        fn = RRoiAlignFunction(pooled_height, pooled_width, spatial_scale)
        pooled = fn(features, rois)            # autograd-aware
        fn.idx_x, fn.idx_y                     # [N, C, PH, PW] sample centres, as the reference kept
        fn.forward(features, rois)             # legacy direct call (no graph), then fn.backward(grad)
    """

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = pooled_width
        self.pooled_height = pooled_height
        self.spatial_scale = spatial_scale
        self.feature_size = None
        self.rois = None
        self._idx = None      # (idx_x, idx_y) compact [N, PH, PW]
        self._layout = None

    def _remember(self, features, rois, idx_x, idx_y, layout):
        self.feature_size = features.size()
        self.rois = rois
        self._idx = (idx_x, idx_y)
        self._layout = layout

    def __call__(self, features, rois):
        return _RRoiAlignOp.apply(features, rois, int(self.pooled_height), int(self.pooled_width),
                                  float(self.spatial_scale), self, None)

    # -- legacy direct methods (reference: forward(ctx, features, rois) / backward(ctx, grad_output)) --
    def forward(self, features, rois):
        pooled, idx_x, idx_y, layout = forward_raw(features, rois, self.pooled_height, self.pooled_width,
                                                   self.spatial_scale, want_idx=True)
        self._remember(features, rois, idx_x, idx_y, layout)
        return pooled

    def backward(self, grad_output):
        assert self.feature_size is not None and grad_output.is_cuda   # reference :33
        idx_x, idx_y = self._idx
        grad_input = backward_raw(grad_output, self.rois.contiguous(), idx_x, idx_y,
                                  tuple(self.feature_size), self.spatial_scale, self._layout)
        return grad_input, None

    def _expanded(self, which):
        if self._idx is None or self._idx[which] is None:
            return None
        compact = self._idx[which]
        N, ph, pw = compact.shape
        C = self.feature_size[1]
        full = torch.empty((N, C, ph, pw), dtype=torch.float32, device=compact.device)
        with torch.cuda.device(compact.device):
            st = _cabi.lib().rroi_b200_expand_idx(compact.data_ptr(), full.data_ptr(), N, C, ph, pw,
                                                  _stream(compact.device))
            _cabi.check(st, "rroi_b200_expand_idx")
        return full

    @property
    def idx_x(self):
        return self._expanded(0)

    @property
    def idx_y(self):
        return self._expanded(1)
