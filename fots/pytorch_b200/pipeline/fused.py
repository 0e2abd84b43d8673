"""Host side of the fused channels-last InstanceNorm(+affine)(+residual)+leaky-ReLU kernel
(include/fots_b200_pipeline.h: fots_b200_instnorm_nhwc_bf16).  Used by pipeline.nets on the CUDA bf16
channels-last inference path; everywhere else (CPU, fp32, training) the networks run torch's own ops, which
are the definition the fused kernel is tested against."""
import ctypes
import os

import torch

from .. import _cabi



def _lib():
    L = _cabi.lib()
    if not getattr(L, "_instnorm_bound", False):
        i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        L.fots_b200_instnorm_nhwc_bf16.restype = i
        L.fots_b200_instnorm_nhwc_bf16.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, f, f, i, vp]
        L.fots_b200_instnorm_apply_nhwc_bf16.restype = i
        L.fots_b200_instnorm_apply_nhwc_bf16.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, f, f, i, vp]
        L.fots_b200_instnorm_set_single_pass.restype = i
        L.fots_b200_instnorm_set_single_pass.argtypes = [i]
        L.fots_b200_maxpool_h2_nhwc_bf16.restype = i
        L.fots_b200_maxpool_h2_nhwc_bf16.argtypes = [vp, vp, i, i, i, i, vp]
        L.fots_b200_fpn_merge_nhwc_bf16.restype = i
        L.fots_b200_fpn_merge_nhwc_bf16.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
        L.fots_b200_fpn_merge_prob_nhwc_bf16.restype = i
        L.fots_b200_fpn_merge_prob_nhwc_bf16.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
        L._instnorm_bound = True
    return L


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def eligible(x, residual=None, weight=None, bias=None):
    """The fused kernels are inference-only: under autograd (grad mode on and ANY of the activation, the residual or the
    norm's affine parameters requiring grad) the caller must take torch's differentiable ops."""
    ok = (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.size(1) % 8 == 0 and x.size(1) <= 1024
          and x.is_contiguous(memory_format=torch.channels_last) and not _needs_grad(x, residual, weight, bias))
    if ok and residual is not None:
        ok = (residual.dtype == torch.bfloat16 and residual.shape == x.shape
              and residual.is_contiguous(memory_format=torch.channels_last))
    return ok


def set_single_pass(mode):
    """A/B switch (sweeps, tests): 0 / False = the three-launch two-pass InstanceNorm everywhere; 1 / True = automatic (the
    default: single-pass cluster kernel for instances of <= 40 KB per CTA); 2 = single-pass whenever the instance fits."""
    _cabi.check(_lib().fots_b200_instnorm_set_single_pass(int(mode)), "fots_b200_instnorm_set_single_pass")


def workspace(device, numel):
    """fp64 statistics workspace [B, C, 2], allocated per call from torch's caching allocator (stream-ordered; inside a
    CUDA-graph capture it belongs to that graph's pool).  No module-level cache: a buffer first allocated during a
    capture must not be shared with eager work on a recycled stream handle."""
    return torch.empty(max(int(numel), 1), dtype=torch.float64, device=device)


def instnorm_act(x, weight, bias, eps, slope, residual=None, crelu=False, stats=None):
    """act(IN(x) * weight + bias [+ residual]); crelu=True: act(IN(concat(x, -x))) with [2C] weight/bias.
    stats: the [B, C, 2] fp64 sums a producer already accumulated (conv.conv2d(..., stats=True)) -- the statistics
    pass over x is then skipped."""
    B, C, H, W = x.shape
    cout = 2 * C if crelu else C
    y = torch.empty((B, cout, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    ws = stats if stats is not None else workspace(x.device, B * C * 2)
    w = weight.float().contiguous() if weight is not None else None
    b = bias.float().contiguous() if bias is not None else None
    fn = _lib().fots_b200_instnorm_apply_nhwc_bf16 if stats is not None else _lib().fots_b200_instnorm_nhwc_bf16
    with torch.cuda.device(x.device):
        st = fn(x.data_ptr(), y.data_ptr(), w.data_ptr() if w is not None else None,
                b.data_ptr() if b is not None else None, residual.data_ptr() if residual is not None else None,
                ws.data_ptr(), B, H * W, C, float(eps), float(slope), 1 if crelu else 0, stream)
    _cabi.check(st, "fots_b200_instnorm_nhwc_bf16")
    return y


def instnorm_stats(x):
    """The statistics pass alone: fp64 [B, C, 2] per-image per-channel sum / sum of squares of x (bf16 channels-last), for a
    consumer that normalises on load (conv.dwconv_norm)."""
    B, C, H, W = x.shape
    ws = workspace(x.device, B * C * 2)
    L = _lib()
    L.fots_b200_instnorm_stats_nhwc_bf16.restype = ctypes.c_int
    L.fots_b200_instnorm_stats_nhwc_bf16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    with torch.cuda.device(x.device):
        st = L.fots_b200_instnorm_stats_nhwc_bf16(x.data_ptr(), ws.data_ptr(), B, H * W, C, torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(st, "fots_b200_instnorm_stats_nhwc_bf16")
    return ws


# ---- the same normalisation under autograd (training step) ------------------------------------------------------------
TRAIN_KERNELS = os.environ.get("FOTS_B200_TRAIN_IN", "1") != "0"   # A/B switch: hand-written InstanceNorm forward + backward in training


def train_eligible(x, residual=None):
    """InstanceNorm (+ residual) + activation through this repo's forward AND backward kernels under autograd: bf16
    channels-last CUDA activations (what the bf16-autocast training step hands over)."""
    ok = (TRAIN_KERNELS and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.size(1) % 8 == 0 and x.size(1) <= 1024
          and x.size(0) <= 65535 and x.is_contiguous(memory_format=torch.channels_last))
    if ok and residual is not None:
        ok = (residual.dtype == torch.bfloat16 and residual.shape == x.shape and residual.is_cuda
              and residual.is_contiguous(memory_format=torch.channels_last))
    return ok


class _InstNormActFn(torch.autograd.Function):
    """y = act(IN(x) * weight + bias [+ residual]) with fots_b200_instnorm_{stats,apply}_nhwc_bf16 forward and
    fots_b200_instnorm_bwd_nhwc_bf16 backward (two HBM passes each way; torch's instance_norm copies a channels-last tensor to
    NCHW and back around its batch-norm kernels in both directions)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, eps, slope):
        ws = instnorm_stats(x)
        y = instnorm_act(x, weight, bias, eps, slope, residual, stats=ws)
        ctx.save_for_backward(x, y, weight, ws)
        ctx.eps, ctx.slope, ctx.has_res, ctx.has_affine = float(eps), float(slope), residual is not None, weight is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, ws = ctx.saved_tensors
        B, C, H, W = x.shape
        if dy.dtype != torch.bfloat16 or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if ctx.has_res else None
        wsb = workspace(x.device, B * C * 2)
        w = weight.float().contiguous() if weight is not None else None
        L = _lib()
        L.fots_b200_instnorm_bwd_nhwc_bf16.restype = ctypes.c_int
        L.fots_b200_instnorm_bwd_nhwc_bf16.argtypes = [ctypes.c_void_p] * 8 + [ctypes.c_int] * 3 + [ctypes.c_float] * 2 + [ctypes.c_void_p]
        with torch.cuda.device(x.device):
            st = L.fots_b200_instnorm_bwd_nhwc_bf16(x.data_ptr(), y.data_ptr(), dy.data_ptr(), w.data_ptr() if w is not None else None,
                                                    ws.data_ptr(), wsb.data_ptr(), dx.data_ptr(), dres.data_ptr() if dres is not None else None,
                                                    B, H * W, C, ctx.eps, ctx.slope, torch.cuda.current_stream(x.device).cuda_stream)
        _cabi.check(st, "fots_b200_instnorm_bwd_nhwc_bf16")
        dgamma = dbeta = None
        if ctx.has_affine:
            sums = wsb[:B * C * 2].view(B, C, 2).sum(0)                   # [C, 2]: (sum g, sum g * xhat) over the batch
            dbeta, dgamma = sums[:, 0].to(weight.dtype), sums[:, 1].to(weight.dtype)
        return dx, dgamma, dbeta, dres, None, None


def instnorm_act_train(x, weight, bias, eps, slope, residual=None):
    """Differentiable act(IN(x) * weight + bias [+ residual]); caller checks train_eligible(x, residual)."""
    return _InstNormActFn.apply(x, weight, bias, residual, eps, slope)


class _CReLUNormFn(torch.autograd.Function):
    """y = act(IN(concat(x, -x)) * weight + bias) (tools/models.py:41-48 CReLU_IN) with this repo's forward kernels and
    fots_b200_instnorm_crelu_bwd_nhwc_bf16 backward; weight / bias have 2C entries."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, slope):
        ws = instnorm_stats(x)
        y = instnorm_act(x, weight, bias, eps, slope, crelu=True, stats=ws)
        ctx.save_for_backward(x, y, weight, ws)
        ctx.eps, ctx.slope, ctx.has_affine = float(eps), float(slope), weight is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, ws = ctx.saved_tensors
        B, C, H, W = x.shape
        if dy.dtype != torch.bfloat16 or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        wsb = workspace(x.device, B * 2 * C * 2)
        w = weight.float().contiguous() if weight is not None else None
        L = _lib()
        L.fots_b200_instnorm_crelu_bwd_nhwc_bf16.restype = ctypes.c_int
        L.fots_b200_instnorm_crelu_bwd_nhwc_bf16.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_int] * 3 + [ctypes.c_float] * 2 + [ctypes.c_void_p]
        with torch.cuda.device(x.device):
            st = L.fots_b200_instnorm_crelu_bwd_nhwc_bf16(x.data_ptr(), y.data_ptr(), dy.data_ptr(), w.data_ptr() if w is not None else None,
                                                          ws.data_ptr(), wsb.data_ptr(), dx.data_ptr(), B, H * W, C, ctx.eps, ctx.slope,
                                                          torch.cuda.current_stream(x.device).cuda_stream)
        _cabi.check(st, "fots_b200_instnorm_crelu_bwd_nhwc_bf16")
        dgamma = dbeta = None
        if ctx.has_affine:
            sums = wsb[:B * 2 * C * 2].view(B, 2 * C, 2).sum(0)
            dbeta, dgamma = sums[:, 0].to(weight.dtype), sums[:, 1].to(weight.dtype)
        return dx, dgamma, dbeta, None, None


def crelu_norm_train(x, weight, bias, eps, slope):
    """Differentiable act(IN(concat(x, -x)) * weight + bias); caller checks train_eligible(x) and C <= 512."""
    return _CReLUNormFn.apply(x, weight, bias, eps, slope)


class _UpsampleFn(torch.autograd.Function):
    """Bilinear (align_corners = True) upsampling of a bf16 channels-last map to `size` with this repo's kernels in both
    directions (forward: the a_lo-only form of fots_b200_fpn_merge_nhwc_bf16; backward: fots_b200_upsample_bilinear_bwd_nhwc_bf16)."""

    @staticmethod
    def forward(ctx, x, H, W):
        ctx.lo_shape = tuple(x.shape)
        return fpn_merge(a_lo=x, size=(H, W))

    @staticmethod
    def backward(ctx, dy):
        B, C, h, w = ctx.lo_shape
        H, W = dy.shape[2], dy.shape[3]
        if dy.dtype != torch.bfloat16 or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dx = torch.empty((B, C, h, w), dtype=torch.bfloat16, device=dy.device, memory_format=torch.channels_last)
        L = _lib()
        L.fots_b200_upsample_bilinear_bwd_nhwc_bf16.restype = ctypes.c_int
        L.fots_b200_upsample_bilinear_bwd_nhwc_bf16.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
        with torch.cuda.device(dy.device):
            st = L.fots_b200_upsample_bilinear_bwd_nhwc_bf16(dy.data_ptr(), dx.data_ptr(), B, h, w, H, W, C,
                                                             torch.cuda.current_stream(dy.device).cuda_stream)
        _cabi.check(st, "fots_b200_upsample_bilinear_bwd_nhwc_bf16")
        return dx, None, None


def upsample_train(x, size):
    """Differentiable bilinear (align_corners) upsampling; caller checks train_eligible(x)."""
    return _UpsampleFn.apply(x, int(size[0]), int(size[1]))


def _cl_bf16(t):
    return (t is None) or (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4
                           and t.is_contiguous(memory_format=torch.channels_last)
                           and not (torch.is_grad_enabled() and t.requires_grad))


def merge_eligible(*tensors):
    ts = [t for t in tensors if t is not None]
    return len(ts) > 0 and all(_cl_bf16(t) for t in ts) and ts[0].size(1) % 8 == 0


def fpn_merge(a_lo=None, c_hi=None, b_hi=None, gate_logits_lo=None, size=None, gate_prob_lo=None):
    """y = (upsample(a_lo) | c_hi) + b_hi * (upsample(sigmoid(gate_logits_lo)) | 1); bilinear, align_corners=True.
    a_lo [B,C,h,w] / c_hi, b_hi [B,C,H,W] / gate_logits_lo [B,1,h,w]; size=(H, W) when only low-res inputs are given.
    gate_prob_lo: the gate as sigmoid(logits) already (bf16 [B,1,h,w], conv.conv1x1_to1(..., sigmoid=True)) instead of logits."""
    if gate_prob_lo is not None and gate_logits_lo is not None:
        raise ValueError("fpn_merge: pass the gate either as logits or as probabilities")
    prob = gate_prob_lo is not None
    if prob:
        gate_logits_lo = gate_prob_lo
    hi = c_hi if c_hi is not None else b_hi
    lo = a_lo if a_lo is not None else gate_logits_lo
    H, W = (hi.shape[2], hi.shape[3]) if hi is not None else size
    ref = hi if hi is not None else lo
    B, C = ref.shape[0], (a_lo if a_lo is not None else hi).shape[1]
    h, w = (lo.shape[2], lo.shape[3]) if lo is not None else (H, W)
    y = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=ref.device, memory_format=torch.channels_last)
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(ref.device):
        fn = _lib().fots_b200_fpn_merge_prob_nhwc_bf16 if prob else _lib().fots_b200_fpn_merge_nhwc_bf16
        st = fn(ptr(a_lo), ptr(c_hi), ptr(b_hi), ptr(gate_logits_lo), y.data_ptr(),
                B, h, w, H, W, C, torch.cuda.current_stream(ref.device).cuda_stream)
    _cabi.check(st, "fots_b200_fpn_merge_nhwc_bf16")
    return y


def maxpool_h2(x):
    """MaxPool2d((2, 1), stride (2, 1)) of a bf16 channels-last tensor in one HBM pass (fots_b200_maxpool_h2_nhwc_bf16);
    the caller checks `eligible(x)` and H >= 2."""
    N, C, H, W = x.shape
    y = torch.empty((N, C, H // 2, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        st = _lib().fots_b200_maxpool_h2_nhwc_bf16(x.data_ptr(), y.data_ptr(), N, H, W, C,
                                                   torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(st, "fots_b200_maxpool_h2_nhwc_bf16")
    return y
