"""Host side of the tcgen05 implicit-GEMM convolution (include/fots_b200_pipeline.h:
fots_b200_conv2d_nhwc_bf16; csrc/conv_tc.cu).  Used by pipeline.nets on the CUDA bf16 channels-last inference path
for the dense stride-1 convolutions (tools/models.py:336-366 forward_ocr, :142-166 BasicBlockIn); everything else
(CPU, fp32, training, strided / grouped / tiny-channel convolutions) stays on torch's own ops, which are the
definition this kernel is tested against."""
import ctypes
import os

import torch

from .. import _cabi

ENABLED = os.environ.get("FOTS_B200_TC_CONV", "1") != "0"   # A/B switch against the library convolution (sweeps)
# Convolutions followed by an InstanceNorm: run them here with the statistics accumulated in the epilogue?  Measured
# on B200 (profiles/r01_conv_tc_bench.txt) the bare library convolution + a separate statistics pass is still a
# little faster for these shapes (the statistics make the 4-warp epilogue the bottleneck), so the default is off.
FUSE_STATS = os.environ.get("FOTS_B200_TC_STATS", "0") != "0"
# Epilogue statistics only for SMALL outputs (MB of bf16 output; 0 = off): there the separate statistics pass is two
# launch-bound launches (memset + kernel, ~8 us for a few MB), which costs more than the longer epilogue.
STATS_MAX_MB = float(os.environ.get("FOTS_B200_TC_STATS_MAX_MB", "0"))
# How much of the networks runs on this kernel (A/B switch for the step-time sweeps, tools/profile_pipeline.py):
#   0 = only the convolutions whose activation it fuses (conv6/8/9, layer0_1[0]);
#   1 = + every other convolution of the recogniser (conv5/7/10_s in front of an InstanceNorm, conv11 with its 89 classes
#       padded to 128 output channels): forward_ocr then contains no library call;
#   2 = + the 3x3 convolutions of stages 1-2 (stride 1 and 2), every 1x1 convolution of the feeder (FPN laterals, separable
#       blocks' pointwise halves, up-convolutions, down-sampling branches with their BatchNorm folded) and the depthwise
#       3x3 convolutions (csrc/dwconv_kernels.cu).  Default: measured 4.79 ms per 8-image step against 4.77 ms at level 1.
LEVEL = int(os.environ.get("FOTS_B200_TC_LEVEL", "2"))
# A/B switch: compute the 2x upsampling of the top-down merge inside the depthwise kernel (1) or as its own kernel (0)
DW_UP = os.environ.get("FOTS_B200_DW_UP", "1") != "0"
# A/B switch: fold the last top-down level into the heads when the caller does not need the 256-channel map (inference)
MERGED_HEADS = os.environ.get("FOTS_B200_MERGED_HEADS", "1") != "0"
GATHER_HEADS = os.environ.get("FOTS_B200_GATHER_HEADS", "1") != "0"    # ... and the depthwise half of upconv2 (heads_gather)


def _lib():
    L = _cabi.lib()
    if not getattr(L, "_conv_bound", False):
        i, f, vp = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
        L.fots_b200_conv2d_nhwc_bf16.restype = i
        L.fots_b200_conv2d_nhwc_bf16.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, f, vp]
        L.fots_b200_conv2d_strided_nhwc_bf16.restype = i
        L.fots_b200_conv2d_strided_nhwc_bf16.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, i, f, vp]
        L.fots_b200_conv2d_stats_nhwc_bf16.restype = i
        L.fots_b200_conv2d_stats_nhwc_bf16.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]
        L.fots_b200_stem_conv3x3_c3_c16.restype = i
        L.fots_b200_stem_conv3x3_c3_c16.argtypes = [vp, vp, vp, vp, i, i, i, vp]
        L.fots_b200_stem_conv3x3_c3_c16_u8.restype = i
        L.fots_b200_stem_conv3x3_c3_c16_u8.argtypes = [vp, vp, vp, vp, i, i, i, vp]
        L.fots_b200_dwconv3x3_nhwc_bf16.restype = i
        L.fots_b200_dwconv3x3_nhwc_bf16.argtypes = [vp, vp, vp, i, i, i, i, i, vp]
        L.fots_b200_dwconv3x3_norm_nhwc_bf16.restype = i
        L.fots_b200_dwconv3x3_norm_nhwc_bf16.argtypes = [vp, vp, vp, vp, vp, vp, f, f, vp, i, i, i, i, i, vp]
        L.fots_b200_heads_nhwc_bf16.restype = i
        L.fots_b200_heads_nhwc_bf16.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp]
        L.fots_b200_conv_set_tile.restype = i
        L.fots_b200_conv_set_tile.argtypes = [i]
        L._conv_bound = True
    return L


def set_halo(mode):
    """-1 automatic, 0 never, 1 whenever the shape allows: halo reuse of the A operand (sweeps, tests)."""
    L = _lib()
    L.fots_b200_conv_set_halo.restype = ctypes.c_int
    L.fots_b200_conv_set_halo.argtypes = [ctypes.c_int]
    _cabi.check(L.fots_b200_conv_set_halo(int(mode)), "fots_b200_conv_set_halo")


def set_tile(bn):
    _cabi.check(_lib().fots_b200_conv_set_tile(int(bn)), "fots_b200_conv_set_tile")


def input_ok(x):
    """bf16 channels-last CUDA activations outside autograd: what the kernel consumes."""
    return (ENABLED and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
            and x.is_contiguous(memory_format=torch.channels_last) and not (torch.is_grad_enabled() and x.requires_grad))


def eligible(x, conv):
    """True when `conv` (an nn.Conv2d) applied to `x` can run on the tensor-core kernel."""
    w = conv.weight
    return (ENABLED and x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 4
            and x.is_contiguous(memory_format=torch.channels_last)
            and conv.stride in ((1, 1), (2, 2)) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.padding_mode == "zeros" and not isinstance(conv.padding, str)
            and w.size(1) % 64 == 0 and w.size(0) % 64 == 0 and w.size(2) <= 7 and w.size(3) <= 7
            and not (torch.is_grad_enabled() and (x.requires_grad or w.requires_grad
                                                  or (conv.bias is not None and conv.bias.requires_grad))))


def conv2d(x, weight, bias=None, padding=(0, 0), slope=1.0, stats=False, stride=1):
    """act(conv2d(x, weight, bias, stride (1 or 2), padding)) -> bf16 channels-last [N, Cout, Ho, Wo].
    x: bf16 channels-last [N, Cin, H, W]; weight: bf16 [Cout, Cin, R, S]; bias: fp32/bf16 [Cout] or None.
    stats=True (slope must be 1): returns (y, ws) where ws is the fp64 [N, Cout, 2] per-image sum / sum of squares
    of y accumulated by the epilogue, for fused.instnorm_act(y, ..., stats=ws)."""
    N, Cin, H, W = x.shape
    Cout, Cin_w, R, S = weight.shape
    if Cin_w != Cin:
        raise ValueError("conv2d: weight expects %d input channels, x has %d" % (Cin_w, Cin))
    ph, pw = (padding, padding) if isinstance(padding, int) else padding
    stride = stride[0] if isinstance(stride, (tuple, list)) else int(stride)
    Ho, Wo = (H + 2 * ph - R) // stride + 1, (W + 2 * pw - S) // stride + 1
    wk = weight if weight.is_contiguous(memory_format=torch.channels_last) else weight.contiguous(memory_format=torch.channels_last)
    b = None if bias is None else bias.float().contiguous()
    y = torch.empty((N, Cout, Ho, Wo), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        if stride != 1:
            if stats:
                raise ValueError("conv2d: fused statistics are only available at stride 1")
            st = _lib().fots_b200_conv2d_strided_nhwc_bf16(
                x.data_ptr(), wk.data_ptr(), b.data_ptr() if b is not None else None, y.data_ptr(),
                N, H, W, Cin, Cout, R, S, ph, pw, stride, float(slope), stream)
        elif stats:
            if slope != 1.0:
                raise ValueError("conv2d: stats=True computes the statistics of the raw convolution (slope must be 1)")
            from . import fused
            ws = fused.workspace(x.device, N * Cout * 2)
            st = _lib().fots_b200_conv2d_stats_nhwc_bf16(
                x.data_ptr(), wk.data_ptr(), b.data_ptr() if b is not None else None, y.data_ptr(), ws.data_ptr(),
                N, H, W, Cin, Cout, R, S, ph, pw, stream)
        else:
            st = _lib().fots_b200_conv2d_nhwc_bf16(
                x.data_ptr(), wk.data_ptr(), b.data_ptr() if b is not None else None, y.data_ptr(),
                N, H, W, Cin, Cout, R, S, ph, pw, float(slope), stream)
    _cabi.check(st, "fots_b200_conv2d_nhwc_bf16")
    return (y, ws) if stats else y


def conv_stats_small(conv, x, level=2):
    """conv(x) for a stride-1 convolution that is followed by an InstanceNorm -> (y, ws or None): ws = the epilogue's fp64
    statistics when the output is small enough for that to pay (STATS_MAX_MB), else None (caller runs the statistics pass)."""
    if LEVEL >= level and eligible(x, conv):
        N, _, H, W = x.shape
        small = N * conv.out_channels * H * W * 2 <= STATS_MAX_MB * 1e6
        if small and conv.stride == (1, 1) and 128 <= conv.out_channels <= 1024:
            return conv2d(x, conv.weight, conv.bias, conv.padding, 1.0, stats=True)
        return conv2d(x, conv.weight, conv.bias, conv.padding, 1.0, stride=conv.stride), None
    return conv(x), None


def apply(conv, x, slope=1.0, level=0):
    """conv(x) followed by leaky-ReLU(slope) through the tensor-core kernel when eligible (and LEVEL >= level), else torch."""
    if LEVEL >= level and eligible(x, conv):
        return conv2d(x, conv.weight, conv.bias, conv.padding, slope, stride=conv.stride)
    y = conv(x)
    if slope == 1.0:
        return y
    return torch.relu(y) if slope == 0.0 else torch.nn.functional.leaky_relu(y, slope)


def dw_eligible(x, conv):
    """Depthwise 3x3, pad 1, stride 1 or 2, no bias, C % 64 == 0, on bf16 channels-last activations outside autograd."""
    w = conv.weight
    C = conv.in_channels
    return (LEVEL >= 2 and input_ok(x) and w.dtype == torch.bfloat16 and conv.groups == C == conv.out_channels and C % 64 == 0
            and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride in ((1, 1), (2, 2))
            and conv.dilation == (1, 1) and conv.bias is None and conv.padding_mode == "zeros"
            and not (torch.is_grad_enabled() and w.requires_grad))


def dwconv(conv, x):
    """conv(x) for a depthwise 3x3 nn.Conv2d through fots_b200_dwconv3x3_nhwc_bf16 when eligible, else the module."""
    if not dw_eligible(x, conv):
        return conv(x)
    N, C, H, W = x.shape
    st = conv.stride[0]
    y = torch.empty((N, C, (H - 1) // st + 1, (W - 1) // st + 1), dtype=torch.bfloat16, device=x.device,
                    memory_format=torch.channels_last)
    w = conv.weight.reshape(C, 9)
    if not w.is_contiguous():
        w = w.contiguous()
    with torch.cuda.device(x.device):
        rc = _lib().fots_b200_dwconv3x3_nhwc_bf16(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, H, W, C, st,
                                                  torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(rc, "fots_b200_dwconv3x3_nhwc_bf16")
    return y


def dwconv_norm(conv, x, stats=None, norm=None, slope=1.0, stats_out=False):
    """Depthwise 3x3 `conv` with the InstanceNorms either side of it fused (fots_b200_dwconv3x3_norm_nhwc_bf16):
    stats / norm / slope: computes conv(act(norm(x))) WITHOUT materialising the normalised tensor -- x is the raw bf16
      channels-last input, `stats` = fused.instnorm_stats(x), the normalisation is applied to the staged tile;
    stats_out=True: also returns the fp64 [N, C, 2] sums of the output (for fused.instnorm_act(y, ..., stats=ws)).
    Caller checks dw_eligible(x, conv)."""
    from . import fused
    N, C, H, W = x.shape
    st = conv.stride[0]
    y = torch.empty((N, C, (H - 1) // st + 1, (W - 1) // st + 1), dtype=torch.bfloat16, device=x.device,
                    memory_format=torch.channels_last)
    w = conv.weight.reshape(C, 9)
    if not w.is_contiguous():
        w = w.contiguous()
    g = norm.weight.float().contiguous() if (norm is not None and norm.weight is not None) else None
    b = norm.bias.float().contiguous() if (norm is not None and norm.bias is not None) else None
    ws = fused.workspace(x.device, N * C * 2) if stats_out else None
    with torch.cuda.device(x.device):
        rc = _lib().fots_b200_dwconv3x3_norm_nhwc_bf16(x.data_ptr(), w.data_ptr(), y.data_ptr(),
                                                       stats.data_ptr() if stats is not None else None,
                                                       g.data_ptr() if g is not None else None,
                                                       b.data_ptr() if b is not None else None,
                                                       float(norm.eps) if norm is not None else 0.0, float(slope),
                                                       ws.data_ptr() if ws is not None else None,
                                                       N, H, W, C, st, torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(rc, "fots_b200_dwconv3x3_norm_nhwc_bf16")
    return (y, ws) if stats_out else y


def dwconv_up(conv, x_lo, size):
    """conv(bilinear_upsample(x_lo, size, align_corners=True)) for a depthwise 3x3 stride-1 `conv` without materialising the
    upsampled map (fots_b200_dwconv3x3_up_nhwc_bf16).  Caller checks dw_eligible(x_lo, conv) and conv.stride == (1, 1)."""
    N, C, h, w = x_lo.shape
    H, W = int(size[0]), int(size[1])
    y = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=x_lo.device, memory_format=torch.channels_last)
    wt = conv.weight.reshape(C, 9)
    if not wt.is_contiguous():
        wt = wt.contiguous()
    L = _lib()
    L.fots_b200_dwconv3x3_up_nhwc_bf16.restype = ctypes.c_int
    L.fots_b200_dwconv3x3_up_nhwc_bf16.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
    with torch.cuda.device(x_lo.device):
        rc = L.fots_b200_dwconv3x3_up_nhwc_bf16(x_lo.data_ptr(), wt.data_ptr(), y.data_ptr(), N, h, w, H, W, C,
                                                torch.cuda.current_stream(x_lo.device).cuda_stream)
    _cabi.check(rc, "fots_b200_dwconv3x3_up_nhwc_bf16")
    return y


def pack_pixel_pairs_s2(weight):
    """Weights of a 3x3 stride-2 pad-1 convolution with FEWER than 64 input channels (layer0's 32 -> 32, tools/models.py:252)
    re-expressed for the 64-channel kernel: two horizontally adjacent pixels are viewed as one pixel with 2*Cin channels
    (a free reinterpretation of a channels-last tensor), on the input AND on the output side.  Output pair xo' (pixels
    2xo', 2xo'+1) reads input pairs 2xo'-1 .. 2xo'+1, i.e. the result is again a 3x3 stride-2 pad-1 convolution,
    [2*Cout, 2*Cin, 3, 3], with W'[(e,o), (p,c), r, s'] = W[o, c, r, 2s' + p - 2e - 1] where that tap exists, else 0."""
    Cout, Cin, R, S = weight.shape
    assert R == 3 and S == 3
    w2 = torch.zeros((2 * Cout, 2 * Cin, 3, 3), dtype=weight.dtype, device=weight.device)
    for e in range(2):
        for p_ in range(2):
            for s2 in range(3):
                s1 = 2 * s2 + p_ - 2 * e - 1
                if 0 <= s1 <= 2:
                    w2[e * Cout:(e + 1) * Cout, p_ * Cin:(p_ + 1) * Cin, :, s2] = weight[:, :, :, s1]
    return w2.contiguous(memory_format=torch.channels_last)


def conv3x3_s2_pixel_pairs(x, w2):
    """x bf16 channels-last [B, Cin, H, W] (W % 4 == 0), w2 = pack_pixel_pairs_s2(weight) -> conv2d(x, weight, stride 2, pad 1)
    as bf16 channels-last [B, Cout, H', W / 2], computed on the 64-channel tcgen05 kernel through pixel-pair views."""
    B, Cin, H, W = x.shape
    Cout = w2.size(0) // 2
    xp = x.permute(0, 2, 3, 1).reshape(B, H, W // 2, 2 * Cin).permute(0, 3, 1, 2)          # view, channels-last
    yp = conv2d(xp, w2, None, (1, 1), 1.0, stride=2)                                          # [B, 2*Cout, H', W/4]
    Ho, Wo2 = yp.shape[2], yp.shape[3]
    return yp.permute(0, 2, 3, 1).reshape(B, Ho, 2 * Wo2, Cout).permute(0, 3, 1, 2)          # view [B, Cout, H', W/2]


def pack_heads(act, rbox, angle):
    """The three head convolutions (Conv2d(C, 1, 1), Conv2d(C, 4, 1), Conv2d(C, 2, 1), all with bias) as the [8, C] bf16
    block + [8] fp32 bias fots_b200_heads_nhwc_bf16 wants: rows {act, zeros, rbox0..3, angle0, angle1}."""
    C = act.in_channels
    w = torch.zeros((8, C), dtype=torch.bfloat16, device=act.weight.device)
    b = torch.zeros((8,), dtype=torch.float32, device=act.weight.device)
    w[0] = act.weight.detach()[0, :, 0, 0]
    w[2:6] = rbox.weight.detach()[:, :, 0, 0]
    w[6:8] = angle.weight.detach()[:, :, 0, 0]
    b[0] = act.bias.detach().float()[0]
    b[2:6] = rbox.bias.detach().float()
    b[6:8] = angle.bias.detach().float()
    return w.contiguous(), b.contiguous()


def pack_to1(conv):
    """Conv2d(C, 1, 1, bias=True) in the 8-row block of fots_b200_conv1x1_to1_nhwc_bf16 (filter in row 0)."""
    C = conv.in_channels
    w = torch.zeros((8, C), dtype=torch.bfloat16, device=conv.weight.device)
    b = torch.zeros((8,), dtype=torch.float32, device=conv.weight.device)
    w[0] = conv.weight.detach()[0, :, 0, 0]
    b[0] = conv.bias.detach().float()[0]
    return w.contiguous(), b.contiguous()


def conv1x1_to1(x, packed, sigmoid=False):
    """x bf16 channels-last [B, C, H, W] -> bf16 logits [B, 1, H, W] (one pass over x, bias added); sigmoid=True: sigmoid(logits)."""
    B, C, H, W = x.shape
    out = torch.empty((B, 1, H, W), dtype=torch.bfloat16, device=x.device)
    L = _lib()
    L.fots_b200_conv1x1_to1_nhwc_bf16.restype = ctypes.c_int
    L.fots_b200_conv1x1_to1_nhwc_bf16.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    with torch.cuda.device(x.device):
        rc = L.fots_b200_conv1x1_to1_nhwc_bf16(x.data_ptr(), packed[0].data_ptr(), packed[1].data_ptr(), out.data_ptr(), B, H, W, C,
                                               1 if sigmoid else 0, torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(rc, "fots_b200_conv1x1_to1_nhwc_bf16")
    return out


def heads(x, packed):
    """x bf16 channels-last [B, C, H, W] -> (seg [B,1,H,W], rbox [B,4,H,W], angle [B,2,H,W]) fp32 in one pass over x."""
    B, C, H, W = x.shape
    seg = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    rb = torch.empty((B, 4, H, W), dtype=torch.float32, device=x.device)
    an = torch.empty((B, 2, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib().fots_b200_heads_nhwc_bf16(x.data_ptr(), packed[0].data_ptr(), packed[1].data_ptr(), seg.data_ptr(),
                                              rb.data_ptr(), an.data_ptr(), B, H, W, C,
                                              torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(rc, "fots_b200_heads_nhwc_bf16")
    return seg, rb, an


def conv3x3_c3_pool(x, weight, bias, pool2x2):
    """Consumer B's first layer in one kernel (fots_b200_conv3x3_c3_pool_nhwc_bf16): x fp32 NCHW [N, 3, H, W] ->
    [maxpool2x2](relu(conv3x3_pad1(x, weight) + bias)) as bf16 channels-last.  weight bf16 [Cout, 3, 3, 3], bias fp32 [Cout]."""
    N, _, H, W = x.shape
    Cout = weight.size(0)
    P = 2 if pool2x2 else 1
    x = x.float().contiguous()
    wt = weight.contiguous()                                  # plain NCHW-contiguous [Cout, 3, 3, 3]
    y = torch.empty((N, Cout, H // P, W // P), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    L = _lib()
    L.fots_b200_conv3x3_c3_pool_nhwc_bf16.restype = ctypes.c_int
    L.fots_b200_conv3x3_c3_pool_nhwc_bf16.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    with torch.cuda.device(x.device):
        rc = L.fots_b200_conv3x3_c3_pool_nhwc_bf16(x.data_ptr(), wt.data_ptr(), bias.data_ptr() if bias is not None else None, y.data_ptr(),
                                                   N, H, W, Cout, 1 if pool2x2 else 0, torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(rc, "fots_b200_conv3x3_c3_pool_nhwc_bf16")
    return y


def maxpool(x, kernel, stride, padding):
    """MaxPool2d(kernel, stride, padding) (floor mode) of a bf16 channels-last tensor on fots_b200_maxpool_nhwc_bf16."""
    N, C, H, W = x.shape
    (kh, kw), (sh, sw), (ph, pw) = kernel, stride, padding
    Ho, Wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    y = torch.empty((N, C, Ho, Wo), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    L = _lib()
    L.fots_b200_maxpool_nhwc_bf16.restype = ctypes.c_int
    L.fots_b200_maxpool_nhwc_bf16.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] * 10 + [ctypes.c_void_p]
    with torch.cuda.device(x.device):
        rc = L.fots_b200_maxpool_nhwc_bf16(x.data_ptr(), y.data_ptr(), N, H, W, C, kh, kw, sh, sw, ph, pw,
                                           torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(rc, "fots_b200_maxpool_nhwc_bf16")
    return y


def pack_merged_heads(act, rbox, angle, pw_conv, lateral_conv):
    """Fold the last level of the top-down merge into the heads (fots_b200_heads_merged_nhwc_bf16): with the head block
    Wh [8, 256] (pack_heads layout), x = pw_conv(d) + lateral_conv(s) * gate gives logits = (Wh Wpw) d + gate * (Wh Wlat) s + bh.
    Returns (w1 bf16 [8, 256], w2 bf16 [8, Cs], bias fp32 [8]) or None when a convolution carries a bias (it would need the gate)."""
    if pw_conv.bias is not None or lateral_conv.bias is not None:
        return None
    C = act.in_channels
    wh = torch.zeros((8, C), dtype=torch.float32, device=act.weight.device)
    bh = torch.zeros((8,), dtype=torch.float32, device=act.weight.device)
    wh[0] = act.weight.detach().float()[0, :, 0, 0]
    wh[2:6] = rbox.weight.detach().float()[:, :, 0, 0]
    wh[6:8] = angle.weight.detach().float()[:, :, 0, 0]
    bh[0] = act.bias.detach().float()[0]
    bh[2:6] = rbox.bias.detach().float()
    bh[6:8] = angle.bias.detach().float()
    w1 = wh @ pw_conv.weight.detach().float()[:, :, 0, 0]                     # [8, 256]
    w2 = wh @ lateral_conv.weight.detach().float()[:, :, 0, 0]                # [8, Cs]
    return w1.to(torch.bfloat16).contiguous(), w2.to(torch.bfloat16).contiguous(), bh.contiguous()


def heads_merged(d, s, gate_prob, packed):
    """(seg, rbox, angle) fp32 from d bf16 [B, 256, H, W], s bf16 [B, 64, H, W] (both channels-last) and the low-resolution
    gate probabilities bf16 [B, 1, gh, gw] in one pass over d and s (see pack_merged_heads)."""
    B, C1, H, W = d.shape
    C2 = s.size(1)
    seg = torch.empty((B, 1, H, W), dtype=torch.float32, device=d.device)
    rbox = torch.empty((B, 4, H, W), dtype=torch.float32, device=d.device)
    ang = torch.empty((B, 2, H, W), dtype=torch.float32, device=d.device)
    L = _lib()
    L.fots_b200_heads_merged_nhwc_bf16.restype = ctypes.c_int
    L.fots_b200_heads_merged_nhwc_bf16.argtypes = [ctypes.c_void_p] * 9 + [ctypes.c_int] * 7 + [ctypes.c_void_p]
    with torch.cuda.device(d.device):
        rc = L.fots_b200_heads_merged_nhwc_bf16(d.data_ptr(), packed[0].data_ptr(), s.data_ptr(), packed[1].data_ptr(), gate_prob.data_ptr(),
                                                packed[2].data_ptr(), seg.data_ptr(), rbox.data_ptr(), ang.data_ptr(), B, H, W, C1, C2,
                                                gate_prob.size(2), gate_prob.size(3), torch.cuda.current_stream(d.device).cuda_stream)
    _cabi.check(rc, "fots_b200_heads_merged_nhwc_bf16")
    return seg, rbox, ang


def pack_gather_heads(act, rbox, angle, pw_conv, dw_conv):
    """Fold the depthwise half of upconv2 into the heads as well (fots_b200_heads_gather_nhwc_bf16): with M = Wh Wpw [8, 256]
    and the depthwise taps w_dw [256, 9], filter tap * 8 + o of the returned 1x1 convolution weight (bf16 [128, 256, 1, 1],
    filters 72.. are zero: the tcgen05 kernel's output tiles are 64 wide) is M[o, :] * w_dw[:, tap] -- it turns the
    low-resolution map f2 into the tap map T.  None when a bias is in the way."""
    if pw_conv.bias is not None or dw_conv.bias is not None or tuple(dw_conv.weight.shape[1:]) != (1, 3, 3):
        return None
    C = act.in_channels
    wh = torch.zeros((8, C), dtype=torch.float32, device=act.weight.device)
    wh[0] = act.weight.detach().float()[0, :, 0, 0]
    wh[2:6] = rbox.weight.detach().float()[:, :, 0, 0]
    wh[6:8] = angle.weight.detach().float()[:, :, 0, 0]
    m = wh @ pw_conv.weight.detach().float()[:, :, 0, 0]                      # [8, 256]
    wdw = dw_conv.weight.detach().float().reshape(-1, 9)                      # [256, 9]
    a = torch.zeros((128, m.size(1)), dtype=torch.float32, device=m.device)
    a[:72] = (m[None, :, :] * wdw.t()[:, None, :]).reshape(72, -1)            # [9, 8, 256]
    return a.to(torch.bfloat16).reshape(128, -1, 1, 1).contiguous(memory_format=torch.channels_last)


def heads_gather(f2, s, gate_prob, a72, packed):
    """(seg, rbox, angle) fp32 at s's resolution from the LOW-resolution map f2 bf16 [B, 256, h, w], s bf16 [B, 64, H, W]
    (channels-last) and the gate probabilities bf16 [B, 1, h, w]: T = conv1x1(f2, a72) on the tcgen05 kernel (bf16
    [B, h, w, 128]), then the 9-tap x 4-sample gather per pixel.  packed = pack_merged_heads(...) (its w2 and bias),
    a72 = pack_gather_heads(...)."""
    B, C1, h, w = f2.shape
    _, C2, H, W = s.shape
    T = conv2d(f2, a72)
    seg = torch.empty((B, 1, H, W), dtype=torch.float32, device=f2.device)
    rbox = torch.empty((B, 4, H, W), dtype=torch.float32, device=f2.device)
    ang = torch.empty((B, 2, H, W), dtype=torch.float32, device=f2.device)
    L = _lib()
    L.fots_b200_heads_gather_nhwc_bf16.restype = ctypes.c_int
    L.fots_b200_heads_gather_nhwc_bf16.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 7 + [ctypes.c_int] * 6 + [ctypes.c_void_p]
    with torch.cuda.device(f2.device):
        rc = L.fots_b200_heads_gather_nhwc_bf16(T.data_ptr(), T.size(1), s.data_ptr(), packed[1].data_ptr(), gate_prob.data_ptr(),
                                                packed[2].data_ptr(), seg.data_ptr(), rbox.data_ptr(), ang.data_ptr(), B, H, W, h, w, C2,
                                                torch.cuda.current_stream(f2.device).cuda_stream)
    _cabi.check(rc, "fots_b200_heads_gather_nhwc_bf16")
    return seg, rbox, ang


def stem_eligible(x, conv):
    """The first layer: fp32 (preprocessed) or uint8 (raw, normalised on load) channels-last image,
    Conv2d(3, 16, 3, 1, 1, bias=False) with bf16 weights."""
    w = conv.weight
    return (ENABLED and x.is_cuda and x.dtype in (torch.float32, torch.uint8) and x.dim() == 4 and x.size(1) == 3
            and x.is_contiguous(memory_format=torch.channels_last) and w.dtype == torch.bfloat16
            and tuple(w.shape) == (16, 3, 3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None
            and not (torch.is_grad_enabled() and (x.requires_grad or w.requires_grad)))


def stem_conv_stats(x, weight):
    """conv2d(bf16(x), weight, pad 1) -> (y bf16 channels-last [B, 16, H, W], ws fp64 [B, 16, 2] statistics of y)
    in one pass (fots_b200_stem_conv3x3_c3_c16)."""
    from . import fused
    B, _, H, W = x.shape
    wk = weight.contiguous(memory_format=torch.channels_last)
    y = torch.empty((B, 16, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    ws = fused.workspace(x.device, B * 32)
    fn = _lib().fots_b200_stem_conv3x3_c3_c16_u8 if x.dtype == torch.uint8 else _lib().fots_b200_stem_conv3x3_c3_c16
    with torch.cuda.device(x.device):
        st = fn(x.data_ptr(), wk.data_ptr(), y.data_ptr(), ws.data_ptr(), B, H, W,
                torch.cuda.current_stream(x.device).cuda_stream)
    _cabi.check(st, "fots_b200_stem_conv3x3_c3_c16")
    return y, ws
