"""GPU parity of the two kernels that join RoIRotate to the recogniser on the bf16 inference path:

* rroi_b200_forward_bf16 against the CPU oracle: bit-exact after one bf16 rounding;
* fots_b200_conv2d_nhwc_bf16 (tcgen05 implicit GEMM, csrc/conv_tc.cu) against torch's fp32 convolution of the same
  bf16-rounded operands: tolerance = one bf16 rounding of the result (2^-8 relative) plus fp32 sum-order noise.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers as Hh
import workloads as WL

pytestmark = pytest.mark.gpu


def _bf16_bits(t):
    return t.contiguous().view(torch.int16).cpu().numpy()


@pytest.mark.parametrize("C,B,n_per,ph,pw,scale", [(64, 1, 64, 8, 64, 0.25), (256, 1, 64, 8, 64, 0.25), (32, 2, 9, 8, 48, 0.25),
                                                    (128, 3, 17, 4, 33, 0.5), (64, 4, 64, 8, 64, 0.25)])
def test_bf16_forward_is_the_rounded_fp32_forward(oracle, cuda, C, B, n_per, ph, pw, scale):
    """pooled_bf16 == bf16(oracle(float(features_bf16))) bit for bit, including the zero tail (pw > roi width) and
    RoIs overhanging the border; small grids take the 64-bin-tile kernel, the 4-image case the 256-bin one."""
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.rroi_align.functions.rroi_align import rroi_align_bf16
    H, W = 180, 320
    feats = torch.from_numpy(WL.features(C + B, B, C, H, W)).to(cuda).to(torch.bfloat16)
    rois = np.concatenate([WL.random_rois(7 + i, n_per, i) for i in range(B)], 0)
    rois[0, 1:3] = (3.0, 5.0)                                   # hangs over the top-left corner
    x = feats.contiguous(memory_format=torch.channels_last)
    want, _, _ = oracle.forward(feats.float().cpu().numpy(), rois, ph, pw, scale, threads=0)
    rois[-1, 0] = B + 3                                          # batch index out of range -> zeros (the oracle,
    want[-1] = 0.0                                               # like the reference, would read out of bounds)
    want_bits = _bf16_bits(torch.from_numpy(want).to(torch.bfloat16))
    for variant in (0, 1, 5, 7):          # 7: whole-RoI 512-bin tiles
        got = rroi_align_bf16(x, torch.from_numpy(rois).to(cuda), ph, pw, scale,
                              opts=_cabi.opts(variant=variant, rois_ready=variant == 5))
        assert got.dtype == torch.bfloat16 and got.shape == (rois.shape[0], C, ph, pw)
        assert got.is_contiguous(memory_format=torch.channels_last)
        got_bits = _bf16_bits(got.contiguous())                  # logical NCHW order, like the oracle
        neq = got_bits != want_bits
        assert not neq.any(), "variant %d: %d / %d bf16 values differ" % (variant, int(neq.sum()), neq.size)


def test_bf16_forward_rejects_what_it_cannot_do(cuda):
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.rroi_align.functions.rroi_align import rroi_align_bf16
    rois = torch.tensor([[0, 10, 10, 8, 16, 0]], dtype=torch.float32, device=cuda)
    with pytest.raises(TypeError):
        rroi_align_bf16(torch.zeros(1, 64, 8, 8, device=cuda), rois, 8, 16, 1.0)           # fp32 features
    with pytest.raises(_cabi.RRoiAlignError):
        rroi_align_bf16(torch.zeros(1, 48, 8, 8, device=cuda, dtype=torch.bfloat16), rois, 8, 16, 1.0)   # C = 48
    out = rroi_align_bf16(torch.zeros(1, 64, 8, 8, device=cuda, dtype=torch.bfloat16), rois[:0], 8, 16, 1.0)
    assert out.shape == (0, 64, 8, 16)


CONV_CASES = [
    # N, H, W, Cin, Cout, R, S, pad_h, pad_w, bias, slope
    (1, 2, 64, 64, 64, 1, 1, 0, 0, False, 1.0),          # plain GEMM, one tile
    (2, 4, 64, 128, 128, 1, 1, 0, 0, True, 1.0),
    (2, 4, 64, 64, 64, 3, 3, 1, 1, False, 1.0),          # zero padding comes from the TMA out-of-bounds fill
    (3, 8, 64, 64, 128, 3, 3, 1, 1, False, 1.0),         # conv5
    (3, 8, 64, 128, 128, 3, 3, 1, 1, True, 0.01),        # conv6 + leaky
    (3, 4, 64, 128, 256, 3, 3, 1, 1, False, 1.0),        # conv7
    (3, 4, 64, 256, 256, 3, 3, 1, 1, False, 0.01),       # conv8 / conv9 + leaky
    (5, 2, 64, 256, 256, 2, 3, 0, 1, True, 1.0),         # conv10_s: 2x3, pad (0,1); tiles span two RoIs
    (1, 45, 80, 64, 64, 3, 3, 1, 1, False, 0.0),         # ragged tiles in h and w, ReLU
    (2, 23, 37, 128, 192, 3, 3, 1, 1, True, 0.0),        # odd sizes, Cout = 3 x 64
    (1, 90, 160, 128, 128, 3, 3, 1, 1, False, 1.0),      # layer2 block at 720p
    (2, 5, 9, 64, 64, 5, 5, 2, 2, True, 1.0),            # 25 taps, image smaller than a tile
    (80, 4, 64, 128, 256, 3, 3, 1, 1, False, 0.01),      # 160 pixel tiles: every CTA (pair) walks several tiles
    (2, 23, 37, 64, 512, 3, 3, 1, 1, True, 0.0),         # two 256-wide cout tiles, ragged pixel tiles
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("bn", [0, 64, 128, 256, 512])
def test_tcgen05_conv_matches_torch_fp32(cuda, case, bn):
    """bn: forced output-channel tile; 512 = a 256-wide tile on a CTA pair (tcgen05 cta_group::2, cluster of two)."""
    from fots.pytorch_b200.pipeline import conv as TC
    N, H, W, Cin, Cout, R, S, ph, pw, bias, slope = case
    if bn and Cout % min(bn, 256):
        pytest.skip("Cout is not a multiple of the forced tile")
    g = torch.Generator().manual_seed(N * 131 + Cin + Cout + R)
    x = torch.randn(N, Cin, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, R, S, generator=g) / (Cin * R * S) ** 0.5).to(cuda).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g).to(cuda) if bias else None
    TC.set_tile(bn)
    try:
        y = TC.conv2d(x, w, b, (ph, pw), slope)
    finally:
        TC.set_tile(0)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float(), w.float(), b, 1, (ph, pw))
    if slope != 1.0:
        ref = F.leaky_relu(ref, slope)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    err = (y.float() - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3 * float(ref.abs().max())      # one bf16 rounding + fp32 sum-order noise
    assert bool((err <= tol).all()), "max excess %.4g" % float((err - tol).max())


STRIDED_CASES = [
    # N, H, W, Cin, Cout, R, S, pad, bias, slope
    (2, 36, 64, 64, 64, 3, 3, 1, False, 0.0),            # layer0_1's second convolution (stride 2, ReLU), even sizes
    (1, 45, 81, 64, 128, 3, 3, 1, True, 1.0),            # first block of a stage: odd sizes, ragged tiles
    (3, 23, 37, 128, 256, 1, 1, 0, True, 1.0),           # 1x1 stride-2 down-sampling branch (BatchNorm folded into bias)
    (2, 7, 9, 64, 64, 3, 3, 1, False, 1.0),              # output smaller than one tile
    (1, 360, 640, 64, 64, 3, 3, 1, False, 0.0),          # full 720p size of layer0_1[1]
]


@pytest.mark.parametrize("case", STRIDED_CASES)
@pytest.mark.parametrize("bn", [0, 64, 256])
def test_tcgen05_conv_stride2_matches_torch_fp32(cuda, case, bn):
    """Stride 2 through the TMA tensor map's traversal stride (fots_b200_conv2d_strided_nhwc_bf16): every second input
    pixel is gathered by the copy engine, the rest is the stride-1 kernel."""
    from fots.pytorch_b200.pipeline import conv as TC
    N, H, W, Cin, Cout, R, S, pad, bias, slope = case
    if bn and Cout % bn:
        pytest.skip("Cout is not a multiple of the forced tile")
    g = torch.Generator().manual_seed(N * 17 + H + Cout)
    x = torch.randn(N, Cin, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, R, S, generator=g) / (Cin * R * S) ** 0.5).to(cuda).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g).to(cuda) if bias else None
    TC.set_tile(bn)
    try:
        y = TC.conv2d(x, w, b, (pad, pad), slope, stride=2)
    finally:
        TC.set_tile(0)
    ref = F.conv2d(x.float(), w.float(), b, 2, (pad, pad))
    if slope != 1.0:
        ref = F.leaky_relu(ref, slope)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    err = (y.float() - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3 * float(ref.abs().max())
    assert bool((err <= tol).all()), "max excess %.4g" % float((err - tol).max())


@pytest.mark.parametrize("N,H,W,Cin,Cout,R,S,ph,pw", [(3, 8, 64, 64, 128, 3, 3, 1, 1), (5, 2, 64, 256, 256, 2, 3, 0, 1),
                                                       (2, 45, 80, 64, 64, 3, 3, 1, 1), (2, 23, 37, 128, 192, 3, 3, 1, 1),
                                                       (7, 3, 5, 64, 64, 3, 3, 1, 1)])
def test_conv_epilogue_statistics_feed_instancenorm(cuda, N, H, W, Cin, Cout, R, S, ph, pw):
    """fots_b200_conv2d_stats_nhwc_bf16: the [N, Cout, 2] sums equal the sums of the bf16 tensor it stored (ragged
    tiles and tiles spanning several images included), and conv -> fused InstanceNorm through them equals the
    two-pass kernel on the same convolution output."""
    from fots.pytorch_b200.pipeline import conv as TC, fused
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn(N, Cin, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, R, S, generator=g) / (Cin * R * S) ** 0.5).to(cuda).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g).to(cuda)
    y, ws = TC.conv2d(x, w, b, (ph, pw), 1.0, stats=True)
    y_plain = TC.conv2d(x, w, b, (ph, pw), 1.0)
    assert torch.equal(y, y_plain)
    got = ws[:N * Cout * 2].view(N, Cout, 2).clone()
    yd = y.double()
    want = torch.stack([yd.sum((2, 3)), (yd * yd).sum((2, 3))], 2)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-3), float((got - want).abs().max())
    gamma, beta = torch.randn(Cout, device=cuda), torch.randn(Cout, device=cuda)
    two_pass = fused.instnorm_act(y, gamma, beta, 1e-5, 0.01)
    y2, ws2 = TC.conv2d(x, w, b, (ph, pw), 1.0, stats=True)
    one_pass = fused.instnorm_act(y2, gamma, beta, 1e-5, 0.01, stats=ws2)
    assert float((one_pass.float() - two_pass.float()).abs().max()) <= 2.0 ** -6 * float(two_pass.float().abs().max())


@pytest.mark.parametrize("B,H,W", [(2, 64, 256), (1, 37, 150), (3, 8, 128), (1, 720, 1280), (9, 17, 33)])
def test_stem_convolution_and_statistics(cuda, B, H, W):
    """fots_b200_stem_conv3x3_c3_c16 vs torch: the 3 -> 16 convolution on the bf16-rounded image (one bf16 rounding of
    the result) and the [B, 16, 2] sums of exactly the tensor it stored; ragged tiles, several images per CTA range."""
    from fots.pytorch_b200.pipeline import conv as TC
    g = torch.Generator().manual_seed(B * 7 + H)
    x = torch.randn(B, 3, H, W, generator=g).to(cuda).contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(3, 16, 3, 1, 1, bias=False).to(cuda).to(torch.bfloat16).to(memory_format=torch.channels_last)
    with torch.no_grad():
        assert TC.stem_eligible(x, conv)
        y, ws = TC.stem_conv_stats(x, conv.weight)
        ref = F.conv2d(x.to(torch.bfloat16).float(), conv.weight.float(), None, 1, 1)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    err = (y.float() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 1e-3 * float(ref.abs().max())).all()), float(err.max())
    yd = y.double()
    want = torch.stack([yd.sum((2, 3)), (yd * yd).sum((2, 3))], 2)
    got = ws[:B * 32].view(B, 16, 2)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-3), float((got - want).abs().max())


@pytest.mark.parametrize("N,C,H,W", [(5, 128, 8, 64), (3, 256, 4, 64), (2, 64, 7, 33), (1, 8, 2, 1)])
def test_maxpool_h2_matches_torch(cuda, N, C, H, W):
    from fots.pytorch_b200.pipeline import fused
    x = torch.randn(N, C, H, W, device=cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x[0, 0, 0, 0] = float("nan")
    got = fused.maxpool_h2(x)
    want = F.max_pool2d(x.float(), (2, 1), (2, 1)).to(torch.bfloat16)
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(torch.nan_to_num(got.float(), nan=-7.0), torch.nan_to_num(want.float(), nan=-7.0))


def test_tcgen05_conv_argument_checks(cuda):
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.pipeline import conv as TC
    x = torch.zeros(1, 48, 8, 8, device=cuda, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.zeros(64, 48, 3, 3, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(_cabi.RRoiAlignError):
        TC.conv2d(x, w, None, (1, 1))                                # Cin % 64 != 0
    conv = torch.nn.Conv2d(48, 64, 3, 1, 1).to(cuda).to(torch.bfloat16)
    assert not TC.eligible(x, conv)
    conv = torch.nn.Conv2d(64, 64, 3, 2, 1).to(cuda).to(torch.bfloat16)
    x = torch.zeros(1, 64, 8, 8, device=cuda, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert not TC.eligible(x, conv)                                  # stride 2 stays on the library path
    conv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(cuda).to(torch.bfloat16)
    with torch.no_grad():
        assert TC.eligible(x, conv)
        assert TC.apply(conv, x, 0.01).shape == (1, 64, 8, 8)


def test_recogniser_on_tensor_cores_matches_library_path(cuda):
    """forward_ocr (tools/models.py:334-379) on bf16 channels-last input: the tcgen05 convolutions (+ fused leaky-ReLU)
    against the same network on torch's library convolutions; both round every layer to bf16 once."""
    from fots.pytorch_b200.pipeline import FOTSNet
    from fots.pytorch_b200.pipeline import conv as TC
    torch.manual_seed(0)
    net = FOTSNet(attention=True, nclass=89).to_b200(cuda, inference=True)
    pooled = torch.randn(6, 64, 8, 64, device=cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert TC.eligible(pooled, net.conv5)
        a = net.forward_ocr(pooled)
        TC.ENABLED = False
        try:
            b = net.forward_ocr(pooled)
        finally:
            TC.ENABLED = True
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        img = torch.randn(2, 3, 64, 128, device=cuda).contiguous(memory_format=torch.channels_last)
        fa = net(img)                                       # backbone: tcgen05 convs + epilogue statistics in layer1/2
        TC.ENABLED = False
        try:
            fb = net(img)
        finally:
            TC.ENABLED = True
    for ta, tb in zip(fa[0] + fa[3], fb[0] + fb[3]):
        assert ta.shape == tb.shape
        # ~60 layers each rounded to bf16 once, by different kernels on the two sides (every convolution, depthwise
        # convolution, head and folded BatchNorm now runs hand-written): the paths drift apart by bf16 noise only
        assert float((ta.float() - tb.float()).abs().mean()) < 0.05 * float(tb.float().abs().mean()) + 1e-3
    assert a.shape == b.shape == (6, 89, 64)
    assert float((a - b).abs().max()) < 0.15 and float((a - b).abs().mean()) < 0.02      # log-probabilities
    assert float((a.argmax(1) == b.argmax(1)).float().mean()) > 0.9


@pytest.mark.parametrize("B,C,H,W,residual", [
    (8, 128, 90, 160, True),      # stage 2 at 720p: 4 channel slices x 8-CTA clusters
    (2, 256, 45, 80, True),       # stage 3
    (2, 512, 23, 40, False),      # stage 4: 64 channel groups per pixel
    (40, 128, 8, 64, False),      # recogniser batch5: one CTA per RoI
    (40, 256, 1, 64, False),      # batch10_s
    (3, 64, 37, 53, True),        # ragged rows per CTA
    (1, 32, 5, 7, False),         # tiny
    (2, 64, 180, 320, False),     # does not fit: two-pass form
])
def test_single_pass_cluster_instancenorm_equals_two_pass_and_torch(cuda, B, C, H, W, residual):
    """in_fused_cluster_kernel (one launch, instance kept in a cluster's shared memory, partial sums exchanged through
    distributed shared memory) against the two-pass kernels and against torch's InstanceNorm in fp32."""
    from fots.pytorch_b200.pipeline import fused
    g = torch.Generator().manual_seed(B * 100 + C + H)
    x = (torch.randn(B, C, H, W, generator=g) * 1.7 + 0.3).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = torch.randn(B, C, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if residual else None
    gamma, beta = torch.randn(C, generator=g).to(cuda), torch.randn(C, generator=g).to(cuda)
    fused.set_single_pass(2)                       # whenever the instance fits (automatic: only small instances)
    try:
        one = fused.instnorm_act(x, gamma, beta, 1e-5, 0.01, r)
        again = fused.instnorm_act(x, gamma, beta, 1e-5, 0.01, r)
        fused.set_single_pass(0)
        two = fused.instnorm_act(x, gamma, beta, 1e-5, 0.01, r)
    finally:
        fused.set_single_pass(1)
    ref = F.instance_norm(x.float(), weight=gamma, bias=beta, eps=1e-5)
    if residual:
        ref = ref + r.float()
    ref = F.leaky_relu(ref, 0.01)
    scale = float(ref.abs().max())
    assert float((one.float() - ref).abs().max()) <= 2.0 ** -7 * scale + 1e-3
    assert float((one.float() - two.float()).abs().max()) <= 2.0 ** -7 * scale      # same statistics up to summation order
    assert torch.equal(one, again)                                                  # deterministic: no atomics


@pytest.mark.parametrize("N,C,H,W,stride", [(2, 256, 45, 80, 1), (1, 128, 90, 160, 2), (3, 512, 23, 40, 1), (2, 256, 45, 80, 2),
                                             (1, 64, 7, 5, 1), (1, 64, 7, 5, 2), (2, 256, 180, 320, 1), (1, 192, 33, 47, 2)])
def test_depthwise_3x3_matches_torch_fp32(cuda, N, C, H, W, stride):
    """fots_b200_dwconv3x3_nhwc_bf16 (shared-memory tile + halo, 4-output strips) vs torch's grouped convolution in fp32
    on the same bf16 operands: one bf16 rounding of an fp32 sum of nine products."""
    from fots.pytorch_b200.pipeline import conv as TC
    g = torch.Generator().manual_seed(N + C + H + stride)
    conv = torch.nn.Conv2d(C, C, 3, stride, 1, groups=C, bias=False)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(C, 1, 3, 3, generator=g) / 3.0)
    conv = conv.to(cuda).to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(N, C, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        assert TC.dw_eligible(x, conv)
        y = TC.dwconv(conv, x)
        ref = F.conv2d(x.float(), conv.weight.float(), None, stride, 1, groups=C)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    err = (y.float() - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-5
    assert bool((err <= tol).all()), "max excess %.4g" % float((err - tol).max())


@pytest.mark.parametrize("B,C,H,W", [(2, 256, 45, 80), (1, 256, 180, 320), (3, 128, 7, 9), (1, 512, 5, 3)])
def test_fused_detection_heads_match_torch_fp32(cuda, B, C, H, W):
    """fots_b200_heads_nhwc_bf16: act / rbox / angle 1x1 convolutions + sigmoid, x128, (sin, cos) normalisation in one pass
    over the feature map, against the three nn.Conv2d + torch ops of FOTSNet._heads evaluated in fp32 on the same bf16
    operands."""
    from fots.pytorch_b200.pipeline import conv as TC
    g = torch.Generator().manual_seed(B + C + H)
    mk = lambda co: torch.nn.Conv2d(C, co, 1, bias=True)
    act, rbox, angle = mk(1), mk(4), mk(2)
    with torch.no_grad():
        for m in (act, rbox, angle):
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) / C ** 0.5)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g))
    act, rbox, angle = (m.to(cuda).to(torch.bfloat16) for m in (act, rbox, angle))
    x = torch.randn(B, C, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        seg, rb, an = TC.heads(x, TC.pack_heads(act, rbox, angle))
        f = lambda m: F.conv2d(x.float(), m.weight.float(), m.bias.float())
        want_seg = torch.sigmoid(f(act))
        want_rb = torch.sigmoid(f(rbox)) * 128
        a = torch.sigmoid(f(angle)) * 2 - 1
        want_an = a / torch.sqrt((a * a).sum(1, keepdim=True))
    assert seg.shape == want_seg.shape and rb.shape == want_rb.shape and an.shape == want_an.shape
    assert float((seg - want_seg).abs().max()) <= 1e-4
    assert float((rb - want_rb).abs().max()) <= 1e-2            # values up to 128
    assert float((an - want_an).abs().max()) <= 2e-3            # the normalisation amplifies near |a| ~ 0
    # the one-channel variant (attention gate logits, bf16 out)
    with torch.no_grad():
        got = TC.conv1x1_to1(x, TC.pack_to1(act))
    assert got.shape == (B, 1, H, W) and got.dtype == torch.bfloat16
    assert float((got.float() - f(act)).abs().max()) <= 2.0 ** -8 * float(f(act).abs().max()) + 1e-3
    with torch.no_grad():
        gs = TC.conv1x1_to1(x, TC.pack_to1(act), sigmoid=True)
    assert float((gs.float() - want_seg).abs().max()) <= 2.0 ** -8


def test_folded_batchnorm_downsample_branch_matches_module(cuda):
    """FOTSNet.to_b200(inference=True) folds the eval-mode BatchNorm of every stage's 1x1 stride-2 down-sampling branch
    (tools/models.py:319-324) into the convolution (bf16 weights, fp32 bias) and runs it on the strided tcgen05 kernel:
    equal to conv -> BatchNorm of the module with non-trivial running statistics."""
    from fots.pytorch_b200.pipeline import FOTSNet
    from fots.pytorch_b200.pipeline import conv as TC
    from fots.pytorch_b200.pipeline.nets import _downsample
    torch.manual_seed(3)
    net = FOTSNet(attention=True, nclass=89)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.4, 1.6); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.3)
    net.to_b200(cuda, inference=True)
    checked = 0
    with torch.no_grad():
        for stage, cin, hw in ((net.layer2, 64, (45, 80)), (net.layer3, 128, (23, 41)), (net.layer4, 256, (12, 20))):
            block = stage[0]
            assert block._ds_pack is not None
            x = torch.randn(2, cin, *hw, device=cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            got = _downsample(block, x)
            want = block.downsample[1](block.downsample[0](x).float())            # library conv, BatchNorm in fp32
            assert got.shape == want.shape
            assert float((got.float() - want).abs().max()) <= 2.0 ** -6 * float(want.abs().max()) + 1e-3
            checked += 1
    assert checked == 3 and TC.LEVEL >= 2


@pytest.mark.parametrize("N,H,W,Cin,Cout,bias,slope", [
    (1, 45, 80, 64, 64, False, 0.0),          # layer1 at reduced size: ragged in h (45 = 32 + 13)
    (2, 90, 160, 128, 128, True, 1.0),        # layer2 block
    (2, 23, 37, 128, 192, True, 0.0),         # odd sizes: ragged in w (37 = 4 x 8 + 5), three 64-wide cout tiles
    (1, 180, 320, 64, 64, False, 1.0),        # layer1 at 720p
    (3, 8, 64, 64, 128, False, 0.01),         # the recogniser's conv5 shape (forced: tiles mostly outside the image)
    (1, 33, 9, 256, 128, True, 1.0),          # four 64-channel chunks, two tile columns
    (2, 37, 21, 64, 64, True, 0.01),          # 64 -> 64 ragged in both directions, bias + leaky
])
def test_tcgen05_conv_halo_reuse_matches_torch_fp32(cuda, N, H, W, Cin, Cout, bias, slope):
    """Halo reuse (one TMA load of the tile + halo rows feeds the three taps of a filter column; the A descriptor of tap
    (r, s) is the same buffer at +r image rows) against torch's fp32 convolution and against the per-tap kernel."""
    from fots.pytorch_b200.pipeline import conv as TC
    g = torch.Generator().manual_seed(N * 31 + H + Cout)
    x = torch.randn(N, Cin, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(cuda).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g).to(cuda) if bias else None
    TC.set_halo(1)                 # 64 -> 64: one 10-pixel-wide box + resident weights; otherwise three x-shifted copies
    try:
        y = TC.conv2d(x, w, b, (1, 1), slope)
        TC.set_halo(2)             # always the three-copy form
        y2 = TC.conv2d(x, w, b, (1, 1), slope)
        TC.set_halo(0)
        y0 = TC.conv2d(x, w, b, (1, 1), slope)
    finally:
        TC.set_halo(-1)
    ref = F.conv2d(x.float(), w.float(), b, 1, (1, 1))
    if slope != 1.0:
        ref = F.leaky_relu(ref, slope)
    err = (y.float() - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3 * float(ref.abs().max())
    assert bool((err <= tol).all()), "max excess %.4g" % float((err - tol).max())
    assert float((y.float() - y0.float()).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())     # fp32 sum order differs
    assert float((y2.float() - y0.float()).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 45, 84), (1, 720, 1280)])
def test_stride2_conv_with_32_channels_through_pixel_pairs(cuda, B, H, W):
    """layer0's second convolution (32 -> 32, 3x3, stride 2, tools/models.py:252) has too few channels for a 64-wide
    k-block: two adjacent pixels are viewed as one 64-channel pixel on both sides (conv.pack_pixel_pairs_s2) and the
    result is again a 3x3 stride-2 convolution on the tcgen05 kernel.  Against torch's fp32 convolution."""
    from fots.pytorch_b200.pipeline import conv as TC
    g = torch.Generator().manual_seed(B + H)
    w = (torch.randn(32, 32, 3, 3, generator=g) / 17.0).to(cuda).to(torch.bfloat16)
    x = torch.randn(B, 32, H, W, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y = TC.conv3x3_s2_pixel_pairs(x, TC.pack_pixel_pairs_s2(w))
        ref = F.conv2d(x.float(), w.float(), None, 2, 1)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    err = (y.float() - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3 * float(ref.abs().max())
    assert bool((err <= tol).all()), "max excess %.4g" % float((err - tol).max())


@pytest.mark.parametrize("N,C,H,W,stride,affine", [(2, 256, 45, 80, 1, False), (1, 128, 33, 47, 2, True), (3, 512, 23, 40, 1, False), (1, 64, 7, 5, 1, True)])
def test_depthwise_with_instancenorm_on_load_equals_two_kernels(cuda, N, C, H, W, stride, affine):
    """fots_b200_dwconv3x3_norm_nhwc_bf16 = depthwise(leaky(InstanceNorm(x))) with the normalisation applied to the staged
    tile: bit-identical to the InstanceNorm kernels followed by the plain depthwise kernel (same coefficients, same bf16
    rounding of the normalised values, zero padding untouched), and within bf16 noise of torch in fp32."""
    from fots.pytorch_b200.pipeline import conv as TC, fused
    g = torch.Generator().manual_seed(N + C + H)
    conv = torch.nn.Conv2d(C, C, 3, stride, 1, groups=C, bias=False)
    norm = torch.nn.InstanceNorm2d(C, eps=1e-5, affine=affine)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(C, 1, 3, 3, generator=g) / 3.0)
        if affine:
            norm.weight.copy_(torch.rand(C, generator=g) + 0.5); norm.bias.copy_(torch.randn(C, generator=g) * 0.3)
    conv = conv.to(cuda).to(torch.bfloat16).to(memory_format=torch.channels_last)
    norm = norm.to(cuda)
    x = (torch.randn(N, C, H, W, generator=g) * 2 + 0.5).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        got = TC.dwconv_norm(conv, x, fused.instnorm_stats(x), norm, 0.01)
        fused.set_single_pass(0)
        try:
            two = TC.dwconv(conv, fused.instnorm_act(x, norm.weight, norm.bias, norm.eps, 0.01))
        finally:
            fused.set_single_pass(1)
        ref = F.conv2d(F.leaky_relu(norm(x.float()), 0.01), conv.weight.float(), None, stride, 1, groups=C)
    assert torch.equal(got, two)
    assert float((got.float() - ref).abs().max()) <= 2.0 ** -6 * float(ref.abs().max()) + 1e-3
    # epilogue statistics: the sums of exactly the tensor it stored, with and without the normalisation on load
    with torch.no_grad():
        y2, ws = TC.dwconv_norm(conv, x, fused.instnorm_stats(x), norm, 0.01, stats_out=True)
        y3, ws3 = TC.dwconv_norm(conv, x, stats_out=True)
    assert torch.equal(y2, got) and torch.equal(y3, TC.dwconv(conv, x))
    for yy, w_ in ((y2, ws), (y3, ws3)):
        yd = yy.double()
        want = torch.stack([yd.sum((2, 3)), (yd * yd).sum((2, 3))], 2)
        assert torch.allclose(w_[:N * C * 2].view(N, C, 2), want, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("N,C,h,w,H,W", [(2, 256, 45, 80, 90, 160), (1, 256, 23, 40, 45, 80), (1, 64, 5, 7, 9, 13), (1, 128, 8, 8, 8, 8)])
def test_depthwise_with_upsample_on_load_equals_two_kernels(cuda, N, C, h, w, H, W):
    """fots_b200_dwconv3x3_up_nhwc_bf16 = depthwise(bilinear_upsample(x_lo)) with the upsampling computed while the tile is
    staged: bit-identical to fots_b200_fpn_merge_nhwc_bf16 (upsample only) followed by the plain depthwise kernel, and within
    bf16 noise of torch's F.interpolate(align_corners=True) + grouped convolution in fp32."""
    from fots.pytorch_b200.pipeline import conv as TC, fused
    g = torch.Generator().manual_seed(N + C + h)
    conv = torch.nn.Conv2d(C, C, 3, 1, 1, groups=C, bias=False)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(C, 1, 3, 3, generator=g) / 3.0)
    conv = conv.to(cuda).to(torch.bfloat16).to(memory_format=torch.channels_last)
    lo = torch.randn(N, C, h, w, generator=g).to(cuda).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        got = TC.dwconv_up(conv, lo, (H, W))
        two = TC.dwconv(conv, fused.fpn_merge(a_lo=lo, size=(H, W)))
        up = F.interpolate(lo.float(), size=(H, W), mode="bilinear", align_corners=True)
        ref = F.conv2d(up, conv.weight.float(), None, 1, 1, groups=C)
    assert got.shape == (N, C, H, W) and torch.equal(got, two)
    assert float((got.float() - ref).abs().max()) <= 2.0 ** -6 * float(ref.abs().max()) + 1e-3
