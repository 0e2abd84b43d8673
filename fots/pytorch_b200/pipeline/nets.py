"""Feature extractor + recognition heads, architecture-compatible with the reference's
`ModelResNetSep2` (tools/models.py:237-457; forward_ocr :334-379) and `CRNN` (:853-909).

Written from the layer inventory in SURVEY.md section 2 (#7-#9), not from the reference's code: the layer
tables below are declarative and the parameter names match the reference's `state_dict()` key for key
(tests/test_pipeline_nets.py checks names, shapes and -- when /root/reference is importable -- outputs), so a
reference checkpoint (`tools/net_utils.py:16-43`) loads unchanged.

B200 notes: convolutions are cuDNN through torch; run the module under `torch.autocast("cuda", torch.bfloat16)`
with channels-last activations (`FOTSNet.to_b200()`), InstanceNorm statistics stay fp32 inside autocast, and the
map handed to RoIRotate (`focr`) is returned as an fp32 channels-last tensor -- the layout the sampler wants.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import conv as tc
from . import fused


def _inorm(ch, affine=True):
    return nn.InstanceNorm2d(ch, eps=1e-5, momentum=0.1, affine=affine)


def _in_act(norm, x, slope, residual=None):
    """act(norm(x) [+ residual]) with leaky slope (0 = ReLU, 1 = none).  On the CUDA bf16 channels-last inference
    path this is ONE fused pair of kernels (statistics + apply); otherwise torch's ops."""
    if fused.eligible(x, residual, norm.weight, norm.bias):
        return fused.instnorm_act(x, norm.weight, norm.bias, norm.eps, slope, residual)
    if torch.is_grad_enabled() and fused.train_eligible(x, residual):
        # training step (bf16 autocast, channels-last): forward AND backward on this repo's kernels
        return fused.instnorm_act_train(x, norm.weight, norm.bias, norm.eps, slope, residual)
    y = norm(x)
    if residual is not None:
        y = y + residual
    return y if slope == 1.0 else F.leaky_relu(y, slope)


def _conv_in_act(conv, norm, x, slope, residual=None, level=1):
    """act(norm(conv(x)) [+ residual]).  On the bf16 channels-last inference path (and conv.LEVEL >= level) the
    convolution runs on the tcgen05 kernel; with conv.FUSE_STATS its epilogue also accumulates the InstanceNorm
    statistics, so the normalisation is a single pass over the convolution's output; otherwise conv + _in_act."""
    if tc.LEVEL >= level and tc.eligible(x, conv):
        if tc.FUSE_STATS and 128 <= conv.out_channels <= 1024 and conv.stride == (1, 1):
            y, ws = tc.conv2d(x, conv.weight, conv.bias, conv.padding, 1.0, stats=True)
            if fused.eligible(y, residual, norm.weight, norm.bias):
                return fused.instnorm_act(y, norm.weight, norm.bias, norm.eps, slope, residual, stats=ws)
            return _in_act(norm, y, slope, residual)
        return _in_act(norm, tc.conv2d(x, conv.weight, conv.bias, conv.padding, 1.0, stride=conv.stride), slope, residual)
    return _in_act(norm, conv(x), slope, residual)


class _CReLUNorm(nn.Module):
    """concat(x, -x) -> affine InstanceNorm -> leaky ReLU (the stem's channel-doubling activation)."""

    def __init__(self, ch):
        super().__init__()
        self.bn = _inorm(2 * ch)

    def forward(self, x):
        if fused.eligible(x, None, self.bn.weight, self.bn.bias):
            return fused.instnorm_act(x, self.bn.weight, self.bn.bias, self.bn.eps, 0.01, crelu=True)
        if torch.is_grad_enabled() and fused.train_eligible(x) and x.size(1) <= 512:
            return fused.crelu_norm_train(x, self.bn.weight, self.bn.bias, self.bn.eps, 0.01)     # training: forward + backward kernels
        return F.leaky_relu(self.bn(torch.cat((x, -x), 1)), 0.01)


def _conv(cin, cout, k, stride=1, pad=0, groups=1, bias=False, dilation=1):
    return nn.Conv2d(cin, cout, k, stride, pad, dilation=dilation, groups=groups, bias=bias)


class _ResIN(nn.Module):
    """3x3 conv - IN - ReLU - 3x3 conv - IN, residual add, ReLU (stages 1 and 2)."""

    def __init__(self, cin, cout, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(cin, cout, 3, stride, 1)
        self.bn1 = _inorm(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv(cout, cout, 3, 1, 1)
        self.bn2 = _inorm(cout)
        self.downsample = downsample

    def forward(self, x):
        res = x if self.downsample is None else _downsample(self, x)
        y = _conv_in_act(self.conv1, self.bn1, x, 0.0, level=2)
        return _conv_in_act(self.conv2, self.bn2, y, 0.0, res, level=2)


class _ResSepIN(nn.Module):
    """Depthwise-separable residual block with InstanceNorm (stages 3 and 4)."""

    def __init__(self, cin, cout, stride=1, downsample=None):
        super().__init__()
        self.conv_sep1 = nn.Sequential(
            _conv(cin, cin, 3, stride, 1, groups=cin), _conv(cin, cout, 1),
            _inorm(cout, affine=False), nn.LeakyReLU(0.01, inplace=True))
        self.conv2 = nn.Sequential(
            _conv(cout, cout, 3, 1, 1, groups=cout), _inorm(cout), nn.LeakyReLU(0.01, inplace=True),
            _conv(cout, cout, 1), _inorm(cout))
        self.downsample = downsample
        self.relu = nn.LeakyReLU(0.01, inplace=True)

    def forward(self, x):
        res = x if self.downsample is None else _downsample(self, x)
        s1, s2 = self.conv_sep1, self.conv2
        # depthwise halves: csrc/dwconv_kernels.cu; pointwise halves: 1x1 GEMMs on the tcgen05 kernel
        z, zs = tc.conv_stats_small(s1[1], tc.dwconv(s1[0], x))
        if tc.dw_eligible(z, s2[0]) and fused.eligible(z, None, s1[2].weight, s1[2].bias):
            # IN + leaky between the two halves is applied on load inside the depthwise kernel: statistics only (from the
            # pointwise convolution's epilogue when the map is small, else their own pass), the normalised tensor is never
            # written ... and the statistics of the InstanceNorm that follows come out of the depthwise kernel's epilogue
            y, ws = tc.dwconv_norm(s2[0], z, zs if zs is not None else fused.instnorm_stats(z), s1[2], 0.01, stats_out=True)
            y = fused.instnorm_act(y, s2[1].weight, s2[1].bias, s2[1].eps, 0.01, stats=ws)
        else:
            y = _in_act(s2[1], tc.dwconv(s2[0], _in_act(s1[2], z, 0.01)), 0.01)
        u, us = tc.conv_stats_small(s2[3], y)
        if us is not None and fused.eligible(u, res, s2[4].weight, s2[4].bias):
            return fused.instnorm_act(u, s2[4].weight, s2[4].bias, s2[4].eps, 0.01, res, stats=us)
        return _in_act(s2[4], u, 0.01, res)


def _downsample(block, x):
    """The residual branch of a stage's first block: Conv2d(1x1, stride 2) + eval-mode BatchNorm2d (tools/models.py:319-324).
    On the B200 inference path the BatchNorm is folded into the convolution's weights and bias (FOTSNet.to_b200) and the
    strided 1x1 convolution runs on the tcgen05 kernel; otherwise the module itself."""
    pack = getattr(block, "_ds_pack", None)
    if pack is not None and tc.LEVEL >= 2 and not block.training and tc.input_ok(x) and x.size(1) % 64 == 0:
        return tc.conv2d(x, pack[0], pack[1], (0, 0), 1.0, stride=pack[2])
    return block.downsample(x)


def _up(x, like):
    # CUDA autocast lists upsample_bilinear2d as an fp32 op: a 256-channel map at 1/4 scale would be blown up to
    # fp32 (and every add/mul after it promoted).  Interpolate in the tensor's own dtype instead.
    if torch.is_grad_enabled() and fused.train_eligible(x):
        return fused.upsample_train(x, like.shape[2:])          # training step: forward + backward on this repo's kernels
    with torch.autocast(device_type=x.device.type, enabled=False):
        return F.interpolate(x, size=like.shape[2:], mode="bilinear", align_corners=True)


class FOTSNet(nn.Module):
    """Shared convolutions + EAST-style heads + fully-convolutional recogniser.

    forward(x[B,3,H,W]) -> ([seg, seg2], [rbox, rbox2], [angle, angle2], [fpn256, focr64])   (tools/models.py:387-457)
    forward_features(x) -> focr64 only                                                        (:381-385)
    forward_ocr(pooled[N,64,PH,PW]) -> log-softmax [N, nclass, T=PW]                          (:334-379)
    """

    STAGES = ((_ResIN, 64, 3, 1), (_ResIN, 128, 4, 2), (_ResSepIN, 256, 6, 2), (_ResSepIN, 512, 4, 2))
    HEADS_ONLY_FORWARD = True          # forward(x, need_features=False) exists (FOTSPipeline asks for it)

    def __init__(self, attention=False, multi_scale=True, nclass=7500):
        super().__init__()
        self.attention, self.multi_scale = attention, multi_scale
        self.layer0 = nn.Sequential(_conv(3, 16, 3, 1, 1), _CReLUNorm(16), _conv(32, 32, 3, 2, 1), _CReLUNorm(32))
        self.layer0_1 = nn.Sequential(_conv(64, 64, 3, 1, 1), nn.ReLU(), _conv(64, 64, 3, 2, 1), nn.ReLU(inplace=True))
        # recogniser (conv6/8/9 are applied twice with shared weights; batch6/8/9 exist in checkpoints but are unused)
        self.conv5, self.conv6 = _conv(64, 128, 3, 1, 1), _conv(128, 128, 3, 1, 1)
        self.conv7, self.conv8, self.conv9 = _conv(128, 256, 3, 1, 1), _conv(256, 256, 3, 1, 1), _conv(256, 256, 3, 1, 1)
        self.conv10_s = _conv(256, 256, (2, 3), 1, (0, 1))
        self.conv11 = _conv(256, nclass, 1, bias=True)
        for name, ch in (("batch5", 128), ("batch6", 128), ("batch7", 256), ("batch8", 256), ("batch9", 256), ("batch10_s", 256)):
            setattr(self, name, _inorm(ch))
        self.max2 = nn.MaxPool2d((2, 1), stride=(2, 1))
        self.leaky = nn.LeakyReLU(0.01, inplace=True)
        # residual stages
        cin = 64
        for i, (block, ch, depth, stride) in enumerate(self.STAGES, start=1):
            down = None
            if stride != 1 or cin != ch:
                down = nn.Sequential(_conv(cin, ch, 1, stride), nn.BatchNorm2d(ch))
            setattr(self, "layer%d" % i, nn.Sequential(block(cin, ch, stride, down), *[block(ch, ch) for _ in range(depth - 1)]))
            cin = ch
        # top-down merge to 256 channels at 1/4 scale
        self.feature4, self.feature3, self.feature2 = _conv(512, 256, 1), _conv(256, 256, 1), _conv(128, 256, 1)
        self.upconv2 = nn.Sequential(_conv(256, 256, 3, 1, 1, groups=256), _conv(256, 256, 1))
        self.upconv1 = nn.Sequential(_conv(256, 256, 3, 1, 1, groups=256), _conv(256, 256, 1))
        self.feature1 = _conv(64, 256, 1)
        self.act, self.rbox, self.angle = _conv(256, 1, 1, bias=True), _conv(256, 4, 1, bias=True), _conv(256, 2, 1, bias=True)
        self.drop1 = nn.Dropout2d(p=0.2)
        if attention:
            self.conv_attenton = _conv(256, 1, 1, bias=True)     # (sic) the reference's spelling is a checkpoint key

    # ---- feeder -------------------------------------------------------------------------------
    def forward_features(self, x):
        conv0, crelu0, conv1, crelu1 = self.layer0
        if tc.stem_eligible(x, conv0) and x.size(0) * 32 <= (1 << 16):
            # first layer + the statistics of its CReLU_IN in one HBM pass (csrc/stem_conv.cu), then the apply pass
            y, ws = tc.stem_conv_stats(x, conv0.weight)
            bn = crelu0.bn
            y = fused.instnorm_act(y, bn.weight, bn.bias, bn.eps, 0.01, crelu=True, stats=ws)
            pairs = getattr(self, "_l0c1_pairs", None)
            if pairs is not None and tc.LEVEL >= 2 and tc.input_ok(y) and y.size(3) % 4 == 0:
                # 32 -> 32 stride 2: too few channels for a 64-wide k-block -- pixel pairs as 64 channels (conv.pack_pixel_pairs_s2)
                y = crelu1(tc.conv3x3_s2_pixel_pairs(y, pairs))
            else:
                y = crelu1(conv1(y))
        else:
            if x.dtype == torch.uint8:                   # raw image: the reference's host-side preprocessing (test.py:80-83)
                x = x.float() / 128 - 1
            y = self.layer0(x)
        c1, _, c2, _ = self.layer0_1
        y = tc.apply(c1, y, 0.0)                        # conv + ReLU in one kernel on the inference path
        return tc.apply(c2, y, 0.0, level=1)            # stride 2: the TMA traversal stride gathers every second pixel

    def _gate(self, x, like):
        return _up(torch.sigmoid(self.conv_attenton(x)), like)

    def _heads(self, x):
        # the 1x1 convolutions may run in bf16 under autocast; the squashing and the (sin, cos) normalisation are
        # done in fp32 (in bf16 both angle components can round to exactly 0 -> 0/0)
        packed = getattr(self, "_heads_pack", None)
        if packed is not None and tc.LEVEL >= 2 and not self.training and tc.input_ok(x) and x.size(1) in (128, 256, 512):
            return tc.heads(x, packed)                   # all three heads + squashing in one pass over x (csrc/heads_kernels.cu)
        seg = torch.sigmoid(self.act(x).float())
        rbox = torch.sigmoid(self.rbox(x).float()) * 128
        ang = torch.sigmoid(self.angle(x).float()) * 2 - 1
        ang = ang / torch.sqrt((ang * ang).sum(1, keepdim=True))
        return seg, rbox, ang

    def forward(self, x, need_features=True):
        """need_features=False (inference): the caller only wants the full-resolution heads and the recogniser's map --
        returns ([seg], [rbox], [angle], [None, focr]) and, on the B200 fast path, folds the last top-down level into the heads
        (conv.pack_merged_heads: the 256-channel map at 1/4 scale and the 1/8-scale auxiliary heads are never computed)."""
        focr = self.forward_features(x)
        s3 = self.layer1(self.drop1(focr))
        s2 = self.layer2(s3)
        s1 = self.layer3(s2)
        pw = lambda conv, t: tc.apply(conv, t, 1.0, level=2)              # 1x1 laterals
        f2, f3 = pw(self.feature2, s2), pw(self.feature3, s1)
        f4 = pw(self.feature4, self.drop1(self.layer4(s1)))
        if fused.merge_eligible(f2, f3, f4) and fused.merge_eligible(s3):
            # inference fast path: each merge step is one fused kernel (upsample + attention gate + add)
            apack = getattr(self, "_att_pack", None)
            if self.attention and apack is not None and tc.LEVEL >= 2 and f4.size(1) in (128, 256, 512):
                # one pass over t -> sigmoid(logits) in bf16 (what torch's autocast path interpolates); the merge kernel then
                # only interpolates the gate
                gate = lambda t: dict(gate_prob_lo=tc.conv1x1_to1(t, apack, sigmoid=True))
            elif self.attention:
                gate = lambda t: dict(gate_logits_lo=self.conv_attenton(t))
            else:
                gate = lambda t: {}
            x = fused.fpn_merge(a_lo=f4, b_hi=f3, **gate(f4))
            def up(seq, lo, size):
                # upconv(F.interpolate(lo)): depthwise 3x3 + pointwise 1x1 on the 2x bilinear upsampling of `lo`.  The upsampling
                # is computed while the depthwise kernel stages its tile, so the upsampled map is never written.
                if tc.DW_UP and tc.dw_eligible(lo, seq[0]) and seq[0].stride == (1, 1):
                    return pw(seq[1], tc.dwconv_up(seq[0], lo, size))
                return pw(seq[1], tc.dwconv(seq[0], fused.fpn_merge(a_lo=lo, size=size)))
            f2 = fused.fpn_merge(c_hi=up(self.upconv1, x, f2.shape[2:]), b_hi=f2, **gate(x))
            mh = getattr(self, "_merged_heads", None)
            g2 = gate(f2)
            if (not need_features and mh is not None and tc.MERGED_HEADS and tc.LEVEL >= 2 and not self.training and "gate_prob_lo" in g2
                    and tc.DW_UP and tc.dw_eligible(f2, self.upconv2[0]) and f2.size(1) == 256 and s3.size(1) == 64):
                a72 = getattr(self, "_gather_heads", None)
                if a72 is not None and tc.GATHER_HEADS and f2.is_contiguous(memory_format=torch.channels_last):
                    # the depthwise half of upconv2 and its upsampling folded in too: a 256 -> 72 GEMM at 1/8 scale + a gather
                    seg, rbox, ang = tc.heads_gather(f2, s3, g2["gate_prob_lo"], a72, mh)
                else:
                    d = tc.dwconv_up(self.upconv2[0], f2, s3.shape[2:])
                    seg, rbox, ang = tc.heads_merged(d, s3, g2["gate_prob_lo"], mh)
                return [seg], [rbox], [ang], [None, focr]
            f1 = pw(self.feature1, s3)
            x = fused.fpn_merge(c_hi=up(self.upconv2, f2, f1.shape[2:]), b_hi=f1, **g2)
            if not need_features:
                seg, rbox, ang = self._heads(x)
                return [seg], [rbox], [ang], [None, focr]
            seg2, rbox2, ang2 = self._heads(f2)
            x = self.drop1(x)
            seg, rbox, ang = self._heads(x)
            return [seg, seg2], [rbox, rbox2], [ang, ang2], [x, focr]
        f1 = pw(self.feature1, s3)
        if self.attention:
            # (the reference upsamples the gate expanded to all channels; upsampling the one-channel gate and broadcasting is
            # the same arithmetic per element)
            x = _up(f4, f3) + f3 * self._gate(f4, f3)
            gate = self._gate(x, f2)
            f2 = self.upconv1(_up(x, f2)) + f2 * gate
            gate = self._gate(f2, f1)
            x = self.upconv2(_up(f2, f1)) + f1 * gate
        else:
            x = _up(f4, f3) + f3
            f2 = self.upconv1(_up(x, f2)) + f2
            x = self.upconv2(_up(f2, f1)) + f1
        seg2, rbox2, ang2 = self._heads(f2)
        x = self.drop1(x)
        seg, rbox, ang = self._heads(x)
        return [seg, seg2], [rbox, rbox2], [ang, ang2], [x, focr]

    # ---- consumer A ---------------------------------------------------------------------------
    def _max2(self, x):
        if fused.eligible(x) and x.size(2) >= 2:
            return fused.maxpool_h2(x)
        return self.max2(x)

    def forward_ocr(self, x, log_probs=True):
        """log_probs=False returns the raw class scores [N, nclass, T] (fp32): the greedy decode only needs the arg-max,
        so the inference step skips the log-softmax pass (tools/ocr_utils.py:183 takes the arg-max of it)."""
        # tc.apply = conv + leaky-ReLU in one tcgen05 kernel on the bf16 channels-last inference path, torch otherwise
        x = _conv_in_act(self.conv5, self.batch5, x, 0.01)
        x = tc.apply(self.conv6, tc.apply(self.conv6, x, 0.01), 0.01)
        x = _conv_in_act(self.conv7, self.batch7, self._max2(x), 0.01)
        x = tc.apply(self.conv8, tc.apply(self.conv8, x, 0.01), 0.01)
        x = tc.apply(self.conv9, tc.apply(self.conv9, x, 0.01), 0.01)
        x = _conv_in_act(self.conv10_s, self.batch10_s, self._max2(x), 0.01)
        pad = getattr(self, "_conv11_pad", None)
        if tc.LEVEL >= 1 and pad is not None and not self.training and tc.input_ok(x) and x.size(1) % 64 == 0:
            # the classifier as a 1x1 convolution on the tcgen05 kernel: nclass padded to a multiple of 64 output channels
            x = tc.conv2d(x, pad[0], pad[1], (0, 0), 1.0)[:, :self.conv11.out_channels]
        else:
            x = self.conv11(self.drop1(x))
        x = x.squeeze(2).float()                             # [N, nclass, T]
        return F.log_softmax(x, dim=1) if log_probs else x

    # ---- B200 placement ------------------------------------------------------------------------
    def to_b200(self, device="cuda", inference=False):
        """Placement for B200: parameters on the device, convolution weights channels-last; call under bf16
        autocast.  inference=True additionally stores the convolution weights in bf16 once, so autocast does not
        re-cast ~100 fp32 weight tensors on every forward (norm parameters stay fp32: the fused kernels and
        batch_norm read them as such).  Keep inference=False for training (fp32 master weights)."""
        self.to(device=device, memory_format=torch.channels_last)
        self._conv11_pad = self._heads_pack = self._l0c1_pairs = self._att_pack = self._merged_heads = self._gather_heads = None
        if inference:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    m.to(dtype=torch.bfloat16)
            self.eval()
            # classifier weights with the class dimension padded to a multiple of 64 (zero rows), bias in fp32
            c11 = self.conv11
            cpad = (c11.out_channels + 63) // 64 * 64
            w = torch.zeros((cpad, c11.in_channels, 1, 1), dtype=torch.bfloat16, device=c11.weight.device)
            w[:c11.out_channels] = c11.weight.detach()
            bias = torch.zeros((cpad,), dtype=torch.float32, device=w.device)
            bias[:c11.out_channels] = c11.bias.detach().float()
            self._conv11_pad = (w.contiguous(memory_format=torch.channels_last), bias)
            self._heads_pack = tc.pack_heads(self.act, self.rbox, self.angle)
            if self.attention:
                self._att_pack = tc.pack_to1(self.conv_attenton)
                # the last top-down level folded into the heads (forward(..., need_features=False))
                self._merged_heads = tc.pack_merged_heads(self.act, self.rbox, self.angle, self.upconv2[1], self.feature1)
                self._gather_heads = tc.pack_gather_heads(self.act, self.rbox, self.angle, self.upconv2[1], self.upconv2[0])
            c01 = self.layer0[2]
            if c01.in_channels == 32 and c01.out_channels == 32 and c01.bias is None:
                self._l0c1_pairs = tc.pack_pixel_pairs_s2(c01.weight.detach())
            # down-sampling branches: eval-mode BatchNorm folded into the 1x1 stride-2 convolution (bf16 weights, fp32 bias)
            for m in self.modules():
                ds = getattr(m, "downsample", None)
                if isinstance(ds, nn.Sequential) and len(ds) == 2 and isinstance(ds[1], nn.BatchNorm2d) and ds[0].bias is None:
                    conv, bn = ds
                    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
                    w = (conv.weight.detach().float() * scale.view(-1, 1, 1, 1)).to(torch.bfloat16)
                    b = (bn.bias.detach().float() - bn.running_mean.detach().float() * scale).contiguous()
                    m._ds_pack = (w.contiguous(memory_format=torch.channels_last), b, conv.stride[0])
        else:
            for m in self.modules():
                if hasattr(m, "_ds_pack"):
                    m._ds_pack = None
        return self


class _BiLSTM(nn.Module):
    def __init__(self, nin, hidden, nout):
        super().__init__()
        self.rnn = nn.LSTM(nin, hidden, bidirectional=True)
        self.embedding = nn.Linear(2 * hidden, nout)

    def forward(self, x):                                    # [T, N, nin]
        y, _ = self.rnn(x)
        return self.embedding(y)                             # Linear acts on the last dim: same as the view/unview


class CRNN(nn.Module):
    """Consumer B (tools/models.py:853-909): 7-conv CNN collapsing H 32 -> 1, two BiLSTMs.  Input is RoIRotate of
    the raw image with PH = 32 (src/utils.py:430-436); output [T = W/4 + 1, N, nclass]."""

    #        cout  k  pad  batchnorm  pool-after
    PLAN = ((64, 3, 1, False, (2, 2, 2, 2, 0, 0)), (128, 3, 1, False, (2, 2, 2, 2, 0, 0)),
            (256, 3, 1, True, None), (256, 3, 1, False, (2, 2, 2, 1, 0, 1)),
            (512, 3, 1, True, None), (512, 3, 1, False, (2, 2, 2, 1, 0, 1)), (512, 2, 0, True, None))

    def __init__(self, nclass=7500, hidden=256):
        super().__init__()
        cnn, cin, pool_i = nn.Sequential(), 3, 0
        for i, (cout, k, pad, bn, pool) in enumerate(self.PLAN):
            cnn.add_module("conv%d" % i, nn.Conv2d(cin, cout, k, 1, pad))
            if bn:
                cnn.add_module("batchnorm%d" % i, nn.BatchNorm2d(cout))
            cnn.add_module("relu%d" % i, nn.ReLU(True))
            if pool is not None:
                kh, kw, sh, sw, ph, pw = pool
                cnn.add_module("pooling%d" % pool_i, nn.MaxPool2d((kh, kw), (sh, sw), (ph, pw)))
                pool_i += 1
            cin = cout
        self.cnn = cnn
        self.rnn = nn.Sequential(_BiLSTM(512, hidden, hidden), _BiLSTM(hidden, hidden, nclass))

    def forward(self, x):
        # the folded-BatchNorm bf16 path is a frozen inference snapshot of the weights: never under training or autograd
        fast = getattr(self, "_b200", None)
        if fast is not None and x.is_cuda and not self.training and not (
                torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            f = self._cnn_b200(x)
            if f.size(2) != 1:
                raise ValueError("CRNN: input height must reduce to 1 (use PH = 32)")
            seq = f.squeeze(2).permute(2, 0, 1).contiguous()             # [T, N, 512] bf16
            if self._rnn_b200 is not None:
                # both BiLSTMs + embeddings on the hand-written kernels (csrc/lstm_kernels.cu): fp32 [T, N, nclass]
                return self._rnn_b200[1](self._rnn_b200[0](seq))
            return self.rnn(seq.float())
        f = self.cnn(x)
        if f.size(2) != 1:
            raise ValueError("CRNN: input height must reduce to 1 (use PH = 32)")
        return self.rnn(f.squeeze(2).permute(2, 0, 1).contiguous())

    def train(self, mode=True):
        if mode:
            self._b200 = self._rnn_b200 = None   # the folded weights are stale as soon as the parameters may change
        return super().train(mode)

    # ---- B200 placement ------------------------------------------------------------------------
    def to_b200(self, device="cuda"):
        """Inference placement: every convolution gets its (eval-mode) BatchNorm folded into bf16 channels-last weights
        and an fp32 bias, so that each of the seven layers is ONE convolution with bias + ReLU in the epilogue -- on the
        tcgen05 kernel (csrc/conv_tc.cu) wherever Cin % 64 == 0 (six of the seven layers, 99 % of the FLOPs), on the
        library for the 3-channel first layer.  The two BiLSTMs + embeddings run on csrc/lstm_kernels.cu (bf16 weights,
        fp32 state; FOTS_B200_LSTM=0 keeps cuDNN for A/B timing).  Call again after loading a checkpoint."""
        self.to(device).eval()
        packs = []
        for i, (cout, k, pad, bn, pool) in enumerate(self.PLAN):
            conv = getattr(self.cnn, "conv%d" % i)
            w, b = conv.weight.detach().float(), conv.bias.detach().float()
            if bn:
                m = getattr(self.cnn, "batchnorm%d" % i)
                scale = m.weight.detach().float() / torch.sqrt(m.running_var.detach().float() + m.eps)
                w = w * scale.view(-1, 1, 1, 1)
                b = (b - m.running_mean.detach().float()) * scale + m.bias.detach().float()
            packs.append((w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last), b.contiguous(), pad, pool))
        self._b200 = packs
        self._w0_plain = packs[0][0].contiguous()             # first layer's [Cout, 3, 3, 3] in plain layout for the 3-channel kernel
        self._rnn_b200 = None
        if os.environ.get("FOTS_B200_LSTM", "1") != "0":
            from .lstm import BiLSTMPack
            self._rnn_b200 = [BiLSTMPack(self.rnn[0]), BiLSTMPack(self.rnn[1])]
        return self

    @torch.no_grad()
    def _cnn_b200(self, x):
        y = x
        for i, (w, b, pad, pool) in enumerate(self._b200):
            pooled = False
            if tc.ENABLED and i == 0 and w.size(1) == 3 and w.shape[2:] == (3, 3) and pad == 1 and w.size(0) % 8 == 0 and w.size(0) <= 256:
                # first layer: conv + bias + ReLU + (2,2)/(2,2) pooling in one kernel straight from the fp32 NCHW crops
                pooled = pool == (2, 2, 2, 2, 0, 0) and y.size(2) % 2 == 0 and y.size(3) % 2 == 0
                y = tc.conv3x3_c3_pool(y, self._w0_plain, b, pooled)
            else:
                if y.dtype != torch.bfloat16 or not y.is_contiguous(memory_format=torch.channels_last):
                    y = y.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                if tc.ENABLED and w.size(1) % 64 == 0:
                    y = tc.conv2d(y, w, b, (pad, pad), 0.0)                  # conv + folded BN + bias + ReLU, one kernel
                else:
                    y = torch.relu(F.conv2d(y, w, b.to(torch.bfloat16), 1, pad))
            if pool is not None and not pooled:
                kh, kw, sh, sw, ph, pw = pool
                if tc.ENABLED and y.size(1) % 8 == 0 and y.is_contiguous(memory_format=torch.channels_last):
                    y = tc.maxpool(y, (kh, kw), (sh, sw), (ph, pw))
                else:
                    y = F.max_pool2d(y, (kh, kw), (sh, sw), (ph, pw))
        return y
