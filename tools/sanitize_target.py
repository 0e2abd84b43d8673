#!/usr/bin/env python
"""Small launches of this round's new kernels for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fots.pytorch_b200 import _cabi  # noqa: E402
from fots.pytorch_b200.pipeline import conv as TC, fused  # noqa: E402
from fots.pytorch_b200.rroi_align.functions.rroi_align import backward_raw, forward_raw  # noqa: E402
import workloads as WL  # noqa: E402

dev = torch.device("cuda:0")
cl = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)

with torch.no_grad():
    # depthwise: plain / norm (+ statistics) / upsample-on-load, ragged sizes
    for (N, C, H, W) in ((2, 128, 21, 37), (1, 64, 7, 5), (1, 256, 45, 80)):
        conv = torch.nn.Conv2d(C, C, 3, 1, 1, groups=C, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        norm = torch.nn.InstanceNorm2d(C, affine=True).to(dev)
        x = cl(N, C, H, W)
        TC.dwconv(conv, x)
        TC.dwconv_norm(conv, x, fused.instnorm_stats(x), norm, 0.01, stats_out=True)
        TC.dwconv_norm(conv, x, stats_out=True)
        lo = cl(N, C, (H + 1) // 2, (W + 1) // 2)
        TC.dwconv_up(conv, lo, (H, W))
    # InstanceNorm variants, merge kernels, gate convolution
    x = cl(2, 64, 33, 20)
    g, b = torch.randn(64, device=dev), torch.randn(64, device=dev)
    fused.set_single_pass(0)
    fused.instnorm_act(x, g, b, 1e-5, 0.01, cl(2, 64, 33, 20))
    fused.instnorm_act(cl(2, 16, 33, 20), torch.randn(32, device=dev), torch.randn(32, device=dev), 1e-5, 0.01, crelu=True)
    fused.set_single_pass(1)
    a_lo, c_hi, b_hi = cl(2, 256, 6, 10), cl(2, 256, 12, 20), cl(2, 256, 12, 20)
    act = torch.nn.Conv2d(256, 1, 1).to(dev).to(torch.bfloat16)
    gate = TC.conv1x1_to1(a_lo, TC.pack_to1(act), sigmoid=True)
    fused.fpn_merge(a_lo=a_lo, b_hi=b_hi, gate_prob_lo=gate)
    fused.fpn_merge(c_hi=c_hi, b_hi=b_hi, gate_prob_lo=gate)
    fused.fpn_merge(c_hi=c_hi, b_hi=b_hi, gate_logits_lo=gate)
    # the last top-down level folded into the heads: tap map (tcgen05 1x1 convolution) + gather, odd sizes and a ragged tile
    for (B_, h_, w_, H_, W_) in ((2, 12, 20, 24, 40), (1, 7, 9, 13, 17)):
        f2, s3 = cl(B_, 256, h_, w_), cl(B_, 64, H_, W_)
        gp = torch.rand(B_, 1, h_, w_, device=dev).to(torch.bfloat16)
        hd = [torch.nn.Conv2d(256, k, 1).to(dev) for k in (1, 4, 2)]
        pw_, lat_ = torch.nn.Conv2d(256, 256, 1, bias=False).to(dev), torch.nn.Conv2d(64, 256, 1, bias=False).to(dev)
        dw_ = torch.nn.Conv2d(256, 256, 3, 1, 1, groups=256, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        TC.heads_gather(f2, s3, gp, TC.pack_gather_heads(hd[0], hd[1], hd[2], pw_, dw_), TC.pack_merged_heads(hd[0], hd[1], hd[2], pw_, lat_))
    # CRNN front end
    w = (torch.randn(64, 3, 3, 3, device=dev) / 5).to(torch.bfloat16)
    y = TC.conv3x3_c3_pool(torch.randn(3, 3, 32, 100, device=dev), w, torch.randn(64, device=dev), True)
    TC.maxpool(y, (2, 2), (2, 1), (0, 1))
    # a small tcgen05 convolution with statistics (shared-memory reduced flush)
    xs = cl(2, 64, 16, 24)
    ws = (torch.randn(128, 64, 3, 3, device=dev) / 24).to(torch.bfloat16)
    TC.conv2d(xs, ws, None, (1, 1), 1.0, stats=True)

# NCHW backward: row segments + gather, aligned and unaligned maps
for (H, W) in ((45, 80), (45, 79)):
    feats = WL.features(3, 2, 16, H, W)
    rois = WL.stress_rois(33, 40, 2, W * 4, H * 4)
    f, r = torch.from_numpy(feats).to(dev), torch.from_numpy(rois).to(dev)
    out, ix, iy, _ = forward_raw(f, r, 8, 64, 0.25, want_idx=True)
    gt = torch.randn_like(out)
    backward_raw(gt, r, ix, iy, feats.shape, 0.25, _cabi.LAYOUT_NCHW, opts=_cabi.opts(bwd_mode=4))
    backward_raw(gt, r, None, None, feats.shape, 0.25, _cabi.LAYOUT_NCHW, opts=_cabi.opts(bwd_mode=4))
# NCHW forward, row segments in fixed-stride channel slots: double buffer and both ring forms; RoI heights chosen so that
# tiles take 8, 4, 2 and 1 channels per stage; C = 7 leaves a partial last stage; W = 79 takes the gather fallback
for (C, H, W) in ((16, 90, 160), (7, 90, 160), (16, 45, 79)):
    feats = WL.features(5, 2, C, H, W)
    rois = np.concatenate([WL.stress_rois(41, 24, 2, W * 4, H * 4),
                           np.array([[0, 320, 180, 40, 300, 10], [1, 300, 200, 90, 600, -25], [0, 320, 180, 160, 640, 40],
                                     [1, 320, 180, 250, 640, 75]], np.float32)], 0)
    f, r = torch.from_numpy(feats).to(dev), torch.from_numpy(rois).to(dev)
    for v in (2, 3, 4):
        forward_raw(f, r, 8, 64, 0.25, opts=_cabi.opts(variant=v))
        forward_raw(f, r, 8, 64, 0.25, want_idx=True, opts=_cabi.opts(variant=v, rois_ready=True))
torch.cuda.synchronize()
print("sanitize target done")
