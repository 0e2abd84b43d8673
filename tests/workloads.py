"""Seeded synthetic workloads for the RoIRotate path (SURVEY.md section 8d / BASELINE.md section 3).

Shared by tests/, bench.py and __graft_entry__.smoke().  Pure numpy; no reference access.
"""
import math

import numpy as np

# cfg0: BASELINE.json configs[0] -- 1x3x64x64 map, 4 axis-aligned RoIs (one overhanging the border)
CFG0_ROIS = np.array([[0, 32, 32, 16, 32, 0],
                      [0, 20, 40, 8, 32, 0],
                      [0, 40, 20, 16, 16, 0],
                      [0, 32, 32, 32, 64, 0]], dtype=np.float32)


def cfg0():
    import torch
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(1, 3, 64, 64, generator=g).numpy()
    ph = 8
    pw = int(math.ceil(ph * float((CFG0_ROIS[:, 4] / CFG0_ROIS[:, 3]).max())))
    return feats, CFG0_ROIS.copy(), ph, pw, 1.0


def random_rois(seed, n=64, batch_idx=0, img_w=1280, img_h=720):
    """cfg1 generator: draw order h, ratio, cx, cy, angle (each a length-n vector)."""
    rng = np.random.default_rng(seed)
    h = rng.uniform(16, 64, n)
    ratio = rng.uniform(1, 8, n)
    cx = rng.uniform(0, img_w, n)
    cy = rng.uniform(0, img_h, n)
    ang = rng.uniform(-90, 90, n)
    rois = np.stack([np.full(n, batch_idx, np.float64), cx, cy, h, h * ratio, ang], axis=1)
    return rois.astype(np.float32)


def batch_rois(batch, per_image=64, seed0=0):
    """cfg2/3/4 generator: seed = image index, batch_idx = image index."""
    return np.concatenate([random_rois(seed0 + i, per_image, i) for i in range(batch)], axis=0)


def features(seed, b, c, h, w):
    rng = np.random.default_rng(1000 + seed)
    return rng.standard_normal((b, c, h, w), dtype=np.float32)


def cfg1(channels=64):
    return features(0, 1, channels, 180, 320), random_rois(0), 8, 64, 0.25


def valid_counts(rois, ph, pw):
    """V_n = PH * (min(PW-1, floor(rpw_n)) + 1): output elements with pw <= rpw (kernel.cu:107)."""
    rois = np.asarray(rois, np.float32)
    rpw = (rois[:, 4] * np.float32(ph)) / rois[:, 3]
    last = np.minimum(pw - 1, np.floor(rpw)).astype(np.int64)
    return ph * (np.maximum(last, -1) + 1)


def algorithmic_bytes_fwd(rois, channels, ph, pw):
    """4*N*C*PH*PW (store) + 4*C*sum V_n (one read per valid element) + 24*N."""
    n = len(rois)
    return 4 * n * channels * ph * pw + 4 * channels * int(valid_counts(rois, ph, pw).sum()) + 24 * n


# rroi_align/test2.py:36-59 scenario: three rotated quads on timg.jpeg, PH=44
TEST2_QUADS = [np.asarray([[206, 111], [199, 95], [349, 60], [355, 80]]),
               np.asarray([[312, 127], [304, 105], [367, 88], [374, 114]]),
               np.asarray([[133, 168], [118, 112], [175, 100], [185, 154]])]


def test2_rois(height_jitter=(2, 1, 2)):
    """RoI rows as rroi_align/test2.py:49-59 builds them.  The script perturbs each height with an
    unseeded random.randint(-2, 2); (+2, +1, +2) is the draw that reproduces the committed
    res0-2.jpg."""
    rois = []
    for gt, r in zip(TEST2_QUADS, height_jitter):
        center = (gt[0] + gt[1] + gt[2] + gt[3]) / 4
        dw = gt[2] - gt[1]
        dh = gt[1] - gt[0]
        w = math.sqrt(dw[0] * dw[0] + dw[1] * dw[1])
        h = math.sqrt(dh[0] * dh[0] + dh[1] * dh[1]) + r
        ang = (math.atan2(gt[2][1] - gt[1][1], gt[2][0] - gt[1][0]) +
               math.atan2(gt[3][1] - gt[0][1], gt[3][0] - gt[0][0])) / 2
        ang = -ang / 3.1415926535 * 180
        rois.append([0, center[0], center[1], h, w, ang])
    rois = np.asarray(rois, np.float32)
    ph = 44
    pw = int(math.ceil(ph * float((rois[:, 4] / rois[:, 3]).max())))
    return rois, ph, pw


def stress_rois(seed, n, batch, img_w, img_h):
    """Rotated boxes that exercise every branch of kernel.cu:58-141: any angle (incl. exact
    multiples of 90 and angles large enough for libdevice's Payne-Hanek path), boxes partly or
    wholly outside the image, 1-pixel-high boxes, aspect ratios < 1 (rpw < PH) and >> PW/PH."""
    rng = np.random.default_rng(seed)
    kind = rng.integers(0, 8, n)
    h = rng.uniform(4, 80, n)
    ratio = rng.uniform(0.2, 12, n)
    cx = rng.uniform(-0.2 * img_w, 1.2 * img_w, n)
    cy = rng.uniform(-0.2 * img_h, 1.2 * img_h, n)
    ang = rng.uniform(-180, 180, n)
    ang = np.where(kind == 1, rng.choice([0.0, 90.0, -90.0, 180.0, 45.0, -45.0, 360.0], n), ang)
    ang = np.where(kind == 2, rng.uniform(-4e7, 4e7, n), ang)          # |rad| > 105615 -> slow path
    h = np.where(kind == 3, rng.uniform(0.5, 3, n), h)                 # sub-pixel bin pitch
    ratio = np.where(kind == 4, rng.uniform(12, 60, n), ratio)         # far wider than PW
    cx = np.where(kind == 5, rng.choice([0.0, 1.0, img_w - 1.0, float(img_w)], n), cx)   # on the border
    cy = np.where(kind == 6, rng.choice([0.0, 1.0, img_h - 1.0, float(img_h)], n), cy)
    # integer-valued parameters make bin corners land exactly on x.5 (round-half-away cases)
    snap = kind == 7
    h = np.where(snap, np.round(h), h)
    cx = np.where(snap, np.round(cx), cx)
    cy = np.where(snap, np.round(cy), cy)
    ratio = np.where(snap, np.round(ratio) + 1, ratio)
    ang = np.where(snap, rng.choice([0.0, 90.0, 180.0, -90.0], n), ang)
    b = rng.integers(0, batch, n).astype(np.float64)
    return np.stack([b, cx, cy, h, h * ratio, ang], axis=1).astype(np.float32)


def planted_detection_maps(h, w, boxes, seed=0, noise=0.15):
    """Synthetic first-scale detector outputs (SURVEY 8d protocol) with `boxes` = [(cx, cy, bw, bh, angle_rad), ...] in
    MAP pixels planted in them: seg [h,w] is 0.9 inside each box shrunk by 30 % and 0.1 elsewhere; rbox [4,h,w] holds
    each inside pixel's distances (top, bottom, left, right) to the box edges and angle [2,h,w] its (sin, cos), with a
    little seeded noise, so every inside pixel decodes (nms/adaptor.cpp:85-113) to nearly the same quadrangle."""
    rng = np.random.default_rng(seed)
    seg = np.full((h, w), 0.1, np.float32)
    rbox = np.zeros((4, h, w), np.float32)
    angle = np.zeros((2, h, w), np.float32)
    angle[1] = 1.0
    ys, xs = np.mgrid[0:h, 0:w]
    px, py = xs + 0.25, ys + 0.25                     # the decode's pixel position
    for (cx, cy, bw, bh, a) in boxes:
        ca, sa = np.cos(a), np.sin(a)
        u = (px - cx) * ca + (py - cy) * sa           # along the width direction
        v = -(px - cx) * sa + (py - cy) * ca          # along the height direction
        inside = (np.abs(u) < 0.35 * bw) & (np.abs(v) < 0.35 * bh)
        seg[inside] = 0.9
        jitter = lambda: 1.0 + noise * (rng.random(int(inside.sum()), dtype=np.float32) - 0.5) * 0.1
        rbox[0][inside] = (bh / 2 + v[inside]) * jitter()        # top
        rbox[1][inside] = (bh / 2 - v[inside]) * jitter()        # bottom
        rbox[2][inside] = (bw / 2 + u[inside]) * jitter()        # left
        rbox[3][inside] = (bw / 2 - u[inside]) * jitter()        # right
        angle[0][inside] = sa
        angle[1][inside] = ca
    return seg, rbox.astype(np.float32), angle.astype(np.float32)


def planted_maps_from_quads(quads, h, w, seed0=0):
    """Detection maps (SURVEY 8d protocol) for a batch of planted quads [b, R, >=8] in IMAGE pixels (corner order of
    fots.pytorch_b200.pipeline.infer.planted_quads: p0->p1 height edge, p1->p2 width edge) at 1/4 scale:
    (seg [b,1,h,w], rbox [b,4,h,w], angle [b,2,h,w]) float32."""
    b = quads.shape[0]
    seg = np.zeros((b, 1, h, w), np.float32)
    rbox = np.zeros((b, 4, h, w), np.float32)
    ang = np.zeros((b, 2, h, w), np.float32)
    for i in range(b):
        boxes = []
        for q in quads[i]:
            p4 = np.asarray(q[:8], np.float64).reshape(4, 2) / 4.0
            c = p4.mean(0)
            dw, dh = p4[2] - p4[1], p4[1] - p4[0]
            boxes.append((c[0], c[1], float(np.hypot(*dw)), float(np.hypot(*dh)), float(np.arctan2(dw[1], dw[0]))))
        s, r, a = planted_detection_maps(h, w, boxes, seed=seed0 + i)
        seg[i, 0], rbox[i], ang[i] = s, r, a
    return seg, rbox, ang
