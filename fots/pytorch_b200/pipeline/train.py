"""Training step (BASELINE.json configs[3]): backbone -> detection losses + RoIRotate -> recogniser -> CTC ->
backward (through the RoIRotate backward kernel) -> Adam.  Mirrors train.py:79-123 / src/ocr_process.py:253-301
with the per-image Python RoI selection replaced by the planted boxes of the measurement protocol.

Losses: the EAST losses of tools/models.py:459-505 are out of scope for the rewrite; this is a
dense, host-sync-free restatement of their single-scale form (dice on the score map, MSE on (sin, cos), -log IoU
on the four distances, total = seg + 2*angle + 0.5*box) using masked means instead of boolean indexing.
CTC: warp-ctc is not vendored by the reference (parity unpinned); torch's ctc_loss(sum)/N takes its place.
"""
import numpy as np
import torch
import torch.nn.functional as F

from ..rroi_align.functions.rroi_align import rroi_align
from .infer import planted_quads
from .rois import boxes_to_rois


def synthetic_targets(batch, per_image, H, W, nclass, device, seed=0, max_label=12):
    """Seeded planted boxes + rasterised score / geometry targets at 1/4 scale + random label strings."""
    rng = np.random.default_rng(10_000 + seed)
    quads = planted_quads(batch, per_image, seed0=seed, img_w=W, img_h=H)
    h4, w4 = H // 4, W // 4
    yy, xx = np.mgrid[0:h4, 0:w4].astype(np.float32)
    score = np.zeros((batch, h4, w4), np.float32)
    geo = np.zeros((batch, 4, h4, w4), np.float32)
    ang = np.zeros((batch, h4, w4), np.float32)
    for b in range(batch):
        for q in quads[b]:
            p = q[:8].reshape(4, 2) / 4.0
            c = p.mean(0)
            u = p[2] - p[1]
            v = p[0] - p[1]
            wl, hl = np.linalg.norm(u) + 1e-6, np.linalg.norm(v) + 1e-6
            du = ((xx - c[0]) * u[0] + (yy - c[1]) * u[1]) / wl
            dv = ((xx - c[0]) * v[0] + (yy - c[1]) * v[1]) / hl
            inside = (np.abs(du) < 0.35 * wl) & (np.abs(dv) < 0.35 * hl)          # shrunk quad
            score[b][inside] = 1.0
            geo[b, 0][inside] = (hl / 2 - dv)[inside]
            geo[b, 1][inside] = (hl / 2 + dv)[inside]
            geo[b, 2][inside] = (wl / 2 + du)[inside]
            geo[b, 3][inside] = (wl / 2 - du)[inside]
            ang[b][inside] = np.arctan2(u[1], u[0])
    lens = rng.integers(1, max_label + 1, batch * per_image)
    labels = rng.integers(1, nclass, int(lens.sum()))
    t = lambda a, dt=torch.float32: torch.as_tensor(a, dtype=dt, device=device)
    # label lengths also as a HOST tensor: ctc_loss wants its lengths on the CPU and would otherwise copy them back every step
    return {"quads": t(quads), "score": t(score), "geo": t(geo), "angle": t(ang),
            "labels": t(labels, torch.int32), "label_lens": t(lens, torch.int32),
            "label_lens_cpu": torch.as_tensor(lens, dtype=torch.int32)}


def detection_loss(seg, rbox, angle, tgt):
    score, geo, ang = tgt["score"], tgt["geo"], tgt["angle"]
    pred = seg.squeeze(1).float()
    inter = (pred * score).sum()
    seg_loss = -((2.0 * inter + 1.0) / (pred.sum() + score.sum() + 1.0))
    m = score
    cnt = m.sum().clamp_min(1.0)
    a = angle.float()
    ang_loss = (((a[:, 0] - torch.sin(ang)) ** 2 + (a[:, 1] - torch.cos(ang)) ** 2) * m).sum() / cnt
    r = rbox.float()
    hsum_g, hsum_p = geo[:, 0] + geo[:, 1], r[:, 0] + r[:, 1]
    hmin = torch.minimum(geo[:, 0], r[:, 0]) + torch.minimum(geo[:, 1], r[:, 1])
    box_loss = 0.0
    for k in (2, 3):
        inter_a = torch.minimum(geo[:, k], r[:, k]) * hmin
        union_a = hsum_g * geo[:, k] + hsum_p * r[:, k] - inter_a
        box_loss = box_loss + (-torch.log((inter_a + 1.0) / (union_a + 1.0)) * m).sum() / cnt
    return seg_loss + 2.0 * ang_loss + 0.5 * box_loss, {"seg": seg_loss, "angle": ang_loss, "box": box_loss}


class TrainStep:
    def __init__(self, net, lr=1e-3, pooled_height=8, pooled_width=64, spatial_scale=0.25, amp_dtype=torch.bfloat16,
                 det_weight=1.0):
        self.net = net.train()
        # train.py:40; fused = one multi-tensor kernel for all ~340 parameters instead of ~10 launches per parameter group
        on_cuda = all(p.is_cuda for p in net.parameters())
        self.opt = torch.optim.Adam(net.parameters(), lr=lr, betas=(0.5, 0.999), **({"fused": True} if on_cuda else {}))
        self.ph, self.pw, self.scale, self.amp_dtype = pooled_height, pooled_width, spatial_scale, amp_dtype
        self.det_weight = det_weight

    def __call__(self, images, tgt):
        net = self.net
        quads = tgt["quads"]
        b, R, _ = quads.shape
        x = images.contiguous(memory_format=torch.channels_last)
        self.opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            seg, rbox, angle, feats = net(x)
        det_loss, parts = detection_loss(seg[0], rbox[0], angle[0], tgt)
        focr = feats[1].float().contiguous(memory_format=torch.channels_last)          # fp32 sampler, grads flow back
        bidx = torch.arange(b, device=quads.device, dtype=torch.int32).repeat_interleave(R)
        rois = boxes_to_rois(quads.reshape(b * R, 9), bidx)
        pooled = rroi_align(focr, rois, self.ph, self.pw, self.scale)
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            logp = net.forward_ocr(pooled)                                              # [N, nclass, T] fp32 log-softmax
        N, _, T = logp.shape
        lens = tgt.get("label_lens_cpu", tgt["label_lens"])                             # host lengths: no per-step copy back
        ctc = F.ctc_loss(logp.permute(2, 0, 1), tgt["labels"], torch.full((N,), T, dtype=torch.int32, device=lens.device),
                         lens, blank=0, reduction="sum", zero_infinity=True) / N       # ocr_process.py:300-301
        total = self.det_weight * det_loss + ctc
        total.backward()
        self.opt.step()
        vals = torch.stack((total.detach(), ctc.detach(), det_loss.detach())).tolist()     # ONE device -> host read per step
        return {"total": vals[0], "ctc": vals[1], "det": vals[2]}
