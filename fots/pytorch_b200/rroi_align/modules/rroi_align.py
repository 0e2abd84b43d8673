"""nn.Module face of RoIRotate.  Mirrors rroi_align/modules/rroi_align.py:5-14 of the reference:
same class name, constructor arguments, attribute names and forward(features, rois) contract."""
from torch import nn

from ..functions.rroi_align import rroi_align


class _RRoiAlign(nn.Module):
    """RoIRotate: crop N rotated boxes out of a shared feature map into [N, C, PH, PW].

    features: fp32 CUDA tensor [B, C, H, W] -- contiguous (reference layout) or channels_last (the
              layout the B200 kernels prefer; the result is then channels_last as well).
    rois:     fp32 CUDA tensor [N, 6] = [batch_idx, cx, cy, h, w, angle_deg] in input-image pixels.
    """

    def __init__(self, pooled_height, pooled_width, spatial_scale, concurrency=0, rois_ready=False):
        """The first three arguments are the reference's.  The optional two are per-call launch hints of the B200 library
        (rroi_b200_opts, include/rroi_align_b200.h): `concurrency` = independent RoIRotate launches the caller keeps in
        flight on other streams (picks the tile size), `rois_ready` = the RoI rows were on the device before the kernel that
        precedes this call on the stream (lets the forward's prologue overlap that kernel's tail).  Results never depend
        on them."""
        super().__init__()
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.concurrency = int(concurrency)
        self.rois_ready = bool(rois_ready)

    def forward(self, features, rois):
        return rroi_align(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale,
                          concurrency=self.concurrency, rois_ready=self.rois_ready)

    def extra_repr(self):
        return "pooled_height=%d, pooled_width=%d, spatial_scale=%g" % (
            self.pooled_height, self.pooled_width, self.spatial_scale)
