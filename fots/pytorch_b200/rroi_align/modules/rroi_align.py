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

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        super().__init__()
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return rroi_align(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale)

    def extra_repr(self):
        return "pooled_height=%d, pooled_width=%d, spatial_scale=%g" % (
            self.pooled_height, self.pooled_width, self.spatial_scale)
