"""fots.pytorch_b200 -- B200-native (sm_100a) RoIRotate hot path of chenjun2hao/FOTS.pytorch.

Host side mirrors the reference's operator interface (rroi_align/modules/rroi_align.py:5-14,
rroi_align/functions/rroi_align.py:6-40); the compute is hand-written CUDA behind the C ABI in
include/rroi_align_b200.h (fots/pytorch_b200/lib/librroi_b200.so).  There is no CPU fallback: the
ops raise if the library is missing or a tensor is not on a CUDA device.
"""
from .rroi_align.functions.rroi_align import RRoiAlignFunction, rroi_align  # noqa: F401
from .rroi_align.modules.rroi_align import _RRoiAlign  # noqa: F401
from . import _cabi  # noqa: F401

__all__ = ["_RRoiAlign", "RRoiAlignFunction", "rroi_align"]
