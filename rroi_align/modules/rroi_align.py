from fots.pytorch_b200.rroi_align.modules.rroi_align import _RRoiAlign  # noqa: F401
