from fots.pytorch_b200.rroi_align.functions.rroi_align import RRoiAlignFunction, rroi_align  # noqa: F401
