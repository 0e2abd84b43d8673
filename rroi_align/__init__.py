"""Drop-in import path of the reference package `rroi_align` (chenjun2hao/FOTS.pytorch).

`from rroi_align.modules.rroi_align import _RRoiAlign` and
`from rroi_align.functions.rroi_align import RRoiAlignFunction` resolve to the B200-native
implementation in fots.pytorch_b200.
"""
