"""What sits either side of RoIRotate in the reference (e, f): the shared
feature extractor that feeds it (tools/models.py:237-457), the two recognisers that consume it
(tools/models.py:334-379 and :853-909), batched RoI construction, greedy CTC decode, and the image-sharded
end-to-end inference / training steps.  Dense layers go through cuDNN/cuBLAS (library calls) in bf16
channels-last; RoIRotate itself is always the hand-written kernel of fots.pytorch_b200 (fp32 sampler)."""
from .nets import FOTSNet, CRNN  # noqa: F401
from .rois import boxes_to_rois, pooled_width_for  # noqa: F401
from .decode import greedy_ctc_decode  # noqa: F401
from .infer import FOTSPipeline  # noqa: F401
