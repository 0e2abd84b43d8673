"""Greedy CTC decode on the GPU: arg-max + collapse for a whole batch of RoIs in one
launch, replacing the per-box host loop of tools/ocr_utils.py:183-186 / src/utils.py:93-97."""
import torch

from .. import _cabi
from .rois import _lib


def greedy_ctc_decode(log_probs):
    """log_probs [N, nclass, T] (what forward_ocr returns) -> (ids int32 [N, T] left-aligned zero-padded, lengths int32 [N])."""
    if not log_probs.is_cuda or log_probs.dim() != 3:
        raise ValueError("greedy_ctc_decode: expected a CUDA tensor [N, nclass, T]")
    lp = log_probs.float().contiguous()
    N, C, T = lp.shape
    ids = torch.empty((N, T), dtype=torch.int32, device=lp.device)
    lens = torch.empty((N,), dtype=torch.int32, device=lp.device)
    with torch.cuda.device(lp.device):
        st = _lib().fots_b200_ctc_greedy(lp.data_ptr(), N, C, T, ids.data_ptr(), lens.data_ptr(),
                                         torch.cuda.current_stream(lp.device).cuda_stream)
    _cabi.check(st, "fots_b200_ctc_greedy")
    return ids, lens


def ids_to_text(ids, lengths, alphabet):
    """Host-side string join (class k -> alphabet[k-1], like src/utils.py:87-97)."""
    ids, lengths = ids.cpu().tolist(), lengths.cpu().tolist()
    return ["".join(alphabet[k - 1] for k in row[:n]) for row, n in zip(ids, lengths)]
