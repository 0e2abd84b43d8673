"""Host side of the CRNN's recurrent half on B200 (include/fots_b200_pipeline.h: fots_b200_gemm_bf16w,
fots_b200_bilstm_recurrent; csrc/lstm_kernels.cu).  Replaces the cuDNN nn.LSTM + nn.Linear of the reference's
BidirectionalLSTM (tools/models.py:17-33) on the inference path: per layer one GEMM for the input projections of all time
steps, ONE persistent cluster kernel for the whole time loop of both directions, one GEMM for the embedding."""
import ctypes

import torch

from .. import _cabi


def _lib():
    L = _cabi.lib()
    if not getattr(L, "_lstm_bound", False):
        i, vp = ctypes.c_int, ctypes.c_void_p
        L.fots_b200_gemm_bf16w.restype = i
        L.fots_b200_gemm_bf16w.argtypes = [vp, i, vp, vp, vp, i, i, i, vp]
        L.fots_b200_bilstm_recurrent.restype = i
        L.fots_b200_bilstm_recurrent.argtypes = [vp, vp, vp, i, i, i, vp]
        L._lstm_bound = True
    return L


def gemm(a, w, bias=None):
    """a [M, K] bf16 or fp32 (CUDA, contiguous), w [N, K] bf16, bias fp32 [N] or None -> fp32 [M, N]."""
    M, K = a.shape
    N = w.size(0)
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        st = _lib().fots_b200_gemm_bf16w(a.data_ptr(), 1 if a.dtype == torch.float32 else 0, w.data_ptr(),
                                         bias.data_ptr() if bias is not None else None, out.data_ptr(), M, N, K,
                                         torch.cuda.current_stream(a.device).cuda_stream)
    _cabi.check(st, "fots_b200_gemm_bf16w")
    return out


class BiLSTMPack:
    """Inference snapshot of one _BiLSTM (nn.LSTM(bidirectional=True) + nn.Linear): bf16 weights, fp32 biases."""

    def __init__(self, module):
        rnn, emb = module.rnn, module.embedding
        if rnn.num_layers != 1 or not rnn.bidirectional or rnn.hidden_size != 256 or rnn.input_size % 32 != 0:
            raise ValueError("BiLSTMPack: single-layer bidirectional LSTM with hidden size 256 expected")
        f32 = lambda t: t.detach().float()
        self.hidden = rnn.hidden_size
        self.w_ih = torch.cat((f32(rnn.weight_ih_l0), f32(rnn.weight_ih_l0_reverse)), 0).to(torch.bfloat16).contiguous()   # [8H, nin]
        self.b = torch.cat((f32(rnn.bias_ih_l0) + f32(rnn.bias_hh_l0),
                            f32(rnn.bias_ih_l0_reverse) + f32(rnn.bias_hh_l0_reverse)), 0).contiguous()                    # [8H]
        self.w_hh = torch.stack((f32(rnn.weight_hh_l0), f32(rnn.weight_hh_l0_reverse)), 0).to(torch.bfloat16).contiguous()  # [2, 4H, H]
        self.w_emb = f32(emb.weight).to(torch.bfloat16).contiguous()                                                       # [nout, 2H]
        self.b_emb = f32(emb.bias).contiguous()

    def __call__(self, x):
        """x [T, N, nin] bf16 or fp32 (CUDA) -> fp32 [T, N, nout]."""
        T, N, nin = x.shape
        H = self.hidden
        x2 = x.contiguous().view(T * N, nin)
        if x2.dtype not in (torch.bfloat16, torch.float32):
            x2 = x2.float()
        g = gemm(x2, self.w_ih, self.b)                                          # [T*N, 8H]
        y = torch.empty((T, N, 2 * H), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            st = _lib().fots_b200_bilstm_recurrent(g.data_ptr(), self.w_hh.data_ptr(), y.data_ptr(), T, N, H,
                                                   torch.cuda.current_stream(x.device).cuda_stream)
        _cabi.check(st, "fots_b200_bilstm_recurrent")
        return gemm(y.view(T * N, 2 * H), self.w_emb, self.b_emb).view(T, N, -1)
