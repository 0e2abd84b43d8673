#!/usr/bin/env python
"""Consumer B (CRNN, tools/models.py:853-909) on 64 crops of 32x256: torch fp32 / torch bf16 autocast / to_b200()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fots.pytorch_b200.pipeline import CRNN


def t(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = CRNN(nclass=7500).to(dev).eval()
    x = torch.randn(64, 3, 32, 256, device=dev)
    xc = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        print("CRNN 64 x 3x32x256 (207 GFLOP): torch fp32 %.2f ms" % t(lambda: net(x)))
        netc = CRNN(nclass=7500).to(dev).eval().to(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            print("torch bf16 autocast channels-last %.2f ms" % t(lambda: netc(xc)))
        net.to_b200(dev)
        print("to_b200 (folded BN, tcgen05 conv + bias + ReLU) %.2f ms" % t(lambda: net(x)))
        print("  cnn part only %.2f ms" % t(lambda: net._cnn_b200(x)))
        seq = net._cnn_b200(x).squeeze(2).permute(2, 0, 1).contiguous()
        print("  BiLSTM x2 + embeddings, hand-written kernels (csrc/lstm_kernels.cu) %.3f ms" % t(lambda: net._rnn_b200[1](net._rnn_b200[0](seq))))
        seqf = seq.float()
        print("  BiLSTM x2 + embeddings, cuDNN nn.LSTM fp32 + cuBLAS %.3f ms" % t(lambda: net.rnn(seqf)))
        from fots.pytorch_b200.pipeline.lstm import gemm
        p0 = net._rnn_b200[0]
        g = gemm(seq.view(-1, 512), p0.w_ih, p0.b)
        print("    layer 1 input-projection GEMM %.3f ms" % t(lambda: gemm(seq.view(-1, 512), p0.w_ih, p0.b)))
        import ctypes
        from fots.pytorch_b200 import _cabi
        y = torch.empty((seq.size(0), seq.size(1), 512), device=dev)
        L = _cabi.lib()
        st = torch.cuda.current_stream().cuda_stream
        print("    layer 1 recurrent kernel (T = %d, N = %d) %.3f ms" % (seq.size(0), seq.size(1), t(lambda: L.fots_b200_bilstm_recurrent(
            g.data_ptr(), p0.w_hh.data_ptr(), y.data_ptr(), seq.size(0), seq.size(1), 256, st))))
        net._rnn_b200 = None
        print("to_b200 with the cuDNN LSTMs %.2f ms" % t(lambda: net(x)))
