"""GPU check of the tcgen05 convolution against torch (fp32 reference on the same bf16-rounded operands).
Run under gpurun; every case prints max abs error and, on a mismatch, which (channel, row, column) pattern is off."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from fots.pytorch_b200.pipeline import conv as TC

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def case(N, H, W, Cin, Cout, R, S, ph, pw, bias, slope, bn=0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, R, S, generator=g) / (Cin * R * S) ** 0.5).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(Cout, generator=g).to(dev) if bias else None
    TC.set_tile(bn)
    y = TC.conv2d(x, w, b, (ph, pw), slope)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), b, 1, (ph, pw))
    if slope != 1.0:
        ref = F.leaky_relu(ref, slope)
    err = (y.float() - ref).abs()
    tol = 2e-2 * ref.abs().max().item()
    ok = err.max().item() <= tol
    print("%s N%d %dx%d Cin%d Cout%d %dx%d pad(%d,%d) bias=%d slope=%g bn=%d: max err %.4g (ref max %.3g)" % (
        "ok  " if ok else "FAIL", N, H, W, Cin, Cout, R, S, ph, pw, bias, slope, bn, err.max().item(), ref.abs().max().item()), flush=True)
    if not ok:
        bad = (err > tol)
        print("   bad fraction %.4f; by channel-block of 32:" % bad.float().mean().item(),
              [round(bad[:, i:i + 32].float().mean().item(), 3) for i in range(0, Cout, 32)][:8])
        print("   by h:", [round(bad[:, :, i].float().mean().item(), 3) for i in range(min(bad.shape[2], 8))],
              " by w:", [round(bad[:, :, :, i].float().mean().item(), 3) for i in range(min(bad.shape[3], 16))])
    return ok


def bench(N, H, W, Cin, Cout, R, S, ph, pw, bn=0, iters=20):
    x = torch.randn(N, Cin, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(Cin, Cout, (R, S), 1, (ph, pw), bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    TC.set_tile(bn)
    flops = 2.0 * N * (H + 2 * ph - R + 1) * (W + 2 * pw - S + 1) * Cout * Cin * R * S
    res = []
    for name, fn in (("tcgen05", lambda: TC.conv2d(x, conv.weight, None, (ph, pw), 0.01)),
                     ("tcgen05+stats", lambda: TC.conv2d(x, conv.weight, None, (ph, pw), 1.0, stats=True)),
                     ("cudnn+leaky", lambda: F.leaky_relu(conv(x), 0.01)), ("cudnn", lambda: conv(x))):
        with torch.no_grad():
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res.append("%s %.3f ms %.0f TF/s" % (name, ms, flops / ms / 1e9))
    print("bench N%d %dx%d %d->%d %dx%d bn=%d: %s" % (N, H, W, Cin, Cout, R, S, bn, " | ".join(res)), flush=True)


if __name__ == "__main__":
    ok = True
    # plain GEMM first (1x1), then padding, then the real shapes
    ok &= case(1, 2, 64, 64, 64, 1, 1, 0, 0, False, 1.0)
    ok &= case(1, 2, 64, 64, 128, 1, 1, 0, 0, False, 1.0)
    ok &= case(2, 4, 64, 128, 128, 1, 1, 0, 0, True, 1.0)
    ok &= case(2, 4, 64, 64, 64, 3, 3, 1, 1, False, 1.0)
    ok &= case(3, 8, 64, 128, 128, 3, 3, 1, 1, True, 0.01)
    ok &= case(3, 4, 64, 256, 256, 3, 3, 1, 1, False, 0.01)
    ok &= case(3, 4, 64, 256, 256, 3, 3, 1, 1, False, 0.01, bn=256)
    ok &= case(2, 4, 64, 256, 256, 3, 3, 1, 1, False, 0.0, bn=64)
    ok &= case(5, 2, 64, 256, 256, 2, 3, 0, 1, True, 1.0)          # conv10_s: 2x3, pad (0,1), Ho = 1, tiles span images
    ok &= case(1, 45, 80, 64, 64, 3, 3, 1, 1, False, 1.0)          # ragged tiles in h and w
    ok &= case(2, 23, 37, 128, 192, 3, 3, 1, 1, True, 0.0)         # odd sizes, Cout = 3 x 64
    ok &= case(1, 180, 320, 64, 64, 3, 3, 1, 1, False, 1.0)
    # CTA-pair variant (cta_group::2): bn = 512 selects it for Cout % 256 == 0
    ok &= case(3, 4, 64, 256, 256, 3, 3, 1, 1, False, 0.01, bn=512)      # 6 pixel tiles -> 3 pairs
    ok &= case(5, 2, 64, 256, 256, 2, 3, 0, 1, True, 1.0, bn=512)        # 3 pixel tiles -> the last pair is half empty
    ok &= case(2, 23, 37, 128, 512, 3, 3, 1, 1, True, 0.0, bn=512)       # ragged, two cout tiles
    ok &= case(64, 4, 64, 256, 256, 3, 3, 1, 1, False, 0.01, bn=512)     # more pairs than clusters: several tiles per CTA
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    if ok and len(sys.argv) > 1 and sys.argv[1] == "bench":
        for bn in (512, 256):
            bench(512, 4, 64, 256, 256, 3, 3, 1, 1, bn)          # conv8 / conv9
            bench(512, 4, 64, 128, 256, 3, 3, 1, 1, bn)          # conv7
            bench(512, 2, 64, 256, 256, 2, 3, 0, 1, bn)          # conv10_s
        for bn in (128, 64):
            bench(512, 4, 64, 256, 256, 3, 3, 1, 1, bn)          # conv8 / conv9
        bench(512, 8, 64, 64, 128, 3, 3, 1, 1, 128)              # conv5
        bench(512, 8, 64, 128, 128, 3, 3, 1, 1, 128)             # conv6
        bench(512, 8, 64, 128, 128, 3, 3, 1, 1, 64)
        bench(512, 4, 64, 128, 256, 3, 3, 1, 1, 256)             # conv7
        bench(512, 4, 64, 128, 256, 3, 3, 1, 1, 128)
        bench(512, 2, 64, 256, 256, 2, 3, 0, 1, 256)             # conv10_s
        bench(8, 360, 640, 64, 64, 3, 3, 1, 1, 64)               # layer0_1[0]
        bench(8, 180, 320, 64, 64, 3, 3, 1, 1, 64)               # layer1 blocks
        bench(8, 90, 160, 128, 128, 3, 3, 1, 1, 128)             # layer2 blocks
        bench(8, 90, 160, 128, 128, 3, 3, 1, 1, 64)
    sys.exit(0 if ok else 1)
