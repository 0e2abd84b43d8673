import torch, torch.nn.functional as F
dev = torch.device("cuda:0")
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
with torch.no_grad():
    for cin in (3, 4, 8, 16):
        x = torch.randn(8, cin, 720, 1280, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        conv = torch.nn.Conv2d(cin, 16, 3, 1, 1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        print("conv %d->16 s1 720p NHWC bf16: %.1f us" % (cin, t(lambda: conv(x))), flush=True)
    x = torch.randn(8, 3, 720, 1280, device=dev).to(torch.bfloat16)
    conv = torch.nn.Conv2d(3, 16, 3, 1, 1, bias=False).to(dev).to(torch.bfloat16)
    print("conv 3->16 NCHW bf16: %.1f us" % t(lambda: conv(x)), flush=True)
    xf = torch.randn(8, 3, 720, 1280, device=dev).contiguous(memory_format=torch.channels_last)
    print("cast fp32->bf16 3ch: %.1f us" % t(lambda: xf.to(torch.bfloat16)))
    print("pad to 8ch + cast: %.1f us" % t(lambda: F.pad(xf, (0, 0, 0, 0, 0, 5)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)))
    for cin, cout, s in ((32, 32, 2), (64, 64, 2)):
        h = 720 if cin == 32 else 360
        x = torch.randn(8, cin, h, h * 16 // 9, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        conv = torch.nn.Conv2d(cin, cout, 3, s, 1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
        print("conv %d->%d s%d at %d: %.1f us" % (cin, cout, s, h, t(lambda: conv(x))), flush=True)
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fots.pytorch_b200.pipeline import conv as TC
    xf = torch.randn(8, 3, 720, 1280, device=dev).contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(3, 16, 3, 1, 1, bias=False).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    print("stem conv+stats (fp32 image in): %.1f us" % t(lambda: TC.stem_conv_stats(xf, conv.weight)))
