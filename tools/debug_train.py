#!/usr/bin/env python
"""Where does the cfg3 training step first go non-finite?  (development tool, run under gpurun)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from fots.pytorch_b200.pipeline import FOTSNet  # noqa: E402
from fots.pytorch_b200.pipeline.train import TrainStep, synthetic_targets  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
net = FOTSNet(attention=True, nclass=89).to(dev).to(memory_format=torch.channels_last)
step = TrainStep(net, lr=1e-4)
images = torch.randn(B, 3, 720, 1280, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
tgt = synthetic_targets(B, 64, 720, 1280, 89, dev, seed=0)


def report(tag):
    bad = [(n, "param") for n, p in net.named_parameters() if not torch.isfinite(p).all()]
    badg = [(n, float(p.grad.abs().max())) for n, p in net.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    big = sorted(((float(p.grad.abs().max()), n) for n, p in net.named_parameters() if p.grad is not None and torch.isfinite(p.grad).all()), reverse=True)[:5]
    print(tag, "non-finite params:", bad[:5], "non-finite grads:", badg[:8], "largest finite grads:", big, flush=True)


hooks = []
def mk(name):
    def h(mod, inp, out):
        o = out[0] if isinstance(out, (tuple, list)) else out
        if torch.is_tensor(o) and not torch.isfinite(o).all():
            print("  non-finite activation out of", name, flush=True)
    return h
for n, m in net.named_modules():
    if len(list(m.children())) == 0:
        hooks.append(m.register_forward_hook(mk(n)))
for i in range(3):
    out = step(images, tgt)
    print("step", i, out, flush=True)
    report("after step %d" % i)
