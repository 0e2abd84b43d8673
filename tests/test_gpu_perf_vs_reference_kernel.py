"""Times the product forward against the reference CUDA kernel itself (oracle/_ref: rroi_align_kernel.cu
compiled unmodified for sm_100a, launched with the three zero-fills functions/rroi_align.py:17-20 needs) on
cfg1, same box, same buffers.  The reference kernel is test infrastructure, so this comparison lives in
tests/; the numbers are written to gpurun_out/ref_kernel_timing.json for profiles/."""
import json
import os

import pytest

import helpers as Hh
import workloads as WL

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, torch, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


@pytest.mark.parametrize("channels", [64, 256])
def test_new_forward_faster_than_reference_kernel(oracle, cuda, channels):
    import torch
    if not oracle.ref_gpu_available():
        pytest.skip("oracle/_ref not built")
    from fots.pytorch_b200.rroi_align.functions.rroi_align import forward_raw
    feats, rois, ph, pw, scale = WL.cfg1(channels)
    f, r = Hh.to_cuda(feats, cuda), Hh.to_cuda(rois, cuda)
    fcl = f.contiguous(memory_format=torch.channels_last)
    t_ref = _time(lambda: oracle.ref_gpu_forward(f, r, ph, pw, scale), torch)
    t_nchw = _time(lambda: forward_raw(f, r, ph, pw, scale, want_idx=True), torch)
    t_nhwc = _time(lambda: forward_raw(fcl, r, ph, pw, scale, want_idx=True), torch)
    rec = {"channels": channels, "reference_kernel_us": t_ref, "b200_nchw_us": t_nchw, "b200_nhwc_us": t_nhwc,
           "note": "eager launches through Python (L2-warm, launch-overhead bound); reference includes its 3 zero-fills"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_kernel_timing_c%d.json" % channels), "w") as fh:
        json.dump(rec, fh)
    print(rec)
    assert t_nchw < t_ref and t_nhwc < t_ref
