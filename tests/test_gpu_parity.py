"""GPU parity tests (the parity tests proper): the sm_100a kernels, called through the Python host
API -> C ABI, against
  (1) the CPU oracle (oracle/rroi_oracle.c, restating rroi_align_kernel.cu:28-162 / :193-278) and
  (2) the reference CUDA kernel itself, compiled unmodified for sm_100a (oracle/_ref), when present.
Bar: sample centres (index work) and sampled values bit-exact in the forward -- the weights are
{0,1/4,1/2,1} so only the sum order matters and it is reproduced; backward within 1e-4 relative
(fp32 atomics: the reference's own sum order is unspecified).
"""
import numpy as np
import pytest

import helpers as Hh
import workloads as WL

pytestmark = pytest.mark.gpu

REL_BWD = 1e-4   # north_star tolerance for sampled / accumulated fp32 values


def _ref_gpu(oracle, feats, rois, ph, pw, scale, device):
    f, r = Hh.to_cuda(feats, device), Hh.to_cuda(rois, device)
    out, ix, iy = oracle.ref_gpu_forward(f, r, ph, pw, scale)
    return out.cpu().numpy(), ix.cpu().numpy(), iy.cpu().numpy()


def _check_forward_all_paths(oracle, feats, rois, ph, pw, scale, device, vs_cpu=True, vs_ref=True):
    C = feats.shape[1]
    got = {}
    got["nchw"] = Hh.run_new_forward(feats, rois, ph, pw, scale, device, channels_last=False)
    got["nhwc"] = Hh.run_new_forward(feats, rois, ph, pw, scale, device, channels_last=True)
    lo, lx, ly = Hh.run_legacy_forward(feats, rois, ph, pw, scale, device)
    # legacy idx is [N,C,PH,PW]; all channels must agree with the compact centres
    Hh.assert_bit_equal(lx, Hh.expand_idx(got["nchw"][1], C), "legacy idx_x vs compact")
    Hh.assert_bit_equal(ly, Hh.expand_idx(got["nchw"][2], C), "legacy idx_y vs compact")
    Hh.assert_bit_equal(lo, got["nchw"][0], "legacy launcher vs v2 NCHW")
    Hh.assert_bit_equal(got["nhwc"][0], got["nchw"][0], "NHWC vs NCHW values")
    Hh.assert_bit_equal(got["nhwc"][1], got["nchw"][1], "NHWC vs NCHW idx_x")
    Hh.assert_bit_equal(got["nhwc"][2], got["nchw"][2], "NHWC vs NCHW idx_y")
    out, ix, iy = got["nchw"]
    if vs_cpu:
        eo, ex, ey = oracle.forward(feats, rois, ph, pw, scale, threads=0)
        Hh.assert_bit_equal(Hh.expand_idx(ix, C), ex, "idx_x vs CPU oracle")
        Hh.assert_bit_equal(Hh.expand_idx(iy, C), ey, "idx_y vs CPU oracle")
        Hh.assert_bit_equal(out, eo, "values vs CPU oracle")
    if vs_ref and oracle.ref_gpu_available():
        ro, rx, ry = _ref_gpu(oracle, feats, rois, ph, pw, scale, device)
        Hh.assert_bit_equal(Hh.expand_idx(ix, C), rx, "idx_x vs reference kernel")
        Hh.assert_bit_equal(Hh.expand_idx(iy, C), ry, "idx_y vs reference kernel")
        Hh.assert_bit_equal(out, ro, "values vs reference kernel")
    return out, ix, iy


def test_library_is_the_cuda_one(cuda):
    from fots.pytorch_b200 import _cabi
    assert "sm_100a" in _cabi.build_info()


def test_cfg0_forward_backward(oracle, cuda):
    """BASELINE.json configs[0]: 1x3x64x64, 4 axis-aligned RoIs (one overhanging), fwd + bwd."""
    feats, rois, ph, pw, scale = WL.cfg0()
    assert (ph, pw) == (8, 32)
    out, ix, iy = _check_forward_all_paths(oracle, feats, rois, ph, pw, scale, cuda)
    rng = np.random.default_rng(5)
    g = rng.standard_normal(out.shape, dtype=np.float32)
    C = feats.shape[1]
    want = oracle.backward(g, rois, Hh.expand_idx(ix, C), Hh.expand_idx(iy, C), feats.shape, scale)
    for cl in (False, True):
        for idx in ((ix, iy), None):
            got = Hh.run_new_backward(g, rois, idx, feats.shape, scale, cuda, channels_last=cl)
            Hh.assert_close_rel(got, want, REL_BWD, "cfg0 backward cl=%s idx=%s" % (cl, idx is not None))
    got = Hh.run_legacy_backward(g, rois, Hh.expand_idx(ix, C), Hh.expand_idx(iy, C), feats.shape, scale, cuda)
    Hh.assert_close_rel(got, want, REL_BWD, "cfg0 legacy backward")
    if oracle.ref_gpu_available():
        import torch
        r = oracle.ref_gpu_backward(Hh.to_cuda(g, cuda), Hh.to_cuda(rois, cuda),
                                    Hh.to_cuda(Hh.expand_idx(ix, C), cuda), Hh.to_cuda(Hh.expand_idx(iy, C), cuda),
                                    feats.shape, scale).cpu().numpy()
        Hh.assert_close_rel(got, r, REL_BWD, "cfg0 backward vs reference kernel")
        Hh.assert_close_rel(want, r, REL_BWD, "cfg0 CPU-oracle backward vs reference kernel")


@pytest.mark.parametrize("channels", [64, 256])
def test_cfg1_forward(oracle, cuda, channels):
    """BASELINE.json configs[1]: 180x320 map (1280x720 / 4), 64 random rotated RoIs, 8x64 output."""
    feats, rois, ph, pw, scale = WL.cfg1(channels)
    _check_forward_all_paths(oracle, feats, rois, ph, pw, scale, cuda, vs_cpu=(channels == 64))


def test_test2_scenario_forward_backward(oracle, cuda):
    """rroi_align/test2.py:22-77 replayed on the committed image: 3 rotated RoIs, PH=44, scale 1,
    loss = pooled.pow(2).sum(); forward against the reference's res*.jpg and both oracles."""
    import torch
    from fots.pytorch_b200 import _RRoiAlign
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ref_test2_jpeg.npz"))
    img = g["timg"].astype(np.float32).transpose(2, 0, 1)[None]
    rois, ph, pw = WL.test2_rois()
    out, ix, iy = _check_forward_all_paths(oracle, img, rois, ph, pw, 1.0, cuda)
    for i in range(3):
        crop = out[i].transpose(1, 2, 0).astype(np.uint8).astype(np.float64)
        mse = ((crop - g["res%d" % i]) ** 2).mean()
        assert 10 * np.log10(255 ** 2 / mse) > 40.0
    # autograd through the module, as test2.py:72-77 does
    x = Hh.to_cuda(img, cuda).requires_grad_(True)
    pooled = _RRoiAlign(ph, pw, 1.0)(x, Hh.to_cuda(rois, cuda).view(-1, 6))
    pooled.pow(2).sum().backward()
    C = img.shape[1]
    want = oracle.backward(2 * out, rois, Hh.expand_idx(ix, C), Hh.expand_idx(iy, C), img.shape, 1.0)
    Hh.assert_close_rel(x.grad.cpu().numpy(), want, REL_BWD, "test2 autograd backward")


@pytest.mark.parametrize("seed,B,C,H,W,N,ph,pw,scale", [
    (1, 2, 5, 45, 80, 96, 8, 64, 0.25),      # odd C -> scalar NHWC path
    (2, 3, 8, 45, 80, 64, 11, 37, 0.25),     # PH=11 as src/ocr_process.py:260; ragged PW
    (3, 1, 3, 72, 128, 40, 32, 100, 1.0),    # CRNN variant: raw image, PH=32 (src/utils.py:430-436)
    (4, 4, 16, 23, 31, 128, 4, 9, 0.125),    # tiny map
    (5, 2, 1, 64, 64, 33, 8, 300, 0.5),      # C=1, PW > 256 (several bin tiles per plane)
    (6, 2, 12, 50, 70, 80, 1, 1, 0.25),      # 1x1 pooling
])
def test_stress_forward_backward(oracle, cuda, seed, B, C, H, W, N, ph, pw, scale):
    feats = WL.features(seed, B, C, H, W)
    rois = WL.stress_rois(seed, N, B, int(W / scale), int(H / scale))
    out, ix, iy = _check_forward_all_paths(oracle, feats, rois, ph, pw, scale, cuda)
    g = np.random.default_rng(seed).standard_normal(out.shape, dtype=np.float32)
    want = oracle.backward(g, rois, Hh.expand_idx(ix, C), Hh.expand_idx(iy, C), feats.shape, scale, threads=0)
    from fots.pytorch_b200 import _cabi
    for cl in (False, True):
        for idx in ((ix, iy), None):
            got = Hh.run_new_backward(g, rois, idx, feats.shape, scale, cuda, channels_last=cl)
            Hh.assert_close_rel(got, want, REL_BWD, "stress backward cl=%s saved_idx=%s" % (cl, idx is not None))
    # the NCHW gather kernel is the automatic choice only for large launches: force it on these shapes too (odd channel
    # counts, PH not a multiple of 8, PW not a multiple of 32, several bin tiles per RoI, unaligned rows -> per-tile fallback)
    for idx in ((ix, iy), None):
        got = Hh.run_new_backward(g, rois, idx, feats.shape, scale, cuda, opts=_cabi.opts(bwd_mode=4))
        Hh.assert_close_rel(got, want, REL_BWD, "stress backward, NCHW gather kernel, saved_idx=%s" % (idx is not None))
    got = Hh.run_legacy_backward(g, rois, Hh.expand_idx(ix, C), Hh.expand_idx(iy, C), feats.shape, scale, cuda)
    Hh.assert_close_rel(got, want, REL_BWD, "stress legacy backward")


def test_degenerate_rois_match_reference_kernel(oracle, cuda):
    """h=0, w=0, negative sizes, NaN/inf parameters: whatever the reference kernel produces (zeros,
    the image centre, NaN) must come out bit-identically -- no special cases of our own."""
    if not oracle.ref_gpu_available():
        pytest.skip("oracle/_ref not built")
    feats = WL.features(9, 1, 4, 40, 60)
    inf, nan = np.inf, np.nan
    rois = np.array([[0, 100, 80, 0, 50, 10], [0, 100, 80, 20, 0, 10], [0, 100, 80, 0, 0, 0],
                     [0, 100, 80, -20, 60, 30], [0, 100, 80, 20, -60, 30], [0, inf, 80, 20, 60, 30],
                     [0, 100, -inf, 20, 60, 30], [0, 100, 80, 20, 60, inf], [0, 100, 80, 20, 60, nan],
                     [0, nan, 80, 20, 60, 0], [0, 100, 80, inf, 60, 0], [0, 100, 80, 20, inf, 0],
                     [0, 1e30, 1e30, 20, 60, 45], [0, 100, 80, 1e-30, 1e-30, 45], [0, 100, 80, 1e20, 1e25, 45]],
                    dtype=np.float32)
    _check_forward_all_paths(oracle, feats, rois, 8, 24, 0.25, cuda, vs_cpu=True, vs_ref=True)


def test_ten_million_bins_index_parity_vs_reference_kernel(oracle, cuda):
    """>= 1e7 random (RoI, bin) pairs: sample centres bit-identical to the reference kernel (GPU vs GPU)."""
    if not oracle.ref_gpu_available():
        pytest.skip("oracle/_ref not built")
    N, ph, pw = 20000, 8, 64       # 10.24 M bins
    feats = WL.features(11, 2, 1, 180, 320)
    rois = np.concatenate([WL.stress_rois(21, N // 2, 2, 1280, 720),
                           np.concatenate([WL.random_rois(100 + i, 64, i % 2) for i in range(N // 2 // 64 + 1)])[:N // 2]])
    out, ix, iy = Hh.run_new_forward(feats, rois, ph, pw, 0.25, cuda)
    ro, rx, ry = _ref_gpu(oracle, feats, rois, ph, pw, 0.25, cuda)
    Hh.assert_bit_equal(ix[:, None], rx, "10M-bin idx_x")
    Hh.assert_bit_equal(iy[:, None], ry, "10M-bin idx_y")
    Hh.assert_bit_equal(out, ro, "10M-bin values")


def test_full_size_properties(oracle, cuda):
    """cfg4 per-GPU size (32 images x 64 RoIs, C=64, 180x320): the oracle itself on all 2 048 RoIs (values and centres
    bit for bit, both layouts), then size-independent properties.
      * zero tail: every element with pw > rpw is exactly 0, and nothing else was left unwritten;
      * layout invariance: NHWC result == NCHW result bitwise;
      * linearity: f(a + b) ~= f(a) + f(b), f(2a) == 2 f(a) exactly (power-of-two scaling commutes);
      * adjointness for RoIs away from the border: <f(x), g> ~= <x, f^T(g)>."""
    import torch
    from fots.pytorch_b200.rroi_align.functions.rroi_align import forward_raw, backward_raw
    from fots.pytorch_b200 import _cabi
    B, C, H, W, ph, pw, scale = 32, 64, 180, 320, 8, 64, 0.25
    gen = torch.Generator(device=cuda).manual_seed(0)
    a = torch.randn(B, C, H, W, device=cuda, generator=gen)
    b = torch.randn(B, C, H, W, device=cuda, generator=gen)
    rois_np = WL.batch_rois(B, 64)
    rois = Hh.to_cuda(rois_np, cuda)
    fa, ix, iy, _ = forward_raw(a, rois, ph, pw, scale)
    fa_cl, _, _, _ = forward_raw(a.contiguous(memory_format=torch.channels_last), rois, ph, pw, scale)
    assert torch.equal(fa, fa_cl.contiguous())
    want, wx, wy = oracle.forward(a.cpu().numpy(), rois_np, ph, pw, scale, threads=0)
    Hh.assert_bit_equal(fa.cpu().numpy(), want, "cfg4-size forward vs oracle")
    Hh.assert_bit_equal(ix.cpu().numpy(), wx[:, 0], "cfg4-size idx_x vs oracle")
    Hh.assert_bit_equal(iy.cpu().numpy(), wy[:, 0], "cfg4-size idx_y vs oracle")
    del want, wx, wy
    rpw = (rois[:, 4] * ph) / rois[:, 3]
    tail = torch.arange(pw, device=cuda)[None, :] > rpw[:, None]            # [N,PW]
    assert (fa.permute(0, 3, 1, 2)[tail] == 0).all()
    assert (ix.permute(0, 2, 1)[tail] == 0).all() and (iy.permute(0, 2, 1)[tail] == 0).all()
    poisoned = torch.full_like(fa, float("nan"))                             # every element is written
    st = _cabi.lib().rroi_b200_forward(a.data_ptr(), rois.data_ptr(), poisoned.data_ptr(), None, None,
                                       rois.shape[0], B, C, H, W, ph, pw, scale, _cabi.LAYOUT_NCHW,
                                       torch.cuda.current_stream(cuda).cuda_stream)
    assert st == 0 and torch.equal(poisoned, fa)
    f2a, _, _, _ = forward_raw(2 * a, rois, ph, pw, scale)
    assert torch.equal(f2a, 2 * fa)
    fb, _, _, _ = forward_raw(b, rois, ph, pw, scale)
    fab, _, _, _ = forward_raw(a + b, rois, ph, pw, scale)
    assert torch.allclose(fab, fa + fb, rtol=1e-4, atol=1e-5)
    # adjointness on interior RoIs (the backward's border rule is stricter than the forward's)
    inner = WL.batch_rois(B, 64)
    inner[:, 1] = np.clip(inner[:, 1], 500, 780)
    inner[:, 2] = np.clip(inner[:, 2], 300, 420)
    inner[:, 3] = np.minimum(inner[:, 3], 40)
    inner[:, 4] = np.minimum(inner[:, 4], 160)
    ri = Hh.to_cuda(inner, cuda)
    fx, jx, jy, lay = forward_raw(a, ri, ph, pw, scale)
    g = torch.randn(fx.shape, device=cuda, generator=gen)
    gt = backward_raw(g, ri, jx, jy, (B, C, H, W), scale, lay)
    lhs = (fx.double() * g.double()).sum().item()
    rhs = (a.double() * gt.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)
    # recomputed centres == saved centres
    gt2 = backward_raw(g, ri, None, None, (B, C, H, W), scale, lay)
    assert torch.allclose(gt, gt2, rtol=1e-4, atol=1e-5)


def test_backward_nchw_variants_agree(oracle, cuda):
    """The three NCHW backward kernels -- row segments + per-granule gather (automatic, bwd_mode 4), one reduction per run of
    equal centres (1) and per tap (3) -- against the CPU oracle and each other, on RoIs whose bin pitch is below one feature
    pixel (many taps per pixel: the case the gather kernel merges in shared memory) and on an unaligned map (W % 4 != 0: the
    gather kernel's per-tile fallback)."""
    from fots.pytorch_b200 import _cabi
    for (H, W) in ((90, 160), (45, 79)):
        feats = WL.features(3, 2, 16, H, W)
        rois = WL.stress_rois(33, 200, 2, W * 4, H * 4)
        rois[:, 3] = np.minimum(rois[:, 3], 6)          # bin pitch < 1 px -> long runs of equal centres
        out, ix, iy = Hh.run_new_forward(feats, rois, 8, 64, 0.25, cuda)
        # positive gradients: with ~100 taps of mixed sign on one pixel the fp32 sum ORDER (unspecified for atomics, in the
        # reference too) alone exceeds 1e-4 of a nearly cancelled sum; the signed case is covered by the stress tests
        g = np.abs(np.random.default_rng(0).standard_normal(out.shape, dtype=np.float32)) + 0.1
        want = oracle.backward(g, rois, Hh.expand_idx(ix, 16), Hh.expand_idx(iy, 16), feats.shape, 0.25, threads=0)
        for mode in (0, 4, 1, 3):
            for idx in ((ix, iy), None):
                got = Hh.run_new_backward(g, rois, idx, feats.shape, 0.25, cuda, opts=_cabi.opts(bwd_mode=mode))
                Hh.assert_close_rel(got, want, REL_BWD, "NCHW backward mode %d, %dx%d, saved centres %s" % (mode, H, W, idx is not None))


@pytest.mark.parametrize("cg", [1, 2, 4, 8, 16])
def test_tuning_variants_bit_exact(cuda, cg):
    """Every per-call kernel variant (rroi_b200_opts.variant / nchw_cg, with and without programmatic dependent launch)
    returns the bits of the default."""
    from fots.pytorch_b200 import _cabi
    feats, rois, ph, pw, scale = WL.cfg1({1: 24, 2: 64, 4: 128, 8: 256, 16: 40}[cg])   # templated and run-time C paths
    base = Hh.run_new_forward(feats, rois, ph, pw, scale, cuda)
    o = _cabi.opts(nchw_cg=cg, variant={1: 1, 2: 2, 4: 3, 8: 4, 16: 5}[cg], pdl=cg % 4 == 0)
    a = Hh.run_new_forward(feats, rois, ph, pw, scale, cuda, opts=o)
    b = Hh.run_new_forward(feats, rois, ph, pw, scale, cuda, channels_last=True, opts=o)
    for x, y in zip(a, base):
        Hh.assert_bit_equal(x, y, "NCHW cg=%d" % cg)
    for x, y in zip(b, base):
        Hh.assert_bit_equal(x, y, "NHWC variant")


@pytest.mark.parametrize("C", [32, 64, 128, 256])
@pytest.mark.parametrize("variant", [0, 1, 5, 6, 11, 12, 13, 14, 15, 16, 17])
def test_nhwc_packed_and_warp_autonomous_variants_bit_exact(oracle, cuda, C, variant):
    """The block-level and the warp-autonomous channels-last kernels, with the RoI rows declared ready (prologue ahead
    of the grid dependency), with a precomputed transform table, and under a concurrency hint: same bits as the oracle,
    centres included; stress RoIs (border, huge angles, out-of-range batch index) on two images."""
    from fots.pytorch_b200 import _cabi
    B, H, W, ph, pw, scale = 2, 45, 80, 8, 64, 0.25
    feats = WL.features(C + variant, B, C, H, W)
    rois = np.concatenate([WL.stress_rois(50 + variant, 40, B, int(W / scale), int(H / scale)),
                           WL.random_rois(9, 9, 1, img_w=int(W / scale), img_h=int(H / scale))], 0)
    want, wx, wy = oracle.forward(feats, rois, ph, pw, scale, threads=0)
    for kw, xf in ((dict(), False), (dict(rois_ready=True), False), (dict(), True), (dict(rois_ready=True, concurrency=8), True),
                   (dict(pdl=False), False)):
        got, ix, iy = Hh.run_new_forward(feats, rois, ph, pw, scale, cuda, channels_last=True,
                                         opts=_cabi.opts(variant=variant, **kw), with_xform=xf)
        Hh.assert_bit_equal(got, want, "variant %d %r xform=%s" % (variant, kw, xf))
        Hh.assert_bit_equal(ix, wx[:, 0], "idx_x")
        Hh.assert_bit_equal(iy, wy[:, 0], "idx_y")
    # ragged: bins not a multiple of any segment size, N not a multiple of the warps per CTA
    r3 = rois[:3]
    want3, wx3, _ = oracle.forward(feats, r3, 5, 13, scale, threads=0)
    got3, ix3, _ = Hh.run_new_forward(feats, r3, 5, 13, scale, cuda, channels_last=True, opts=_cabi.opts(variant=variant))
    Hh.assert_bit_equal(got3, want3, "ragged variant %d" % variant)
    Hh.assert_bit_equal(ix3, wx3[:, 0], "ragged idx_x")


def test_unknown_variant_and_bad_opts_are_rejected(cuda):
    import torch
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.rroi_align.functions.rroi_align import forward_raw
    f = torch.zeros(1, 64, 8, 8, device=cuda).contiguous(memory_format=torch.channels_last)
    r = torch.tensor([[0, 10, 10, 8, 16, 0]], dtype=torch.float32, device=cuda)
    with pytest.raises(_cabi.RRoiAlignError):
        forward_raw(f, r, 8, 16, 1.0, opts=_cabi.opts(variant=9))
    forward_raw(f, r, 8, 16, 1.0, opts=_cabi.opts(variant=12))
    torch.cuda.synchronize()


def test_module_api_and_function_attributes(oracle, cuda):
    """Drop-in surface: rroi_align.modules.rroi_align._RRoiAlign / rroi_align.functions.rroi_align.RRoiAlignFunction."""
    import torch
    from rroi_align.modules.rroi_align import _RRoiAlign
    from rroi_align.functions.rroi_align import RRoiAlignFunction
    feats, rois, ph, pw, scale = WL.cfg0()
    f = Hh.to_cuda(feats, cuda).requires_grad_(True)
    r = Hh.to_cuda(rois, cuda)
    eo, ex, ey = oracle.forward(feats, rois, ph, pw, scale)
    m = _RRoiAlign(ph, pw, scale)
    assert (m.pooled_height, m.pooled_width, m.spatial_scale) == (8, 32, 1.0)
    y = m(f, r)
    assert y.shape == (4, 3, 8, 32) and y.requires_grad
    Hh.assert_bit_equal(y.detach().cpu().numpy(), eo, "module forward")
    fn = RRoiAlignFunction(ph, pw, scale)
    y2 = fn(f, r)
    Hh.assert_bit_equal(y2.detach().cpu().numpy(), eo, "function forward")
    assert tuple(fn.feature_size) == (1, 3, 64, 64) and fn.rois is r
    Hh.assert_bit_equal(fn.idx_x.cpu().numpy(), ex, "fn.idx_x")     # [N,C,PH,PW] like the reference's ctx.idx_x
    Hh.assert_bit_equal(fn.idx_y.cpu().numpy(), ey, "fn.idx_y")
    g = torch.randn_like(y2)
    y2.backward(g)
    want = oracle.backward(g.cpu().numpy(), rois, ex, ey, feats.shape, scale)
    Hh.assert_close_rel(f.grad.cpu().numpy(), want, REL_BWD, "autograd backward")
    # legacy direct forward()/backward() pair returns (grad_input, None) like functions/rroi_align.py:40
    fn2 = RRoiAlignFunction(ph, pw, scale)
    y3 = fn2.forward(f.detach(), r)
    gi, none = fn2.backward(g)
    assert none is None
    Hh.assert_bit_equal(y3.cpu().numpy(), eo, "legacy forward()")
    Hh.assert_close_rel(gi.cpu().numpy(), want, REL_BWD, "legacy backward()")
    # channels_last features -> channels_last result, same logical values
    fcl = f.detach().contiguous(memory_format=torch.channels_last)
    ycl = m(fcl, r)
    assert ycl.is_contiguous(memory_format=torch.channels_last)
    Hh.assert_bit_equal(ycl.cpu().numpy(), eo, "channels_last forward")
    # empty RoI set
    y0 = m(f, r[:0])
    assert y0.shape == (0, 3, 8, 32)
    y0.sum().backward()


def test_module_launch_hints_do_not_change_results(oracle, cuda):
    """_RRoiAlign(ph, pw, scale, concurrency=, rois_ready=): the optional launch hints pick kernels, never values (forward
    bit-exact, backward to 1e-4), in both layouts."""
    import torch
    from fots.pytorch_b200 import _RRoiAlign
    feats, rois, ph, pw, scale = WL.cfg1(64)
    want, _, _ = oracle.forward(feats, rois, ph, pw, scale, threads=0)
    r = Hh.to_cuda(rois, cuda)
    for cl in (False, True):
        f = Hh.to_cuda(feats, cuda, cl).requires_grad_(True)
        base = None
        for kw in (dict(), dict(concurrency=8), dict(rois_ready=True), dict(concurrency=8, rois_ready=True)):
            f.grad = None
            y = _RRoiAlign(ph, pw, scale, **kw)(f, r)
            Hh.assert_bit_equal(y.detach().cpu().numpy(), want, "module forward %r cl=%s" % (kw, cl))
            y.backward(torch.ones_like(y))
            g = f.grad.detach().cpu().numpy()
            if base is None:
                base = g
            else:
                Hh.assert_close_rel(g, base, REL_BWD, "module backward %r" % (kw,))


def test_error_behaviour(cuda):
    import torch
    from fots.pytorch_b200 import _RRoiAlign, _cabi
    m = _RRoiAlign(8, 32, 1.0)
    f = torch.zeros(1, 3, 16, 16, device=cuda)
    with pytest.raises(ValueError):
        m(f, torch.zeros(4, 5, device=cuda))                 # reference: wrapper returns 0, silently
    with pytest.raises(RuntimeError):
        m(f.cpu(), torch.zeros(4, 6))                        # reference: NameError on the CPU branch
    with pytest.raises(TypeError):
        m(f.double(), torch.zeros(4, 6, device=cuda))
    # out-of-range batch index: zeros instead of the reference's out-of-bounds read
    r = torch.tensor([[3, 8, 8, 4, 8, 0], [-1, 8, 8, 4, 8, 0]], device=cuda, dtype=torch.float32)
    y = m(f + 1, r)
    assert (y == 0).all()
    assert _cabi.lib().rroi_b200_forward(None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1.0, 0, None) == _cabi.ERR_INVALID_ARG
    assert _cabi.lib().rroi_b200_forward(f.data_ptr(), r.data_ptr(), y.data_ptr(), None, None, 2, 1, 3, 16, 16, 8, 32, 1.0, 7, None) == _cabi.ERR_INVALID_ARG


@pytest.mark.parametrize("C,H,W,scale,rois", [
    (64, 180, 320, 0.25, None),                                            # cfg1: boxes of 8..32 pixels
    (7, 90, 160, 0.25, None),                                              # C smaller than every box's channel group
    (16, 64, 64, 1.0, [[0, 32, 32, 60, 200, 33], [0, 10, 50, 40, 64, -70]]),   # footprints > 32 px: gather fallback per CTA
    (5, 33, 50, 0.5, None),                                                # W % 4 != 0: no tensor map, gather kernel
])
def test_nchw_tma_staged_forward_equals_gather_kernel_and_oracle(oracle, cuda, C, H, W, scale, rois):
    """The opt-in NCHW forward that stages each patch's footprint with TMA box loads (rroi_fwd_nchw_tma_kernel,
    rroi_b200_opts.nchw_tma = 1, and 5 = largest box forced) must agree bit for bit with the gather kernel and with
    the oracle, centres included."""
    from fots.pytorch_b200 import _cabi
    B = 2
    feats = WL.features(C, B, C, H, W)
    if rois is None:
        r = np.concatenate([WL.random_rois(3 + i, 24, i, img_w=int(W / scale), img_h=int(H / scale)) for i in range(B)], 0)
        r[0, 1:3] = (2.0, 3.0)
    else:
        r = np.array(rois, np.float32)
    ph, pw = 8, 64
    want, wx, wy = oracle.forward(feats, r, ph, pw, scale, threads=0)
    outs = []
    for mode in (0, 1, 5):
        outs.append(Hh.run_new_forward(feats, r, ph, pw, scale, cuda, channels_last=False, opts=_cabi.opts(nchw_tma=mode)))
    for got, ix, iy in outs:
        Hh.assert_bit_equal(got, want, "NCHW forward")
        Hh.assert_bit_equal(ix, wx[:, 0], "idx_x")
        Hh.assert_bit_equal(iy, wy[:, 0], "idx_y")


@pytest.mark.parametrize("C,H,W,scale,ph,pw,rois", [
    (64, 180, 320, 0.25, 8, 64, None),                                         # cfg1
    (7, 90, 160, 0.25, 8, 64, None),                                           # fewer channels than one stage holds
    (3, 276, 500, 1.0, 44, 349, "test2"),                                      # the reference's test2.py scenario: PH = 44, ragged blocks
    (16, 64, 64, 1.0, 8, 64, [[0, 32, 32, 60, 200, 33], [0, 10, 50, 40, 64, -70], [0, 30, 30, 62, 500, 90]]),  # footprints > 96 rows: gather fallback
    (5, 33, 50, 0.5, 5, 13, None),                                             # W % 4 != 0: gather fallback; ragged block
    (40, 45, 80, 0.25, 8, 64, "stress"),
])
@pytest.mark.parametrize("variant", [2, 3, 4])
def test_nchw_row_segment_staged_forward_equals_oracle(oracle, cuda, C, H, W, scale, ph, pw, rois, variant):
    """The NCHW forward that stages exactly the row segments a tile touches (rroi_fwd_nchw_rows_kernel, opts.variant = 2)
    agrees bit for bit with the oracle and the gather kernel, centres included, also through the legacy [N,C,PH,PW] centres."""
    from fots.pytorch_b200 import _cabi
    B = 2
    feats = WL.features(C + H, B, C, H, W)
    if rois is None:
        r = np.concatenate([WL.random_rois(3 + i, 24, i, img_w=int(W / scale), img_h=int(H / scale)) for i in range(B)], 0)
        r[0, 1:3] = (2.0, 3.0)
    elif rois == "test2":
        r, _, _ = WL.test2_rois()
    elif rois == "stress":
        r = WL.stress_rois(91, 150, B, int(W / scale), int(H / scale))
    else:
        r = np.array(rois, np.float32)
    want, wx, wy = oracle.forward(feats, r, ph, pw, scale, threads=0)
    got, ix, iy = Hh.run_new_forward(feats, r, ph, pw, scale, cuda, channels_last=False, opts=_cabi.opts(variant=variant))
    Hh.assert_bit_equal(got, want, "row-staged NCHW forward")
    Hh.assert_bit_equal(ix, wx[:, 0], "idx_x")
    Hh.assert_bit_equal(iy, wy[:, 0], "idx_y")
    got2, _, _ = Hh.run_new_forward(feats, r, ph, pw, scale, cuda, channels_last=False, opts=_cabi.opts(variant=variant, rois_ready=True, pdl=True))
    Hh.assert_bit_equal(got2, want, "row-staged NCHW forward, RoIs ready")


@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
@pytest.mark.parametrize("order", ["grouped", "grouped_with_empty_images", "shuffled", "crowded_image"])
def test_backward_chunked_zero_fill_and_scatter(oracle, cuda, order, layout):
    """zero_fill on a map larger than 96 MB: the gradient map is cleared in chunks of a few images on a side stream and
    each chunk's RoIs are scattered as soon as its chunk is clear (launch_bwd_zero_scatter, rroi_bwd.cu), for any order
    of the RoI rows.  Every element of the map is defined (the buffer starts as NaN) and equals the oracle's to 1e-4,
    for the automatic chunk size, one image per chunk, three, and the unchunked memset + scatter."""
    import torch
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.rroi_align.functions.rroi_align import backward_raw
    B, C, H, W, ph, pw, scale = 8, 64, 180, 320, 8, 64, 0.25          # 118 MB gradient map
    per = {"grouped": [5, 3, 4, 2, 6, 1, 3, 4], "grouped_with_empty_images": [6, 0, 0, 5, 0, 7, 0, 0],
           "shuffled": [3, 3, 3, 3, 3, 3, 3, 3], "crowded_image": [300, 2, 0, 1, 0, 0, 0, 1]}[order]
    rois = np.concatenate([WL.random_rois(11 + i, n, i) for i, n in enumerate(per) if n > 0], 0)
    if order == "shuffled":
        rois = rois[np.random.default_rng(0).permutation(len(rois))]
        rois[0, 0] = B + 2                                                # and one row whose image does not exist
    N = rois.shape[0]
    rng = np.random.default_rng(5)
    top = rng.standard_normal((N, C, ph, pw), dtype=np.float32)
    feats = np.zeros((B, C, H, W), np.float32)
    ok = rois[:, 0] < B
    _, ix, iy = oracle.forward(feats, rois[ok], ph, pw, scale, threads=0)
    want = oracle.backward(top[ok], rois[ok], ix, iy, (B, C, H, W), scale, threads=0)
    fmt = torch.channels_last if layout == "nhwc" else torch.contiguous_format
    lay = _cabi.LAYOUT_NHWC if layout == "nhwc" else _cabi.LAYOUT_NCHW
    g = torch.from_numpy(top).to(cuda).contiguous(memory_format=fmt)
    r = torch.from_numpy(rois).to(cuda)
    for chunk in (0, 1, 3, -1):
        # poison the allocator's next block so that an element the kernel forgets to define shows up as NaN
        poison = torch.full((B, C, H, W), float("nan"), device=cuda).contiguous(memory_format=fmt)
        del poison
        got = backward_raw(g, r, None, None, (B, C, H, W), scale, lay, opts=_cabi.opts(zero_chunk_images=chunk))
        torch.cuda.synchronize()
        assert bool(torch.isfinite(got).all())
        Hh.assert_close_rel(got.cpu().numpy(), want, 1e-4, "chunked zero-fill backward (%s, %s, chunk %d)" % (order, layout, chunk))
        del got
