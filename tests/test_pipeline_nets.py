"""Feeder / consumer networks: checkpoint-key compatibility with the
reference (fixture tests/golden/ref_model_keys.json), and -- when /root/reference is importable, i.e. in the
build container -- identical outputs to the reference modules with shared weights (fp32 CPU, 1e-5)."""
import importlib.util
import json
import os

import pytest
import torch

from fots.pytorch_b200.pipeline.nets import CRNN, FOTSNet

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF_MODELS = "/root/reference/tools/models.py"
TOL = 1e-5   # fp32 CPU, same op sequence: differences are summation-order noise only


def _ref():
    spec = importlib.util.spec_from_file_location("ref_models", REF_MODELS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,build", [
    ("ModelResNetSep2_att_89", lambda: FOTSNet(attention=True, nclass=89)),
    ("ModelResNetSep2_noatt_89", lambda: FOTSNet(attention=False, nclass=89)),
    ("CRNN_89", lambda: CRNN(nclass=89)),
])
def test_state_dict_keys_match_reference(name, build):
    want = json.load(open(os.path.join(GOLDEN, "ref_model_keys.json")))[name]
    got = {k: list(v.shape) for k, v in build().state_dict().items()}
    assert got == want


needs_ref = pytest.mark.skipif(not os.path.exists(REF_MODELS), reason="reference checkout not present")


@needs_ref
@pytest.mark.parametrize("attention", [True, False])
def test_fotsnet_matches_reference_outputs(attention):
    torch.manual_seed(0)
    ref = _ref().ModelResNetSep2(attention=attention, nclass=89).eval()
    mine = FOTSNet(attention=attention, nclass=89).eval()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 96, 160)
    with torch.no_grad():
        a, b = ref(x), mine(x)
    for group_a, group_b in zip(a, b):
        for ta, tb in zip(group_a, group_b):
            assert ta.shape == tb.shape
            assert torch.allclose(ta, tb, rtol=TOL, atol=TOL), float((ta - tb).abs().max())
    pooled = torch.randn(3, 64, 8, 64)
    with torch.no_grad():
        assert torch.allclose(ref.forward_ocr(pooled), mine.forward_ocr(pooled), rtol=TOL, atol=TOL)
        p11 = torch.randn(2, 64, 11, 96)                                          # PH=11 (src/ocr_process.py:260)
        assert torch.allclose(ref.forward_ocr(p11), mine.forward_ocr(p11), rtol=TOL, atol=TOL)
        assert torch.allclose(ref.forward_features(x), mine.forward_features(x), rtol=TOL, atol=TOL)


@needs_ref
def test_crnn_matches_reference_outputs():
    torch.manual_seed(1)
    ref = _ref().CRNN(nclass=89).eval()
    mine = CRNN(nclass=89).eval()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(3, 3, 32, 120)
    with torch.no_grad():
        a, b = ref(x), mine(x)
    assert a.shape == b.shape == (31, 3, 89)
    assert torch.allclose(a, b, rtol=TOL, atol=TOL)


def test_shapes_and_channels_last():
    net = FOTSNet(attention=True, nclass=89).eval()
    x = torch.randn(1, 3, 64, 96).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        seg, rbox, ang, feats = net(x)
        assert seg[0].shape == (1, 1, 16, 24) and seg[1].shape == (1, 1, 8, 12)
        assert rbox[0].shape == (1, 4, 16, 24) and ang[0].shape == (1, 2, 16, 24)
        assert feats[0].shape == (1, 256, 16, 24) and feats[1].shape == (1, 64, 16, 24)
        assert torch.allclose((ang[0] ** 2).sum(1), torch.ones(1, 16, 24), atol=1e-5)     # unit (sin, cos)
        for ph in (8, 9, 10, 11):                                                           # SURVEY #8: PH in 8..11
            assert net.forward_ocr(torch.randn(2, 64, ph, 40)).shape == (2, 89, 40)
        lp = net.forward_ocr(torch.randn(2, 64, 8, 40))
        assert torch.allclose(lp.exp().sum(1), torch.ones(2, 40), atol=1e-4)
    with pytest.raises(ValueError):
        CRNN(nclass=10)(torch.randn(1, 3, 48, 64))


def test_folded_head_weights_reproduce_the_last_merge_level():
    """conv.pack_merged_heads (host algebra behind fots_b200_heads_merged_nhwc_bf16): with x = upconv2_pw(d) + feature1(s) * gate
    (tools/models.py:430-438) the head logits Wh x + bh equal (Wh Wpw) d + gate * (Wh Wf1) s + bh.  Checked on the CPU in fp32
    against the modules themselves; the packed weights are bf16, hence the tolerance."""
    from fots.pytorch_b200.pipeline import conv as TC
    torch.manual_seed(5)
    net = FOTSNet(attention=True, nclass=20).eval()
    d, s = torch.randn(2, 256, 6, 10), torch.randn(2, 64, 6, 10)
    gate = torch.rand(2, 1, 6, 10)
    with torch.no_grad():
        x = net.upconv2[1](d) + net.feature1(s) * gate
        want = torch.cat((net.act(x), torch.zeros_like(net.act(x)), net.rbox(x), net.angle(x)), 1)       # the kernel's 8-column block
        w1, w2, b = TC.pack_merged_heads(net.act, net.rbox, net.angle, net.upconv2[1], net.feature1)
        got = (torch.einsum("jc,bchw->bjhw", w1.float(), d) + gate * torch.einsum("jc,bchw->bjhw", w2.float(), s)
               + b.view(1, 8, 1, 1))
    assert w1.shape == (8, 256) and w2.shape == (8, 64) and b.shape == (8,)
    live = [0, 2, 3, 4, 5, 6, 7]
    err = (got[:, live] - want[:, live]).abs().max()
    assert float(err) <= 2.0 ** -7 * float(want.abs().max()) + 1e-3, float(err)
    assert float(got[:, 1].abs().max()) == 0.0                                                           # the padding column stays empty
    # a convolution with a bias cannot be folded behind the gate: the packer declines
    biased = torch.nn.Conv2d(64, 256, 1, bias=True)
    assert TC.pack_merged_heads(net.act, net.rbox, net.angle, net.upconv2[1], biased) is None


def test_gather_head_weights_reproduce_upsample_depthwise_and_pointwise():
    """conv.pack_gather_heads (host algebra behind fots_b200_heads_gather_nhwc_bf16): the head logits of
    upconv2(upsample(f2)) -- tools/models.py:436-438, F.interpolate(bilinear, align_corners=True) -> depthwise 3x3 -> 1x1 --
    equal  sum_tap shift_tap(upsample(T_tap)),  T = conv1x1(f2, A),  A[tap * 8 + o] = (Wh Wpw)[o, :] * w_dw[:, tap]:
    the upsampling commutes with the channel mix and the depthwise taps become shifts with zero padding.  CPU, fp32, against
    the modules themselves, on an odd size; the packed weights are bf16, hence the tolerance."""
    import torch.nn.functional as F
    from fots.pytorch_b200.pipeline import conv as TC
    torch.manual_seed(6)
    net = FOTSNet(attention=True, nclass=20).eval()
    B, h, w, H, W = 2, 5, 7, 9, 13
    f2 = torch.randn(B, 256, h, w)
    with torch.no_grad():
        x = net.upconv2(F.interpolate(f2, size=(H, W), mode="bilinear", align_corners=True))
        want = torch.cat((net.act(x) - net.act.bias.view(1, -1, 1, 1), torch.zeros(B, 1, H, W), net.rbox(x) - net.rbox.bias.view(1, -1, 1, 1),
                          net.angle(x) - net.angle.bias.view(1, -1, 1, 1)), 1)                           # 8-column block without the bias
        a = TC.pack_gather_heads(net.act, net.rbox, net.angle, net.upconv2[1], net.upconv2[0])
        assert a.shape == (128, 256, 1, 1) and float(a[72:].abs().max()) == 0.0
        T = F.conv2d(f2, a.float())[:, :72]                                                              # [B, 72, h, w]
        up = F.pad(F.interpolate(T, size=(H, W), mode="bilinear", align_corners=True), (1, 1, 1, 1)).reshape(B, 9, 8, H + 2, W + 2)
        got = sum(up[:, r * 3 + c, :, r:r + H, c:c + W] for r in range(3) for c in range(3))
    live = [0, 2, 3, 4, 5, 6, 7]
    err = (got[:, live] - want[:, live]).abs().max()
    assert float(err) <= 2.0 ** -7 * float(want.abs().max()) + 1e-3, float(err)
    assert float(got[:, 1].abs().max()) == 0.0
    # a depthwise convolution with a bias is not folded
    biased = torch.nn.Conv2d(256, 256, 3, 1, 1, groups=256, bias=True)
    assert TC.pack_gather_heads(net.act, net.rbox, net.angle, net.upconv2[1], biased) is None
