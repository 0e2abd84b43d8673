"""The host merge of the detector post-processing (fots_b200_merge_candidates_host) against the
reference's OWN merge: nms/nms.h + its vendored Clipper, compiled unmodified into oracle/_ref/libref_nms.so by
oracle/Makefile behind oracle/nms_ref_shim.cpp.  CPU-only tests (the merge is host code in the reference too); the
candidates come from the CPU restatement of the decode.  A GPU variant (decode on the device) is in
tests/test_gpu_pipeline.py."""
import numpy as np
import pytest
import torch

import workloads as WL


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_nms_available():
        pytest.skip("oracle/_ref/libref_nms.so not built (needs /root/reference at build time)")
    return oracle


def _both(oracle, cand, n, w, h, t1=0.4, t2=0.2, max_boxes=1024):
    from fots.pytorch_b200.pipeline import detect
    want = oracle.ref_merge_candidates(cand[:n], w, h, t1, t2, 8192)
    want[:, :8] /= 10000.0
    got = detect.merge_candidates(torch.tensor([n], dtype=torch.int32), torch.from_numpy(cand[None]), w, h, t1, t2, max_boxes)[0]
    return got, want


SCENES = {
    "separate": [(40, 20, 50, 10, 0.2), (100, 50, 60, 14, -0.5), (60, 70, 30, 8, 1.2), (130, 20, 40, 12, 0.0)],
    "overlapping_copies": [(40, 20, 50, 10, 0.2), (42, 24, 50, 10, 0.2), (100, 50, 60, 14, -0.5), (101, 50, 60, 14, -0.45)],
    "touching_borders": [(10, 5, 30, 8, 0.0), (150, 84, 28, 8, 0.1), (155, 40, 20, 40, 1.5)],
    "steep": [(80, 45, 70, 12, 1.4), (30, 45, 40, 10, -1.3)],
}


@pytest.mark.parametrize("scene", sorted(SCENES))
@pytest.mark.parametrize("noise", [0.0, 0.15, 1.0])
def test_planted_boxes_merge_like_the_reference(ref, scene, noise):
    h, w = 90, 160
    seg, rbox, ang = WL.planted_detection_maps(h, w, SCENES[scene], seed=len(scene), noise=noise)
    n, cand = ref.decode_candidates(seg, rbox, ang, 0.5, 16384)
    assert 0 < n <= 16384
    got, want = _both(ref, cand, n, w, h)
    assert got.shape == want.shape and got.shape[0] >= 1
    assert np.array_equal(got, want), np.abs(got - want).max()


@pytest.mark.parametrize("seed,thr", [(0, 0.97), (1, 0.95), (2, 0.9)])
def test_random_maps_many_fragments(ref, seed, thr):
    """Unstructured maps: every positive pixel is its own random quadrangle, so the merge chains, the duplicate appends
    and the second stage all see thousands of unrelated polygons; the two implementations may differ only where an
    IoU lands within rounding of a threshold (Clipper rounds intersection vertices to integers)."""
    rng = np.random.default_rng(seed)
    h, w = 48, 64
    seg = rng.random((h, w), dtype=np.float32)
    rbox = (rng.random((4, h, w), dtype=np.float32) * 24).astype(np.float32)
    a = rng.uniform(-np.pi, np.pi, (h, w)).astype(np.float32)
    ang = np.stack([np.sin(a), np.cos(a)]).astype(np.float32)
    n, cand = ref.decode_candidates(seg, rbox, ang, thr, 8192)
    got, want = _both(ref, cand, n, w, h, max_boxes=8192)
    assert got.shape == want.shape
    same = np.all(got == want, axis=1).mean()
    assert same > 0.98, same


def test_thresholds_empty_input_and_truncation(ref):
    from fots.pytorch_b200 import _cabi
    from fots.pytorch_b200.pipeline import detect
    h, w = 90, 160
    seg, rbox, ang = WL.planted_detection_maps(h, w, SCENES["overlapping_copies"], seed=3)
    n, cand = ref.decode_candidates(seg, rbox, ang, 0.5, 16384)
    for t1, t2 in ((0.4, 0.2), (0.9, 0.8), (0.05, 0.01), (0.4, 0.99)):
        got, want = _both(ref, cand, n, w, h, t1, t2)
        assert np.array_equal(got, want), (t1, t2)
    # nothing above the threshold -> no boxes
    got = detect.merge_candidates(torch.zeros(2, dtype=torch.int32), torch.zeros(2, 8, 16, dtype=torch.int32), w, h)
    assert [g.shape for g in got] == [(0, 9), (0, 9)]
    # more polygons than the caller's buffer: the first max_boxes, in the reference's order
    got_all, want = _both(ref, cand, n, w, h, 0.9, 0.8)
    assert got_all.shape[0] > 2
    got2, _ = _both(ref, cand, n, w, h, 0.9, 0.8, max_boxes=2)
    assert np.array_equal(got2, want[:2])
    # a candidate whose pixel lies outside the map is an argument error, not a wild write
    bad = cand[:1].copy()
    bad[0, 13] = w
    with pytest.raises(_cabi.RRoiAlignError):
        detect.merge_candidates(torch.tensor([1], dtype=torch.int32), torch.from_numpy(bad[None]), w, h)
