"""Host-side logic of the end-to-end path: sharding, record packing, the single collective (gloo, world 2),
RoI-row construction rules."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fots.pytorch_b200.pipeline import shard
from fots.pytorch_b200.pipeline.infer import planted_quads
from fots.pytorch_b200.pipeline.rois import pooled_width_for


def ref_roi_row(q, b=0):
    """tools/ocr_utils.py:133-145 restated on numpy float32 scalars + Python floats, exactly as the reference mixes them."""
    boxr = np.asarray(q[:8], np.float32).reshape(-1, 2)
    center = (boxr[0, :] + boxr[1, :] + boxr[2, :] + boxr[3, :]) / 4
    dw = boxr[2, :] - boxr[1, :]
    dh = boxr[1, :] - boxr[0, :]
    w = math.sqrt(dw[0] * dw[0] + dw[1] * dw[1])
    h = math.sqrt(dh[0] * dh[0] + dh[1] * dh[1])
    angle = math.atan2((boxr[2][1] - boxr[1][1]), boxr[2][0] - boxr[1][0])
    angle = -angle / 3.1415926535 * 180
    return np.asarray([b, int(center[0]), int(center[1]), h, w, angle], np.float32)


def test_shard_range_partitions_the_batch():
    for batch in (1, 5, 8, 32, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard.shard_range(256, 8, 3) == (96, 128)          # cfg4: 32 images per GPU


def test_pack_unpack_round_trip():
    g = torch.Generator().manual_seed(0)
    quads = torch.randn(3, 5, 9, generator=g)
    ids = torch.randint(0, 89, (3, 5, 16), generator=g, dtype=torch.int32)
    lens = torch.randint(0, 16, (3, 5), generator=g, dtype=torch.int32)
    rec = shard.pack_records(quads, ids, lens)
    assert rec.dtype == torch.int32 and rec.shape == (3, 5, 9 + 16 + 1)
    q2, i2, l2 = shard.unpack_records(rec, 16)
    assert torch.equal(q2, quads) and torch.equal(i2, ids) and torch.equal(l2, lens)


def test_planted_quads_recover_generator_parameters():
    """The planted boxes must come back out of the reference's RoI rule as the (h, w, -angle) they were drawn with."""
    q = planted_quads(2, 64)
    assert q.shape == (2, 64, 9)
    for b in range(2):
        rng = np.random.default_rng(b)
        h = rng.uniform(16, 64, 64)
        w = h * rng.uniform(1, 8, 64)
        cx, cy = rng.uniform(0, 1280, 64), rng.uniform(0, 720, 64)
        ang = rng.uniform(-90, 90, 64)
        for i in range(64):
            row = ref_roi_row(q[b, i], b)
            assert abs(row[3] - h[i]) < 1e-2 and abs(row[4] - w[i]) < 1e-2
            assert abs(row[5] - (-ang[i])) < 1e-2
            assert abs(row[1] - int(cx[i])) <= 1 and abs(row[2] - int(cy[i])) <= 1


def test_pooled_width_rules():
    assert pooled_width_for(20.0, 200.0, 11, "infer") == max(2, (int(200 * 11 / 20) + 11) // 32) * 32 == 96
    assert pooled_width_for(0.5, 3.0, 11, "infer") == 64            # h clamps to 1, floor of 2 buckets
    assert pooled_width_for([10, 20], [100, 60], 11, "train") == 110  # ceil(11 * max(w/h)) (src/ocr_process.py:260-263)
    assert pooled_width_for([19.464249], [154.02922], 44, "train") == 349   # rroi_align/test2.py:64-67 fixture size


def _worker(rank, world, batch, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        R, T = 4, 6
        full = torch.arange(batch * R * (9 + T + 1), dtype=torch.int32).view(batch, R, 9 + T + 1)
        lo, hi = shard.shard_range(batch, world, rank)
        got = shard.all_gather_records(full[lo:hi].clone(), batch)
        q.put((rank, bool(torch.equal(got, full)), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5])
def test_all_gather_records_world2_gloo(batch):
    """N>1 path on CPU: two gloo ranks, even and uneven shards, one collective, image order preserved."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + batch) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, batch, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(ok for _, ok, _ in res)
