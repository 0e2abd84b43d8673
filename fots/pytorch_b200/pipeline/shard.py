"""Image sharding across ranks and the ONE collective of the end-to-end path.

Images are independent units: every RoI row carries its image index and reads only that image's planes, so
rank r owns a contiguous slice of the batch and nothing is exchanged inside RoIRotate.  After recognition each
rank holds fixed-size per-image records; a single all_gather makes the full result visible everywhere."""
import torch
import torch.distributed as dist


def shard_range(batch, world, rank):
    """Contiguous slice [lo, hi) of `batch` images owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_records(quads, ids, lengths):
    """quads [b, R, 9] fp32, ids [b, R, T] int32, lengths [b, R] int32 -> int32 [b, R, 9 + T + 1] (bit-cast, no copy of meaning)."""
    b, R, _ = quads.shape
    return torch.cat((quads.contiguous().view(torch.int32), ids, lengths.view(b, R, 1)), dim=2).contiguous()


def unpack_records(rec, T):
    quads = rec[..., :9].contiguous().view(torch.float32)
    return quads, rec[..., 9:9 + T], rec[..., 9 + T]


def all_gather_records(rec, batch, group=None):
    """One collective: gather every rank's [b_r, R, F] int32 records into [batch, R, F] (rank order = image order).
    Ranks may own different numbers of images (batch % world != 0): records are padded to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rec
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(batch, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    pad = rec
    if rec.size(0) < most:
        pad = torch.cat((rec, rec.new_zeros((most - rec.size(0),) + tuple(rec.shape[1:]))), 0)
    out = rec.new_empty((world * most,) + tuple(rec.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(hi - lo == most for lo, hi in sizes):
        return out
    return torch.cat([out[r * most: r * most + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)
