"""Multi-GPU plumbing: one process per GPU over torch.distributed.

The path shards by (user,item) segment: contiguous CSR segment ranges per rank, balanced by path count, never splitting
a segment -> scoring needs no data-path collective; the training step needs one all-reduce(sum) over the flat gradient
buffer (and a scalar for the loss mean).  The reference is single-device (model/OneModel.lua:177); this is new design.
"""
import numpy as np


def partition_segments(seg, world):
    """Split CSR offsets seg[S+1] into `world` contiguous segment ranges with ~equal path counts.
    Returns a list of (seg_begin, seg_end) per rank (possibly empty ranges for tiny inputs)."""
    seg = np.asarray(seg, dtype=np.int64)
    S, N = len(seg) - 1, int(seg[-1])
    bounds = [0]
    for r in range(1, world):
        target = N * r / world
        k = int(np.searchsorted(seg, target, side="left"))
        k = min(max(k, bounds[-1]), S)
        bounds.append(k)
    bounds.append(S)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard(ids, seg, labels, rank, world):
    """The rank's slice of a CSR batch: (ids_local, seg_local (rebased to 0), labels_local, (seg_begin, seg_end))."""
    a, b = partition_segments(seg, world)[rank]
    seg = np.asarray(seg, dtype=np.int64)
    p0, p1 = int(seg[a]), int(seg[b])
    seg_local = (seg[a:b + 1] - p0).astype(np.int32)
    lab = None if labels is None else np.asarray(labels)[a:b]
    return np.asarray(ids)[p0:p1], seg_local, lab, (a, b)


def shard_paths(ids, seg, rank, world):
    """Path-sharded view of a CSR batch (kprn_set_option "path_sharded"): path p goes to rank p % world, every rank keeps all S
    segments (its share of each, possibly empty).  Returns (ids_local, seg_local)."""
    ids = np.asarray(ids)
    seg = np.asarray(seg, dtype=np.int64)
    mine = np.arange(ids.shape[0]) % world == rank
    counts = np.add.reduceat(mine.astype(np.int64), seg[:-1]) if len(seg) > 1 else np.zeros(0, np.int64)
    counts[seg[1:] == seg[:-1]] = 0                     # reduceat on an empty segment returns the next element
    seg_local = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return ids[mine], seg_local


def allreduce_lse(seg_logits_local, seg_local, inv_gamma=1.0, group=None):
    """Host mirror of pool_allreduce_lse (csrc/comm.cu): combines per-rank pooled logits [S, L] of a path-sharded batch as
    log-sum-exp partials -- all-reduce(max) of m_r = ig * pooled_r (-inf where the rank holds no path of the segment), then
    all-reduce(sum) of exp(m_r - M)."""
    import torch
    import torch.distributed as dist
    x = np.asarray(seg_logits_local, dtype=np.float64) * inv_gamma
    seg_local = np.asarray(seg_local)
    empty = seg_local[1:] == seg_local[:-1]
    x = np.where(empty[:, None], -np.inf, x)
    M = torch.from_numpy(x.copy())
    dist.all_reduce(M, op=dist.ReduceOp.MAX, group=group)
    Mn = M.numpy()
    with np.errstate(invalid="ignore"):
        z = torch.from_numpy(np.where(np.isfinite(x), np.exp(x - Mn), 0.0))
    dist.all_reduce(z, op=dist.ReduceOp.SUM, group=group)
    with np.errstate(divide="ignore"):
        return np.where(np.isfinite(Mn), (Mn + np.log(z.numpy())) / inv_gamma, 0.0)


def gather_scores(local_scores, seg_range, n_segments, group=None):
    """All ranks obtain the full score vector [S] (all-gather of the ragged per-rank slices via a zero-padded sum)."""
    import torch
    import torch.distributed as dist
    full = torch.zeros(n_segments, dtype=torch.float32)
    a, b = seg_range
    full[a:b] = torch.as_tensor(np.asarray(local_scores, dtype=np.float32))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            full = full.cuda()
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full.cpu().numpy()


class _DevView:
    """__cuda_array_interface__ view of a raw device buffer so torch can wrap it without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def allreduce_grads(engine, average=False, group=None):
    """One NCCL all-reduce(sum) over the engine's flat gradient buffer, in place on the device."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    ptr, n = engine.grads_device()
    engine.sync()
    t = torch.as_tensor(_DevView(ptr, n), device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if average:
        t.div_(dist.get_world_size(group))
    torch.cuda.current_stream().synchronize()


def allreduce_host(vec, group=None):
    """Host (gloo) all-reduce(sum) of a float vector -- used by the CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(vec, dtype=np.float64))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
