"""
Ray sharding across the GPUs of one box (one process per GPU, torch.distributed). The reference
is single-device; this is the new multi-GPU functionality of SURVEY.md 8e:

  * rendering: rank r takes the contiguous ray range shard_range(H*W, r, G) of a view (or whole
    views of a multi-view job), renders with replicated weights and no communication; the only
    collective is the final gather of the [n,5] (rgb, depth, acc) image rows;
  * training: the 4096-ray batch is split B/G per rank, each rank's loss is scaled by
    1/(B_global*3) and ONE sum of the flat gradient buffer over the ranks precedes the fused Adam
    step (NeRF.train_step does this when set_distributed() was called: on GPUs through the library's
    own peer-memory kernel, csrc/peer.cu, with an NCCL all-reduce as the fallback).

Everything here works on any backend (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [start, stop) of n units for `rank`; the union over ranks is [0, n)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_views(num_views, rank, world):
    """Round-robin view ids of a multi-view job (evaluate.py-shaped 200-view test set)."""
    return list(range(rank, num_views, world))


def gather_rows(local, n_total, group=None, dst=None):
    """The final image gather of a ray-sharded render: row shards produced with shard_range ->
    [n_total, ...] on every rank (or on `dst` only; other ranks get None). ONE collective: an all-gather into a single
    [world, rows_max, ...] buffer (NCCL: one kernel over NVLink; shards are padded to the largest one, which differs
    from the others by at most one row)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(b - a for a, b in sizes)
    local = local.contiguous()
    if local.shape[0] == pad:
        buf = local
    else:
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
    out = torch.empty((world, pad) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    try:
        dist.all_gather_into_tensor(out, buf, group=group)
    except (RuntimeError, NotImplementedError):       # a backend without the single-buffer form
        dist.all_gather(list(out.unbind(0)), buf, group=group)
    if dst is not None and rank != dst:
        return None
    if n_total == world * pad:
        return out.reshape((n_total,) + tuple(local.shape[1:]))
    return torch.cat([out[r, :b - a] for r, (a, b) in enumerate(sizes)], dim=0)


def allreduce_flat(flat, group=None):
    """Sum-all-reduce of the single flat fp32 gradient buffer (1 191 688 floats = 4.77 MB)."""
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def max_over_ranks(value, device, group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
