"""Process-per-GPU data parallelism for the DREAM hot path.

The reference's only parallelism is single-process `torch.nn.DataParallel` (dream/network.py:244-256,
281-284): parameters re-broadcast every forward, inputs scattered, outputs gathered, gradients
reduce-added to GPU 0.  Here each rank owns one GPU and a full model replica:

  * inference : frames are sharded by rank, no collective on the data path; per-frame results are
                gathered on the host only for reporting (`gather_rows`).
  * training  : each rank runs forward/backward on its own batch shard; `allreduce_gradients` averages
                the fp32 gradients with ONE flat NCCL all-reduce per bucket over NVLink/NVSwitch
                (equal shards => mean of per-rank means == the reference's global MSE mean,
                network.py:359).  Parameters are never re-broadcast after `broadcast_parameters`.
Works with the `gloo` backend on CPU tensors too (used by the world_size-2 tests).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items, rank=None, world_size=None):
    """Frame indices owned by `rank`: round-robin i % world == rank."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, n_items, world_size))


def broadcast_parameters(module, src=0):
    """One-time sync of parameters and buffers from `src` (replaces DataParallel's per-forward replicate)."""
    _, w = world()
    if w == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def allreduce_gradients(module, bucket_bytes=64 << 20):
    """Average .grad of all parameters across ranks using flat fp32 buckets."""
    _, w = world()
    if w == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    bucket, size = [], 0
    pending = []

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        pending.append((work, flat, bucket))
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    for work, flat, bucket_ in pending:
        work.wait()
        flat.div_(w)
        off = 0
        for g in bucket_:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


def gather_rows(local_rows, n_items):
    """Collect per-frame python rows from all ranks on every rank, restoring global frame order."""
    r, w = world()
    if w == 1:
        return list(local_rows)
    gathered = [None] * w
    dist.all_gather_object(gathered, list(local_rows))
    out = [None] * n_items
    for rk, rows in enumerate(gathered):
        for j, idx in enumerate(range(rk, n_items, w)):
            out[idx] = rows[j]
    return out
