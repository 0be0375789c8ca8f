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
    """Average .grad of all parameters across ranks using flat fp32 buckets, AFTER backward has finished (the simple
    form; `GradReducer` below overlaps the buckets with backward).  Buckets are built from `requires_grad`, not from
    which gradients happen to exist on this rank, and a missing gradient counts as zero: every rank then reduces
    identical bucket sizes even if a parameter was unused on one of them."""
    _, w = world()
    if w == 1:
        return
    params = [p for p in module.parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
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


class GradReducer:
    """Bucketed gradient all-reduce that runs WHILE backward is still producing gradients (what
    DataParallel's reduce-add to GPU 0, dream/network.py:244-256, becomes with one process per GPU).

    * One persistent flat fp32 buffer holds every trainable parameter's gradient; `p.grad` is a view into it (no
      `torch.cat` into a staging buffer, no copy back).  The buffer is laid out in REVERSE parameter order -- the
      order in which backward finishes gradients (head first) -- and cut into buckets of ~`bucket_bytes`.
    * A bucket's all-reduce is issued the moment its last gradient lands, strictly in bucket order (identical launch
      order on every rank).  Gradients land either through `deposit()` -- called from inside the hand-written
      backward passes (dream_b200.autograd) as each layer's weight gradient kernel has been queued -- or, for any
      other autograd graph, through torch's post-accumulate-grad hooks.
    * NCCL runs the collective on its own stream behind the gradient kernels already queued; `finish()` makes the
      compute stream wait for the buckets and (gloo: SUM + scale, NCCL: AVG) leaves the rank average in `p.grad`.
    Usage per step: `begin_step()`, forward, `loss.backward()`, `finish()`, `optimizer.step()`."""

    def __init__(self, module, bucket_bytes=24 << 20, process_group=None, tail_bytes=1 << 20):
        self.group = process_group
        self.params = [p for p in module.parameters() if p.requires_grad]
        assert self.params, "GradReducer: the module has no trainable parameter"
        dev, dt = self.params[0].device, self.params[0].dtype
        assert all(p.device == dev and p.dtype == dt for p in self.params), \
            "GradReducer: parameters must share one device and dtype"
        order = list(reversed(self.params))
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros((total,), dtype=dt, device=dev)
        self.views, self.bucket_of, self.buckets = {}, {}, []      # buckets: [start, end, n_params]
        # The LAST bucket is the only one whose all-reduce cannot hide behind backward (its gradients are the last to
        # be produced), so it is kept small: the trailing parameters -- the network's first layers -- up to `tail_bytes`
        # get a bucket of their own, and what is exposed is the latency of a ~1 MB collective instead of the transfer
        # of whatever the greedy cut left over (measured at 8 GPUs: 1.0 ms of a 54 ms step).
        tail_from, acc = len(order), 0
        while tail_from > 1 and acc + order[tail_from - 1].numel() * self.flat.element_size() <= tail_bytes:
            tail_from -= 1
            acc += order[tail_from].numel() * self.flat.element_size()
        off, start, count = 0, 0, 0
        for i, p in enumerate(order):
            n = p.numel()
            self.views[id(p)] = self.flat[off:off + n].view_as(p)
            self.bucket_of[id(p)] = len(self.buckets)
            off += n
            count += 1
            if (off - start) * self.flat.element_size() >= bucket_bytes or (i + 1 == tail_from and i + 1 < len(order)):
                self.buckets.append([start, off, count])
                start, count = off, 0
        if count:
            self.buckets.append([start, off, count])
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_accumulated) for p in self.params]
        self._ready = [0] * len(self.buckets)
        self._seen = set()
        self._next = 0
        self._works = []
        self._active = False
        self.record_exposed = False                 # bench.py: CUDA-event pairs around finish()'s stream waits
        self.exposed_events = []
        self.world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        self._avg = self.world > 1 and dist.get_backend(self.group) == "nccl"
        for mod in module.modules():                # the hand-written backward passes look this up on their model
            mod._grad_sink = self

    def allreduce_alone_ms(self, reps=3):
        """The buckets' all-reduces back to back with nothing else running (CUDA events, ms): what the collective
        would cost if it were NOT overlapped with backward."""
        if self.world == 1 or not self.flat.is_cuda:
            return 0.0
        keep = self.flat.clone()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        dist.barrier(group=self.group)
        e0.record()
        for _ in range(reps):
            for start, end, _n in self.buckets:
                dist.all_reduce(self.flat[start:end], group=self.group)
        e1.record()
        torch.cuda.synchronize()
        self.flat.copy_(keep)
        return e0.elapsed_time(e1) / reps

    def remove(self, module):
        for h in self._hooks:
            h.remove()
        for mod in module.modules():
            if getattr(mod, "_grad_sink", None) is self:
                del mod._grad_sink

    # -- per step -------------------------------------------------------------------------------
    def begin_step(self):
        """Zero the flat buffer and (re-)attach every `p.grad` to its view (`optimizer.zero_grad()` detaches them)."""
        self.flat.zero_()
        for p in self.params:
            p.grad = self.views[id(p)]
        self._ready = [0] * len(self.buckets)
        self._seen = set()
        self._next = 0
        self._works = []
        self._active = True

    def accepts(self, param):
        return self._active and id(param) in self.views

    def deposit(self, param, grad):
        """Write a finished gradient (any layout broadcastable to the parameter's shape) straight into the flat
        buffer and count it towards its bucket.  The caller returns None for this parameter to autograd."""
        self.views[id(param)].copy_(grad)
        self._mark(param)

    def _on_accumulated(self, param):
        if self._active:
            self._mark(param)

    def _mark(self, param):
        if id(param) in self._seen:
            return
        self._seen.add(id(param))
        b = self.bucket_of[id(param)]
        self._ready[b] += 1
        self._launch_ready()

    def _launch_ready(self, force=False):
        while self._next < len(self.buckets) and (force or self._ready[self._next] >= self.buckets[self._next][2]):
            start, end, _ = self.buckets[self._next]
            if self.world > 1:
                op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
                self._works.append(dist.all_reduce(self.flat[start:end], op=op, group=self.group, async_op=True))
            self._next += 1

    def finish(self):
        """Issue whatever has not been issued (parameters without a gradient this step count as zero), wait for all
        buckets on the current stream, leave the average over ranks in every `p.grad`."""
        self._launch_ready(force=True)
        timed = self.record_exposed and self.flat.is_cuda
        if timed:                                   # how long the compute stream stalls for gradients still in flight
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        for w in self._works:
            w.wait()
        if self.world > 1 and not self._avg:
            self.flat.div_(self.world)
        if timed:
            e1.record()
            self.exposed_events.append((e0, e1))
        self._works = []
        self._active = False


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
