"""CPU, world_size 2, gloo: the host-side logic of the N>1 path (sharding, gradient averaging, gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dream_b200 import distributed as D
    torch.manual_seed(rank)
    model = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    D.broadcast_parameters(model)
    ref = [p.detach().clone() for p in model.parameters()]
    x = torch.randn(4, 5)
    model(x).pow(2).mean().backward()
    local = [p.grad.clone() for p in model.parameters()]
    D.allreduce_gradients(model, bucket_bytes=32)          # tiny buckets: exercises the multi-bucket path
    idx = D.shard_indices(7)
    rows = D.gather_rows([("frame", i, rank) for i in idx], 7)
    torch.save({"params": ref, "local": local, "avg": [p.grad.clone() for p in model.parameters()],
                "idx": idx, "rows": rows}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_world2_gradient_average_and_sharding(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "r1.pt", weights_only=False)
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)                            # broadcast made replicas identical
    for l0, l1, a0, a1 in zip(r0["local"], r1["local"], r0["avg"], r1["avg"]):
        assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-7) and torch.equal(a0, a1)
    assert r0["idx"] == [0, 2, 4, 6] and r1["idx"] == [1, 3, 5]
    assert [r[1] for r in r0["rows"]] == list(range(7)) and r0["rows"] == r1["rows"]
    assert [r[2] for r in r0["rows"]] == [0, 1, 0, 1, 0, 1, 0]


def _reducer_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dream_b200 import distributed as D
    torch.manual_seed(rank)
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.ones(3))                # a trainable parameter no loss touches
    model.register_parameter("unused", unused)
    D.broadcast_parameters(model)
    red = D.GradReducer(model, bucket_bytes=48)               # several small buckets
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    out = {"n_buckets": len(red.buckets)}
    for step in range(2):
        x = torch.randn(4, 5)
        # local reference gradient (plain autograd, no reducer involved)
        ref = torch.autograd.grad(model(x).pow(2).mean(), [p for p in model.parameters() if p is not unused])
        opt.zero_grad()                                       # detaches p.grad (set_to_none): begin_step re-attaches
        red.begin_step()
        loss = model(x).pow(2).mean()
        if step == 1:
            # the hand-written backward passes hand gradients over with deposit(); emulate that for one parameter:
            # autograd then sees no gradient for it (detached use), the reducer gets it directly
            w0 = model[0].weight
            assert red.accepts(w0)
            red.deposit(w0, ref[0])
            with torch.no_grad():
                saved = w0.detach().clone()
            loss = torch.nn.functional.linear(x, saved, model[0].bias)
            loss = model[2](model[1](loss)).pow(2).mean()
        loss.backward()
        assert all(p.grad.data_ptr() == red.views[id(p)].data_ptr() for p in model.parameters())
        red.finish()
        out["local%d" % step] = [g.clone() for g in ref]
        out["avg%d" % step] = [p.grad.clone() for p in model.parameters() if p is not unused]
        out["unused%d" % step] = unused.grad.clone()
        opt.step()
    out["params"] = [p.detach().clone() for p in model.parameters()]
    torch.save(out, os.path.join(out_dir, "red%d.pt" % rank))
    dist.destroy_process_group()


def test_world2_grad_reducer_buckets_hooks_and_deposit(tmp_path):
    """GradReducer: p.grad are views of one flat buffer, buckets are reduced as gradients land (autograd hooks and
    direct deposits), parameters without a gradient count as zero, replicas stay identical after optimizer steps."""
    port = _free_port()
    mp.spawn(_reducer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "red0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "red1.pt", weights_only=False)
    assert r0["n_buckets"] >= 3
    for step in range(2):
        for l0, l1, a0, a1 in zip(r0["local%d" % step], r1["local%d" % step], r0["avg%d" % step], r1["avg%d" % step]):
            assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-7) and torch.equal(a0, a1)
        assert torch.equal(r0["unused%d" % step], torch.zeros(3))
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)


def test_grad_reducer_single_process_leaves_plain_gradients():
    from dream_b200 import distributed as D
    m = torch.nn.Linear(3, 2)
    x = torch.ones(2, 3)
    ref = torch.autograd.grad(m(x).sum(), list(m.parameters()))
    red = D.GradReducer(m)
    red.begin_step()
    m(x).sum().backward()
    red.finish()
    for p, g in zip(m.parameters(), ref):
        assert torch.equal(p.grad, g)
    red.remove(m)
    assert not hasattr(m, "_grad_sink")


def test_single_process_is_a_no_op():
    from dream_b200 import distributed as D
    m = torch.nn.Linear(2, 2)
    m(torch.ones(1, 2)).sum().backward()
    g = m.weight.grad.clone()
    D.allreduce_gradients(m)
    assert torch.equal(g, m.weight.grad)
    assert D.shard_indices(5) == [0, 1, 2, 3, 4]
    assert D.gather_rows([1, 2], 2) == [1, 2]


def test_grad_reducer_bucket_layout_keeps_the_last_bucket_small():
    """Bucket layout only (no process group needed): reverse parameter order, ~bucket_bytes per bucket, and the trailing
    parameters -- the first layers, whose gradients are the last to be produced -- in a small bucket of their own, so
    that the one all-reduce that cannot hide behind backward is a short one."""
    from dream_b200 import distributed as D
    sizes = [100, 2000, 6000, 9000, 9000, 500]                       # parameters in forward order, numels
    model = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(n)) for n in sizes])
    r = D.GradReducer(model, bucket_bytes=10000 * 4, tail_bytes=2500 * 4)
    assert r.flat.numel() == sum(sizes)
    # reverse order, a bucket closes once it holds >= 10000 elements: 500 + 9000 + 9000 | 6000 (closed early by the tail
    # cut) | 2000 + 100 = the tail (<= 2500 elements)
    assert [b[1] - b[0] for b in r.buckets] == [18500, 6000, 2100]
    assert [b[2] for b in r.buckets] == [3, 1, 2]
    params = list(model)
    assert r.bucket_of[id(params[0])] == r.bucket_of[id(params[1])] == 2
    assert r.views[id(params[5])].data_ptr() == r.flat.data_ptr()    # the last parameter's gradient comes first
    # a single huge first parameter cannot be split: it simply is the tail
    model2 = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(n)) for n in (50000, 10, 10)])
    r2 = D.GradReducer(model2, bucket_bytes=10000 * 4, tail_bytes=2500 * 4)
    assert sum(b[1] - b[0] for b in r2.buckets) == 50020 and r2.buckets[-1][1] == 50020
