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


def test_single_process_is_a_no_op():
    from dream_b200 import distributed as D
    m = torch.nn.Linear(2, 2)
    m(torch.ones(1, 2)).sum().backward()
    g = m.weight.grad.clone()
    D.allreduce_gradients(m)
    assert torch.equal(g, m.weight.grad)
    assert D.shard_indices(5) == [0, 1, 2, 3, 4]
    assert D.gather_rows([1, 2], 2) == [1, 2]
