"""CPU, gloo, world_size 2: the N>1 host logic -- batch sharding for inference (no collective on the
data path) and the gradient all-reduce of the training path (reference distributed.py:126-173)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tacotron2_b200.distributed import (allreduce_gradients, apply_gradient_allreduce, gather_rows,
                                        shard_rows)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. shard / gather round trip with an uneven batch (7 rows over 2 ranks)
        full = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)
        a, b = shard_rows(7, rank, world)
        got = gather_rows(full[a:b] * 1.0, 7)
        assert torch.equal(got, full)
        # 2. wrap: state is broadcast from rank 0, gradients are averaged after backward
        torch.manual_seed(100 + rank)
        net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 2))
        apply_gradient_allreduce(net)
        apply_gradient_allreduce(net)               # train.py wraps twice (:79 and :179)
        w0 = [p.detach().clone() for p in net.parameters()]
        x = torch.randn(6, 5, generator=torch.Generator().manual_seed(7 + rank))
        local = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.BatchNorm1d(4), torch.nn.Linear(4, 2))
        local.load_state_dict(net.state_dict())
        local(x).pow(2).sum().backward()
        lg = [p.grad.clone() for p in local.parameters()]
        net(x).pow(2).sum().backward()
        q.put((rank, [t.numpy() for t in w0], [t.numpy() for t in lg], [p.grad.numpy() for p in net.parameters()]))
        # 3. bucketing with tiny buckets gives the same result
        for p, g in zip(net.parameters(), lg):
            p.grad = g.clone()
        allreduce_gradients(net, bucket_bytes=16)
        q.put((rank + 10, [p.grad.numpy() for p in net.parameters()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_sharding_and_gradient_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2 * world):
        item = q.get(timeout=100)
        res[item[0]] = item[1:]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    w0_r0, lg0, red0 = res[0]
    w0_r1, lg1, red1 = res[1]
    import numpy as np
    for a, b in zip(w0_r0, w0_r1):
        assert np.array_equal(a, b)                           # broadcast made the replicas identical
    for g0, g1, r0, r1 in zip(lg0, lg1, red0, red1):
        assert np.allclose(r0, (g0 + g1) / 2, atol=1e-6) and np.array_equal(r0, r1)
    for r0, b0, b1 in zip(red0, res[10][0], res[11][0]):
        assert np.allclose(b0, r0, atol=1e-6) and np.array_equal(b0, b1)


def test_shard_rows_partitions_exactly():
    for n in (1, 7, 64, 256):
        for world in (1, 2, 4, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
