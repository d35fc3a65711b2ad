"""N > 1 host logic on CPU: world_size-2 gloo processes (the data path itself has no collective)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rpo_b200.dist import FlatGradBucket, broadcast_parameters, global_loop_test, shard_rows
    torch.manual_seed(100 + rank)                                  # deliberately different initial weights
    net = torch.nn.Sequential(torch.nn.Linear(57, 32), torch.nn.Tanh(), torch.nn.Linear(32, 14))
    broadcast_parameters([net])
    bucket = FlatGradBucket(net.parameters())
    g = torch.Generator().manual_seed(0)
    X = torch.randn(64, 57, generator=g)
    Y = torch.randn(64, 14, generator=g)
    lo, hi = shard_rows(64, rank, world)
    bucket.zero_()
    loss = ((net(X[lo:hi]) - Y[lo:hi]) ** 2).sum() / 64 * world      # local mean over the GLOBAL batch, times world
    loss.backward()
    bucket.all_reduce_mean()
    flat = bucket.flat.clone()
    # loop-test flags: OR across shards
    flags = torch.tensor([1, 1, 0, 0] if rank == 0 else [1, 0, 1, 0], dtype=torch.int32)
    global_loop_test(flags)
    w0 = torch.cat([p.data.reshape(-1) for p in net.parameters()])
    q.put((rank, flat, flags, w0, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_flat_bucket_allreduce_equals_single_process_gradient():
    world, port = 2, _free_port()                                    # generous timeouts: a loaded CI host spawns slowly
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, f0, fl0, w0, s0), (_, f1, fl1, w1, s1) = res
    assert torch.equal(w0, w1), "broadcast_parameters must equalise the replicas"
    assert torch.equal(f0, f1) and fl0.tolist() == fl1.tolist() == [1, 1, 1, 0]
    assert s0 == (0, 32) and s1 == (32, 64)
    # single-process gradient on the concatenated batch
    net = torch.nn.Sequential(torch.nn.Linear(57, 32), torch.nn.Tanh(), torch.nn.Linear(32, 14))
    off = 0
    for p in net.parameters():
        p.data.copy_(w0[off:off + p.numel()].view_as(p))
        off += p.numel()
    g = torch.Generator().manual_seed(0)
    X = torch.randn(64, 57, generator=g)
    Y = torch.randn(64, 14, generator=g)
    (((net(X) - Y) ** 2).sum() / 64).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(f0, ref, rtol=1e-5, atol=1e-6)
