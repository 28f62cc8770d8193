"""Multi-GPU (needs >= 2 GPUs on the box; skipped otherwise): the peer-memory result gather (pn_gather_*, csrc/gather.cu)
- every rank pushes its per-environment results straight into the root's slab over NVLink, the root waits for the flags -
must deliver every rank's bytes exactly, step after step (two alternating slots, ack back-pressure), and a world of 1 must
degenerate to a local copy."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    from peanut_b200 import _lib, parallel
    torch.cuda.set_device(rank)
    parallel.init_from_env("nccl")
    ctx = _lib.Context(rank)
    dev = torch.device("cuda", rank)
    local = torch.zeros((3, 6, 40, 40), dtype=torch.float32, device=dev)
    g = parallel.PeerGather(ctx, local)
    ok = True
    for step in range(7):
        local.copy_(torch.arange(local.numel(), dtype=torch.float32, device=dev).view_as(local) * 0.5 + 1000.0 * rank + step)
        if rank != 0 and step % 2 == 1:
            torch.cuda._sleep(20_000_000)          # a slow rank: the root's wait kernel really waits
        g.step(local)
        if rank == 0:
            res = g.result().clone()               # stream-ordered after the wait
            torch.cuda.synchronize()
            for r in range(world):
                want = torch.arange(local.numel(), dtype=torch.float32, device=dev).view_as(local) * 0.5 + 1000.0 * r + step
                ok = ok and bool(torch.equal(res[r], want))
    torch.cuda.synchronize()
    st = g.status()
    g.close()
    dist.barrier()
    if rank == 0:
        q.put((ok, st))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_gather_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    ok, st = q.get()
    assert ok and st == 0


def test_peer_gather_single_gpu(ctx):
    from peanut_b200 import parallel
    local = torch.randn((4, 6, 24, 24), device="cuda:0")
    g = parallel.PeerGather(ctx, local)
    for _ in range(3):
        local.normal_()
        g.step(local)
        res = g.result()
        torch.cuda.synchronize()
        assert tuple(res.shape) == (1, 4, 6, 24, 24) and torch.equal(res[0], local)
    assert g.status() == 0
    g.close()
