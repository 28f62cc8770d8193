"""CPU, world_size 2 over gloo: the environment partition and the result gather used by bench.py for N > 1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from peanut_b200 import parallel as P


def test_env_partition():
    assert [P.env_range(64, 8, r) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
    assert [P.env_range(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert [P.env_range(1, 4, r) for r in range(4)] == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert P.env_range(0, 2, 1) == (0, 0)
    for n, w in ((64, 8), (5, 2), (7, 3)):
        for e in range(n):
            lo, hi = P.env_range(n, w, P.owner_of(e, n, w))
            assert lo <= e < hi
    with pytest.raises(ValueError):
        P.env_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    r, _, w = P.init_from_env("gloo")
    lo, hi = P.env_range(num_envs, w, r)
    local = torch.stack([torch.full((2, 3), float(e)) for e in range(lo, hi)]) if hi > lo else torch.zeros((0, 2, 3))
    got = P.gather_env_results(local, num_envs, dst=0)
    slow = P.max_over_ranks([1.0 + r, 5.0 - r])
    if r == 0:
        q.put((got.tolist(), slow))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_envs", [4, 5])
def test_gather_world2_gloo(num_envs):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_envs, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got, slow = q.get()
    assert got == [[[float(e)] * 3] * 2 for e in range(num_envs)]
    assert slow == [2.0, 5.0]


def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    from peanut_b200 import _lib
    P.init_from_env("gloo")
    _lib.load().pn_conv_tuning_clear()
    order = []

    def build():
        # what a *_build call does to the process-wide table: rank 0 "tunes" (adds its choices); the others must find
        # rank 0's entries already there when they build
        order.append(_lib.tuning_export())
        if rank == 0:
            _lib.tuning_import("1|100|10|64|64|1|1|1|1|1|0|0|1|0|64|0|0|0|148|0 128 1 0 0\n")
        return rank

    assert P.build_synchronised(build) == rank
    q.put((rank, order[0], _lib.tuning_export()))
    dist.barrier()
    dist.destroy_process_group()


def test_build_synchronised_world2_gloo():
    """Rank 0 builds first, its launch-configuration table reaches the other ranks before they build."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r, before, after = q.get()
        got[r] = (before, after)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    line = "1|100|10|64|64|1|1|1|1|1|0|0|1|0|64|0|0|0|148|0 128 1 0 0\n"
    assert got[0] == ("", line)
    assert got[1] == (line, line)


def test_tuning_table_roundtrip_and_validation():
    from peanut_b200 import _lib
    lib = _lib.load()
    lib.pn_conv_tuning_clear()
    assert _lib.tuning_import("# comment\nk1 64 2 0 0\nk2 256 1 1 0\n") == 2
    assert _lib.tuning_export() == "k1 64 2 0 0\nk2 256 1 1 0\n"
    with pytest.raises(RuntimeError):
        _lib.tuning_import("k3 100 1 0 0\n")   # not a tile width
    with pytest.raises(RuntimeError):
        _lib.tuning_import("k4 64\n")          # malformed
    lib.pn_conv_tuning_clear()
    assert _lib.tuning_export() == ""
