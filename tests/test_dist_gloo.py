"""world_size-2 gloo tests of the multi-GPU host logic (CPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import igdm_b200  # noqa
    from igdm_b200 import dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # communicator bootstrap payload: rank 0's 128 bytes reach everybody
        payload = bytes(range(128)) if rank == 0 else b"\x00" * 128
        got = dist.broadcast_bytes(payload, 128, src=0)
        ok_b = got == bytes(range(128))
        # sampling shards: disjoint cover, gathered in global order
        n_total = 11
        b, e = dist.shard_range(n_total, rank, world)
        local = np.arange(b, e, dtype=np.float32).reshape(-1, 1, 1, 1) * np.ones((1, 1, 2, 2), np.float32)
        full = dist.gather_shards(local, n_total)
        ok_g = full.shape == (n_total, 1, 2, 2) and np.array_equal(full[:, 0, 0, 0], np.arange(n_total))
        # data-parallel gradient averaging contract: sum of per-rank grads scaled by 1/(B*world)
        g_local = torch.full((4,), float(rank + 1)) / world
        td.all_reduce(g_local)
        ok_a = torch.allclose(g_local, torch.full((4,), sum(range(1, world + 1)) / world))
        q.put((rank, ok_b, ok_g, bool(ok_a)))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res
