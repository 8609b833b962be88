"""world_size-2 gloo test of the multi-GPU control plane (scene sharding, weight broadcast, result gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from umgen_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_scenes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    weights = [torch.randn(257, 33, generator=g), torch.randn(19, generator=g).half()]
    if rank != 0:
        for w in weights:
            w.zero_()
    moved = dp.broadcast_tensors(weights, src=0)
    ref = torch.Generator().manual_seed(5)
    ok = torch.equal(weights[0], torch.randn(257, 33, generator=ref)) and torch.equal(weights[1], torch.randn(19, generator=ref).half())
    mine = dp.shard_scenes(n_scenes, world, rank)
    local = {i: {"map": np.full((1, 2, 4), i, dtype=np.int64)} for i in mine}
    merged = dp.gather_results(local, dst=0)
    q.put((rank, ok, moved, mine, sorted(merged)))
    dist.destroy_process_group()


def test_two_rank_control_plane():
    world, n_scenes = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                                   # broadcast delivered rank 0's weights
    assert res[0][2] == 257 * 33 * 4 + 19 * 2
    assert res[0][3] == [0, 2, 4] and res[1][3] == [1, 3]           # scene i -> rank i mod world, each exactly once
    assert res[0][4] == [0, 1, 2, 3, 4] and res[1][4] == []         # gathered on rank 0 only


def test_shard_edge_cases():
    assert dp.shard_scenes(0, 4, 1) == []
    assert dp.shard_scenes(3, 8, 5) == []
    assert sorted(sum((dp.shard_scenes(8, 8, r) for r in range(8)), [])) == list(range(8))
