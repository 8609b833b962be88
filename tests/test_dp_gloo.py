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


def _scene_worker(rank, world, port, root, out, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.test_runner import Recorder
    from tests._cases import DATASET_CASES
    from umgen_b200 import runner as R
    from umgen_b200.dataset import NuPlanTokenScenes
    _, _, block, gap, _ = DATASET_CASES["short_clip"]
    scenes = NuPlanTokenScenes([root], block_size=block, sampling_gap=gap)
    model = Recorder()
    merged = dp.run_sharded(model, scenes, R.RunSettings(new_frames=2, cond_frames=13, input_cond_frames=13, token_save_path=out))
    q.put((rank, len(model.calls), {k: (v["name"], tuple(v["tokens"]["map"].shape)) for k, v in merged.items()}))
    dist.destroy_process_group()


def test_two_ranks_evaluate_a_dataset(tmp_path):
    """Scene-level data parallelism end to end on the host side: 3 scenes, 2 ranks, each rank runs its share through the scene runner (dataset
    front-end -> inference kwargs -> token pickle), rank 0 ends up with every scene's tokens."""
    import pickle
    from tests._cases import DATASET_CASES, raw_scene
    seed, n, _, _, n_tracks = DATASET_CASES["short_clip"]
    root = tmp_path / "scenes"
    root.mkdir()
    for k in range(3):
        with open(root / f"synthetic_scene_{k:04d}_clip_a.pkl", "wb") as f:
            pickle.dump(raw_scene(seed + k, n, n_tracks), f)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scene_worker, args=(r, 2, port, str(root), str(tmp_path / "tokens"), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 2 and res[1][1] == 1                        # scenes 0, 2 on rank 0; scene 1 on rank 1
    assert res[1][2] == {}
    assert res[0][2] == {i: (f"synthetic_scene_{i:04d}_clip_a", (1, 15, 1024)) for i in range(3)}
    assert sorted(os.listdir(tmp_path / "tokens")) == [f"synthetic_scene_{i:04d}_clip_a_tokens.pkl" for i in range(3)]
