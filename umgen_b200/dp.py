"""Scene-level data parallelism: rollouts of different scenes are independent, so scene i
runs on rank i mod world with a full weight replica and NO data-path collective.  torch.distributed is used
for two control-plane steps only: the weight broadcast from rank 0 at start-up (NCCL over NVLink on GPUs) and
the gather of the finished int64 token arrays."""
from __future__ import annotations

from typing import Dict, Iterable, List

import torch
import torch.distributed as dist


def shard_scenes(n_scenes: int, world: int, rank: int) -> List[int]:
    """Indices of the scenes rank `rank` decodes (round robin, like scene i -> GPU i mod G)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_scenes, world))


def broadcast_tensors(tensors: Iterable[torch.Tensor], src: int = 0) -> int:
    """In-place broadcast of every tensor from `src`; returns the number of bytes moved per receiving rank."""
    n = 0
    for t in tensors:
        dist.broadcast(t, src=src)
        n += t.numel() * t.element_size()
    return n


def engine_tensors(engine) -> List[torch.Tensor]:
    """Every packed device tensor of a UMGenEngine (TAR stacks, ego decoders, OAR/head blobs, tables)."""
    out: List[torch.Tensor] = []
    for blocks in engine.tar.stacks.values():
        for blk in blocks:
            for sub in blk:
                out += [v for v in sub.values() if torch.is_tensor(v)]
    for d in engine.tar.ego_dec:
        out += list(d.values())
    out += list(engine.tar.ln.values()) + list(engine.tar.tables.values())
    out += [engine.tar.head_ego, engine.tar.egoe, engine.tar.map_table, engine.tar.grid_pos]
    out += [v for v in engine.dec.w.values() if torch.is_tensor(v)]
    return out


def gather_results(local: Dict[int, dict], dst: int = 0) -> Dict[int, dict]:
    """Collect {scene index: token dict} from every rank on `dst` (small: <= 0.9 MB per 50-frame scene)."""
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local, bucket, dst=dst)
    merged: Dict[int, dict] = {}
    if bucket is not None:
        for part in bucket:
            merged.update(part)
    return merged


def run_sharded(model, scenes, settings, mapdecoder=None, imagedecoder=None, device=None, video=None, dst: int = 0) -> Dict[int, dict]:
    """A dataset evaluated data-parallel: this rank runs its share of the scenes through ``runner.run_dataset`` (token pickles, decoded values /
    pixels and videos are written where the rank runs), then the token arrays of all scenes are gathered on ``dst`` as
    ``{scene index: {"name", "tokens"}}`` (empty dict elsewhere; scenes whose token pickle existed are skipped like in the reference).  Without an initialised process group it is the one-rank loop."""
    from . import runner
    on = dist.is_available() and dist.is_initialized()
    world, rank = (dist.get_world_size(), dist.get_rank()) if on else (1, 0)
    mine = shard_scenes(len(scenes), world, rank)
    res = runner.run_dataset(model, scenes, settings, mapdecoder, imagedecoder, device, indices=mine, video=video)
    local = {r["index"]: {"name": r["name"], "tokens": r["tokens"]} for r in res}          # scenes already processed are skipped by the runner and absent here
    return gather_results(local, dst=dst) if on else local
