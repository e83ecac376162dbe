"""Frame sharding across GPUs and the single reduce that combines the shards.

The reference parallelises over frames: thread t of N analyses frames t, t+N, ... with a private
``SystemTopology`` and the clones are summed (``topology/mod.rs:139-144, 236-278``).  Here every
rank (one process per GPU) gets a CONTIGUOUS range of analysed frames; range starts are rounded up
to a multiple of the leaflet-assignment period so that every shard contains its own assignment
frames (``leaflets.rs:435-441, 1438-1473``); ``Frequency::Once`` needs the frame-0 table, which is
broadcast.  All accumulators are integers, so the merge is exact and order-free: ONE sum-reduce of
the contiguous accumulator block (NCCL over NVLink on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import abi


def frame_ranges(n_frames: int, world: int, period: int = 1) -> List[Tuple[int, int]]:
    """[(first, last)) per rank over analysed-frame ordinals 0..n_frames; starts are multiples of ``period``."""
    period = max(1, int(period))
    out, start = [], 0
    for r in range(world):
        end = n_frames if r == world - 1 else min(n_frames, -(-((r + 1) * n_frames // world) // period) * period)
        end = max(end, start)
        out.append((start, end))
        start = end
    return out


def assignment_period(setup: abi.EngineSetup) -> int:
    """Period of leaflet assignment in analysed frames (real frequency / step)."""
    if setup.leaflet_mode == abi.LEAFLET_NONE or setup.leaflet_freq_kind == abi.FREQ_ONCE:
        return 1
    return max(1, setup.leaflet_freq // max(1, setup.step))


def pack_block(raw: abi.RawResults) -> np.ndarray:
    """The accumulator block as one int64 vector: [sum][count][map_sum][map_count] (C ABI layout)."""
    parts = [raw.sum.reshape(-1).astype(np.int64), raw.count.reshape(-1).astype(np.int64)]
    if raw.map_sum is not None:
        parts += [raw.map_sum.reshape(-1).astype(np.int64), raw.map_count.reshape(-1).astype(np.int64)]
    return np.concatenate(parts)


def unpack_block(block: np.ndarray, like: abi.RawResults) -> None:
    n = like.sum.size
    like.sum[...] = block[:n].reshape(like.sum.shape)
    like.count[...] = block[n:2 * n].reshape(like.count.shape).astype(np.uint64)
    if like.map_sum is not None:
        m = like.map_sum.size
        like.map_sum[...] = block[2 * n:2 * n + m].reshape(like.map_sum.shape)
        like.map_count[...] = block[2 * n + m:2 * n + 2 * m].reshape(like.map_count.shape).astype(np.uint64)


def reduce_results(raw: abi.RawResults, setup: abi.EngineSetup, dst: int = 0, device: Optional[str] = None) -> Optional[abi.RawResults]:
    """Combine the shards of all ranks on ``dst``: one sum-reduce of the accumulator block, plus a
    gather of the per-frame rows / leaflet tables / normals (each rank owns disjoint frames)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    t = torch.from_numpy(pack_block(raw))
    if device:
        t = t.to(device)
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    payload = dict(n_frames=raw.n_frames, tw_sum=raw.tw_sum, tw_count=raw.tw_count, tw_frame_index=raw.tw_frame_index,
                   leaflets=raw.leaflets, leaflet_frame_index=raw.leaflet_frame_index, normals=raw.normals)
    dist.gather_object(payload, gathered if rank == dst else None, dst=dst)
    if rank != dst:
        return None
    unpack_block(t.cpu().numpy(), raw)
    order = np.argsort([int(g["tw_frame_index"][0]) if g["n_frames"] else 1 << 62 for g in gathered], kind="stable")
    parts = [gathered[i] for i in order if gathered[i]["n_frames"]]

    def cat(key):
        xs = [p[key] for p in parts if p[key] is not None]
        return np.concatenate(xs, axis=0) if xs else None

    raw.n_frames = sum(p["n_frames"] for p in parts)
    raw.tw_frame_index = cat("tw_frame_index")
    raw.tw_sum, raw.tw_count = cat("tw_sum"), cat("tw_count")
    raw.normals = cat("normals")
    lf = [(p["leaflet_frame_index"], p["leaflets"]) for p in parts if p["leaflets"] is not None]
    if lf:
        raw.leaflet_frame_index = np.concatenate([a for a, _ in lf])
        raw.leaflets = np.concatenate([b for _, b in lf], axis=0)
    return raw
