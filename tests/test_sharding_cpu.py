"""N>1 path on CPU: world_size-2 gloo processes shard the frames, analyse their shard (the oracle
stands in for the device engine: same RawResults contract) and combine with ONE sum-reduce."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from gorder_b200 import abi, sharding, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_ranges_cover_and_align():
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            for p in (1, 3, 10):
                rs = sharding.frame_ranges(n, w, p)
                assert rs[0][0] == 0 and rs[-1][1] == n and len(rs) == w
                for (a, b), (c, d) in zip(rs[:-1], rs[1:]):
                    assert b == c and a <= b
                for a, b in rs:
                    assert a % p == 0 or a == b or a == n


def _worker(rank, world, port, kw, n_frames, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from oracle import oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = synthetic.s_cg(120, **kw)
    lo, hi = sharding.frame_ranges(n_frames, world, sharding.assignment_period(s.setup))[rank]
    eng = orc.Oracle(s.setup, n_threads=1)
    if s.setup.leaflet_freq_kind == abi.FREQ_ONCE and s.setup.leaflet_mode != abi.LEAFLET_NONE:
        # Frequency::Once: the owner of frame 0 assigns, everybody else receives the table (one broadcast)
        import torch
        table = torch.zeros(s.setup.n_molecules_total, dtype=torch.uint8)
        if rank == 0:
            x0, b0, i0 = s.frames(0, 1)
            probe = orc.Oracle(abi.EngineSetup.from_dict({**s.setup.to_dict(), "collect_leaflets": True}), n_threads=1)
            probe.analyze_frames(x0, b0, i0)
            table = torch.from_numpy(probe.finish().leaflets[0].copy())
        dist.broadcast(table, src=0)
        if rank != 0:
            eng.set_leaflets(table.numpy(), 0)
    if hi > lo:
        xyz, box, idx = s.frames(lo, hi - lo)
        eng.analyze_frames(xyz, box, idx)
    raw = eng.finish()
    merged = sharding.reduce_results(raw, s.setup, dst=0)
    if rank == 0:
        np.savez(out, sum=merged.sum, count=merged.count, tw_sum=merged.tw_sum, tw_frame_index=merged.tw_frame_index, n_frames=merged.n_frames)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kw", [dict(leaflet_mode=abi.LEAFLET_GLOBAL, leaflet_freq=3, timewise=True),
                                dict(leaflet_mode=abi.LEAFLET_GLOBAL, leaflet_freq_kind=abi.FREQ_ONCE, timewise=True)])
def test_two_rank_shards_equal_single_process(tmp_path, kw):
    from oracle import oracle as orc
    n_frames, world = 10, 2
    out = str(tmp_path / "merged.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, kw, n_frames, out), nprocs=world, join=True)
    z = np.load(out)
    s = synthetic.s_cg(120, **kw)
    ref = orc.Oracle(s.setup, n_threads=2)
    xyz, box, idx = s.frames(0, n_frames)
    ref.analyze_frames(xyz, box, idx)
    r = ref.finish()
    assert int(z["n_frames"]) == n_frames
    np.testing.assert_array_equal(z["sum"], r.sum)
    np.testing.assert_array_equal(z["count"], r.count)
    np.testing.assert_array_equal(z["tw_frame_index"], r.tw_frame_index)
    np.testing.assert_array_equal(z["tw_sum"], r.tw_sum)
