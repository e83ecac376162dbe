"""Frames sharded over two GPUs and merged behind the C ABI (gorder_gpu_reduce: one process, peer access;
gorder_gpu_reduce_comm: one process per GPU, NCCL) equal the single-GPU analysis bit for bit.

Reference semantics: ParallelTrajData::reduce + the Add chain (topology/mod.rs:236-278), the interleave of per-frame
vectors (common.rs:380-404), leaflet assignment frequency across workers (leaflets.rs:435-441, 1438-1473, 1523-1577).
Needs two devices: `gpurun --gpus 2`; skipped otherwise."""
import dataclasses
import os
import subprocess
import sys

import numpy as np
import pytest

from gorder_b200 import SystemTopology, abi, sharding, synthetic
from gorder_b200.topology import reduce_handles

pytestmark = pytest.mark.gpu


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs (gpurun --gpus 2)")


def _single(setup, xyz, box, idx):
    e = SystemTopology(setup)
    e.analyze_frames(xyz, box, idx)
    r = e.finish()
    e.close()
    return r


def _assert_same(a: abi.RawResults, b: abi.RawResults):
    assert a.n_frames == b.n_frames
    for k in ("sum", "count", "tw_sum", "tw_count", "tw_frame_index", "map_sum", "map_count", "leaflets", "leaflet_frame_index"):
        x, y = getattr(a, k), getattr(b, k)
        assert (x is None) == (y is None), k
        if x is not None:
            np.testing.assert_array_equal(x, y, err_msg=k)
    if a.normals is not None:
        np.testing.assert_array_equal(np.isnan(a.normals), np.isnan(b.normals))
        np.testing.assert_array_equal(np.nan_to_num(a.normals), np.nan_to_num(b.normals))


def _cases():
    g = dict(leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True)
    return {
        "every3": (synthetic.s_cg(700, leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=3, **g), 10),
        "every1_fast_kernel": (synthetic.s_cg(2600, **g), 9),
        "once": (synthetic.s_cg(700, leaflet_freq_kind=abi.FREQ_ONCE, **g), 10),
        "maps": (synthetic.s_aa(40, n_water=0, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=(0.5, 0.5), map_span_x=(0.0, 4.0),
                                map_span_y=(0.0, 4.0), timewise=True), 7),
        "dynamic_normals": (synthetic.s_cg(300, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0, collect_normals=True, timewise=True), 6),
        "ua_error_blocks": (synthetic.s_ua(100, timewise=True, leaflet_mode=abi.LEAFLET_INDIVIDUAL), 8),
        # BASELINE configs[4]: vesicle, dynamic normals, spherical-clustering leaflets assigned once, sharded by frame ranges
        "vesicle_once": (synthetic.s_ves(2500, timewise=True, collect_leaflets=True, collect_normals=True), 6),
    }


@needs2
@pytest.mark.parametrize("root", [1, 0])
@pytest.mark.parametrize("name", sorted(_cases()))
def test_two_devices_one_process(name, root):
    """root = 1: the merge happens on the shard that does NOT hold frame 0 (the rows must still come out in frame order);
    root = 0 with the whole trajectory announced (reserve_frames): the peers' rows land behind the root's own, in place."""
    s, n = _cases()[name]
    xyz, box, idx = s.frames(0, n)
    want = _single(s.setup, xyz, box, idx)
    ranges = sharding.frame_ranges(n, 2, sharding.assignment_period(s.setup))
    engines = []
    for dev, (lo, hi) in enumerate(ranges):
        st = dataclasses.replace(s.setup, device=dev)
        engines.append(SystemTopology(st))
    once = s.setup.leaflet_mode != abi.LEAFLET_NONE and s.setup.leaflet_freq_kind == abi.FREQ_ONCE
    if root == 0:
        engines[0].reserve_frames(n)
    for dev, (lo, hi) in enumerate(ranges):
        if once and dev > 0:   # the table of analysed frame 0 (shard 0) reaches the other shards before they accumulate
            engines[dev].set_leaflets(engines[0].finish().leaflets[0], 0)
        if hi > lo:
            engines[dev].analyze_frames(xyz[lo:hi], box[lo:hi], idx[lo:hi])
    reduce_handles(engines, root=root)
    got = engines[root].finish()
    _assert_same(got, want)
    for e in engines:
        e.close()


@needs2
def test_reduce_reports_the_first_shard_error():
    s = synthetic.s_cg(300, leaflet_mode=abi.LEAFLET_GLOBAL)
    xyz, box, idx = s.frames(0, 4)
    bad = xyz.copy()
    bad[3, 7, 2] = np.nan
    engines = [SystemTopology(dataclasses.replace(s.setup, device=d)) for d in range(2)]
    engines[0].analyze_frames(xyz[:2], box[:2], idx[:2])
    engines[1].analyze_frames(bad[2:], box[2:], idx[2:])
    with pytest.raises(abi.GorderError) as e:
        reduce_handles(engines, root=0)
    assert e.value.code == abi.ERR_UNDEFINED_POSITION
    for x in engines:
        x.close()


_WORKER = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, sys.argv[1])
from gorder_b200 import SystemTopology, abi, sharding, synthetic
from gorder_b200.topology import Comm
rank, world, idfile, out, mode = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], sys.argv[6]
kw = dict(leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True)
if mode == "once":
    kw.update(leaflet_freq_kind=abi.FREQ_ONCE)
else:
    kw.update(leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=3)
s = synthetic.s_cg(700, **kw)
n = 10
xyz, box, idx = s.frames(0, n)
if rank == 0:
    uid = Comm.unique_id()
    open(idfile + ".tmp", "wb").write(uid)
    os.replace(idfile + ".tmp", idfile)
else:
    while not os.path.exists(idfile):
        time.sleep(0.05)
    uid = open(idfile, "rb").read()
comm = Comm(uid, world, rank, rank)
s.setup.device = rank
eng = SystemTopology(s.setup)
if rank == 0 and mode != "once":
    eng.reserve_frames(n)   # the root of this mode: the other shard's per-frame rows land behind its own, in place
lo, hi = sharding.frame_ranges(n, world, sharding.assignment_period(s.setup))[rank]
if mode == "once":
    if rank == 0:
        eng.analyze_frames(xyz[lo:hi], box[lo:hi], idx[lo:hi])
    eng.broadcast_leaflets(comm, 0)
    if rank != 0:
        eng.analyze_frames(xyz[lo:hi], box[lo:hi], idx[lo:hi])
else:
    eng.analyze_frames(xyz[lo:hi], box[lo:hi], idx[lo:hi])
root = world - 1 if mode == "once" else 0
eng.reduce_comm(comm, root)
if rank == root:
    r = eng.finish()
    np.savez(out, sum=r.sum, count=r.count, tw_sum=r.tw_sum, tw_count=r.tw_count, tw_frame_index=r.tw_frame_index, leaflets=r.leaflets,
             leaflet_frame_index=r.leaflet_frame_index)
eng.close()
comm.close()
'''


@needs2
@pytest.mark.parametrize("mode", ["every3", "once"])
def test_two_processes_nccl(tmp_path, mode):
    """One process per GPU, the library's own NCCL communicator (unique id handed over through a file, as an MPI / torchrun
    host would broadcast it): merged result on the last rank ("once") or on rank 0 with the rows gathered in place ("every3")
    == single-GPU result."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    idfile, out = str(tmp_path / "nccl.id"), str(tmp_path / "merged.npz")
    procs = [subprocess.Popen([sys.executable, str(script), root, str(r), "2", idfile, out, mode], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o.decode(errors="replace"))
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    kw = dict(leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True)
    kw.update(dict(leaflet_freq_kind=abi.FREQ_ONCE) if mode == "once" else dict(leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=3))
    s = synthetic.s_cg(700, **kw)
    xyz, box, idx = s.frames(0, 10)
    want = _single(s.setup, xyz, box, idx)
    got = np.load(out)
    for k in ("sum", "count", "tw_sum", "tw_count", "tw_frame_index", "leaflets", "leaflet_frame_index"):
        np.testing.assert_array_equal(got[k], getattr(want, k), err_msg=k)
