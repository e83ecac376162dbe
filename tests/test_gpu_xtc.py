"""End-to-end trajectory path: XTC file -> host decode threads -> pinned plane batches -> engine (gorder_gpu_run_xtc)."""
import numpy as np
import pytest

from gorder_b200 import SystemTopology, abi, synthetic
from gorder_b200.xtc import XtcFile, write_xtc

from parity import assert_raw_parity

pytestmark = pytest.mark.gpu


def _direct(setup, xyz, box, idx):
    eng = SystemTopology(setup)
    eng.analyze_frames(xyz, box, idx)
    r = eng.finish()
    eng.close()
    return r


def test_run_xtc_equals_decoded_frames(tmp_path):
    """The feed gives the engine exactly the frames a reader would decode (bit-identical accumulators), for several
    batch sizes / thread counts, with begin / stride, and matches the oracle on them."""
    from oracle import oracle as orc
    s = synthetic.s_cg(2600, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True)
    xyz, box, idx = s.frames(0, 13)
    path = str(tmp_path / "cg.xtc")
    write_xtc(path, xyz, box)
    with XtcFile(path) as x:
        dec, box9, _, _ = x.read()
        dbox = np.ascontiguousarray(box9[:, [0, 4, 8]])
        want = _direct(s.setup, dec, dbox, idx)
        for batch, threads in ((4, 3), (32, 1), (5, 8)):
            eng = SystemTopology(s.setup)
            eng.run_xtc(x, n_threads=threads, batch_frames=batch)
            got = eng.finish()
            eng.close()
            np.testing.assert_array_equal(got.sum, want.sum)
            np.testing.assert_array_equal(got.count, want.count)
            np.testing.assert_array_equal(got.tw_sum, want.tw_sum)
            np.testing.assert_array_equal(got.leaflets, want.leaflets)
        # frames 2, 5, 8, 11 (begin = 2, step = 3): analysed-frame indices 0, 3, 6, 9 as in the reference's frame counter
        s2 = synthetic.s_cg(2600, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True, step=3)
        eng = SystemTopology(s2.setup)
        eng.run_xtc(x, first=2, last=12, stride=3, batch_frames=3)
        got = eng.finish()
        eng.close()
        sel = np.arange(2, 12, 3)
        want2 = _direct(s2.setup, dec[sel], dbox[sel], np.arange(4, dtype=np.int64) * 3)
        np.testing.assert_array_equal(got.sum, want2.sum)
        np.testing.assert_array_equal(got.tw_frame_index, [0, 3, 6, 9])
    ref = orc.Oracle(s.setup, n_threads=4)
    ref.analyze_frames(dec, dbox, idx)
    r = ref.finish()
    ref.close()
    assert_raw_parity(want, r, s.setup, what="xtc-decoded frames")


def test_run_xtc_atom_map_and_solvent(tmp_path):
    """The trajectory holds more atoms than the analysis (water after the lipids; the engine's atoms in another order):
    atom_of_slot maps them, the decoder stops after the last atom it needs."""
    s = synthetic.s_aa(64, n_water=3000)
    xyz, box, idx = s.frames(0, 4)
    n = s.n_atoms
    rng = np.random.default_rng(3)
    perm = rng.permutation(n).astype(np.int32)          # engine atom s is trajectory atom perm[s]
    traj = np.empty_like(xyz)
    traj[:, perm] = xyz
    path = str(tmp_path / "aa.xtc")
    write_xtc(path, traj, box)
    with XtcFile(path) as x:
        dec, box9, _, _ = x.read()
        want = _direct(s.setup, dec[:, perm], np.ascontiguousarray(box9[:, [0, 4, 8]]), idx)
        eng = SystemTopology(s.setup)
        eng.run_xtc(x, atom_of_slot=perm, batch_frames=3, n_threads=2)
        got = eng.finish()
        eng.close()
    np.testing.assert_array_equal(got.sum, want.sum)
    np.testing.assert_array_equal(got.count, want.count)


def test_device_decode_falls_back_to_the_host_decoder(tmp_path):
    """2e6 lattice points per nm: bonded displacements need more than 64 bits per small triple, which the device path
    does not cover -- the host decoder takes over transparently (no bytes of compressed data cross PCIe)."""
    s = synthetic.s_cg(1100, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
    xyz, box, idx = s.frames(0, 5)
    path = str(tmp_path / "w.xtc")
    write_xtc(path, xyz, box, precision=2.0e6)
    with XtcFile(path) as x:
        a = SystemTopology(s.setup)
        a.run_xtc(x, batch_frames=2)
        want = a.finish()
        a.close()
        b = SystemTopology(s.setup)
        moved = b.run_xtc_device(x, batch_frames=2)
        got = b.finish()
        b.close()
    assert moved == 0
    np.testing.assert_array_equal(got.sum, want.sum)
    np.testing.assert_array_equal(got.tw_sum, want.tw_sum)


@pytest.mark.parametrize("case", ["cg", "aa_map", "wide_lattice", "cg_prec100"])
def test_device_decode_equals_host_decode(tmp_path, case):
    """xtc_scan_kernel + xtc_decode_kernel give the engine the same coordinates as the host decoder, bit for bit:
    identical accumulators for runs of bonded beads, scattered solvent with an atom map, per-coordinate bit sizes
    (lattice wider than 2^24) and a coarse lattice; several batch sizes, begin / stride."""
    precision, perm = 1000.0, None
    if case == "aa_map":
        s = synthetic.s_aa(64, n_water=3000, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
        xyz, box, idx = s.frames(0, 7)
        rng = np.random.default_rng(11)
        perm = rng.permutation(s.n_atoms).astype(np.int32)
        traj = np.empty_like(xyz)
        traj[:, perm] = xyz
    else:
        s = synthetic.s_cg(2600, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True)
        xyz, box, idx = s.frames(0, 9)
        traj = xyz
        # 1e6 lattice points per nm: the box needs > 2^24 points per axis (per-coordinate bit sizes), the bonded
        # displacements still fit the 64 bits per triple of the device path (2e6 would not: host fallback, tested below)
        precision = {"cg": 1000.0, "wide_lattice": 1.0e6, "cg_prec100": 100.0}[case]
    path = str(tmp_path / "t.xtc")
    write_xtc(path, traj, box, precision=precision)
    with XtcFile(path) as x:
        eng = SystemTopology(s.setup)
        eng.run_xtc(x, atom_of_slot=perm, batch_frames=4, n_threads=4)
        want = eng.finish()
        eng.close()
        for batch in (2, 64):
            eng = SystemTopology(s.setup)
            moved = eng.run_xtc_device(x, atom_of_slot=perm, batch_frames=batch, n_threads=3)
            got = eng.finish()
            eng.close()
            assert 0 < moved < (1.2 if case == "wide_lattice" else 0.8) * traj.nbytes   # 2 x 10^6 lattice points per nm: ~10 B per atom
            np.testing.assert_array_equal(got.sum, want.sum)
            np.testing.assert_array_equal(got.count, want.count)
            np.testing.assert_array_equal(got.tw_sum, want.tw_sum)
            if got.leaflets is not None:
                np.testing.assert_array_equal(got.leaflets, want.leaflets)
        # begin / stride, two calls continuing the frame count (concatenated reading)
        a = SystemTopology(s.setup)
        a.run_xtc(x, atom_of_slot=perm, first=1, stride=2, batch_frames=3)
        wa = a.finish()
        a.close()
        b = SystemTopology(s.setup)
        b.run_xtc_device(x, atom_of_slot=perm, first=1, last=4, stride=2, batch_frames=3)
        b.run_xtc_device(x, atom_of_slot=perm, first=5, stride=2, batch_frames=3, frame_index0=4)
        wb = b.finish()
        b.close()
        np.testing.assert_array_equal(wa.sum, wb.sum)
        np.testing.assert_array_equal(wa.tw_frame_index, wb.tw_frame_index)


GROMACS_FIXTURES = ["pcpepg_selected.xtc", "cg3.xtc", "pcpepg4.xtc", "multiple_resid_same_name.xtc", "ua_first5.xtc"]


@pytest.mark.parametrize("name", GROMACS_FIXTURES)
def test_device_decode_of_gromacs_written_files(name):
    """xtc_decode_kernel on streams GROMACS wrote (tests/golden/xtc: the reference's pcpepg_selected.xtc, split/cg3.xtc,
    split/pcpepg4.xtc, multiple_resid_same_name.xtc, the first five frames of ua.xtc), checked against the ORACLE's reader
    (oracle/xtc.c), not against this repository's host decoder or writer.  The probe engine uses EVERY atom of the file --
    consecutive atoms form two-atom 'molecules' with one bond -- and keeps per-frame sums: the coordinates live on a
    lattice of 1 / precision nm, so a single atom decoded one lattice step off moves S of its bond by ~1e-2 and the
    frame's integer sum with it."""
    import os
    from oracle import fixtures
    path = os.path.join(os.path.dirname(__file__), "golden", "xtc", name)
    ref = fixtures.read_xtc(path)
    xyz = np.ascontiguousarray(np.asarray(ref.xyz, np.float32))
    n_frames, n_atoms = xyz.shape[0], xyz.shape[1]
    box = np.ascontiguousarray(np.asarray(ref.box, np.float32).reshape(n_frames, 3))
    pairs = abi.MolType(name="PAIR", mol_base=np.arange(0, n_atoms - 1, 2), bond_rel=[(0, 1)])
    moltypes = [pairs]
    if n_atoms % 2:   # the last atom joins the two before it
        pairs.mol_base = pairs.mol_base[:-1]
        moltypes.append(abi.MolType(name="TRIPLE", mol_base=np.array([n_atoms - 3]), bond_rel=[(0, 1), (1, 2)]))
    setup = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=n_atoms, moltypes=moltypes, timewise=True)
    want = _direct(setup, xyz, box, np.arange(n_frames))
    assert int(want.count[:, 0].sum()) == (n_atoms + 1) // 2 * n_frames   # every atom is in exactly one 'molecule'
    with XtcFile(path) as x:
        assert (x.n_atoms, x.n_frames) == (n_atoms, n_frames)
        for batch in (1, 3):
            eng = SystemTopology(setup)
            moved = eng.run_xtc_device(x, batch_frames=batch, n_threads=2)
            got = eng.finish()
            eng.close()
            assert moved > 0   # the device path took the frames (no fall-back to the host decoder)
            np.testing.assert_array_equal(got.tw_sum, want.tw_sum)
            np.testing.assert_array_equal(got.count, want.count)
        host = SystemTopology(setup)
        host.run_xtc(x, batch_frames=2, n_threads=2)
        got = host.finish()
        host.close()
        np.testing.assert_array_equal(got.tw_sum, want.tw_sum)


def test_device_walk_equals_host_walk(tmp_path, monkeypatch):
    """Round 2: the control bits are walked by xtc_walk_kernel (one thread per frame); GORDER_XTC_HOST_WALK=1 keeps the host
    threads' walk.  Same bookmarks, hence the same accumulators; fewer bytes cross PCIe (no bookmarks, no slack)."""
    s = synthetic.s_cg(2600, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
    xyz, box, idx = s.frames(0, 9)
    path = str(tmp_path / "t.xtc")
    write_xtc(path, xyz, box)
    out = {}
    with XtcFile(path) as x:
        for mode in ("device", "host"):
            if mode == "host":
                monkeypatch.setenv("GORDER_XTC_HOST_WALK", "1")
            eng = SystemTopology(s.setup)
            moved = eng.run_xtc_device(x, batch_frames=4, n_threads=3)
            out[mode] = (eng.finish(), moved)
            eng.close()
    np.testing.assert_array_equal(out["device"][0].sum, out["host"][0].sum)
    np.testing.assert_array_equal(out["device"][0].tw_sum, out["host"][0].tw_sum)
    assert 0 < out["device"][1] < out["host"][1]


def test_corrupt_stream_is_an_error_on_the_device_path(tmp_path):
    """Garbage in the compressed bytes of one frame: the device walk loses the thread (atom count / stream length do not
    come out), the decoder skips the frame and the engine reports GORDER_ERR_INVALID_ARGUMENT with the frame's number."""
    s = synthetic.s_cg(600, leaflet_mode=abi.LEAFLET_GLOBAL)
    xyz, box, idx = s.frames(0, 6)
    path = str(tmp_path / "t.xtc")
    write_xtc(path, xyz, box)
    raw = bytearray(open(path, "rb").read())
    with XtcFile(path) as x:
        n_frames = x.n_frames
    per = len(raw) // n_frames
    rng = np.random.default_rng(3)
    lo = 4 * per + 200   # inside the stream of frame 4 (headers are ~92 bytes)
    raw[lo:lo + 400] = bytes(rng.integers(0, 256, 400, dtype=np.uint8))
    open(path, "wb").write(bytes(raw))
    with XtcFile(path) as x:
        eng = SystemTopology(s.setup)
        try:
            with pytest.raises(abi.GorderError) as e:
                eng.run_xtc_device(x, batch_frames=4)
                eng.finish()
            assert e.value.code == abi.ERR_INVALID_ARGUMENT
        finally:
            eng.close()
