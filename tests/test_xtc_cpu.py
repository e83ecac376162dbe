"""Host-side trajectory feed (gorder_xtc_* in the C ABI): XTC writer / reader, no GPU needed.

The product reader is checked against (i) its own writer (round trip on the lattice), (ii) the oracle's independent
reader (oracle/xtc.c, the restatement of xdrfile's xdr3dfcoord used for the fixtures) on the same files and
(iii) -- only where /root/reference is mounted -- the reference's own trajectory tests/files/ua.xtc.
"""
import os

import numpy as np
import pytest

from gorder_b200 import synthetic
from gorder_b200.xtc import XtcFile, write_xtc


def _lattice(xyz, precision):
    lf = xyz.astype(np.float32) * np.float32(precision)
    li = np.where(lf >= 0, lf + np.float32(0.5), lf - np.float32(0.5)).astype(np.int32)
    return li.astype(np.float32) * (np.float32(1.0) / np.float32(precision))


def _oracle_read(path):
    from oracle import fixtures
    return fixtures.read_xtc(path)


@pytest.mark.parametrize("kind,precision", [("cg", 1000.0), ("aa", 1000.0), ("cg", 100.0), ("gas", 1000.0), ("tiny", 1000.0)])
def test_round_trip_and_oracle_reader(tmp_path, kind, precision):
    rng = np.random.default_rng(7)
    if kind == "cg":
        s = synthetic.s_cg(700)
        xyz, box, _ = s.frames(0, 5)
    elif kind == "aa":
        s = synthetic.s_aa(40, n_water=900)     # bonded hydrogens (runs) + scattered water (no runs)
        xyz, box, _ = s.frames(0, 4)
    elif kind == "gas":                          # huge box: per-coordinate bit sizes (sizeint > 2^24), negative coordinates
        xyz = (rng.random((3, 500, 3)) * 40000.0 - 20000.0).astype(np.float32)
        box = np.full((3, 3), 40000.0, np.float32)
    else:                                        # <= 9 atoms: stored uncompressed
        xyz = rng.random((4, 7, 3)).astype(np.float32)
        box = np.full((4, 3), 3.0, np.float32)
    path = str(tmp_path / "t.xtc")
    write_xtc(path, xyz[:2], box[:2], precision=precision, dt=2.0)
    write_xtc(path, xyz[2:], box[2:], precision=precision, append=True, first_step=2, dt=2.0)
    with XtcFile(path) as x:
        assert (x.n_atoms, x.n_frames) == (xyz.shape[1], xyz.shape[0])
        got, box9, time, step = x.read(n_threads=3)
        want = xyz if kind == "tiny" else _lattice(xyz, precision)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(box9[:, [0, 4, 8]], box)
        assert not box9[:, [1, 2, 3, 5, 6, 7]].any()
        np.testing.assert_array_equal(step, np.arange(xyz.shape[0]))
        np.testing.assert_allclose(time, 2.0 * np.arange(xyz.shape[0]))
        # strided random access, single thread
        sub, _, _, st2 = x.read(first=1, stride=2, n_threads=1)
        np.testing.assert_array_equal(sub, want[1::2])
        np.testing.assert_array_equal(st2, np.arange(xyz.shape[0])[1::2])
    if kind in ("cg", "aa"):
        assert os.path.getsize(path) < 0.6 * xyz.nbytes     # it does compress
    ref = _oracle_read(path)
    np.testing.assert_array_equal(np.asarray(ref.xyz, np.float32).reshape(got.shape), got)


def test_compression_uses_runs(tmp_path):
    """Bonded beads 0.47 nm apart are stored as small displacements (runs): well under the 6+ bytes / atom of the
    run-free encoding for this box."""
    s = synthetic.s_cg(3000)
    xyz, box, _ = s.frames(0, 2)
    path = str(tmp_path / "c.xtc")
    write_xtc(path, xyz, box)
    assert os.path.getsize(path) / (2 * xyz.shape[1]) < 5.0


def test_truncated_and_bad_files(tmp_path):
    s = synthetic.s_cg(100)
    xyz, box, _ = s.frames(0, 3)
    path = str(tmp_path / "t.xtc")
    write_xtc(path, xyz, box)
    raw = open(path, "rb").read()
    cut = str(tmp_path / "cut.xtc")
    open(cut, "wb").write(raw[: len(raw) - 100])           # last frame incomplete: ignored
    with XtcFile(cut) as x:
        assert x.n_frames == 2
    # a forged 64-bit length in a magic-2023 header: 0xFFFFFFFFFFFFFFF0 wraps the cursor, 2^40 points far
    # outside the mapping; both must stop the index at the frames before, not crash or loop
    import struct
    for forged in (0xFFFFFFFFFFFFFFF0, 1 << 40, len(raw)):
        n_first = raw.index(struct.pack(">i", 1995), 4)          # start of the second frame
        head = bytearray(raw[n_first:n_first + 56 + 32])
        head[0:4] = struct.pack(">i", 2023)
        evil = str(tmp_path / "evil.xtc")
        open(evil, "wb").write(raw[:n_first] + bytes(head) + struct.pack(">Q", forged) + raw[n_first + 92:])
        with XtcFile(evil) as x:
            assert x.n_frames == 1
            got, _, _, _ = x.read()
            assert got.shape[0] == 1
    bad = str(tmp_path / "bad.xtc")
    open(bad, "wb").write(b"\x00" * 200)
    with pytest.raises(OSError):
        XtcFile(bad)
    with pytest.raises(OSError):
        XtcFile(str(tmp_path / "missing.xtc"))


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/files/ua.xtc"), reason="reference tree not mounted")
@pytest.mark.parametrize("name,shape", [("ua.xtc", (51, 19790)), ("pcpepg_selected.xtc", (4, 68375)), ("ua_whole_nobox.xtc", None)])
def test_reference_trajectories_match_oracle_reader(name, shape):
    """The reference's own trajectories (precision 1000 and 100, with water runs): product reader == oracle reader, bit for bit."""
    path = "/root/reference/tests/files/" + name
    if not os.path.exists(path):
        pytest.skip(name + " not in the reference tree")
    ref = _oracle_read(path)
    with XtcFile(path) as x:
        got, box9, time, step = x.read()
    np.testing.assert_array_equal(got, np.asarray(ref.xyz, np.float32).reshape(got.shape))
    np.testing.assert_array_equal(time, np.asarray(ref.time, np.float32))
    if shape is not None:
        assert got.shape[:2] == shape


GROMACS_FIXTURES = ["pcpepg_selected.xtc", "cg3.xtc", "pcpepg4.xtc", "multiple_resid_same_name.xtc", "ua_first5.xtc"]


@pytest.mark.parametrize("name", GROMACS_FIXTURES)
def test_gromacs_written_fixtures_match_oracle_reader(name):
    """The committed GROMACS-written streams (tests/golden/xtc, copied from the reference's tests/files by make_golden.py):
    product reader == the oracle's independent reader, bit for bit, and the control-bit walk of the device path accepts
    them (the same group structure the decoder kernel will follow)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "xtc", name)
    ref = _oracle_read(path)
    with XtcFile(path) as x:
        got, box9, time, step = x.read(n_threads=2)
        np.testing.assert_array_equal(got, np.asarray(ref.xyz, np.float32).reshape(got.shape))
        np.testing.assert_array_equal(time, np.asarray(ref.time, np.float32))
        groups, marks = x.scan()
        assert np.all(groups > 0) and np.all(marks == (groups + 31) // 32)
        assert np.all(groups <= x.n_atoms)


_FUZZ = r'''
import os, sys, tempfile
import numpy as np
from gorder_b200 import xtc
rng = np.random.default_rng(int(sys.argv[1]))
d = tempfile.mkdtemp()
n = 500
xyz = (rng.normal(0, 1.5, (6, n, 3)) + 5).astype(np.float32)
xyz[:, 1::2] = xyz[:, ::2] + rng.normal(0, 0.1, (6, n // 2, 3)).astype(np.float32)   # pairs of close atoms: runs of small triples
xtc.write_xtc(os.path.join(d, "a.xtc"), xyz, np.full((6, 3), 10, np.float32), precision=1000.0)
raw = np.fromfile(os.path.join(d, "a.xtc"), np.uint8)
ok = 0
for it in range(int(sys.argv[2])):
    b = raw.copy()
    if it % 4 == 0:                                   # scattered byte flips
        for _ in range(rng.integers(1, 8)):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
    elif it % 4 == 1:                                 # truncation anywhere
        b = b[: rng.integers(0, len(b))]
    elif it % 4 == 2:                                 # a header word of the first frame
        w = rng.integers(0, 24)
        b[4 * w: 4 * w + 4] = rng.integers(0, 256, 4)
    else:                                             # 64 bytes of noise in the bit stream
        i = rng.integers(0, len(b) - 64)
        b[i: i + 64] = rng.integers(0, 256, 64)
    p = os.path.join(d, "f.xtc")
    b.tofile(p)
    try:
        with xtc.XtcFile(p) as f:
            if f.n_frames > 0 and f.n_atoms < 10 ** 7:
                f.read(0, f.n_frames, n_threads=2)
                try:                                  # the host stage of the device-decode path walks the same bits
                    f.scan()
                except OSError:
                    pass
        ok += 1
    except (OSError, ValueError, RuntimeError):
        pass
print("survived", ok)
'''


@pytest.mark.parametrize("seed", [1, 2])
def test_reader_survives_corrupt_files(seed):
    """Flipped bytes, truncation, corrupt headers and noise in the bit stream end in an error code or in garbage
    coordinates, never in a crash of the host process (the reader runs inside the caller's analysis)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _FUZZ, str(seed), "400"], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    assert "survived" in r.stdout


def test_scan_counts_the_groups_of_a_frame(tmp_path):
    """gorder_xtc_scan (the host stage of the device-decode path, without a GPU): every atom is either the large atom of a
    group or one of the <= 8 small ones behind it; one bookmark per 32 groups; uncompressed frames have no groups; walking two
    frames in one loop (even counts) gives what walking them one by one (count = 1) gives."""
    s = synthetic.s_cg(700)
    xyz, box, _ = s.frames(0, 5)
    xyz[3] += 0.37                                   # frames of different lengths next to each other
    path = str(tmp_path / "t.xtc")
    write_xtc(path, xyz, box)
    with XtcFile(path) as x:
        groups, marks = x.scan()
        assert groups.shape == (5,) and np.all(groups >= x.n_atoms / 9) and np.all(groups <= x.n_atoms)
        np.testing.assert_array_equal(marks, (groups + 31) // 32)
        for k in range(5):
            g1, m1 = x.scan(k, 1)
            assert (g1[0], m1[0]) == (groups[k], marks[k])
        g2, m2 = x.scan(1, 4)
        np.testing.assert_array_equal(g2, groups[1:])
        with pytest.raises(OSError):
            x.scan(3, 3)
        assert x.scan(5, 0)[0].shape == (0,)
    # a gas: the writer widens the small triples until runs form again, so the same bounds hold
    rng = np.random.default_rng(2)
    far = rng.uniform(0, 50, (2, 640, 3)).astype(np.float32)
    write_xtc(path, far, np.full((2, 3), 50, np.float32))
    with XtcFile(path) as x:
        groups, marks = x.scan()
        assert np.all(groups >= 640 / 9) and np.all(groups <= 640)
        np.testing.assert_array_equal(marks, (groups + 31) // 32)
    write_xtc(path, far[:, :7], np.full((2, 3), 50, np.float32))
    with XtcFile(path) as x:
        groups, marks = x.scan()
        assert groups.tolist() == [0, 0] and marks.tolist() == [0, 0]


def test_scan_of_corrupt_frames_next_to_good_ones(tmp_path):
    """A corrupt stream walked beside a good one (two frames per loop) is reported for that frame only."""
    s = synthetic.s_cg(300)
    xyz, box, _ = s.frames(0, 4)
    path = str(tmp_path / "t.xtc")
    write_xtc(path, xyz, box)
    raw = bytearray(open(path, "rb").read())
    with XtcFile(path) as x:
        good, _ = x.scan()
    per = len(raw) // 4
    rng = np.random.default_rng(0)
    hit = 0
    for trial in range(30):
        b = bytearray(raw)
        at = per + 200 + int(rng.integers(0, per - 400))          # somewhere inside the stream of frame 1
        for k in range(24):
            b[at + k] = int(rng.integers(0, 256))
        open(path, "wb").write(b)
        with XtcFile(path) as x:
            if x.n_frames != 4:
                continue
            L = __import__("gorder_b200._lib", fromlist=["lib"]).lib()
            ng, nb = np.zeros(4, np.int32), np.zeros(4, np.int32)
            rc = L.gorder_xtc_scan(x._x, 0, 4, ng.ctypes.data, nb.ctypes.data)
            assert ng[0] == good[0] and ng[2] == good[2] and ng[3] == good[3]
            if rc:
                assert ng[1] == -1
                hit += 1
    assert hit > 0
