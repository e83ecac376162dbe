"""GPU parity tests proper: CUDA engine (through the C ABI) vs the CPU oracle on seeded inputs."""
import numpy as np
import pytest

from gorder_b200 import abi, synthetic

from parity import assert_raw_parity, run_both

pytestmark = pytest.mark.gpu


def _frames(sys_, n, step=1):
    xyz, box, idx = sys_.frames(0, n, step)
    return xyz, box, idx


@pytest.mark.parametrize("n_lipids,mpt", [(37, 1), (1000, 1), (5000, 2), (5000, 4)])
def test_cg_basic(n_lipids, mpt, monkeypatch):
    monkeypatch.setenv("GORDER_MPT", str(mpt))
    s = synthetic.s_cg(n_lipids, leaflet_mode=abi.LEAFLET_NONE)
    xyz, box, idx = _frames(s, 5)
    g, r = run_both(s.setup, xyz, box, idx)
    assert r.count[:, 0].min() == n_lipids * 5
    assert_raw_parity(g, r, s.setup, what=f"cg basic {n_lipids}")


@pytest.mark.parametrize("mode", [abi.LEAFLET_GLOBAL, abi.LEAFLET_INDIVIDUAL])
@pytest.mark.parametrize("freq", [("every", 1), ("every", 3), ("once", 1)])
def test_cg_leaflets(mode, freq):
    kind, n = freq
    s = synthetic.s_cg(600, leaflet_mode=mode, leaflet_freq_kind=abi.FREQ_ONCE if kind == "once" else abi.FREQ_EVERY,
                       leaflet_freq=n, collect_leaflets=True, timewise=True, split_types=3)
    xyz, box, idx = _frames(s, 8)
    g, r = run_both(s.setup, xyz, box, idx, batches=3)
    assert g.leaflets is not None and g.leaflets.shape[1] == 600
    assert_raw_parity(g, r, s.setup, what=f"cg leaflets {mode} {freq}")


def test_cg_leaflets_flip_nopbc():
    s = synthetic.s_cg(300, leaflet_mode=abi.LEAFLET_GLOBAL, leaflet_flip=True, handle_pbc=False, collect_leaflets=True)
    xyz, box, idx = _frames(s, 4)
    g, r = run_both(s.setup, xyz, box, idx)
    assert_raw_parity(g, r, s.setup, what="flip nopbc")


def test_native_layout_matches_aos():
    s = synthetic.s_cg(700, leaflet_mode=abi.LEAFLET_GLOBAL)
    xyz, box, idx = _frames(s, 4)
    g1, r = run_both(s.setup, xyz, box, idx)
    g2, _ = run_both(s.setup, xyz, box, idx, native=True)
    np.testing.assert_array_equal(g1.sum, g2.sum)
    np.testing.assert_array_equal(g1.count, g2.count)
    assert_raw_parity(g2, r, s.setup, what="native")


def test_aa_basic_and_maps():
    s = synthetic.s_aa(64, n_water=500, leaflet_mode=abi.LEAFLET_GLOBAL, map_enabled=True, map_plane=abi.PLANE_XY,
                       map_bin=(0.5, 0.5))
    s.setup.map_span_x = (0.0, float(s.box[0]))
    s.setup.map_span_y = (0.0, float(s.box[1]))
    xyz, box, idx = _frames(s, 6)
    g, r = run_both(s.setup, xyz, box, idx, batches=2)
    assert_raw_parity(g, r, s.setup, what="aa maps")


@pytest.mark.parametrize("geom", ["cuboid", "cylinder", "sphere", "cylinder_sel", "cuboid_inverted"])
def test_geometry(geom):
    kw = dict(leaflet_mode=abi.LEAFLET_GLOBAL)
    if geom.startswith("cuboid"):
        kw.update(geom_kind=abi.GEOM_CUBOID, geom_ref_kind=abi.GEOMREF_BOX_CENTER,
                  geom_dims=(-3.0, 2.5, float("-inf"), float("inf"), -1.0, 4.0), geom_invert=geom.endswith("inverted"))
    elif geom == "cylinder":
        kw.update(geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_POINT, geom_ref_point=(1.0, 2.0, 3.0),
                  geom_dims=(4.0, float("-inf"), float("inf")), geom_axis=abi.AXIS_Z)
    elif geom == "cylinder_sel":
        kw.update(geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_SELECTION, geom_dims=(3.5, -2.0, 5.0),
                  geom_axis=abi.AXIS_Z)
    else:
        kw.update(geom_kind=abi.GEOM_SPHERE, geom_ref_kind=abi.GEOMREF_BOX_CENTER, geom_dims=(4.5,))
    s = synthetic.s_cg(800, **kw)
    if geom == "cylinder_sel":
        s.setup.geom_ref = np.arange(0, 40 * 12, dtype=np.int32)   # first 40 lipids
    xyz, box, idx = _frames(s, 5)
    g, r = run_both(s.setup, xyz, box, idx)
    assert 0 < r.count[:, 0].min() < 800 * 5, "geometry filter should keep some, not all"
    assert_raw_parity(g, r, s.setup, what=f"geometry {geom}")


@pytest.mark.parametrize("with_sat", [False, True])
def test_ua(with_sat):
    s = synthetic.s_ua(96, with_ch1_sat=with_sat, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
    xyz, box, idx = _frames(s, 6)
    g, r = run_both(s.setup, xyz, box, idx, batches=2)
    assert g.n_slots == (63 if with_sat else 64)
    assert_raw_parity(g, r, s.setup, what="ua")


def test_dynamic_normals():
    s = synthetic.s_cg(400, leaflet_mode=abi.LEAFLET_GLOBAL, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0,
                       collect_normals=True)
    xyz, box, idx = _frames(s, 3)
    g, r = run_both(s.setup, xyz, box, idx)
    assert_raw_parity(g, r, s.setup, what="dynamic normals")


@pytest.mark.parametrize("n_lipids", [400, 6000])
def test_dynamic_normals_cell_list(n_lipids, monkeypatch):
    """Force the cell-list neighbour search (default only for >= 2048 heads) and compare with the oracle's brute force."""
    monkeypatch.setenv("GORDER_CELL_MIN_HEADS", "0")
    s = synthetic.s_cg(n_lipids, leaflet_mode=abi.LEAFLET_GLOBAL, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=1.7,
                       collect_normals=True, geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_BOX_CENTER,
                       geom_dims=(5.0, float("-inf"), float("inf")), geom_axis=abi.AXIS_Z)
    xyz, box, idx = _frames(s, 3)
    g, r = run_both(s.setup, xyz, box, idx, oracle_threads=8)
    assert_raw_parity(g, r, s.setup, what="dynamic normals (cell list)")
    assert np.isnan(r.normals).any() and not np.isnan(r.normals).all()   # lazily computed: NaN outside the cylinder


def test_manual_normals_and_leaflets():
    s = synthetic.s_cg(200, leaflet_mode=abi.LEAFLET_MANUAL, normal_mode=abi.NORMAL_MANUAL)
    rng = np.random.default_rng(5)
    m = s.setup.moltypes[0]
    m.manual_leaflets = rng.integers(0, 2, (4, 200)).astype(np.uint8)
    nrm = rng.normal(size=(4, 200, 3)).astype(np.float32)
    m.manual_normals = nrm
    xyz, box, idx = _frames(s, 4)
    g, r = run_both(s.setup, xyz, box, idx)
    assert_raw_parity(g, r, s.setup, what="manual")


def test_local_leaflets():
    s = synthetic.s_cg(150, leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=2.5, collect_leaflets=True)
    xyz, box, idx = _frames(s, 2)
    g, r = run_both(s.setup, xyz, box, idx)
    assert_raw_parity(g, r, s.setup, what="local leaflets")


def test_determinism():
    s = synthetic.s_cg(3000, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
    xyz, box, idx = _frames(s, 6)
    g1, _ = run_both(s.setup, xyz, box, idx)
    g2, _ = run_both(s.setup, xyz, box, idx, batches=3)
    np.testing.assert_array_equal(g1.sum, g2.sum)
    np.testing.assert_array_equal(g1.tw_sum, g2.tw_sum)


def test_errors():
    from gorder_b200 import SystemTopology
    s = synthetic.s_cg(100, leaflet_mode=abi.LEAFLET_GLOBAL)
    xyz, box, idx = _frames(s, 2)
    bad = xyz.copy()
    bad[1, 5, 1] = np.nan
    eng = SystemTopology(s.setup)
    eng.analyze_frames(bad, box, idx)
    with pytest.raises(abi.GorderError) as e:
        eng.finish()
    assert e.value.code == abi.ERR_UNDEFINED_POSITION and e.value.index == 5
    eng.close()
    eng = SystemTopology(s.setup)
    zb = box.copy()
    zb[0] = 0
    eng.analyze_frames(xyz, zb, idx)
    with pytest.raises(abi.GorderError) as e:
        eng.finish()
    assert e.value.code == abi.ERR_ZERO_BOX
    eng.close()
    # a shard that does not hold the assignment frame
    s2 = synthetic.s_cg(100, leaflet_mode=abi.LEAFLET_GLOBAL, leaflet_freq_kind=abi.FREQ_ONCE)
    eng = SystemTopology(s2.setup)
    with pytest.raises(abi.GorderError) as e:
        eng.analyze_frames(xyz, box, np.array([4, 5]))
    assert e.value.code == abi.ERR_LEAFLET_FRAME_UNAVAILABLE
    eng.close()


def test_edge_cases_empty_and_ragged():
    """Empty submissions, a batch larger than max_batch_frames, per-frame rows growing past their first allocation,
    molecule types of very different sizes (1 molecule .. hundreds)."""
    from gorder_b200 import SystemTopology
    from oracle import oracle as orc
    s = synthetic.s_cg(333, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True, max_batch_frames=4)
    # re-cut the molecule types raggedly: 1, 2, 30 and the rest
    base = list(s.setup.moltypes[0].mol_base)
    proto = s.setup.moltypes[0]
    cuts = [(0, 1), (1, 3), (3, 33), (33, 333)]
    s.setup.moltypes = [abi.MolType(name=f"T{i}", mol_base=base[a:b], bond_rel=proto.bond_rel, head_rel=proto.head_rel,
                                    methyl_rel=proto.methyl_rel, normal_head_rel=proto.normal_head_rel) for i, (a, b) in enumerate(cuts)]
    xyz, box, idx = s.frames(0, 11)
    eng = SystemTopology(s.setup)
    ref = orc.Oracle(s.setup, n_threads=2)
    eng.analyze_frames(xyz[:0], box[:0], idx[:0])          # empty
    eng.analyze_frames(xyz[:9], box[:9], idx[:9])          # 9 frames with max_batch_frames = 4 -> 3 internal batches
    eng.analyze_frames(xyz[9:], box[9:], idx[9:])
    ref.analyze_frames(xyz, box, idx)
    g, r = eng.finish(), ref.finish()
    assert g.n_frames == 11 and g.n_slots == 44
    assert_raw_parity(g, r, s.setup, what="ragged")
    eng.close()
    ref.close()


def test_no_frames_gives_zero_results():
    from gorder_b200 import SystemTopology
    s = synthetic.s_cg(50)
    eng = SystemTopology(s.setup)
    g = eng.finish()
    assert g.n_frames == 0 and not g.sum.any() and not g.count.any()
    eng.close()


def test_large_system_properties():
    """BASELINE-size frame (83 334 lipids): size-independent properties instead of the (slow) oracle:
    counts are exact, upper + lower == total, S in [-0.5, 1], two shards sum to the whole (linearity)."""
    from gorder_b200 import SystemTopology
    s = synthetic.s_cg(83334, leaflet_mode=abi.LEAFLET_GLOBAL)
    xyz, box, idx = s.frames(0, 4)
    whole = SystemTopology(s.setup)
    whole.analyze_frames(xyz, box, idx)
    w = whole.finish()
    whole.close()
    assert np.all(w.count[:, 0] == 83334 * 4)
    np.testing.assert_array_equal(w.count[:, 1] + w.count[:, 2], w.count[:, 0])
    np.testing.assert_array_equal(w.sum[:, 1] + w.sum[:, 2], w.sum[:, 0])
    mean = w.sum[:, 0] / w.count[:, 0] / 1e6
    assert np.all(mean > -0.5) and np.all(mean < 1.0)
    assert abs(int(w.count[0, 1]) - int(w.count[0, 2])) <= 4   # 41 667 lipids per leaflet and frame
    parts = []
    for lo, hi in ((0, 2), (2, 4)):
        e = SystemTopology(s.setup)
        e.analyze_frames(xyz[lo:hi], box[lo:hi], idx[lo:hi])
        parts.append(e.finish())
        e.close()
    np.testing.assert_array_equal(parts[0].sum + parts[1].sum, w.sum)
    np.testing.assert_array_equal(parts[0].count + parts[1].count, w.count)


def test_nan_in_native_frames_is_reported():
    """Frames handed over in the plane layout skip the relayout check: the accumulation kernels must still
    report AnalysisError::UndefinedPosition for a NaN coordinate (bond and UA engines)."""
    from gorder_b200 import SystemTopology
    for s in (synthetic.s_cg(300, leaflet_mode=abi.LEAFLET_NONE), synthetic.s_ua(64, leaflet_mode=abi.LEAFLET_NONE, timewise=False)):
        xyz, box, idx = s.frames(0, 2)
        eng = SystemTopology(s.setup)
        planes = eng.to_native(xyz)
        _, off, cs = eng.native_layout()
        victim = int(s.setup.moltypes[0].mol_base[7]) + (int(s.setup.moltypes[0].bond_rel[0][0]) if s.setup.kind != abi.KIND_UA else int(s.setup.moltypes[0].ua_rel[0][0]))
        planes[1, off[victim] + cs[victim]] = np.nan
        eng.analyze_frames_native(planes, box, idx)
        with pytest.raises(abi.GorderError) as e:
            eng.finish()
        assert e.value.code == abi.ERR_UNDEFINED_POSITION
        eng.close()


@pytest.mark.parametrize("mpt", [2, 4])
@pytest.mark.parametrize("mode,freq", [(abi.LEAFLET_GLOBAL, ("every", 3)), (abi.LEAFLET_GLOBAL, ("once", 1)),
                                       (abi.LEAFLET_INDIVIDUAL, ("every", 1)), (abi.LEAFLET_NONE, ("every", 1))])
def test_fast_kernel_leaflet_tables(mode, freq, mpt, monkeypatch):
    """bond_fast_kernel (full tiles) + the generic body (partial last tile) with leaflets read from the table
    (assignment not on every frame / Individual method) and without leaflets; several batches, per-frame rows."""
    monkeypatch.setenv("GORDER_MPT", str(mpt))
    kind, n = freq
    s = synthetic.s_cg(4500, leaflet_mode=mode, leaflet_freq_kind=abi.FREQ_ONCE if kind == "once" else abi.FREQ_EVERY,
                       leaflet_freq=n, collect_leaflets=mode != abi.LEAFLET_NONE, timewise=True, split_types=2)
    xyz, box, idx = _frames(s, 7)
    g, r = run_both(s.setup, xyz, box, idx, batches=3)
    assert_raw_parity(g, r, s.setup, what=f"fast kernel, leaflet table {mode} {freq} mpt {mpt}")
    monkeypatch.setenv("GORDER_NO_FAST", "1")
    g0, _ = run_both(s.setup, xyz, box, idx, batches=2)
    np.testing.assert_array_equal(g.sum, g0.sum)      # the two kernels agree bit for bit
    np.testing.assert_array_equal(g.tw_sum, g0.tw_sum)
    np.testing.assert_array_equal(g.count, g0.count)


@pytest.mark.parametrize("axis", [abi.AXIS_X, abi.AXIS_Y])
@pytest.mark.parametrize("mpt", [1, 4])
def test_membrane_normal_along_x_or_y(axis, mpt, monkeypatch):
    """Membrane normal / leaflet axis other than z: the kernels read the components in a permuted order."""
    monkeypatch.setenv("GORDER_MPT", str(mpt))
    s = synthetic.s_cg(2100, leaflet_mode=abi.LEAFLET_GLOBAL, collect_leaflets=True)
    s.setup.normal_axis = axis
    s.setup.leaflet_axis = axis
    xyz, box, idx = _frames(s, 4)
    perm = [0, 1, 2]
    perm[axis], perm[2] = 2, axis
    xyz, box = np.ascontiguousarray(xyz[..., perm]), np.ascontiguousarray(box[..., perm])
    g, r = run_both(s.setup, xyz, box, idx)
    assert abs(int(r.leaflets[0].astype(int).sum()) - 1050) <= 2
    assert_raw_parity(g, r, s.setup, what=f"normal axis {axis}")


@pytest.mark.parametrize("n_lipids,axis", [(150, abi.AXIS_Z), (2600, abi.AXIS_Z), (900, abi.AXIS_Y)])
def test_local_leaflets_cell_list(n_lipids, axis, monkeypatch):
    """Local leaflet method with the 2-D cell list over the membrane atoms (default for >= 4096 membrane atoms), against
    the oracle's brute-force cylinder search; also with the membrane normal along y."""
    monkeypatch.setenv("GORDER_LCELL_MIN_ATOMS", "0")
    s = synthetic.s_cg(n_lipids, leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=2.5, collect_leaflets=True, timewise=True)
    s.setup.normal_axis = axis
    s.setup.leaflet_axis = axis
    xyz, box, idx = _frames(s, 3)
    if axis != abi.AXIS_Z:
        perm = [0, 1, 2]
        perm[axis], perm[2] = 2, axis
        xyz, box = np.ascontiguousarray(xyz[..., perm]), np.ascontiguousarray(box[..., perm])
    g, r = run_both(s.setup, xyz, box, idx, batches=2, oracle_threads=8)
    up = r.leaflets.astype(int).sum(axis=1)
    assert np.all(np.abs(up - (n_lipids + 1) // 2) <= 2), up
    assert_raw_parity(g, r, s.setup, what=f"local leaflets, cell list, {n_lipids} lipids, axis {axis}")


def test_ua_streaming_hydrogens_vs_exact(monkeypatch):
    """UA without geometry / maps builds the hydrogens with rsqrt-normalised directions: against the
    bit-exact construction (GORDER_UA_EXACT=1) the per-frame, per-leaflet means move by << 1e-5, counts not at all."""
    from gorder_b200 import SystemTopology
    s = synthetic.s_ua(200, with_ch1_sat=True, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
    xyz, box, idx = _frames(s, 5)
    res = []
    for exact in (False, True):
        if exact:
            monkeypatch.setenv("GORDER_UA_EXACT", "1")
        eng = SystemTopology(s.setup)
        eng.analyze_frames(xyz, box, idx)
        res.append(eng.finish())
        eng.close()
    fast, exact = res
    np.testing.assert_array_equal(fast.count, exact.count)
    np.testing.assert_array_equal(fast.tw_count, exact.tw_count)
    assert not np.array_equal(fast.sum, exact.sum)   # the streaming path is really a different computation
    dev = np.abs(fast.sum - exact.sum) / np.maximum(exact.count, 1) / 1e6
    assert dev.max() < 1e-6, dev.max()
    tw = np.abs(fast.tw_sum - exact.tw_sum) / np.maximum(exact.tw_count, 1) / 1e6
    assert tw.max() < 3e-6, tw.max()


@pytest.mark.parametrize("leaflets", [abi.LEAFLET_NONE, abi.LEAFLET_GLOBAL])
@pytest.mark.parametrize("bad", [np.nan, np.inf])
def test_fast_kernel_reports_undefined_positions(leaflets, bad, monkeypatch):
    """bond_fast_kernel has no per-sample NaN test: an integer max over the bit patterns of |d|^2 must still turn a
    NaN / Inf coordinate of a resident frame into AnalysisError::UndefinedPosition (with and without the
    speculative leaflets, whose sums the same coordinate poisons)."""
    from gorder_b200 import SystemTopology
    monkeypatch.setenv("GORDER_MPT", "2")
    s = synthetic.s_cg(1300, leaflet_mode=leaflets)
    xyz, box, idx = s.frames(0, 3)
    eng = SystemTopology(s.setup)
    planes = eng.to_native(xyz)
    _, off, cs = eng.native_layout()
    victim = int(s.setup.moltypes[0].mol_base[700]) + 5   # a tail bead of a lipid in a full tile
    planes[2, off[victim] + 2 * cs[victim]] = bad
    eng.analyze_frames_native(planes, box, idx)
    with pytest.raises(abi.GorderError) as e:
        eng.finish()
    assert e.value.code == abi.ERR_UNDEFINED_POSITION
    eng.close()


@pytest.mark.parametrize("leaflets", [abi.LEAFLET_NONE, abi.LEAFLET_GLOBAL])
def test_fast_kernel_aa_64_bond_types(leaflets, monkeypatch):
    """AA lipids (64 C-H bond types: the register-resident warp sums are flushed twice per molecule) through
    bond_fast_kernel; with Global leaflets the membrane group holds head-group atoms no bond loads, so the centre comes
    from the pre-pass (inline classification, not speculative)."""
    from gorder_b200 import SystemTopology
    monkeypatch.setenv("GORDER_MPT", "2")
    s = synthetic.s_aa(1100, n_water=200, leaflet_mode=leaflets, timewise=True, collect_leaflets=leaflets != abi.LEAFLET_NONE)
    xyz, box, idx = _frames(s, 3)
    g, r = run_both(s.setup, xyz, box, idx, batches=2, oracle_threads=8)
    assert g.n_slots == 64
    assert_raw_parity(g, r, s.setup, what=f"aa fast kernel, leaflets {leaflets}")
    monkeypatch.setenv("GORDER_NO_FAST", "1")
    g0, _ = run_both(s.setup, xyz, box, idx, oracle_threads=8)
    np.testing.assert_array_equal(g.sum, g0.sum)
    np.testing.assert_array_equal(g.count, g0.count)


# ---- oracle comparisons at BASELINE.json's sizes (the oracle needs seconds for a few frames of them) ----
def test_baseline_size_s_cg_against_oracle():
    """configs[1] S-CG: 83 334 lipids, 1 000 008 beads, Global leaflets every frame (bond_fast_kernel + speculative centre),
    3 frames against the oracle: counts and leaflet tables bit-exact, every sample within one unit of 1e-6."""
    s = synthetic.s_cg(83334, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True, collect_leaflets=True)
    xyz, box, idx = s.frames(0, 3)
    g, r = run_both(s.setup, xyz, box, idx, batches=2, oracle_threads=8)
    assert_raw_parity(g, r, s.setup, what="S-CG full size")
    assert int(g.count[:, 0].sum()) == 916674 * 3


def test_baseline_size_s_aa_large_maps_and_cylinder_against_oracle():
    """configs[3] S-AA-large: 4 096 lipids, 64 C-H bond types, Global leaflets, XY order maps (0.1 nm bins over the box) and
    a cylinder of 8 nm around the box centre: 2 frames against the oracle, maps included (64 x 3 maps of ~130 000 bins)."""
    s = synthetic.s_aa(4096, n_water=0, leaflet_mode=abi.LEAFLET_GLOBAL, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=(0.1, 0.1),
                       geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_BOX_CENTER, geom_dims=(8.0, float("-inf"), float("inf")),
                       geom_axis=abi.AXIS_Z)
    s.setup.map_span_x = (0.0, float(s.box[0]))
    s.setup.map_span_y = (0.0, float(s.box[1]))
    xyz, box, idx = s.frames(0, 2)
    g, r = run_both(s.setup, xyz, box, idx, oracle_threads=8)
    assert_raw_parity(g, r, s.setup, what="S-AA-large full size")
    assert g.map_count.sum() == g.count[:, 0].sum() * 2   # with leaflets: total map = upper + lower, every sample in one bin
    assert 0 < g.count[:, 0].sum() < 4096 * 64 * 2        # the cylinder leaves lipids out


def test_baseline_size_s_ua_error_blocks_against_oracle():
    """configs[2] S-UA: 256 Berger-like lipids (CH3 / CH2 / CH1), per-frame sums for the error blocks, 40 frames: per-frame
    counts bit-exact, per-frame order within 1e-5, and the block errors of the converted results equal."""
    from gorder_b200 import results
    s = synthetic.s_ua(256, timewise=True, leaflet_mode=abi.LEAFLET_GLOBAL)
    xyz, box, idx = s.frames(0, 40)
    g, r = run_both(s.setup, xyz, box, idx, batches=3, oracle_threads=8)
    assert_raw_parity(g, r, s.setup, what="S-UA full size")
    cg_, cr_ = results.convert(g, s.setup, n_blocks=5, native=True), results.convert(r, s.setup, n_blocks=5)
    for key in ("total", "upper", "lower"):
        a, b = getattr(cg_.average, key), getattr(cr_.average, key)
        assert abs(a.value - b.value) < 1e-5 and abs(a.error - b.error) < 1e-5, key


def test_wave_frames_hint():
    """gorder_gpu_wave_frames: tiles x frames is a whole number of waves of the accumulation kernel for the hinted batch."""
    from gorder_b200 import SystemTopology
    import torch
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    for s, per_sm in ((synthetic.s_cg(4096, leaflet_mode=abi.LEAFLET_GLOBAL), 4),                      # K1f, 256-thread CTAs, 4 tiles of 1024
                      (synthetic.s_aa(256, n_water=0), 16),                                            # K1f, 64-thread CTAs, one tile
                      (synthetic.s_cg(4096, leaflet_mode=abi.LEAFLET_GLOBAL, geom_kind=abi.GEOM_SPHERE, geom_ref_kind=abi.GEOMREF_BOX_CENTER,
                                      geom_dims=(3.0,)), 4)):                                           # geometry: two molecules per lane, 8 tiles
        eng = SystemTopology(s.setup)
        f = eng.wave_frames()
        eng.close()
        assert f > 0 and (per_sm * n_sm) % f == 0 or f == per_sm * n_sm, (f, per_sm)
    eng = SystemTopology(synthetic.s_ua(256).setup)
    assert eng.wave_frames() == 0     # persistent kernel: no preference
    eng.close()
