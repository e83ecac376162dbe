"""Pins of the CPU oracle against the reference's own golden vectors, known-answer tests and
fixtures.  CPU only.  Every expected number below comes from the reference tree:

  calc_sch known answer                      src/analysis/mod.rs:94-105
  UA hydrogen positions (18 coordinates)     src/analysis/uaorder.rs:1113-1200      (<= 1 ulp)
  CG single-frame sums + leaflet counts      src/analysis/cgorder.rs:188-302
  AA single-frame sums + leaflet counts      src/analysis/aaorder.rs:226-415
  OrderValue rounding / integer division     src/analysis/order.rs:21-41; converter.rs:787-794
  block error 0.0514468, prefix averages     src/analysis/timewise.rs:595-648
  cuboid construct_shape / inside            src/analysis/geometry.rs:527-664
  UA end-to-end YAML / maps / leaflets / normals fixtures   tests/files/ua_order_*.yaml, ordermaps_ua/, ...
"""
import json
import os

import numpy as np
import pytest

from gorder_b200 import abi, results
from oracle import oracle

import golden_cases as gc


def test_calc_sch_known_answer():
    v = oracle.vector_to([1.7, 2.1, 9.7], [1.9, 2.4, 0.8], [10, 10, 10])
    assert abs(oracle.calc_sch(v, [0, 0, 1]) - 0.8544775) < 1e-6


def test_ua_hydrogens_bit_exact():
    d = json.load(open(os.path.join(gc.GOLDEN, "ua_hydrogens.json")))
    box = np.array([np.frombuffer(bytes.fromhex(h), np.float32)[0] for h in d["box_hex"]], np.float32)
    for name, c in d["cases"].items():
        atoms = np.array([[np.frombuffer(bytes.fromhex(h), np.float32)[0] for h in a] for a in c["atoms_hex"]], np.float32)
        h3 = atoms[3] if len(atoms) > 3 else None
        got = oracle.predict_hydrogens(c["kind"], atoms[0], atoms[1], atoms[2], h3, box)
        gold = np.array(c["gold"], np.float32)
        ulps = np.abs(got.view(np.int32).astype(np.int64) - gold.view(np.int32).astype(np.int64))
        assert ulps.max() <= 1, (name, ulps)   # the reference asserts relative 1.2e-7 (= 1 ulp); we hit 0


def test_min_image_variant_is_pinned():
    """The alternatives to the frozen fold / wrap do NOT reproduce the hydrogen goldens."""
    import ctypes as C
    L = oracle.lib()
    vm = C.c_int.in_dll(L, "gorder_oracle_variant_minimage")
    d = json.load(open(os.path.join(gc.GOLDEN, "ua_hydrogens.json")))
    box = np.array([np.frombuffer(bytes.fromhex(h), np.float32)[0] for h in d["box_hex"]], np.float32)

    def worst():
        w = 0
        for c in d["cases"].values():
            atoms = np.array([[np.frombuffer(bytes.fromhex(h), np.float32)[0] for h in a] for a in c["atoms_hex"]], np.float32)
            got = oracle.predict_hydrogens(c["kind"], atoms[0], atoms[1], atoms[2], atoms[3] if len(atoms) > 3 else None, box)
            gold = np.array(c["gold"], np.float32)
            w = max(w, int(np.abs(got.view(np.int32).astype(np.int64) - gold.view(np.int32).astype(np.int64)).max()))
        return w

    assert vm.value == 1 and worst() == 0
    vm.value = 0
    try:
        assert worst() > 1
    finally:
        vm.value = 1


def test_order_free_group_center_agrees_with_the_sequential_fold():
    """group_center: engine and oracle add the terms of the refined Bai-Breen centre as fixed-point integers
    and seed it with polynomial sin / cos / atan, so that the result does not depend on the order of the atoms.  The previous
    evaluation (sequential f32 folds, libm: what groan most likely does) stays available as variant 1: on the reference's own
    AA and CG trajectories both give the same centre to a few ulp -- the sequential fold itself is only good to ~1e-6 -- and
    the leaflet fixtures are reproduced bit for bit either way (the whole module runs with variant 0)."""
    import ctypes as C
    L = oracle.lib()
    vc = C.c_int.in_dll(L, "gorder_oracle_variant_center")
    assert vc.value == 0
    rng = np.random.default_rng(3)
    for which in ("aa", "cg"):
        setup, xyz, box, fi, case = gc.full_case(which, "leaflets_global")
        mem = np.asarray(setup.membrane)
        for f in (0, len(xyz) // 2, len(xyz) - 1):
            a = oracle.group_center(xyz[f], mem, box[f])
            b = oracle.group_center(xyz[f], rng.permutation(mem), box[f])
            np.testing.assert_array_equal(a, b)               # order-free
            vc.value = 1
            try:
                c = oracle.group_center(xyz[f], mem, box[f])
            finally:
                vc.value = 0
            assert np.all(np.abs(a - c) <= 2e-6 * box[f]), (which, f, a, c)
    # no PBC: the naive mean, same fixed-point sum
    pts = rng.random((1000, 3)).astype(np.float32) * 7
    a = oracle.group_center(pts, np.arange(1000), np.zeros(3), pbc=False)
    np.testing.assert_allclose(a, pts.astype(np.float64).mean(axis=0), atol=1e-6)
    assert np.all(np.isnan(oracle.group_center(np.full((4, 3), np.nan, np.float32), np.arange(4), np.ones(3))))


@pytest.mark.parametrize("name,counts", [("cg_single_frame", ([242, 242, 24], [121, 121, 12], [121, 121, 12])),
                                         ("aa_single_frame", ([131, 128, 15], [65, 64, 8], [66, 64, 7]))])
def test_single_frame_goldens(name, counts):
    setup, xyz, box, exp = gc.single_frame(name)
    o = oracle.Oracle(setup, n_threads=2)
    o.analyze_frames(xyz, box)
    r = o.finish()
    o.close()
    for col, key in enumerate(("total", "upper", "lower")):
        got = r.sum[:, col] / 1e6
        # the goldens are f32 prints: tolerance = the reference's own (epsilon 1e-5, relative f32 eps)
        tol = np.maximum(1e-5, np.abs(exp[key]) * 1.2e-7)
        assert np.all(np.abs(got - exp[key]) <= tol), (key, np.abs(got - exp[key]).max())
        for (s0, n), c in zip(setup.slot_ranges(), counts[col]):
            assert np.all(r.count[s0:s0 + n, col] == c)


def test_order_value_and_integer_division():
    assert oracle.order_value(0.1234565) == round(float(np.float32(0.1234565)) * 1e6)
    assert oracle.order_value(-0.5) == -500000 and oracle.order_value(1.0) == 1000000
    assert oracle.order_value(np.float32(2.5e-7)) == 0 and oracle.order_value(np.float32(5.000001e-7)) == 1
    # converter.rs:787-794: AnalysisOrder::new(45.32, 56) -> 0.8092857 (UA reports the negative)
    total = oracle.order_value(45.32)
    assert total == 45320000
    assert abs(oracle.calc_order(total, 56) - 0.8092857) < 1e-6
    assert results.calc_order(total, 56) == pytest.approx(0.8092857, abs=1e-6)
    # integer division truncates toward zero (order.rs:34-41)
    assert oracle.calc_order(-7, 2) == pytest.approx(-3e-6, abs=1e-12) and results.calc_order(-7, 2) == pytest.approx(-3e-6, abs=1e-12)
    assert np.isnan(oracle.calc_order(10, 3, 5))


def test_block_error_and_prefix_average_known_answers():
    # timewise.rs:594-616: estimate_error(5) == 0.0514468 (f32 relative eq)
    order = [10.0, 15.0, 18.0, 12.0, 14.0, 15.0, 16.0, 20.0, 21.0, 18.0, 9.0, 11.0, 13.0, 14.0, 19.0, 16.0, 17.0]
    samples = np.array([10, 12, 15, 11, 13, 11, 11, 17, 18, 15, 8, 10, 12, 13, 17, 14, 15], np.uint64)
    sums = np.array([oracle.order_value(x) for x in order], np.int64)
    e = oracle.estimate_error(sums, samples, 5)
    assert e == pytest.approx(0.0514468, abs=1.2e-7)   # assert_relative_eq! default: absolute f32::EPSILON
    assert results.estimate_error(sums, samples, 5) == pytest.approx(0.0514468, abs=1.2e-7)
    assert oracle.estimate_error(np.zeros(0, np.int64), np.zeros(0, np.uint64), 5) is None
    # timewise.rs:624-647: prefix averages
    order = [10.0, 12.0, 15.0, 10.0, 9.0, 12.0, 98432.0]
    samples = np.array([13, 15, 20, 12, 11, 14, 98432], np.uint64)
    sums = np.array([oracle.order_value(x) for x in order], np.int64)
    expected = [0.769230769, 0.785714286, 0.770833333, 0.783333333, 0.788732394, 0.8, 0.999827441]
    np.testing.assert_allclose(oracle.prefix_average(sums, samples), expected, atol=1e-5)
    np.testing.assert_allclose(results.prefix_average(sums, samples), expected, atol=1e-5)


def test_cuboid_shape_construction():
    # geometry.rs:527-604
    s = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=1, moltypes=[], geom_kind=abi.GEOM_CUBOID,
                        geom_dims=(2.5, 3.1, -1.5, 3.5, -1.0, 1.0))
    o = oracle.shape_origin(s, [0, 0, 0], [10, 6, 8])
    np.testing.assert_allclose(o[:6], [2.5, 4.5, 7.0, 0.6, 5.0, 2.0], rtol=2e-7)
    o = oracle.shape_origin(s, [8.0, 5.5, 2.0], [10, 6, 8])
    np.testing.assert_allclose(o[:3], [0.5, 4.0, 1.0], rtol=2e-7)
    s.geom_dims = (2.5, 3.1, float("-inf"), float("inf"), -1.0, 1.0)
    o = oracle.shape_origin(s, [15.0, 5.5, 1.0], [10, 6, 8])
    np.testing.assert_allclose(o[:3], [7.5, 0.0, 0.0], rtol=2e-7)
    assert np.isinf(o[4])
    assert oracle.shape_inside(s, [15.0, 5.5, 1.0], [7.8, -124.4, 1.5], [10, 6, 8])
    assert oracle.shape_inside(s, [15.0, 5.5, 1.0], [7.8, 124.4, 1.5], [10, 6, 8])


def test_cuboid_inside_random():
    # geometry.rs:607-664: PBC-aware predicate equals the naive open-interval one for points in the box
    rng = np.random.default_rng(1288746347198273 % (2 ** 32))
    for i in range(40):
        lo, hi = np.sort(rng.uniform(0, 10, (2, 3)), axis=0)
        dims = [lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]]
        if i % 8 == 0:
            dims[0:2] = [float("-inf"), float("inf")]
        if i % 6 == 0:
            dims[4:6] = [float("-inf"), float("inf")]
        s = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=1, moltypes=[], geom_kind=abi.GEOM_CUBOID, geom_dims=tuple(float(x) for x in dims))
        for _ in range(100):
            p = rng.uniform(0, 10, 3).astype(np.float32)
            naive = all(np.float32(dims[2 * a]) < p[a] < np.float32(dims[2 * a + 1]) for a in range(3))
            assert oracle.shape_inside(s, [0, 0, 0], p, [10, 10, 10]) == naive


UA_YAML_CASES = ["basic", "leaflets_global", "leaflets_individual", "leaflets_local", "error", "error_leaflets", "begin_end_step",
                 "cylinder_center", "cuboid_point", "dynamic_normals", "basic_saturated", "basic_unsaturated", "leaflets_flipped",
                 "manual_normals"]


@pytest.mark.parametrize("name", UA_YAML_CASES)
def test_ua_trajectory_fixtures(name):
    """51-frame Berger POPC/POPS trajectory: the oracle reproduces the reference's YAML outputs."""
    setup, xyz, box, fi, case = gc.ua_case(name)
    o = oracle.Oracle(setup, n_threads=8)
    o.analyze_frames(xyz, box, fi)
    raw = o.finish()
    o.close()
    gc.assert_matches_yaml(raw, setup, case)
    if name == "dynamic_normals":   # ua_normals.yaml: signed components at 1e-5 (tests/common/mod.rs:55-91); sign is nalgebra's
        for (mt, (m0, n)) in zip(setup.moltypes, _mol_ranges(setup)):
            exp = np.array(case["normals"][mt.name], np.float32)
            got = raw.normals[:, m0:m0 + n, :]
            assert exp.shape == got.shape
            # SIGNED components, as the reference's comparator (tests/common/mod.rs:84-87, epsilon 1e-5): every one of the
            # 51 x 128 normals has the reference's sign; 99.5 % agree to 1e-5, the rest -- clouds of a few heads, where one ulp
            # of a re-imaged position (the neighbour order of groan's CellGrid is not known) is amplified -- to 2e-4, with the
            # f32 SVD restatement and with an f64 eigen-decomposition of the same cloud alike
            diff = np.abs(exp - got).max(axis=-1)
            assert np.all(np.sum(exp * got, axis=-1) > 0.99), mt.name
            assert np.mean(diff < 1e-5) > 0.99 and float(diff.max()) < 2e-4, (mt.name, float(diff.max()), float(np.mean(diff < 1e-5)))


def test_signed_normals_of_the_reference_unit_test():
    """membrane_normal_from_cloud (normal.rs:421-458) takes the last row of V^T of nalgebra's SVD, sign included.  The oracle
    restates that algorithm (oracle/gorder_oracle.c nalgebra_svd_last_row); here it reproduces the 274 signed normals of the
    reference's own unit test (normal.rs:664-963: pcpepg.tpr, the P atoms within 2 nm of every P atom, positions NOT re-imaged)
    -- tests/golden/normals_planar.npz holds the P positions, the box and the expected vectors (make_golden.py)."""
    d = np.load(os.path.join(gc.GOLDEN, "normals_planar.npz"))
    pts, L, exp = d["P"], d["box"], d["expected"]
    assert exp.shape == (274, 3)
    worst = 0.0
    for i in range(len(exp)):
        dd = pts - pts[i]
        dd -= np.round(dd / L) * L                      # groan's Sphere filter measures the distance through the box
        got = oracle.normal_from_cloud(pts[np.linalg.norm(dd, axis=1) < 2.0])
        worst = max(worst, float(np.abs(got - exp[i]).max()))
    assert worst < 5e-6, worst
    # the eigenvector route (variant 1) gives the same directions with arbitrary signs
    import ctypes as C
    vn = C.c_int.in_dll(oracle.lib(), "gorder_oracle_variant_normal")
    vn.value = 1
    try:
        flipped = 0
        for i in range(0, len(exp), 7):
            dd = pts - pts[i]
            dd -= np.round(dd / L) * L
            got = oracle.normal_from_cloud(pts[np.linalg.norm(dd, axis=1) < 2.0])
            assert abs(abs(float(np.dot(got, exp[i]))) - 1) < 1e-5
            flipped += float(np.dot(got, exp[i])) < 0
        assert flipped > 0
    finally:
        vn.value = 0


def _mol_ranges(setup):
    out, m = [], 0
    for mt in setup.moltypes:
        out.append((m, mt.n_molecules))
        m += mt.n_molecules
    return out


def test_ua_leaflets_once_export_bit_exact():
    setup, xyz, box, fi, case = gc.ua_case("leaflets_once_export")
    o = oracle.Oracle(setup, n_threads=4)
    o.analyze_frames(xyz, box, fi)
    raw = o.finish()
    o.close()
    assert raw.leaflets.shape[0] == 1
    for mt, (m0, n) in zip(setup.moltypes, _mol_ranges(setup)):
        np.testing.assert_array_equal(raw.leaflets[0, m0:m0 + n], np.array(case["leaflets"][mt.name][0], np.uint8))


def test_ua_ordermaps_fixture():
    setup, xyz, box, fi, case = gc.ua_case("maps_basic")
    o = oracle.Oracle(setup, n_threads=8)
    o.analyze_frames(xyz, box, fi)
    raw = o.finish()
    o.close()
    check_maps(raw, setup, case)


def check_maps(raw, setup, case):
    """ordermaps_ua/ordermap_<atom>--<H>_full.dat: x y value rows, x-major; NaN below min_samples."""
    nx, ny = raw.map_shape
    # slots in order: per carbon (sorted by relative index), per hydrogen
    labels = []
    mt = setup.moltypes[0]
    for i, kind in enumerate(mt.ua_kind):
        atom = mt.bond_names[i].split()   # "POPC C13 (12)"
        for h in range(abi.ua_hydrogens(kind)):
            labels.append(f"ordermap_{atom[0]}-{atom[1]}-{atom[2].strip('()')}--{atom[0]}-H{h + 1}-{atom[2].strip('()')}_full.dat")
    assert set(labels) == set(case["maps"].keys()), (labels, list(case["maps"].keys()))
    for s, lab in enumerate(labels):
        rows = np.array(case["maps"][lab], np.float64)
        assert rows.shape[0] == nx * ny, (rows.shape, nx, ny)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = -(raw.map_sum[s, 0].astype(np.float64) / 1e6) / raw.map_count[s, 0].astype(np.float64)
        val = np.where(raw.map_count[s, 0] < case["map_min_samples"], np.nan, val).reshape(-1)
        np.testing.assert_allclose(val, rows[:, 2], atol=gc.FIXTURE_TOL, rtol=0, equal_nan=True, err_msg=lab)
        # node coordinates: min + i * bin (GridMap is node-centred)
        xs = np.repeat(np.arange(nx) * setup.map_bin[0], ny)
        ys = np.tile(np.arange(ny) * setup.map_bin[1], nx)
        np.testing.assert_allclose(rows[:, 0], xs, atol=1e-3)
        np.testing.assert_allclose(rows[:, 1], ys, atol=1e-3)


def test_aa_trajectory_fixture():
    """AA end to end (tests_aa.rs:1019-1040): pcpepg_selected.xtc, 4 frames, 229 C-H bond types in POPE / POPC / POPG,
    Global leaflets -> aa_order_selected.yaml (total / upper / lower of every atom and bond, reference tolerance)."""
    setup, xyz, box, case = gc.aa_traj()
    assert [m.name for m in setup.moltypes] == case["molecules"]
    o = oracle.Oracle(setup, n_threads=8)
    o.analyze_frames(xyz, box, np.arange(xyz.shape[0], dtype=np.int64))
    raw = o.finish()
    o.close()
    gc.assert_matches_yaml(raw, setup, case)


# ---- the reference's full AA / CG test trajectories (re-joined from tests/files/split/*) against its YAML fixtures ----
AA_FULL_CASES = ["basic", "leaflets_global", "leaflets_individual", "leaflets_local", "leaflets_every5", "leaflets_once", "error",
                 "error_leaflets", "begin_end", "begin_end_step", "limit", "leaflets_limit", "sphere_center", "maps_basic",
                 "maps_cuboid_square", "maps_cylinder", "cuboid_dynamic", "cylinder_dynamic", "sphere_dynamic", "sphere_dynamic_inverted", "cuboid_patch",
                 "cylinder_x", "cylinder_z_inverted", "cuboid_square_inverted", "leaflets_dynamic", "export_once_global", "export_every5_local",
                 "export_every1_individual", "export_every1_global", "error_blocks10", "step5_leaflets", "convergence", "convergence_leaflets",
                 "maps_leaflets", "error_limit", "error_leaflets_limit", "sphere_static", "manual_once", "manual_every10", "manual_every",
                 "manual_every10_stepping", "manual_begin_end_step"]
CG_FULL_CASES = ["basic", "leaflets_global", "leaflets_individual", "leaflets_local", "leaflets_every5", "leaflets_once", "error",
                 "error_leaflets", "begin_end_step", "leaflets_dynamic", "cuboid_square", "cylinder", "sphere_dynamic", "cylinder_z_inverted", "limit",
                 "leaflets_limit", "maps_basic", "maps_leaflets", "error_limit", "error_leaflets_limit", "begin_end", "leaflets_only_upper",
                 "leaflets_only_upper_individual", "leaflets_only_upper_local", "redefined_bonds", "manual_once", "manual_every20",
                 "manual_every", "manual_not_enough_frames"]


def check_maps_aa(raw, setup, case):
    """ordermaps*/ordermap_<res>-<A>-<i>--<res>-<B>-<j>_{full,upper,lower}.dat (one per bond type) and, for AA,
    ordermap_<res>-<C>-<i>_*.dat (the heavy atom: its bonds merged sample by sample): x y value rows, x-major; NaN below
    min_samples.  AA / UA report -S, CG reports S."""
    nx, ny = raw.map_shape
    mt = setup.moltypes[0]
    sign = 1.0 if setup.kind == abi.KIND_CG else -1.0
    suffixes = [("full", 0)] + ([("upper", 1), ("lower", 2)] if any(k.endswith("_upper.dat") for k in case["maps"]) else [])
    per_atom = {}
    seen = set()
    vals = {}

    def compare(lab, sm, cn):
        rows = np.array(case["maps"][lab], np.float64)
        assert rows.shape[0] == nx * ny, (rows.shape, nx, ny)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = sign * (sm.astype(np.float64) / 1e6) / cn.astype(np.float64)
        val = np.where(cn < case["map_min_samples"], np.nan, val).reshape(-1)
        np.testing.assert_allclose(val, rows[:, 2], atol=gc.FIXTURE_TOL, rtol=0, equal_nan=True, err_msg=lab)
        xs = np.repeat(np.arange(nx) * setup.map_bin[0], ny)
        ys = np.tile(np.arange(ny) * setup.map_bin[1], nx)
        np.testing.assert_allclose(rows[:, 0], xs, atol=1e-3)
        np.testing.assert_allclose(rows[:, 1], ys, atol=1e-3)
        seen.add(lab)
        vals[lab] = val.reshape(nx, ny)

    for b, name in enumerate(mt.bond_names):    # "POPC C22 (32) - POPC H2R (33)"
        a, h = name.split(" - ")
        fa, fh = ("-".join(x.replace("(", "").replace(")", "").split()) for x in (a, h))
        for suf, k in suffixes:
            compare(f"ordermap_{fa}--{fh}_{suf}.dat", raw.map_sum[b, k], raw.map_count[b, k].astype(np.int64))
            acc = per_atom.setdefault((fa, suf), [np.zeros((nx, ny), np.int64), np.zeros((nx, ny), np.int64)])
            acc[0] += raw.map_sum[b, k]
            acc[1] += raw.map_count[b, k].astype(np.int64)
    if setup.kind == abi.KIND_AA:
        for (fa, suf), (sm, cn) in per_atom.items():
            compare(f"ordermap_{fa}_{suf}.dat", sm, cn)
    assert seen == set(case["maps"].keys()), seen ^ set(case["maps"].keys())
    # the converters (numpy and the shared library's gorder_results_map) produce the same maps: per bond and, for AA, per atom
    for native in (False, True):
        res = results.convert(raw, setup, map_min_samples=case["map_min_samples"], native=native)
        items = next(iter(res.molecules.values())).items
        b = 0
        for it in items:
            per_bond = it.bond_maps if setup.kind != abi.KIND_CG else [it.maps]
            for bm in per_bond:
                a, h = mt.bond_names[b].split(" - ")
                fa, fh = ("-".join(x.replace("(", "").replace(")", "").split()) for x in (a, h))
                for suf, k in suffixes:
                    np.testing.assert_allclose(bm[k], vals[f"ordermap_{fa}--{fh}_{suf}.dat"], atol=1e-6, rtol=0, equal_nan=True)
                    if setup.kind == abi.KIND_AA:
                        np.testing.assert_allclose(it.maps[k], vals[f"ordermap_{fa}_{suf}.dat"], atol=1e-6, rtol=0, equal_nan=True)
                b += 1
        assert b == len(mt.bond_names)


def check_convergence(raw, setup, case):
    """*_convergence.xvg: frame number, then the prefix average of every molecule type (total [, upper, lower] blocks)."""
    res = results.convert(raw, setup, n_blocks=case.get("n_blocks"))
    exp = np.array(case["convergence"], np.float64)
    keys = case["keys"]
    cols = [np.asarray(m.convergence[k], np.float64) for m in res.molecules.values() for k in keys]   # molecule-major
    got = np.stack(cols, axis=1)
    assert exp.shape == (got.shape[0], 1 + got.shape[1]), (exp.shape, got.shape)
    np.testing.assert_array_equal(exp[:, 0], np.arange(1, got.shape[0] + 1))
    np.testing.assert_allclose(got, exp[:, 1:], atol=gc.FIXTURE_TOL, rtol=0)
    nat = results.convert(raw, setup, n_blocks=case.get("n_blocks"), native=True)
    for name, m in res.molecules.items():
        for k in keys:
            np.testing.assert_array_equal(nat.molecules[name].convergence[k], m.convergence[k], err_msg=f"native prefix average {name} {k}")


def check_leaflet_export(raw, setup, case):
    """aa_leaflets_*.yaml: one row of 0 / 1 per assignment frame and molecule type -- bit exact."""
    for mt, (m0, n) in zip(setup.moltypes, _mol_ranges(setup)):
        exp = np.array(case["leaflets"][mt.name], np.uint8)
        assert raw.leaflets.shape[0] == exp.shape[0], (raw.leaflets.shape, exp.shape)
        np.testing.assert_array_equal(raw.leaflets[:, m0:m0 + n], exp, err_msg=f"{case['leaflet_source']} {mt.name}")


def _oracle_full(which, name):
    setup, xyz, box, fi, case = gc.full_case(which, name)
    o = oracle.Oracle(setup, n_threads=8)
    if "expect_error" in case:   # e.g. "could not get leaflet assignment for frame" (leaflets.rs:816-874)
        with pytest.raises(abi.GorderError) as e:
            o.analyze_frames(xyz, box, fi)
            o.finish()
        o.close()
        assert e.value.code == case["expect_error"]
        return
    o.analyze_frames(xyz, box, fi)
    raw = o.finish()
    o.close()
    gc.assert_matches_yaml(raw, setup, case)
    if "maps" in case:
        check_maps_aa(raw, setup, case)
    if "leaflets" in case:
        check_leaflet_export(raw, setup, case)
    if "convergence" in case:
        check_convergence(raw, setup, case)


@pytest.mark.parametrize("name", AA_FULL_CASES)
def test_aa_full_trajectory_fixtures(name):
    """pcpepg.xtc (51 frames, 35 432 lipid atoms, 229 C-H bond types): tests_aa.rs:25-45, 289-316, 548-582, 1099-1149,
    1202-1232, 1398-1423, 2170-2247, 3239-3260 -> tests/files/aa_order_*.yaml."""
    assert set(AA_FULL_CASES) == set(gc.full_case_names("aa"))
    _oracle_full("aa", name)


@pytest.mark.parametrize("name", CG_FULL_CASES)
def test_cg_full_trajectory_fixtures(name):
    """cg.xtc (101 frames, 6 096 beads): tests_cg.rs:26-43, 180-213, 746-772, 1367-1435, 3356-3388 -> tests/files/cg_order_*.yaml."""
    assert set(CG_FULL_CASES) == set(gc.full_case_names("cg"))
    _oracle_full("cg", name)


def test_ua_no_pbc_fixture():
    """handle_pbc(false): naive centre of geometry, no minimum image, no wrap of the rebuilt hydrogens (pbc.rs:163-199)."""
    setup, xyz, box, case = gc.ua_nopbc()
    assert not setup.handle_pbc
    o = oracle.Oracle(setup, n_threads=8)
    o.analyze_frames(xyz, box, np.arange(xyz.shape[0], dtype=np.int64))
    raw = o.finish()
    o.close()
    gc.assert_matches_yaml(raw, setup, case)


# ---- spherical clustering: oracle restatement, pinned by the reference's own unit tests ----
def _gmm(data):
    import ctypes as C
    L = oracle.lib()
    L.gorder_oracle_gmm_fit.restype = C.c_float
    L.gorder_oracle_gmm_fit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    data = np.ascontiguousarray(data, np.float32)
    resp, params = np.zeros(len(data), np.float32), np.zeros(5, np.float32)
    ll = L.gorder_oracle_gmm_fit(data.ctypes.data, len(data), resp.ctypes.data, params.ctypes.data)
    return resp, params, ll


def test_spherical_clusters_from_responsibilities_known_answer():
    """spherical_clustering.rs:299-316: r < 0.5 -> one cluster, the one farther from the centre is the upper (outer) leaflet."""
    import ctypes as C
    L = oracle.lib()
    L.gorder_oracle_gmm_clusters.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    resp = np.array([0.9998, 0.1, 0.42, 0.834, 0.932], np.float32)
    dist = np.array([10.5, 1.3, 2.8, 7.8, 8.4], np.float32)
    upper = np.zeros(5, np.uint8)
    L.gorder_oracle_gmm_clusters(resp.ctypes.data, dist.ctypes.data, 5, upper.ctypes.data)
    assert upper.tolist() == [1, 0, 0, 1, 1]
    # the other orientation: cluster 1 (r < 0.5) is the outer one
    L.gorder_oracle_gmm_clusters(resp.ctypes.data, (12 - dist).astype(np.float32).ctypes.data, 5, upper.ctypes.data)
    assert upper.tolist() == [0, 1, 1, 0, 0]


@pytest.mark.parametrize("seed", [424242, 67676767, 12345678, 1111111, 999999])
def test_spherical_gmm_separates_two_normals(seed):
    """spherical_clustering.rs:346-361 (test_fit_gmm): 50 points from N(5, 1) / N(20, 2), p = 0.5: component A
    (initialised from the 25th percentile) takes the points near 5.  Same property, numpy's generator."""
    rng = np.random.default_rng(seed)
    pick = rng.random(50) < 0.5
    data = np.where(pick, rng.normal(5.0, 1.0, 50), rng.normal(20.0, 2.0, 50)).astype(np.float32)
    resp, params, ll = _gmm(data)
    assert np.all(resp[pick] > 0.5) and np.all(resp[~pick] < 0.5)   # (the components are > 5 sigma apart)
    assert params[1] == pytest.approx(5.0, abs=0.6) and params[3] == pytest.approx(20.0, abs=1.2) and np.isfinite(ll)
    # against an independent float64 EM with the same initialisation and stopping rule
    x = data.astype(np.float64)
    srt = np.sort(x)
    ma, mb, va, vb, w, prev = srt[len(x) // 4], srt[3 * len(x) // 4], x.var(ddof=1), x.var(ddof=1), 0.5, -np.inf
    for _ in range(50):
        ja = np.log(w) - 0.5 * (np.log(2 * np.pi) + np.log(va) + (x - ma) ** 2 / va)
        jb = np.log(1 - w) - 0.5 * (np.log(2 * np.pi) + np.log(vb) + (x - mb) ** 2 / vb)
        lp = np.logaddexp(ja, jb)
        r = np.exp(ja - lp)
        if abs(lp.mean() - prev) < 1e-4:
            break
        prev = lp.mean()
        sa, sb = r.sum(), len(x) - r.sum()
        w = min(max(sa / len(x), 1e-4), 1 - 1e-4)
        ma, mb = (r * x).sum() / sa, ((1 - r) * x).sum() / sb
        va, vb = max((r * (x - ma) ** 2).sum() / sa, 1e-6), max(((1 - r) * (x - mb) ** 2).sum() / sb, 1e-6)
    np.testing.assert_allclose(resp, r, atol=2e-4)
    np.testing.assert_allclose(params, [w, ma, va, mb, vb], rtol=2e-4)


def test_spherical_leaflets_on_a_vesicle():
    """Vesicle across the periodic boundary: the outer leaflet is `upper`, whatever the shift; `flip` swaps them."""
    rng = np.random.default_rng(5)
    n_out, n_in = 420, 260

    def sphere(n, r):
        v = rng.normal(size=(n, 3))
        return (v / np.linalg.norm(v, axis=1)[:, None]) * (r + rng.normal(0, 0.12, (n, 1)))

    heads = np.concatenate([sphere(n_out, 9.0), sphere(n_in, 5.5)])
    outward = heads / np.linalg.norm(heads, axis=1)[:, None]
    is_outer = np.arange(n_out + n_in) < n_out
    tails = heads - outward * np.where(is_outer, 1.0, -1.0)[:, None] * 0.45     # chains point into the bilayer
    L = 30.0
    perm = rng.permutation(n_out + n_in)
    heads, tails, is_outer = heads[perm], tails[perm], is_outer[perm]
    xyz = np.empty((1, 2 * len(heads), 3), np.float32)
    shift = np.array([13.0, -4.0, 29.0])
    xyz[0, 0::2], xyz[0, 1::2] = np.mod(heads + shift, L), np.mod(tails + shift, L)
    box = np.full((1, 3), L, np.float32)
    mt = abi.MolType(name="LIP", mol_base=np.arange(0, 2 * len(heads), 2), bond_rel=[(0, 1)], head_rel=0)
    for flip in (False, True):
        setup = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=xyz.shape[1], moltypes=[mt], leaflet_mode=abi.LEAFLET_SPHERICAL,
                                membrane=np.arange(0, 2 * len(heads), 2), leaflet_flip=flip, collect_leaflets=True)
        o = oracle.Oracle(setup, n_threads=1)
        o.analyze_frames(xyz, box, np.arange(1))
        raw = o.finish()
        o.close()
        np.testing.assert_array_equal(raw.leaflets[0].astype(bool), is_outer ^ flip)
        assert raw.count[0].tolist() == [len(heads), int((is_outer ^ flip).sum()), int((~(is_outer ^ flip)).sum())]


def test_spherical_gmm_parallel_model_with_sequential_folds():
    """The device kernel (csrc/gorder_spherical.cuh) evaluates the terms of every phase in parallel and folds each sum on one
    lane in index order.  numpy model of exactly that (f32 terms, np.cumsum = sequential f32 fold, correctly rounded log / exp
    instead of libm's) against the oracle, on leaflets whose distance distributions overlap: same assignments, |dr| < 1e-4.
    (A tree / f64 reduction stops the EM at another iteration and flips up to 0.1 % of the heads here: measured)"""
    f32 = np.float32
    seq = lambda v: np.cumsum(v.astype(f32), dtype=f32)[-1]                         # noqa: E731
    lg = lambda v: np.log(np.asarray(v, np.float64)).astype(f32)                    # noqa: E731
    ex = lambda v: np.exp(np.asarray(v, np.float64)).astype(f32)                    # noqa: E731

    def model(x):
        n, nf = len(x), f32(len(x))
        srt = np.sort(x)
        ma, mb = srt[n // 4], srt[3 * n // 4]
        gm = seq(x) / nf
        t = (x - gm).astype(f32)
        gv = seq((t * t).astype(f32)) / (nf - f32(1))
        va = vb = max(gv, f32(1e-6))
        w, prev, c2pi = f32(0.5), f32(-np.inf), lg(f32(2) * f32(3.14159274101257324))
        for _ in range(50):
            def lgauss(m, v):
                d = (x - m).astype(f32)
                return (f32(-0.5) * ((c2pi + lg(v)).astype(f32) + ((d * d).astype(f32) / v).astype(f32)).astype(f32)).astype(f32)
            ja, jb = (lg(w) + lgauss(ma, va)).astype(f32), (lg(f32(1) - w) + lgauss(mb, vb)).astype(f32)
            m = np.maximum(ja, jb)
            lp = (m + lg((ex((ja - m).astype(f32)) + ex((jb - m).astype(f32))).astype(f32))).astype(f32)
            r = ex((ja - lp).astype(f32))
            avg = seq(lp) / nf
            if abs(avg - prev) < f32(1e-4):
                break
            prev = avg
            sa = seq(r)
            sb = nf - sa
            sa, sb = max(sa, f32(1e-6)), max(sb, f32(1e-6))
            w = min(max(sa / nf, f32(1e-4)), f32(1) - f32(1e-4))
            ma, mb = seq((r * x).astype(f32)) / sa, seq(((f32(1) - r).astype(f32) * x).astype(f32)) / sb
            da, db = (x - ma).astype(f32), (x - mb).astype(f32)
            va = max(seq(((r * da).astype(f32) * da).astype(f32)) / sa, f32(1e-6))
            vb = max(seq((((f32(1) - r).astype(f32) * db).astype(f32) * db).astype(f32)) / sb, f32(1e-6))
        return r

    rng = np.random.default_rng(1)
    for _ in range(8):
        n_out, n_in = rng.integers(200, 30000), rng.integers(100, 20000)
        gap, sig = rng.uniform(1.0, 4.0), rng.uniform(0.05, 0.5)
        x = np.concatenate([rng.normal(6 + gap, sig, n_out), rng.normal(6, sig, n_in)]).astype(f32)
        rng.shuffle(x)
        resp, _params, _ll = _gmm(x)
        r = model(x)
        assert np.abs(r - resp).max() < 1e-4
        np.testing.assert_array_equal(r < 0.5, resp < 0.5)
