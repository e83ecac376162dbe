"""GPU vs the reference's own goldens (tests/golden, see make_golden.py) and vs the oracle on the
same real-system inputs, through the C ABI."""
import numpy as np
import pytest

from gorder_b200 import abi

import golden_cases as gc
from parity import assert_raw_parity, run_both
from test_oracle_pins import UA_YAML_CASES, _mol_ranges, check_maps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,counts", [("cg_single_frame", ([242, 242, 24], [121, 121, 12], [121, 121, 12])),
                                         ("aa_single_frame", ([131, 128, 15], [65, 64, 8], [66, 64, 7]))])
def test_single_frame_goldens(name, counts):
    setup, xyz, box, exp = gc.single_frame(name)
    g, r = run_both(setup, xyz, box)
    assert_raw_parity(g, r, setup, what=name)
    for col, key in enumerate(("total", "upper", "lower")):
        # reference tolerance 1e-5 on the raw sums + one unit of 1e-6 per sample for cos(acos(c)) -> c
        tol = np.maximum(1e-5, np.abs(exp[key]) * 1.2e-7) + g.count[:, col] * 1e-6
        assert np.all(np.abs(g.sum[:, col] / 1e6 - exp[key]) <= tol), key
        for (s0, n), c in zip(setup.slot_ranges(), counts[col]):
            assert np.all(g.count[s0:s0 + n, col] == c)


@pytest.mark.parametrize("name", UA_YAML_CASES)
def test_ua_trajectory_fixtures(name):
    setup, xyz, box, fi, case = gc.ua_case(name)
    g, r = run_both(setup, xyz, box, fi, batches=2, oracle_threads=8)
    assert_raw_parity(g, r, setup, what=name)
    gc.assert_matches_yaml(g, setup, case)
    if name == "dynamic_normals":
        for mt, (m0, n) in zip(setup.moltypes, _mol_ranges(setup)):
            exp = np.array(case["normals"][mt.name], np.float32)
            # signed components against the reference's ua_normals.yaml (see tests/test_oracle_pins.py for the bars)
            diff = np.abs(exp - g.normals[:, m0:m0 + n, :]).max(axis=-1)
            assert np.all(np.sum(exp * g.normals[:, m0:m0 + n, :], axis=-1) > 0.99), mt.name
            assert np.mean(diff < 1e-5) > 0.99 and float(diff.max()) < 2e-4, (mt.name, float(diff.max()))


def test_ua_leaflets_once_export_bit_exact():
    setup, xyz, box, fi, case = gc.ua_case("leaflets_once_export")
    g, r = run_both(setup, xyz, box, fi, batches=3)
    assert_raw_parity(g, r, setup, what="once export")
    for mt, (m0, n) in zip(setup.moltypes, _mol_ranges(setup)):
        np.testing.assert_array_equal(g.leaflets[0, m0:m0 + n], np.array(case["leaflets"][mt.name][0], np.uint8))


def test_ua_ordermaps_fixture():
    setup, xyz, box, fi, case = gc.ua_case("maps_basic")
    g, r = run_both(setup, xyz, box, fi)
    assert_raw_parity(g, r, setup, what="maps")
    check_maps(g, setup, case)


def test_aa_trajectory_fixture():
    """AA end to end on the GPU: the reference's aa_order_selected.yaml (tests_aa.rs:1019-1040), and the oracle."""
    setup, xyz, box, case = gc.aa_traj()
    g, r = run_both(setup, xyz, box, batches=2, oracle_threads=8)
    assert_raw_parity(g, r, setup, what="aa trajectory")
    gc.assert_matches_yaml(g, setup, case)


from test_oracle_pins import AA_FULL_CASES, CG_FULL_CASES, check_convergence, check_leaflet_export, check_maps_aa  # noqa: E402
from parity import mean_order  # noqa: E402

def _check_full(which, name, batches):
    setup, xyz, box, fi, case = gc.full_case(which, name)
    g, r = run_both(setup, xyz, box, fi, batches=batches, oracle_threads=8)
    # counts bit-exact everywhere, also for geometry selections around the PBC centre of a group: engine and oracle share
    # the order-free centre arithmetic
    assert_raw_parity(g, r, setup, what=f"{which} full {name}")
    gc.assert_matches_yaml(g, setup, case)
    if "maps" in case:
        check_maps_aa(g, setup, case)
    if "leaflets" in case:
        check_leaflet_export(g, setup, case)
    if "convergence" in case:
        check_convergence(g, setup, case)


@pytest.mark.parametrize("name", AA_FULL_CASES)
def test_aa_full_trajectory_fixtures(name):
    """The reference's full AA test trajectory on the GPU: its aa_order_*.yaml / ordermaps fixtures, and the oracle."""
    _check_full("aa", name, 2)


def _expect_error(which, name):
    from gorder_b200 import SystemTopology
    setup, xyz, box, fi, case = gc.full_case(which, name)
    eng = SystemTopology(setup)
    with pytest.raises(abi.GorderError) as e:
        eng.analyze_frames(xyz, box, fi)
        eng.finish()
    eng.close()
    assert e.value.code == case["expect_error"]


@pytest.mark.parametrize("name", CG_FULL_CASES)
def test_cg_full_trajectory_fixtures(name):
    """The reference's full CG test trajectory on the GPU: its cg_order_*.yaml fixtures, and the oracle."""
    if "expect_error" in gc.full_case("cg", name)[4]:   # a leaflet file with too few rows (tests_cg.rs:3054-3074)
        _expect_error("cg", name)
    else:
        _check_full("cg", name, 3)


def test_ua_no_pbc_fixture():
    """handle_pbc(false) on the GPU: ua_order_leaflets_nopbc.yaml (tests_ua.rs:688-714), and the oracle."""
    setup, xyz, box, case = gc.ua_nopbc()
    g, r = run_both(setup, xyz, box, batches=2, oracle_threads=8)
    assert_raw_parity(g, r, setup, what="ua nopbc")
    gc.assert_matches_yaml(g, setup, case)
