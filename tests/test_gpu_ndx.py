"""Leaflets read from GROMACS index files (LeafletClassification::from_ndx, tests_aa.rs:5210-5360): C++ ndx reader -> manual
leaflet tables -> engine, against the reference's expected output and the oracle."""
import pytest

import golden_cases as gc
from parity import assert_raw_parity, run_both
from test_ndx_cpu import ndx_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which,variant", [("cg", "once"), ("cg", "every"), ("cg", "every10"), ("aa", "once"), ("aa", "every"), ("aa", "every10")])
def test_leaflets_from_ndx(which, variant):
    st, xyz, box, fi, case = ndx_case(which, variant)
    g, r = run_both(st, xyz, box, fi, batches=3, oracle_threads=8)
    assert_raw_parity(g, r, st, what=f"{which} ndx {variant}")
    gc.assert_matches_yaml(g, st, case)
