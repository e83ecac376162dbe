"""TPR -> C++ reader -> C++ classifier -> engine on the GPU, against the reference's expected output for a system whose
topology exists only as a TPR file (tests_cg.rs:2182-2309: asymmetric CG membrane, cg_order_asymmetric*.yaml)."""
import dataclasses

import pytest

import golden_cases as gc
from parity import assert_raw_parity, run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["leaflets", "errors"])
def test_cg_asym_from_tpr(name):
    setup, xyz, box, cases = gc.cg_asym()
    case = cases[name]
    setup = dataclasses.replace(setup, timewise="n_blocks" in case)
    g, r = run_both(setup, xyz, box, batches=2, oracle_threads=8)
    assert_raw_parity(g, r, setup, what=f"cg_asym {name}")
    gc.assert_matches_yaml(g, setup, case)
    assert g.count[:, 1].sum() != g.count[:, 2].sum()   # the membrane IS asymmetric: 144 + 135 lipids split unevenly
