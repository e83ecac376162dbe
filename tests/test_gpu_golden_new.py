"""Reference fixtures added after the last GPU session of round 1 (tests/test_oracle_pins.py::*_CASES_NEW): pinned with the
oracle on the CPU, not yet run on a device.  Opt-in so that an untested expectation cannot stop the verified suite:
    GORDER_NEW_GPU_CASES=1 python -m pytest tests/test_gpu_golden_new.py -m gpu
Once green on a B200 the names move into AA_FULL_CASES / CG_FULL_CASES / UA_YAML_CASES and this file goes away."""
import os

import numpy as np
import pytest

from gorder_b200 import SystemTopology, abi

import golden_cases as gc
from parity import assert_raw_parity, run_both
from test_gpu_golden import _check_full
from test_oracle_pins import AA_FULL_CASES_NEW, CG_FULL_CASES_NEW, UA_YAML_CASES_NEW

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("GORDER_NEW_GPU_CASES"), reason="not yet verified on a device: set GORDER_NEW_GPU_CASES=1")]


def _expect_error(which, name):
    setup, xyz, box, fi, case = gc.full_case(which, name)
    eng = SystemTopology(setup)
    with pytest.raises(abi.GorderError) as e:
        eng.analyze_frames(xyz, box, fi)
        eng.finish()
    eng.close()
    assert e.value.code == case["expect_error"]


@pytest.mark.parametrize("name", AA_FULL_CASES_NEW)
def test_aa_new_cases(name):
    _check_full("aa", name, 2)


@pytest.mark.parametrize("name", CG_FULL_CASES_NEW)
def test_cg_new_cases(name):
    if "expect_error" in gc.full_case("cg", name)[4]:
        _expect_error("cg", name)
    else:
        _check_full("cg", name, 3)


@pytest.mark.parametrize("name", UA_YAML_CASES_NEW)
def test_ua_new_cases(name):
    setup, xyz, box, fi, case = gc.ua_case(name)
    g, r = run_both(setup, xyz, box, fi, batches=2, oracle_threads=8)
    assert_raw_parity(g, r, setup, what=name)
    gc.assert_matches_yaml(g, setup, case)
