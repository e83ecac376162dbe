"""BASELINE configs[4] (S-VES): CGOrder on a vesicle with dynamic local membrane normals (cell-list PCA around every lipid's
PO4, r = 2 nm) and spherical-clustering leaflets, against the oracle at a size it finishes in seconds.  Reference:
examples/coarse_grained/3_vesicles.yaml; tests_cg.rs:3391-3417 (dynamic normals) and the SphericalClustering variants after it
(their fixtures need vesicle.tpr / vesicle.xtc, absent from the reference tree, so the synthetic vesicle stands in)."""
import numpy as np
import pytest

from gorder_b200 import abi, synthetic

from parity import assert_raw_parity, run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["once", "every", "no_leaflets", "cells"])
def test_vesicle_against_oracle(variant, monkeypatch):
    kw = dict(timewise=True, collect_leaflets=True)
    n = 3000
    if variant == "every":
        kw.update(leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=2)
    elif variant == "no_leaflets":
        kw.update(leaflet_mode=abi.LEAFLET_NONE, collect_leaflets=False, collect_normals=True)
    elif variant == "cells":
        monkeypatch.setenv("GORDER_CELL_MIN_HEADS", "64")   # the cell-list neighbour search (default above 2048 heads)
        n = 2500
    s = synthetic.s_ves(n, **kw)
    xyz, box, idx = s.frames(0, 5)
    g, r = run_both(s.setup, xyz, box, idx, batches=2, oracle_threads=8)
    assert_raw_parity(g, r, s.setup, what=f"vesicle {variant}")
    if s.setup.leaflet_mode == abi.LEAFLET_SPHERICAL:   # outer leaflet = upper, from the clustering alone
        np.testing.assert_array_equal(g.leaflets[0].astype(bool), s.is_outer)
    # normals point along the radius: S of the tail bonds is high although the vesicle has every orientation
    tails = [i for i, nm in enumerate(s.setup.moltypes[0].bond_names) if nm.startswith(("C", "D"))]
    mean = g.sum[tails, 0] / g.count[tails, 0] / 1e6
    assert np.all(mean > 0.5), mean
