"""Dynamic membrane normals (normal.rs:160-199, 421-458; pbc.rs:321-351) through the cell-sorted kernel of round 2
(`dynamic_normal_sorted_kernel`) and its fall-backs, against the oracle's brute force."""
import dataclasses

import numpy as np
import pytest

from gorder_b200 import SystemTopology, abi, synthetic

from parity import assert_raw_parity, run_both

pytestmark = pytest.mark.gpu

DYN = dict(leaflet_mode=abi.LEAFLET_GLOBAL, normal_mode=abi.NORMAL_DYNAMIC, collect_normals=True)


def _frames(s, n):
    return s.frames(0, n)


def _run(setup, xyz, box, idx):
    eng = SystemTopology(setup)
    try:
        eng.analyze_frames(xyz, box, idx)
        return eng.finish()
    finally:
        eng.close()


def test_sorted_walk_equals_molecule_order(monkeypatch):
    """Lanes over the cell-sorted heads (default) vs one lane per lipid in molecule order: same neighbours, same moments."""
    monkeypatch.setenv("GORDER_CELL_MIN_HEADS", "0")
    s = synthetic.s_cg(3000, dynamic_radius=2.0, split_types=3, **DYN)
    xyz, box, idx = _frames(s, 4)
    a = _run(s.setup, xyz, box, idx)
    monkeypatch.setenv("GORDER_NO_SORTED_NORMALS", "1")
    b = _run(s.setup, xyz, box, idx)
    np.testing.assert_array_equal(a.count, b.count)
    assert np.abs(a.sum - b.sum).max() <= 1          # f64 moments: the order of the neighbours does not reach the f32 result
    np.testing.assert_allclose(a.normals, b.normals, atol=2e-7, rtol=0)


@pytest.mark.parametrize("n_lipids,radius", [(60, 1.7), (140, 2.0), (2500, 3.6)])
def test_small_boxes_and_long_lists(n_lipids, radius, monkeypatch):
    """Fewer than three cells along an axis (every cell of the axis is visited once, no image shortcut); a radius that puts
    more heads into a cloud than a lane's list holds (48: the list is drained early)."""
    monkeypatch.setenv("GORDER_CELL_MIN_HEADS", "0")
    s = synthetic.s_cg(n_lipids, dynamic_radius=radius, **DYN)
    xyz, box, idx = _frames(s, 3)
    g, r = run_both(s.setup, xyz, box, idx, oracle_threads=8)
    assert_raw_parity(g, r, s.setup, what=f"normals {n_lipids} lipids, r = {radius}")


def test_cloud_only_heads_and_repeated_members(monkeypatch):
    """NormalHeads may hold heads of lipids that are not analysed (they count in the clouds of the others) and repeated
    members (counted twice in a cloud, as the reference's group iteration would not -- so the host must not pass them; the
    engine takes the molecule-order kernel for such a setup and agrees with the oracle either way)."""
    monkeypatch.setenv("GORDER_CELL_MIN_HEADS", "0")
    s = synthetic.s_cg(1200, dynamic_radius=2.0, split_types=2, **DYN)
    xyz, box, idx = _frames(s, 3)
    only_first = dataclasses.replace(s.setup, moltypes=s.setup.moltypes[:1])   # the second type's heads stay in the group
    g, r = run_both(only_first, xyz, box, idx, oracle_threads=8)
    assert_raw_parity(g, r, only_first, what="cloud-only heads")
    assert g.normals.shape[1] == len(s.setup.moltypes[0].mol_base)
    whole = _run(s.setup, xyz, box, idx)
    n0 = len(s.setup.moltypes[0].mol_base)
    np.testing.assert_allclose(g.normals, whole.normals[:, :n0], atol=2e-7, rtol=0)   # the same clouds as in the full analysis


def test_too_few_points_is_an_error(monkeypatch):
    """DynamicNormalError::NotEnoughPoints (normal.rs:424): a radius that leaves a head alone."""
    monkeypatch.setenv("GORDER_CELL_MIN_HEADS", "0")
    s = synthetic.s_cg(400, dynamic_radius=0.3, **DYN)
    xyz, box, idx = _frames(s, 2)
    with pytest.raises(abi.GorderError) as e:
        _run(s.setup, xyz, box, idx)
    assert e.value.code == abi.ERR_DYNAMIC_NORMAL_POINTS
