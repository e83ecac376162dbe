"""Spherical-clustering leaflets on the device (gorder_spherical.cuh; reference spherical_clustering.rs:36-275) against the
oracle's restatement, which is pinned by the reference's unit tests (spherical_clustering.rs:299-361, tests/test_oracle_pins.py)."""
import numpy as np
import pytest

from gorder_b200 import abi

from parity import assert_raw_parity, run_both

pytestmark = pytest.mark.gpu


def vesicle(n_out, n_in, n_frames, seed=5, box=30.0, pbc=True):
    rng = np.random.default_rng(seed)
    n = n_out + n_in
    is_outer = rng.permutation(np.arange(n) < n_out)
    xyz = np.empty((n_frames, 2 * n, 3), np.float32)
    for f in range(n_frames):
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1)[:, None]
        r = np.where(is_outer, 9.0, 5.5)[:, None] + rng.normal(0, 0.12, (n, 1))
        heads = v * r
        tails = heads - v * np.where(is_outer, 1.0, -1.0)[:, None] * 0.45
        shift = rng.uniform(0, box, 3) if pbc else np.full(3, box / 2)
        wrap = (lambda p: np.mod(p + shift, box)) if pbc else (lambda p: p + shift)
        xyz[f, 0::2], xyz[f, 1::2] = wrap(heads), wrap(tails)
    mt = abi.MolType(name="LIP", mol_base=np.arange(0, 2 * n, 2), bond_rel=[(0, 1)], head_rel=0)
    return mt, xyz, np.full((n_frames, 3), box, np.float32), is_outer


@pytest.mark.parametrize("n_out,n_in,pbc,flip", [(420, 260, True, False), (5000, 3100, True, True), (700, 300, False, False)])
def test_vesicle_leaflets_match_oracle(n_out, n_in, pbc, flip):
    mt, xyz, box, is_outer = vesicle(n_out, n_in, 6, pbc=pbc)
    setup = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=xyz.shape[1], moltypes=[mt], leaflet_mode=abi.LEAFLET_SPHERICAL, handle_pbc=pbc,
                            membrane=np.arange(0, xyz.shape[1], 2), leaflet_flip=flip, collect_leaflets=True, timewise=True)
    g, r = run_both(setup, xyz, box, np.arange(6), batches=2)
    assert_raw_parity(g, r, setup, what="spherical clustering")
    np.testing.assert_array_equal(g.leaflets.astype(bool), np.tile(is_outer ^ flip, (6, 1)))


def test_assignment_once_and_every_n():
    mt, xyz, box, is_outer = vesicle(300, 200, 8, seed=9)
    for kw in (dict(leaflet_freq_kind=abi.FREQ_ONCE), dict(leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=3)):
        setup = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=xyz.shape[1], moltypes=[mt], leaflet_mode=abi.LEAFLET_SPHERICAL,
                                membrane=np.arange(0, xyz.shape[1], 2), collect_leaflets=True, **kw)
        g, r = run_both(setup, xyz, box, np.arange(8), batches=3)
        assert_raw_parity(g, r, setup, what=f"spherical clustering {kw}")
