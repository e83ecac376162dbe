"""Speculative Global leaflets (bond_order_kernel<SPEC> + spec_repair_kernel).

The bond kernel classifies every lipid against a provisional membrane centre while it sums the
membrane's displacements from it; frames whose leaflets are not provably those of the exact
(two-pass, refined Bai-Breen: leaflets.rs:187 -> groan group_get_center) centre are repaired.
Whatever happens, the accumulators must be bit-identical to the non-speculative path and match the
oracle like every other configuration.
"""
import numpy as np
import pytest

from gorder_b200 import SystemTopology, abi, synthetic

from parity import assert_raw_parity, run_both

pytestmark = pytest.mark.gpu


def _run(setup, xyz, box, idx, batches=1, native=False):
    eng = SystemTopology(setup)
    try:
        edges = np.linspace(0, len(idx), batches + 1).astype(int)
        for a, b in zip(edges[:-1], edges[1:]):
            if native:
                eng.analyze_frames_native(eng.to_native(xyz[a:b]), box[a:b], idx[a:b])
            else:
                eng.analyze_frames(xyz[a:b], box[a:b], idx[a:b])
        st = eng.speculation_stats()
        return eng.finish(), st
    finally:
        eng.close()


def _same(a, b):
    np.testing.assert_array_equal(a.sum, b.sum)
    np.testing.assert_array_equal(a.count, b.count)
    if a.leaflets is not None:
        np.testing.assert_array_equal(a.leaflets, b.leaflets)
    if a.tw_sum is not None:
        np.testing.assert_array_equal(a.tw_sum, b.tw_sum)
        np.testing.assert_array_equal(a.tw_count, b.tw_count)


@pytest.mark.parametrize("n_lipids,mpt,split", [(700, 1, 1), (5000, 2, 3), (5000, 4, 1)])
def test_spec_is_used_and_identical(n_lipids, mpt, split, monkeypatch):
    monkeypatch.setenv("GORDER_MPT", str(mpt))
    s = synthetic.s_cg(n_lipids, leaflet_mode=abi.LEAFLET_GLOBAL, collect_leaflets=True, timewise=True, split_types=split)
    xyz, box, idx = s.frames(0, 9)
    g, st = _run(s.setup, xyz, box, idx, batches=3)
    assert st == {"enabled": True, "frames_speculated": 9, "frames_repaired": 0}
    monkeypatch.setenv("GORDER_NO_SPEC", "1")
    e, st0 = _run(s.setup, xyz, box, idx, batches=2)
    assert st0["frames_speculated"] == 0 and not st0["enabled"]
    _same(g, e)
    monkeypatch.delenv("GORDER_NO_SPEC")
    g2, r = run_both(s.setup, xyz, box, idx)
    _same(g, g2)
    assert_raw_parity(g2, r, s.setup, what="speculative leaflets")


def test_spec_repairs_drifting_membrane(monkeypatch):
    """The membrane jumps along z from frame to frame (and across the periodic boundary): the provisional centre
    is useless for most frames, which must be detected and repaired with the exact centre."""
    s = synthetic.s_cg(1500, leaflet_mode=abi.LEAFLET_GLOBAL, collect_leaflets=True, timewise=True)
    xyz, box, idx = s.frames(0, 8)
    shifts = np.array([0.0, 0.02, 1.7, 3.1, -2.9, 5.5, 0.4, -0.3], np.float32)
    for f in range(8):
        z = xyz[f, :, 2] + shifts[f]
        xyz[f, :, 2] = z - np.floor(z / box[f, 2]) * box[f, 2]   # every atom wrapped: lipids broken at the boundary
    g, st = _run(s.setup, xyz, box, idx)
    assert st["frames_speculated"] == 8 and 3 <= st["frames_repaired"] <= 6, st
    monkeypatch.setenv("GORDER_NO_SPEC", "1")
    e, _ = _run(s.setup, xyz, box, idx)
    monkeypatch.delenv("GORDER_NO_SPEC")
    _same(g, e)
    g2, r = run_both(s.setup, xyz, box, idx, batches=2)
    assert_raw_parity(g2, r, s.setup, what="drifting membrane")
    up = g.leaflets.astype(int).sum(axis=1)
    assert np.all(np.abs(up - 750) <= 2), up    # both leaflets found in every frame


def test_spec_repairs_heads_at_the_centre(monkeypatch):
    """Lipids whose head sits at the membrane centre (flip-flop in progress): the side they fall on depends on
    the last bits of the centre, so the frame has to take the exact path."""
    s = synthetic.s_cg(900, leaflet_mode=abi.LEAFLET_GLOBAL, collect_leaflets=True)
    xyz, box, idx = s.frames(0, 5)
    head = np.asarray(s.setup.moltypes[0].mol_base)[[5, 77, 400]] + 1
    for f in (1, 3):
        centre = xyz[f, : 900 * 12, 2].mean()
        xyz[f, head, 2] = centre + np.array([1e-6, -2e-6, 3e-5], np.float32)
    g, st = _run(s.setup, xyz, box, idx)
    assert st["frames_repaired"] == 2, st
    monkeypatch.setenv("GORDER_NO_SPEC", "1")
    e, _ = _run(s.setup, xyz, box, idx)
    monkeypatch.delenv("GORDER_NO_SPEC")
    _same(g, e)
    g2, r = run_both(s.setup, xyz, box, idx)
    # the oracle's centre is an f32 running sum in atom order: a head within 1e-5 nm of it may legitimately differ
    far = np.ones(900, bool)
    far[[5, 77, 400]] = False
    np.testing.assert_array_equal(g2.leaflets[:, far], r.leaflets[:, far])


def test_spec_switches_itself_off(monkeypatch):
    """Thin water layer: the membrane fills most of the box along z, the moment bound on the circular mean cannot
    exclude a re-imaged atom, every frame is repaired -- after a few batches the engine returns to the pre-pass."""
    s = synthetic.s_cg(600, leaflet_mode=abi.LEAFLET_GLOBAL, collect_leaflets=True, max_batch_frames=8)
    xyz, box, idx = s.frames(0, 40)
    box = box.copy()
    box[:, 2] = 6.4   # membrane thickness ~6 nm in a 6.4 nm box
    xyz = xyz.copy()
    xyz[:, :, 2] -= 2.8
    xyz[:, :, 2] -= np.floor(xyz[:, :, 2] / 6.4) * 6.4
    eng = SystemTopology(s.setup)
    for a in range(0, 40, 8):
        eng.analyze_frames(xyz[a:a + 8], box[a:a + 8], idx[a:a + 8])
        eng.sync()
    st = eng.speculation_stats()
    g = eng.finish()
    eng.close()
    assert not st["enabled"] and 16 <= st["frames_speculated"] < 40 and st["frames_repaired"] >= 8, st
    monkeypatch.setenv("GORDER_NO_SPEC", "1")
    e, _ = _run(s.setup, xyz, box, idx, batches=5)
    _same(g, e)


@pytest.mark.parametrize("n_lipids,batches", [(64, 1), (1100, 3)])
def test_spec_with_membrane_atoms_no_bond_loads(n_lipids, batches, monkeypatch):
    """AA: the membrane group holds atoms no bond loads (head group, lipids of a type without bonds): their
    displacements come from the side pass (spec_leftover_kernel), the result is that of the pre-pass path."""
    s = synthetic.s_aa(n_lipids, n_water=100, leaflet_mode=abi.LEAFLET_GLOBAL, collect_leaflets=True)
    xyz, box, idx = s.frames(0, 6)
    g, st = _run(s.setup, xyz, box, idx, batches=batches)
    assert st == {"enabled": True, "frames_speculated": 6, "frames_repaired": 0}
    monkeypatch.setenv("GORDER_NO_SPEC_LEFTOVER", "1")
    e, st0 = _run(s.setup, xyz, box, idx)
    assert not st0["enabled"] and st0["frames_speculated"] == 0
    monkeypatch.delenv("GORDER_NO_SPEC_LEFTOVER")
    _same(g, e)
    g2, r = run_both(s.setup, xyz, box, idx)
    assert_raw_parity(g2, r, s.setup, what="speculative leaflets, side pass")


def test_spec_side_pass_sees_what_the_bond_kernel_does_not(monkeypatch):
    """Atoms of the side pass half a box away from the membrane centre (which image they belong to depends on the
    exact centre) must flag the frame."""
    s = synthetic.s_aa(300, n_water=50, leaflet_mode=abi.LEAFLET_GLOBAL, collect_leaflets=True)
    xyz, box, idx = s.frames(0, 5)
    mt = s.setup.moltypes[0]
    rel = set(np.asarray(mt.bond_rel).ravel().tolist()) | {mt.head_rel}
    left = [a for a in np.asarray(s.setup.membrane) if (a - np.asarray(mt.mol_base)[0]) not in rel and a < np.asarray(mt.mol_base)[1]]
    assert left, "the synthetic AA lipid has membrane atoms outside its bonds"
    stride = int(np.asarray(mt.mol_base)[1] - np.asarray(mt.mol_base)[0])
    far = np.array([left[0] + stride * m for m in (3, 150, 299)])
    centre = xyz[2, np.asarray(s.setup.membrane), 2].mean()
    xyz[2, far, 2] = centre + 0.5 * box[2, 2] - np.array([0.02, 0.01, -0.01], np.float32)
    g, st = _run(s.setup, xyz, box, idx)
    assert st["frames_speculated"] == 5 and st["frames_repaired"] == 1, st
    monkeypatch.setenv("GORDER_NO_SPEC", "1")
    e, _ = _run(s.setup, xyz, box, idx)
    monkeypatch.delenv("GORDER_NO_SPEC")
    _same(g, e)
    # ... and an undefined position that only the side pass reads ends as it does on the pre-pass path:
    # "invalid global membrane center" (leaflets.rs:187 -> AnalysisError::InvalidGlobalMembraneCenter)
    xyz[3, far[0], 2] = np.nan
    for no_spec in (False, True):
        if no_spec:
            monkeypatch.setenv("GORDER_NO_SPEC", "1")
        with pytest.raises(abi.GorderError) as err:
            _run(s.setup, xyz, box, idx, native=True)   # (the AoS upload would already refuse the NaN)
        assert err.value.code == abi.ERR_INVALID_GLOBAL_CENTER, (no_spec, str(err.value))
    monkeypatch.delenv("GORDER_NO_SPEC")


@pytest.mark.parametrize("timewise", [False, True])
def test_spec_pipelined_on_resident_frames(timewise, monkeypatch):
    """Frames resident in HBM in the plane layout: setup | bond kernel | tail of consecutive batches run on three
    streams.  Many small batches, with a frame that needs repair in the middle, against the one-shot host path."""
    import torch
    s = synthetic.s_cg(3000, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=timewise, max_batch_frames=3)
    xyz, box, idx = s.frames(0, 24)
    xyz[13, :, 2] += 2.2                                   # the membrane jumps in frame 13
    xyz[13, :, 2] -= np.floor(xyz[13, :, 2] / box[13, 2]) * box[13, 2]
    want, _ = _run(s.setup, xyz, box, idx)
    eng = SystemTopology(s.setup)
    planes = torch.from_numpy(eng.to_native(xyz)).cuda()
    dbox = torch.from_numpy(box).cuda()
    ff = eng.frame_floats
    for a in range(0, 24, 3):
        eng.analyze_frames_device(planes.data_ptr() + 4 * ff * a, dbox.data_ptr() + 12 * a, 3, frame_index=idx[a:a + 3], native=True)
    st = eng.speculation_stats()
    got = eng.finish()
    eng.close()
    assert st["frames_speculated"] == 24 and 1 <= st["frames_repaired"] <= 3, st
    _same(got, want)
