"""Loader of the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py
from the reference's own test files and expected values)."""
from __future__ import annotations

import json
import os

import numpy as np

from gorder_b200 import abi, results

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURE_TOL = 2e-4   # the reference's own comparator (tests/common/mod.rs:139-150); YAML holds 4 decimals


def single_frame(name: str):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    setup = abi.EngineSetup.from_dict(json.loads(str(z["setup"])))
    exp = {k: np.array(v) for k, v in json.loads(str(z["expected"])).items()}
    return setup, z["xyz"][None].astype(np.float32), z["box"][None].astype(np.float32), exp


def aa_traj():
    """AA end-to-end fixture (tests_aa.rs:1019-1040 -> tests/files/aa_order_selected.yaml): setup, frames, expected."""
    z = np.load(os.path.join(GOLDEN, "aa_traj.npz"))
    setup = abi.EngineSetup.from_dict(json.loads(str(z["setup"])))
    xyz = z["q"].astype(np.float32) * np.float32(1.0 / float(z["precision"]))   # exactly the XTC decoder's arithmetic
    case = dict(expected=json.loads(str(z["expected"])), keys=json.loads(str(z["keys"])), source="aa_order_selected.yaml",
                molecules=json.loads(str(z["molecules"])))
    return setup, xyz, z["box"].astype(np.float32), case


def ua_nopbc():
    """UA without PBC (tests_ua.rs:688-714 -> ua_order_leaflets_nopbc.yaml): setup, frames, (unused) boxes, case."""
    z = np.load(os.path.join(GOLDEN, "ua_nopbc.npz"))
    q = np.cumsum(z["dq"].astype(np.int32), axis=0)
    xyz = q.astype(np.float32) * np.float32(1.0 / float(z["precision"]))
    case = json.loads(str(z["case"]))
    return abi.EngineSetup.from_dict(case["setup"]), xyz, z["box"].astype(np.float32), case


_FULL = {}


def full_case(which: str, name: str):
    """A case of the full AA (pcpepg, 51 frames) or CG (cg, 101 frames) test trajectory of the reference, re-joined from
    tests/files/split/* (make_golden.concatenated): setup, frames, boxes, frame indices, case dict."""
    if which not in _FULL:
        z = np.load(os.path.join(GOLDEN, f"{which}_full.npz"))
        q = np.cumsum(z["dq"].astype(np.int32), axis=0)
        xyz = q.astype(np.float32) * np.float32(1.0 / float(z["precision"]))   # exactly the XTC decoder's arithmetic
        _FULL[which] = (xyz, z["box"].astype(np.float32), json.loads(str(z["cases"])))
    xyz, box, cases = _FULL[which]
    c = cases[name]
    setup = abi.EngineSetup.from_dict(c["setup"])
    fr = np.array(c["frames"], dtype=np.int64)
    return setup, xyz[fr], box[fr], fr - fr[0], c


def full_case_names(which: str):
    z = np.load(os.path.join(GOLDEN, f"{which}_full.npz"))
    return list(json.loads(str(z["cases"])))


def cg_asym():
    """The reference's asymmetric CG membrane (tests_cg.rs:2182-2309: beads ``@membrane``, Global leaflets, heads ``name
    PO4``).  Its topology exists only as a TPR file, so the setup is built HERE by the C++ reader and classifier
    (gorder_b200.structure): TPR -> Master group (the membrane atoms, renumbered densely: common.rs:92-103) -> molecule types.
    Returns setup, frames, boxes, cases."""
    from gorder_b200.structure import System
    z = np.load(os.path.join(GOLDEN, "cg_asym.npz"))
    s = System.from_tpr(os.path.join(GOLDEN, "tpr", "cg_asym.tpr"))
    names, resn, resid, *_ = s.atoms()
    keep = z["keep"].astype(np.int64)
    new = -np.ones(s.n_atoms, np.int64)
    new[keep] = np.arange(keep.size)
    b = s.bonds().astype(np.int64)
    b = b[(new[b[:, 0]] >= 0) & (new[b[:, 1]] >= 0)]
    master = System.from_arrays([names[i] for i in keep], [resn[i] for i in keep], res_ids=resid[keep])
    master.set_bonds(new[b])
    allm = np.arange(keep.size)
    heads = np.array([i for i in allm if names[keep[i]] == "PO4"])
    mts = master.classify_bonds(abi.KIND_CG, allm, allm, heads=heads)
    setup = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=int(keep.size), moltypes=mts, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL)
    q = np.cumsum(z["dq"].astype(np.int32), axis=0)
    xyz = q.astype(np.float32) * np.float32(1.0 / float(z["precision"]))   # exactly the XTC decoder's arithmetic
    return setup, xyz, z["box"].astype(np.float32), json.loads(str(z["cases"]))


_UA = None


def ua_traj():
    global _UA
    if _UA is None:
        z = np.load(os.path.join(GOLDEN, "ua_traj.npz"))
        xyz = z["q"].astype(np.float32) * np.float32(1.0 / 1000.0)   # exactly the XTC decoder's arithmetic
        _UA = (xyz, z["box"].astype(np.float32), z["time"], z["structure_box"], json.loads(str(z["cases"])))
    return _UA


def ua_case(name: str):
    xyz, box, _time, _sbox, cases = ua_traj()
    c = cases[name]
    setup = abi.EngineSetup.from_dict(c["setup"])
    fr = np.array(c["frames"], dtype=np.int64)
    # SystemTopology::frame counts trajectory frames from `begin` (topology/mod.rs:41-43)
    frame_index = fr - fr[0]
    return setup, xyz[fr], box[fr], frame_index, c


def flatten_results(res: results.AnalysisResults, keys=("total",), with_error=False):
    """Same traversal as make_golden.flatten_yaml."""
    out = []

    def push(coll):
        for k in keys:
            o = getattr(coll, k)
            if o is None:
                continue
            out.append(o.value)
            if with_error:
                out.append(o.error)

    push(res.average)
    for m in res.molecules.values():
        push(m.average)
        for it in m.items:
            push(it.order)
            for b in it.bonds:
                push(b)
    return np.array(out, dtype=np.float64)


def assert_matches_yaml(raw: abi.RawResults, setup: abi.EngineSetup, case: dict, tol: float = FIXTURE_TOL):
    tol = max(tol, case.get("tol", 0.0))
    nb = case.get("n_blocks")
    res = results.convert(raw, setup, n_blocks=nb, min_samples=case.get("min_samples", 1))
    got = flatten_results(res, tuple(case["keys"]), with_error=nb is not None)
    exp = np.array(case["expected"], dtype=np.float64)
    assert got.shape == exp.shape, (got.shape, exp.shape)
    np.testing.assert_allclose(got, exp, atol=tol, rtol=0, equal_nan=True, err_msg=f"fixture {case['source']}")
    # the converter of the shared library (gorder_results_*) gives the same bits as the numpy one
    nat = results.convert(raw, setup, n_blocks=nb, min_samples=case.get("min_samples", 1), native=True)
    got_native = flatten_results(nat, tuple(case["keys"]), with_error=nb is not None)
    np.testing.assert_array_equal(got_native.astype(np.float32), got.astype(np.float32), err_msg=f"native converter, {case['source']}")
