"""GROMACS index files and leaflets read from them (``LeafletClassification::from_ndx``, leaflets.rs:1030-1215; groan_rs
``Groups::from_ndx``) in C++ behind the C ABI: ``gorder_ndx_*``, ``gorder_leaflets_from_ndx``.  Fixtures: the reference's own
ndx files (tests/files/ndx, copied to tests/golden/ndx) and its expected output for the runs that use them
(tests_aa.rs:5210-5360, tests_cg.rs analogues: the same ``*_order_leaflets.yaml`` as with Global leaflets)."""
import dataclasses
import os

import numpy as np
import pytest

from gorder_b200 import abi
from gorder_b200.structure import leaflets_from_ndx, read_ndx

import golden_cases as gc

NDX = os.path.join(gc.GOLDEN, "ndx")


def test_read_ndx_groups():
    g = read_ndx(os.path.join(NDX, "pcpepg_leaflets_all.ndx"), 68375)
    assert list(g) == ["SomeGroup", "Lower", "Upper", "IrrelevantGroup"]          # "[ Upper]" and "[Upper ]" are "Upper"
    assert g["Upper"].size == 6921 and g["Lower"].size == 6919 and g["Lower"][0] == 8125   # atom numbers start at 1 in the file
    p = read_ndx(os.path.join(NDX, "pcpepg_leaflets.ndx"), 68375)                # the P atoms only
    assert list(p) == ["Lower", "Upper"] and p["Upper"].size + p["Lower"].size == 274 and p["Lower"][0] == 8135
    d = read_ndx(os.path.join(NDX, "pcpepg_leaflets_duplicate_irrelevant.ndx"))
    assert list(d) == ["Irrelevant Group", "Lower", "Upper"]                      # a repeated name replaces the earlier group
    i = read_ndx(os.path.join(NDX, "pcpepg_leaflets_invalid_irrelevant.ndx"))
    assert list(i) == ["Lower", "Upper"]                                          # "Inv@alidName" is refused
    with pytest.raises(abi.GorderError) as e:
        read_ndx(os.path.join(NDX, "pcpepg_leaflets.ndx"), 9000)                  # atom numbers beyond the system
    assert e.value.code == abi.ERR_NDX_PARSE
    with pytest.raises(abi.GorderError) as e:
        read_ndx(os.path.join(NDX, "nothing.ndx"))
    assert e.value.code == abi.ERR_IO


def test_ndx_leaflet_errors(tmp_path):
    """NdxLeafletClassificationError: InvalidName / DuplicateName only when the name is one of the two that are asked for,
    GroupNotFound, AssignmentNotFound (leaflets.rs:1090-1175; the reference's leaflets_*_main / missing_* files)."""
    heads = [0, 1, 2]
    # the reference's own cases (leaflets.rs:3109-3225)
    for fn, up, code, what in (("leaflets_duplicate_main.ndx", "Upper", abi.ERR_NDX_DUPLICATE_NAME, "'Upper'"),
                               ("leaflets_invalid_main.ndx", "U!pper", abi.ERR_NDX_INVALID_NAME, "'U!pper'"),
                               ("leaflets_missing_upper.ndx", "Upper", abi.ERR_NDX_GROUP_NOT_FOUND, "'Upper' expected to specify upper-leaflet"),
                               ("leaflets_missing_lower.ndx", "Upper", abi.ERR_NDX_GROUP_NOT_FOUND, "'Lower' expected to specify lower-leaflet")):
        with pytest.raises(abi.GorderError) as e:
            leaflets_from_ndx([os.path.join(NDX, fn)], heads, up, "Lower")
        assert e.value.code == code and what in str(e.value) and fn in str(e.value), (fn, e.value)
    assert leaflets_from_ndx([os.path.join(NDX, "leaflets_only_upper.ndx")], [0, 1, 2, 3]).tolist() == [[1, 1, 1, 1]]   # an empty leaflet is fine
    f = tmp_path / "x.ndx"
    f.write_text("[ Up@per ]\n1 2\n[ Lower ]\n3\n")
    with pytest.raises(abi.GorderError) as e:
        leaflets_from_ndx([str(f)], heads, "Up@per", "Lower")
    assert e.value.code == abi.ERR_NDX_INVALID_NAME
    f.write_text("[ Upper ]\n1 2\n[ Lower ]\n4\n")
    with pytest.raises(abi.GorderError) as e:
        leaflets_from_ndx([str(f)], heads)
    assert e.value.code == abi.ERR_NDX_ASSIGNMENT_NOT_FOUND and "molecule 2" in str(e.value)
    f.write_text("[ Upper ]\n1 2 x\n[ Lower ]\n3\n")
    with pytest.raises(abi.GorderError) as e:
        leaflets_from_ndx([str(f)], heads)
    assert e.value.code == abi.ERR_NDX_PARSE
    f.write_text("[ Upper ]\n1 2\n[ Lower ]\n3\n[ Lower ]\n")   # the repeated group is empty: DuplicateName wins over the assignment
    with pytest.raises(abi.GorderError) as e:
        leaflets_from_ndx([str(f)], heads)
    assert e.value.code == abi.ERR_NDX_DUPLICATE_NAME
    f.write_text("[ Upper ]\n1 2\n[ Lower ]\n3\n")
    assert leaflets_from_ndx([str(f), str(f)], heads).tolist() == [[1, 1, 0], [1, 1, 0]]


def ndx_case(which: str, variant: str):
    """The reference's `leaflets` run of the AA / CG test trajectory with the leaflets read from ndx files instead of computed:
    tests_aa.rs:5210-5360 (once; one file per frame; every 10th frame with files that hold repeated / refused groups)."""
    setup, xyz, box, fi, case = gc.full_case(which, "leaflets_global")
    pre = "pcpepg" if which == "aa" else "cg"
    one, all_ = os.path.join(NDX, f"{pre}_leaflets.ndx"), os.path.join(NDX, f"{pre}_leaflets_all.ndx")
    n = xyz.shape[0]
    if variant == "once":
        files, kind, freq = [one], abi.FREQ_ONCE, 1
    elif variant == "every":
        files, kind, freq = ([one, all_] * n)[:n], abi.FREQ_EVERY, 1
    else:
        irr = [os.path.join(NDX, f"pcpepg_leaflets_{x}_irrelevant.ndx") for x in ("duplicate", "invalid")] if which == "aa" else [one, all_]
        files, kind, freq = ([one, all_] + irr + [one, all_, one, all_, one, all_, one])[: (n + 9) // 10], abi.FREQ_EVERY, 10
    mts = []
    for m in setup.moltypes:
        heads = np.asarray(m.mol_base) + m.head_rel     # the Master group of these fixtures starts at atom 0 of the system
        mts.append(dataclasses.replace(m, manual_leaflets=leaflets_from_ndx(files, heads)))
    st = dataclasses.replace(setup, moltypes=mts, leaflet_mode=abi.LEAFLET_MANUAL, leaflet_freq_kind=kind, leaflet_freq=freq, membrane=())
    return st, xyz, box, fi, case


@pytest.mark.parametrize("which,variant", [("cg", "once"), ("cg", "every"), ("aa", "every10")])
def test_leaflets_from_ndx_oracle(which, variant):
    from oracle import oracle
    st, xyz, box, fi, case = ndx_case(which, variant)
    o = oracle.Oracle(st, n_threads=8)
    o.analyze_frames(xyz, box, fi)
    gc.assert_matches_yaml(o.finish(), st, case)
