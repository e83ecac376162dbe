"""C++ structure / topology / classification (``csrc/gorder_topology.inl``) -- host only, runs without a GPU.

Pinned in two ways:
  * on the small GROMACS-written TPR files committed under ``tests/golden/tpr`` (tpx 103, 122, 127) against
    ``tests/golden/tpr/expected.json`` (written by ``tests/golden/make_golden.py tpr`` after the reader agreed, file by
    file, with the independent parser ``tests/golden/tpr_independent.py`` and with the reference's .gro / .bnd / .pdb files:
    a regression pin for machines without the reference tree) and against the reference's expected output for
    ``cg_asym.tpr`` (``tests/test_gpu_topology.py``, ``test_cg_asym_oracle`` below);
  * where the reference tree is mounted, on all 14 TPR files of ``/root/reference/tests/files`` against the .gro / .bnd /
    .pdb files next to them, and the C++ classifier against the Python restatement in ``oracle/fixtures.py`` (which the
    reference's YAML fixtures pin end to end).
"""
import dataclasses
import json
import os

import numpy as np
import pytest

from gorder_b200 import abi
from gorder_b200.structure import System

HERE = os.path.dirname(os.path.abspath(__file__))
TPR = os.path.join(HERE, "golden", "tpr")
REF = "/root/reference/tests/files"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")

LIPIDS = {"POPC", "POPE", "POPG", "POPS", "OPC"}


def _membrane(names, resn):
    return [i for i in range(len(names)) if resn[i] in LIPIDS]


# ---- committed fixtures -------------------------------------------------------------------------------------------------
def test_tpr_fixtures_match_expected():
    exp = json.load(open(os.path.join(TPR, "expected.json")))
    assert len(exp) >= 5
    for fn, e in exp.items():
        s = System.from_tpr(os.path.join(TPR, fn))
        assert s.tpx_version == e["tpx"], fn
        assert s.n_atoms == e["n_atoms"] and s.n_bonds == e["n_bonds"], fn
        names, resn, resid, z, m, q = s.atoms()
        assert names[:24] == e["names_head"] and resn[:24] == e["resn_head"], fn
        import zlib
        assert zlib.crc32(" ".join(names).encode()) == e["names_crc"], fn
        assert zlib.crc32(" ".join(resn).encode()) == e["resn_crc"], fn
        assert zlib.crc32(s.bonds().astype("<i4").tobytes()) == e["bonds_crc"], fn
        x = s.positions()
        assert x is not None and np.allclose(x[:4].reshape(-1), e["xyz_head"], atol=0, rtol=0), fn
        b = s.box9()
        assert np.array_equal(b.astype(np.float32), np.array(e["box9"], np.float32)), fn
        assert [int(v) for v in resid[:24]] == e["resid_head"], fn
        s.close()


def test_tpr_classification_cg_fixture():
    """cg_asym.tpr: ``@membrane`` beads, Global leaflets with heads ``name PO4`` (tests_cg.rs:2199-2212)."""
    s = System.from_tpr(os.path.join(TPR, "cg_asym.tpr"))
    names, resn, *_ = s.atoms()
    mem = _membrane(names, resn)
    heads = [i for i in mem if names[i] == "PO4"]
    mts = s.classify_bonds(abi.KIND_CG, mem, mem, heads=heads)
    assert [m.name for m in mts] == ["POPE", "POPG"]
    assert [m.n_molecules for m in mts] == [144, 135]
    assert all(len(m.bond_rel) == 11 and m.head_rel == 1 for m in mts)
    assert mts[0].bond_names[0] == "POPE NH3 (0) - POPE PO4 (1)"
    assert mts[0].mol_base[:3] == [0, 12, 24] and mts[1].mol_base[0] == 144 * 12
    # bond.rs:77-81: sorted by the relative indices
    assert mts[0].bond_rel == sorted(mts[0].bond_rel)


def test_same_name_types_are_renamed():
    """classify.rs:262-294 on same_name.tpr: two POPC topologies -> POPC1, POPC2."""
    s = System.from_tpr(os.path.join(TPR, "same_name.tpr"))
    names, resn, *_ = s.atoms()
    allb = list(range(s.n_atoms))
    mts = s.classify_bonds(abi.KIND_CG, allb, allb)
    assert [m.name for m in mts] == ["POPC1", "POPC2"]
    assert [m.n_molecules for m in mts] == [2, 1]


def test_cyclic_molecule():
    s = System.from_tpr(os.path.join(TPR, "cyclic.tpr"))
    allb = list(range(s.n_atoms))
    mts = s.classify_bonds(abi.KIND_CG, allb, allb)
    assert len(mts) == 1 and mts[0].n_molecules == 3 and len(mts[0].bond_rel) == 14


def test_multiple_residues_one_molecule():
    s = System.from_tpr(os.path.join(TPR, "multiple_resid.tpr"))
    allb = list(range(s.n_atoms))
    mts = s.classify_bonds(abi.KIND_CG, allb, allb)
    assert sum(m.n_molecules for m in mts) == 3


# ---- errors -------------------------------------------------------------------------------------------------------------
def test_tpr_errors(tmp_path):
    with pytest.raises(abi.GorderError) as e:
        System.from_tpr(str(tmp_path / "missing.tpr"))
    assert e.value.code == abi.ERR_IO
    raw = open(os.path.join(TPR, "cyclic.tpr"), "rb").read()
    for cut in (0, 3, 40, 99, 500, 5000, 14500):   # 14500: inside the coordinates (the file goes on with velocities and run parameters)
        p = tmp_path / f"cut{cut}.tpr"
        p.write_bytes(raw[:cut])
        with pytest.raises(abi.GorderError) as e:
            System.from_tpr(str(p))
        assert e.value.code == abi.ERR_TPR_FORMAT
    bad = bytearray(raw)
    ver_at = 8 + (int.from_bytes(raw[4:8], "big") + 3) // 4 * 4 + 4   # VERSION string, precision, then the tpx version
    assert int.from_bytes(raw[ver_at:ver_at + 4], "big") == 127
    bad[ver_at:ver_at + 4] = (200).to_bytes(4, "big")   # tpx version of the future
    p = tmp_path / "future.tpr"
    p.write_bytes(bytes(bad))
    with pytest.raises(abi.GorderError) as e:
        System.from_tpr(str(p))
    assert e.value.code == abi.ERR_TPR_FORMAT and "200" in str(e.value)
    p = tmp_path / "noise.tpr"
    p.write_bytes(bytes(np.random.default_rng(1).integers(0, 256, 4096, dtype=np.uint8)))
    with pytest.raises(abi.GorderError):
        System.from_tpr(str(p))


def test_tpr_fuzz_does_not_crash(tmp_path):
    """Flipped bytes anywhere in the file: either a system comes back or GORDER_ERR_TPR_FORMAT, never a crash."""
    raw = open(os.path.join(TPR, "cyclic.tpr"), "rb").read()
    rng = np.random.default_rng(7)
    ok = 0
    for it in range(300):
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, min(len(b), 14300)))   # the topology part
            b[pos] = int(rng.integers(0, 256))
        p = tmp_path / "f.tpr"
        p.write_bytes(bytes(b))
        try:
            s = System.from_tpr(str(p))
            assert s.n_atoms == 36
            s.bonds(); s.atoms(); s.close()
            ok += 1
        except abi.GorderError as e:
            assert e.code == abi.ERR_TPR_FORMAT
    assert ok > 0


def _toy():
    names = ["P", "C1", "C2", "C3"] * 3
    resn = ["LIP"] * 12
    s = System.from_arrays(names, resn, res_ids=[1] * 4 + [2] * 4 + [3] * 4)
    s.set_bonds([(0, 1), (1, 2), (2, 3), (4, 5), (5, 6), (6, 7), (8, 9), (9, 10), (10, 11)])
    return s


def test_bonds_file(tmp_path):
    s = _toy()
    f = tmp_path / "b.bnd"
    f.write_text("# comment\n1 2\n2 3 1   # both directions\n\n3 4\n5 6\n6 7\n7 8\n9 10\n10 11\n11 12 10\n")
    s.read_bonds(str(f))
    assert s.n_bonds == 9 and s.bonds().tolist()[:3] == [[0, 1], [1, 2], [2, 3]]
    for text, code in (("1 x\n", abi.ERR_BONDS_PARSE), ("1 1\n", abi.ERR_BONDS_SELF), ("1 13\n", abi.ERR_BONDS_ATOM_NOT_FOUND),
                       ("14 1\n", abi.ERR_BONDS_ATOM_NOT_FOUND), ("1 -2\n", abi.ERR_BONDS_PARSE)):
        f.write_text(text)
        with pytest.raises(abi.GorderError) as e:
            s.read_bonds(str(f))
        assert e.value.code == code, text
    assert s.n_bonds == 9   # a failed read leaves the bonds alone
    with pytest.raises(abi.GorderError) as e:
        s.read_bonds(str(tmp_path / "none.bnd"))
    assert e.value.code == abi.ERR_IO


def test_classification_errors_and_warnings():
    s = _toy()
    allb = list(range(12))
    mts = s.classify_bonds(abi.KIND_CG, allb, allb, heads=[0, 4, 8], methyls=[3, 7, 11], normal_heads=[0, 4, 8])
    assert len(mts) == 1 and mts[0].n_molecules == 3 and mts[0].head_rel == 0 and mts[0].methyl_rel == [3] and mts[0].normal_head_rel == 0
    assert mts[0].bond_rel == [(0, 1), (1, 2), (2, 3)]
    with pytest.raises(abi.GorderError) as e:
        s.classify_bonds(abi.KIND_CG, allb, allb, heads=[0, 4])
    assert e.value.code == abi.ERR_TOPOLOGY_NO_HEAD and "'8'" in str(e.value)
    with pytest.raises(abi.GorderError) as e:
        s.classify_bonds(abi.KIND_CG, allb, allb, heads=[0, 1, 4, 8])
    assert e.value.code == abi.ERR_TOPOLOGY_MULTIPLE_HEADS
    with pytest.raises(abi.GorderError) as e:
        s.classify_bonds(abi.KIND_CG, allb, allb, methyls=[3, 7])
    assert e.value.code == abi.ERR_TOPOLOGY_NO_METHYL
    with pytest.raises(abi.GorderError) as e:
        s.classify_bonds(abi.KIND_CG, allb, allb, methyls=[3, 6, 7, 11])
    assert e.value.code == abi.ERR_TOPOLOGY_INCONSISTENT_METHYLS
    with pytest.raises(abi.GorderError) as e:
        s.classify_ua([], [])
    assert e.value.code == abi.ERR_TOPOLOGY_NO_UA_CARBONS
    # classify.rs:297-315: a molecule type without order bonds -> nothing is analysed, with a warning
    mts = s.classify_bonds(abi.KIND_AA, [1, 5, 9], [])
    assert mts == [] and "No bonds/atoms" in s.last_warning
    mts = s.classify_bonds(abi.KIND_AA, [], [])
    assert mts == [] and "No molecules" in s.last_warning
    # AA: only bonds between the two groups count
    mts = s.classify_bonds(abi.KIND_AA, [1, 5, 9], [2, 6, 10])
    assert mts[0].bond_rel == [(1, 2)] and mts[0].bond_names == ["LIP C1 (1) - LIP C2 (2)"]


def test_ua_carbon_types_toy():
    """uaorder.rs:580-665 on a 4-atom chain: P-C1-C2-C3 with C1..C3 saturated: C1, C2 = CH2, C3 = CH3 (helpers C2, C1)."""
    s = _toy()
    sat = [1, 2, 3, 5, 6, 7, 9, 10, 11]
    mts = s.classify_ua(sat)
    m = mts[0]
    assert m.ua_kind == [abi.UA_CH2, abi.UA_CH2, abi.UA_CH3]
    assert m.ua_rel == [(1, 0, 2, -1), (2, 1, 3, -1), (3, 2, 1, -1)]
    assert m.bond_names == ["LIP C1 (1)", "LIP C2 (2)", "LIP C3 (3)"]
    m = s.classify_ua(sat, ignore=[0, 4, 8])[0]
    assert m.ua_kind[0] == abi.UA_CH3 and m.ua_rel[0] == (1, 2, 3, -1)
    m = s.classify_ua([3, 7, 11], unsaturated=[1, 2, 5, 6, 9, 10])[0]
    assert m.ua_kind == [abi.UA_CH1_UNSAT, abi.UA_CH1_UNSAT, abi.UA_CH3]


# ---- against the reference tree -----------------------------------------------------------------------------------------
ALL_TPR = ["cg.tpr", "pcpepg.tpr", "ua.tpr", "asymmetric/cg_asym.tpr", "asymmetric/aa_asym.tpr", "cyclic.tpr", "scrambling/cg_scrambling.tpr",
           "cg_buckled.tpr", "pepg_cg.tpr", "multiple_resid.tpr", "multiple_resid_same_name.tpr", "same_name.tpr", "pcpepg_switched_xz.tpr",
           "pcpepg_switched_yz.tpr"]


@needs_ref
def test_reference_tprs_all_parse():
    for fn in ALL_TPR:
        s = System.from_tpr(os.path.join(REF, fn))
        assert s.n_atoms > 0 and s.n_bonds > 0 and s.positions() is not None and s.box9() is not None, fn
        s.close()


@needs_ref
@pytest.mark.parametrize("tpr,gro,bnd,off", [("cg.tpr", "cg.gro", "cg.bnd", 37646), ("pcpepg.tpr", "pcpepg.gro", "pcpepg.bnd", 249767)])
def test_reference_tpr_vs_gro_and_bnd(tpr, gro, bnd, off):
    from oracle import fixtures
    s = System.from_tpr(os.path.join(REF, tpr))
    st = fixtures.read_gro(os.path.join(REF, gro))
    fixtures.read_bnd(os.path.join(REF, bnd), st)
    names, resn, resid, z, m, q = s.atoms()
    assert names == st.name and resn == st.resname
    assert np.array_equal(resid % 100000, st.resid)   # GROMACS' renumbering of one-residue molecule types (the .gro holds 5 digits)
    lip = set(i for i in range(s.n_atoms) if resn[i] in LIPIDS)
    mine = set(map(tuple, s.bonds().tolist()))
    assert set(st.bonds) <= mine
    assert all(not (i in lip or j in lip) for i, j in mine - set(st.bonds))   # the .bnd files leave out the water (SETTLE) bonds
    # coordinates: the block SURVEY.md §8c located by matching against the GRO
    x, box, where = fixtures.tpr_coordinates(os.path.join(REF, tpr), st.xyz, hint=off)
    assert where == off and np.array_equal(s.positions(), x)
    assert np.array_equal(s.box9()[[0, 4, 8]], box)
    # the bonds file through the C++ parser gives the .bnd bonds exactly
    s.read_bonds(os.path.join(REF, bnd))
    assert set(map(tuple, s.bonds().tolist())) == set(st.bonds)


@needs_ref
def test_reference_ua_tpr_vs_pdb():
    from oracle import fixtures
    s = System.from_tpr(os.path.join(REF, "ua.tpr"))
    st = fixtures.read_pdb(os.path.join(REF, "ua_nobox.pdb"))
    names, resn, resid, *_ = s.atoms()
    assert names == st.name and resn == st.resname and np.array_equal(resid % 10000, st.resid % 10000)
    assert np.abs(s.positions() - st.xyz).max() < 1e-6
    lip = set(i for i in range(s.n_atoms) if resn[i] in LIPIDS)
    mine = set(t for t in map(tuple, s.bonds().tolist()) if t[0] in lip)
    assert mine == set(st.bonds)


def _same_types(a, b, kind):
    assert [m.name for m in a] == [m.name for m in b]
    for x, y in zip(a, b):
        assert list(x.mol_base) == list(y.mol_base), x.name
        assert [tuple(t) for t in x.bond_rel] == [tuple(t) for t in y.bond_rel], x.name
        assert list(x.ua_kind) == list(y.ua_kind) and [tuple(t) for t in x.ua_rel] == [tuple(t) for t in y.ua_rel], x.name
        assert x.head_rel == y.head_rel and list(x.methyl_rel) == list(y.methyl_rel) and x.normal_head_rel == y.normal_head_rel, x.name
        assert list(x.bond_names) == list(y.bond_names), x.name


@needs_ref
def test_reference_classification_matches_python_restatement():
    """The C++ classifier against oracle/fixtures.py (pinned end to end by the reference's YAML fixtures) on CG, AA and UA."""
    from oracle import fixtures
    # CG: cg.tpr, beads = membrane, heads PO4 (tests_cg.rs)
    s = System.from_tpr(os.path.join(REF, "cg.tpr"))
    st = fixtures.read_gro(os.path.join(REF, "cg.gro"))
    fixtures.read_bnd(os.path.join(REF, "cg.bnd"), st)
    mem = st.select(lambda r, n: r in LIPIDS)
    heads = st.select(lambda r, n: r in LIPIDS and n == "PO4")
    meth = st.select(lambda r, n: r in LIPIDS and n in ("C4A", "C4B"))
    ref = fixtures.build_bond_setup(st, abi.KIND_CG, mem, mem, heads=heads, methyls=meth, normal_heads=heads).moltypes
    _same_types(s.classify_bonds(abi.KIND_CG, mem, mem, heads=heads, methyls=meth, normal_heads=heads), ref, abi.KIND_CG)
    # AA: pcpepg, heavy atoms = lipid carbons, hydrogens = lipid hydrogens bonded to them
    s = System.from_tpr(os.path.join(REF, "pcpepg.tpr"))
    st = fixtures.read_gro(os.path.join(REF, "pcpepg.gro"))
    fixtures.read_bnd(os.path.join(REF, "pcpepg.bnd"), st)
    heavy = st.select(lambda r, n: r in LIPIDS and n.startswith("C"))
    hyd = st.select(lambda r, n: r in LIPIDS and n.startswith("H"))
    heads = st.select(lambda r, n: r in LIPIDS and n == "P")
    meth = st.select(lambda r, n: r in LIPIDS and n in ("C218", "C316"))
    ref = fixtures.build_bond_setup(st, abi.KIND_AA, heavy, hyd, heads=heads, methyls=meth).moltypes
    _same_types(s.classify_bonds(abi.KIND_AA, heavy, hyd, heads=heads, methyls=meth), ref, abi.KIND_AA)
    z = s.atoms()[3]
    assert all(z[i] == 6 for i in heavy) and all(z[i] == 1 for i in hyd)   # atomic numbers of the TPR agree with the names
    # UA: ua.tpr with the selections of tests_ua.rs
    s = System.from_tpr(os.path.join(REF, "ua.tpr"))
    st = fixtures.read_pdb(os.path.join(REF, "ua_nobox.pdb"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    sat, unsat = mg.ua_selections(st)
    heads = st.select(lambda r, n: r in LIPIDS and n.startswith("P"))
    meth = st.select(lambda r, n: (r == "POPC" and n in ("CA2", "C50")) or (r == "POPS" and n in ("C36", "C55")))
    ref = fixtures.build_ua_setup(st, sat, unsat, heads=heads, methyls=meth).moltypes
    _same_types(s.classify_ua(sat, unsat, heads=heads, methyls=meth), ref, abi.KIND_UA)


def test_cg_asym_oracle():
    """The asymmetric CG membrane of the reference (tests_cg.rs:2182-2309), whose topology exists only as a TPR file: C++ TPR
    reader -> Master group -> C++ classifier -> oracle, against cg_order_asymmetric.yaml and ..._errors.yaml."""
    from oracle import oracle
    import golden_cases
    setup, xyz, box, cases = golden_cases.cg_asym()
    assert [m.name for m in setup.moltypes] == ["POPE", "POPG"] and setup.n_atoms == 279 * 12
    for name, case in cases.items():
        st = dataclasses.replace(setup, timewise="n_blocks" in case)
        o = oracle.Oracle(st)
        o.analyze_frames(xyz, box)
        golden_cases.assert_matches_yaml(o.finish(), st, case)


# ---- GRO / PDB / read_structure_and_topology ------------------------------------------------------------------------------
GRO = """toy
    4
    1LIP      P    1   1.000   2.000   3.000
    1LIP     C1    2   1.100   2.000   3.000
    1LIP     C2    3   1.200   2.000   3.000
    2SOL     OW    4   0.500   0.500   0.500
   4.00000   5.00000   6.00000
"""
PDB = """CRYST1   40.000   50.000   60.000  90.00  90.00  90.00 P 1           1
ATOM      1  P   LIP     1      10.000  20.000  30.000  1.00  0.00
ATOM      2  C1  LIP     1      11.000  20.000  30.000  1.00  0.00
ATOM      3  C2  LIP     1      12.000  20.000  30.000  1.00  0.00
ATOM      4  OW  SOL     2       5.000   5.000   5.000  1.00  0.00
ENDMDL
CONECT    1    2
CONECT    2    1    3
CONECT    3    2
"""


def test_structure_files_and_topology_rules(tmp_path):
    """structure.rs:27-88: a GRO file has no topology (NoTopology) unless a bonds file comes with it; a PDB brings CONECT
    records (none: NoTopology, repeated atom numbers: InvalidPdbTopology); unknown extensions are refused."""
    gro, pdb, bnd = tmp_path / "s.gro", tmp_path / "s.pdb", tmp_path / "s.bnd"
    gro.write_text(GRO); pdb.write_text(PDB); bnd.write_text("1 2\n2 3\n")
    with pytest.raises(abi.GorderError) as e:
        System.from_file(str(gro))
    assert e.value.code == abi.ERR_NO_TOPOLOGY
    s = System.from_file(str(gro), str(bnd))
    names, resn, resid, *_ = s.atoms()
    assert names == ["P", "C1", "C2", "OW"] and resn == ["LIP", "LIP", "LIP", "SOL"] and resid.tolist() == [1, 1, 1, 2]
    assert s.bonds().tolist() == [[0, 1], [1, 2]]
    assert np.allclose(s.positions()[1], [1.1, 2.0, 3.0]) and np.allclose(s.box9()[[0, 4, 8]], [4, 5, 6])
    p = System.from_file(str(pdb))
    assert p.atoms()[0] == names and p.bonds().tolist() == [[0, 1], [1, 2]]
    assert np.allclose(p.positions(), s.positions(), atol=1e-6) and np.allclose(p.box9()[[0, 4, 8]], [4, 5, 6])
    (tmp_path / "n.pdb").write_text("".join(ln + "\n" for ln in PDB.splitlines() if not ln.startswith("CONECT")))
    with pytest.raises(abi.GorderError) as e:
        System.from_file(str(tmp_path / "n.pdb"))
    assert e.value.code == abi.ERR_NO_TOPOLOGY
    assert System.from_file(str(tmp_path / "n.pdb"), str(bnd)).n_bonds == 2
    (tmp_path / "d.pdb").write_text(PDB.replace("ATOM      4", "ATOM      1"))
    with pytest.raises(abi.GorderError) as e:
        System.from_file(str(tmp_path / "d.pdb"))
    assert e.value.code == abi.ERR_PDB_TOPOLOGY
    for bad in ("s.xyz", "s"):
        (tmp_path / bad).write_text(GRO)
        with pytest.raises(abi.GorderError) as e:
            System.from_file(str(tmp_path / bad))
        assert e.value.code == abi.ERR_STRUCTURE_FORMAT
    (tmp_path / "t.gro").write_text("toy\n   9\n    1LIP      P    1   1.000\n")
    with pytest.raises(abi.GorderError) as e:
        System.from_file(str(tmp_path / "t.gro"), str(bnd))
    assert e.value.code == abi.ERR_STRUCTURE_FORMAT
    with pytest.raises(abi.GorderError) as e:
        System.from_file(str(tmp_path / "missing.gro"))
    assert e.value.code == abi.ERR_IO
    t = System.from_file(os.path.join(TPR, "cyclic.tpr"))
    assert t.n_atoms == 36 and t.tpx_version == 127
    t2 = System.from_file(os.path.join(TPR, "cyclic.tpr"), str(bnd))   # the bonds file wins over the run file's topology
    assert t2.n_bonds == 2


@needs_ref
def test_reference_gro_pdb_match_python_readers():
    from oracle import fixtures
    for gro, bnd in (("cg.gro", "cg.bnd"), ("pcpepg.gro", "pcpepg.bnd")):
        s = System.from_file(os.path.join(REF, gro), os.path.join(REF, bnd))
        st = fixtures.read_gro(os.path.join(REF, gro))
        fixtures.read_bnd(os.path.join(REF, bnd), st)
        names, resn, resid, *_ = s.atoms()
        assert names == st.name and resn == st.resname and np.array_equal(resid, st.resid)
        assert np.array_equal(s.positions(), st.xyz) and np.array_equal(s.box9()[[0, 4, 8]], st.box)
        assert list(map(tuple, s.bonds().tolist())) == st.bonds
    s = System.from_file(os.path.join(REF, "ua_nobox.pdb"))
    st = fixtures.read_pdb(os.path.join(REF, "ua_nobox.pdb"))
    names, resn, resid, *_ = s.atoms()
    assert names == st.name and resn == st.resname and np.array_equal(resid, st.resid)
    # Angstrom -> nm: an f32 division here, a double division rounded to f32 in the Python reader
    assert np.abs(s.positions() - st.xyz).max() < 1e-6 and list(map(tuple, s.bonds().tolist())) == st.bonds
    c = System.from_file(os.path.join(REF, "cg.pdb"), os.path.join(REF, "cg.bnd"))
    g = System.from_file(os.path.join(REF, "cg.gro"), os.path.join(REF, "cg.bnd"))
    assert c.atoms()[0] == g.atoms()[0] and np.abs(c.positions() - g.positions()).max() < 1e-3
    assert np.allclose(c.box9(), g.box9(), atol=1e-3)


def test_weird_molecules_fixture():
    """tests_aa.rs:1961-2035 (test_aa_order_maps_basic_weird_molecules): molecules that share a name and span several residues.
    The reference names its order-map files after the molecule types and bonds the classifier finds; those names, from
    multiple_resid_same_name.tpr with heavy atoms `name C1A C3A C1B C3B` and hydrogens `name D2A C4A C2B C4B`."""
    s = System.from_tpr(os.path.join(TPR, "multiple_resid_same_name.tpr"))
    names, resn, *_ = s.atoms()
    sel = lambda want: [i for i in range(s.n_atoms) if resn[i] in ("POPC", "POPE") and names[i] in want]
    mts = s.classify_bonds(abi.KIND_AA, sel(("C1A", "C3A", "C1B", "C3B")), sel(("D2A", "C4A", "C2B", "C4B")))
    got = {m.name: ["--".join(part.replace(" (", "-").replace(")", "").replace(" ", "-") for part in b.split(" - ")) for b in m.bond_names] for m in mts}
    assert got == {
        "POPC-POPE1": ["POPC-C1A-4--POPC-D2A-5", "POPC-D2A-5--POPE-C3A-6", "POPE-C3A-6--POPE-C4A-7", "POPE-C1B-8--POPE-C2B-9", "POPE-C2B-9--POPE-C3B-10",
                       "POPE-C3B-10--POPE-C4B-11"],
        "POPC-POPE2": ["POPC-C1A-4--POPC-D2A-5", "POPC-D2A-5--POPE-C3A-6", "POPE-C3A-6--POPE-C4A-7", "POPE-C3B-10--POPE-C4B-11"],
        "POPC": ["POPC-C1A-4--POPC-D2A-5", "POPC-D2A-5--POPC-C3A-6", "POPC-C3A-6--POPC-C4A-7", "POPC-C1B-8--POPC-C2B-9", "POPC-C2B-9--POPC-C3B-10",
                 "POPC-C3B-10--POPC-C4B-11"],
    }, got


def test_text_readers_fuzz(tmp_path):
    """GRO / PDB / bonds / ndx readers on damaged files: an error code or a result, never a crash."""
    from gorder_b200.structure import read_ndx
    rng = np.random.default_rng(11)
    texts = {"s.gro": GRO, "s.pdb": PDB, "s.ndx": "[ Upper ]\n1 2 3\n[ Lower ]\n4\n", "s.bnd": "1 2 3\n2 3\n# c\n4 1\n"}
    base = System.from_arrays(["A"] * 4, ["R"] * 4)
    for it in range(400):
        fn = list(texts)[it % 4]
        b = bytearray(texts[fn].encode())
        for _ in range(int(rng.integers(1, 6))):
            k = int(rng.integers(0, 3))
            if not b:
                break
            pos = int(rng.integers(0, len(b)))
            if k == 0:
                b[pos] = int(rng.integers(0, 256))
            elif k == 1:
                del b[pos:pos + int(rng.integers(1, 20))]
            else:
                b[pos:pos] = bytes(rng.integers(32, 127, int(rng.integers(1, 12)), dtype=np.uint8))
        p = tmp_path / fn
        p.write_bytes(bytes(b))
        try:
            if fn == "s.ndx":
                read_ndx(str(p), 4)
            elif fn == "s.bnd":
                base.read_bonds(str(p))
            else:
                s = System.from_file(str(p), None if fn == "s.pdb" else str(tmp_path / "ok.bnd") if (tmp_path / "ok.bnd").exists() else None)
                s.atoms(); s.bonds(); s.close()
        except abi.GorderError as e:
            assert e.code in (abi.ERR_STRUCTURE_FORMAT, abi.ERR_NO_TOPOLOGY, abi.ERR_PDB_TOPOLOGY, abi.ERR_NDX_PARSE, abi.ERR_BONDS_PARSE,
                              abi.ERR_BONDS_ATOM_NOT_FOUND, abi.ERR_BONDS_SELF, abi.ERR_IO), e
        except UnicodeDecodeError:
            pass   # a name that is no longer UTF-8 (the C side handed it over byte for byte)
