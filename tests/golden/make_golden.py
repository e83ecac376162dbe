#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference checkout.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py

Every fixture carries (a) inputs dug out of the reference's own test files (coordinates of the
Master group only, renumbered densely as the reference does, common.rs:92-103), (b) the engine
setup as JSON, and (c) the EXPECTED values copied from the reference: the hard-coded vectors of its
unit tests (cgorder.rs:188-241, aaorder.rs:226-350, uaorder.rs:1113-1200) or its YAML / map /
leaflet / normals fixtures (tests/files/ua_order_*.yaml, ordermaps_ua/*, ua_leaflets_once.yaml,
ua_normals.yaml).  Nothing in the expected values comes from this repository's own code.
"""
from __future__ import annotations

import json
import os
import re
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gorder_b200 import abi  # noqa: E402
from oracle import fixtures  # noqa: E402

REF = "/root/reference"
FILES = os.path.join(REF, "tests", "files")
LIPIDS = {"POPC", "POPE", "POPG", "POPS"}


def rust_vectors(path: str, fn: str):
    src = open(path).read()
    body = src[src.index("fn " + fn):]
    body = body[: body.index("\n    }\n")]
    return [np.array([float(x) for x in v.replace("\n", " ").split(",") if x.strip()]) for v in re.findall(r"vec!\[(.*?)\]", body, re.S)]


def single_frame(name: str, gro: str, bnd: str, tpr: str, kind: int, rust: str, sign: float, head: str):
    st = fixtures.read_gro(os.path.join(FILES, gro))
    fixtures.read_bnd(os.path.join(FILES, bnd), st)
    xyz, box, _ = fixtures.tpr_coordinates(os.path.join(FILES, tpr), st.xyz)
    st.xyz = xyz
    mem = st.select(lambda r, n: r in LIPIDS)
    cst, keep = fixtures.compact(st, mem)
    allm = np.arange(cst.n_atoms)
    if kind == abi.KIND_CG:
        g1 = g2 = allm
    else:   # "@membrane and element name carbon / hydrogen" (aaorder.rs:170-181)
        g1 = cst.select(lambda r, n: n.startswith("C"))
        g2 = cst.select(lambda r, n: n.startswith("H"))
    heads = cst.select(lambda r, n: n == head)
    setup = fixtures.build_bond_setup(cst, kind, g1, g2, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL)
    exp = {k: (sign * np.concatenate(rust_vectors(os.path.join(REF, "src", "analysis", rust), f"expected_{k}_orders"))).tolist()
           for k in ("total", "upper", "lower")}
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), xyz=cst.xyz.astype(np.float32), box=box.astype(np.float32),
                        setup=json.dumps(setup.to_dict()), expected=json.dumps(exp))
    print(name, cst.n_atoms, "atoms,", setup.n_slots, "bond types")


def ua_selections(st):
    sat_excl = {"POPC": {"C15", "C34", "C24", "C25"}, "POPS": {"C6", "C18", "C39", "C27", "C28"}}
    unsat_sel = {"POPC": {"C24", "C25"}, "POPS": {"C27", "C28"}}
    sat = st.select(lambda r, n: r in sat_excl and n.startswith("C") and n not in sat_excl[r])
    unsat = st.select(lambda r, n: r in unsat_sel and n in unsat_sel[r])
    return sat, unsat


def flatten_yaml(doc: dict, keys=("total",)):
    """Expected numbers of an order YAML in traversal order (system, then per molecule: average,
    per atom: atom value(s), per bond value(s)); {mean, error} dicts are expanded to two numbers."""
    out = []

    def push(node):
        for k in keys:
            if k in node:
                v = node[k]
                if isinstance(v, dict):
                    out.append(float(v["mean"]) if v["mean"] == v["mean"] else float("nan"))
                    out.append(float(v["error"]))
                else:
                    out.append(float(v))

    push(doc["average order"])
    for mol, body in doc.items():
        if mol == "average order":
            continue
        push(body["average order"])
        for _atom, node in body["order parameters"].items():
            push(node)
            bonds = node.get("bonds", [])
            for b in (bonds.values() if isinstance(bonds, dict) else bonds):   # AA: mapping hydrogen -> values; UA: list
                push(b)
    return out


def ua_golden():
    st = fixtures.read_pdb(os.path.join(FILES, "ua_nobox.pdb"))
    tpr_xyz, sbox, _ = fixtures.tpr_coordinates(os.path.join(FILES, "ua.tpr"), st.xyz)
    traj = fixtures.read_xtc(os.path.join(FILES, "ua.xtc"))
    mem = st.select(lambda r, n: r in LIPIDS)
    cst, keep = fixtures.compact(st, mem)
    q = np.round(traj.xyz[:, keep, :].astype(np.float64) * 1000.0).astype(np.int32)   # XTC precision 1000: exact
    assert np.array_equal((q.astype(np.float32) * np.float32(1.0 / 1000.0)), traj.xyz[:, keep, :]), "XTC coordinates are not k/1000"
    sat, unsat = ua_selections(cst)
    heads = cst.select(lambda r, n: n.startswith("P"))
    allm = np.arange(cst.n_atoms)
    base = fixtures.build_ua_setup(cst, sat, unsat)
    cases = {}

    def add(name, setup, yaml_file, keys=("total",), frames=None, **extra):
        doc = yaml.safe_load(open(os.path.join(FILES, yaml_file)))
        cases[name] = dict(setup=setup.to_dict(), expected=flatten_yaml(doc, keys), keys=list(keys),
                           frames=frames if frames is not None else list(range(traj.xyz.shape[0])), source=yaml_file, **extra)

    add("basic", base, "ua_order_basic.yaml")
    leaf = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL)
    add("leaflets_global", leaf, "ua_order_leaflets.yaml", keys=("total", "upper", "lower"))
    # Individual / Local give the same fixture (tests_ua.rs:147-211); methyl selection of that test:
    # "(resname POPC and name CA2 C50) or (resname POPS and name C36 C55)"
    methyls = cst.select(lambda r, n: (r == "POPC" and n in ("CA2", "C50")) or (r == "POPS" and n in ("C36", "C55")))
    ind = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, methyls=methyls, leaflet_mode=abi.LEAFLET_INDIVIDUAL)
    add("leaflets_individual", ind, "ua_order_leaflets.yaml", keys=("total", "upper", "lower"))
    loc = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=2.5)
    add("leaflets_local", loc, "ua_order_leaflets.yaml", keys=("total", "upper", "lower"))
    err = fixtures.build_ua_setup(cst, sat, unsat, timewise=True)
    add("error", err, "ua_order_error.yaml", n_blocks=5)
    errl = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL, timewise=True)
    add("error_leaflets", errl, "ua_order_leaflets_error.yaml", keys=("total", "upper", "lower"), n_blocks=5)
    # begin 199200 ps, end 199800 ps, step 3 (tests_ua.rs:301-333): 11 frames
    sel = [i for i, t in enumerate(traj.time) if 199200.0 <= t <= 199800.0][::3]
    assert len(sel) == 11
    bes = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL, step=3)
    add("begin_end_step", bes, "ua_order_begin_end_step.yaml", keys=("total", "upper", "lower"), frames=sel)
    cyl = fixtures.build_ua_setup(cst, sat, unsat, geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_BOX_CENTER,
                                  geom_dims=(2.5, float("-inf"), float("inf")), geom_axis=abi.AXIS_Z)
    add("cylinder_center", cyl, "ua_order_cylinder_center.yaml")
    cub = fixtures.build_ua_setup(cst, sat, unsat, geom_kind=abi.GEOM_CUBOID, geom_ref_kind=abi.GEOMREF_POINT, geom_ref_point=(1.5, 2.5, 0.0),
                                  geom_dims=(-1.0, 2.0, 0.0, 1.0, float("-inf"), float("inf")), structure_box=tuple(float(x) for x in sbox))
    add("cuboid_point", cub, "ua_order_cuboid_point.yaml")
    dyn = fixtures.build_ua_setup(cst, sat, unsat, normal_heads=heads, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0, collect_normals=True)
    ndoc = yaml.safe_load(open(os.path.join(FILES, "ua_normals.yaml")))
    add("dynamic_normals", dyn, "ua_order_dynamic_normals.yaml",
        normals={k: np.array(v, np.float32).tolist() for k, v in ndoc.items()})
    # saturated / unsaturated carbons only (tests_ua.rs:69-118)
    add("basic_saturated", fixtures.build_ua_setup(cst, sat, ()), "ua_order_basic_saturated.yaml")
    add("basic_unsaturated", fixtures.build_ua_setup(cst, (), unsat), "ua_order_basic_unsaturated.yaml")
    # tests_ua.rs:215-251 (spectral clustering of the P atoms names the leaflets the other way round): the Global
    # assignment with `flip`
    flip = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL, leaflet_flip=True)
    add("leaflets_flipped", flip, "ua_order_leaflets_flipped.yaml", keys=("total", "upper", "lower"))
    # the exported normals read back as manual normals (ManualMembraneNormal, normal.rs:259-298; the reference's own
    # round trip: tests_aa.rs:4964-5022) reproduce the dynamic-normal result to the 6 printed decimals of the file
    man = fixtures.build_ua_setup(cst, sat, unsat, normal_mode=abi.NORMAL_MANUAL)
    for mt in man.moltypes:
        mt.manual_normals = np.array(ndoc[mt.name], np.float32)
    add("manual_normals", man, "ua_order_dynamic_normals.yaml")
    # leaflet export, Once (bit-exact fixture)
    once = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL,
                                   leaflet_freq_kind=abi.FREQ_ONCE, collect_leaflets=True)
    ldoc = yaml.safe_load(open(os.path.join(FILES, "ua_leaflets_once.yaml")))
    cases["leaflets_once_export"] = dict(setup=once.to_dict(), frames=list(range(traj.xyz.shape[0])), source="ua_leaflets_once.yaml",
                                         leaflets={k: np.array(v, np.uint8).tolist() for k, v in ldoc.items()})
    # order maps (tests_ua.rs:352-415): bin 0.5 x 2.0, min_samples 5, auto span = box of the structure file
    msat = cst.select(lambda r, n: r == "POPC" and n in ("C50", "C20", "C13"))
    munsat = cst.select(lambda r, n: r == "POPC" and n == "C24")
    maps = fixtures.build_ua_setup(cst, msat, munsat, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=(0.5, 2.0),
                                   map_span_x=(0.0, float(sbox[0])), map_span_y=(0.0, float(sbox[1])))
    mexp = {}
    mdir = os.path.join(FILES, "ordermaps_ua")
    for fn in sorted(os.listdir(mdir)):
        if not fn.endswith("_full.dat") or "--" not in fn:
            continue
        rows = [ln.split() for ln in open(os.path.join(mdir, fn)) if ln[0] not in "#@$"]
        mexp[fn] = [[float(a), float(b), float(c)] for a, b, c in rows]
    cases["maps_basic"] = dict(setup=maps.to_dict(), frames=list(range(traj.xyz.shape[0])), source="ordermaps_ua/*_full.dat",
                               maps=mexp, map_min_samples=5)
    np.savez_compressed(os.path.join(HERE, "ua_traj.npz"), q=q, box=traj.box.astype(np.float32), time=traj.time,
                        structure_box=sbox.astype(np.float32), cases=json.dumps(cases))
    print("ua_traj", q.shape, "cases:", list(cases))
    # hydrogen goldens (uaorder.rs:1113-1200): atom indices are absolute in ua.tpr
    hyd = {
        "ch2": dict(kind=abi.UA_CH2, idx=[39, 38, 40], gold=[[2.3435528, 2.1503785, 2.1272178], [2.35857, 2.3045487, 2.039533]]),
        "ch3": dict(kind=abi.UA_CH3, idx=[49, 48, 47], gold=[[3.3708375, 2.7527616, 2.257202], [3.254057, 2.8633823, 2.3334126], [3.3182635, 2.8995805, 2.1713943]]),
        "ch1_unsat": dict(kind=abi.UA_CH1_UNSAT, idx=[23, 22, 24], gold=[[1.0985602, 2.994375, 2.7727659]]),
        "ch1_sat": dict(kind=abi.UA_CH1_SAT, idx=[12, 11, 31, 13], gold=[[1.5022101, 2.6938448, 1.7839708]]),
    }
    for v in hyd.values():
        v["atoms"] = [[float(np.float32(c)) for c in tpr_xyz[i]] for i in v["idx"]]
        v["atoms_hex"] = [[np.float32(c).tobytes().hex() for c in tpr_xyz[i]] for i in v["idx"]]
    json.dump(dict(box=[float(x) for x in sbox], box_hex=[np.float32(x).tobytes().hex() for x in sbox], cases=hyd),
              open(os.path.join(HERE, "ua_hydrogens.json"), "w"), indent=1)
    print("ua_hydrogens ok")


def aa_traj_golden():
    """AA end to end: tests_aa.rs:1019-1040 (test_aa_order_leaflets_yaml_supershort): pcpepg + pcpepg_selected.xtc
    (4 frames, precision 100), Global leaflets (@membrane, name P) -> tests/files/aa_order_selected.yaml."""
    st = fixtures.read_gro(os.path.join(FILES, "pcpepg.gro"))
    fixtures.read_bnd(os.path.join(FILES, "pcpepg.bnd"), st)
    traj = fixtures.read_xtc(os.path.join(FILES, "pcpepg_selected.xtc"))
    assert traj.xyz.shape[1] == st.n_atoms
    mem = st.select(lambda r, n: r in LIPIDS)
    cst, keep = fixtures.compact(st, mem)
    prec = 100.0
    q = np.round(traj.xyz[:, keep, :].astype(np.float64) * prec).astype(np.int32)
    assert np.array_equal(q.astype(np.float32) * np.float32(1.0 / prec), traj.xyz[:, keep, :]), "XTC coordinates are not k/100"
    allm = np.arange(cst.n_atoms)
    g1 = cst.select(lambda r, n: n.startswith("C"))
    g2 = cst.select(lambda r, n: n.startswith("H"))
    heads = cst.select(lambda r, n: n == "P")
    setup = fixtures.build_bond_setup(cst, abi.KIND_AA, g1, g2, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL)
    doc = yaml.safe_load(open(os.path.join(FILES, "aa_order_selected.yaml")))
    keys = ("total", "upper", "lower")
    np.savez_compressed(os.path.join(HERE, "aa_traj.npz"), q=q, precision=np.float32(prec), box=traj.box.astype(np.float32),
                        setup=json.dumps(setup.to_dict()), expected=json.dumps(flatten_yaml(doc, keys)), keys=json.dumps(list(keys)),
                        molecules=json.dumps([k for k in doc if k != "average order"]))
    print("aa_traj", q.shape, setup.n_slots, "bond types,", [m.name for m in setup.moltypes])


def ua_nopbc_golden():
    """UA without periodic boundary conditions (tests_ua.rs:688-714): ua_nobox.pdb + ua_whole_nobox.xtc (whole molecules, no
    box), Global leaflets, handle_pbc(false) -> ua_order_leaflets_nopbc.yaml."""
    st = fixtures.read_pdb(os.path.join(FILES, "ua_nobox.pdb"))
    traj = fixtures.read_xtc(os.path.join(FILES, "ua_whole_nobox.xtc"))
    assert traj.xyz.shape[1] == st.n_atoms
    mem = st.select(lambda r, n: r in LIPIDS)
    cst, keep = fixtures.compact(st, mem)
    prec = float(traj.precision)
    q = np.round(traj.xyz[:, keep, :].astype(np.float64) * prec).astype(np.int32)
    assert np.array_equal(q.astype(np.float32) * np.float32(1.0 / prec), traj.xyz[:, keep, :]), "XTC coordinates are not k/precision"
    sat, unsat = ua_selections(cst)
    heads = cst.select(lambda r, n: n.startswith("P"))
    setup = fixtures.build_ua_setup(cst, sat, unsat, heads=heads, membrane=np.arange(cst.n_atoms), leaflet_mode=abi.LEAFLET_GLOBAL, handle_pbc=False)
    doc = yaml.safe_load(open(os.path.join(FILES, "ua_order_leaflets_nopbc.yaml")))
    keys = ("total", "upper", "lower")
    case = dict(setup=setup.to_dict(), expected=flatten_yaml(doc, keys), keys=list(keys), source="ua_order_leaflets_nopbc.yaml")
    np.savez_compressed(os.path.join(HERE, "ua_nopbc.npz"), precision=np.float32(prec), box=np.zeros((q.shape[0], 3), np.float32),
                        case=json.dumps(case), **pack_lattice(q))
    print("ua_nopbc", q.shape, "precision", prec)


def concatenated(base: str, n: int = 5):
    """The reference ships its AA / CG test trajectories only in pieces (tests/files/split/<base>1..5.xtc: the inputs of
    the concatenation tests, tests_aa.rs:48-78, tests_cg.rs:45-76); joined with the duplicated boundary frames dropped
    (what traj_iter_cat_map_reduce does) they are the full pcpepg.xtc / cg.xtc every other fixture was made from."""
    xs, bs, ts = [], [], []
    last = None
    for i in range(1, n + 1):
        t = fixtures.read_xtc(os.path.join(FILES, "split", f"{base}{i}.xtc"))
        lo = 1 if last is not None and t.time[0] == last else 0
        last = t.time[-1]
        xs.append(t.xyz[lo:]); bs.append(t.box[lo:]); ts.append(np.asarray(t.time, np.float64)[lo:])
    return np.concatenate(xs), np.concatenate(bs), np.concatenate(ts)


def pack_lattice(q: np.ndarray) -> dict:
    """int16 frame-to-frame differences of the lattice coordinates (they compress 3x better than the coordinates)."""
    d = np.diff(q.astype(np.int32), axis=0, prepend=np.zeros((1,) + q.shape[1:], np.int32))
    assert np.abs(d).max() < 32768
    return dict(dq=d.astype(np.int16))


def full_traj_golden(name: str, base: str, gro: str, bnd: str, kind: int, head: str, methyl_names, n_expected: int):
    st = fixtures.read_gro(os.path.join(FILES, gro))
    fixtures.read_bnd(os.path.join(FILES, bnd), st)
    xyz, box, time = concatenated(base)
    assert xyz.shape[0] == n_expected and xyz.shape[1] == st.n_atoms, xyz.shape
    mem = st.select(lambda r, n: r in LIPIDS)
    cst, keep = fixtures.compact(st, mem)
    prec = 100.0
    q = np.round(xyz[:, keep, :].astype(np.float64) * prec).astype(np.int32)
    assert np.array_equal(q.astype(np.float32) * np.float32(1.0 / prec), xyz[:, keep, :]), "XTC coordinates are not k/100"
    allm = np.arange(cst.n_atoms)
    if kind == abi.KIND_CG:
        g1 = g2 = allm
    else:
        g1 = cst.select(lambda r, n: n.startswith("C"))
        g2 = cst.select(lambda r, n: n.startswith("H"))
    heads = cst.select(lambda r, n: n == head)
    methyls = cst.select(lambda r, n: n in methyl_names)
    cases = {}

    def add(case, yaml_file, keys=("total",), frames=None, **kw):
        extra = {k: kw.pop(k) for k in ("n_blocks", "min_samples", "tol") if k in kw}
        setup = fixtures.build_bond_setup(cst, kind, g1, g2, **kw)
        doc = yaml.safe_load(open(os.path.join(FILES, yaml_file)))
        cases[case] = dict(setup=setup.to_dict(), expected=flatten_yaml(doc, keys), keys=list(keys),
                           frames=frames if frames is not None else list(range(xyz.shape[0])), source=yaml_file, **extra)

    tul = ("total", "upper", "lower")
    glob = dict(heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_GLOBAL)
    pre = "aa" if kind == abi.KIND_AA else "cg"
    add("basic", f"{pre}_order_basic.yaml")
    add("leaflets_global", f"{pre}_order_leaflets.yaml", tul, **glob)
    add("leaflets_individual", f"{pre}_order_leaflets.yaml", tul, heads=heads, methyls=methyls, leaflet_mode=abi.LEAFLET_INDIVIDUAL)
    add("leaflets_local", f"{pre}_order_leaflets.yaml", tul, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=2.5)
    add("leaflets_every5", f"{pre}_order_leaflets.yaml", tul, leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=5, **glob)
    add("leaflets_once", f"{pre}_order_leaflets.yaml", tul, leaflet_freq_kind=abi.FREQ_ONCE, **glob)
    add("error", f"{pre}_order_error.yaml", n_blocks=5, timewise=True)
    add("error_leaflets", f"{pre}_order_error_leaflets.yaml", tul, n_blocks=5, timewise=True, **glob)

    # leaflets read from files (ManualClassification, leaflets.rs:816-874; 1 / Upper = upper leaflet)
    def add_manual(case, table_file, yaml_file, frames=None, **kw):
        add(case, yaml_file, tul, frames=frames, leaflet_mode=abi.LEAFLET_MANUAL, **kw)
        tab = yaml.safe_load(open(os.path.join(FILES, "inputs", "leaflets_files", table_file)))
        for md in cases[case]["setup"]["moltypes"]:
            md["manual_leaflets"] = [[1 if x in (1, "Upper", "upper") else 0 for x in row] for row in tab[md["name"]]]

    if kind == abi.KIND_AA:
        # begin 450 200 ps, end 450 400 ps (, step 3): tests_aa.rs:1202-1232, 1398-1423
        sel = [i for i, t in enumerate(time) if 450200.0 <= t <= 450400.0]
        add("begin_end", "aa_order_begin_end.yaml", tul, frames=sel, **glob)
        add("begin_end_step", "aa_order_begin_end_step.yaml", tul, frames=sel[::3], step=3, **glob)
        add("limit", "aa_order_limit.yaml", min_samples=2000)                                   # tests_aa.rs:1099-1120
        add("leaflets_limit", "aa_order_leaflets_limit.yaml", tul, min_samples=500, **glob)     # tests_aa.rs:1123-1149
        add("sphere_center", "aa_order_sphere_center.yaml", geom_kind=abi.GEOM_SPHERE, geom_ref_kind=abi.GEOMREF_BOX_CENTER,
            geom_dims=(2.5,))                                                                   # tests_aa.rs:3239-3260
        add("error_blocks10", "aa_order_error_blocks10.yaml", n_blocks=10, timewise=True)        # tests_aa.rs:2530-2552
        add("error_limit", "aa_order_error_limit.yaml", n_blocks=5, timewise=True, min_samples=2000)             # tests_aa.rs:2444-2477
        add("error_leaflets_limit", "aa_order_error_leaflets_limit.yaml", tul, n_blocks=5, timewise=True, min_samples=500, **glob)   # :2480-2527
        # step 5, leaflets assigned on every analysed frame (real frequency = 1 x step): tests_aa.rs:1307-1346
        add("step5_leaflets", "aa_order_step.yaml", tul, frames=list(range(0, xyz.shape[0], 5)), step=5, leaflet_freq_kind=abi.FREQ_EVERY,
            leaflet_freq=5, **glob)
        # convergence of the molecule averages (prefix averages over frames): tests_aa.rs:2580-2660
        def read_xvg(fn):
            return [[float(x) for x in ln.split()] for ln in open(os.path.join(FILES, fn)) if ln[0] not in "#@"]
        add("convergence", "aa_order_error.yaml", n_blocks=5, timewise=True)
        cases["convergence"].update(convergence=read_xvg("aa_order_convergence.xvg"))
        add("convergence_leaflets", "aa_order_error_leaflets.yaml", tul, n_blocks=5, timewise=True, **glob)
        cases["convergence_leaflets"].update(convergence=read_xvg("aa_order_leaflets_convergence.xvg"))
        # exported leaflet tables (bit-exact fixtures; every method writes the same file): tests_aa.rs:588-720
        def add_export(case, yaml_file, **kw):
            add(case, "aa_order_leaflets.yaml", tul, collect_leaflets=True, **kw)
            ldoc = yaml.safe_load(open(os.path.join(FILES, yaml_file)))
            cases[case].update(leaflets={k: np.array(v, np.uint8).tolist() for k, v in ldoc.items()}, leaflet_source=yaml_file)

        add_export("export_once_global", "aa_leaflets_once.yaml", leaflet_freq_kind=abi.FREQ_ONCE, **glob)
        add_export("export_every5_local", "aa_leaflets_every5.yaml", heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_LOCAL, leaflet_radius=2.5,
                   leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=5)
        add_export("export_every1_individual", "aa_leaflets_every1.yaml", heads=heads, methyls=methyls, leaflet_mode=abi.LEAFLET_INDIVIDUAL)
        add_export("export_every1_global", "aa_leaflets_every1.yaml", **glob)
        # leaflets read from files: tests_aa.rs:3619-3893
        add_manual("manual_once", "pcpepg_once.yaml", "aa_order_leaflets.yaml", leaflet_freq_kind=abi.FREQ_ONCE)
        add_manual("manual_every10", "pcpepg_every10.yaml", "aa_order_leaflets.yaml", leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=10)
        add_manual("manual_every", "pcpepg_every.yaml", "aa_order_leaflets.yaml", leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=1)
        # every(2) x step 5 = every 10th trajectory frame (leaflets.rs:261-262)
        add_manual("manual_every10_stepping", "pcpepg_every10.yaml", "aa_order_step.yaml", frames=list(range(0, xyz.shape[0], 5)), step=5,
                   leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=10)
        add_manual("manual_begin_end_step", "pcpepg_every_begin_end_step.yaml", "aa_order_begin_end_step.yaml", frames=sel[::3], step=3,
                   leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=3)
        # geometry selections: reference = centre of residue 1 (PBC centre of geometry of a group: the refined Bai-Breen
        # estimate, otherwise pinned only through leaflets), box centre, fixed point; inverted: tests_aa.rs:3183-3345, 3508-3615
        inf = (float("-inf"), float("inf"))
        res1 = np.array([i for i in range(cst.n_atoms) if cst.resid[i] == 1], np.int64)
        assert 50 < len(res1) < 200
        dyn = dict(geom_ref_kind=abi.GEOMREF_SELECTION, geom_ref=res1)
        add("cuboid_dynamic", "aa_order_cuboid_dynamic.yaml", geom_kind=abi.GEOM_CUBOID, geom_dims=(-1.0, 3.0, 1.0, 4.0, -3.0, 3.0), **dyn)
        add("cylinder_dynamic", "aa_order_cylinder_dynamic.yaml", geom_kind=abi.GEOM_CYLINDER, geom_dims=(2.1,) + inf, geom_axis=abi.AXIS_Y, **dyn)
        # tol: ONE of the 159 921 samples lies within an ulp of the sphere's surface and falls on the other side with the oracle's
        # summation order of the group centre (a 900-sample bond moves by 3.4e-4; the reference's own comparator allows 2e-4)
        add("sphere_dynamic", "aa_order_sphere_dynamic.yaml", geom_kind=abi.GEOM_SPHERE, geom_dims=(2.5,), tol=5e-4, **dyn)
        add("sphere_dynamic_inverted", "aa_order_sphere_dynamic_inverted.yaml", geom_kind=abi.GEOM_SPHERE, geom_dims=(2.5,), geom_invert=True, **dyn)
        cen = dict(geom_ref_kind=abi.GEOMREF_BOX_CENTER)
        add("cuboid_patch", "aa_order_cuboid_patch.yaml", geom_kind=abi.GEOM_CUBOID, geom_dims=(-1.0, 3.0) + inf + inf, **cen)
        add("cylinder_x", "aa_order_cylinder_x.yaml", geom_kind=abi.GEOM_CYLINDER, geom_dims=(3.0, -1.0, 3.0), geom_axis=abi.AXIS_X, **cen)
        add("cylinder_z_inverted", "aa_order_cylinder_z_inverted.yaml", geom_kind=abi.GEOM_CYLINDER, geom_dims=(3.0,) + inf, geom_axis=abi.AXIS_Z,
            geom_invert=True, **cen)
        _, sbox0, _ = fixtures.tpr_coordinates(os.path.join(FILES, "pcpepg.tpr"), st.xyz)
        add("cuboid_square_inverted", "aa_order_cuboid_square_inverted.yaml", geom_kind=abi.GEOM_CUBOID, geom_ref_kind=abi.GEOMREF_POINT,
            geom_ref_point=(8.0, 2.0, 0.0), geom_dims=(-2.0, 4.0, -4.0, 1.0) + inf, geom_invert=True, structure_box=tuple(float(x) for x in sbox0))
        # sphere around a fixed point: tests_aa.rs:3154-3180
        add("sphere_static", "aa_order_sphere_static.yaml", geom_kind=abi.GEOM_SPHERE, geom_ref_kind=abi.GEOMREF_POINT, geom_ref_point=(8.0, 2.0, 4.5),
            geom_dims=(2.5,), structure_box=tuple(float(x) for x in sbox0))
        # dynamic PCA normals (P atoms, 2 nm) + Individual leaflets assigned once: tests_aa.rs:4774-4806
        add("leaflets_dynamic", "aa_order_leaflets_dynamic.yaml", tul, heads=heads, methyls=methyls, leaflet_mode=abi.LEAFLET_INDIVIDUAL,
            leaflet_freq_kind=abi.FREQ_ONCE, normal_heads=heads, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0)
        # order maps of three POPC carbons (+ static geometry): tests_aa.rs:1560-1624, 3023-3087, 3090-3153
        _, sbox, _ = fixtures.tpr_coordinates(os.path.join(FILES, "pcpepg.tpr"), st.xyz)
        g1 = cst.select(lambda r, n: r == "POPC" and n in ("C22", "C24", "C218"))

        def add_maps(case, yaml_file, mdir, bin_, min_samples, keys=("total",), **kw):
            add(case, yaml_file, keys, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=bin_, map_span_x=(0.0, float(sbox[0])),
                map_span_y=(0.0, float(sbox[1])), **kw)
            mexp = {}
            ends = ("_full.dat",) if len(keys) == 1 else ("_full.dat", "_upper.dat", "_lower.dat")
            for fn in sorted(os.listdir(os.path.join(FILES, mdir))):
                if fn.endswith(ends) and "average" not in fn:
                    rows = [ln.split() for ln in open(os.path.join(FILES, mdir, fn)) if ln[0] not in "#@$"]
                    mexp[fn] = [[float(a), float(b), float(c)] for a, b, c in rows]
            cases[case].update(maps=mexp, map_min_samples=min_samples)

        add_maps("maps_basic", "aa_order_small.yaml", "ordermaps", (0.1, 4.0), 5)
        add_maps("maps_leaflets", "aa_order_leaflets_small.yaml", "ordermaps", (0.1, 4.0), 5, tul, **glob)   # tests_aa.rs:1627-1730
        add_maps("maps_cuboid_square", "aa_order_cuboid_square.yaml", "ordermaps_cuboid", (0.5, 0.5), 5, geom_kind=abi.GEOM_CUBOID,
                 geom_ref_kind=abi.GEOMREF_POINT, geom_ref_point=(8.0, 2.0, 0.0), geom_dims=(-2.0, 4.0, -4.0, 1.0, float("-inf"), float("inf")),
                 structure_box=tuple(float(x) for x in sbox))
        add_maps("maps_cylinder", "aa_order_cylinder.yaml", "ordermaps_cylinder", (0.5, 0.5), 1, geom_kind=abi.GEOM_CYLINDER,
                 geom_ref_kind=abi.GEOMREF_POINT, geom_ref_point=(8.0, 2.0, 0.0), geom_dims=(2.5, float("-inf"), float("inf")), geom_axis=abi.AXIS_Z,
                 structure_box=tuple(float(x) for x in sbox))
    else:
        # begin 352 000 ps, end 358 000 ps, step 5: tests_cg.rs:746-772
        sel = [i for i, t in enumerate(time) if 352000.0 <= t <= 358000.0][::5]
        add("begin_end_step", "cg_order_begin_end_step.yaml", tul, frames=sel, step=5, **glob)
        # geometry selections (tests_cg.rs:2468-2553, 2682-2713) and min_samples limits (tests_cg.rs:620-665)
        inf = (float("-inf"), float("inf"))
        _, sbox_cg, _ = fixtures.tpr_coordinates(os.path.join(FILES, "cg.tpr"), st.xyz)
        sb = tuple(float(x) for x in sbox_cg)
        res1 = np.array([i for i in range(cst.n_atoms) if cst.resid[i] == 1], np.int64)
        add("cuboid_square", "cg_order_cuboid_square.yaml", geom_kind=abi.GEOM_CUBOID, geom_ref_kind=abi.GEOMREF_BOX_CENTER,
            geom_dims=(-8.0, -2.0, 2.0, 8.0) + inf)
        add("cylinder", "cg_order_cylinder.yaml", geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_POINT, geom_ref_point=(2.0, 1.0, 0.0),
            geom_dims=(3.25,) + inf, geom_axis=abi.AXIS_Z, structure_box=sb)
        add("sphere_dynamic", "cg_order_sphere.yaml", geom_kind=abi.GEOM_SPHERE, geom_ref_kind=abi.GEOMREF_SELECTION, geom_ref=res1, geom_dims=(2.5,))
        add("cylinder_z_inverted", "cg_order_cylinder_z_inverted.yaml", geom_kind=abi.GEOM_CYLINDER, geom_ref_kind=abi.GEOMREF_POINT,
            geom_ref_point=(3.0, 3.0, 3.0), geom_dims=(4.0,) + inf, geom_axis=abi.AXIS_Z, geom_invert=True, structure_box=sb)
        add("limit", "cg_order_limit.yaml", min_samples=5000)
        add("leaflets_limit", "cg_order_leaflets_limit.yaml", tul, min_samples=2000, **glob)
        # leaflets read from files: tests_cg.rs:2719-2876; a table with too few rows (every 16 frames, 6 rows): tests_cg.rs:3054-3074
        add_manual("manual_once", "cg_once.yaml", "cg_order_leaflets.yaml", leaflet_freq_kind=abi.FREQ_ONCE)
        add_manual("manual_every20", "cg_every20.yaml", "cg_order_leaflets.yaml", leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=20)
        add_manual("manual_every", "cg_every.yaml", "cg_order_leaflets.yaml", leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=1)
        add_manual("manual_not_enough_frames", "cg_every20.yaml", "cg_order_leaflets.yaml", leaflet_freq_kind=abi.FREQ_EVERY, leaflet_freq=16)
        cases["manual_not_enough_frames"].update(expect_error=abi.ERR_MANUAL_LEAFLET_FRAME)
        add("error_limit", "cg_order_error_limit.yaml", n_blocks=5, timewise=True, min_samples=5000)                     # tests_cg.rs:1612-1641
        add("error_leaflets_limit", "cg_order_error_leaflets_limit.yaml", tul, n_blocks=5, timewise=True, min_samples=2000, **glob)   # :1644-1690
        # begin 352 000 ps, end 358 000 ps: 61 frames (tests_cg.rs:893-915)
        sel61 = [i for i, t in enumerate(time) if 352000.0 <= t <= 358000.0]
        assert len(sel61) == 61
        add("begin_end", "cg_order_begin_end.yaml", tul, frames=sel61, **glob)
        # only the lipids of the upper leaflet are analysed ("resid 1 to 254"), leaflets assigned once: tests_cg.rs:206-236
        gsave = (g1, g2)
        g1 = g2 = np.array([i for i in range(cst.n_atoms) if 1 <= cst.resid[i] <= 254], np.int64)
        add("leaflets_only_upper", "cg_order_leaflets_only_upper.yaml", tul, leaflet_freq_kind=abi.FREQ_ONCE, **glob)
        add("leaflets_only_upper_individual", "cg_order_leaflets_only_upper.yaml", tul, heads=heads, methyls=methyls,
            leaflet_mode=abi.LEAFLET_INDIVIDUAL, leaflet_freq_kind=abi.FREQ_ONCE)
        add("leaflets_only_upper_local", "cg_order_leaflets_only_upper.yaml", tul, heads=heads, membrane=allm, leaflet_mode=abi.LEAFLET_LOCAL,
            leaflet_radius=2.5, leaflet_freq_kind=abi.FREQ_ONCE)
        g1, g2 = gsave
        # order maps of the POPC B-chain bonds, bin 1 x 1 nm, min_samples 10 (tests_cg.rs:1040-1180)
        sbox = sbox_cg
        gsave = (g1, g2)
        g1 = g2 = cst.select(lambda r, n: r == "POPC" and n in ("C1B", "C2B", "C3B", "C4B"))

        def add_maps(case, yaml_file, mdir, bin_, min_samples, keys=("total",), **kw):
            add(case, yaml_file, keys, map_enabled=True, map_plane=abi.PLANE_XY, map_bin=bin_, map_span_x=(0.0, float(sbox[0])),
                map_span_y=(0.0, float(sbox[1])), **kw)
            mexp = {}
            ends = ("_full.dat",) if len(keys) == 1 else ("_full.dat", "_upper.dat", "_lower.dat")
            for fn in sorted(os.listdir(os.path.join(FILES, mdir))):
                if fn.endswith(ends) and "average" not in fn:
                    rows = [ln.split() for ln in open(os.path.join(FILES, mdir, fn)) if ln[0] not in "#@$"]
                    mexp[fn] = [[float(a), float(b), float(c)] for a, b, c in rows]
            cases[case].update(maps=mexp, map_min_samples=min_samples)

        add_maps("maps_basic", "cg_order_small.yaml", "ordermaps_cg", (1.0, 1.0), 10)
        add_maps("maps_leaflets", "cg_order_leaflets_small.yaml", "ordermaps_cg", (1.0, 1.0), 10, tul, **glob)
        g1, g2 = gsave
        # bonds redefined by a bond file: a few lipids lose / gain a bond and become molecule types of their own, named
        # POPE1, POPG1, POPG2, POPE2, POPE3 (classify.rs:262-294): tests_cg.rs:382-406
        st2 = fixtures.read_gro(os.path.join(FILES, gro))
        fixtures.read_bnd(os.path.join(FILES, "cg_redefined.bnd"), st2)
        cst2, keep2 = fixtures.compact(st2, mem)
        assert np.array_equal(keep, keep2)
        doc = yaml.safe_load(open(os.path.join(FILES, "cg_order_redefined_bonds.yaml")))
        all2 = np.arange(cst2.n_atoms)
        setup2 = fixtures.build_bond_setup(cst2, kind, all2, all2)
        cases["redefined_bonds"] = dict(setup=setup2.to_dict(), expected=flatten_yaml(doc, ("total",)), keys=["total"],
                                        frames=list(range(xyz.shape[0])), source="cg_order_redefined_bonds.yaml")
        # Individual leaflets assigned once + dynamic normals (PO4, 2 nm): tests_cg.rs:3356-3388
        add("leaflets_dynamic", "cg_order_leaflets_dynamic.yaml", tul, heads=heads, methyls=methyls, leaflet_mode=abi.LEAFLET_INDIVIDUAL,
            leaflet_freq_kind=abi.FREQ_ONCE, normal_heads=heads, normal_mode=abi.NORMAL_DYNAMIC, dynamic_radius=2.0)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), precision=np.float32(prec), box=box.astype(np.float32), time=time.astype(np.float64),
                        cases=json.dumps(cases), **pack_lattice(q))
    print(name, q.shape, "cases:", list(cases))


def xtc_fixtures():
    """GROMACS-written trajectories of the reference's test tree, as they are (tests/golden/xtc/): the parity tests of the
    XTC readers (host and device) must not only see streams of this repository's own writer.  Small files whole; of ua.xtc
    (3.7 MB) the first five frames -- a byte prefix of an XTC file is the file of its first frames."""
    import shutil
    import struct
    out = os.path.join(HERE, "xtc")
    os.makedirs(out, exist_ok=True)
    for rel in ("pcpepg_selected.xtc", "split/cg3.xtc", "split/pcpepg4.xtc", "multiple_resid_same_name.xtc"):
        shutil.copyfile(os.path.join(FILES, rel), os.path.join(out, os.path.basename(rel)))
    raw = open(os.path.join(FILES, "ua.xtc"), "rb").read()
    pos, n = 0, 0
    while n < 5:   # frame: 56-byte header, 32 bytes of compression parameters, length, padded stream
        assert struct.unpack(">i", raw[pos:pos + 4])[0] == 1995
        nbytes = struct.unpack(">i", raw[pos + 88:pos + 92])[0]
        pos += 92 + (nbytes + 3) // 4 * 4
        n += 1
    open(os.path.join(out, "ua_first5.xtc"), "wb").write(raw[:pos])
    print("xtc fixtures:", sorted(os.listdir(out)))


def normals_planar():
    """The reference's unit test of membrane_normal_from_cloud (normal.rs:664-963): positions of the 274 P atoms of
    pcpepg.tpr, its box, and the 274 expected SIGNED normals copied from the Rust source."""
    st = fixtures.read_gro(os.path.join(FILES, "pcpepg.gro"))
    xyz, box, _ = fixtures.tpr_coordinates(os.path.join(FILES, "pcpepg.tpr"), st.xyz)
    p = np.array([i for i in range(st.n_atoms) if st.name[i] == "P"])
    src = open(os.path.join(REF, "src", "analysis", "normal.rs")).read()
    body = src[src.index("fn test_real_planar"):src.index("fn test_real_vesicle")]
    exp = np.array([[float(x) for x in m] for m in re.findall(r"Vector3D::new\(\s*([-0-9.e]+),\s*([-0-9.e]+),\s*([-0-9.e]+)\s*,?\s*\)", body)], np.float32)
    assert exp.shape == (len(p), 3)
    np.savez_compressed(os.path.join(HERE, "normals_planar.npz"), P=xyz[p].astype(np.float32), box=np.asarray(box, np.float32), expected=exp)
    print("normals_planar", exp.shape)


def tpr_golden():
    """tests/golden/tpr/: small GROMACS-written TPR files of the reference's test tree as they are (tpx 103, 122, 127) for the
    C++ reader of csrc/gorder_topology.inl, `expected.json` with what they hold, and cg_asym.npz -- the trajectory and the
    reference's expected order parameters of its asymmetric CG membrane (tests_cg.rs:2182-2309), a fixture whose topology
    exists ONLY as a TPR file.  expected.json is written from the C++ reader after it has been checked here (a) against the
    independent Python parser tests/golden/tpr_independent.py on every file and (b) against the reference's .gro / .bnd / .pdb
    files for the three systems that have them; on the GPU box, where the reference tree is absent, it is a regression pin."""
    import shutil
    import zlib
    import importlib.util
    from gorder_b200.structure import System
    spec = importlib.util.spec_from_file_location("tpr_independent", os.path.join(HERE, "tpr_independent.py"))
    ind = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ind)
    out = os.path.join(HERE, "tpr")
    os.makedirs(out, exist_ok=True)

    def independent(path):
        d = ind.parse(path, False)
        names, resn, bonds, off = [], [], set(), 0
        btypes = ["BONDS", "G96BONDS", "MORSE", "CUBICBONDS", "CONNBONDS", "HARMONIC", "FENEBONDS", "TABBONDS", "TABBONDSNC", "RESTRBONDS", "CONSTR", "CONSTRNC"]
        for (t, nmol, nat) in d["mbs"]:
            mname, atoms, an, res, il = d["mts"][t]
            for _ in range(nmol):
                for k in btypes:
                    if k in il:
                        for _, i, j in il[k].reshape(-1, 3):
                            bonds.add((min(i, j) + off, max(i, j) + off))
                if "SETTLE" in il:
                    for _, o, h1, h2 in il["SETTLE"].reshape(-1, 4):
                        bonds.add((o + off, h1 + off)); bonds.add((o + off, h2 + off))
                for a in range(nat):
                    names.append(an[a]); resn.append(res[atoms[a][2]][0])
                off += nat
        return names, resn, sorted(bonds), d["x"], d["box"]

    exp = {}
    every = ["cg.tpr", "pcpepg.tpr", "ua.tpr", "asymmetric/cg_asym.tpr", "asymmetric/aa_asym.tpr", "cyclic.tpr", "scrambling/cg_scrambling.tpr",
             "cg_buckled.tpr", "pepg_cg.tpr", "multiple_resid.tpr", "multiple_resid_same_name.tpr", "same_name.tpr"]
    keep = {"asymmetric/cg_asym.tpr", "cyclic.tpr", "pepg_cg.tpr", "multiple_resid.tpr", "multiple_resid_same_name.tpr", "same_name.tpr"}
    for rel in every:
        path = os.path.join(FILES, rel)
        s = System.from_tpr(path)
        names, resn, resid, z, m, q = s.atoms()
        inames, iresn, ibonds, ix, ibox = independent(path)
        assert names == inames and resn == iresn, rel
        assert [tuple(b) for b in s.bonds().tolist()] == [tuple(int(v) for v in b) for b in ibonds], rel
        assert np.array_equal(s.positions(), np.asarray(ix, np.float32)) and np.array_equal(s.box9(), np.asarray(ibox, np.float32)), rel
        if rel in keep:
            fn = os.path.basename(rel)
            shutil.copyfile(path, os.path.join(out, fn))
            exp[fn] = dict(tpx=s.tpx_version, n_atoms=s.n_atoms, n_bonds=s.n_bonds, names_head=names[:24], resn_head=resn[:24],
                           resid_head=[int(v) for v in resid[:24]], names_crc=zlib.crc32(" ".join(names).encode()),
                           resn_crc=zlib.crc32(" ".join(resn).encode()), bonds_crc=zlib.crc32(s.bonds().astype("<i4").tobytes()),
                           xyz_head=[float(v) for v in s.positions()[:4].reshape(-1)], box9=[float(v) for v in s.box9()])
    for tpr, gro, bnd in (("cg.tpr", "cg.gro", "cg.bnd"), ("pcpepg.tpr", "pcpepg.gro", "pcpepg.bnd")):
        s = System.from_tpr(os.path.join(FILES, tpr))
        st = fixtures.read_gro(os.path.join(FILES, gro))
        fixtures.read_bnd(os.path.join(FILES, bnd), st)
        names, resn, *_ = s.atoms()
        assert names == st.name and resn == st.resname and set(st.bonds) <= set(map(tuple, s.bonds().tolist())), tpr
    json.dump(exp, open(os.path.join(out, "expected.json"), "w"), indent=1)

    # the asymmetric CG membrane: trajectory of the membrane beads + the reference's expected values
    s = System.from_tpr(os.path.join(FILES, "asymmetric/cg_asym.tpr"))
    names, resn, *_ = s.atoms()
    keep_atoms = np.array([i for i in range(s.n_atoms) if resn[i] in LIPIDS], dtype=np.int64)   # @membrane = the Master group
    tr = fixtures.read_xtc(os.path.join(FILES, "asymmetric/cg_asym.xtc"))
    assert tr.xyz.shape[1] == s.n_atoms
    prec = 100.0
    q = np.round(tr.xyz[:, keep_atoms, :].astype(np.float64) * prec).astype(np.int32)
    assert np.array_equal(q.astype(np.float32) * np.float32(1.0 / prec), tr.xyz[:, keep_atoms, :]), "XTC coordinates are not k/100"
    tul = ("total", "upper", "lower")
    cases = {}
    for case, yf, extra in (("leaflets", "asymmetric/cg_order_asymmetric.yaml", {}),
                            ("errors", "asymmetric/cg_order_asymmetric_errors.yaml", dict(n_blocks=5))):
        doc = yaml.safe_load(open(os.path.join(FILES, yf)))
        cases[case] = dict(expected=flatten_yaml(doc, tul), keys=list(tul), source=yf, **extra)
    np.savez_compressed(os.path.join(HERE, "cg_asym.npz"), keep=keep_atoms, box=tr.box.astype(np.float32), precision=np.float32(prec),
                        cases=json.dumps(cases), **pack_lattice(q))
    print("tpr fixtures:", sorted(os.listdir(out)), "cg_asym.npz", os.path.getsize(os.path.join(HERE, "cg_asym.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tpr":
        tpr_golden()
        sys.exit(0)
    xtc_fixtures()
    normals_planar()
    single_frame("cg_single_frame", "cg.gro", "cg.bnd", "cg.tpr", abi.KIND_CG, "cgorder.rs", 1.0, "PO4")
    single_frame("aa_single_frame", "pcpepg.gro", "pcpepg.bnd", "pcpepg.tpr", abi.KIND_AA, "aaorder.rs", -1.0, "P")
    ua_golden()
    aa_traj_golden()
    ua_nopbc_golden()
    full_traj_golden("aa_full", "pcpepg", "pcpepg.gro", "pcpepg.bnd", abi.KIND_AA, "P", ("C218", "C316"), 51)
    full_traj_golden("cg_full", "cg", "cg.gro", "cg.bnd", abi.KIND_CG, "PO4", ("C4A", "C4B"), 101)
