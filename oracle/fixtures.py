"""Readers and a molecule classifier for the reference's own test fixtures.

TEST INFRASTRUCTURE ONLY (see ``oracle/gorder_oracle.c``).  Used (a) in this container to pin the
oracle against the reference's fixtures under ``/root/reference/tests/files`` and to generate the
small golden vectors committed under ``tests/golden`` (``tests/golden/make_golden.py``), and (b) by
the tests to load those golden vectors.  Nothing here runs on the product path.

Restated reference logic:
  * molecule classification: ``src/analysis/topology/classify.rs:140-315`` (iterate the order
    group in index order, molecule = connected component, ``min_index`` = lowest atom index,
    molecule types keyed by topology, in order of first appearance),
  * bond-based order bonds: ``classify.rs:355-420``; sorted by relative indices ``bond.rs:77-81``,
  * united-atom typing: ``src/analysis/uaorder.rs:580-665`` (``get_atom_type``),
  * reference head / methyls: ``src/analysis/common.rs:345-376``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from gorder_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------
# structure / topology files
# ---------------------------------------------------------------------------------------------
@dataclass
class Structure:
    resid: np.ndarray
    resname: List[str]
    name: List[str]
    xyz: np.ndarray            # [n][3] f32, nm
    box: np.ndarray            # [3] f32 (zeros if unknown)
    bonds: List[Tuple[int, int]] = field(default_factory=list)   # 0-based, i < j

    @property
    def n_atoms(self) -> int:
        return len(self.name)

    def neighbours(self) -> List[List[int]]:
        nb: List[List[int]] = [[] for _ in range(self.n_atoms)]
        for i, j in self.bonds:
            nb[i].append(j)
            nb[j].append(i)
        return [sorted(set(x)) for x in nb]

    def select(self, pred: Callable[[str, str], bool]) -> np.ndarray:
        """Indices of atoms for which ``pred(resname, atomname)`` holds (stand-in for GSL queries)."""
        return np.array([i for i in range(self.n_atoms) if pred(self.resname[i], self.name[i])], dtype=np.int64)


def read_gro(path: str) -> Structure:
    with open(path) as f:
        lines = f.read().split("\n")
    n = int(lines[1])
    resid = np.zeros(n, np.int64)
    resname, name = [], []
    xyz = np.zeros((n, 3), np.float32)
    for i in range(n):
        ln = lines[2 + i]
        resid[i] = int(ln[0:5])
        resname.append(ln[5:10].strip())
        name.append(ln[10:15].strip())
        xyz[i] = [float(ln[20:28]), float(ln[28:36]), float(ln[36:44])]
    b = [float(x) for x in lines[2 + n].split()]
    return Structure(resid, resname, name, xyz, np.array(b[:3], np.float32))


def read_bnd(path: str, st: Structure) -> None:
    """``.bnd`` bonds file: ``i j k ...`` (1-based) = atom i bonded to j, k, ... (structure.rs:140-200)."""
    bonds = set()
    with open(path) as f:
        for ln in f:
            ln = ln.split("#")[0].split()
            if len(ln) < 2:
                continue
            i = int(ln[0]) - 1
            for t in ln[1:]:
                j = int(t) - 1
                if i != j:
                    bonds.add((min(i, j), max(i, j)))
    st.bonds = sorted(bonds)


def read_pdb(path: str) -> Structure:
    """PDB with CONECT records (Angstrom -> nm)."""
    resid, resname, name, xyz, bonds = [], [], [], [], set()
    serial2idx: Dict[int, int] = {}
    box = np.zeros(3, np.float32)
    done = False
    with open(path) as f:
        for ln in f:
            rec = ln[0:6]
            if rec in ("ATOM  ", "HETATM") and not done:
                serial2idx[int(ln[6:11])] = len(name)
                name.append(ln[12:16].strip())
                resname.append(ln[17:21].strip())
                resid.append(int(ln[22:26]))
                xyz.append([float(ln[30:38]) / 10.0, float(ln[38:46]) / 10.0, float(ln[46:54]) / 10.0])
            elif rec == "CRYST1":
                box = np.array([float(ln[6:15]) / 10, float(ln[15:24]) / 10, float(ln[24:33]) / 10], np.float32)
            elif rec == "CONECT":
                nums = [int(ln[k:k + 5]) for k in range(6, len(ln.rstrip()), 5) if ln[k:k + 5].strip()]
                i = serial2idx[nums[0]]
                for s in nums[1:]:
                    j = serial2idx[s]
                    if i != j:
                        bonds.add((min(i, j), max(i, j)))
            elif rec.startswith("ENDMDL"):
                done = True   # first model only; CONECT records follow
    return Structure(np.array(resid), resname, name, np.array(xyz, np.float32), box, sorted(bonds))


def tpr_coordinates(path: str, approx_xyz: np.ndarray, hint: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, int]:
    """Dig the f32 big-endian coordinate block (and the box at byte 100) out of a GROMACS TPR by
    matching it against approximate coordinates (GRO/PDB of the same system).  SURVEY.md §8c."""
    raw = np.fromfile(path, dtype=np.uint8)
    n = approx_xyz.shape[0]
    flat = approx_xyz.reshape(-1).astype(np.float32)
    probe = flat[:12]

    def block(off):
        return np.frombuffer(raw[off:off + 4 * 3 * n].tobytes(), dtype=">f4").astype(np.float32)

    cands = [hint] if hint is not None else []
    for align in range(4):
        v = np.frombuffer(raw[align:align + (len(raw) - align) // 4 * 4].tobytes(), dtype=">f4")
        with np.errstate(invalid="ignore"):
            hits = np.nonzero(np.abs(v[:-12] - probe[0]) < 6e-4)[0]
        for h in hits:
            if np.all(np.abs(v[h:h + 12] - probe) < 6e-4):
                cands.append(align + 4 * int(h))
    for off in cands:
        if off is None or off + 12 * n > len(raw):
            continue
        b = block(off)
        if np.all(np.isfinite(b)) and np.max(np.abs(b - flat)) < 1e-3:
            box9 = np.frombuffer(raw[100:136].tobytes(), dtype=">f4").astype(np.float32)
            return b.reshape(n, 3), np.array([box9[0], box9[4], box9[8]], np.float32), off
    raise RuntimeError(f"coordinate block not found in {path}")


# ---------------------------------------------------------------------------------------------
# XTC
# ---------------------------------------------------------------------------------------------
_XTC = None


def _xtc_lib():
    global _XTC
    if _XTC is None:
        from . import oracle as _o
        _o.build()
        _XTC = C.CDLL(os.path.join(_HERE, "libxtc.so"))
        _XTC.xtc_scan.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _XTC.xtc_read.argtypes = [C.c_char_p, C.c_int, C.c_int] + [C.c_void_p] * 5
    return _XTC


@dataclass
class Trajectory:
    xyz: np.ndarray    # [frames][atoms][3]
    box: np.ndarray    # [frames][3] (diagonal)
    box9: np.ndarray   # [frames][9]
    time: np.ndarray
    step: np.ndarray
    precision: float


def read_xtc(path: str) -> Trajectory:
    lib = _xtc_lib()
    na, nf = C.c_int(0), C.c_int(0)
    rc = lib.xtc_scan(path.encode(), C.byref(na), C.byref(nf))
    if rc:
        raise RuntimeError(f"xtc_scan({path}) failed: {rc}")
    xyz = np.zeros((nf.value, na.value, 3), np.float32)
    box9 = np.zeros((nf.value, 9), np.float32)
    time = np.zeros(nf.value, np.float32)
    step = np.zeros(nf.value, np.int32)
    prec = C.c_float(0)
    rc = lib.xtc_read(path.encode(), na.value, nf.value, xyz.ctypes.data, box9.ctypes.data, time.ctypes.data,
                      step.ctypes.data, C.byref(prec))
    if rc:
        raise RuntimeError(f"xtc_read({path}) failed: {rc}")
    return Trajectory(xyz, box9[:, [0, 4, 8]].copy(), box9, time, step, float(prec.value))


# ---------------------------------------------------------------------------------------------
# molecule classification
# ---------------------------------------------------------------------------------------------
def _component(nb: List[List[int]], start: int) -> List[int]:
    seen = {start}
    stack = [start]
    while stack:
        a = stack.pop()
        for b in nb[a]:
            if b not in seen:
                seen.add(b)
                stack.append(b)
    return sorted(seen)


@dataclass
class ClassifiedType:
    name: str
    atoms_rel: List[int]                  # relative indices of all atoms of the molecule
    names_rel: Dict[int, Tuple[str, str]]  # rel -> (resname, atomname)
    mol_base: List[int]
    bonds_rel: List[Tuple[int, int]]      # all bonds (relative)


def classify(st: Structure, order_atoms: Sequence[int]) -> List[ClassifiedType]:
    """classify.rs:140-242: molecule types in order of first appearance; molecules in index order."""
    nb = st.neighbours()
    visited = set()
    types: List[ClassifiedType] = []
    keys: List[tuple] = []
    comp_bonds: Dict[int, List[Tuple[int, int]]] = {}
    bonds_by_atom: Dict[int, List[Tuple[int, int]]] = {}
    for i, j in st.bonds:
        bonds_by_atom.setdefault(i, []).append((i, j))
    for a in sorted(int(x) for x in order_atoms):
        if a in visited:
            continue
        comp = _component(nb, a)
        visited.update(comp)
        if len(comp) < 2:
            continue  # an atom without bonds has no molecule bonds (min_index undefined)
        mn = comp[0]
        rel_bonds = sorted((i - mn, j - mn) for c in comp for (i, j) in bonds_by_atom.get(c, []))
        key = (tuple((c - mn, st.resname[c], st.name[c]) for c in comp), tuple(rel_bonds))
        if key in keys:
            types[keys.index(key)].mol_base.append(mn)
        else:
            residues: List[str] = []
            for c in comp:
                if st.resname[c] not in residues:
                    residues.append(st.resname[c])
            keys.append(key)
            types.append(ClassifiedType("-".join(residues), [c - mn for c in comp],
                                        {c - mn: (st.resname[c], st.name[c]) for c in comp}, [mn], rel_bonds))
    # solve_name_conflicts (classify.rs:262-294)
    counts: Dict[str, int] = {}
    for t in types:
        counts[t.name] = counts.get(t.name, 0) + 1
    counts = {k: v for k, v in counts.items() if v > 1}
    for t in reversed(types):
        if t.name in counts:
            c = counts[t.name]
            counts[t.name] -= 1
            t.name = f"{t.name}{c}"
    return types


def _single_rel(t: ClassifiedType, base_sel: set, what: str) -> int:
    """get_reference_head (common.rs:345-360): exactly one atom of the molecule in the group."""
    hits = [r for r in t.atoms_rel if (t.mol_base[0] + r) in base_sel]
    if len(hits) != 1:
        raise ValueError(f"molecule type {t.name}: expected exactly one {what}, found {len(hits)}")
    return hits[0]


def build_bond_setup(st: Structure, kind: int, group1: Sequence[int], group2: Sequence[int], *,
                     heads: Optional[Sequence[int]] = None, methyls: Optional[Sequence[int]] = None,
                     normal_heads: Optional[Sequence[int]] = None, **kw) -> abi.EngineSetup:
    """AA (heavy atoms x hydrogens) or CG (beads x beads) setup from a structure with bonds."""
    g1, g2 = set(int(x) for x in group1), set(int(x) for x in group2)
    types = classify(st, sorted(g1))
    hs = set(int(x) for x in heads) if heads is not None else None
    ms = set(int(x) for x in methyls) if methyls is not None else None
    nh = set(int(x) for x in normal_heads) if normal_heads is not None else None
    mts = []
    for t in types:
        base = t.mol_base[0]
        ob = sorted((i, j) for (i, j) in t.bonds_rel
                    if ((base + i) in g1 and (base + j) in g2) or ((base + i) in g2 and (base + j) in g1))
        if not ob:
            continue
        names = [f"{t.names_rel[i][0]} {t.names_rel[i][1]} ({i}) - {t.names_rel[j][0]} {t.names_rel[j][1]} ({j})" for i, j in ob]
        mts.append(abi.MolType(
            name=t.name, mol_base=t.mol_base, bond_rel=ob, bond_names=names,
            head_rel=_single_rel(t, hs, "head") if hs is not None else -1,
            methyl_rel=[r for r in t.atoms_rel if (base + r) in ms] if ms is not None else (),
            normal_head_rel=_single_rel(t, nh, "normal head") if nh is not None else -1))
    if normal_heads is not None:
        kw.setdefault("normal_heads", sorted(nh))
    return abi.EngineSetup(kind=kind, n_atoms=st.n_atoms, moltypes=mts, **kw)


def build_ua_setup(st: Structure, saturated: Sequence[int], unsaturated: Sequence[int] = (),
                   ignore: Sequence[int] = (), *, heads: Optional[Sequence[int]] = None,
                   methyls: Optional[Sequence[int]] = None, normal_heads: Optional[Sequence[int]] = None,
                   **kw) -> abi.EngineSetup:
    """UA setup: carbon typing as ``UAOrderAtomType::get_atom_type`` (uaorder.rs:580-665)."""
    sat, unsat, ign = (set(int(x) for x in s) for s in (saturated, unsaturated, ignore))
    nb = st.neighbours()
    types = classify(st, sorted(sat | unsat))
    hs = set(int(x) for x in heads) if heads is not None else None
    ms = set(int(x) for x in methyls) if methyls is not None else None
    nh = set(int(x) for x in normal_heads) if normal_heads is not None else None
    mts = []
    for t in types:
        base = t.mol_base[0]
        kinds, rels, names = [], [], []
        for r in sorted(x for x in t.atoms_rel if (base + x) in sat or (base + x) in unsat):
            a = base + r
            bonded = [b for b in nb[a] if b not in ign]
            missing = max(0, 4 - len(bonded))
            is_sat = a in sat
            entry = None
            if missing == 0 or (not is_sat and missing == 1):
                entry = None
            elif is_sat and missing == 1:
                entry = (abi.UA_CH1_SAT, (a, bonded[0], bonded[1], bonded[2]))
            elif is_sat and missing == 2:
                entry = (abi.UA_CH2, (a, bonded[0], bonded[1], -1))
            elif is_sat and missing == 3:
                h1 = bonded[0]
                h2 = next((x for x in nb[h1] if x != a), None)
                entry = (abi.UA_CH3, (a, h1, h2, -1)) if h2 is not None else None
            elif (not is_sat) and missing == 2:
                entry = (abi.UA_CH1_UNSAT, (a, bonded[0], bonded[1], -1))
            if entry is None:
                continue
            kinds.append(entry[0])
            rels.append(tuple(x - base if x >= 0 else -1 for x in entry[1]))
            names.append(f"{t.names_rel[r][0]} {t.names_rel[r][1]} ({r})")
        if not kinds:
            continue
        mts.append(abi.MolType(
            name=t.name, mol_base=t.mol_base, ua_kind=kinds, ua_rel=rels, bond_names=names,
            head_rel=_single_rel(t, hs, "head") if hs is not None else -1,
            methyl_rel=[r for r in t.atoms_rel if (base + r) in ms] if ms is not None else (),
            normal_head_rel=_single_rel(t, nh, "normal head") if nh is not None else -1))
    if normal_heads is not None:
        kw.setdefault("normal_heads", sorted(nh))
    return abi.EngineSetup(kind=abi.KIND_UA, n_atoms=st.n_atoms, moltypes=mts, **kw)


def compact(st: Structure, keep: Sequence[int]) -> Tuple[Structure, np.ndarray]:
    """Keep only ``keep`` atoms (the reference's Master group, common.rs:92-103) and renumber.
    Returns the compacted structure and the kept original indices."""
    keep = np.array(sorted(set(int(x) for x in keep)), dtype=np.int64)
    new = -np.ones(st.n_atoms, np.int64)
    new[keep] = np.arange(keep.size)
    bonds = [(int(new[i]), int(new[j])) for i, j in st.bonds if new[i] >= 0 and new[j] >= 0]
    out = Structure(st.resid[keep], [st.resname[i] for i in keep], [st.name[i] for i in keep], st.xyz[keep].copy(), st.box.copy(), bonds)
    return out, keep
