"""Host-side mirror of the reference's structure / topology step (``src/analysis/structure.rs:27-165``) and of its molecule
classification (``src/analysis/topology/classify.rs:45-315, 318-580``).

All the work happens in C++ behind the C ABI (``gorder_system_*``, ``gorder_classify_*`` of ``include/gorder_b200.h``,
``csrc/gorder_topology.inl``): this module only marshals strings and index arrays.  Atom groups are index arrays -- the
selection language of the reference (GSL) is not part of this repository.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import abi
from ._lib import lib


def _fail(rc: int):
    raise abi.GorderError(rc, lib().gorder_topology_last_error().decode(errors="replace"))


def _idx(a) -> Optional[np.ndarray]:
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32).reshape(-1)


def _p(a: Optional[np.ndarray]):
    return C.c_void_p(None) if a is None else C.c_void_p(a.ctypes.data)


class System:
    """Atoms, bonds, box and coordinates of a system (the part of groan_rs' ``System`` gorder uses)."""

    def __init__(self, handle):
        self._h = handle

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_tpr(cls, path: str) -> "System":
        """``System::from_file`` on a TPR (structure.rs:31): GROMACS 5.1 - 2022 (tpx 103 - 127)."""
        h = C.c_void_p()
        rc = lib().gorder_system_from_tpr(path.encode(), C.byref(h))
        if rc != abi.OK:
            _fail(rc)
        return cls(h)

    @classmethod
    def from_file(cls, structure: str, bonds_file: Optional[str] = None) -> "System":
        """``read_structure_and_topology`` (structure.rs:27-88): TPR, PDB (CONECT) or GRO, bonds file optional."""
        h = C.c_void_p()
        rc = lib().gorder_system_from_file(structure.encode(), bonds_file.encode() if bonds_file else None, C.byref(h))
        if rc != abi.OK:
            _fail(rc)
        return cls(h)

    @classmethod
    def from_arrays(cls, atom_names: Sequence[str], res_names: Sequence[str], res_ids=None, xyz=None, box9=None) -> "System":
        n = len(atom_names)
        an = (C.c_char_p * n)(*[s.encode() for s in atom_names])
        rn = (C.c_char_p * n)(*[s.encode() for s in res_names])
        ri = _idx(res_ids)
        x = None if xyz is None else np.ascontiguousarray(xyz, dtype=np.float32).reshape(n, 3)
        b = None if box9 is None else np.ascontiguousarray(box9, dtype=np.float32).reshape(9)
        h = C.c_void_p()
        rc = lib().gorder_system_from_arrays(n, an, rn, _p(ri), _p(x), _p(b), C.byref(h))
        if rc != abi.OK:
            _fail(rc)
        return cls(h)

    def close(self):
        if self._h:
            lib().gorder_system_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- contents -------------------------------------------------------------------------------
    @property
    def n_atoms(self) -> int:
        return int(lib().gorder_system_n_atoms(self._h))

    @property
    def n_bonds(self) -> int:
        return int(lib().gorder_system_n_bonds(self._h))

    @property
    def tpx_version(self) -> int:
        return int(lib().gorder_system_tpx_version(self._h))

    def atoms(self):
        """(names, residue names, residue numbers, atomic numbers, masses, charges)."""
        n = self.n_atoms
        an = np.zeros((n, 8), np.uint8)
        rn = np.zeros((n, 8), np.uint8)
        ri = np.zeros(n, np.int32)
        z = np.zeros(n, np.int32)
        m = np.zeros(n, np.float32)
        q = np.zeros(n, np.float32)
        rc = lib().gorder_system_atoms(self._h, _p(an), _p(rn), _p(ri), _p(z), _p(m), _p(q))
        if rc != abi.OK:
            _fail(rc)
        dec = lambda a: [bytes(r).split(b"\0", 1)[0].decode() for r in a]
        return dec(an), dec(rn), ri, z, m, q

    def bonds(self) -> np.ndarray:
        out = np.zeros((self.n_bonds, 2), np.int32)
        rc = lib().gorder_system_bonds(self._h, _p(out))
        if rc != abi.OK:
            _fail(rc)
        return out

    def positions(self) -> Optional[np.ndarray]:
        out = np.zeros((self.n_atoms, 3), np.float32)
        has = C.c_int32(0)
        lib().gorder_system_positions(self._h, _p(out), C.byref(has))
        return out if has.value else None

    def box9(self) -> Optional[np.ndarray]:
        out = np.zeros(9, np.float32)
        has = C.c_int32(0)
        lib().gorder_system_box(self._h, _p(out), C.byref(has))
        return out if has.value else None

    # -- bonds ----------------------------------------------------------------------------------
    def set_bonds(self, pairs) -> None:
        p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        rc = lib().gorder_system_set_bonds(self._h, _p(p), p.shape[0])
        if rc != abi.OK:
            _fail(rc)

    def read_bonds(self, bonds_file: str) -> None:
        """``read_bonds`` (structure.rs:91-165)."""
        rc = lib().gorder_system_read_bonds(self._h, bonds_file.encode())
        if rc != abi.OK:
            _fail(rc)

    # -- classification -------------------------------------------------------------------------
    def _collect(self, h, kind: int) -> List[abi.MolType]:
        L = lib()
        try:
            n = L.gorder_classification_n_types(h)
            arr = C.cast(L.gorder_classification_moltypes(h), C.POINTER(abi.CGorderMolType))
            out = []
            for t in range(n):
                m = arr[t]
                take = lambda p, k: [int(p[i]) for i in range(k)]
                n_items = m.n_ua_atoms if kind == abi.KIND_UA else m.n_bond_types
                out.append(abi.MolType(
                    name=L.gorder_classification_type_name(h, t).decode(),
                    mol_base=take(m.mol_base, m.n_molecules),
                    bond_rel=[(int(m.bond_rel[2 * i]), int(m.bond_rel[2 * i + 1])) for i in range(m.n_bond_types)],
                    ua_kind=take(m.ua_kind, m.n_ua_atoms),
                    ua_rel=[tuple(int(m.ua_rel[4 * i + k]) for k in range(4)) for i in range(m.n_ua_atoms)],
                    head_rel=int(m.head_rel), methyl_rel=take(m.methyl_rel, m.n_methyls), normal_head_rel=int(m.normal_head_rel),
                    bond_names=[L.gorder_classification_item_name(h, t, i).decode() for i in range(n_items)]))
            self.last_warning = L.gorder_classification_warning(h).decode()
            return out
        finally:
            L.gorder_classification_free(h)

    def classify_bonds(self, kind: int, group1, group2, heads=None, methyls=None, normal_heads=None) -> List[abi.MolType]:
        """``MoleculesClassifier::classify`` for AA (heavy atoms x hydrogens) / CG (beads x beads), classify.rs:53-76."""
        g1, g2, hd, me, nh = _idx(group1), _idx(group2), _idx(heads), _idx(methyls), _idx(normal_heads)
        h = C.c_void_p()
        rc = lib().gorder_classify_bonds(self._h, _p(g1), g1.size, _p(g2), g2.size, _p(hd), 0 if hd is None else hd.size,
                                         _p(me), 0 if me is None else me.size, _p(nh), 0 if nh is None else nh.size, C.byref(h))
        if rc != abi.OK:
            _fail(rc)
        return self._collect(h, kind)

    def classify_ua(self, saturated, unsaturated=(), ignore=(), heads=None, methyls=None, normal_heads=None) -> List[abi.MolType]:
        """``MoleculesClassifier::classify`` for UA (classify.rs:77-90; uaorder.rs:580-665)."""
        sa, un, ig, hd, me, nh = _idx(saturated), _idx(unsaturated), _idx(ignore), _idx(heads), _idx(methyls), _idx(normal_heads)
        h = C.c_void_p()
        rc = lib().gorder_classify_ua(self._h, _p(sa), sa.size, _p(un), un.size, _p(ig), ig.size, _p(hd), 0 if hd is None else hd.size,
                                      _p(me), 0 if me is None else me.size, _p(nh), 0 if nh is None else nh.size, C.byref(h))
        if rc != abi.OK:
            _fail(rc)
        return self._collect(h, abi.KIND_UA)


def read_ndx(path: str, n_atoms: int = -1):
    """GROMACS index file -> {group name: atom indices from 0} (groan_rs ``Groups::from_ndx``)."""
    L = lib()
    h = C.c_void_p()
    rc = L.gorder_ndx_open(path.encode(), n_atoms, C.byref(h))
    if rc != abi.OK:
        _fail(rc)
    try:
        out = {}
        for g in range(L.gorder_ndx_n_groups(h)):
            n = L.gorder_ndx_group_size(h, g)
            p = L.gorder_ndx_group_atoms(h, g)
            out[L.gorder_ndx_group_name(h, g).decode()] = np.array([p[i] for i in range(n)], dtype=np.int32)
        return out
    finally:
        L.gorder_ndx_close(h)


def leaflets_from_ndx(ndx_files: Sequence[str], heads, upper: str = "Upper", lower: str = "Lower", n_atoms: int = -1) -> np.ndarray:
    """``LeafletClassification::from_ndx`` (leaflets.rs:1030-1215) for the molecules whose head atoms are ``heads`` (absolute
    indices, molecule order): uint8 table [len(ndx_files)][len(heads)] for ``MolType.manual_leaflets``."""
    hd = np.ascontiguousarray(heads, dtype=np.int32).reshape(-1)
    files = (C.c_char_p * len(ndx_files))(*[f.encode() for f in ndx_files])
    table = np.zeros((len(ndx_files), hd.size), np.uint8)
    rc = lib().gorder_leaflets_from_ndx(files, len(ndx_files), n_atoms, upper.encode(), lower.encode(), _p(hd), hd.size, _p(table))
    if rc != abi.OK:
        _fail(rc)
    return table
