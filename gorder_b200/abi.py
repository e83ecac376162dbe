"""ctypes mirror of ``include/gorder_b200.h`` (the C ABI of the engine).

Only declarations live here: struct layouts, enum values and a builder that turns the Python
description of a classified system (:class:`MolType`, :class:`EngineSetup`) into a
``GorderSetup`` whose arrays stay alive as long as the Python object does.

Reference counterparts: the fields mirror what ``SystemTopology::new`` receives
(``src/analysis/topology/mod.rs:70-118``) and what the per-molecule-type structures hold
(``topology/molecule.rs:147-169``, ``topology/bond.rs:221-247``, ``uaorder.rs:234-239``,
``leaflets.rs:571-811``, ``normal.rs:131-141``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

ABI_VERSION = 2

# enums -----------------------------------------------------------------------------------
KIND_AA, KIND_CG, KIND_UA = 0, 1, 2
AXIS_X, AXIS_Y, AXIS_Z = 0, 1, 2
NORMAL_STATIC, NORMAL_DYNAMIC, NORMAL_MANUAL = 0, 1, 2
LEAFLET_NONE, LEAFLET_GLOBAL, LEAFLET_LOCAL, LEAFLET_INDIVIDUAL, LEAFLET_MANUAL = 0, 1, 2, 3, 4
LEAFLET_SPHERICAL = 5   # spherical clustering (spherical_clustering.rs:36-275; csrc/gorder_spherical.cuh)
FREQ_EVERY, FREQ_ONCE = 0, 1
GEOM_NONE, GEOM_CUBOID, GEOM_CYLINDER, GEOM_SPHERE = 0, 1, 2, 3
GEOMREF_POINT, GEOMREF_SELECTION, GEOMREF_BOX_CENTER = 0, 1, 2
PLANE_XY, PLANE_XZ, PLANE_YZ = 0, 1, 2
UA_CH3, UA_CH2, UA_CH1_UNSAT, UA_CH1_SAT = 0, 1, 2, 3
LOWER, UPPER = 0, 1
TOTAL, ACC_UPPER, ACC_LOWER = 0, 1, 2

OK = 0
ERR_UNDEFINED_BOX = 1
ERR_NOT_ORTHOGONAL_BOX = 2
ERR_ZERO_BOX = 3
ERR_UNDEFINED_POSITION = 4
ERR_INVALID_GLOBAL_CENTER = 5
ERR_INVALID_LOCAL_CENTER = 6
ERR_MANUAL_LEAFLET_FRAME = 7
ERR_DYNAMIC_NORMAL_POINTS = 8
ERR_DYNAMIC_NORMAL_SVD = 9
ERR_MANUAL_NORMAL_FRAME = 10
ERR_LEAFLET_FRAME_UNAVAILABLE = 11
ERR_ORDER_OVERFLOW = 12
ERR_INVALID_ARGUMENT = 20
ERR_ORDERMAP_BIN_TOO_LARGE = 21
ERR_ORDERMAP_NO_BOX = 22
ERR_NO_DEVICE = 30
ERR_CUDA = 31
ERR_OUT_OF_MEMORY = 32
ERR_NCCL = 33
ERR_IO = 40
ERR_TPR_FORMAT = 41
ERR_BONDS_PARSE = 42
ERR_BONDS_ATOM_NOT_FOUND = 43
ERR_BONDS_SELF = 44
ERR_TOPOLOGY_NO_HEAD = 45
ERR_TOPOLOGY_MULTIPLE_HEADS = 46
ERR_TOPOLOGY_NO_METHYL = 47
ERR_TOPOLOGY_INCONSISTENT_METHYLS = 48
ERR_TOPOLOGY_NO_UA_CARBONS = 49
ERR_NO_TOPOLOGY = 50
ERR_PDB_TOPOLOGY = 51
ERR_STRUCTURE_FORMAT = 52
ERR_NDX_PARSE = 53
ERR_NDX_INVALID_NAME = 54
ERR_NDX_DUPLICATE_NAME = 55
ERR_NDX_GROUP_NOT_FOUND = 56
ERR_NDX_ASSIGNMENT_NOT_FOUND = 57

ERROR_NAMES = {
    ERR_UNDEFINED_BOX: "AnalysisError::UndefinedBox",
    ERR_NOT_ORTHOGONAL_BOX: "AnalysisError::NotOrthogonalBox",
    ERR_ZERO_BOX: "AnalysisError::ZeroBox",
    ERR_UNDEFINED_POSITION: "AnalysisError::UndefinedPosition",
    ERR_INVALID_GLOBAL_CENTER: "AnalysisError::InvalidGlobalMembraneCenter",
    ERR_INVALID_LOCAL_CENTER: "AnalysisError::InvalidLocalMembraneCenter",
    ERR_MANUAL_LEAFLET_FRAME: "ManualLeafletClassificationError::FrameNotFound",
    ERR_DYNAMIC_NORMAL_POINTS: "DynamicNormalError::NotEnoughPoints",
    ERR_DYNAMIC_NORMAL_SVD: "DynamicNormalError::SVDFailed",
    ERR_MANUAL_NORMAL_FRAME: "ManualNormalError::FrameNotFound",
    ERR_LEAFLET_FRAME_UNAVAILABLE: "leaflet assignment frame not available on this shard",
    ERR_ORDER_OVERFLOW: "OrderValue overflowed",
    ERR_INVALID_ARGUMENT: "invalid argument",
    ERR_ORDERMAP_BIN_TOO_LARGE: "OrderMapConfigError::BinTooLarge",
    ERR_ORDERMAP_NO_BOX: "OrderMapConfigError::InvalidBoxAuto",
    ERR_NO_DEVICE: "no CUDA device (there is no CPU fallback)",
    ERR_CUDA: "CUDA error",
    ERR_OUT_OF_MEMORY: "out of device memory",
    ERR_NCCL: "NCCL unavailable or a collective failed",
    ERR_IO: "file missing or unreadable",
    ERR_TPR_FORMAT: "not a supported TPR file",
    ERR_BONDS_PARSE: "BondsError::CouldNotParse",
    ERR_BONDS_ATOM_NOT_FOUND: "BondsError::AtomNotFound",
    ERR_BONDS_SELF: "BondsError::SelfBonding",
    ERR_TOPOLOGY_NO_HEAD: "TopologyError::NoHead",
    ERR_TOPOLOGY_MULTIPLE_HEADS: "TopologyError::MultipleHeads",
    ERR_TOPOLOGY_NO_METHYL: "TopologyError::NoMethyl",
    ERR_TOPOLOGY_INCONSISTENT_METHYLS: "TopologyError::InconsistentNumberOfMethyls",
    ERR_TOPOLOGY_NO_UA_CARBONS: "TopologyError::NoUACarbons",
    ERR_NO_TOPOLOGY: "ConfigError::NoTopology",
    ERR_PDB_TOPOLOGY: "ConfigError::InvalidPdbTopology",
    ERR_STRUCTURE_FORMAT: "ConfigError::InvalidStructureFormat",
    ERR_NDX_PARSE: "NdxLeafletClassificationError::CouldNotParse",
    ERR_NDX_INVALID_NAME: "NdxLeafletClassificationError::InvalidName",
    ERR_NDX_DUPLICATE_NAME: "NdxLeafletClassificationError::DuplicateName",
    ERR_NDX_GROUP_NOT_FOUND: "NdxLeafletClassificationError::GroupNotFound",
    ERR_NDX_ASSIGNMENT_NOT_FOUND: "NdxLeafletClassificationError::AssignmentNotFound",
}

_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)


class CGorderMolType(C.Structure):
    _fields_ = [
        ("n_molecules", C.c_int32),
        ("mol_base", _i32p),
        ("n_bond_types", C.c_int32),
        ("bond_rel", _i32p),
        ("n_ua_atoms", C.c_int32),
        ("ua_kind", _i32p),
        ("ua_rel", _i32p),
        ("head_rel", C.c_int32),
        ("n_methyls", C.c_int32),
        ("methyl_rel", _i32p),
        ("normal_head_rel", C.c_int32),
        ("n_manual_leaflet_frames", C.c_int32),
        ("manual_leaflets", _u8p),
        ("n_manual_normal_frames", C.c_int32),
        ("manual_normals", _f32p),
    ]


class CGorderSetup(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("kind", C.c_int32),
        ("n_atoms", C.c_int32),
        ("handle_pbc", C.c_int32),
        ("step", C.c_int32),
        ("n_moltypes", C.c_int32),
        ("moltypes", C.POINTER(CGorderMolType)),
        ("normal_mode", C.c_int32),
        ("normal_axis", C.c_int32),
        ("dynamic_radius", C.c_float),
        ("n_normal_heads", C.c_int32),
        ("normal_heads", _i32p),
        ("collect_normals", C.c_int32),
        ("leaflet_mode", C.c_int32),
        ("leaflet_axis", C.c_int32),
        ("leaflet_freq_kind", C.c_int32),
        ("leaflet_freq", C.c_int32),
        ("leaflet_flip", C.c_int32),
        ("leaflet_radius", C.c_float),
        ("n_membrane", C.c_int32),
        ("membrane", _i32p),
        ("collect_leaflets", C.c_int32),
        ("geom_kind", C.c_int32),
        ("geom_invert", C.c_int32),
        ("geom_ref_kind", C.c_int32),
        ("geom_ref_point", C.c_float * 3),
        ("n_geom_ref", C.c_int32),
        ("geom_ref", _i32p),
        ("geom_dims", C.c_float * 6),
        ("geom_axis", C.c_int32),
        ("structure_box", C.c_float * 3),
        ("map_enabled", C.c_int32),
        ("map_plane", C.c_int32),
        ("map_span_x", C.c_float * 2),
        ("map_span_y", C.c_float * 2),
        ("map_bin", C.c_float * 2),
        ("timewise", C.c_int32),
        ("device", C.c_int32),
        ("max_batch_frames", C.c_int32),
    ]


class CGorderResults(C.Structure):
    _fields_ = [
        ("n_slots", C.c_int64),
        ("n_frames", C.c_int64),
        ("n_map_bins", C.c_int64),
        ("map_nx", C.c_int64),
        ("map_ny", C.c_int64),
        ("n_leaflet_frames", C.c_int64),
        ("n_molecules_total", C.c_int64),
        ("sum", _i64p),
        ("count", _u64p),
        ("tw_sum", _i64p),
        ("tw_count", _u64p),
        ("tw_frame_index", _i64p),
        ("map_sum", _i64p),
        ("map_count", _u64p),
        ("leaflets", _u8p),
        ("leaflet_frame_index", _i64p),
        ("normals", _f32p),
    ]


class CGorderRaw(C.Structure):
    """``GorderRaw`` (include/gorder_b200.h): views of the accumulators for ``gorder_results_*``."""

    _fields_ = [
        ("n_slots", C.c_int32),
        ("n_frames", C.c_int64),
        ("sum", C.c_void_p),
        ("count", C.c_void_p),
        ("tw_sum", C.c_void_p),
        ("tw_count", C.c_void_p),
    ]


# Python-side description ---------------------------------------------------------------------

def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def ua_hydrogens(kind: int) -> int:
    """Virtual hydrogens per carbon kind (uaorder.rs:253-272)."""
    return {UA_CH3: 3, UA_CH2: 2, UA_CH1_UNSAT: 1, UA_CH1_SAT: 1}[int(kind)]


@dataclass
class MolType:
    """One molecule type (reference: ``MoleculeType<O>``, topology/molecule.rs:147-169)."""

    name: str
    mol_base: Sequence[int]
    bond_rel: Sequence[Sequence[int]] = ()          # AA / CG
    ua_kind: Sequence[int] = ()                     # UA
    ua_rel: Sequence[Sequence[int]] = ()            # UA: (target, helper1, helper2, helper3|-1)
    head_rel: int = -1
    methyl_rel: Sequence[int] = ()
    normal_head_rel: int = -1
    manual_leaflets: Optional[np.ndarray] = None    # [frames][n_molecules] uint8
    manual_normals: Optional[np.ndarray] = None     # [frames][n_molecules][3] f32
    bond_names: Sequence[str] = ()                  # labels for presentation only

    @property
    def n_molecules(self) -> int:
        return len(self.mol_base)

    def n_orders(self, kind: int) -> int:
        if kind == KIND_UA:
            return sum(ua_hydrogens(k) for k in self.ua_kind)
        return len(self.bond_rel)


@dataclass
class EngineSetup:
    """Everything ``gorder_gpu_create`` needs (reference: ``SystemTopology::new`` arguments)."""

    kind: int
    n_atoms: int
    moltypes: List[MolType]
    handle_pbc: bool = True
    step: int = 1
    normal_mode: int = NORMAL_STATIC
    normal_axis: int = AXIS_Z
    dynamic_radius: float = 2.0
    normal_heads: Sequence[int] = ()
    collect_normals: bool = False
    leaflet_mode: int = LEAFLET_NONE
    leaflet_axis: int = AXIS_Z
    leaflet_freq_kind: int = FREQ_EVERY
    leaflet_freq: int = 1
    leaflet_flip: bool = False
    leaflet_radius: float = 2.5
    membrane: Sequence[int] = ()
    collect_leaflets: bool = False
    geom_kind: int = GEOM_NONE
    geom_invert: bool = False
    geom_ref_kind: int = GEOMREF_POINT
    geom_ref_point: Sequence[float] = (0.0, 0.0, 0.0)
    geom_ref: Sequence[int] = ()
    geom_dims: Sequence[float] = (0.0,) * 6
    geom_axis: int = AXIS_Z
    structure_box: Sequence[float] = (0.0, 0.0, 0.0)
    map_enabled: bool = False
    map_plane: int = PLANE_XY
    map_span_x: Sequence[float] = (0.0, 0.0)
    map_span_y: Sequence[float] = (0.0, 0.0)
    map_bin: Sequence[float] = (0.1, 0.1)
    timewise: bool = False
    device: int = 0
    max_batch_frames: int = 0
    _keep: list = field(default_factory=list, repr=False)

    # -- derived sizes ------------------------------------------------------------------
    @property
    def n_slots(self) -> int:
        return sum(m.n_orders(self.kind) for m in self.moltypes)

    @property
    def n_molecules_total(self) -> int:
        return sum(m.n_molecules for m in self.moltypes)

    def slot_ranges(self):
        """[(first_slot, n_slots)] per molecule type, in input order."""
        out, s = [], 0
        for m in self.moltypes:
            n = m.n_orders(self.kind)
            out.append((s, n))
            s += n
        return out

    # -- (de)serialisation: golden fixtures store the setup as JSON ---------------------------------
    _ARRAY_FIELDS = ("normal_heads", "membrane", "geom_ref")

    def to_dict(self) -> dict:
        d = {}
        for k, v in self.__dict__.items():
            if k in ("_keep", "moltypes"):
                continue
            if isinstance(v, np.ndarray):
                v = v.tolist()
            elif isinstance(v, (tuple, list)):
                v = [float(x) if isinstance(x, (float, np.floating)) else int(x) for x in v]
            elif isinstance(v, (np.integer,)):
                v = int(v)
            elif isinstance(v, (np.floating,)):
                v = float(v)
            d[k] = v
        d["moltypes"] = []
        for m in self.moltypes:
            md = dict(name=m.name, mol_base=[int(x) for x in m.mol_base], bond_rel=[[int(a), int(b)] for a, b in m.bond_rel],
                      ua_kind=[int(x) for x in m.ua_kind], ua_rel=[[int(x) for x in r] for r in m.ua_rel], head_rel=int(m.head_rel),
                      methyl_rel=[int(x) for x in m.methyl_rel], normal_head_rel=int(m.normal_head_rel),
                      bond_names=list(m.bond_names))
            if m.manual_leaflets is not None:
                md["manual_leaflets"] = np.asarray(m.manual_leaflets, np.uint8).reshape(-1, m.n_molecules).tolist()
            if m.manual_normals is not None:
                md["manual_normals"] = np.asarray(m.manual_normals, np.float32).reshape(-1, m.n_molecules, 3).tolist()
            d["moltypes"].append(md)
        return d

    @classmethod
    def from_dict(cls, d: dict) -> "EngineSetup":
        d = dict(d)
        mts = [MolType(**m) for m in d.pop("moltypes")]
        return cls(moltypes=mts, **d)

    def samples_per_frame(self) -> int:
        """Upper bound of S evaluations per frame (no geometry filter)."""
        return sum(m.n_orders(self.kind) * m.n_molecules for m in self.moltypes)

    # -- C struct -------------------------------------------------------------------------
    def to_c(self) -> CGorderSetup:
        keep = self._keep
        keep.clear()

        def ptr(arr, ctype):
            if arr is None or arr.size == 0:
                return C.cast(None, C.POINTER(ctype))
            keep.append(arr)
            return arr.ctypes.data_as(C.POINTER(ctype))

        mts = (CGorderMolType * max(1, len(self.moltypes)))()
        for i, m in enumerate(self.moltypes):
            c = mts[i]
            c.n_molecules = m.n_molecules
            c.mol_base = ptr(_i32(m.mol_base), C.c_int32)
            br = _i32(m.bond_rel).reshape(-1, 2) if len(m.bond_rel) else _i32([])
            c.n_bond_types = br.shape[0] if br.size else 0
            c.bond_rel = ptr(br, C.c_int32)
            c.n_ua_atoms = len(m.ua_kind)
            c.ua_kind = ptr(_i32(m.ua_kind), C.c_int32)
            ur = _i32(m.ua_rel).reshape(-1, 4) if len(m.ua_rel) else _i32([])
            c.ua_rel = ptr(ur, C.c_int32)
            c.head_rel = int(m.head_rel)
            c.n_methyls = len(m.methyl_rel)
            c.methyl_rel = ptr(_i32(m.methyl_rel), C.c_int32)
            c.normal_head_rel = int(m.normal_head_rel)
            if m.manual_leaflets is not None:
                ml = np.ascontiguousarray(m.manual_leaflets, dtype=np.uint8).reshape(-1, m.n_molecules)
                c.n_manual_leaflet_frames = ml.shape[0]
                c.manual_leaflets = ptr(ml, C.c_uint8)
            if m.manual_normals is not None:
                mn = np.ascontiguousarray(m.manual_normals, dtype=np.float32).reshape(-1, m.n_molecules, 3)
                c.n_manual_normal_frames = mn.shape[0]
                c.manual_normals = ptr(mn, C.c_float)
        keep.append(mts)

        s = CGorderSetup()
        s.abi_version = ABI_VERSION
        s.kind = self.kind
        s.n_atoms = int(self.n_atoms)
        s.handle_pbc = int(bool(self.handle_pbc))
        s.step = int(self.step)
        s.n_moltypes = len(self.moltypes)
        s.moltypes = C.cast(mts, C.POINTER(CGorderMolType))
        s.normal_mode = self.normal_mode
        s.normal_axis = self.normal_axis
        s.dynamic_radius = float(self.dynamic_radius)
        nh = _i32(self.normal_heads)
        s.n_normal_heads = nh.size
        s.normal_heads = ptr(nh, C.c_int32)
        s.collect_normals = int(bool(self.collect_normals))
        s.leaflet_mode = self.leaflet_mode
        s.leaflet_axis = self.leaflet_axis
        s.leaflet_freq_kind = self.leaflet_freq_kind
        s.leaflet_freq = int(self.leaflet_freq)
        s.leaflet_flip = int(bool(self.leaflet_flip))
        s.leaflet_radius = float(self.leaflet_radius)
        mem = _i32(self.membrane)
        s.n_membrane = mem.size
        s.membrane = ptr(mem, C.c_int32)
        s.collect_leaflets = int(bool(self.collect_leaflets))
        s.geom_kind = self.geom_kind
        s.geom_invert = int(bool(self.geom_invert))
        s.geom_ref_kind = self.geom_ref_kind
        s.geom_ref_point = (C.c_float * 3)(*[float(x) for x in self.geom_ref_point])
        gr = _i32(self.geom_ref)
        s.n_geom_ref = gr.size
        s.geom_ref = ptr(gr, C.c_int32)
        dims = list(self.geom_dims) + [0.0] * (6 - len(self.geom_dims))
        s.geom_dims = (C.c_float * 6)(*[float(x) for x in dims])
        s.geom_axis = self.geom_axis
        s.structure_box = (C.c_float * 3)(*[float(x) for x in self.structure_box])
        s.map_enabled = int(bool(self.map_enabled))
        s.map_plane = self.map_plane
        s.map_span_x = (C.c_float * 2)(*[float(x) for x in self.map_span_x])
        s.map_span_y = (C.c_float * 2)(*[float(x) for x in self.map_span_y])
        s.map_bin = (C.c_float * 2)(*[float(x) for x in self.map_bin])
        s.timewise = int(bool(self.timewise))
        s.device = int(self.device)
        s.max_batch_frames = int(self.max_batch_frames)
        return s


@dataclass
class RawResults:
    """Accumulators as returned through ``GorderResults`` (numpy views owned by Python)."""

    n_slots: int
    n_frames: int
    sum: np.ndarray                 # [n_slots][3] int64
    count: np.ndarray               # [n_slots][3] uint64
    tw_sum: Optional[np.ndarray] = None     # [n_frames][n_slots][3]
    tw_count: Optional[np.ndarray] = None
    tw_frame_index: Optional[np.ndarray] = None
    map_sum: Optional[np.ndarray] = None    # [n_slots][3][nx][ny]
    map_count: Optional[np.ndarray] = None
    map_shape: tuple = (0, 0)
    leaflets: Optional[np.ndarray] = None   # [n_leaflet_frames][n_molecules_total]
    leaflet_frame_index: Optional[np.ndarray] = None
    normals: Optional[np.ndarray] = None    # [n_frames][n_molecules_total][3]


def fetch_results(lib, handle, prefix: str, setup: EngineSetup, totals_only: bool = False) -> RawResults:
    """Allocate arrays from ``*_result_sizes`` and call ``*_finish`` (shared by engine and oracle).  ``totals_only``: fetch the
    running sums / counts only (the arrays of GorderResults left NULL are skipped by the library)."""
    r = CGorderResults()
    rc = getattr(lib, prefix + "_result_sizes")(handle, C.byref(r))
    if rc != OK:
        raise GorderError(rc)
    ns, nf, nb, nm = r.n_slots, r.n_frames, r.n_map_bins, r.n_molecules_total
    out = RawResults(n_slots=ns, n_frames=nf, sum=np.zeros((ns, 3), np.int64), count=np.zeros((ns, 3), np.uint64))
    keep = [out.sum, out.count]
    r.sum = out.sum.ctypes.data_as(_i64p)
    r.count = out.count.ctypes.data_as(_u64p)
    out.tw_frame_index = np.zeros(nf, np.int64)
    r.tw_frame_index = out.tw_frame_index.ctypes.data_as(_i64p)
    if totals_only:
        rc = getattr(lib, prefix + "_finish")(handle, C.byref(r))
        if rc != OK:
            raise GorderError(rc)
        return out
    if setup.timewise:
        out.tw_sum = np.zeros((nf, ns, 3), np.int64)
        out.tw_count = np.zeros((nf, ns, 3), np.uint64)
        r.tw_sum = out.tw_sum.ctypes.data_as(_i64p)
        r.tw_count = out.tw_count.ctypes.data_as(_u64p)
    if setup.map_enabled:
        out.map_shape = (int(r.map_nx), int(r.map_ny))
        out.map_sum = np.zeros((ns, 3, int(r.map_nx), int(r.map_ny)), np.int64)
        out.map_count = np.zeros((ns, 3, int(r.map_nx), int(r.map_ny)), np.uint64)
        r.map_sum = out.map_sum.ctypes.data_as(_i64p)
        r.map_count = out.map_count.ctypes.data_as(_u64p)
    if setup.collect_leaflets and r.n_leaflet_frames:
        out.leaflets = np.zeros((int(r.n_leaflet_frames), nm), np.uint8)
        out.leaflet_frame_index = np.zeros(int(r.n_leaflet_frames), np.int64)
        r.leaflets = out.leaflets.ctypes.data_as(_u8p)
        r.leaflet_frame_index = out.leaflet_frame_index.ctypes.data_as(_i64p)
    if setup.collect_normals and setup.normal_mode == NORMAL_DYNAMIC:
        out.normals = np.zeros((nf, nm, 3), np.float32)
        r.normals = out.normals.ctypes.data_as(_f32p)
    rc = getattr(lib, prefix + "_finish")(handle, C.byref(r))
    if rc != OK:
        raise GorderError(rc)
    del keep
    return out


class GorderError(RuntimeError):
    """Raised for any non-zero return code of the C ABI (maps onto ``AnalysisError``)."""

    def __init__(self, code: int, detail: str = "", index: int = -1):
        self.code = int(code)
        self.index = int(index)
        name = ERROR_NAMES.get(self.code, f"error {self.code}")
        msg = f"{name} (code {self.code})"
        if index >= 0:
            msg += f" [detail {index}]"
        if detail:
            msg += f": {detail}"
        super().__init__(msg)
