"""Conversion of raw accumulators into presentable order parameters.

Host-side restatement of the step that follows the hot path in the reference
(``src/presentation/converter.rs:52-559``): integer mean, block-averaging error, sample-weighted
atom / molecule / system averages (``OrderSummer``), order-map division and convergence prefix
averages.  In the drop-in deployment this is done by the untouched Rust converter from the
back-filled ``SystemTopology``; it is provided here so that the C++/Python harness can produce
final numbers (and so that the parity tests can compare against the reference's YAML fixtures).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import abi

PRECISION = 1_000_000


def calc_order(total: int, n: int, min_samples: int = 1) -> float:
    """``AnalysisOrder::calc_order`` (order.rs:97-107): integer division truncating toward zero
    (order.rs:34-41), then ``/1e6`` in f64 rounded to f32."""
    total, n = int(total), int(n)
    if n < max(1, min_samples):
        return float("nan")
    q = abs(total) // n
    q = q if total >= 0 else -q
    return float(np.float32(q / 1e6))


def estimate_error(tw_sum: np.ndarray, tw_count: np.ndarray, n_blocks: int) -> Optional[float]:
    """``TimeWiseData::estimate_error`` (timewise.rs:191-231): sample standard deviation (n-1, f32,
    ``statistical::standard_deviation``) of the block means; NaN if a block has no samples."""
    n_frames = len(tw_sum)
    if n_frames == 0:
        return None
    if n_blocks < 2 or n_blocks > n_frames:   # timewise.rs:196-199 panics below two blocks; block_size 0 divides by zero
        raise abi.GorderError(abi.ERR_INVALID_ARGUMENT, f"cannot estimate the error from {n_blocks} blocks of {n_frames} frames")
    block = n_frames // n_blocks
    vals = []
    for b in range(n_blocks):
        lo, hi = b * block, (b + 1) * block
        s, c = int(np.sum(tw_sum[lo:hi], dtype=np.int64)), int(np.sum(tw_count[lo:hi], dtype=np.uint64))
        if c == 0:
            return float("nan")
        vals.append(np.float32(calc_order(s, c)))
    # sequential f32 folds, as `statistical 1.0.0` does them (mean = fold(+) / n, variance = fold(+ d^2) / (n - 1))
    with np.errstate(divide="ignore", invalid="ignore"):
        total = np.float32(0)
        for x in vals:
            total = np.float32(total + x)
        mean = np.float32(total / np.float32(n_blocks))
        dev2 = np.float32(0)
        for x in vals:
            d = np.float32(x - mean)
            dev2 = np.float32(dev2 + np.float32(d * d))
        var = np.float32(dev2 / np.float32(n_blocks - 1))
        return float(np.sqrt(var, dtype=np.float32))


def prefix_average(tw_sum: np.ndarray, tw_count: np.ndarray) -> np.ndarray:
    """``TimeWiseData::prefix_average`` (timewise.rs:259-274)."""
    cs, cc = np.cumsum(tw_sum.astype(np.int64)), np.cumsum(tw_count.astype(np.uint64))
    return np.array([calc_order(s, c) if c > 0 else np.nan for s, c in zip(cs, cc)], np.float32)


@dataclass
class Order:
    value: float
    error: Optional[float] = None


@dataclass
class OrderCollection:
    total: Optional[Order] = None
    upper: Optional[Order] = None
    lower: Optional[Order] = None


@dataclass
class ItemResults:
    """A bond type (AA/CG) or a united atom with its virtual bonds."""

    label: str
    order: OrderCollection
    bonds: List[OrderCollection] = field(default_factory=list)
    maps: Optional[np.ndarray] = None   # [3][nx][ny] f32 (NaN where samples < min); AA / UA: the bonds of the atom merged
    bond_maps: List[np.ndarray] = field(default_factory=list)   # AA / UA: one [3][nx][ny] per bond of the atom


@dataclass
class MoleculeResults:
    name: str
    average: OrderCollection
    items: List[ItemResults]
    convergence: Optional[Dict[str, np.ndarray]] = None


@dataclass
class AnalysisResults:
    average: OrderCollection
    molecules: Dict[str, MoleculeResults]
    n_frames: int


class _PythonBackend:
    """``OrderSummer`` (converter.rs:513-559) + the divisions, in numpy."""

    def __init__(self, raw: abi.RawResults, sign: float, leaflets: bool, n_blocks: Optional[int], min_samples: int, map_min_samples: int):
        self.raw, self.sign, self.leaflets, self.nb, self.min_samples, self.map_min = raw, sign, leaflets, n_blocks, min_samples, map_min_samples
        self.tw = raw.tw_sum is not None

    def _sums(self, slots):
        raw, sl = self.raw, np.asarray(slots, np.int64)
        sm, ct = raw.sum[sl].sum(axis=0, dtype=np.int64), raw.count[sl].sum(axis=0, dtype=np.uint64)
        if not self.tw:
            return sm, ct, None, None
        return sm, ct, raw.tw_sum[:, sl, :].sum(axis=1, dtype=np.int64), raw.tw_count[:, sl, :].sum(axis=1, dtype=np.uint64)

    def collection(self, slots) -> OrderCollection:
        sm, ct, ts, tc = self._sums(slots)
        out = []
        for k in range(3 if self.leaflets else 1):
            v = calc_order(sm[k], ct[k], self.min_samples)
            err = None
            if self.nb is not None and ts is not None:
                err = estimate_error(ts[:, k], tc[:, k], self.nb)
                if err is not None and int(ct[k]) < self.min_samples:
                    err = float("nan")
            out.append(Order(self.sign * v if v == v else v, err))
        return OrderCollection(*out)

    def convergence(self, slots):
        _sm, _ct, ts, tc = self._sums(slots)
        return {k: np.array([self.sign * x for x in prefix_average(ts[:, j], tc[:, j])], np.float32)
                for j, k in enumerate(["total", "upper", "lower"][: 3 if self.leaflets else 1])}

    def slots_map(self, slots):
        """Map of the samples of `slots` together (one bond; the bonds of an atom: ordermap.rs:116-138 `Add`)."""
        raw, sl = self.raw, np.asarray(slots, np.int64)
        ms, mc = raw.map_sum[sl].sum(axis=0, dtype=np.int64), raw.map_count[sl].sum(axis=0, dtype=np.uint64)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = (ms.astype(np.float64) / 1e6).astype(np.float32) / mc.astype(np.float32)
        return np.where(mc < self.map_min, np.nan, self.sign * val).astype(np.float32)


class _NativeBackend:
    """The same through the C ABI (``gorder_results_*``, include/gorder_b200.h; csrc/gorder_results.inl)."""

    def __init__(self, raw: abi.RawResults, sign: float, leaflets: bool, n_blocks: Optional[int], min_samples: int, map_min_samples: int):
        import ctypes as C
        from . import _lib
        self.C, self.lib = C, _lib.lib()
        self.raw, self.sign, self.leaflets, self.nb, self.min_samples, self.map_min = raw, sign, leaflets, n_blocks, min_samples, map_min_samples
        self.tw = raw.tw_sum is not None
        self._keep = [np.ascontiguousarray(raw.sum, np.int64), np.ascontiguousarray(raw.count, np.uint64)]
        r = abi.CGorderRaw()
        r.n_slots, r.n_frames = raw.n_slots, (raw.n_frames if self.tw else 0)
        r.sum, r.count = self._keep[0].ctypes.data, self._keep[1].ctypes.data
        if self.tw:
            self._keep += [np.ascontiguousarray(raw.tw_sum, np.int64), np.ascontiguousarray(raw.tw_count, np.uint64)]
            r.tw_sum, r.tw_count = self._keep[2].ctypes.data, self._keep[3].ctypes.data
        self.r = r

    def _check(self, rc):
        if rc != abi.OK:
            raise abi.GorderError(rc)

    def collection(self, slots) -> OrderCollection:
        C = self.C
        sl = np.ascontiguousarray(slots, np.int32)
        val, err = np.zeros(3, np.float32), np.zeros(3, np.float32)
        want_err = self.nb is not None and self.tw
        self._check(self.lib.gorder_results_order(C.byref(self.r), sl.ctypes.data, len(sl), self.nb if want_err else 0, self.min_samples,
                                                  self.sign, val.ctypes.data, err.ctypes.data if want_err else None))
        n = 3 if self.leaflets else 1
        no_frames = self.raw.n_frames == 0
        return OrderCollection(*[Order(float(val[k]), (None if no_frames else float(err[k])) if want_err else None) for k in range(n)])

    def convergence(self, slots):
        sl = np.ascontiguousarray(slots, np.int32)
        out = np.zeros((self.raw.n_frames, 3), np.float32)
        self._check(self.lib.gorder_results_convergence(self.C.byref(self.r), sl.ctypes.data, len(sl), self.sign, out.ctypes.data))
        return {k: out[:, j].copy() for j, k in enumerate(["total", "upper", "lower"][: 3 if self.leaflets else 1])}

    def slots_map(self, slots):
        sl = np.asarray(slots, np.int64)
        ms = np.ascontiguousarray(self.raw.map_sum[sl].sum(axis=0, dtype=np.int64))
        mc = np.ascontiguousarray(self.raw.map_count[sl].sum(axis=0, dtype=np.uint64))
        out = np.zeros(ms.shape, np.float32)
        self._check(self.lib.gorder_results_map(ms.ctypes.data, mc.ctypes.data, ms.size, self.map_min, self.sign, out.ctypes.data))
        return out


def convert(raw: abi.RawResults, setup: abi.EngineSetup, *, n_blocks: Optional[int] = None, min_samples: int = 1,
            map_min_samples: int = 1, native: bool = False) -> AnalysisResults:
    """``ResultsConverter::convert_topology`` (converter.rs:52-85).  ``native``: the divisions, block errors and prefix
    averages run in the shared library (``gorder_results_*``) instead of numpy; the tree is built here either way."""
    sign = 1.0 if setup.kind == abi.KIND_CG else -1.0   # AA / UA report -S_CH (presentation/mod.rs:618-691)
    leaflets = setup.leaflet_mode != abi.LEAFLET_NONE
    tw = setup.timewise and raw.tw_sum is not None
    if not tw and raw.tw_sum is not None:
        raw = abi.RawResults(**{**raw.__dict__, "tw_sum": None, "tw_count": None})
    be = (_NativeBackend if native else _PythonBackend)(raw, sign, leaflets, n_blocks if tw else None, min_samples, map_min_samples)
    molecules: Dict[str, MoleculeResults] = {}
    system_slots: List[int] = []

    def slots_map(slots):
        return None if raw.map_sum is None else be.slots_map(slots)

    def bond_maps(slots):
        return [] if raw.map_sum is None else [be.slots_map([q]) for q in slots]

    for (s0, n), mt in zip(setup.slot_ranges(), setup.moltypes):
        mol_slots: List[int] = []
        items: List[ItemResults] = []
        if setup.kind == abi.KIND_UA:
            s = s0
            for i, kind in enumerate(mt.ua_kind):
                atom_slots = list(range(s, s + abi.ua_hydrogens(kind)))
                s += len(atom_slots)
                mol_slots += atom_slots
                label = mt.bond_names[i] if i < len(mt.bond_names) else f"atom {i}"
                items.append(ItemResults(label, be.collection(atom_slots), [be.collection([q]) for q in atom_slots],
                                         maps=slots_map(atom_slots), bond_maps=bond_maps(atom_slots)))
        elif setup.kind == abi.KIND_CG:
            for b in range(n):
                mol_slots.append(s0 + b)
                label = mt.bond_names[b] if b < len(mt.bond_names) else f"bond {b}"
                items.append(ItemResults(label, be.collection([s0 + b]), maps=slots_map([s0 + b])))
        else:
            # AA: bonds grouped by their heavy atom (converter.rs:325-352); bonds are sorted by atom1 (bond.rs:77-81)
            by_atom: Dict[int, List[int]] = {}
            for b, (a1, _a2) in enumerate(mt.bond_rel):
                by_atom.setdefault(int(a1), []).append(b)
            for a1 in sorted(by_atom):
                atom_slots = [s0 + b for b in by_atom[a1]]
                mol_slots += atom_slots
                label = (mt.bond_names[by_atom[a1][0]].split(" - ")[0] if by_atom[a1][0] < len(mt.bond_names) else f"atom {a1}")
                items.append(ItemResults(label, be.collection(atom_slots), [be.collection([q]) for q in atom_slots],
                                         maps=slots_map(atom_slots), bond_maps=bond_maps(atom_slots)))
        conv = be.convergence(mol_slots) if tw else None
        molecules[mt.name] = MoleculeResults(mt.name, be.collection(mol_slots), items, conv)
        system_slots += mol_slots
    return AnalysisResults(be.collection(system_slots), molecules, raw.n_frames)
