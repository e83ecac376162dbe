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
    block = n_frames // n_blocks
    vals = []
    for b in range(n_blocks):
        lo, hi = b * block, (b + 1) * block
        s, c = int(np.sum(tw_sum[lo:hi], dtype=np.int64)), int(np.sum(tw_count[lo:hi], dtype=np.uint64))
        if c == 0:
            return float("nan")
        vals.append(np.float32(calc_order(s, c)))
    v = np.array(vals, np.float32)
    mean = np.float32(np.float32(v.sum(dtype=np.float32)) / np.float32(n_blocks))
    dev = (v - mean).astype(np.float32)
    var = np.float32(np.float32((dev * dev).sum(dtype=np.float32)) / np.float32(n_blocks - 1))
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
    maps: Optional[np.ndarray] = None   # [3][nx][ny] f32 (NaN where samples < min)


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


class _Summer:
    """``OrderSummer`` (converter.rs:513-559): element-wise sums of accumulators."""

    def __init__(self, n_frames: int, timewise: bool):
        self.sum = np.zeros(3, np.int64)
        self.cnt = np.zeros(3, np.uint64)
        self.tw_sum = np.zeros((n_frames, 3), np.int64) if timewise else None
        self.tw_cnt = np.zeros((n_frames, 3), np.uint64) if timewise else None

    def add_slot(self, raw: abi.RawResults, s: int):
        self.sum += raw.sum[s]
        self.cnt += raw.count[s]
        if self.tw_sum is not None:
            self.tw_sum += raw.tw_sum[:, s, :]
            self.tw_cnt += raw.tw_count[:, s, :]

    def add(self, other: "_Summer"):
        self.sum += other.sum
        self.cnt += other.cnt
        if self.tw_sum is not None:
            self.tw_sum += other.tw_sum
            self.tw_cnt += other.tw_cnt


def _collection(sign: float, sum3, cnt3, tw_sum, tw_cnt, leaflets: bool, n_blocks: Optional[int], min_samples: int) -> OrderCollection:
    out = []
    for k in range(3 if leaflets else 1):
        v = calc_order(sum3[k], cnt3[k], min_samples)
        err = None
        if n_blocks is not None and tw_sum is not None:
            err = estimate_error(tw_sum[:, k], tw_cnt[:, k], n_blocks)
            if err is not None and int(cnt3[k]) < min_samples:
                err = float("nan")
        out.append(Order(sign * v if v == v else v, err))
    return OrderCollection(*out)


def convert(raw: abi.RawResults, setup: abi.EngineSetup, *, n_blocks: Optional[int] = None, min_samples: int = 1,
            map_min_samples: int = 1) -> AnalysisResults:
    """``ResultsConverter::convert_topology`` (converter.rs:52-85)."""
    sign = 1.0 if setup.kind == abi.KIND_CG else -1.0   # AA / UA report -S_CH (presentation/mod.rs:618-691)
    leaflets = setup.leaflet_mode != abi.LEAFLET_NONE
    tw = setup.timewise and raw.tw_sum is not None
    nb = n_blocks if tw else None
    nf = raw.n_frames
    system = _Summer(nf, tw)
    molecules: Dict[str, MoleculeResults] = {}

    def coll(sm: _Summer) -> OrderCollection:
        return _collection(sign, sm.sum, sm.cnt, sm.tw_sum, sm.tw_cnt, leaflets, nb, min_samples)

    def slot_coll(s: int) -> OrderCollection:
        return _collection(sign, raw.sum[s], raw.count[s], raw.tw_sum[:, s, :] if tw else None, raw.tw_count[:, s, :] if tw else None,
                           leaflets, nb, min_samples)

    def slot_map(s: int):
        if raw.map_sum is None:
            return None
        with np.errstate(divide="ignore", invalid="ignore"):
            val = (raw.map_sum[s].astype(np.float64) / 1e6).astype(np.float32) / raw.map_count[s].astype(np.float32)
        val = np.where(raw.map_count[s] < map_min_samples, np.nan, sign * val).astype(np.float32)
        return val

    for (s0, n), mt in zip(setup.slot_ranges(), setup.moltypes):
        mol = _Summer(nf, tw)
        items: List[ItemResults] = []
        if setup.kind == abi.KIND_UA:
            s = s0
            for i, kind in enumerate(mt.ua_kind):
                atom = _Summer(nf, tw)
                bonds = []
                for _ in range(abi.ua_hydrogens(kind)):
                    atom.add_slot(raw, s)
                    bonds.append(slot_coll(s))
                    s += 1
                mol.add(atom)
                label = mt.bond_names[i] if i < len(mt.bond_names) else f"atom {i}"
                items.append(ItemResults(label, coll(atom), bonds))
        elif setup.kind == abi.KIND_CG:
            for b in range(n):
                mol.add_slot(raw, s0 + b)
                label = mt.bond_names[b] if b < len(mt.bond_names) else f"bond {b}"
                items.append(ItemResults(label, slot_coll(s0 + b), maps=slot_map(s0 + b)))
        else:
            # AA: bonds grouped by their heavy atom (converter.rs:325-352); bonds are sorted by atom1 (bond.rs:77-81)
            by_atom: Dict[int, List[int]] = {}
            for b, (a1, _a2) in enumerate(mt.bond_rel):
                by_atom.setdefault(int(a1), []).append(b)
            for a1 in sorted(by_atom):
                atom = _Summer(nf, tw)
                bonds = []
                for b in by_atom[a1]:
                    atom.add_slot(raw, s0 + b)
                    bonds.append(slot_coll(s0 + b))
                mol.add(atom)
                label = (mt.bond_names[by_atom[a1][0]].split(" - ")[0] if by_atom[a1][0] < len(mt.bond_names) else f"atom {a1}")
                items.append(ItemResults(label, coll(atom), bonds))
        conv = None
        if tw:
            conv = {k: np.array([sign * x for x in prefix_average(mol.tw_sum[:, j], mol.tw_cnt[:, j])], np.float32)
                    for j, k in enumerate(["total", "upper", "lower"][: 3 if leaflets else 1])}
        molecules[mt.name] = MoleculeResults(mt.name, coll(mol), items, conv)
        system.add(mol)
    return AnalysisResults(coll(system), molecules, nf)
