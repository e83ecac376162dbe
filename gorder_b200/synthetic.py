"""Deterministic synthetic membranes of the shapes named in BASELINE.json.

There are no trajectories of the benchmark sizes in the repository (and no network), so the bench
and the large parity tests run on generated bilayers: lipids on a lattice in two leaflets, every
lipid a chain of beads/atoms laid along a per-lipid, per-frame director with thermal noise, whole
lipids wrapped into the box so that a few per cent of the bonds cross the periodic boundary, and a
per-frame box jitter.  Frames are reproducible independently (counter-based Philox stream keyed by
``(seed, frame)``), so any GPU can regenerate exactly the frames of its own shard.

Only inputs are generated here; nothing in this module computes order parameters.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi

SEED = 20240917


@dataclass
class LipidTemplate:
    """A lipid as a set of sites laid along its director.

    ``depth[k]`` (in units of ``spacing``) along the director and ``lateral[k]`` (nm) along a fixed
    perpendicular place site k; ``bonds`` are (i, j) site pairs, i < j.
    """

    name: str
    site_names: List[str]
    depth: np.ndarray
    lateral: np.ndarray
    spacing: float
    bonds: List[Tuple[int, int]]
    head: int
    methyls: List[int] = field(default_factory=list)
    # united-atom carbons: (kind, target, helper1, helper2, helper3)
    ua: List[Tuple[int, int, int, int, int]] = field(default_factory=list)
    # sites of the full lipid that are not part of the analysis (heads etc.) still occupy slots
    n_sites: int = 0

    def __post_init__(self):
        if not self.n_sites:
            self.n_sites = len(self.site_names)


def martini_popc() -> LipidTemplate:
    """Martini POPC: 12 beads, 11 bonds (bond order as validation/cg_martini/.../order.yaml)."""
    names = ["NC3", "PO4", "GL1", "GL2", "C1A", "D2A", "C3A", "C4A", "C1B", "C2B", "C3B", "C4B"]
    depth = np.array([-1, 0, 1, 1, 2, 3, 4, 5, 2, 3, 4, 5], np.float32)
    lateral = np.array([0, 0, 0, 0.35, 0, 0, 0, 0, 0.35, 0.35, 0.35, 0.35], np.float32)
    bonds = [(0, 1), (1, 2), (2, 3), (2, 4), (3, 8), (4, 5), (5, 6), (6, 7), (8, 9), (9, 10), (10, 11)]
    return LipidTemplate("POPC", names, depth, lateral, 0.47, bonds, head=1, methyls=[7, 11])


def charmm_like_popc() -> LipidTemplate:
    """All-atom POPC-like lipid: 134 atoms, of which 32 carbons carry 64 hydrogens
    (28 CH2, 2 CH, 2 CH3) = 64 C-H bond types; the remaining 38 atoms are head-group filler."""
    names, depth, lateral, bonds = [], [], [], []
    # 38 filler atoms first (head group): site 0 is "P"
    for k in range(38):
        names.append("P" if k == 0 else f"X{k}")
        depth.append(-1.0 + 0.05 * k)
        lateral.append(0.1 * (k % 5))
    n_h = [2] * 14 + [1, 1] + [2] * 14 + [3, 3]
    methyls = []
    for ci, nh in enumerate(n_h):
        c = len(names)
        chain = 0 if ci < 16 else 1
        pos = ci if ci < 16 else ci - 16
        names.append(f"C{ci + 1}")
        depth.append(1.0 + pos)
        lateral.append(0.45 * chain)
        if nh == 3:
            methyls.append(c)
        for hi in range(nh):
            names.append(f"H{ci + 1}{'ABC'[hi]}")
            # hydrogens sit ~0.109 nm off the carbon: 0.7 spacing units sideways/along
            depth.append(1.0 + pos + (0.35 if hi == 2 else 0.0))
            lateral.append(0.45 * chain + (0.109 if hi == 0 else (-0.109 if hi == 1 else 0.0)))
            bonds.append((c, len(names) - 1))
    return LipidTemplate("POPC", names, np.array(depth, np.float32), np.array(lateral, np.float32), 0.13, bonds,
                         head=0, methyls=methyls)


def berger_like_popc(with_ch1_sat: bool = False) -> LipidTemplate:
    """United-atom POPC-like lipid: 52 heavy atoms; two zig-zag chains of 16 carbons:
    28 CH2 + 2 CH3 + 2 CH1 (double bond) = 64 virtual C-H bonds (optionally one CH2 -> CH1 saturated)."""
    names, depth, lateral = [], [], []
    for k in range(20):   # head-group filler; site 0 = "P"
        names.append("P" if k == 0 else f"X{k}")
        depth.append(-1.0 + 0.08 * k)
        lateral.append(0.12 * (k % 4))
    ua = []
    methyls = []
    chains = []
    for chain in range(2):
        idx = []
        for pos in range(16):
            idx.append(len(names))
            names.append(f"C{chain + 1}{pos + 1}")
            depth.append(1.0 + pos)
            # zig-zag: alternate sideways displacement so that helper vectors are not collinear
            lateral.append(0.5 * chain + (0.06 if pos % 2 else -0.06))
        chains.append(idx)
    anchor = [19, 18]   # filler atoms bonded to the first carbon of each chain
    branch = 17         # a third neighbour for the optional CH1-saturated carbon
    for chain, idx in enumerate(chains):
        for pos, c in enumerate(idx):
            prev = idx[pos - 1] if pos > 0 else anchor[chain]
            if pos == 15:
                ua.append((abi.UA_CH3, c, prev, idx[pos - 2], -1))
                methyls.append(c)
            elif chain == 0 and pos in (8, 9):
                ua.append((abi.UA_CH1_UNSAT, c, prev, idx[pos + 1], -1))
            elif with_ch1_sat and chain == 1 and pos == 3:
                ua.append((abi.UA_CH1_SAT, c, prev, idx[pos + 1], branch))
            else:
                ua.append((abi.UA_CH2, c, prev, idx[pos + 1], -1))
    ua.sort(key=lambda e: e[1])
    return LipidTemplate("POPC", names, np.array(depth, np.float32), np.array(lateral, np.float32), 0.127, [],
                         head=0, methyls=methyls, ua=ua)


@dataclass
class SyntheticSystem:
    """A generated bilayer: topology (``setup``) + a frame generator."""

    template: LipidTemplate
    n_lipids: int
    n_water: int
    box: np.ndarray
    head_xy: np.ndarray       # [n_lipids][2]
    leaflet_sign: np.ndarray  # +1 upper, -1 lower
    setup: abi.EngineSetup
    seed: int = SEED
    tilt_sigma_deg: float = 25.0
    noise: float = 0.05
    half_thickness: float = 2.0

    @property
    def n_atoms(self) -> int:
        return self.setup.n_atoms

    def frame(self, f: int) -> Tuple[np.ndarray, np.ndarray]:
        """(xyz [n_atoms][3] f32, box [3] f32) of trajectory frame ``f``."""
        rng = np.random.Generator(np.random.Philox(key=self.seed, counter=[0, 0, 0, f]))
        t = self.template
        n, ns = self.n_lipids, t.n_sites
        box = (self.box * (1.0 + 0.005 * (2.0 * rng.random() - 1.0))).astype(np.float32)
        scale = box / self.box
        # director: tilted about -sign * z
        tilt = np.abs(rng.normal(0.0, np.deg2rad(self.tilt_sigma_deg), n))
        phi = rng.uniform(0.0, 2.0 * np.pi, n)
        ux, uy, uz = np.sin(tilt) * np.cos(phi), np.sin(tilt) * np.sin(phi), -self.leaflet_sign * np.cos(tilt)
        u = np.stack([ux, uy, uz], 1)
        # a perpendicular for the lateral offsets
        perp = np.cross(u, np.array([0.0, 0.0, 1.0]))
        pn = np.linalg.norm(perp, axis=1, keepdims=True)
        perp = np.where(pn > 1e-6, perp / np.maximum(pn, 1e-6), np.array([1.0, 0.0, 0.0]))
        head = np.empty((n, 3))
        head[:, :2] = self.head_xy * scale[:2] + rng.normal(0.0, 0.05, (n, 2))
        head[:, 2] = 0.5 * box[2] + self.leaflet_sign * self.half_thickness + rng.normal(0.0, 0.1, n)
        pos = (head[:, None, :] + (t.depth * t.spacing)[None, :, None] * u[:, None, :]
               + t.lateral[None, :, None] * perp[:, None, :] + rng.normal(0.0, self.noise, (n, ns, 3)))
        # whole lipids are wrapped by their head: bonds near the box edge cross the boundary
        shift = np.floor(pos[:, t.head, :] / box) * box
        pos -= shift[:, None, :]
        # additionally wrap every site of ~3% of the lipids individually (broken molecules)
        broken = rng.random(n) < 0.03
        pos[broken] -= np.floor(pos[broken] / box) * box
        xyz = np.empty((self.n_atoms, 3), np.float32)
        xyz[: n * ns] = pos.reshape(n * ns, 3)
        if self.n_water:
            xyz[n * ns:] = rng.random((self.n_water, 3)) * box
        return xyz, box

    def frames(self, first: int, count: int, step: int = 1) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        xyz = np.empty((count, self.n_atoms, 3), np.float32)
        box = np.empty((count, 3), np.float32)
        idx = np.arange(count, dtype=np.int64) * step + first
        for i, f in enumerate(idx):
            xyz[i], box[i] = self.frame(int(f))
        return xyz, box, idx


def make_bilayer(template: LipidTemplate, n_lipids: int, kind: int, *, area_per_lipid: float = 0.61, box_z: float = 12.0,
                 n_water: int = 0, seed: int = SEED, split_types: int = 1, **setup_kw) -> SyntheticSystem:
    """Build a bilayer of ``n_lipids`` (half per leaflet) and its ``EngineSetup``.

    ``split_types`` > 1 declares the lipids as that many molecule types (round-robin blocks), which
    exercises the multi-type paths with otherwise identical chemistry.
    """
    per_leaf = (n_lipids + 1) // 2
    side = int(np.ceil(np.sqrt(per_leaf)))
    a = float(np.sqrt(area_per_lipid))
    box = np.array([side * a, side * a, box_z], np.float32)
    ij = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2)
    xy_leaf = (ij[:per_leaf] + 0.5) * a
    head_xy = np.concatenate([xy_leaf, xy_leaf[: n_lipids - per_leaf]], 0)
    sign = np.concatenate([np.ones(per_leaf), -np.ones(n_lipids - per_leaf)])
    ns = template.n_sites
    n_atoms = n_lipids * ns + n_water
    bases = np.arange(n_lipids, dtype=np.int64) * ns
    mts = []
    for k in range(split_types):
        # blocks: upper-leaflet part and lower-leaflet part of every type, like insane-built membranes
        sel = np.concatenate([np.arange(per_leaf)[k::split_types], per_leaf + np.arange(n_lipids - per_leaf)[k::split_types]])
        sel.sort()
        name = template.name if split_types == 1 else f"{template.name}{k + 1}"
        common = dict(name=name, mol_base=bases[sel].tolist(), head_rel=template.head, methyl_rel=template.methyls,
                      normal_head_rel=template.head)
        if kind == abi.KIND_UA:
            mts.append(abi.MolType(ua_kind=[e[0] for e in template.ua], ua_rel=[e[1:] for e in template.ua],
                                   bond_names=[template.site_names[e[1]] for e in template.ua], **common))
        else:
            mts.append(abi.MolType(bond_rel=sorted(template.bonds),
                                   bond_names=[f"{template.site_names[i]}-{template.site_names[j]}" for i, j in sorted(template.bonds)],
                                   **common))
    lipid_atoms = np.arange(n_lipids * ns, dtype=np.int32)
    heads = (bases + template.head).astype(np.int32)
    setup_kw.setdefault("membrane", lipid_atoms if setup_kw.get("leaflet_mode", 0) in (abi.LEAFLET_GLOBAL, abi.LEAFLET_LOCAL) else ())
    if setup_kw.get("normal_mode", abi.NORMAL_STATIC) == abi.NORMAL_DYNAMIC:
        setup_kw.setdefault("normal_heads", heads)
    setup = abi.EngineSetup(kind=kind, n_atoms=n_atoms, moltypes=mts, **setup_kw)
    return SyntheticSystem(template, n_lipids, n_water, box, head_xy, sign, setup, seed=seed)


# -- the named benchmark shapes -----------------------------------------------------------------

def s_cg(n_lipids: int = 83334, **kw) -> SyntheticSystem:
    """BASELINE config 2: CG Martini bilayer, 1 000 008 beads, 916 674 bonds / frame, Global leaflets."""
    kw.setdefault("leaflet_mode", abi.LEAFLET_GLOBAL)
    return make_bilayer(martini_popc(), n_lipids, abi.KIND_CG, area_per_lipid=0.61, box_z=12.0, **kw)


def s_aa(n_lipids: int = 256, n_water: int = 30000, **kw) -> SyntheticSystem:
    """BASELINE config 1 (256 lipids) / config 4 (4096 lipids): AA POPC-like, 64 C-H bond types."""
    return make_bilayer(charmm_like_popc(), n_lipids, abi.KIND_AA, area_per_lipid=0.64, box_z=9.0, n_water=n_water, **kw)


def s_ua(n_lipids: int = 256, **kw) -> SyntheticSystem:
    """BASELINE config 3: UA Berger-like bilayer, 64 virtual C-H per lipid, error blocks."""
    kw.setdefault("timewise", True)
    with_sat = kw.pop("with_ch1_sat", False)
    return make_bilayer(berger_like_popc(with_sat), n_lipids, abi.KIND_UA, area_per_lipid=0.64, box_z=8.0, **kw)


# -- vesicle (BASELINE config 5, S-VES) ----------------------------------------------------------

@dataclass
class SyntheticVesicle:
    """A generated vesicle: lipids of the outer / inner leaflet on two concentric spheres (Fibonacci lattice + jitter), every
    lipid a chain of beads along its radial director (tails towards the mid-surface) with a tilt and thermal noise; the
    vesicle is centred at a per-frame random point and every bead is wrapped into the cubic box, so the vesicle is cut by the
    periodic boundary in most frames."""

    template: LipidTemplate
    n_lipids: int
    box: np.ndarray
    unit: np.ndarray           # [n_lipids][3] radial unit vector of every lipid
    is_outer: np.ndarray       # [n_lipids] bool
    r_outer: float
    r_inner: float
    setup: abi.EngineSetup
    seed: int = SEED
    tilt_sigma_deg: float = 20.0
    noise: float = 0.05
    n_water: int = 0

    @property
    def n_atoms(self) -> int:
        return self.setup.n_atoms

    def frame(self, f: int) -> Tuple[np.ndarray, np.ndarray]:
        rng = np.random.Generator(np.random.Philox(key=self.seed + 1, counter=[0, 0, 0, f]))
        t = self.template
        n, ns = self.n_lipids, t.n_sites
        box = (self.box * (1.0 + 0.004 * (2.0 * rng.random() - 1.0))).astype(np.float32)
        v = self.unit + rng.normal(0.0, 0.01, (n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        sign = np.where(self.is_outer, -1.0, 1.0)[:, None]        # director: from the head towards the mid-surface
        tilt = np.abs(rng.normal(0.0, np.deg2rad(self.tilt_sigma_deg), n))
        phi = rng.uniform(0.0, 2.0 * np.pi, n)
        e1 = np.cross(v, np.array([0.0, 0.0, 1.0]))
        e1n = np.linalg.norm(e1, axis=1, keepdims=True)
        e1 = np.where(e1n > 1e-6, e1 / np.maximum(e1n, 1e-6), np.array([1.0, 0.0, 0.0]))
        e2 = np.cross(v, e1)
        u = sign * v * np.cos(tilt)[:, None] + (e1 * np.cos(phi)[:, None] + e2 * np.sin(phi)[:, None]) * np.sin(tilt)[:, None]
        radius = np.where(self.is_outer, self.r_outer, self.r_inner)[:, None] + rng.normal(0.0, 0.1, (n, 1))
        head = v * radius
        pos = (head[:, None, :] + (t.depth * t.spacing)[None, :, None] * u[:, None, :] + t.lateral[None, :, None] * e1[:, None, :]
               + rng.normal(0.0, self.noise, (n, ns, 3)))
        centre = rng.uniform(0.0, 1.0, 3) * box
        pos = pos + centre
        pos -= np.floor(pos / box) * box                          # every bead on its own: broken molecules at the boundary
        xyz = np.empty((self.n_atoms, 3), np.float32)
        xyz[: n * ns] = pos.reshape(n * ns, 3)
        if self.n_water:
            xyz[n * ns:] = rng.random((self.n_water, 3)) * box
        return xyz, box

    def frames(self, first: int, count: int, step: int = 1) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        xyz = np.empty((count, self.n_atoms, 3), np.float32)
        box = np.empty((count, 3), np.float32)
        idx = np.arange(count, dtype=np.int64) * step + first
        for i, f in enumerate(idx):
            xyz[i], box[i] = self.frame(int(f))
        return xyz, box, idx


def _fibonacci_sphere(n: int) -> np.ndarray:
    k = np.arange(n) + 0.5
    z = 1.0 - 2.0 * k / n
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = np.pi * (1.0 + 5.0 ** 0.5) * k
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], 1)


def s_ves(n_lipids: int = 100000, *, area_per_lipid: float = 0.61, thickness: float = 4.0, margin: float = 6.0, **kw) -> SyntheticVesicle:
    """BASELINE config 5 (S-VES): CGOrder on a Martini vesicle with dynamic local membrane normals (heads = PO4, r = 2 nm) and
    spherical-clustering leaflets (reference: examples/coarse_grained/3_vesicles.yaml, tests_cg.rs:3391-3417 and the
    clustering variant after it).  The radii follow from the lipid count at 0.61 nm^2 per lipid and a 4 nm bilayer (100 000
    lipids: outer radius 54 nm, box 120 nm; SURVEY.md §8d's "R = 20 / 16 nm, ~100 k lipids" would pack 7 lipids per nm^2)."""
    t = martini_popc()
    # n_out / n_in = (R / (R - thickness))^2 and n_out * area = 4 pi R^2
    r_out = 0.5 * thickness + np.sqrt(max(0.0, n_lipids * area_per_lipid / (8.0 * np.pi) - 0.25 * thickness ** 2))
    r_out = float(max(r_out, thickness + 1.0))
    r_in = r_out - thickness
    n_out = int(round(n_lipids * r_out ** 2 / (r_out ** 2 + r_in ** 2)))
    n_in = n_lipids - n_out
    unit = np.concatenate([_fibonacci_sphere(n_out), _fibonacci_sphere(max(n_in, 1))[:n_in]], 0)
    is_outer = np.arange(n_lipids) < n_out
    rng = np.random.default_rng(SEED)
    perm = rng.permutation(n_lipids)                     # leaflets interleaved in the topology, as after self-assembly
    unit, is_outer = unit[perm], is_outer[perm]
    side = 2.0 * r_out + 2.0 * margin
    box = np.array([side, side, side], np.float32)
    ns = t.n_sites
    bases = np.arange(n_lipids, dtype=np.int64) * ns
    heads = (bases + t.head).astype(np.int32)
    mt = abi.MolType(name=t.name, mol_base=bases.tolist(), bond_rel=sorted(t.bonds), head_rel=t.head, methyl_rel=t.methyls, normal_head_rel=t.head,
                     bond_names=[f"{t.site_names[i]}-{t.site_names[j]}" for i, j in sorted(t.bonds)])
    kw.setdefault("normal_mode", abi.NORMAL_DYNAMIC)
    kw.setdefault("dynamic_radius", 2.0)
    kw.setdefault("normal_heads", heads)
    kw.setdefault("leaflet_mode", abi.LEAFLET_SPHERICAL)
    kw.setdefault("leaflet_freq_kind", abi.FREQ_ONCE)
    if kw["leaflet_mode"] == abi.LEAFLET_SPHERICAL:
        kw.setdefault("membrane", heads)                 # the ClusterHeads group
    setup = abi.EngineSetup(kind=abi.KIND_CG, n_atoms=n_lipids * ns, moltypes=[mt], **kw)
    return SyntheticVesicle(t, n_lipids, box, unit, is_outer, r_out, r_in, setup)
