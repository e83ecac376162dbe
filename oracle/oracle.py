"""Python wrapper of the CPU oracle (``oracle/gorder_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by ``gorder_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from gorder_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None


def build(force: bool = False) -> str:
    """Compile ``libgorder_oracle.so`` / ``libxtc.so`` with the Makefile next to this file."""
    so = os.path.join(_HERE, "libgorder_oracle.so")
    srcs = [os.path.join(_HERE, "gorder_oracle.c"), os.path.join(_HERE, "..", "include", "gorder_b200.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    xtc_so, xtc_c = os.path.join(_HERE, "libxtc.so"), os.path.join(_HERE, "xtc.c")
    if os.path.exists(xtc_c) and (not os.path.exists(xtc_so) or os.path.getmtime(xtc_c) > os.path.getmtime(xtc_so)):
        stale = True
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.gorder_oracle_calc_sch.restype = C.c_float
        L.gorder_oracle_calc_sch.argtypes = [C.POINTER(C.c_float)] * 2
        L.gorder_oracle_order_value.restype = C.c_int64
        L.gorder_oracle_order_value.argtypes = [C.c_float]
        L.gorder_oracle_calc_order.restype = C.c_float
        L.gorder_oracle_calc_order.argtypes = [C.c_int64, C.c_uint64, C.c_uint64]
        L.gorder_oracle_create.argtypes = [C.POINTER(abi.CGorderSetup), C.POINTER(C.c_void_p)]
        L.gorder_oracle_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.gorder_oracle_finish.argtypes = [C.c_void_p, C.POINTER(abi.CGorderResults)]
        L.gorder_oracle_result_sizes.argtypes = [C.c_void_p, C.POINTER(abi.CGorderResults)]
        L.gorder_oracle_destroy.argtypes = [C.c_void_p]
        L.gorder_oracle_error_detail.restype = C.c_int64
        L.gorder_oracle_error_detail.argtypes = [C.c_void_p]
        L.gorder_oracle_set_leaflets.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.gorder_oracle_estimate_error.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float)]
        L.gorder_oracle_prefix_average.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.gorder_oracle_predict_hydrogens.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
        L.gorder_oracle_vector_to.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gorder_oracle_group_center.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.gorder_oracle_normal_from_cloud.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.gorder_oracle_shape_inside.argtypes = [C.POINTER(abi.CGorderSetup), C.c_void_p, C.c_void_p, C.c_void_p]
        L.gorder_oracle_shape_origin.argtypes = [C.POINTER(abi.CGorderSetup), C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# unit functions ------------------------------------------------------------------------------

def calc_sch(vec, normal) -> float:
    v, n = _f32(vec), _f32(normal)
    return float(lib().gorder_oracle_calc_sch(v.ctypes.data_as(C.POINTER(C.c_float)), n.ctypes.data_as(C.POINTER(C.c_float))))


def vector_to(p1, p2, box, pbc=True):
    out = np.zeros(3, np.float32)
    a, b, bx = _f32(p1), _f32(p2), _f32(box)
    lib().gorder_oracle_vector_to(_p(a), _p(b), _p(bx), int(pbc), _p(out))
    return out


def order_value(s: float) -> int:
    return int(lib().gorder_oracle_order_value(C.c_float(s)))


def calc_order(total: int, n: int, min_samples: int = 1) -> float:
    return float(lib().gorder_oracle_calc_order(int(total), int(n), int(min_samples)))


def predict_hydrogens(kind, target, h1, h2, h3=None, box=(0, 0, 0), pbc=True):
    out = np.zeros(9, np.float32)
    t, a, b, bx = _f32(target), _f32(h1), _f32(h2), _f32(box)
    c = _f32(h3) if h3 is not None else None
    n = lib().gorder_oracle_predict_hydrogens(int(kind), _p(t), _p(a), _p(b), _p(c) if c is not None else None, _p(bx), int(pbc), _p(out))
    return out.reshape(3, 3)[:n].copy()


def group_center(xyz, idx, box, pbc=True):
    out = np.zeros(3, np.float32)
    x, i, bx = _f32(xyz), np.ascontiguousarray(idx, dtype=np.int32), _f32(box)
    lib().gorder_oracle_group_center(_p(x), _p(i), int(i.size), _p(bx), int(pbc), _p(out))
    return out


def normal_from_cloud(points):
    out = np.zeros(3, np.float32)
    p = _f32(points).reshape(-1, 3)
    rc = lib().gorder_oracle_normal_from_cloud(_p(p), int(p.shape[0]), _p(out))
    if rc:
        raise abi.GorderError(rc, index=p.shape[0])
    return out


def shape_inside(setup: abi.EngineSetup, ref, point, box) -> bool:
    cs = setup.to_c()
    r, p, b = _f32(ref), _f32(point), _f32(box)
    return bool(lib().gorder_oracle_shape_inside(C.byref(cs), _p(r), _p(p), _p(b)))


def shape_origin(setup: abi.EngineSetup, ref, box):
    cs = setup.to_c()
    out = np.zeros(8, np.float32)
    r, b = _f32(ref), _f32(box)
    lib().gorder_oracle_shape_origin(C.byref(cs), _p(r), _p(b), _p(out))
    return out


def estimate_error(order, n_samples, n_blocks: int) -> Optional[float]:
    o = np.ascontiguousarray(order, dtype=np.int64)
    n = np.ascontiguousarray(n_samples, dtype=np.uint64)
    out = C.c_float(0)
    rc = lib().gorder_oracle_estimate_error(_p(o), _p(n), int(o.size), int(n_blocks), C.byref(out))
    return None if rc else float(out.value)


def prefix_average(order, n_samples):
    o = np.ascontiguousarray(order, dtype=np.int64)
    n = np.ascontiguousarray(n_samples, dtype=np.uint64)
    out = np.zeros(o.size, np.float32)
    lib().gorder_oracle_prefix_average(_p(o), _p(n), int(o.size), _p(out))
    return out


# engine ----------------------------------------------------------------------------------------

class Oracle:
    """CPU engine with the same call shape as ``gorder_b200.SystemTopology``."""

    def __init__(self, setup: abi.EngineSetup, n_threads: int = 1):
        self.setup = setup
        self.n_threads = int(n_threads)
        self._c = setup.to_c()
        self._h = C.c_void_p()
        rc = lib().gorder_oracle_create(C.byref(self._c), C.byref(self._h))
        if rc:
            raise abi.GorderError(rc)

    def analyze_frames(self, xyz, box, frame_index=None):
        xyz = _f32(xyz).reshape(-1, self.setup.n_atoms, 3)
        nf = xyz.shape[0]
        box = _f32(box).reshape(nf, 3)
        if frame_index is None:
            base = getattr(self, "_next", 0)
            frame_index = base + np.arange(nf, dtype=np.int64) * self.setup.step
        fi = np.ascontiguousarray(frame_index, dtype=np.int64)
        self._next = int(fi[-1]) + self.setup.step if nf else getattr(self, "_next", 0)
        rc = lib().gorder_oracle_analyze(self._h, _p(xyz), _p(box), _p(fi), nf, self.n_threads)
        if rc:
            raise abi.GorderError(rc, index=int(lib().gorder_oracle_error_detail(self._h)))

    def set_leaflets(self, table, frame_index=0):
        t = np.ascontiguousarray(table, dtype=np.uint8)
        lib().gorder_oracle_set_leaflets(self._h, _p(t), int(frame_index))

    def finish(self) -> abi.RawResults:
        return abi.fetch_results(lib(), self._h, "gorder_oracle", self.setup)

    def close(self):
        if self._h:
            lib().gorder_oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
