"""ctypes mirror of the host-side trajectory feed (``gorder_xtc_*`` in include/gorder_b200.h).

The product's stand-in for the reference's reader (read_trajectory -> groan_rs GroupXtcReader -> molly,
src/analysis/common.rs:281-304): open / decode / write GROMACS XTC files and run a whole trajectory through the
engine with ``SystemTopology.run_xtc``.  Reading and writing need no GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class XtcFile:
    """An open, memory-mapped, frame-indexed XTC trajectory."""

    def __init__(self, path: str):
        self._x = C.c_void_p()
        rc = lib().gorder_xtc_open(os.fsencode(path), C.byref(self._x))
        if rc:
            raise OSError(f"gorder_xtc_open({path!r}) failed with code {rc}")
        na, nf, pr = C.c_int32(0), C.c_int64(0), C.c_float(0)
        lib().gorder_xtc_info(self._x, C.byref(na), C.byref(nf), C.byref(pr))
        self.path, self.n_atoms, self.n_frames, self.precision = path, int(na.value), int(nf.value), float(pr.value)

    def read(self, first: int = 0, count: int | None = None, stride: int = 1, n_threads: int = 0):
        """(xyz [count][n_atoms][3] f32, box9 [count][9], time [count], step [count])."""
        if count is None:
            count = max(0, (self.n_frames - first + stride - 1) // stride)
        xyz = np.empty((count, self.n_atoms, 3), np.float32)
        box9 = np.empty((count, 9), np.float32)
        time = np.empty(count, np.float32)
        step = np.empty(count, np.int32)
        rc = lib().gorder_xtc_read(self._x, first, count, stride, n_threads or (os.cpu_count() or 1), _ptr(xyz), _ptr(box9), _ptr(time), _ptr(step))
        if rc:
            raise OSError(f"gorder_xtc_read failed with code {rc}")
        return xyz, box9, time, step

    def scan(self, first: int = 0, count: int | None = None):
        """(groups [count], bookmarks [count]) of frames first .. first + count - 1 as the device-decode path's host stage sees
        them (``gorder_xtc_scan``); groups is -2 for a frame that path hands to the host decoder.  Raises OSError when a
        stream is inconsistent."""
        if count is None:
            count = self.n_frames - first
        ng, nb = np.zeros(max(count, 0), np.int32), np.zeros(max(count, 0), np.int32)
        rc = lib().gorder_xtc_scan(self._x, first, count, _ptr(ng), _ptr(nb))
        if rc:
            raise OSError(f"gorder_xtc_scan: corrupt frame(s) {np.flatnonzero(ng == -1) + first} (code {rc})")
        return ng, nb

    def close(self):
        if self._x:
            lib().gorder_xtc_close(self._x)
            self._x = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_xtc(path: str, xyz, box, precision: float = 1000.0, append: bool = False, first_step: int = 0, dt: float = 1.0, n_threads: int = 0):
    """Write frames (orthogonal boxes) with xdrfile's compression: xyz [F][n_atoms][3], box [F][3]."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    box = np.ascontiguousarray(box, np.float32).reshape(xyz.shape[0], 3)
    rc = lib().gorder_xtc_write(os.fsencode(path), _ptr(xyz), _ptr(box), xyz.shape[1], xyz.shape[0], float(precision), int(append), int(first_step),
                                float(dt), n_threads or (os.cpu_count() or 1))
    if rc:
        raise OSError(f"gorder_xtc_write({path!r}) failed with code {rc}")
