"""Host-side mirror of the reference's per-worker analysis state.

``SystemTopology`` here plays the role of ``gorder``'s ``SystemTopology``
(``src/analysis/topology/mod.rs:35-65``): it is built once from the classified molecule types
(``SystemTopology::new``, :70-118), is fed frames (``analyze_frame``, ``src/analysis/common.rs:201-235``)
and is reduced into raw accumulators (``ParallelTrajData::reduce``, ``topology/mod.rs:256-272``).
All the work happens in the CUDA library behind the C ABI of ``include/gorder_b200.h``; this class
only marshals numpy / torch buffers into plain pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import abi
from ._lib import lib


def _ptr(a) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


class SystemTopology:
    """One engine instance on one GPU (the reference clones one ``SystemTopology`` per thread)."""

    def __init__(self, setup: abi.EngineSetup):
        self.setup = setup
        self._c = setup.to_c()
        self._h = C.c_void_p()
        rc = lib().gorder_gpu_create(C.byref(self._c), C.byref(self._h))
        if rc != abi.OK:
            self._h = C.c_void_p()
            raise abi.GorderError(rc)
        self._next_frame = 0

    # -- frames ---------------------------------------------------------------------------------
    def _frame_index(self, n: int, frame_index) -> np.ndarray:
        if frame_index is None:
            fi = self._next_frame + np.arange(n, dtype=np.int64) * self.setup.step
        else:
            fi = np.ascontiguousarray(frame_index, dtype=np.int64)
            if fi.size != n:
                raise ValueError("frame_index must have one entry per frame")
        if n:
            self._next_frame = int(fi[-1]) + self.setup.step
        return fi

    def _check(self, rc: int):
        if rc != abi.OK:
            buf = C.create_string_buffer(512)
            lib().gorder_gpu_last_error(self._h, buf, 512)
            raise abi.GorderError(rc, buf.value.decode(errors="replace"), int(lib().gorder_gpu_error_detail(self._h)))

    def analyze_frames(self, xyz, box, frame_index=None):
        """``analyze_frame`` for a batch: ``xyz`` [F][n_atoms][3] f32 host array, ``box`` [F][3]."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, self.setup.n_atoms, 3)
        n = xyz.shape[0]
        box = np.ascontiguousarray(box, dtype=np.float32).reshape(n, 3) if box is not None else None
        fi = self._frame_index(n, frame_index)
        self._check(lib().gorder_gpu_submit(self._h, _ptr(xyz), _ptr(box), _ptr(fi), n))

    def analyze_frames_native(self, planes, box, frame_index=None):
        """Frames already in the native plane layout (host memory)."""
        planes = np.ascontiguousarray(planes, dtype=np.float32).reshape(-1, self.frame_floats)
        n = planes.shape[0]
        box = np.ascontiguousarray(box, dtype=np.float32).reshape(n, 3) if box is not None else None
        fi = self._frame_index(n, frame_index)
        self._check(lib().gorder_gpu_submit_native(self._h, _ptr(planes), _ptr(box), _ptr(fi), n))

    def analyze_frames_device(self, d_ptr: int, d_box: int, n_frames: int, frame_index=None, native: bool = False):
        """Frames resident in device memory (raw device pointers, e.g. ``tensor.data_ptr()``)."""
        fi = self._frame_index(n_frames, frame_index)
        fn = lib().gorder_gpu_submit_native_device if native else lib().gorder_gpu_submit_device
        self._check(fn(self._h, C.c_void_p(d_ptr), C.c_void_p(d_box), _ptr(fi), n_frames))

    def wave_frames(self) -> int:
        """Frames per batch that make the accumulation kernel's grid a whole number of waves (0: no preference)."""
        n = C.c_int32(0)
        self._check(lib().gorder_gpu_wave_frames(self._h, C.byref(n)))
        return int(n.value)

    def run_xtc(self, xtc, atom_of_slot=None, first: int = 0, last: int | None = None, stride: int = 1, n_threads: int = 0,
                batch_frames: int = 0, frame_index0: int = 0) -> float:
        """``read_trajectory`` for an open :class:`gorder_b200.xtc.XtcFile`: host threads decode, the engine analyses.
        Returns the host seconds spent decoding (summed over threads)."""
        import os
        m = None if atom_of_slot is None else np.ascontiguousarray(atom_of_slot, dtype=np.int32)
        if m is not None and m.size != self.setup.n_atoms:
            raise ValueError("atom_of_slot must have one entry per engine atom")
        sec = C.c_double(0)
        self._check(lib().gorder_gpu_run_xtc(self._h, xtc._x, _ptr(m), first, xtc.n_frames if last is None else last, stride, frame_index0,
                                             n_threads or (os.cpu_count() or 1), batch_frames, C.byref(sec)))
        return float(sec.value)

    def run_xtc_device(self, xtc, atom_of_slot=None, first: int = 0, last: int | None = None, stride: int = 1, n_threads: int = 0,
                       batch_frames: int = 0, frame_index0: int = 0) -> int:
        """``run_xtc`` with the XTC decode on the device; returns the bytes that crossed PCIe."""
        import os
        m = None if atom_of_slot is None else np.ascontiguousarray(atom_of_slot, dtype=np.int32)
        if m is not None and m.size != self.setup.n_atoms:
            raise ValueError("atom_of_slot must have one entry per engine atom")
        moved = C.c_int64(0)
        self._check(lib().gorder_gpu_run_xtc_device(self._h, xtc._x, _ptr(m), first, xtc.n_frames if last is None else last, stride, frame_index0,
                                                    n_threads or (os.cpu_count() or 1), batch_frames, C.byref(moved)))
        return int(moved.value)

    # -- layout -----------------------------------------------------------------------------------
    @property
    def frame_floats(self) -> int:
        n = C.c_int64(0)
        lib().gorder_gpu_native_layout(self._h, C.byref(n), None, None)
        return int(n.value)

    def native_layout(self):
        """(frame_floats, plane_offset[n_atoms], plane_cstride[n_atoms]); offset -1 = atom not needed."""
        n = C.c_int64(0)
        off = np.zeros(self.setup.n_atoms, np.int32)
        cs = np.zeros(self.setup.n_atoms, np.int32)
        lib().gorder_gpu_native_layout(self._h, C.byref(n), _ptr(off), _ptr(cs))
        return int(n.value), off, cs

    def to_native(self, xyz) -> np.ndarray:
        """Host-side gather of AoS frames into the native layout (what the Rust shim does while it
        copies the Master group into its pinned batch buffer)."""
        xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, self.setup.n_atoms, 3)
        ff, off, cs = self.native_layout()
        out = np.zeros((xyz.shape[0], ff), np.float32)
        sel = np.nonzero(off >= 0)[0]
        for c in range(3):
            out[:, off[sel] + c * cs[sel]] = xyz[:, sel, c]
        return out

    # -- leaflets / reduce ---------------------------------------------------------------------------
    def reserve_frames(self, n_frames: int):
        """Pre-allocate the per-frame rows for ``n_frames`` analysed frames (error estimation)."""
        self._check(lib().gorder_gpu_reserve_frames(self._h, int(n_frames)))

    def set_leaflets(self, table, frame_index: int = 0):
        t = np.ascontiguousarray(table, dtype=np.uint8)
        self._check(lib().gorder_gpu_set_leaflets(self._h, _ptr(t), int(frame_index)))

    def sync(self):
        self._check(lib().gorder_gpu_sync(self._h))

    def accumulator_block(self):
        """(device pointer, n int64 words) of the contiguous accumulator block (for the NCCL reduce)."""
        p, n = C.c_void_p(), C.c_int64(0)
        self._check(lib().gorder_gpu_accumulator_block(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def read_block(self, d_dst: int):
        self._check(lib().gorder_gpu_read_block(self._h, C.c_void_p(d_dst)))

    def write_block(self, d_src: int):
        self._check(lib().gorder_gpu_write_block(self._h, C.c_void_p(d_src)))

    def reduce_comm(self, comm: "Comm", root: int = 0):
        """``ParallelTrajData::reduce`` across ranks (one process per GPU): collective ``gorder_gpu_reduce_comm``."""
        self._check(lib().gorder_gpu_reduce_comm(self._h, comm._c, int(root)))

    def broadcast_leaflets(self, comm: "Comm", root: int = 0):
        """``Frequency::Once``: the table of analysed frame 0 (rank ``root``) reaches every shard (collective)."""
        self._check(lib().gorder_comm_broadcast_leaflets(self._h, comm._c, int(root)))

    def profile(self, enable: bool = True):
        self._check(lib().gorder_gpu_profile(self._h, int(enable)))

    def profile_read(self):
        """(summed ms, launches) of the accumulation kernel since the last read (CUDA events)."""
        ms, n = C.c_double(0), C.c_int64(0)
        self._check(lib().gorder_gpu_profile_read(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def profile_read_normals(self):
        """(summed ms, batches) of the membrane-normal stage (cell list + PCA kernels) since the last read."""
        ms, n = C.c_double(0), C.c_int64(0)
        self._check(lib().gorder_gpu_profile_read_normals(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def stats(self):
        k, f = C.c_int64(0), C.c_int64(0)
        lib().gorder_gpu_stats(self._h, C.byref(k), C.byref(f))
        return {"kernel_launches": int(k.value), "frames": int(f.value)}

    def fence(self):
        """Main stream waits for the helper streams (asynchronous); record closing timing events after it."""
        self._check(lib().gorder_gpu_fence(self._h))

    def speculation_stats(self):
        """Frames classified without a centre pre-pass / frames that needed the exact centre afterwards."""
        e, a, b = C.c_int32(0), C.c_int64(0), C.c_int64(0)
        lib().gorder_gpu_speculation_stats(self._h, C.byref(e), C.byref(a), C.byref(b))
        return {"enabled": bool(e.value), "frames_speculated": int(a.value), "frames_repaired": int(b.value)}

    @property
    def stream(self) -> int:
        return int(lib().gorder_gpu_stream(self._h) or 0)

    def finish(self, totals_only: bool = False) -> abi.RawResults:
        """``ParallelTrajData::reduce`` for one GPU: fetch the raw accumulators (``totals_only``: without per-frame rows, maps,
        leaflet tables and normals)."""
        try:
            return abi.fetch_results(lib(), self._h, "gorder_gpu", self.setup, totals_only)
        except abi.GorderError as e:
            buf = C.create_string_buffer(512)
            lib().gorder_gpu_last_error(self._h, buf, 512)
            raise abi.GorderError(e.code, buf.value.decode(errors="replace"), int(lib().gorder_gpu_error_detail(self._h)))

    def close(self):
        if self._h:
            lib().gorder_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reduce_handles(engines, root: int = 0):
    """``ParallelTrajData::reduce`` for one process that drives several GPUs (``gorder_gpu_reduce``): afterwards
    ``engines[root].finish()`` returns the merged result."""
    arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
    rc = lib().gorder_gpu_reduce(arr, len(engines), int(root))
    if rc != abi.OK:
        engines[root]._check(rc)


class Comm:
    """The library's own NCCL communicator (``gorder_comm_*``): rank 0 makes the id, the host broadcasts its 128 bytes."""

    ID_BYTES = 128

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(Comm.ID_BYTES)
        rc = lib().gorder_comm_unique_id(buf)
        if rc != abi.OK:
            raise abi.GorderError(rc)
        return buf.raw

    def __init__(self, uid: bytes, n_ranks: int, rank: int, device: int):
        self._c = C.c_void_p()
        buf = C.create_string_buffer(bytes(uid), Comm.ID_BYTES)
        rc = lib().gorder_comm_create(buf, int(n_ranks), int(rank), int(device), C.byref(self._c))
        if rc != abi.OK:
            self._c = C.c_void_p()
            raise abi.GorderError(rc)
        self.n_ranks, self.rank = n_ranks, rank

    def close(self):
        if self._c:
            lib().gorder_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
