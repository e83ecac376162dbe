"""Loader of the C-ABI shared library.  Fails loudly: there is no Python / CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

from . import abi
from .build import SO_PATH

_LIB = None

#: every symbol ``include/gorder_b200.h`` declares
SYMBOLS = [
    "gorder_gpu_create", "gorder_gpu_submit", "gorder_gpu_submit_device", "gorder_gpu_native_layout",
    "gorder_gpu_submit_native", "gorder_gpu_submit_native_device", "gorder_gpu_reserve_frames", "gorder_gpu_set_leaflets", "gorder_gpu_sync",
    "gorder_gpu_result_sizes", "gorder_gpu_finish", "gorder_gpu_accumulator_block", "gorder_gpu_stats",
    "gorder_gpu_read_block", "gorder_gpu_write_block", "gorder_gpu_profile", "gorder_gpu_profile_read", "gorder_gpu_profile_read_normals",
    "gorder_gpu_speculation_stats", "gorder_gpu_fence", "gorder_gpu_wave_frames", "gorder_gpu_stream",
    "gorder_xtc_open", "gorder_xtc_info", "gorder_xtc_read", "gorder_xtc_write", "gorder_xtc_close", "gorder_xtc_scan", "gorder_gpu_run_xtc", "gorder_gpu_run_xtc_device",
    "gorder_results_order", "gorder_results_convergence", "gorder_results_map", "gorder_gpu_last_error", "gorder_gpu_error_detail", "gorder_gpu_destroy", "gorder_gpu_version",
    "gorder_topology_last_error", "gorder_system_from_tpr", "gorder_system_from_file", "gorder_system_from_arrays", "gorder_system_free", "gorder_system_n_atoms", "gorder_system_n_bonds",
    "gorder_system_tpx_version", "gorder_system_atoms", "gorder_system_bonds", "gorder_system_positions", "gorder_system_box", "gorder_system_set_bonds",
    "gorder_system_read_bonds", "gorder_classify_bonds", "gorder_classify_ua", "gorder_classification_free", "gorder_classification_n_types",
    "gorder_classification_moltypes", "gorder_classification_type_name", "gorder_classification_item_name", "gorder_classification_warning",
    "gorder_classification_n_atoms_rel", "gorder_classification_atoms_rel",
    "gorder_ndx_open", "gorder_ndx_close", "gorder_ndx_n_groups", "gorder_ndx_group_name", "gorder_ndx_group_size", "gorder_ndx_group_atoms", "gorder_ndx_find",
    "gorder_leaflets_from_ndx",
    "gorder_gpu_reduce", "gorder_comm_unique_id", "gorder_comm_create", "gorder_gpu_reduce_comm", "gorder_comm_broadcast_leaflets", "gorder_comm_destroy",
]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: build it with `python -m gorder_b200.build` "
            "(or __graft_entry__.build()).  gorder_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.gorder_gpu_create.argtypes = [C.POINTER(abi.CGorderSetup), C.POINTER(vp)]
    L.gorder_gpu_submit.argtypes = [vp, vp, vp, vp, i32]
    L.gorder_gpu_submit_device.argtypes = [vp, vp, vp, vp, i32]
    L.gorder_gpu_native_layout.argtypes = [vp, C.POINTER(i64), vp, vp]
    L.gorder_gpu_submit_native.argtypes = [vp, vp, vp, vp, i32]
    L.gorder_gpu_submit_native_device.argtypes = [vp, vp, vp, vp, i32]
    L.gorder_gpu_set_leaflets.argtypes = [vp, vp, i64]
    L.gorder_gpu_reserve_frames.argtypes = [vp, i64]
    L.gorder_gpu_reserve_frames.restype = C.c_int
    L.gorder_gpu_sync.argtypes = [vp]
    L.gorder_gpu_reduce.argtypes = [C.POINTER(vp), i32, i32]
    L.gorder_comm_unique_id.argtypes = [vp]
    L.gorder_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    L.gorder_gpu_reduce_comm.argtypes = [vp, vp, i32]
    L.gorder_comm_broadcast_leaflets.argtypes = [vp, vp, i32]
    L.gorder_comm_destroy.argtypes = [vp]
    L.gorder_comm_destroy.restype = None
    L.gorder_gpu_result_sizes.argtypes = [vp, C.POINTER(abi.CGorderResults)]
    L.gorder_gpu_finish.argtypes = [vp, C.POINTER(abi.CGorderResults)]
    L.gorder_gpu_accumulator_block.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.gorder_gpu_read_block.argtypes = [vp, vp]
    L.gorder_gpu_write_block.argtypes = [vp, vp]
    L.gorder_gpu_profile.argtypes = [vp, C.c_int]
    L.gorder_gpu_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    L.gorder_gpu_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.gorder_gpu_speculation_stats.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
    L.gorder_gpu_speculation_stats.restype = C.c_int
    L.gorder_xtc_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.gorder_xtc_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(C.c_float)]
    L.gorder_xtc_read.argtypes = [vp, i64, i64, i64, i32, vp, vp, vp, vp]
    L.gorder_xtc_write.argtypes = [C.c_char_p, vp, vp, i32, i64, C.c_float, i32, i32, C.c_float, i32]
    L.gorder_xtc_close.argtypes = [vp]
    L.gorder_xtc_scan.argtypes = [vp, i64, i64, vp, vp]
    L.gorder_xtc_scan.restype = C.c_int
    L.gorder_xtc_close.restype = None
    L.gorder_gpu_run_xtc.argtypes = [vp, vp, vp, i64, i64, i64, i64, i32, i32, C.POINTER(C.c_double)]
    L.gorder_gpu_run_xtc_device.argtypes = [vp, vp, vp, i64, i64, i64, i64, i32, i32, C.POINTER(i64)]
    for name in ("gorder_xtc_open", "gorder_xtc_info", "gorder_xtc_read", "gorder_xtc_write", "gorder_gpu_run_xtc", "gorder_gpu_run_xtc_device"):
        getattr(L, name).restype = C.c_int
    L.gorder_results_order.argtypes = [C.POINTER(abi.CGorderRaw), vp, i32, i32, i32, C.c_float, vp, vp]
    L.gorder_results_convergence.argtypes = [C.POINTER(abi.CGorderRaw), vp, i32, C.c_float, vp]
    L.gorder_results_map.argtypes = [vp, vp, i64, i32, C.c_float, vp]
    for name in ("gorder_results_order", "gorder_results_convergence", "gorder_results_map"):
        getattr(L, name).restype = C.c_int
    # structure / topology / classification (host only)
    L.gorder_topology_last_error.restype = C.c_char_p
    L.gorder_system_from_tpr.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.gorder_system_from_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.gorder_system_from_file.restype = C.c_int
    L.gorder_system_from_arrays.argtypes = [i32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), vp, vp, vp, C.POINTER(vp)]
    L.gorder_system_free.argtypes = [vp]
    L.gorder_system_free.restype = None
    L.gorder_system_n_atoms.argtypes = [vp]
    L.gorder_system_n_atoms.restype = i32
    L.gorder_system_n_bonds.argtypes = [vp]
    L.gorder_system_n_bonds.restype = i64
    L.gorder_system_tpx_version.argtypes = [vp]
    L.gorder_system_tpx_version.restype = i32
    L.gorder_system_atoms.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.gorder_system_bonds.argtypes = [vp, vp]
    L.gorder_system_positions.argtypes = [vp, vp, C.POINTER(i32)]
    L.gorder_system_box.argtypes = [vp, vp, C.POINTER(i32)]
    L.gorder_system_set_bonds.argtypes = [vp, vp, i64]
    L.gorder_system_read_bonds.argtypes = [vp, C.c_char_p]
    L.gorder_classify_bonds.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, C.POINTER(vp)]
    L.gorder_classify_ua.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, C.POINTER(vp)]
    L.gorder_classification_free.argtypes = [vp]
    L.gorder_classification_free.restype = None
    L.gorder_classification_n_types.argtypes = [vp]
    L.gorder_classification_n_types.restype = i32
    L.gorder_classification_moltypes.argtypes = [vp]
    L.gorder_classification_moltypes.restype = vp
    L.gorder_classification_type_name.argtypes = [vp, i32]
    L.gorder_classification_type_name.restype = C.c_char_p
    L.gorder_classification_item_name.argtypes = [vp, i32, i32]
    L.gorder_classification_item_name.restype = C.c_char_p
    L.gorder_classification_warning.argtypes = [vp]
    L.gorder_classification_warning.restype = C.c_char_p
    L.gorder_classification_n_atoms_rel.argtypes = [vp, i32]
    L.gorder_classification_n_atoms_rel.restype = i32
    L.gorder_classification_atoms_rel.argtypes = [vp, i32]
    L.gorder_classification_atoms_rel.restype = C.POINTER(i32)
    for name in ("gorder_system_from_tpr", "gorder_system_from_file", "gorder_system_from_arrays", "gorder_system_atoms", "gorder_system_bonds", "gorder_system_positions",
                 "gorder_system_box", "gorder_system_set_bonds", "gorder_system_read_bonds", "gorder_classify_bonds", "gorder_classify_ua"):
        getattr(L, name).restype = C.c_int
    L.gorder_ndx_open.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
    L.gorder_ndx_open.restype = C.c_int
    L.gorder_ndx_close.argtypes = [vp]
    L.gorder_ndx_close.restype = None
    L.gorder_ndx_n_groups.argtypes = [vp]
    L.gorder_ndx_n_groups.restype = i32
    L.gorder_ndx_group_name.argtypes = [vp, i32]
    L.gorder_ndx_group_name.restype = C.c_char_p
    L.gorder_ndx_group_size.argtypes = [vp, i32]
    L.gorder_ndx_group_size.restype = i64
    L.gorder_ndx_group_atoms.argtypes = [vp, i32]
    L.gorder_ndx_group_atoms.restype = C.POINTER(i32)
    L.gorder_ndx_find.argtypes = [vp, C.c_char_p]
    L.gorder_ndx_find.restype = i32
    L.gorder_leaflets_from_ndx.argtypes = [C.POINTER(C.c_char_p), i32, i32, C.c_char_p, C.c_char_p, vp, i32, vp]
    L.gorder_leaflets_from_ndx.restype = C.c_int
    L.gorder_gpu_fence.argtypes = [vp]
    L.gorder_gpu_fence.restype = C.c_int
    L.gorder_gpu_wave_frames.argtypes = [vp, C.POINTER(i32)]
    L.gorder_gpu_wave_frames.restype = C.c_int
    L.gorder_gpu_stream.argtypes = [vp]
    L.gorder_gpu_stream.restype = vp
    L.gorder_gpu_last_error.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.gorder_gpu_error_detail.argtypes = [vp]
    L.gorder_gpu_error_detail.restype = i64
    L.gorder_gpu_destroy.argtypes = [vp]
    L.gorder_gpu_destroy.restype = None
    L.gorder_gpu_version.restype = C.c_char_p
    for name in ("gorder_gpu_create", "gorder_gpu_submit", "gorder_gpu_submit_device", "gorder_gpu_native_layout",
                 "gorder_gpu_submit_native", "gorder_gpu_submit_native_device", "gorder_gpu_set_leaflets",
                 "gorder_gpu_sync", "gorder_gpu_result_sizes", "gorder_gpu_finish", "gorder_gpu_accumulator_block",
                 "gorder_gpu_stats", "gorder_gpu_last_error", "gorder_gpu_read_block", "gorder_gpu_write_block",
                 "gorder_gpu_profile", "gorder_gpu_profile_read"):
        getattr(L, name).restype = C.c_int
    _LIB = L
    return L
