"""Loader for the in-tree native library (platform C ABI + host driver).

One shared object per value type, as in PanguLU where the value type is a compile-time choice:
``pangulu_b200/lib/libpangulu_sm100_<r64|r32|cr64|cr32>.so`` built by ``pangulu_b200/csrc/Makefile``
(``__graft_entry__.build()``).  There is no Python or CPU fallback: a missing library is an error.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (PANGULU_SM100_LIB_DIR: an instrumented build of the same library, e.g. lib_prof built with EXTRA=-DPSM_CT_PROFILE)
LIB_DIR = os.environ.get("PANGULU_SM100_LIB_DIR") or os.path.join(_HERE, "lib")
PLATFORM_ID = "0201001"

VALUE_DTYPES = {"r64": np.float64, "r32": np.float32, "cr64": np.complex128, "cr32": np.complex64}

#: the 21 functions of the PanguLU platform interface (reference build_helper.py:8-32)
PLATFORM_FUNCTIONS = [
    "malloc", "malloc_pinned", "synchronize", "memset", "create_stream", "memcpy", "memcpy_async", "free",
    "get_device_num", "set_default_device", "get_device_name", "get_device_memory_usage",
    "getrf", "tstrf", "gessm", "ssssm", "ssssm_batched", "hybrid_batched", "spmv", "vecadd", "sptrsv",
]
#: additive entry points declared in include/pangulu_platform_0201001.h
EXTENSION_FUNCTIONS = [
    "pangulu_sm100_set_option", "pangulu_sm100_get_option", "pangulu_sm100_free_pinned",
    "pangulu_sm100_download_blocks", "pangulu_sm100_upload_blocks",
    "pangulu_sm100_plan_create", "pangulu_sm100_plan_execute", "pangulu_sm100_plan_execute_range",
    "pangulu_sm100_plan_destroy", "pangulu_sm100_stats_reset", "pangulu_sm100_stats_launches",
    "pangulu_sm100_stats_ms", "pangulu_sm100_stream", "pangulu_sm100_version", "pangulu_sm100_plan_trace",
    "pangulu_sm100_plan_sync_base", "pangulu_sm100_plan_flag_offset", "pangulu_sm100_plan_expect_remote",
    "pangulu_sm100_plan_add_push", "pangulu_sm100_plan_commit", "pangulu_sm100_plan_reset", "pangulu_sm100_plan_launch",
    "pangulu_sm100_ipc_export", "pangulu_sm100_ipc_open", "pangulu_sm100_plan_dense_base", "pangulu_sm100_plan_dense_offset",
    "pangulu_sm100_plan_add_push2", "pangulu_sm100_set_diag_groups", "pangulu_sm100_plan_stat",
    "pangulu_sm100_plan_flag_offset2", "pangulu_sm100_plan_expect_remote2", "pangulu_sm100_plan_add_push3",
    "pangulu_sm100_ipc_close", "pangulu_sm100_plan_task_work",
    "pangulu_sm100_solve_plan_create", "pangulu_sm100_solve_plan_execute", "pangulu_sm100_solve_plan_destroy",
    "pangulu_sm100_scatter_values", "pangulu_sm100_mirror_patterns", "pangulu_sm100_plan_execute_download",
]
#: host driver, include/pangulu_sm100_driver.h
DRIVER_FUNCTIONS = [
    "pangulu_init", "pangulu_gstrf", "pangulu_gstrs", "pangulu_gssv", "pangulu_finalize",
    "pangulu_sm100_driver_config", "pangulu_sm100_driver_info", "pangulu_sm100_driver_slot",
    "pangulu_sm100_driver_perm", "pangulu_sm100_driver_tasks", "pangulu_sm100_driver_reset_values",
    "pangulu_sm100_driver_download", "pangulu_sm100_driver_factor_check", "pangulu_sm100_driver_gstrf_enqueue",
    "pangulu_sm100_driver_trace", "pangulu_sm100_driver_set_barrier", "pangulu_sm100_driver_ipc_handles",
    "pangulu_sm100_driver_mirror_table", "pangulu_sm100_driver_connect_peer", "pangulu_sm100_driver_commit",
    "pangulu_sm100_driver_gstrf_reset", "pangulu_sm100_driver_gstrf_launch", "pangulu_sm100_driver_set_mode",
    "pangulu_sm100_driver_rebuild_plan", "pangulu_sm100_driver_diag_groups", "pangulu_sm100_driver_disconnect", "pangulu_sm100_driver_task_work",
    "pangulu_sm100_driver_save_factors", "pangulu_sm100_driver_load_factors",
    "pangulu_sm100_reference_gstrf", "pangulu_sm100_reference_gstrf_time", "pangulu_sm100_reference_release",
    "pangulu_sm100_driver_check_ux", "pangulu_sm100_driver_check_lux", "pangulu_sm100_driver_check_ax",
    "pangulu_sm100_driver_update_values", "pangulu_sm100_driver_device_mirrors",
]


class Slot(C.Structure):
    """pangulu_storage_slot_t (include/pangulu_sm100_abi.h; reference src/pangulu_common.h:207-231, GPU build)."""

    _fields_ = [
        ("brow_pos", C.c_uint32), ("bcol_pos", C.c_uint32),
        ("columnpointer", C.c_void_p), ("rowindex", C.c_void_p), ("value", C.c_void_p),
        ("rowpointer", C.c_void_p), ("columnindex", C.c_void_p), ("idx_of_csc_value_for_csr", C.c_void_p),
        ("data_status", C.c_char), ("related_block", C.c_void_p),
        ("is_upper", C.c_int32), ("bin_id", C.c_int32), ("slot_idx", C.c_int32), ("task_queue", C.c_void_p),
        ("d_columnpointer", C.c_void_p), ("d_rowindex", C.c_void_p), ("d_value", C.c_void_p),
        ("d_rowpointer", C.c_void_p), ("d_columnindex", C.c_void_p), ("d_idx_of_csc_value_for_csr", C.c_void_p),
    ]


class Task(C.Structure):
    """pangulu_task_t (reference src/pangulu_common.h:233-243)."""

    _fields_ = [
        ("row", C.c_uint32), ("col", C.c_uint32), ("kernel_id", C.c_int16), ("task_level", C.c_uint32),
        ("compare_flag", C.c_int64), ("opdst", C.c_void_p), ("op1", C.c_void_p), ("op2", C.c_void_p),
    ]


class InitOptions(C.Structure):
    """pangulu_init_options (reference include/pangulu_interface_common.h:3-12)."""

    _fields_ = [
        ("nthread", C.c_int), ("nb", C.c_int), ("gpu_kernel_warp_per_block", C.c_int),
        ("gpu_data_move_warp_per_block", C.c_int), ("sizeof_value", C.c_int), ("is_complex_matrix", C.c_int),
        ("mpi_recv_buffer_level", C.c_float),
    ]


assert C.sizeof(Slot) == 144 and C.sizeof(Task) == 48

TASK_GETRF, TASK_TSTRF, TASK_GESSM, TASK_SSSSM = 1, 2, 3, 4


def lib_path(vtype: str = "r64") -> str:
    return os.path.join(LIB_DIR, f"libpangulu_sm100_{vtype}.so")


_cache: dict[str, C.CDLL] = {}


def load(vtype: str = "r64") -> C.CDLL:
    """Load (once) the native library of one value type and declare its prototypes."""
    if vtype in _cache:
        return _cache[vtype]
    if vtype not in VALUE_DTYPES:
        raise ValueError(f"unknown value type {vtype!r}")
    path = lib_path(vtype)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the sm_100a extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C pangulu_b200/csrc). There is no CPU fallback for the numeric factorisation.")
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    P = lambda fn: getattr(L, f"pangulu_platform_{PLATFORM_ID}_{fn}")
    vp, sz, u64, i64 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int64
    sig = {
        "malloc": [C.POINTER(vp), sz], "malloc_pinned": [C.POINTER(vp), sz], "synchronize": [], "memset": [vp, C.c_int, sz],
        "create_stream": [C.POINTER(vp)], "memcpy": [vp, vp, sz, C.c_uint], "memcpy_async": [vp, vp, sz, C.c_uint, vp],
        "free": [vp], "get_device_num": [C.POINTER(C.c_int)], "set_default_device": [C.c_int],
        "get_device_name": [C.c_char_p, C.c_int], "get_device_memory_usage": [C.POINTER(sz)],
        "getrf": [C.c_uint16, vp, C.c_int], "tstrf": [C.c_uint16, vp, vp, C.c_int], "gessm": [C.c_uint16, vp, vp, C.c_int],
        "ssssm": [C.c_uint16, vp, vp, vp, C.c_int], "ssssm_batched": [C.c_uint16, u64, vp], "hybrid_batched": [C.c_uint16, u64, vp],
        "spmv": [C.c_uint16, vp, vp, vp], "vecadd": [i64, vp, vp], "sptrsv": [C.c_uint16, vp, vp, i64],
    }
    for fn, args in sig.items():
        f = P(fn)
        f.argtypes = args
        f.restype = None
    L.pangulu_sm100_set_option.argtypes = [C.c_char_p, C.c_longlong]
    L.pangulu_sm100_set_option.restype = C.c_int
    L.pangulu_sm100_get_option.argtypes = [C.c_char_p]
    L.pangulu_sm100_get_option.restype = C.c_longlong
    L.pangulu_sm100_free_pinned.argtypes = [vp]
    L.pangulu_sm100_download_blocks.argtypes = [C.c_uint16, u64, vp]
    L.pangulu_sm100_upload_blocks.argtypes = [C.c_uint16, u64, vp]
    L.pangulu_sm100_plan_create.argtypes = [C.c_uint16, u64, vp, vp]
    L.pangulu_sm100_plan_create.restype = vp
    L.pangulu_sm100_plan_execute.argtypes = [vp]
    L.pangulu_sm100_plan_execute_range.argtypes = [vp, u64, u64]
    L.pangulu_sm100_plan_destroy.argtypes = [vp]
    L.pangulu_sm100_stats_launches.argtypes = [C.c_int]
    L.pangulu_sm100_stats_launches.restype = C.c_longlong
    L.pangulu_sm100_stats_ms.argtypes = [C.c_int]
    L.pangulu_sm100_stats_ms.restype = C.c_double
    L.pangulu_sm100_stream.restype = vp
    L.pangulu_sm100_version.restype = C.c_char_p
    # driver
    L.pangulu_init.argtypes = [C.c_uint32, u64, vp, vp, vp, C.POINTER(InitOptions), C.POINTER(vp)]
    L.pangulu_init.restype = None
    L.pangulu_gstrf.argtypes = [vp, C.POINTER(vp)]
    L.pangulu_gstrs.argtypes = [vp, vp, C.POINTER(vp)]
    L.pangulu_gssv.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.pangulu_finalize.argtypes = [C.POINTER(vp)]
    for f in (L.pangulu_gstrf, L.pangulu_gstrs, L.pangulu_gssv, L.pangulu_finalize):
        f.restype = None
    L.pangulu_sm100_driver_config.argtypes = [C.c_char_p, C.c_longlong]
    L.pangulu_sm100_driver_config.restype = C.c_int
    L.pangulu_sm100_driver_info.argtypes = [vp, C.c_char_p]
    L.pangulu_sm100_driver_info.restype = C.c_double
    L.pangulu_sm100_driver_slot.argtypes = [vp, C.c_longlong]
    L.pangulu_sm100_driver_slot.restype = C.POINTER(Slot)
    L.pangulu_sm100_driver_perm.argtypes = [vp, C.c_int]
    L.pangulu_sm100_driver_perm.restype = vp
    L.pangulu_sm100_driver_tasks.argtypes = [vp, C.POINTER(vp)]
    L.pangulu_sm100_driver_tasks.restype = vp
    L.pangulu_sm100_driver_reset_values.argtypes = [vp, C.c_int]
    L.pangulu_sm100_driver_download.argtypes = [vp]
    L.pangulu_sm100_driver_gstrf_enqueue.argtypes = [vp]
    L.pangulu_sm100_driver_set_barrier.argtypes = [vp]
    L.pangulu_sm100_driver_ipc_handles.argtypes = [vp, vp]
    L.pangulu_sm100_driver_mirror_table.argtypes = [vp, vp, C.c_longlong]
    L.pangulu_sm100_driver_mirror_table.restype = C.c_longlong
    L.pangulu_sm100_driver_connect_peer.argtypes = [vp, vp, vp, C.c_longlong]
    L.pangulu_sm100_driver_connect_peer.restype = C.c_longlong
    L.pangulu_sm100_driver_commit.argtypes = [vp]
    L.pangulu_sm100_driver_set_mode.argtypes = [vp, C.c_int]
    L.pangulu_sm100_driver_rebuild_plan.argtypes = [vp]
    L.pangulu_sm100_driver_diag_groups.argtypes = [vp, C.c_longlong, vp]
    L.pangulu_sm100_driver_diag_groups.restype = C.c_int
    L.pangulu_sm100_driver_gstrf_reset.argtypes = [vp]
    L.pangulu_sm100_driver_gstrf_launch.argtypes = [vp]
    L.pangulu_sm100_driver_trace.argtypes = [vp, vp, C.c_longlong]
    L.pangulu_sm100_driver_trace.restype = C.c_longlong
    L.pangulu_sm100_plan_trace.argtypes = [vp, vp, C.c_longlong]
    L.pangulu_sm100_plan_trace.restype = C.c_longlong
    L.pangulu_sm100_driver_factor_check.argtypes = [vp]
    L.pangulu_sm100_scatter_values.argtypes = [vp, C.c_size_t, vp, vp, vp, u64]
    L.pangulu_sm100_mirror_patterns.argtypes = [C.c_uint16, u64, vp]
    L.pangulu_sm100_plan_execute_download.argtypes = [vp, u64, vp]
    L.pangulu_sm100_driver_update_values.argtypes = [vp, vp, vp, vp]
    L.pangulu_sm100_driver_update_values.restype = C.c_double
    L.pangulu_sm100_driver_device_mirrors.argtypes = [vp, C.c_int]
    L.pangulu_sm100_driver_device_mirrors.restype = C.c_longlong
    L.pangulu_sm100_driver_factor_check.restype = C.c_double
    L.pangulu_sm100_plan_task_work.argtypes = [vp, vp, C.c_longlong]
    L.pangulu_sm100_plan_task_work.restype = C.c_longlong
    L.pangulu_sm100_driver_task_work.argtypes = [vp, vp, C.c_longlong]
    L.pangulu_sm100_driver_task_work.restype = C.c_longlong
    for fn in (L.pangulu_sm100_driver_save_factors, L.pangulu_sm100_driver_load_factors):
        fn.argtypes = [vp, C.c_char_p]
        fn.restype = C.c_longlong
    L.pangulu_sm100_driver_disconnect.argtypes = [vp]
    L.pangulu_sm100_driver_check_ux.argtypes = [vp, vp]
    L.pangulu_sm100_driver_check_lux.argtypes = [vp, vp, vp]
    L.pangulu_sm100_driver_check_ax.argtypes = [vp, vp]
    L.pangulu_sm100_driver_check_ax.restype = C.c_longlong
    L.platform = P
    _cache[vtype] = L
    return L
