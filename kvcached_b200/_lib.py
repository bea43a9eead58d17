"""ctypes binding of libkvb200.so (include/kvb200.h).

The library is the product: if it is missing and cannot be built, importing this module raises —
there is no Python or CPU fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkvb200.so")

i32, i64, u64, f64 = C.c_int32, C.c_int64, C.c_uint64, C.c_double
p_i64, p_u64, p_f64, p_f32 = C.POINTER(i64), C.POINTER(u64), C.POINTER(f64), C.POINTER(C.c_float)


class MgrConfig(C.Structure):          # kvb_mgr_config
    _fields_ = [("num_blocks", i64), ("block_size", i64), ("cell_size", i64), ("num_layers", i64),
                ("num_kv_buffers", i64), ("page_size", i64), ("group_id", i64),
                ("min_reserved_pages", i64), ("max_reserved_pages", i64),
                ("reserve_null_block", i32), ("async_sched", i32), ("contiguous_layout", i32),
                ("attach_vmm", i32), ("device_index", i32), ("world_size", i32),
                ("gpu_utilization", f64), ("max_cached_blocks", i64),
                ("publish_shm", i32), ("pad_", i32), ("ipc_name", C.c_char * 64)]


class Op(C.Structure):                 # kvb_op
    _fields_ = [("kind", i32), ("n", i32), ("arg", i64)]


class Ref(C.Structure):                # kvb_ref
    _fields_ = [("op", i32), ("lo", i32), ("hi", i32)]


class Res(C.Structure):                # kvb_res
    _fields_ = [("status", i32), ("n", i32), ("off", i64), ("avail", i64), ("aux", i64)]


class MgrStats(C.Structure):           # kvb_mgr_stats
    _fields_ = [(k, i64) for k in (
        "num_free_pages", "num_inuse_pages", "num_reserved_pages", "num_total_pages",
        "num_avail_blocks", "num_reserved_blocks", "in_shrink", "target_num_blocks",
        "num_avail_pages", "num_full_pages", "num_mapped_pages", "has_null_block",
        "blocks_per_page", "block_mem_size", "page_size", "mem_size",
        "bad_frees", "unknown_page_frees", "launches", "events", "h2d_bytes", "d2h_bytes",
        "num_evictable_blocks")]


class KvLayout(C.Structure):           # kvb_kv_layout
    _fields_ = [("base", u64), ("stride_block", i64), ("stride_token", i64), ("stride_layer", i64),
                ("stride_kv", i64), ("cell_bytes", i64), ("block_size", i32), ("num_layers", i32),
                ("num_kv", i32), ("pad_", i32)]


OP_ALLOC, OP_FREE_IDS, OP_FREE_REF, OP_RESIZE, OP_TRIM, OP_TRY_RESERVE, OP_FREE_RESERVED, OP_AVAIL, OP_CLEAR, \
    OP_ALLOC_PAGE, OP_FREE_PAGES, OP_PA_RESIZE, OP_RESET_ORDER = range(1, 14)
OP_PIN_NEW, OP_PIN_TOUCH, OP_PIN_MARK, OP_PIN_FREE, OP_PIN_EVICT, OP_PIN_RESET = range(14, 20)
ST_OK, ST_FALSE, ST_NONE, ST_NO_PAGES, ST_ASSERT, ST_NULL_BLOCK, ST_CAPACITY = 1, 0, -1, -2, -3, -4, -5

# name -> (restype, argtypes): every symbol include/kvb200.h declares
SIGNATURES = {
    "kvb_last_error": (C.c_char_p, []),
    "kvb_abi_version": (C.c_int, []),
    "kvb_total_launches": (i64, []),
    "kvb_init": (C.c_int, [C.c_int, i64, C.c_int]),
    "kvb_shutdown": (C.c_int, []),
    "kvb_is_initialized": (C.c_int, []),
    "kvb_page_size": (i64, []),
    "kvb_contiguous_layout": (C.c_int, []),
    "kvb_create_kv_tensors": (C.c_int, [i64, i64, i64, i64, C.c_int, p_u64, p_i64, C.c_int]),
    "kvb_kv_tensors_created": (C.c_int, [i64]),
    "kvb_map_to_kv_tensors": (C.c_int, [i64, p_i64, i64]),
    "kvb_unmap_from_kv_tensors": (C.c_int, [i64, p_i64, i64]),
    "kvb_map_pages": (C.c_int, [i64, p_i64, i64]),
    "kvb_unmap_pages": (C.c_int, [i64, p_i64, i64]),
    "kvb_apply_page_events": (C.c_int, [i64, p_i64, i64]),
    "kvb_mgr_alloc_extend": (C.c_int, [C.c_void_p, i64, u64, u64, u64, i64, u64, C.c_void_p, p_i64]),
    "kvb_mgr_alloc_decode": (C.c_int, [C.c_void_p, i64, u64, u64, u64, C.c_void_p, p_i64]),
    "kvb_mgr_free_slots": (i64, [C.c_void_p, u64, i64, C.c_void_p]),
    "kvb_mgr_alloc_device": (i64, [C.c_void_p, i64, u64, C.c_void_p]),
    "kvb_mgr_free_device": (C.c_int, [C.c_void_p, u64, i64, C.c_void_p]),
    "kvb_mgr_profile": (C.c_int, [C.c_void_p, p_i64, C.c_int]),
    "kvb_audit_stress": (C.c_int, [C.c_int, i64, i32, i32, p_f64]),
    "kvb_mgr_shm_name": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "kvb_mgr_shm_publishes": (i64, [C.c_void_p]),
    "kvb_shm_read": (C.c_int, [C.c_char_p, p_i64]),
    "kvb_shm_set_limit": (C.c_int, [C.c_char_p, i64, C.c_int]),
    "kvb_vmm_enqueue_page_events": (C.c_int, [i64, p_i64, i64]),
    "kvb_vmm_wait_maps": (C.c_int, [i64]),
    "kvb_vmm_fence": (C.c_int, [i64]),
    "kvb_vmm_set_async": (C.c_int, [i64, C.c_int]),
    "kvb_vmm_set_retention_ms": (C.c_int, [i64, f64]),
    "kvb_vmm_settle": (C.c_int, [i64]),
    "kvb_vmm_set_retain_free_fraction": (C.c_int, [i64, f64, f64]),
    "kvb_vmm_set_retain_max_fraction": (C.c_int, [i64, f64]),
    "kvb_vmm_footprint": (C.c_int, [i64, p_i64]),
    "kvb_vmm_zero_init_mode": (C.c_int, [C.c_int]),
    "kvb_vmm_wait_init": (C.c_int, [i64]),
    "kvb_vmm_init_progress": (C.c_int, [i64, p_i64]),
    "kvb_vmm_pending_unmaps": (i64, [i64]),
    "kvb_vmm_prefetch_config": (C.c_int, [i64, C.c_int]),
    "kvb_vmm_set_prefetch_hints": (C.c_int, [i64, p_i64, i64]),
    "kvb_vmm_prefetched_pages": (i64, [i64, p_i64, i64]),
    "kvb_vmm_engine_stats": (C.c_int, [i64, p_i64]),
    "kvb_vmm_pool_config": (C.c_int, [i64, i64]),
    "kvb_vmm_pool_prefill": (C.c_int, [i64, i64]),
    "kvb_vmm_stats": (C.c_int, [i64, p_i64]),
    "kvb_vmm_timing": (C.c_int, [i64, p_f64]),
    "kvb_vmm_mapped_pages": (i64, [i64, p_i64, i64]),
    "kvb_vmm_avail_physical_pages": (i64, [i64, i64, i64, f64]),
    "kvb_vmm_raw_control": (C.c_int, [i64, i64, i64, C.c_int, p_f64, p_f64]),
    "kvb_mgr_create": (C.c_void_p, [C.POINTER(MgrConfig)]),
    "kvb_mgr_destroy": (C.c_int, [C.c_void_p]),
    # pointer arguments as plain addresses: the per-step call passes ndarray.ctypes.data (no POINTER objects built per call)
    "kvb_mgr_step": (i64, [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p, i64]),
    "kvb_mgr_alloc": (i64, [C.c_void_p, i64, C.c_void_p, i64]),
    "kvb_mgr_free": (C.c_int, [C.c_void_p, C.c_void_p, i64]),
    "kvb_mgr_sync": (C.c_int, [C.c_void_p]),
    "kvb_mgr_resize": (C.c_int, [C.c_void_p, i64]),
    "kvb_mgr_trim": (C.c_int, [C.c_void_p]),
    "kvb_mgr_try_to_reserve": (C.c_int, [C.c_void_p, i64]),
    "kvb_mgr_free_reserved": (C.c_int, [C.c_void_p]),
    "kvb_mgr_clear": (C.c_int, [C.c_void_p]),
    "kvb_mgr_available_size": (i64, [C.c_void_p]),
    "kvb_mgr_get_stats": (C.c_int, [C.c_void_p, C.POINTER(MgrStats)]),
    "kvb_mgr_set_phys_free_pages": (C.c_int, [C.c_void_p, i64]),
    "kvb_mgr_trace_upload": (C.c_int, [C.c_void_p, C.POINTER(Op), i64, p_i64, i64, C.POINTER(Ref), i64, p_i64, i64, i64]),
    "kvb_mgr_trace_run": (C.c_int, [C.c_void_p, i64, i64, C.c_int, p_f32, p_i64]),
    "kvb_mgr_trace_results": (C.c_int, [C.c_void_p, C.POINTER(Res), i64, p_i64, i64, p_i64]),
    "kvb_mgr_trace_reset_cursor": (C.c_int, [C.c_void_p]),
    "kvb_mgr_dump_list": (i64, [C.c_void_p, C.c_int, p_i64, i64]),
    "kvb_mgr_page_free_blocks": (i64, [C.c_void_p, i64, p_i64, i64]),
    "kvb_mgr_drain_events": (i64, [C.c_void_p, p_i64, i64]),
    "kvb_mgr_audit": (C.c_int, [C.c_void_p, p_i64, p_f32]),
    "kvb_mgr_bitmap": (C.c_int, [C.c_void_p, p_u64, p_i64]),
    "kvb_tp_buffer_words": (i64, []),
    "kvb_mgr_tp_export": (C.c_int, [C.c_void_p, u64, C.c_void_p]),
    "kvb_tp_apply": (C.c_int, [i64, u64, C.c_void_p, p_i64, p_i64]),
    "kvb_mgr_tp_export_host": (C.c_int, [C.c_void_p, p_i64]),
    "kvb_tp_pack": (C.c_int, [i64, p_i64, i64, p_i64, p_i64, i64, p_i64]),
    "kvb_tp_apply_host": (C.c_int, [i64, p_i64, p_i64, p_i64]),
    "kvb_tp_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int, i64, C.c_void_p, C.c_int]),
    "kvb_tp_fence": (C.c_int, [i64]),
    "kvb_tp_publish": (C.c_int, [i64, i64, p_i64, i64, p_i64, i64, C.c_int]),
    "kvb_tp_stats": (C.c_int, [i64, p_i64]),
    "kvb_tp_shutdown": (C.c_int, [i64]),
    "kvb_mgr_tp_detach": (C.c_int, [C.c_void_p]),
    "kvb_page_block_range": (None, [i64, i64, i64, p_i64, p_i64]),
    "kvb_page_num_blocks": (i64, [i64, i64]),
    "kvb_page_id_of_block": (i64, [i64, i64, i64]),
    "kvb_group_indices_by_page": (i64, [C.c_void_p, p_i64, i64, p_i64, p_i64, p_i64]),
    "kvb_kv_scatter": (C.c_int, [C.POINTER(KvLayout), u64, i64, u64, u64, i32, i32, C.c_void_p]),
    "kvb_kv_gather": (C.c_int, [C.POINTER(KvLayout), u64, i64, u64, u64, i32, i32, C.c_void_p]),
    "kvb_kv_scatter_host": (C.c_int, [C.POINTER(KvLayout), p_i64, i64, C.c_void_p, C.c_void_p]),
    "kvb_kv_gather_host": (C.c_int, [C.POINTER(KvLayout), p_i64, i64, C.c_void_p, C.c_void_p]),
}

# include/kvb200_test.h: exported by libkvb200_test.so only (mock VMM driver, driver-cost probes)
TEST_SIGNATURES = {
    "kvb_vmm_mock_driver": (C.c_int, [i64, C.c_int]),
    "kvb_vmm_mock_set_other_bytes": (C.c_int, [i64]),
    "kvb_vmm_mock_stats": (C.c_int, [p_i64]),
    "kvb_vmm_mock_last_violation": (C.c_char_p, []),
    "kvb_vmm_probe": (C.c_int, [C.c_int, i64, i64, i64, p_f64]),
    "kvb_vmm_probe2": (C.c_int, [C.c_int, i64, i64, C.c_int, i64, p_f64]),
    "kvb_vmm_probe3": (C.c_int, [C.c_int, i64, i64, i64, p_f64]),
}
TEST_LIB_PATH = os.path.join(HERE, "libkvb200_test.so")

_lib = None


def load() -> C.CDLL:
    """Load (building first if the sources are newer and nvcc is here) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH) or os.environ.get("KVB200_REBUILD") == "1":
        from . import build as _build
        _build.build()
    # KVB200_LIB: an alternative build (tuning experiments) or "test" = libkvb200_test.so (tests/, tools/ only)
    path = os.environ.get("KVB200_LIB", LIB_PATH)
    if path == "test":
        path = TEST_LIB_PATH
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = the library does not match the header
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in TEST_SIGNATURES.items():   # present in the test library only
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return (load().kvb_last_error() or b"").decode(errors="replace")


class KvbError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise KvbError(f"{what or 'kvb200'} failed ({rc}): {last_error()}")
    return rc
