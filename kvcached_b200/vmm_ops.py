"""`kvcached.vmm_ops` re-created on the C ABI of libkvb200.so.

Same names, arguments, defaults and return values as the reference's pybind module
(csrc/torch_bindings.cpp:182-258), so `kvcached/kv_cache_manager.py`, `tp_ipc_util.py` and both
`integration/*/interfaces.py` import it unchanged (`sys.modules["kvcached.vmm_ops"]`).
torch is used only to wrap the reserved device range in a tensor (the reference's at::from_blob).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import KvbError

_MAX_RANGES = 4096


def _dev_index(dev_str: str) -> int:
    """"cuda", "cuda:3" -> index (-1 = current device)."""
    s = str(dev_str)
    if not s.startswith("cuda"):
        raise RuntimeError(f"kvcached_b200 is CUDA-only (B200); got device {dev_str!r}")
    return int(s.split(":", 1)[1]) if ":" in s else -1


def init_kvcached(dev_str: str, page_size: int = 0, contiguous_layout: bool = True) -> None:
    """torch_bindings.cpp:20-24"""
    _lib.check(_lib.load().kvb_init(_dev_index(dev_str), page_size, int(contiguous_layout)), "init_kvcached")


def shutdown_kvcached() -> None:
    """torch_bindings.cpp:26-29: every virtual tensor handed out before dangles afterwards."""
    _lib.check(_lib.load().kvb_shutdown(), "shutdown_kvcached")


class _DeviceRange:
    """Exposes a reserved VA range through __cuda_array_interface__ (the at::from_blob of ftensor.cpp:66-75)."""

    def __init__(self, ptr: int, nbytes: int, itemsize: int):
        typestr = {1: "|i1", 2: "<i2", 4: "<i4", 8: "<i8"}.get(itemsize)
        if typestr is None:   # impl/torch_utils.ipp:32-46
            raise RuntimeError(f"Unsupported dtype size: {itemsize}")
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr,
                                         "data": (ptr, False), "version": 3, "strides": None, "stream": None}


def create_kv_tensors_raw(size: int, num_layers: int, num_kv_buffers: int = 2, group_id: int = 0,
                          unified_pool: bool = False) -> List[Tuple[int, int]]:
    """[(device pointer, bytes)] of the reserved ranges — no torch involved."""
    ptrs = (C.c_uint64 * _MAX_RANGES)()
    sizes = (C.c_int64 * _MAX_RANGES)()
    n = _lib.check(_lib.load().kvb_create_kv_tensors(group_id, size, num_layers, num_kv_buffers, int(unified_pool),
                                                     ptrs, sizes, _MAX_RANGES), "create_kv_tensors")
    return [(int(ptrs[i]), int(sizes[i])) for i in range(n)]


def create_kv_tensors(size: int, dtype_size: int, dev_str: str, num_layers: int, num_kv_buffers: int = 2,
                      group_id: int = 0, unified_pool: bool = False):
    """torch_bindings.cpp:31-40: ONE flat tensor (contiguous layout) or `num_layers` flat tensors, integer
    dtype chosen by itemsize; callers .view() them."""
    import torch
    idx = _dev_index(dev_str)
    dev = torch.device("cuda", idx if idx >= 0 else torch.cuda.current_device())
    out = []
    for ptr, nbytes in create_kv_tensors_raw(size, num_layers, num_kv_buffers, group_id, unified_pool):
        out.append(torch.as_tensor(_DeviceRange(ptr, nbytes, dtype_size), device=dev))
    return out


def kv_tensors_created(group_id: int = 0) -> bool:
    return bool(_lib.load().kvb_kv_tensors_created(group_id))


def _arr(v) -> np.ndarray:
    return np.ascontiguousarray(v, dtype=np.int64)


def map_to_kv_tensors(offsets: List[int], group_id: int = 0) -> bool:
    """torch_bindings.cpp:48-53: byte offsets = page_id * page_size * (L * nKV if contiguous else 1)."""
    a = _arr(offsets)
    return bool(_lib.check(_lib.load().kvb_map_to_kv_tensors(group_id, a.ctypes.data_as(_lib.p_i64), a.size),
                           "map_to_kv_tensors"))


def unmap_from_kv_tensors(offsets: List[int], group_id: int = 0) -> bool:
    a = _arr(offsets)
    return bool(_lib.check(_lib.load().kvb_unmap_from_kv_tensors(group_id, a.ctypes.data_as(_lib.p_i64), a.size),
                           "unmap_from_kv_tensors"))


def mapped_pages(group_id: int = 0) -> List[int]:
    buf = np.empty(1 << 20, dtype=np.int64)
    n = _lib.check(_lib.load().kvb_vmm_mapped_pages(group_id, buf.ctypes.data_as(_lib.p_i64), buf.size))
    return buf[:n].tolist()


def fence(group_id: int = 0) -> None:
    """Wait until the group's VMM thread has applied every queued map AND unmap (kvb_vmm_fence)."""
    _lib.check(_lib.load().kvb_vmm_fence(group_id), "vmm fence")


def settle(group_id: int = 0) -> None:
    """fence() + every retained page back on the zero page (what trim / clear / resize do)."""
    _lib.check(_lib.load().kvb_vmm_settle(group_id), "vmm settle")


def set_retention(ms: float, group_id: int = 0) -> None:
    """How long an unmapped page keeps its physical handle mapped (0 = unmap at once)."""
    _lib.check(_lib.load().kvb_vmm_set_retention_ms(group_id, float(ms)), "vmm set_retention_ms")


def set_retain_free_fraction(fraction: float, quiet_margin: float = -1.0, group_id: int = 0) -> None:
    """The free-memory reserve of retained pages: `fraction` of the device while somebody else's footprint grows,
    headroom + `quiet_margin` while nobody else moves (quiet_margin < 0 keeps the current one)."""
    _lib.check(_lib.load().kvb_vmm_set_retain_free_fraction(group_id, float(fraction), float(quiet_margin)),
               "vmm set_retain_free_fraction")


def set_retain_max_fraction(fraction: float, group_id: int = 0) -> None:
    """Retained pages never hold more than this fraction of the device (default 0.10)."""
    _lib.check(_lib.load().kvb_vmm_set_retain_max_fraction(group_id, float(fraction)), "vmm set_retain_max_fraction")


def footprint(group_id: int = 0) -> Dict[str, int]:
    """Physical footprint against the bookkeeping (peaks are since the previous call)."""
    o = np.zeros(8, dtype=np.int64)
    _lib.check(_lib.load().kvb_vmm_footprint(group_id, o.ctypes.data_as(_lib.p_i64)), "vmm footprint")
    keys = ["physical_pages", "wanted_pages", "peak_overshoot_pages", "retained_pages", "retain_cap_pages",
            "peak_retained_pages", "pooled_pages", "pool_cap_pages"]
    return dict(zip(keys, o.tolist()))


def zero_init_mode(mode: int) -> int:
    """-1 follow KVB200_VMM_ZERO_INIT, 0 eager (default), 1 background, 2 none (no zero page: unmapped pages fault);
    call before create_kv_tensors.  Returns the old mode."""
    return int(_lib.load().kvb_vmm_zero_init_mode(int(mode)))


def wait_init(group_id: int = 0) -> None:
    """Background zero-page init: returns when every unit of the range is backed."""
    _lib.check(_lib.load().kvb_vmm_wait_init(group_id), "vmm wait_init")


def init_progress(group_id: int = 0) -> Dict[str, float]:
    o = np.zeros(3, dtype=np.int64)
    _lib.check(_lib.load().kvb_vmm_init_progress(group_id, o.ctypes.data_as(_lib.p_i64)), "vmm init_progress")
    return {"units_done": int(o[0]), "units": int(o[1]), "init_ms": float(o[2]) / 1e3}


def wait_maps(group_id: int = 0) -> None:
    _lib.check(_lib.load().kvb_vmm_wait_maps(group_id), "vmm wait_maps")


def set_async(on: bool, group_id: int = 0) -> None:
    _lib.check(_lib.load().kvb_vmm_set_async(group_id, 1 if on else 0), "vmm set_async")


def pending_unmaps(group_id: int = 0) -> int:
    return int(_lib.check(_lib.load().kvb_vmm_pending_unmaps(group_id)))


def set_prefetch(depth: int, group_id: int = 0) -> None:
    """Pages the VMM thread maps ahead of the bookkeeping's next allocations (0 = off)."""
    _lib.check(_lib.load().kvb_vmm_prefetch_config(group_id, int(depth)), "vmm prefetch_config")


def prefetched_pages(group_id: int = 0) -> List[int]:
    buf = np.empty(64, dtype=np.int64)
    n = _lib.check(_lib.load().kvb_vmm_prefetched_pages(group_id, buf.ctypes.data_as(_lib.p_i64), buf.size))
    return buf[:n].tolist()


def engine_stats(group_id: int = 0) -> Dict[str, int]:
    s = np.zeros(8, dtype=np.int64)
    _lib.check(_lib.load().kvb_vmm_engine_stats(group_id, s.ctypes.data_as(_lib.p_i64)))
    keys = ["elided", "prefetch_hits", "mapped_ahead", "need_map", "need_unmap", "need_hinted", "retained", "retention_hits"]
    return dict(zip(keys, s.tolist()))


def vmm_stats(group_id: int = 0) -> Dict[str, int]:
    s = np.zeros(8, dtype=np.int64)
    _lib.check(_lib.load().kvb_vmm_stats(group_id, s.ctypes.data_as(_lib.p_i64)))
    keys = ["mapped_pages", "pooled_handles", "handles_created", "handles_released", "map_calls", "unmap_calls",
            "bytes_mapped", "page_bytes"]
    return dict(zip(keys, s.tolist()))


def vmm_timing(group_id: int = 0) -> Dict[str, float]:
    t = np.zeros(10)
    _lib.check(_lib.load().kvb_vmm_timing(group_id, t.ctypes.data_as(_lib.p_f64)))
    names = ["create", "release", "unmap", "map", "setaccess"]
    out = {}
    for i, n in enumerate(names):
        out[n + "_ms"] = float(t[i]) / 1e3
        out[n + "_calls"] = int(t[5 + i])
    return out


# ------------------------------------------------------------------------------------------------
# InternalPage — csrc/page_allocator.cpp:40-100.  A host-side value object in the reference too (a
# page id plus its FIFO of free block ids); the device manager keeps the same FIFO in HBM, this class
# exists for callers of the page-level API.
class InternalPage:
    def __init__(self, page_id: int, page_size: int):
        self.page_id = page_id
        self.page_size = page_size
        self.start_block: Optional[int] = None
        self.end_block: Optional[int] = None
        self.num_kv_blocks: Optional[int] = None
        self.free_list: List[int] = []

    def init(self, block_mem_size: int) -> None:
        self.start_block, self.end_block = self.get_block_range(self.page_id, self.page_size, block_mem_size)
        self.num_kv_blocks = self.end_block - self.start_block
        self.free_list = list(range(self.start_block, self.end_block))

    def alloc(self, num_blocks: int = 1) -> List[int]:
        if len(self.free_list) < num_blocks:
            raise RuntimeError("Not enough free blocks in page")   # page_allocator.cpp:58-60
        out, self.free_list = self.free_list[:num_blocks], self.free_list[num_blocks:]
        return out

    def free(self, block_id: int) -> None:
        self.free_list.append(block_id)

    def free_batch(self, block_ids: List[int]) -> None:
        self.free_list.extend(block_ids)

    def empty(self) -> bool:
        return len(self.free_list) == self.num_kv_blocks

    def full(self) -> bool:
        return not self.free_list

    def num_free_blocks(self) -> int:
        return len(self.free_list)

    def get_free_blocks(self) -> List[int]:
        return list(self.free_list)

    @staticmethod
    def get_block_range(page_id: int, page_size: int, block_mem_size: int) -> Tuple[int, int]:
        s, e = C.c_int64(), C.c_int64()
        _lib.load().kvb_page_block_range(page_id, page_size, block_mem_size, C.byref(s), C.byref(e))
        return s.value, e.value

    @staticmethod
    def get_num_blocks(page_size: int, block_mem_size: int) -> int:
        return int(_lib.load().kvb_page_num_blocks(page_size, block_mem_size))


# ------------------------------------------------------------------------------------------------
class PageAllocator:
    """csrc/page_allocator.cpp:103-782 on the device-resident page lists: the free / reserved / reclaimed
    deques and counters live in HBM and are advanced by the step kernel (requests ALLOC_PAGE,
    FREE_PAGES, PA_RESIZE, TRIM, RESET_ORDER); map/unmap goes to the batched VMM engine or, when
    callbacks are installed, to the caller (the reference's broadcast hooks)."""

    def __init__(self, num_layers: int, mem_size_per_layer: int, page_size: int, world_size: int = 1,
                 pp_rank: int = 0, async_sched: bool = False, contiguous_layout: bool = True,
                 enable_page_prealloc: bool = True, num_kv_buffers: int = 2, group_id: int = 0,
                 ipc_name: str = ""):
        from .kv_cache_manager import KVCacheManager
        self.num_layers = num_layers
        self.page_size = page_size
        self.world_size = world_size
        self.num_kv_buffers = num_kv_buffers
        self.group_id = group_id
        self.contiguous_layout = contiguous_layout
        self._map_cb: Optional[Callable] = None
        self._unmap_cb: Optional[Callable] = None
        self._ipc_cb: Optional[Callable] = None
        # one block per page: the block manager's page lists ARE the page allocator
        lib = _lib.load()
        attach = bool(lib.kvb_is_initialized())   # tensors may not exist yet: the manager waits for them (post-init)
        self._m = KVCacheManager(mem_size_per_layer // page_size, 1, page_size, num_layers, world_size, pp_rank,
                                 async_sched, False, num_kv_buffers, group_id, page_size=page_size,
                                 attach_vmm=attach, ipc_name=ipc_name)
        self._attach = attach

    # -- requests ------------------------------------------------------------------------------
    def _one(self, kind: int, n: int = 0, arg: int = 0, ids=None):
        from .kv_cache_manager import OP_DTYPE
        ops = np.zeros(1, dtype=OP_DTYPE)
        ops["kind"], ops["n"], ops["arg"] = kind, n, arg
        res, _ = self._m.run_ops(ops, None if ids is None else _arr(ids), None, 0)
        self._forward_events()
        return res[0]

    def _forward_events(self) -> None:
        ev = self._m.drain_events()
        if not ev or self._attach and not self._use_callbacks():
            return
        per = self.page_size * (self.num_layers * self.num_kv_buffers if self.contiguous_layout else 1)
        i = 0
        while i < len(ev):          # runs of equal kind keep the order
            j = i
            while j < len(ev) and ev[j][1] == ev[i][1]:
                j += 1
            offs = [p * per for p, _ in ev[i:j]]
            cb = self._unmap_cb if ev[i][1] else self._map_cb
            if cb is not None:
                cb(self.world_size, offs)
            i = j

    def _use_callbacks(self) -> bool:
        if self._map_cb is None and self._unmap_cb is None:
            return False
        return self.world_size > 1 or (self._ipc_cb is not None and bool(self._ipc_cb()))

    def alloc_page(self) -> InternalPage:
        r = self._one(_lib.OP_ALLOC_PAGE)
        if r["status"] != _lib.ST_OK:
            raise RuntimeError("No free pages left")
        return InternalPage(int(r["off"]), self.page_size)

    def free_page(self, page_id: int) -> None:
        self.free_pages([page_id])

    def free_pages(self, page_ids: List[int]) -> None:
        if len(page_ids):
            self._one(_lib.OP_FREE_PAGES, len(page_ids), 0, page_ids)

    def resize(self, new_mem_size: int) -> bool:
        return self._one(_lib.OP_PA_RESIZE, 0, new_mem_size)["status"] == _lib.ST_OK

    def trim(self) -> None:
        self._one(_lib.OP_TRIM)

    def reset_free_page_order(self) -> None:
        self._one(_lib.OP_RESET_ORDER)

    # -- getters (page_allocator.cpp:429-460) -----------------------------------------------------
    def get_num_free_pages(self) -> int:
        return self._m.stats().num_free_pages

    def get_num_inuse_pages(self) -> int:
        return self._m.stats().num_inuse_pages

    def get_num_total_pages(self) -> int:
        return self._m.stats().num_total_pages

    def get_num_reserved_pages(self) -> int:
        return self._m.stats().num_reserved_pages

    def get_avail_physical_pages(self) -> int:
        return self._m.page_allocator.get_avail_physical_pages()

    def check_and_get_resize_target(self, cur_mem_size: int) -> int:
        return self._m.page_allocator.check_and_get_resize_target(cur_mem_size)

    def get_resize_target(self) -> int:
        return self._m.page_allocator.get_resize_target()

    def get_page_id(self, block_id: int, block_mem_size: int) -> int:
        return int(_lib.load().kvb_page_id_of_block(block_id, block_mem_size, self.page_size))

    def group_indices_by_page(self, indices: List[int], block_mem_size: int) -> Dict[int, List[int]]:
        """page_allocator.cpp:471-498 — dict in the reference's (libstdc++) iteration order, computed by the
        device code path free() uses."""
        from .kv_cache_manager import KVCacheManager
        key = ("g", block_mem_size)
        g = getattr(self, "_grp", {}).get(key)
        if g is None:
            nb = self._m.num_blocks * self.page_size // block_mem_size
            g = KVCacheManager(max(nb, 1), 1, block_mem_size, self.num_layers, page_size=self.page_size,
                               attach_vmm=False)
            self._grp = {**getattr(self, "_grp", {}), key: g}
        return {p: ids for p, ids in g.group_indices_by_page(indices)}

    # -- hooks -----------------------------------------------------------------------------------
    def set_broadcast_map_callback(self, fn: Callable) -> None:
        self._map_cb = fn

    def set_broadcast_unmap_callback(self, fn: Callable) -> None:
        self._unmap_cb = fn

    def set_should_use_worker_ipc_callback(self, fn: Callable) -> None:
        self._ipc_cb = fn

    def start_prealloc_thread(self) -> None:
        pass

    def stop_prealloc_thread(self, timeout: Optional[float] = None) -> None:
        pass

    # kv_cache_manager.py:454 calls the private name with a timeout; torch_bindings.cpp:212-213 binds only the public
    # one, so the reference's own clear() raises AttributeError there.  Both names are served here.
    _stop_prealloc_thread = stop_prealloc_thread
