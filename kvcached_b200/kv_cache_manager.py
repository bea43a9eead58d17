"""KVCacheManager — the reference's block-level allocator API (kvcached/kv_cache_manager.py:58-506)
on top of the device-resident manager of libkvb200.so.

Same constructor and method set as the reference class, so the vLLM / SGLang integration code
(kvcached/integration/*/patches.py) can hold it unchanged:

    alloc(n) -> Optional[List[int]]      free(ids)        available_size()      resize(bytes)
    trim()   try_to_reserve(n)   free_reserved()   get_mapped_memory_size(unit)   clear()

plus the batched entry point the device design exists for:

    step(frees, allocs) -> List[Optional[np.ndarray]]    one kernel launch for a whole scheduler step,
                                                         results identical to the sequential calls.

No bookkeeping happens in this file: every id and counter comes out of the sm_100a kernel through
the C ABI.  If the library is missing the import fails; there is no Python fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import time
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (OP_ALLOC, OP_AVAIL, OP_CLEAR, OP_FREE_IDS, OP_FREE_REF, OP_FREE_RESERVED, OP_RESIZE, OP_TRIM,
                   OP_TRY_RESERVE, ST_ASSERT, ST_NO_PAGES, ST_NONE, ST_OK, KvbError, MgrConfig, MgrStats)

OP_DTYPE = np.dtype([("kind", "<i4"), ("n", "<i4"), ("arg", "<i8")])
REF_DTYPE = np.dtype([("op", "<i4"), ("lo", "<i4"), ("hi", "<i4")])
RES_DTYPE = np.dtype([("status", "<i4"), ("n", "<i4"), ("off", "<i8"), ("avail", "<i8"), ("aux", "<i8")])

KV_TENSOR_WAIT_TIMEOUT = float(os.environ.get("KVCACHED_KV_TENSOR_WAIT_TIMEOUT", "10"))  # kv_cache_manager.py:41
SANITY_CHECK = os.environ.get("KVCACHED_SANITY_CHECK", "false").lower() in ("1", "true")


def shm_read(ipc_name: str):
    """get_kv_cache_limit (kvcached/cli/utils.py:104-113): (total_size, used_size, prealloc_size) or None."""
    out = np.zeros(3, dtype=np.int64)
    rc = _lib.load().kvb_shm_read(ipc_name.encode(), out.ctypes.data_as(_lib.p_i64))
    return None if rc != 0 else tuple(int(x) for x in out)


def shm_set_limit(ipc_name: str, total_bytes: int, create: bool = False) -> bool:
    """update_kv_cache_limit (kvcached/cli/utils.py:116-141): what `kvctl limit` writes."""
    return _lib.load().kvb_shm_set_limit(ipc_name.encode(), int(total_bytes), 1 if create else 0) == 0


class KVCachedConfigError(ValueError):
    """Block larger than a page (kv_cache_manager.py:104-116)."""


def _p64(a: np.ndarray):
    return a.ctypes.data_as(_lib.p_i64)


class _PageAllocatorView:
    """The getters callers and tests read off `manager.page_allocator`
    (csrc/page_allocator.cpp:429-460); all served from the counters the last launch published."""

    def __init__(self, mgr: "KVCacheManager"):
        self._m = mgr

    def get_num_free_pages(self) -> int:
        return self._m.stats().num_free_pages

    def get_num_inuse_pages(self) -> int:
        return self._m.stats().num_inuse_pages

    def get_num_total_pages(self) -> int:
        return self._m.stats().num_total_pages

    def get_num_reserved_pages(self) -> int:
        return self._m.stats().num_reserved_pages

    def get_avail_physical_pages(self) -> int:
        return int(_lib.load().kvb_vmm_avail_physical_pages(self._m.group_id, self._m.num_layers,
                                                            self._m.num_kv_buffers, 0.0))

    def get_page_id(self, block_id: int, block_mem_size: int) -> int:
        return block_id * block_mem_size // self._m.page_size

    def get_resize_target(self) -> int:
        """page_allocator.cpp:780-782: the per-layer size the shm limit asks for, -1 if it equals the current one."""
        return self.check_and_get_resize_target(self._m.mem_size)

    def check_and_get_resize_target(self, cur_mem_size: int) -> int:
        """mem_info_tracker.hpp:187-200"""
        name = self._m.shm_name()
        if not name:
            return -1
        info = shm_read(name)
        if info is None:
            return -1
        per_layer = info[0] // self._m.num_layers // self._m.num_kv_buffers
        return per_layer if per_layer != cur_mem_size else -1

    def trim(self) -> None:
        self._m.trim()

    def resize(self, new_mem_size: int) -> bool:
        return self._m.resize(new_mem_size)

    def start_prealloc_thread(self) -> None:   # handle prefill replaces the prealloc thread
        pass

    def stop_prealloc_thread(self, timeout: Optional[float] = None) -> None:
        pass

    _stop_prealloc_thread = stop_prealloc_thread   # the name kv_cache_manager.py:454 calls (with timeout=...)


class KVCacheManager:
    def __init__(self, num_blocks: int, block_size: int, cell_size: int, num_layers: int, world_size: int = 1,
                 pp_rank: int = 0, async_sched: bool = False, reserve_null_block: bool = False,
                 num_kv_buffers: int = 2, group_id: int = 0, *, page_size: Optional[int] = None,
                 attach_vmm: Optional[bool] = None, device_index: int = -1,
                 min_reserved_pages: int = -1, max_reserved_pages: int = -1, max_cached_blocks: int = -1,
                 publish_shm: Optional[bool] = None, ipc_name: str = ""):
        lib = _lib.load()
        self._lib = lib
        self.num_blocks = num_blocks
        self.block_size = block_size
        self.cell_size = cell_size
        self.block_mem_size = block_size * cell_size
        self.num_layers = num_layers
        self.num_kv_buffers = num_kv_buffers
        self.world_size = world_size
        self.pp_rank = pp_rank
        self.group_id = group_id
        self.async_sched = async_sched
        self.reserve_null_block = reserve_null_block
        if page_size is None:
            page_size = int(lib.kvb_page_size()) if lib.kvb_is_initialized() else \
                int(os.environ.get("KVCACHED_PAGE_SIZE_MB", "2")) << 20
        self.page_size = page_size
        if self.block_mem_size > self.page_size:
            need_mb = -(-self.block_mem_size // (2 << 20)) * 2
            raise KVCachedConfigError(
                f"Block memory size ({self.block_mem_size} bytes = block_size {block_size} x cell_size {cell_size}) "
                f"exceeds page size ({self.page_size} bytes). Set KVCACHED_PAGE_SIZE_MB to at least {need_mb}.")
        self.mem_size = num_blocks * self.block_mem_size            # kv_cache_manager.py:118
        if attach_vmm is None:
            attach_vmm = bool(lib.kvb_is_initialized())
        self.attach_vmm = attach_vmm
        self._cfg = MgrConfig(num_blocks, block_size, cell_size, num_layers, num_kv_buffers, page_size, group_id,
                              min_reserved_pages, max_reserved_pages, int(reserve_null_block), int(async_sched),
                              -1, int(attach_vmm), device_index, world_size, 0.0, max_cached_blocks,
                              -1 if publish_shm is None else int(bool(publish_shm)), 0, ipc_name.encode()[:63])
        self._h = None
        self.null_block: Optional[List[int]] = None
        self.page_allocator = _PageAllocatorView(self)
        self._post_init_done = threading.Event()
        self._post_init_error: Optional[BaseException] = None
        # Like the reference (kv_cache_manager.py:190-227) the manager may be built before the KV
        # tensors exist: a thread waits for them, then creates the device state and reserves the null block.
        if attach_vmm and not lib.kvb_kv_tensors_created(group_id):
            threading.Thread(target=self._post_init, daemon=True).start()
        else:
            self._post_init(wait=False)

    # ------------------------------------------------------------------ lifecycle
    def _post_init(self, wait: bool = True) -> None:
        try:
            if wait:
                t0 = time.time()
                while not self._lib.kvb_kv_tensors_created(self.group_id):
                    if time.time() - t0 > KV_TENSOR_WAIT_TIMEOUT:
                        raise TimeoutError(f"KV tensors not created after {KV_TENSOR_WAIT_TIMEOUT} seconds")
                    time.sleep(0.001)
            h = self._lib.kvb_mgr_create(C.byref(self._cfg))
            if not h:
                raise KvbError(f"kvb_mgr_create failed: {_lib.last_error()}")
            self._h = h
            if self.reserve_null_block:
                self.null_block = [0]
        except BaseException as e:   # surfaced by _wait_post_init
            self._post_init_error = e
        finally:
            self._post_init_done.set()

    def _wait_post_init(self) -> None:
        if not self._post_init_done.is_set():
            self._post_init_done.wait()
        if self._post_init_error is not None:
            raise self._post_init_error

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.kvb_mgr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API
    def alloc(self, need_size: int) -> Optional[List[int]]:
        r = self.alloc_np(need_size)
        return None if r is None else r.tolist()

    def alloc_np(self, need_size: int) -> Optional[np.ndarray]:
        self._wait_post_init()
        # a fresh array per call: ctypes releases the GIL during the C call, so a shared scratch buffer could be
        # overwritten by a second thread's alloc() before it is copied (the reference is @synchronized)
        buf = np.empty(max(int(need_size), 1), dtype=np.int64)
        rc = self._lib.kvb_mgr_alloc(self._h, need_size, buf.ctypes.data, buf.size)
        if rc >= 0:
            return buf[:rc]
        if rc == ST_NONE:
            return None
        if rc == ST_NO_PAGES:
            raise RuntimeError("No free pages left")
        raise KvbError(f"alloc({need_size}) failed ({rc}): {_lib.last_error()}")

    def free(self, indices: Sequence[int]) -> None:
        self._wait_post_init()
        a = np.ascontiguousarray(indices, dtype=np.int64)
        if a.size == 0:
            return
        if SANITY_CHECK:
            rb = set(self.reserved_blocks)
            for idx in a.tolist():
                if idx in rb:
                    raise ValueError(f"Freed index {idx} is in  reserved_blocks, which is not allowed.")
            bad0 = self.stats().bad_frees
        _lib.check(self._lib.kvb_mgr_free(self._h, a.ctypes.data, a.size), "free")
        if SANITY_CHECK and self.stats().bad_frees != bad0:
            raise ValueError("free(): some indices were not allocated (state inconsistency)")

    def sync(self) -> None:
        """free() returns before its kernel does; sync() waits for everything launched so far."""
        self._wait_post_init()
        _lib.check(self._lib.kvb_mgr_sync(self._h), "sync")

    def try_to_reserve(self, need_size: int) -> bool:
        self._wait_post_init()
        rc = self._lib.kvb_mgr_try_to_reserve(self._h, need_size)
        if rc == ST_NO_PAGES:
            raise RuntimeError("No free pages left")
        if rc < -50:
            raise KvbError(f"try_to_reserve failed: {_lib.last_error()}")
        return rc == ST_OK

    def free_reserved(self) -> None:
        self._wait_post_init()
        _lib.check(self._lib.kvb_mgr_free_reserved(self._h), "free_reserved")

    def resize(self, new_mem_size: int) -> bool:
        self._wait_post_init()
        assert new_mem_size > 0, "new_mem_size must be positive"
        rc = self._lib.kvb_mgr_resize(self._h, new_mem_size)
        if rc == ST_ASSERT:
            raise AssertionError("Reserved blocks must be freed before resizing.")
        if rc < -50:
            raise KvbError(f"resize failed: {_lib.last_error()}")
        return rc == ST_OK

    def trim(self) -> None:
        self._wait_post_init()
        _lib.check(self._lib.kvb_mgr_trim(self._h), "trim")

    def available_size(self) -> int:
        self._wait_post_init()
        return int(self._lib.kvb_mgr_available_size(self._h))

    def get_mapped_memory_size(self, unit: str = "bytes") -> float:
        b = self.stats().num_inuse_pages * self.num_layers * self.page_size * self.num_kv_buffers
        div = {"bytes": 1, "kb": 1024, "mb": 1024 ** 2, "gb": 1024 ** 3}
        if unit not in div:
            raise ValueError(f"Unknown unit: {unit}")
        return b if unit == "bytes" else b / div[unit]

    def clear(self) -> None:
        self._wait_post_init()
        rc = self._lib.kvb_mgr_clear(self._h)
        if rc < 0:
            raise RuntimeError(f"Failed to reserve null block at index 0 ({rc}): {_lib.last_error()}")
        self.null_block = [0] if self.reserve_null_block else None

    # ------------------------------------------------------------------ attributes of the reference object
    @property
    def reserved_blocks(self) -> List[int]:
        return self._dump(5)

    @property
    def num_avail_blocks(self) -> int:
        return self.stats().num_avail_blocks

    @property
    def in_shrink(self) -> bool:
        return bool(self.stats().in_shrink)

    @property
    def target_num_blocks(self) -> Optional[int]:
        t = self.stats().target_num_blocks
        return None if t < 0 else t

    # ------------------------------------------------------------------ batched fast path
    def run_ops(self, ops: np.ndarray, ids: Optional[np.ndarray] = None, refs: Optional[np.ndarray] = None,
                out_cap: Optional[int] = None):
        """Run a batch of requests (OP_DTYPE array) in one launch; returns (res, out_ids)."""
        self._wait_post_init()
        n = int(ops.shape[0])
        res = np.empty(n, dtype=RES_DTYPE)
        if out_cap is None:
            k = ops["kind"] == OP_ALLOC
            out_cap = int(ops["n"][k].sum()) if k.any() else 0
        out = np.empty(max(out_cap, 1), dtype=np.int64)
        n_ids = 0 if ids is None else int(ids.size)
        n_refs = 0 if refs is None else int(refs.shape[0])
        rc = self._lib.kvb_mgr_step(
            self._h, ops.ctypes.data, n,
            ids.ctypes.data if n_ids else None, n_ids,
            refs.ctypes.data if n_refs else None, n_refs,
            res.ctypes.data, out.ctypes.data, out.size)
        if rc < 0:
            raise KvbError(f"kvb_mgr_step failed ({rc}): {_lib.last_error()}")
        return res, out[:rc]

    def step(self, frees: Iterable[Sequence[int]], allocs: Sequence[int]) -> List[Optional[np.ndarray]]:
        """One scheduler step: every free(ids) of `frees` in order, then every alloc(n) of `allocs` in
        order — one kernel launch, sequential semantics.  Returns one entry per alloc (None = the
        reference's None)."""
        frees = [np.ascontiguousarray(f, dtype=np.int64) for f in frees]
        frees = [f for f in frees if f.size]
        nf, na = len(frees), len(allocs)
        ops = np.zeros(nf + na, dtype=OP_DTYPE)
        if nf:
            lens = np.fromiter((f.size for f in frees), dtype=np.int64, count=nf)
            ops["kind"][:nf] = OP_FREE_IDS
            ops["n"][:nf] = lens
            ops["arg"][:nf] = np.concatenate(([0], np.cumsum(lens)[:-1]))
            ids = np.concatenate(frees)
        else:
            ids = None
        ops["kind"][nf:] = OP_ALLOC
        ops["n"][nf:] = np.asarray(allocs, dtype=np.int64)
        res, out = self.run_ops(ops, ids)
        ret: List[Optional[np.ndarray]] = []
        for r in res[nf:]:
            if r["status"] == ST_OK:
                ret.append(out[r["off"]:r["off"] + r["n"]])
            elif r["status"] == ST_NO_PAGES:
                raise RuntimeError("No free pages left")
            else:
                ret.append(None)
        return ret

    def shm_name(self) -> str:
        """Name of the usage segment under /dev/shm ('' = publishing is off)."""
        self._wait_post_init()
        buf = C.create_string_buffer(128)
        n = self._lib.kvb_mgr_shm_name(self._h, buf, 128)
        return buf.value.decode() if n > 0 else ""

    # ------------------------------------------------------------------ inspection
    def stats(self) -> MgrStats:
        self._wait_post_init()
        s = MgrStats()
        _lib.check(self._lib.kvb_mgr_get_stats(self._h, C.byref(s)), "get_stats")
        return s

    def _dump(self, which: int) -> List[int]:
        self._wait_post_init()
        cap = max((2 if which == 9 else 1) * self.num_blocks + 64, 4096)
        buf = np.empty(cap, dtype=np.int64)
        n = self._lib.kvb_mgr_dump_list(self._h, which, _p64(buf), cap)
        if n < 0:
            raise KvbError(f"dump_list({which}) failed ({n}): {_lib.last_error()}")
        return buf[:n].tolist()

    def free_page_list(self) -> List[int]:
        return self._dump(0)

    def reserved_page_list(self) -> List[int]:
        return self._dump(1)

    def reclaimed_page_list(self) -> List[int]:
        return self._dump(2)

    def avail_pages(self) -> List[int]:
        return self._dump(3)

    def full_pages(self) -> List[int]:
        return self._dump(4)

    def mapped_pages(self) -> List[int]:
        return self._dump(6)

    def page_free_blocks(self, page_id: int) -> Optional[List[int]]:
        self._wait_post_init()
        buf = np.empty(self.stats().blocks_per_page + 1, dtype=np.int64)
        n = self._lib.kvb_mgr_page_free_blocks(self._h, page_id, _p64(buf), buf.size)
        if n == -1:
            return None
        if n < 0:
            raise KvbError(f"page_free_blocks failed ({n}): {_lib.last_error()}")
        return buf[:n].tolist()

    def drain_events(self) -> List[tuple]:
        """Ordered (page_id, is_unmap) list the kernel emitted since the last drain."""
        self._wait_post_init()
        buf = np.empty(1 << 16, dtype=np.int64)
        while True:
            n = self._lib.kvb_mgr_drain_events(self._h, _p64(buf), buf.size)
            if n == -5:
                buf = np.empty(buf.size * 4, dtype=np.int64)
                continue
            break
        if n < 0:
            raise KvbError(f"drain_events failed ({n})")
        return [(int(e) >> 1, int(e) & 1) for e in buf[:n]]

    def set_phys_free_pages(self, n: int) -> None:
        self._wait_post_init()
        self._lib.kvb_mgr_set_phys_free_pages(self._h, n)

    def audit(self):
        """Bitmap sweep kernel: (used blocks, in-use pages, fully-free in-use pages, mismatches, ms)."""
        self._wait_post_init()
        out = np.zeros(4, dtype=np.int64)
        ms = C.c_float(0)
        _lib.check(self._lib.kvb_mgr_audit(self._h, _p64(out), C.byref(ms)), "audit")
        return int(out[0]), int(out[1]), int(out[2]), int(out[3]), float(ms.value)

    def group_indices_by_page(self, ids: Sequence[int]):
        """csrc/page_allocator.cpp:471-498 computed by the free() kernel path: [(page, [ids...]), ...]
        in the reference's dict order."""
        self._wait_post_init()
        a = np.ascontiguousarray(ids, dtype=np.int64)
        if a.size == 0:
            return []
        pages = np.empty(a.size, dtype=np.int64)
        counts = np.empty(a.size, dtype=np.int64)
        grouped = np.empty(a.size, dtype=np.int64)
        d = self._lib.kvb_group_indices_by_page(self._h, _p64(a), a.size, _p64(pages), _p64(counts), _p64(grouped))
        if d < 0:
            raise KvbError(f"group_indices_by_page failed ({d}): {_lib.last_error()}")
        out, w = [], 0
        for j in range(d):
            out.append((int(pages[j]), grouped[w:w + counts[j]].tolist()))
            w += int(counts[j])
        return out

    def snapshot(self) -> dict:
        """Same record oracle.OracleManager.snapshot() / oracle/ref_replay.py produce."""
        s = self.stats()
        return {
            "avail": self.available_size(),
            "pages": [s.num_free_pages, s.num_inuse_pages, s.num_reserved_pages, s.num_total_pages],
            "nab": s.num_avail_blocks,
            "ap": self.avail_pages(),
            "fp": self.full_pages(),
            "rb": self.reserved_blocks,
            "shrink": bool(s.in_shrink),
            "mapped": self.mapped_pages(),
        }
