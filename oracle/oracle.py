"""ctypes front end of oracle/kv_oracle.cpp (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by kvcached_b200/.

`OracleManager` mirrors the reference `KVCacheManager` (kvcached/kv_cache_manager.py:58)
method for method; `replay(trace)` produces the same per-op records as
oracle/ref_replay.py does for the real reference, so the two can be diffed directly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kv_oracle.cpp")
LIB = os.path.join(HERE, "libkvoracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return LIB


class _Cfg(C.Structure):
    _fields_ = [("num_blocks", C.c_int64), ("block_mem_size", C.c_int64), ("page_size", C.c_int64),
                ("num_layers", C.c_int64), ("num_kv_buffers", C.c_int64),
                ("min_reserved_pages", C.c_int64), ("max_reserved_pages", C.c_int64),
                ("reserve_null_block", C.c_int32), ("prealloc", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        p64 = C.POINTER(C.c_int64)
        _lib.kvo_create.restype = C.c_void_p
        _lib.kvo_create.argtypes = [C.POINTER(_Cfg)]
        _lib.kvo_destroy.argtypes = [C.c_void_p]
        _lib.kvo_alloc.restype = C.c_int64
        _lib.kvo_alloc.argtypes = [C.c_void_p, C.c_int64, p64, C.c_int64]
        _lib.kvo_free.argtypes = [C.c_void_p, p64, C.c_int64]
        _lib.kvo_resize.argtypes = [C.c_void_p, C.c_int64]
        _lib.kvo_trim.argtypes = [C.c_void_p]
        _lib.kvo_try_to_reserve.argtypes = [C.c_void_p, C.c_int64]
        _lib.kvo_free_reserved.argtypes = [C.c_void_p]
        _lib.kvo_available_size.restype = C.c_int64
        _lib.kvo_available_size.argtypes = [C.c_void_p]
        _lib.kvo_clear.argtypes = [C.c_void_p]
        _lib.kvo_set_phys_free_pages.argtypes = [C.c_void_p, C.c_int64]
        _lib.kvo_set_resize_target.argtypes = [C.c_void_p, C.c_int64]
        _lib.kvo_scalars.argtypes = [C.c_void_p, p64]
        for name in ("kvo_avail_pages", "kvo_full_pages", "kvo_reserved_blocks", "kvo_mapped_pages",
                     "kvo_drain_events"):
            f = getattr(_lib, name)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p, p64, C.c_int64]
        _lib.kvo_page_list.restype = C.c_int64
        _lib.kvo_page_list.argtypes = [C.c_void_p, C.c_int, p64, C.c_int64]
        _lib.kvo_page_free_blocks.restype = C.c_int64
        _lib.kvo_page_free_blocks.argtypes = [C.c_void_p, C.c_int64, p64, C.c_int64]
        _lib.kvo_group_indices_by_page.restype = C.c_int64
        _lib.kvo_group_indices_by_page.argtypes = [p64, C.c_int64, C.c_int64, C.c_int64, p64, p64, p64]
        _lib.kvo_block_range.argtypes = [C.c_int64, C.c_int64, C.c_int64, p64, p64]
        _lib.kvo_num_blocks.restype = C.c_int64
        _lib.kvo_num_blocks.argtypes = [C.c_int64, C.c_int64]
    return _lib


def _p64(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class OracleManager:
    """Reference KVCacheManager semantics (kv_cache_manager.py:58-506) on the CPU."""

    def __init__(self, num_blocks: int, block_size: int, cell_size: int, num_layers: int,
                 reserve_null_block: bool = False, num_kv_buffers: int = 2,
                 page_size: int = 2 << 20, prealloc: bool = False,
                 min_reserved_pages: int = 5, max_reserved_pages: int = 10):
        self.block_mem_size = block_size * cell_size
        self.page_size = page_size
        self.num_layers = num_layers
        self.num_kv_buffers = num_kv_buffers
        cfg = _Cfg(num_blocks, self.block_mem_size, page_size, num_layers, num_kv_buffers,
                   min_reserved_pages, max_reserved_pages, int(reserve_null_block), int(prealloc))
        self._h = lib().kvo_create(C.byref(cfg))
        if not self._h:
            raise RuntimeError("oracle: invalid config (block larger than page, or null block != 0)")
        self._cap = max(num_blocks + 64, 4096)
        self._buf = np.empty(self._cap, dtype=np.int64)
        self.null_block = [0] if reserve_null_block else None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kvo_destroy(self._h)
            self._h = None

    def _ensure(self, n: int):
        if n > self._cap:
            self._cap = int(n * 2)
            self._buf = np.empty(self._cap, dtype=np.int64)

    def alloc(self, n: int) -> Optional[List[int]]:
        self._ensure(n)
        rc = lib().kvo_alloc(self._h, n, _p64(self._buf), self._cap)
        if rc == -1:
            return None
        if rc == -2:
            raise RuntimeError("No free pages left")
        if rc < 0:
            raise RuntimeError(f"oracle alloc rc={rc}")
        return self._buf[:rc].tolist()

    def alloc_np(self, n: int) -> Optional[np.ndarray]:
        self._ensure(n)
        rc = lib().kvo_alloc(self._h, n, _p64(self._buf), self._cap)
        if rc < 0:
            return None
        return self._buf[:rc].copy()

    def free(self, ids) -> None:
        a = np.ascontiguousarray(ids, dtype=np.int64)
        lib().kvo_free(self._h, _p64(a), a.size)

    def resize(self, new_mem_size: int) -> bool:
        rc = lib().kvo_resize(self._h, new_mem_size)
        if rc == -3:
            raise AssertionError("Reserved blocks must be freed before resizing.")
        return bool(rc)

    def trim(self) -> None:
        lib().kvo_trim(self._h)

    def try_to_reserve(self, n: int) -> bool:
        rc = lib().kvo_try_to_reserve(self._h, n)
        if rc == -2:
            raise RuntimeError("No free pages left")
        return bool(rc)

    def free_reserved(self) -> None:
        lib().kvo_free_reserved(self._h)

    def available_size(self) -> int:
        return int(lib().kvo_available_size(self._h))

    def clear(self) -> None:
        if lib().kvo_clear(self._h) != 0:
            raise RuntimeError("Failed to reserve null block at index 0")

    def set_phys_free_pages(self, n: int) -> None:
        lib().kvo_set_phys_free_pages(self._h, n)

    def set_resize_target(self, nbytes: int) -> None:
        lib().kvo_set_resize_target(self._h, nbytes)

    # ---- observation ---------------------------------------------------------
    def scalars(self) -> dict:
        s = np.zeros(12, dtype=np.int64)
        lib().kvo_scalars(self._h, _p64(s))
        keys = ["free_pages", "inuse_pages", "reserved_pages", "total_pages", "num_avail_blocks", "in_shrink",
                "n_avail_pages", "n_full_pages", "n_reserved_blocks", "n_mapped", "has_null", "target_num_blocks"]
        return dict(zip(keys, s.tolist()))

    def _list(self, fn, *pre) -> List[int]:
        while True:
            rc = fn(self._h, *pre, _p64(self._buf), self._cap)
            if rc == -9:
                self._ensure(self._cap * 2)
                continue
            return self._buf[:rc].tolist()

    def avail_pages(self) -> List[int]:
        return self._list(lib().kvo_avail_pages)

    def full_pages(self) -> List[int]:
        return self._list(lib().kvo_full_pages)

    @property
    def reserved_blocks(self) -> List[int]:
        return self._list(lib().kvo_reserved_blocks)

    def mapped_pages(self) -> List[int]:
        return self._list(lib().kvo_mapped_pages)

    def page_list(self, which: int) -> List[int]:
        return self._list(lib().kvo_page_list, which)

    def page_free_blocks(self, pid: int) -> Optional[List[int]]:
        rc = lib().kvo_page_free_blocks(self._h, pid, _p64(self._buf), self._cap)
        return None if rc < 0 else self._buf[:rc].tolist()

    def drain_events(self) -> List[tuple]:
        ev = self._list(lib().kvo_drain_events)
        return [(e >> 1, e & 1) for e in ev]

    def snapshot(self) -> dict:
        s = self.scalars()
        return {
            "avail": self.available_size(),
            "pages": [s["free_pages"], s["inuse_pages"], s["reserved_pages"], s["total_pages"]],
            "nab": s["num_avail_blocks"],
            "ap": self.avail_pages(),
            "fp": self.full_pages(),
            "rb": self.reserved_blocks,
            "shrink": bool(s["in_shrink"]),
            "mapped": self.mapped_pages(),
        }


def group_indices_by_page(ids, block_mem_size: int, page_size: int = 2 << 20):
    """csrc/page_allocator.cpp:471-498 with libstdc++'s own iteration order."""
    a = np.ascontiguousarray(ids, dtype=np.int64)
    pages = np.empty(max(a.size, 1), dtype=np.int64)
    counts = np.empty(max(a.size, 1), dtype=np.int64)
    grouped = np.empty(max(a.size, 1), dtype=np.int64)
    d = lib().kvo_group_indices_by_page(_p64(a), a.size, block_mem_size, page_size,
                                        _p64(pages), _p64(counts), _p64(grouped))
    out, w = [], 0
    for j in range(d):
        out.append((int(pages[j]), grouped[w:w + counts[j]].tolist()))
        w += int(counts[j])
    return out


def manager_from_config(cfg: dict) -> OracleManager:
    return OracleManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], cfg["num_layers"],
                         reserve_null_block=cfg.get("reserve_null_block", False),
                         num_kv_buffers=cfg.get("num_kv_buffers", 2),
                         page_size=cfg.get("page_size", 2 << 20),
                         prealloc=cfg.get("prealloc", False),
                         min_reserved_pages=cfg.get("min_reserved_pages", 5),
                         max_reserved_pages=cfg.get("max_reserved_pages", 10))


def _ids(r):
    return r if isinstance(r, list) else []


def apply_op(mgr, op, results):
    """One trace op against any manager with the KVCacheManager method set."""
    kind = op[0]
    if kind == "alloc":
        try:
            return mgr.alloc(op[1])
        except RuntimeError:
            return "ERR:RuntimeError"
    if kind == "free_ids":
        mgr.free(list(op[1]))
    elif kind == "free_cat":
        mgr.free([b for k in op[1] for b in _ids(results[k])])
    elif kind == "free_slice":
        mgr.free(list(_ids(results[op[1]])[op[2]:op[3]]))
    elif kind == "resize":
        try:
            return bool(mgr.resize(op[1]))
        except AssertionError:
            return "ERR:AssertionError"
    elif kind == "trim":
        mgr.trim()
    elif kind == "try_reserve":
        try:
            return bool(mgr.try_to_reserve(op[1]))
        except RuntimeError:
            return "ERR:RuntimeError"
    elif kind == "free_reserved":
        mgr.free_reserved()
    elif kind == "avail":
        return mgr.available_size()
    elif kind == "clear":
        mgr.clear()
    elif kind == "step":
        pass
    else:
        raise ValueError(f"unknown op {op}")
    return None


def replay(trace: dict, snapshots: bool = True) -> dict:
    mgr = manager_from_config(trace["config"])
    results, records = [], []
    for op in trace["ops"]:
        r = apply_op(mgr, op, results)
        results.append(r)
        if snapshots:
            records.append({"r": r, **mgr.snapshot()})
    return {"records": records, "results": results, "null_block": mgr.null_block}
