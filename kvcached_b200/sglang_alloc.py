"""ElasticPagedTokenToKVPoolAllocator — SGLang's paged token-to-KV-pool allocator on the device-resident
block manager.  Mirrors the class the reference injects into sglang.srt.mem_cache.allocator
(kvcached/integration/sglang/patches.py:143-297): `available_size`, `alloc`, `alloc_extend`, `alloc_decode`,
`free`, `clear`, with the same arguments and the same `None` on exhaustion.

The reference counts new pages on CPU copies of the lengths, calls kv_cache_manager.alloc, ships the ids to the
GPU with torch.tensor(list) and expands them with SGLang's Triton kernels.  Here alloc_extend / alloc_decode /
free are ONE launch of the step kernel each (kvb_mgr_alloc_extend / _alloc_decode / _free_slots): lengths and
slot indices never leave the device.  torch is used for device memory and the current stream only.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

from . import _lib
from ._lib import ST_NO_PAGES, ST_NONE, KvbError
from .kv_cache_manager import KVCacheManager


class ElasticTokenToKVPoolAllocator:
    """SGLang's token-level allocator (page_size 1) — the class `inject_elastic_allocator` builds
    (kvcached/integration/sglang/patches.py:60-118): alloc(need_size) -> int64 device tensor of slot ids (or None),
    free(free_index), available_size(), clear(), free-group batching.  Ids go kernel -> device tensor and back without
    ever becoming a Python list."""

    def __init__(self, size: int, dtype, device: str, kvcache, *args, **kwargs) -> None:
        import torch
        self._torch = torch
        mgr = getattr(kvcache, "kvcached_allocator", kvcache)
        if not isinstance(mgr, KVCacheManager):
            raise ValueError("ElasticTokenToKVPoolAllocator requires an elastic KV pool")
        self.size = size
        self.page_size = 1
        self.dtype = dtype
        self.device = device
        self.kvcached_allocator = mgr
        self.is_not_in_free_group = True
        self.free_group: List = []
        self._lib = _lib.load()

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream().cuda_stream)

    def available_size(self) -> int:
        return min(self.kvcached_allocator.available_size(), self.size)      # patches.py:92-99

    def alloc(self, need_size: int):
        torch = self._torch
        out = torch.empty((need_size,), dtype=torch.int64, device=self.device)
        if need_size == 0:
            return out
        m = self.kvcached_allocator
        m._wait_post_init()
        rc = self._lib.kvb_mgr_alloc_device(m._h, need_size, out.data_ptr(), self._stream())
        if rc == need_size:
            return out
        if rc == ST_NONE:
            return None
        if rc == ST_NO_PAGES:
            raise RuntimeError("No free pages left")
        raise KvbError(f"alloc_device failed ({rc}): {_lib.last_error()}")

    def free(self, free_index) -> None:
        if free_index.numel() == 0:
            return
        if not self.is_not_in_free_group:
            self.free_group.append(free_index)
            return
        idx = free_index if free_index.dtype == self._torch.int64 and free_index.is_contiguous() else \
            free_index.to(self._torch.int64).contiguous()
        m = self.kvcached_allocator
        m._wait_post_init()
        rc = self._lib.kvb_mgr_free_device(m._h, idx.data_ptr(), idx.numel(), self._stream())
        if rc < 0:
            raise KvbError(f"free_device failed ({rc}): {_lib.last_error()}")

    def free_group_begin(self) -> None:
        self.is_not_in_free_group = False
        self.free_group = []

    def free_group_end(self) -> None:
        self.is_not_in_free_group = True
        if self.free_group:
            self.free(self._torch.cat(self.free_group))

    def clear(self) -> None:
        self.kvcached_allocator.clear()
        self.is_not_in_free_group = True
        self.free_group = []


class ElasticPagedTokenToKVPoolAllocator:
    def __init__(self, size: int, page_size: int, dtype, device: str, kvcache, *args, **kwargs) -> None:
        import torch
        self._torch = torch
        mgr = getattr(kvcache, "kvcached_allocator", kvcache)
        if not isinstance(mgr, KVCacheManager):
            raise ValueError("ElasticPagedTokenToKVPoolAllocator requires elastic MHA pool")   # patches.py:156-159
        if not str(device).startswith("cuda"):
            raise ValueError("ElasticPagedTokenToKVPoolAllocator only supports GPU devices (cuda/hip)")
        if mgr.block_size != page_size:
            raise ValueError(f"page_size {page_size} != the manager's block size {mgr.block_size}")
        self.size = size
        self.page_size = page_size
        self.dtype = dtype
        self.device = device
        self.kvcached_allocator = mgr
        self.num_pages = size // page_size
        self.is_not_in_free_group = True
        self.free_group: List = []
        self.free_pages = torch.empty((0,), dtype=torch.int64, device=device)
        self.release_pages = torch.empty((0,), dtype=torch.int64, device=device)
        self._lib = _lib.load()

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream().cuda_stream)

    def available_size(self) -> int:
        return self.kvcached_allocator.available_size() * self.page_size

    def alloc(self, need_size: int):
        """patches.py:176-186"""
        torch = self._torch
        ids = self.kvcached_allocator.alloc_np(need_size // self.page_size)
        if ids is None:
            return None
        page_ids = torch.from_numpy(ids).to(self.device)
        return (page_ids[:, None] * self.page_size + torch.arange(self.page_size, device=self.device)).reshape(-1)

    def alloc_extend(self, prefix_lens, prefix_lens_cpu, seq_lens, seq_lens_cpu, last_loc, extend_num_tokens: int):
        """patches.py:188-236.  The *_cpu copies are not needed: pages are counted on the device."""
        torch = self._torch
        bs = len(prefix_lens)
        out = torch.empty((extend_num_tokens,), dtype=torch.int64, device=self.device)
        if bs == 0:
            return out
        n_pages = C.c_int64(0)
        m = self.kvcached_allocator
        m._wait_post_init()
        prefix_lens, seq_lens, last_loc = self._i64(prefix_lens), self._i64(seq_lens), self._i64(last_loc)
        rc = self._lib.kvb_mgr_alloc_extend(m._h, bs, prefix_lens.data_ptr(), seq_lens.data_ptr(), last_loc.data_ptr(),
                                            extend_num_tokens, out.data_ptr(), self._stream(), C.byref(n_pages))
        return self._result(rc, out)

    def alloc_decode(self, seq_lens, seq_lens_cpu, last_loc):
        """patches.py:238-272"""
        torch = self._torch
        bs = len(seq_lens)
        out = torch.empty((bs,), dtype=torch.int64, device=self.device)
        if bs == 0:
            return out
        n_pages = C.c_int64(0)
        m = self.kvcached_allocator
        m._wait_post_init()
        seq_lens, last_loc = self._i64(seq_lens), self._i64(last_loc)
        rc = self._lib.kvb_mgr_alloc_decode(m._h, bs, seq_lens.data_ptr(), last_loc.data_ptr(), out.data_ptr(),
                                            self._stream(), C.byref(n_pages))
        return self._result(rc, out)

    def _i64(self, t):
        """The kernel reads contiguous int64 on the allocator's device; SGLang versions differ (int32 lengths,
        views).  The reference's Triton kernels are dtype-agnostic, so normalise instead of asserting."""
        torch = self._torch
        if t.dtype != torch.int64 or not t.is_contiguous() or t.device != torch.device(self.device):
            t = t.to(device=self.device, dtype=torch.int64).contiguous()
        return t

    @staticmethod
    def _result(rc: int, out):
        if rc == 1:
            return out
        if rc == ST_NONE:
            return None
        if rc == ST_NO_PAGES:
            raise RuntimeError("No free pages left")
        raise KvbError(f"slot allocation failed ({rc}): {_lib.last_error()}")

    def free(self, free_index) -> None:
        """patches.py:274-288: kv_cache_manager.free(torch.unique(free_index // page_size)), on the device."""
        if free_index.numel() == 0:
            return
        if not self.is_not_in_free_group:
            self.free_group.append(free_index)
            return
        idx = free_index if free_index.dtype == self._torch.int64 and free_index.is_contiguous() else \
            free_index.to(self._torch.int64).contiguous()
        m = self.kvcached_allocator
        m._wait_post_init()
        rc = self._lib.kvb_mgr_free_slots(m._h, idx.data_ptr(), idx.numel(), self._stream())
        if rc < 0:
            raise KvbError(f"free_slots failed ({rc}): {_lib.last_error()}")

    def free_group_begin(self) -> None:
        self.is_not_in_free_group = False
        self.free_group = []

    def free_group_end(self) -> None:
        self.is_not_in_free_group = True
        if self.free_group:
            self.free(self._torch.cat(self.free_group))

    def clear(self) -> None:
        self.kvcached_allocator.clear()
        self.free_pages = self._torch.empty((0,), dtype=self._torch.int64, device=self.device)
        self.is_not_in_free_group = True
        self.free_group = []
