"""ElasticBlockPool — prefix-cache pinning on top of the device-resident block manager.

Mirrors the class the reference injects into vLLM's block-pool module
(kvcached/integration/vllm/patches.py:308-614): same constructor arguments, same method names, argument
meaning and error behaviour (`get_new_blocks` raises ValueError, `cache_full_blocks` accepts every vLLM call
signature the reference accepts).  What differs is where the state lives:

  device (libkvb200.so, KVB_OP_PIN_* requests of the step kernel)
      ref_cnt of every block, the "has a cache key" flag, the insertion-ordered LRU of evictable blocks,
      eviction (oldest first) fused with kv_cache_manager.free / alloc in the same launch
  host (this file)
      the hash -> block dictionary: hashes are host data and lookups are the scheduler's; evicted ids come
      back in the request's result so their keys can be dropped.

There is no Python bookkeeping of ref counts or LRU order here: `KVCacheBlock.ref_cnt` reads the device.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional

import numpy as np

from ._lib import (OP_PIN_EVICT, OP_PIN_FREE, OP_PIN_MARK, OP_PIN_NEW, OP_PIN_RESET, OP_PIN_TOUCH, ST_NO_PAGES,
                   ST_NONE, ST_OK)
from .kv_cache_manager import OP_DTYPE, KVCacheManager


def _make_cache_key(block_hash: Any, group_id: int) -> bytes:
    """hash ‖ be32(group) — patches.py:260-274 (str hashes are encoded so both spellings give one key)."""
    if isinstance(block_hash, str):
        block_hash = block_hash.encode()
    return bytes(block_hash) + group_id.to_bytes(4, "big", signed=False)


def get_max_cached_blocks(block_size: int, max_cached_tokens: Optional[int] = None) -> int:
    """MAX_CACHED_TOKENS // block_size; -1 unlimited, 0 evict on free — patches.py:247-257."""
    import os
    if max_cached_tokens is None:
        max_cached_tokens = int(os.getenv("KVCACHED_MAX_CACHED_TOKENS", "16000"))
    if max_cached_tokens < 0:
        return -1
    return max_cached_tokens // block_size


class KVCacheBlock:
    """vLLM's KVCacheBlock as far as the pool is concerned: block_id, is_null, ref_cnt (device-resident)."""
    __slots__ = ("block_id", "is_null", "_pool")

    def __init__(self, block_id: int, pool: "ElasticBlockPool"):
        self.block_id = block_id
        self.is_null = False
        self._pool = pool

    @property
    def ref_cnt(self) -> int:
        return self._pool.ref_counts().get(self.block_id, 0)

    def __repr__(self) -> str:
        return f"KVCacheBlock({self.block_id})"


class ElasticBlockPool:
    def __init__(self, num_gpu_blocks: int, block_size: int, cell_size: int, num_layers: int, enable_caching: bool,
                 enable_kv_cache_events: bool = False, num_kv_buffers: int = 2, max_cached_blocks: int = 1000,
                 *, kv_cache_manager: Optional[KVCacheManager] = None, **manager_kwargs) -> None:
        assert isinstance(num_gpu_blocks, int) and num_gpu_blocks > 0
        assert not enable_kv_cache_events, "KV cache events are not supported in ElasticBlockPool"
        self.enable_prefix_cache = enable_caching
        self.max_cached_blocks = max_cached_blocks
        self.num_gpu_blocks = num_gpu_blocks
        self.enable_kv_cache_events = enable_kv_cache_events
        self.kv_event_queue: list = []
        self.kv_block_pool = [KVCacheBlock(i, self) for i in range(num_gpu_blocks)]
        if kv_cache_manager is None:
            kv_cache_manager = KVCacheManager(num_gpu_blocks, block_size, cell_size, num_layers,
                                              num_kv_buffers=num_kv_buffers, max_cached_blocks=max_cached_blocks,
                                              **manager_kwargs)
        self.kv_cache_manager = kv_cache_manager
        # the null block: a real block, so its id is valid on the GPU (patches.py:341-352)
        ids = self.kv_cache_manager.alloc(1)
        assert ids is not None and len(ids) == 1
        self.null_block = self.kv_block_pool[ids[0]]
        self.null_block.is_null = True
        self._cached_blocks: Dict[bytes, KVCacheBlock] = {}
        self._block_id_to_key: Dict[int, bytes] = {}

    # ------------------------------------------------------------------ one request = one launch
    def _run(self, kind: int, n: int, ids: Optional[np.ndarray] = None, out_cap: int = 0):
        ops = np.zeros(1, dtype=OP_DTYPE)
        ops["kind"][0] = kind
        ops["n"][0] = n
        res, out = self.kv_cache_manager.run_ops(ops, ids, out_cap=out_cap)
        return res[0], out

    def _drop_keys(self, evicted: Iterable[int]) -> None:
        for bid in evicted:
            key = self._block_id_to_key.pop(int(bid), None)
            if key is not None:
                self._cached_blocks.pop(key, None)

    @staticmethod
    def _ids(blocks: Iterable[Any]) -> np.ndarray:
        return np.fromiter((b.block_id for b in blocks if b is not None and not getattr(b, "is_null", False)),
                           dtype=np.int64)

    # ------------------------------------------------------------------ the reference's method set
    def get_cached_block(self, block_hash: Any, kv_cache_group_ids: Optional[Iterable[int]] = None):
        """patches.py:365-394"""
        if not self.enable_prefix_cache:
            return None
        if kv_cache_group_ids is None:
            return self._cached_blocks.get(_make_cache_key(block_hash, 0))
        if isinstance(kv_cache_group_ids, int):
            kv_cache_group_ids = [int(kv_cache_group_ids)]
        out: List[KVCacheBlock] = []
        for gid in kv_cache_group_ids:
            b = self._cached_blocks.get(_make_cache_key(block_hash, int(gid)))
            if b is None:
                return None          # all groups must hit
            out.append(b)
        return out or None

    def cache_full_blocks(self, request: Any, blocks: list, *args: Any, **kwargs: Any) -> None:
        """patches.py:396-465 — every call signature the reference accepts."""
        if not self.enable_prefix_cache:
            return
        block_hashes = kwargs.pop("block_hashes", None)
        num_cached_blocks = kwargs.pop("num_cached_blocks", None)
        num_full_blocks = kwargs.pop("num_full_blocks", None)
        _block_size = kwargs.pop("block_size", None)
        kv_cache_group_id = kwargs.pop("kv_cache_group_id", 0)
        kwargs.pop("hash_fn", None)
        rest = list(args)
        if block_hashes is None and rest and isinstance(rest[0], (list, tuple)):
            block_hashes = rest.pop(0)
        if num_cached_blocks is None and rest:
            num_cached_blocks = rest.pop(0)
        if num_full_blocks is None and rest:
            num_full_blocks = rest.pop(0)
        if _block_size is None and rest:
            _block_size = rest.pop(0)
        if rest and isinstance(rest[0], int):
            kv_cache_group_id = rest.pop(0)
        if num_cached_blocks is None or num_full_blocks is None:
            raise TypeError("cache_full_blocks requires num_cached_blocks and num_full_blocks")
        num_cached_blocks, num_full_blocks = int(num_cached_blocks), int(num_full_blocks)
        kv_cache_group_id = int(kv_cache_group_id)
        if num_cached_blocks >= num_full_blocks:
            return
        if block_hashes is None:
            assert hasattr(request, "block_hashes"), "Request missing block_hashes attribute"
            block_hashes = request.block_hashes
        assert len(block_hashes) >= num_full_blocks, \
            f"Request has {len(block_hashes)} hashes but need {num_full_blocks}"
        newly: List[int] = []
        for i, block in enumerate(blocks[num_cached_blocks:num_full_blocks]):
            if getattr(block, "is_null", False):
                continue
            key = _make_cache_key(block_hashes[num_cached_blocks + i], kv_cache_group_id)
            if key in self._cached_blocks:
                continue             # idempotent
            self._cached_blocks[key] = block
            self._block_id_to_key[block.block_id] = key
            newly.append(block.block_id)
        if newly:
            self._run(OP_PIN_MARK, len(newly), np.asarray(newly, dtype=np.int64))

    def get_new_blocks(self, num_blocks: int) -> list:
        """patches.py:483-517: evict-if-short, alloc and ref_cnt = 1 in ONE launch."""
        # caching off: the LRU is empty, so the same request is a plain alloc that also sets ref_cnt = 1 (the
        # reference sets it in both modes, patches.py:511-516)
        r, out = self._run(OP_PIN_NEW, num_blocks, out_cap=2 * max(num_blocks, 0) + 1)
        nev = int(r["aux"])
        st = int(r["status"])
        if st == ST_OK:
            off = int(r["off"])
            self._drop_keys(out[off - nev:off].tolist())
            return [self.kv_block_pool[b] for b in out[off:off + int(r["n"])].tolist()]
        if nev:
            self._drop_keys(out[int(r["off"]) - nev:int(r["off"])].tolist())
        if st == ST_NO_PAGES:
            raise RuntimeError("No free pages left")
        if st == ST_NONE and nev == 0 and num_blocks > self.get_num_free_blocks():
            raise ValueError(f"Cannot get {num_blocks} free blocks from the pool")
        raise ValueError("Unable to allocate KV cache blocks from physical pool; "
                         f"requested={num_blocks}, available={self.kv_cache_manager.available_size()}")

    def touch(self, blocks) -> None:
        """patches.py:519-533"""
        if not self.enable_prefix_cache:
            return
        if isinstance(blocks, tuple):
            ids = np.fromiter((b.block_id for bl in blocks for b in bl), dtype=np.int64)
        else:
            ids = np.fromiter((b.block_id for b in blocks), dtype=np.int64)
        if ids.size:
            self._run(OP_PIN_TOUCH, ids.size, ids)

    def free_blocks(self, ordered_blocks: Iterable[Any]) -> None:
        """patches.py:535-568"""
        ids = self._ids(ordered_blocks)
        if not self.enable_prefix_cache:
            if ids.size:
                self.kv_cache_manager.free(ids)
            return
        if ids.size == 0:
            return               # nothing can have become evictable: the cap already holds
        r, out = self._run(OP_PIN_FREE, ids.size, ids, out_cap=int(ids.size) + 1)
        nev = int(r["aux"])
        if nev:
            self._drop_keys(out[int(r["off"]):int(r["off"]) + nev].tolist())

    def evict_blocks(self, block_ids: set) -> None:
        """patches.py:570-588 — the set is walked in CPython's iteration order, like the reference does."""
        if not self.enable_prefix_cache:
            return
        ids = list(block_ids)
        for bid in ids:
            key = self._block_id_to_key.pop(bid, None)
            if key is not None:
                self._cached_blocks.pop(key, None)
        if ids:
            self._run(OP_PIN_EVICT, len(ids), np.asarray(ids, dtype=np.int64))

    def reset_prefix_cache(self) -> bool:
        """patches.py:590-603"""
        if not self.enable_prefix_cache:
            return True
        self._run(OP_PIN_RESET, 0)
        self._cached_blocks.clear()
        self._block_id_to_key.clear()
        return True

    def get_num_free_blocks(self) -> int:
        """patches.py:605-606"""
        a = self.kv_cache_manager.available_size()
        return a + self.kv_cache_manager.stats().num_evictable_blocks if self.enable_prefix_cache else a

    def get_usage(self) -> float:
        return 1.0 - (self.get_num_free_blocks() / self.num_gpu_blocks)

    def take_events(self) -> list:
        return []

    # ------------------------------------------------------------------ inspection (tests)
    @property
    def _evictable_blocks(self):
        """Read-only view with the reference attribute's shape (OrderedDict block_id -> block, oldest first),
        rebuilt from the device LRU on every access — for tests and debugging, not for the serving path."""
        from collections import OrderedDict
        return OrderedDict((b, self.kv_block_pool[b]) for b in self.evictable_ids())

    def evictable_ids(self) -> List[int]:
        return self.kv_cache_manager._dump(7)

    def keyed_ids(self) -> List[int]:
        return self.kv_cache_manager._dump(8)

    def ref_counts(self) -> Dict[int, int]:
        v = self.kv_cache_manager._dump(9)
        return dict(zip(v[0::2], v[1::2]))
