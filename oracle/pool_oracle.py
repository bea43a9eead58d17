"""CPU restatement of the reference's prefix-cache pinning rules (ElasticBlockPool,
kvcached/integration/vllm/patches.py:308-614) over the oracle block manager (oracle/oracle.py).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and never by kvcached_b200/).  Pinned to the unmodified reference
by tests/golden/pool_*.json.gz (tests/test_pool_oracle_golden.py).

State, in the reference's terms:
  ref[b]        KVCacheBlock.ref_cnt
  key_of[b]     _block_id_to_key        by_key[key] = b   _cached_blocks
  lru           _evictable_blocks (an OrderedDict used as an insertion-ordered LRU; here a dict of ids)
Rules restated (line numbers of patches.py):
  :349        one real block is allocated up front as the null block and never freed
  :463-465    a full block that gets a NEW key is recorded in both maps; an existing key is left alone
  :467-481    eviction pops the oldest LRU entries, drops their keys, frees them with ONE manager.free call
  :483-517    get_new_blocks: refuse if n > available + len(lru); up to two rounds of
              [evict (n - available) if short] + manager.alloc(n); ref = 1 on success
  :519-533    touch: ref += 1, block leaves the LRU
  :546-568    free_blocks: ref -= 1 in list order; at 0 a keyed block joins the LRU, an unkeyed one is collected
              and freed in ONE call afterwards; then the LRU is cut back to max_cached_blocks (>= 0)
  :570-588    evict_blocks: keys dropped; members of the LRU are removed and freed (set iteration order)
  :590-603    reset: LRU freed in order, every key dropped
"""
from __future__ import annotations

from typing import Dict, List, Optional


class _Blk:
    __slots__ = ("block_id", "is_null")

    def __init__(self, i):
        self.block_id = i
        self.is_null = False


def _key(h, gid: int) -> bytes:
    if isinstance(h, str):
        h = h.encode()
    return bytes(h) + int(gid).to_bytes(4, "big")


class OraclePool:
    def __init__(self, mgr, num_blocks: int, enable_caching: bool = True, max_cached_blocks: int = 1000):
        self.m = mgr
        self.on = enable_caching
        self.cap = max_cached_blocks
        self.blk = [_Blk(i) for i in range(num_blocks)]
        self.ref = [0] * num_blocks
        got = mgr.alloc(1)
        self.null_block = self.blk[got[0]]
        self.null_block.is_null = True
        self.by_key: Dict[bytes, int] = {}
        self.key_of: Dict[int, bytes] = {}
        self.lru: Dict[int, None] = {}

    # -- helpers
    def _evict_oldest(self, k: int) -> None:
        victims = []
        for b in list(self.lru)[:max(k, 0)]:
            del self.lru[b]
            kk = self.key_of.pop(b, None)
            if kk is not None:
                self.by_key.pop(kk, None)
            victims.append(b)
        if victims:
            self.m.free(victims)

    # -- pool interface
    def get_cached_block(self, h, gids=None):
        if not self.on:
            return None
        if gids is None:
            b = self.by_key.get(_key(h, 0))
            return None if b is None else self.blk[b]
        if isinstance(gids, int):
            gids = [gids]
        out = []
        for g in gids:
            b = self.by_key.get(_key(h, g))
            if b is None:
                return None
            out.append(self.blk[b])
        return out or None

    def cache_full_blocks(self, request, blocks, num_cached, num_full, block_size=None, gid=0):
        if not self.on or num_cached >= num_full:
            return
        hs = request.block_hashes
        assert len(hs) >= num_full
        for i in range(num_cached, num_full):
            b = blocks[i]
            if b.is_null:
                continue
            kk = _key(hs[i], gid)
            if kk not in self.by_key:
                self.by_key[kk] = b.block_id
                self.key_of[b.block_id] = kk

    def get_num_free_blocks(self) -> int:
        return self.m.available_size() + (len(self.lru) if self.on else 0)

    def get_new_blocks(self, n: int):
        if n > self.get_num_free_blocks():
            raise ValueError("pool exhausted")
        ids = None
        for _ in (0, 1):
            if self.on:
                a = self.m.available_size()
                if a < n and self.lru:
                    self._evict_oldest(n - a)
            ids = self.m.alloc(n)
            if ids is not None:
                break
        if ids is None:
            raise ValueError("physical pool exhausted")
        for b in ids:
            self.ref[b] = 1
        return [self.blk[b] for b in ids]

    def touch(self, blocks) -> None:
        if not self.on:
            return
        groups = blocks if isinstance(blocks, tuple) else (blocks,)
        for g in groups:
            for b in g:
                self.ref[b.block_id] += 1
                self.lru.pop(b.block_id, None)

    def free_blocks(self, ordered) -> None:
        if not self.on:
            ids = [b.block_id for b in ordered if b is not None and not b.is_null]
            if ids:
                self.m.free(ids)
            return
        loose: List[int] = []
        for b in ordered:
            if b is None or b.is_null:
                continue
            i = b.block_id
            self.ref[i] -= 1
            if self.ref[i] == 0:
                if i in self.key_of:
                    self.lru[i] = None
                else:
                    loose.append(i)
        if loose:
            self.m.free(loose)
        if self.cap >= 0 and len(self.lru) > self.cap:
            self._evict_oldest(len(self.lru) - self.cap)

    def evict_blocks(self, ids: set) -> None:
        if not self.on:
            return
        out = []
        for i in ids:
            kk = self.key_of.pop(i, None)
            if kk is not None:
                self.by_key.pop(kk, None)
            if i in self.lru:
                del self.lru[i]
                out.append(i)
        if out:
            self.m.free(out)

    def reset_prefix_cache(self) -> bool:
        if not self.on:
            return True
        if self.lru:
            ids = list(self.lru)
            self.lru.clear()
            self.m.free(ids)
        self.by_key.clear()
        self.key_of.clear()
        return True

    # -- what the fixtures record after every event
    def snap(self) -> dict:
        sc = self.m.scalars()
        refs = [(i, c) for i, c in enumerate(self.ref) if c]
        return {"free": self.get_num_free_blocks(), "evictable": list(self.lru), "ncached": len(self.by_key),
                "nkeyed": len(self.key_of), "nref": len(refs),
                "refsum": sum((i + 1) * c for i, c in refs) % (1 << 61),
                "avail": self.m.available_size(),
                "pages": [sc["free_pages"], sc["inuse_pages"], sc["reserved_pages"]],
                "ap": self.m.avail_pages()}
