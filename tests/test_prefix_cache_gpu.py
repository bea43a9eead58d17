"""The behaviours the reference's own prefix-cache unit tests pin (tests/test_prefix_cache.py: 28 cases, same names
here), run against kvcached_b200.block_pool.ElasticBlockPool on a real device manager instead of the reference pool on
a mock manager.  The scenarios are restated from the behaviours those cases describe; expectations are the
reference's (lazy eviction, LRU order, cap, explicit eviction, reset, caching disabled, edge cases)."""
import numpy as np
import pytest

from kvcached_b200.block_pool import ElasticBlockPool
from kvcached_b200.kv_cache_manager import KVCacheManager
from kvcached_b200.pool_trace import Request

pytestmark = pytest.mark.gpu
BPP = 128                       # blocks per 2 MiB page at block 16 x cell 1024


def make_pool(pages=8, enable_caching=True, cap=-1):
    n = pages * BPP
    mgr = KVCacheManager(n, 16, 1024, 1, page_size=2 << 20, attach_vmm=False, max_cached_blocks=cap)
    mgr.set_phys_free_pages(1 << 40)
    return ElasticBlockPool(n, 16, 1024, 1, enable_caching, max_cached_blocks=cap, kv_cache_manager=mgr), mgr


@pytest.fixture
def pool_and_manager():
    pool, mgr = make_pool()
    yield pool, mgr
    mgr.close()


@pytest.fixture
def pool_factory():
    made = []

    def _make(pages=1, enable_caching=True, cap=-1):
        pool, mgr = make_pool(pages, enable_caching, cap)
        made.append(mgr)
        return pool, mgr
    yield _make
    for m in made:
        m.close()


def run_request(pool, hashes):
    """alloc -> cache every block -> return the blocks (a request that filled len(hashes) blocks)."""
    blocks = pool.get_new_blocks(len(hashes))
    pool.cache_full_blocks(Request(0, list(hashes)), blocks, 0, len(hashes), 16, 0)
    return blocks


def H(prefix, n):
    return [f"{prefix}{i}" for i in range(n)]


class TestLazyEviction:
    def test_cross_request_cache_hit(self, pool_and_manager):
        pool, _ = pool_and_manager
        hs = H("h", 4)
        pool.free_blocks(run_request(pool, hs))
        assert len(pool._evictable_blocks) == 4 and len(pool._cached_blocks) == 4
        assert all(pool.get_cached_block(h, [0]) is not None for h in hs)

    def test_evictable_blocks_not_freed_to_kvcached(self, pool_and_manager):
        pool, mgr = pool_and_manager
        a0 = mgr.available_size()
        blocks = run_request(pool, H("h", 2))
        assert mgr.available_size() == a0 - 2
        pool.free_blocks(blocks)
        assert mgr.available_size() == a0 - 2            # still held: pages stay mapped

    def test_uncached_blocks_freed_immediately(self, pool_and_manager):
        pool, mgr = pool_and_manager
        a0 = mgr.available_size()
        blocks = pool.get_new_blocks(3)
        assert mgr.available_size() == a0 - 3
        pool.free_blocks(blocks)
        assert mgr.available_size() == a0 and len(pool._evictable_blocks) == 0

    def test_touch_removes_from_evictable_pool(self, pool_and_manager):
        pool, _ = pool_and_manager
        blocks = run_request(pool, H("h", 3))
        pool.free_blocks(blocks)
        assert len(pool._evictable_blocks) == 3
        pool.touch([blocks[0], blocks[2]])
        assert list(pool._evictable_blocks) == [blocks[1].block_id]
        assert blocks[0].ref_cnt == 1 and blocks[1].ref_cnt == 0 and blocks[2].ref_cnt == 1

    def test_touch_tuple_form(self, pool_and_manager):
        pool, _ = pool_and_manager
        blocks = run_request(pool, H("h", 2))
        pool.free_blocks(blocks)
        pool.touch(([blocks[0]], [blocks[1]]))
        assert len(pool._evictable_blocks) == 0 and [b.ref_cnt for b in blocks] == [1, 1]

    def test_ref_cnt_shared_prefix(self, pool_and_manager):
        pool, mgr = pool_and_manager
        a0 = mgr.available_size()
        shared = run_request(pool, H("p", 2))               # request A holds the prefix
        pool.touch(shared)                                  # request B hits it
        assert [b.ref_cnt for b in shared] == [2, 2]
        pool.free_blocks(shared)                            # A finishes: still referenced by B
        assert len(pool._evictable_blocks) == 0 and [b.ref_cnt for b in shared] == [1, 1]
        pool.free_blocks(shared)                            # B finishes
        assert list(pool._evictable_blocks) == [b.block_id for b in shared]
        assert mgr.available_size() == a0 - 2


class TestEviction:
    def test_evict_on_alloc_pressure(self, pool_factory):
        pool, mgr = pool_factory(pages=1)
        pool.free_blocks(run_request(pool, H("a", 100)))
        assert len(pool._evictable_blocks) == 100 and mgr.available_size() == 27
        got = pool.get_new_blocks(40)                       # 13 blocks must be evicted to make room
        assert len(got) == 40 and len(pool._evictable_blocks) == 87 and mgr.available_size() == 0

    def test_evict_lru_order(self, pool_factory):
        pool, _ = pool_factory(pages=1)
        a = run_request(pool, H("a", 50))
        b = run_request(pool, H("b", 50))
        pool.free_blocks(a)                                 # a is older than b in the LRU
        pool.free_blocks(b)
        pool.get_new_blocks(27 + 10)                        # evicts the 10 oldest: the head of a
        assert all(pool.get_cached_block(h, [0]) is None for h in H("a", 10))
        assert all(pool.get_cached_block(h, [0]) is not None for h in H("a", 50)[10:] + H("b", 50))

    def test_evict_all_then_alloc(self, pool_factory):
        pool, _ = pool_factory(pages=1)
        pool.free_blocks(run_request(pool, H("a", 127)))
        assert pool.get_num_free_blocks() == 127
        got = pool.get_new_blocks(127)
        assert len(got) == 127 and len(pool._evictable_blocks) == 0 and len(pool._cached_blocks) == 0

    def test_partial_eviction(self, pool_factory):
        pool, _ = pool_factory(pages=1)
        pool.free_blocks(run_request(pool, H("a", 127)))
        pool.get_new_blocks(5)
        assert len(pool._evictable_blocks) == 122 and len(pool._cached_blocks) == 122

    def test_no_eviction_when_kvcached_has_space(self, pool_factory):
        pool, _ = pool_factory(pages=2)
        pool.free_blocks(run_request(pool, H("a", 20)))
        pool.get_new_blocks(30)
        assert len(pool._evictable_blocks) == 20

    def test_get_num_free_blocks_includes_evictable(self, pool_factory):
        pool, mgr = pool_factory(pages=1)
        blocks = run_request(pool, H("a", 30))
        assert pool.get_num_free_blocks() == 97
        pool.free_blocks(blocks)
        assert pool.get_num_free_blocks() == 127 and mgr.available_size() == 97


class TestResetAndExplicitEvict:
    def test_reset_frees_evictable_blocks(self, pool_factory):
        pool, mgr = pool_factory(pages=1)
        a0 = mgr.available_size()
        pool.free_blocks(run_request(pool, H("a", 10)))
        assert pool.reset_prefix_cache() is True
        assert mgr.available_size() == a0 and not pool._evictable_blocks and not pool._cached_blocks

    def test_reset_with_no_evictable(self, pool_and_manager):
        pool, _ = pool_and_manager
        assert pool.reset_prefix_cache() is True

    def test_evict_blocks_explicit(self, pool_factory):
        pool, mgr = pool_factory(pages=1)
        a0 = mgr.available_size()
        blocks = run_request(pool, H("a", 4))
        pool.free_blocks(blocks)
        pool.evict_blocks({blocks[1].block_id, blocks[3].block_id})
        assert sorted(pool._evictable_blocks) == sorted([blocks[0].block_id, blocks[2].block_id])
        assert pool.get_cached_block("a1", [0]) is None and pool.get_cached_block("a0", [0]) is not None
        assert mgr.available_size() == a0 - 2

    def test_evict_blocks_active_not_freed(self, pool_factory):
        pool, mgr = pool_factory(pages=1)
        blocks = run_request(pool, H("a", 3))               # still referenced
        a1 = mgr.available_size()
        pool.evict_blocks({b.block_id for b in blocks})
        assert mgr.available_size() == a1 and len(pool._cached_blocks) == 0
        pool.free_blocks(blocks)                            # keys are gone: freed at once
        assert mgr.available_size() == a1 + 3 and not pool._evictable_blocks


class TestCachingDisabled:
    def test_no_cache_free_immediately(self, pool_factory):
        pool, mgr = pool_factory(pages=1, enable_caching=False)
        a0 = mgr.available_size()
        blocks = run_request(pool, H("a", 5))
        pool.free_blocks(blocks)
        assert mgr.available_size() == a0

    def test_no_cache_hit(self, pool_factory):
        pool, _ = pool_factory(pages=1, enable_caching=False)
        pool.free_blocks(run_request(pool, H("a", 3)))
        assert pool.get_cached_block("a0", [0]) is None

    def test_get_num_free_blocks_no_cache(self, pool_factory):
        pool, mgr = pool_factory(pages=1, enable_caching=False)
        run_request(pool, H("a", 7))
        assert pool.get_num_free_blocks() == mgr.available_size() == 120


class TestEdges:
    def test_get_new_blocks_raises_when_alloc_returns_none(self, pool_and_manager):
        pool, _ = pool_and_manager
        with pytest.raises(ValueError):
            pool.get_new_blocks(8 * BPP)                    # one more than the pool can ever give

    def test_get_new_blocks_retries_after_eviction_when_alloc_returns_none(self, pool_factory):
        pool, mgr = pool_factory(pages=1)
        pool.free_blocks(run_request(pool, H("a", 127)))    # everything evictable, nothing available
        assert mgr.available_size() == 0
        assert len(pool.get_new_blocks(64)) == 64

    def test_free_none_blocks(self, pool_and_manager):
        pool, mgr = pool_and_manager
        a0 = mgr.available_size()
        pool.free_blocks([None, pool.null_block, None])
        assert mgr.available_size() == a0

    def test_evict_more_than_pool_size(self, pool_factory):
        pool, mgr = pool_factory(pages=1, cap=0)            # cap 0: everything is evicted on free
        a0 = mgr.available_size()
        pool.free_blocks(run_request(pool, H("a", 6)))
        assert not pool._evictable_blocks and mgr.available_size() == a0

    def test_evict_empty_pool(self, pool_and_manager):
        pool, _ = pool_and_manager
        pool.evict_blocks({5, 6, 7})                        # nothing cached: no-op
        assert not pool._evictable_blocks

    def test_cache_full_blocks_idempotent(self, pool_and_manager):
        pool, _ = pool_and_manager
        blocks = pool.get_new_blocks(3)
        req = Request(0, H("x", 3))
        pool.cache_full_blocks(req, blocks, 0, 3, 16, 0)
        pool.cache_full_blocks(req, blocks, 0, 3, 16, 0)
        pool.cache_full_blocks(req, blocks, num_cached_blocks=0, num_full_blocks=3, block_size=16, kv_cache_group_id=0)
        assert len(pool._cached_blocks) == 3 and len(pool.keyed_ids()) == 3

    def test_reuse_after_eviction_and_realloc(self, pool_factory):
        pool, _ = pool_factory(pages=1)
        pool.free_blocks(run_request(pool, H("a", 127)))
        again = pool.get_new_blocks(127)                    # evicts all of a, reuses the ids
        assert pool.get_cached_block("a0", [0]) is None
        pool.cache_full_blocks(Request(1, H("b", 127)), again, 0, 127, 16, 0)
        assert pool.get_cached_block("b5", [0])[0] is again[5]

    def test_mixed_cached_and_uncached_free(self, pool_and_manager):
        pool, mgr = pool_and_manager
        a0 = mgr.available_size()
        blocks = pool.get_new_blocks(5)
        pool.cache_full_blocks(Request(0, H("m", 3)), blocks, 0, 3, 16, 0)     # last two stay unhashed
        pool.free_blocks(blocks)
        assert list(pool._evictable_blocks) == [b.block_id for b in blocks[:3]] and mgr.available_size() == a0 - 3

    def test_get_usage(self, pool_factory):
        pool, _ = pool_factory(pages=1)
        assert pool.get_usage() == pytest.approx(1 - 127 / 128)
        run_request(pool, H("a", 32))
        assert pool.get_usage() == pytest.approx(1 - 95 / 128)


def test_cap_keeps_the_newest_blocks():
    pool, mgr = make_pool(pages=2, cap=10)
    try:
        pool.free_blocks(run_request(pool, H("a", 8)))
        pool.free_blocks(run_request(pool, H("b", 8)))     # 16 evictable > cap 10: the 6 oldest go
        ev = list(pool._evictable_blocks)
        assert len(ev) == 10 and all(pool.get_cached_block(h, [0]) is None for h in H("a", 6))
        assert pool.get_cached_block("a6", [0]) is not None and pool.get_cached_block("b7", [0]) is not None
    finally:
        mgr.close()
