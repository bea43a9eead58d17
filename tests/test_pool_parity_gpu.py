"""Prefix-cache pinning on the device (kvcached_b200.block_pool.ElasticBlockPool, KVB_OP_PIN_* requests of the
step kernel) against (1) the fixtures produced by the UNMODIFIED reference ElasticBlockPool and (2) the oracle
restatement on fresh seeds — after EVERY event: returned block ids, get_num_free_blocks, LRU order, key counts,
ref counts, available_size, page counters, avail-page order.  Bit-exact (integers)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from test_pool_oracle_golden import (POOL_FIELDS, compare_records, load_pool_golden, oracle_pool_from,  # noqa: E402
                                     pool_golden_names)

from kvcached_b200 import pool_trace as PT  # noqa: E402
from kvcached_b200.block_pool import ElasticBlockPool  # noqa: E402
from kvcached_b200.kv_cache_manager import KVCacheManager  # noqa: E402

pytestmark = pytest.mark.gpu


def our_pool_from(fix):
    cfg = fix["config"]
    mgr = KVCacheManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], cfg["num_layers"],
                         num_kv_buffers=cfg["num_kv_buffers"], page_size=2 << 20, attach_vmm=False,
                         max_cached_blocks=cfg["max_cached_blocks"])
    mgr.set_phys_free_pages(1 << 40)
    return ElasticBlockPool(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], cfg["num_layers"],
                            cfg.get("enable_caching", True), num_kv_buffers=cfg["num_kv_buffers"],
                            max_cached_blocks=cfg["max_cached_blocks"], kv_cache_manager=mgr)


def our_snap(p):
    m = p.kv_cache_manager
    st = m.stats()
    refs = sorted(p.ref_counts().items())
    keyed = p.keyed_ids()
    assert keyed == sorted(p._block_id_to_key), "device key flags must mirror _block_id_to_key"
    return {"free": p.get_num_free_blocks(), "evictable": p.evictable_ids(), "ncached": len(p._cached_blocks),
            "nkeyed": len(keyed), "nref": len(refs), "refsum": sum((i + 1) * c for i, c in refs) % (1 << 61),
            "avail": m.available_size(), "pages": [st.num_free_pages, st.num_inuse_pages, st.num_reserved_pages],
            "ap": m.avail_pages()}


@pytest.mark.parametrize("name", pool_golden_names())
def test_device_pool_matches_reference_fixture(name):
    fix = load_pool_golden(name)
    pool = our_pool_from(fix)
    try:
        assert pool.null_block.block_id == fix["null_block"]
        s0 = our_snap(pool)
        assert {k: s0[k] for k in POOL_FIELDS[1:]} == {k: fix["init"][k] for k in POOL_FIELDS[1:]}
        got = PT.drive_pool(pool, fix["scenario"], our_snap)
        compare_records(got, fix["records"], name)
        assert pool.kv_cache_manager.audit()[3] == 0          # bitmap sweep agrees with the page counters
    finally:
        pool.kv_cache_manager.close()


@pytest.mark.parametrize("seed,geom,blocks,cap", [(101, "llama", 512, 40), (102, "llama", 200, -1), (103, "mla64", 900, 120),
                                                  (104, "llama", 4096, 0), (105, "tp8", 3000, 300)])
def test_device_pool_matches_oracle_on_fresh_scenarios(seed, geom, blocks, cap):
    g = {"llama": dict(block_size=16, cell_size=2048, num_layers=32, num_kv_buffers=2),
         "mla64": dict(block_size=64, cell_size=1152, num_layers=61, num_kv_buffers=1),
         "tp8": dict(block_size=16, cell_size=256, num_layers=80, num_kv_buffers=2)}[geom]
    fix = {"config": dict(g, num_blocks=blocks, max_cached_blocks=cap),
           "scenario": PT.make_scenario(seed=seed, n_requests=80, concurrency=7, block_size=g["block_size"])}
    want = PT.drive_pool(oracle_pool_from(fix), fix["scenario"], lambda p: p.snap())
    pool = our_pool_from(fix)
    try:
        got = PT.drive_pool(pool, fix["scenario"], our_snap)
        compare_records(got, want, f"seed{seed}")
    finally:
        pool.kv_cache_manager.close()


def test_pool_api_edges():
    """The reference's own unit-test expectations (tests/test_prefix_cache.py): idempotent caching, touch
    reactivates, double key ignored, get_new_blocks beyond capacity raises ValueError, reset empties everything."""
    fix = {"config": dict(block_size=16, cell_size=1024, num_layers=1, num_kv_buffers=2, num_blocks=128, max_cached_blocks=-1)}
    pool = our_pool_from(fix)
    try:
        assert pool.get_num_free_blocks() == 127
        req = PT.Request(0, [bytes([i]) * 32 for i in range(4)])
        bl = pool.get_new_blocks(4)
        assert [b.ref_cnt for b in bl] == [1, 1, 1, 1]
        pool.cache_full_blocks(req, bl, 0, 3, 16, 0)
        pool.cache_full_blocks(req, bl, 0, 3, 16, 0)                   # idempotent
        assert len(pool._cached_blocks) == 3 and pool.keyed_ids() == [b.block_id for b in bl[:3]]
        assert pool.get_cached_block(req.block_hashes[0], [0])[0] is bl[0]
        assert pool.get_cached_block(req.block_hashes[0]) is bl[0]     # old single-group signature
        assert pool.get_cached_block(req.block_hashes[3], [0]) is None
        assert pool.get_cached_block(req.block_hashes[0], [0, 1]) is None   # all groups must hit
        pool.free_blocks(reversed(bl))
        assert pool.evictable_ids() == [b.block_id for b in reversed(bl[:3])]   # the unhashed tail was freed
        assert pool.get_num_free_blocks() == 127
        pool.touch([bl[1]])
        assert pool.evictable_ids() == [bl[2].block_id, bl[0].block_id] and bl[1].ref_cnt == 1
        pool.touch(([bl[1]], [bl[0]]))                                  # tuple-of-lists signature
        assert bl[1].ref_cnt == 2 and pool.evictable_ids() == [bl[2].block_id]
        with pytest.raises(ValueError):
            pool.get_new_blocks(10_000)
        assert pool.reset_prefix_cache() is True
        assert pool.evictable_ids() == [] and pool.keyed_ids() == [] and not pool._cached_blocks
        pool.free_blocks([bl[1], bl[1], bl[0], pool.null_block, None])  # keys are gone: freed at once; null/None skipped
        assert pool.ref_counts() == {} and pool.get_num_free_blocks() == 127
    finally:
        pool.kv_cache_manager.close()
