"""Usage publishing and the kvctl limit on a live manager: the segment carries
used = in-use pages * L * P * nKV and prealloc = reserved pages * L * P * nKV (page_allocator.cpp:688-701) after
every step, and a new total_size is applied as KVCacheManager.resize before the next alloc
(kv_cache_manager.py:258-260) — state compared with the oracle driven by set_resize_target at the same points."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

from kvcached_b200 import kv_cache_manager as K  # noqa: E402

pytestmark = pytest.mark.gpu
P = 2 << 20


def test_segment_follows_the_manager_and_limit_resizes_it():
    L, nkv, bs, cell, blocks = 4, 2, 16, 2048, 64 * 40            # 64 blocks / page, 40 pages
    name = f"kvb200_gpu_shm_{os.getpid()}"
    m = K.KVCacheManager(blocks, bs, cell, L, num_kv_buffers=nkv, page_size=P, attach_vmm=False, publish_shm=True,
                         ipc_name=name)
    m.set_phys_free_pages(1 << 40)
    o = oracle.manager_from_config(dict(num_blocks=blocks, block_size=bs, cell_size=cell, num_layers=L, num_kv_buffers=nkv,
                                        page_size=P, reserve_null_block=False))
    o.set_phys_free_pages(1 << 40)
    per_page = L * P * nkv
    try:
        assert m.shm_name() == name
        total0 = blocks * bs * cell * L * nkv
        assert K.shm_read(name) == (total0, 0, 0)                # init_kv_cache_limit(total_memory_size_)

        def same():
            so, st = o.scalars(), m.stats()
            assert (st.num_free_pages, st.num_total_pages, st.num_reserved_pages, st.in_shrink) == \
                   (so["free_pages"], so["total_pages"], so["reserved_pages"], so["in_shrink"])
            assert m.available_size() == o.available_size()
            assert m.avail_pages() == o.avail_pages()
            t, used, pre = K.shm_read(name)
            assert used == st.num_inuse_pages * per_page and pre == st.num_reserved_pages * per_page

        a = m.alloc(64 * 10 + 5)
        assert a == o.alloc(64 * 10 + 5)
        same()
        assert K.shm_read(name)[1] == 11 * per_page
        m.free(a[:64 * 6]); o.free(a[:64 * 6])                    # six whole pages go to the reserved list
        same()
        assert K.shm_read(name)[2] == 6 * per_page
        assert m.page_allocator.get_resize_target() == -1
        # kvctl limit: 20 of 40 pages
        assert K.shm_set_limit(name, total0 // 2)
        assert m.page_allocator.get_resize_target() == blocks * bs * cell // 2
        o.set_resize_target(blocks * bs * cell // 2)
        b = m.alloc(3)
        assert b == o.alloc(3)
        assert m.stats().num_total_pages == 20
        same()
        # a limit below current usage: the shrink stays pending until enough is freed
        assert K.shm_set_limit(name, 3 * per_page)
        o.set_resize_target(3 * P)
        c = m.alloc(2)
        assert c == o.alloc(2) and m.in_shrink
        same()
        m.free(a[64 * 6:]); o.free(a[64 * 6:])
        same()
        m.free(b + c); o.free(b + c)
        same()
        assert not m.in_shrink and m.stats().num_total_pages == 3
        n_pub = m._lib.kvb_mgr_shm_publishes(m._h)
        assert 0 < n_pub <= 8                                     # at most one segment write per call
    finally:
        m.close()
    assert not os.path.exists("/dev/shm/" + name)                # unlinked with the manager (mem_info_tracker.hpp:228-232)
