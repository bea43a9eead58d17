"""The VMM engine and the KV scatter/gather kernels on a real device, through the C ABI.

Mirrors the reference's own GPU tests: tests/test_kvcache_manager.py (alloc/free conservation, trim,
reserve) and tests/test_paged_allocator_aliasing.py (unmapped pages alias the zero page; data written
through distinct physical pages reads back exactly).  KV contents: bit-exact against torch indexing
through the same virtual tensor (the reference publishes no scatter kernel — SURVEY §8(c))."""
import numpy as np
import pytest

from kvcached_b200 import _lib, kv_ops, vmm_ops
from kvcached_b200 import trace as T
from kvcached_b200.kv_cache_manager import KVCacheManager

pytestmark = pytest.mark.gpu

PAGE = 2 << 20


@pytest.fixture()
def kvcached_fixture():
    """tests/test_kvcache_manager.py:43-85 of the reference, scaled to 4 layers."""
    import torch
    torch.cuda.init()
    vmm_ops.init_kvcached("cuda:0", PAGE, True)
    L, blocks, bs, H, D = 4, 4096, 16, 8, 64
    cell = H * D * 2
    per_layer = blocks * bs * cell * 2           # K and V
    (flat,) = vmm_ops.create_kv_tensors(per_layer, 2, "cuda:0", L)
    vmm_ops.set_prefetch(0)                      # zero-page expectations below need pages to be mapped on demand only
    mgr = KVCacheManager(blocks, bs, cell, L)
    yield dict(mgr=mgr, flat=flat, L=L, blocks=blocks, bs=bs, cell=cell, H=H, D=D)
    mgr.close()
    vmm_ops.shutdown_kvcached()


def test_alloc_free_conservation_trim_reserve(kvcached_fixture):
    m = kvcached_fixture["mgr"]
    a0 = m.available_size()
    ids = m.alloc(256)
    assert ids is not None and len(ids) == 256 and m.available_size() == a0 - 256
    assert sorted(vmm_ops.mapped_pages()) == m.mapped_pages()
    m.free(ids)
    assert m.available_size() == a0
    assert m.alloc(a0 + 1) is None
    m.trim()
    assert m.page_allocator.get_num_reserved_pages() == 0 and vmm_ops.mapped_pages() == []
    assert m.try_to_reserve(512) and len(m.reserved_blocks) == 512
    m.free_reserved()
    assert m.reserved_blocks == [] and m.available_size() == a0


def test_unmapped_pages_alias_zero_page_and_mapped_pages_hold_data(kvcached_fixture):
    import torch
    f = kvcached_fixture
    m, flat, L, bs, cell = f["mgr"], f["flat"], f["L"], f["bs"], f["cell"]
    # layout: [num_blocks, L, 2, block, cell/2] int16 (vllm contiguous)
    kv = flat.view(torch.int16).view(f["blocks"], L, 2, bs, cell // 2)
    # nothing mapped: writes land in the shared zero page and alias across pages
    kv[0, 0, 0, 0, 0] = 7
    torch.cuda.synchronize()
    bpp = PAGE // (bs * cell)
    assert int(kv[bpp, 0, 0, 0, 0]) == 7      # same offset inside another (unmapped) page aliases
    kv[0, 0, 0, 0, 0] = 0
    ids = m.alloc(3 * bpp)                     # three physical pages
    assert ids == list(range(3 * bpp))
    for p in range(3):
        kv[p * bpp, :, :, 0, 0] = 100 + p
    torch.cuda.synchronize()
    for p in range(3):
        assert torch.all(kv[p * bpp, :, :, 0, 0] == 100 + p)
    assert int(kv[3 * bpp, 0, 0, 0, 0]) == 0   # still the zero page
    m.free(ids)
    m.trim()
    torch.cuda.synchronize()
    assert int(kv[0, 0, 0, 0, 0]) == 0         # unmapped again: reads the zero page


def test_map_unmap_by_offsets_is_idempotent(kvcached_fixture):
    f = kvcached_fixture
    per = PAGE * f["L"] * 2
    assert vmm_ops.map_to_kv_tensors([0, 3 * per])
    assert vmm_ops.map_to_kv_tensors([3 * per])          # double map tolerated
    assert vmm_ops.mapped_pages() == [0, 3]
    assert vmm_ops.unmap_from_kv_tensors([0, 3 * per, 5 * per])   # unmapping an unmapped page tolerated
    assert vmm_ops.mapped_pages() == []
    st = vmm_ops.vmm_stats()
    assert st["pooled_handles"] == 2 and st["handles_created"] >= 3   # zero page + 2, parked in the pool


def test_map_ahead_follows_the_free_list(kvcached_fixture):
    """The VMM thread maps the next pages of the bookkeeping's free list in its idle time: the map event that
    follows is a no-op on the scheduler's path, the bookkeeping's mapped set never shows the extra pages, and
    trim() hands everything back."""
    import time
    import torch
    f = kvcached_fixture
    m, L, bs, cell = f["mgr"], f["L"], f["bs"], f["cell"]
    kv = f["flat"].view(torch.int16).view(f["blocks"], L, 2, bs, cell // 2)
    bpp = PAGE // (bs * cell)
    vmm_ops.set_prefetch(4)

    def settle():
        vmm_ops.fence()
        t0 = time.time()
        while vmm_ops.engine_stats()["need_hinted"] and time.time() - t0 < 60:
            time.sleep(0.001)

    try:
        a = m.alloc(2 * bpp)                              # pages 0, 1
        settle()
        assert m.free_page_list()[:4] == [2, 3, 4, 5] == vmm_ops.prefetched_pages()
        assert vmm_ops.mapped_pages() == [0, 1] == m.mapped_pages()
        assert vmm_ops.vmm_stats()["mapped_pages"] == 6   # physically: 2 wanted + 4 ahead
        e0, t0 = vmm_ops.engine_stats(), vmm_ops.vmm_timing()
        b = m.alloc(2 * bpp)                              # pages 2, 3: already in place
        e1, t1 = vmm_ops.engine_stats(), vmm_ops.vmm_timing()
        assert e1["prefetch_hits"] == e0["prefetch_hits"] + 2
        assert b == list(range(2 * bpp, 4 * bpp))
        kv[2 * bpp, :, :, 0, 0] = 321
        kv[3 * bpp, :, :, 0, 0] = 654
        torch.cuda.synchronize()
        settle()
        assert vmm_ops.prefetched_pages() == [4, 5, 6, 7]  # the window moved with the free list
        assert int(kv[2 * bpp, 0, 0, 0, 0]) == 321 and int(kv[3 * bpp, L - 1, 1, 0, 0]) == 654
        assert vmm_ops.mapped_pages() == [0, 1, 2, 3]
        m.free(a + b)
        m.trim()                                          # settle point: hints dropped, everything unmapped
        assert vmm_ops.prefetched_pages() == [] and vmm_ops.mapped_pages() == []
        assert vmm_ops.vmm_stats()["mapped_pages"] == 0
        torch.cuda.synchronize()
        assert int(kv[2 * bpp, 0, 0, 0, 0]) == 0          # back on the zero page
    finally:
        vmm_ops.set_prefetch(0)


def test_retention_keeps_the_handle_until_it_is_needed_or_times_out(kvcached_fixture):
    """A page the bookkeeping unmapped stays on its own physical handle: mapped again soon it costs no driver call;
    after the retention time, or at a settle point, it is back on the zero page and the memory back on the device."""
    import time
    import torch
    f = kvcached_fixture
    m, L, bs, cell = f["mgr"], f["L"], f["bs"], f["cell"]
    kv = f["flat"].view(torch.int16).view(f["blocks"], L, 2, bs, cell // 2)
    bpp = PAGE // (bs * cell)
    lib = _lib.load()
    try:
        vmm_ops.set_retention(60_000)
        ids = np.arange(12, 16, dtype=np.int64)               # pages 12..15 through the event interface
        ev_map, ev_unmap = ids * 2, ids * 2 + 1
        assert lib.kvb_vmm_enqueue_page_events(0, ev_map.ctypes.data_as(_lib.p_i64), 4) == 1
        vmm_ops.wait_maps()
        kv[12 * bpp, 0, 0, 0, 0] = 77
        torch.cuda.synchronize()
        t0 = vmm_ops.vmm_timing()
        assert lib.kvb_vmm_enqueue_page_events(0, ev_unmap.ctypes.data_as(_lib.p_i64), 4) == 1
        vmm_ops.fence()
        e = vmm_ops.engine_stats()
        assert e["retained"] == 4 and vmm_ops.mapped_pages() == [] and vmm_ops.vmm_stats()["mapped_pages"] == 4
        assert lib.kvb_vmm_enqueue_page_events(0, ev_map.ctypes.data_as(_lib.p_i64), 4) == 1
        vmm_ops.wait_maps()
        t1 = vmm_ops.vmm_timing()
        assert t1["unmap_calls"] == t0["unmap_calls"] and t1["setaccess_calls"] == t0["setaccess_calls"]   # no driver call at all
        assert vmm_ops.engine_stats()["retention_hits"] == e["retention_hits"] + 4
        assert int(kv[12 * bpp, 0, 0, 0, 0]) == 77           # the same physical page (a fresh one would be equally valid)
        # time-out: the VMM thread hands the memory back on its own
        vmm_ops.set_retention(150)
        assert lib.kvb_vmm_enqueue_page_events(0, ev_unmap.ctypes.data_as(_lib.p_i64), 4) == 1
        released0 = vmm_ops.vmm_stats()["handles_released"]
        t_end = time.time() + 60
        while time.time() < t_end and (vmm_ops.engine_stats()["retained"] or vmm_ops.pending_unmaps()):
            time.sleep(0.02)
        assert vmm_ops.engine_stats()["retained"] == 0 and vmm_ops.vmm_stats()["mapped_pages"] == 0
        assert vmm_ops.vmm_stats()["handles_released"] >= released0 + 4
        torch.cuda.synchronize()
        assert int(kv[12 * bpp, 0, 0, 0, 0]) == 0            # zero page again
        # the reserve: retained pages are handed back while less than the given fraction of the device is free
        vmm_ops.set_retention(60_000)
        vmm_ops.set_retain_free_fraction(0.9999, 0.9999)
        assert lib.kvb_vmm_enqueue_page_events(0, ev_map.ctypes.data_as(_lib.p_i64), 4) == 1
        assert lib.kvb_vmm_enqueue_page_events(0, ev_unmap.ctypes.data_as(_lib.p_i64), 4) == 1
        t_end = time.time() + 60
        while time.time() < t_end and (vmm_ops.engine_stats()["retained"] or vmm_ops.pending_unmaps()):
            time.sleep(0.02)
        assert vmm_ops.engine_stats()["retained"] == 0 and vmm_ops.vmm_stats()["mapped_pages"] == 0
        vmm_ops.set_retain_free_fraction(0.20, 0.02)
        # retention off: every unmap is applied at once
        vmm_ops.set_retention(0)
        assert lib.kvb_vmm_enqueue_page_events(0, ev_map.ctypes.data_as(_lib.p_i64), 4) == 1
        assert lib.kvb_vmm_enqueue_page_events(0, ev_unmap.ctypes.data_as(_lib.p_i64), 4) == 1
        vmm_ops.fence()
        assert vmm_ops.vmm_stats()["mapped_pages"] == 0 and vmm_ops.engine_stats()["retained"] == 0
    finally:
        vmm_ops.set_retention(2000)
        vmm_ops.set_retain_free_fraction(0.20, 0.02)


@pytest.mark.parametrize("use_thread", [True, False])
def test_async_engine_settles_to_the_wanted_state(kvcached_fixture, use_thread):
    """Events only move the wanted state; the VMM thread (or the caller, KVB200_VMM_ASYNC=0) reconciles.
    Opposite events on one page cancel; wait_maps() guarantees every wanted page is backed; fence() settles."""
    import torch
    f = kvcached_fixture
    lib = _lib.load()
    vmm_ops.set_async(use_thread)
    L, bs, cell = f["L"], f["bs"], f["cell"]
    kv = f["flat"].view(torch.int16).view(f["blocks"], L, 2, bs, cell // 2)
    bpp = PAGE // (bs * cell)
    rng = np.random.default_rng(5)
    n_pages = 32
    want = np.zeros(n_pages, dtype=bool)
    created0 = vmm_ops.vmm_stats()["handles_created"]
    for rnd in range(40):
        k = int(rng.integers(1, 12))
        pages = rng.integers(0, n_pages, size=k)
        ev = []
        for p in pages:                       # alternate per page, like the bookkeeping does
            ev.append(int(p) * 2 + (1 if want[p] else 0))
            want[p] = ~want[p]
        a = np.asarray(ev, dtype=np.int64)
        assert lib.kvb_vmm_enqueue_page_events(0, a.ctypes.data_as(_lib.p_i64), a.size) == 1, _lib.last_error()
        vmm_ops.wait_maps()
        for p in np.nonzero(want)[0][:4]:     # wanted pages are usable right after wait_maps()
            kv[int(p) * bpp, 0, 0, 0, 0] = 1000 + int(p)
        torch.cuda.synchronize()
        for p in np.nonzero(want)[0][:4]:
            assert int(kv[int(p) * bpp, 0, 0, 0, 0]) == 1000 + int(p)
    vmm_ops.fence()
    assert vmm_ops.mapped_pages() == np.nonzero(want)[0].tolist()
    vmm_ops.settle()                          # retained pages give their handles back
    assert vmm_ops.pending_unmaps() == 0 and vmm_ops.engine_stats()["retained"] == 0
    assert vmm_ops.vmm_stats()["mapped_pages"] == int(want.sum())
    for p in np.nonzero(~want)[0][:4]:        # settled unmapped pages read the zero page again
        assert int(kv[int(p) * bpp, 0, 0, 0, 1]) == 0
    # a page unmapped and mapped again before the thread got to it costs nothing
    st = vmm_ops.vmm_stats()
    assert st["handles_created"] - created0 <= n_pages
    ids = np.arange(n_pages, dtype=np.int64)
    assert lib.kvb_unmap_pages(0, ids.ctypes.data_as(_lib.p_i64), ids.size) == 1
    assert vmm_ops.mapped_pages() == []
    vmm_ops.set_async(True)


@pytest.mark.parametrize("layout", ["vllm_contig", "sglang_contig", "per_layer", "mla"])
@pytest.mark.parametrize("T_tokens", [1, 37, 4096])
def test_kv_scatter_gather_bit_exact(layout, T_tokens):
    import torch
    torch.cuda.init()
    contiguous = layout != "per_layer"
    vmm_ops.init_kvcached("cuda:0", PAGE, contiguous)
    try:
        L, bs, blocks = 4, 16, 2048
        nkv = 1 if layout == "mla" else 2
        cell = 576 * 2 if layout == "mla" else 8 * 128 * 2
        if layout == "sglang_contig":
            bs = 1
            blocks = 2048 * 16
        blk_bytes = bs * cell
        per_layer = -(-blocks * blk_bytes * nkv // (2 * PAGE)) * 2 * PAGE
        ranges = vmm_ops.create_kv_tensors_raw(per_layer, L, nkv)
        tens = vmm_ops.create_kv_tensors(per_layer, 2, "cuda:0", L, nkv)
        m = KVCacheManager(blocks, bs, cell, L, num_kv_buffers=nkv)
        need_blocks = -(-T_tokens // bs) + 3
        ids = np.asarray(m.alloc(need_blocks), dtype=np.int64)
        rng = np.random.default_rng(T_tokens)
        all_slots = (ids[:, None] * bs + np.arange(bs)[None, :]).reshape(-1)
        slots_np = rng.choice(all_slots, size=T_tokens, replace=False)
        slots = torch.from_numpy(slots_np).cuda()
        g = torch.Generator(device="cuda").manual_seed(1)
        k_src = torch.randint(-32768, 32767, (L, T_tokens, cell // 2), dtype=torch.int16, device="cuda", generator=g)
        v_src = torch.randint(-32768, 32767, (L, T_tokens, cell // 2), dtype=torch.int16, device="cuda", generator=g) \
            if nkv == 2 else None
        base = ranges[0][0]
        if layout in ("vllm_contig", "mla"):
            lay = kv_ops.layout_vllm_contiguous(base, L, bs, cell, nkv)
            view = tens[0].view(torch.int16).view(-1, L, nkv, bs, cell // 2)

            def ref_write(kv_i, l, src):
                view[slots // bs, l, kv_i, slots % bs] = src
            def ref_read(kv_i, l):
                return view[slots // bs, l, kv_i, slots % bs]
        elif layout == "sglang_contig":
            lay = kv_ops.layout_sglang_contiguous(base, L, cell, nkv)
            view = tens[0].view(torch.int16).view(-1, L, nkv, cell // 2)

            def ref_write(kv_i, l, src):
                view[slots, l, kv_i] = src
            def ref_read(kv_i, l):
                return view[slots, l, kv_i]
        else:
            assert all(ranges[i + 1][0] - ranges[i][0] == ranges[0][1] for i in range(L - 1))
            lay = kv_ops.layout_per_layer(base, ranges[0][1], per_layer, L, bs, cell, nkv)
            views = [t.view(torch.int16).view(2, -1, cell // 2) for t in tens]

            def ref_write(kv_i, l, src):
                views[l][kv_i, slots] = src
            def ref_read(kv_i, l):
                return views[l][kv_i, slots]
        # 1. scatter with the kernel, read back with torch indexing
        kv_ops.kv_scatter(lay, slots, k_src, v_src)
        torch.cuda.synchronize()
        for l in range(L):
            assert torch.equal(ref_read(0, l), k_src[l])
            if nkv == 2:
                assert torch.equal(ref_read(1, l), v_src[l])
        # 2. write with torch indexing, gather with the kernel
        k2 = k_src.flip(1).contiguous()
        v2 = v_src.flip(1).contiguous() if nkv == 2 else None
        for l in range(L):
            ref_write(0, l, k2[l])
            if nkv == 2:
                ref_write(1, l, v2[l])
        k_out = torch.empty_like(k_src)
        v_out = torch.empty_like(k_src) if nkv == 2 else None
        kv_ops.kv_gather(lay, slots, k_out, v_out)
        torch.cuda.synchronize()
        assert torch.equal(k_out, k2)
        if nkv == 2:
            assert torch.equal(v_out, v2)
        # 3. per-layer launch range == fused launch
        k_out.zero_()
        for l in range(L):
            kv_ops.kv_gather(lay, slots, k_out[l:l + 1], None if v_out is None else v_out[l:l + 1], l, l + 1)
        torch.cuda.synchronize()
        assert torch.equal(k_out, k2)
        m.close()
    finally:
        vmm_ops.shutdown_kvcached()


def test_raw_control_and_pool(kvcached_fixture):
    lib = _lib.load()
    n, iters = 4, 5
    a = np.zeros(iters)
    b = np.zeros(iters)
    rc = lib.kvb_vmm_raw_control(0, n, iters, 1, a.ctypes.data_as(_lib.p_f64), b.ctypes.data_as(_lib.p_f64))
    assert rc == 0, _lib.last_error()
    assert (a > 0).all() and (b > 0).all()


@pytest.mark.parametrize("unified", [False, True])
def test_per_layer_and_unified_pool_mapping(unified):
    """Non-contiguous layouts (allocator.cpp:178-207): one flat tensor per layer; an offset maps, in every layer, the K
    page at `off` and the V page at `off + size / 2` (per-layer), or the single page at `off` (unified pool, K and V
    interleaved inside it).  Everything else keeps reading the zero page."""
    import torch
    torch.cuda.init()
    L, pages = 3, 8
    vmm_ops.init_kvcached("cuda:0", PAGE, False)
    try:
        vmm_ops.set_prefetch(0)
        per_layer = pages * PAGE * 2
        tens = vmm_ops.create_kv_tensors(per_layer, 2, "cuda:0", L, 2, 0, unified)
        assert len(tens) == L and all(t.numel() * 2 == per_layer for t in tens)
        half = per_layer // 2 // 2                       # int16 elements to the V half
        epp = PAGE // 2                                  # int16 elements per page
        assert vmm_ops.map_to_kv_tensors([2 * PAGE, 5 * PAGE])
        assert vmm_ops.mapped_pages() == [2, 5]
        for l, t in enumerate(tens):
            t[2 * epp + 7] = 100 + l                     # K page 2 (per-layer) / page 2 (unified)
            t[5 * epp] = 200 + l
            if not unified:
                t[half + 2 * epp + 9] = 300 + l          # V page 2
        torch.cuda.synchronize()
        for l, t in enumerate(tens):
            assert int(t[2 * epp + 7]) == 100 + l and int(t[5 * epp]) == 200 + l
            assert int(t[3 * epp + 7]) == 0 and int(t[0]) == 0          # unmapped pages: zero page
            if not unified:
                assert int(t[half + 2 * epp + 9]) == 300 + l
        st = vmm_ops.vmm_stats()
        assert st["bytes_mapped"] == 2 * PAGE * (L if unified else 2 * L)
        assert vmm_ops.unmap_from_kv_tensors([2 * PAGE, 5 * PAGE])
        torch.cuda.synchronize()
        assert vmm_ops.mapped_pages() == [] and int(tens[0][2 * epp + 7]) == 0
    finally:
        vmm_ops.shutdown_kvcached()


@pytest.mark.parametrize("cell", [256, 512, 1024])
@pytest.mark.parametrize("T_tokens", [5, 4096])
def test_kv_scatter_gather_tp_shard_cells_80_layers(cell, T_tokens):
    """The per-rank shapes of Llama-3.3-70B under TP 8 / 4 / 2 (BASELINE configs[4]: 80 layers, 1 / 2 / 4 KV heads per
    rank -> cell 256 / 512 / 1024 B): several cells share a warp in the kernel, so these sizes take other code paths
    than the 2 KiB cell.  Also: padding slots (-1) move nothing, and per-layer scatter launches == the fused launch."""
    import torch
    torch.cuda.init()
    vmm_ops.init_kvcached("cuda:0", PAGE, True)
    try:
        L, bs, blocks = 80, 16, 1024
        per_layer = -(-blocks * bs * cell * 2 // (2 * PAGE)) * 2 * PAGE
        ranges = vmm_ops.create_kv_tensors_raw(per_layer, L, 2)
        tens = vmm_ops.create_kv_tensors(per_layer, 2, "cuda:0", L, 2)
        m = KVCacheManager(blocks, bs, cell, L)
        ids = np.asarray(m.alloc(-(-T_tokens // bs) + 2), dtype=np.int64)
        rng = np.random.default_rng(cell + T_tokens)
        slots_np = rng.choice((ids[:, None] * bs + np.arange(bs)[None, :]).reshape(-1), size=T_tokens, replace=False)
        pad = rng.random(T_tokens) < 0.2                      # every fifth token is padding
        slots_np = np.where(pad, -1, slots_np)
        slots = torch.from_numpy(slots_np).cuda()
        live = torch.from_numpy(~pad).cuda()
        lslots = slots[live]
        g = torch.Generator(device="cuda").manual_seed(3)
        k_src = torch.randint(-32768, 32767, (L, T_tokens, cell // 2), dtype=torch.int16, device="cuda", generator=g)
        v_src = torch.randint(-32768, 32767, (L, T_tokens, cell // 2), dtype=torch.int16, device="cuda", generator=g)
        lay = kv_ops.layout_vllm_contiguous(ranges[0][0], L, bs, cell, 2)
        view = tens[0].view(torch.int16).view(-1, L, 2, bs, cell // 2)
        before = view[ids[0]].clone()
        kv_ops.kv_scatter(lay, slots, k_src, v_src)
        torch.cuda.synchronize()
        for l in (0, 1, 39, 79):
            assert torch.equal(view[lslots // bs, l, 0, lslots % bs], k_src[l][live])
            assert torch.equal(view[lslots // bs, l, 1, lslots % bs], v_src[l][live])
        # every layer at once, bytewise
        got_k = view[lslots // bs, :, 0, lslots % bs].transpose(0, 1)
        assert torch.equal(got_k, k_src[:, live])
        # gather: padding rows stay untouched
        k_out = torch.full_like(k_src, 7)
        v_out = torch.full_like(v_src, 7)
        kv_ops.kv_gather(lay, slots, k_out, v_out)
        torch.cuda.synchronize()
        assert torch.equal(k_out[:, live], k_src[:, live]) and torch.equal(v_out[:, live], v_src[:, live])
        assert bool((k_out[:, ~live] == 7).all()) and bool((v_out[:, ~live] == 7).all())
        # per-layer scatter launches (the engine's own call shape [T, H, D] per layer) == the fused launch
        k2, v2 = k_src.flip(1).contiguous(), v_src.flip(1).contiguous()
        for l in range(L):
            kv_ops.kv_scatter(lay, slots, k2[l], v2[l], l, l + 1)
        k_out.zero_(); v_out.zero_()
        kv_ops.kv_gather(lay, slots, k_out, v_out)
        torch.cuda.synchronize()
        assert torch.equal(k_out[:, live], k2[:, live]) and torch.equal(v_out[:, live], v2[:, live])
        m.close()
    finally:
        vmm_ops.shutdown_kvcached()


def test_background_zero_page_init_on_the_device():
    """kvb_vmm_zero_init_mode(1): create returns before the range is backed; pages mapped meanwhile hold data; after
    wait_init() every unmapped unit reads the zero page (ftensor.cpp:160-176's guarantee, reached later)."""
    import time
    import torch
    torch.cuda.init()
    vmm_ops.zero_init_mode(1)
    try:
        vmm_ops.init_kvcached("cuda:0", PAGE, False)           # per-layer: 8 layers x K,V = 16 units per page
        L, pages = 8, 256
        t0 = time.perf_counter()
        tens = vmm_ops.create_kv_tensors(pages * PAGE * 2, 2, "cuda:0", L)
        t_create = time.perf_counter() - t0
        pr = vmm_ops.init_progress()
        assert pr["units"] == pages * 2 * L and pr["units_done"] < pr["units"], (pr, t_create)
        lib = _lib.load()
        ids = np.array([200, 201], dtype=np.int64)             # far ahead of the background init
        assert lib.kvb_map_pages(0, ids.ctypes.data_as(_lib.p_i64), ids.size) == 1, _lib.last_error()
        per_page = PAGE // 2                                   # int16 elements per page
        for l in (0, L - 1):
            k_half = tens[l].view(torch.int16)[: pages * per_page]
            k_half[200 * per_page: 202 * per_page] = 1234 + l
        torch.cuda.synchronize()
        vmm_ops.wait_init()
        pr = vmm_ops.init_progress()
        assert pr["units_done"] == pr["units"]
        for l in (0, L - 1):
            k_half = tens[l].view(torch.int16)[: pages * per_page]
            assert bool((k_half[200 * per_page: 202 * per_page] == 1234 + l).all())
            assert int(k_half[5 * per_page]) == 0 and int(k_half[255 * per_page + 7]) == 0     # zero page
        assert lib.kvb_unmap_pages(0, ids.ctypes.data_as(_lib.p_i64), ids.size) == 1
        torch.cuda.synchronize()
        assert int(tens[0].view(torch.int16)[200 * per_page]) == 0
        assert vmm_ops.mapped_pages() == []
    finally:
        vmm_ops.shutdown_kvcached()
        vmm_ops.zero_init_mode(-1)
