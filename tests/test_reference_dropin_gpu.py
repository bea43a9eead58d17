"""Drop-in check of the boundary: the UNMODIFIED reference `kvcached/kv_cache_manager.py` (Python bookkeeping
and all) runs on top of `kvcached_b200.vmm_ops` installed as `kvcached.vmm_ops` — i.e. our PageAllocator,
InternalPage, create_kv_tensors, map/unmap replace the reference's pybind extension with no source change —
and produces the reference's own results (oracle-checked), with real pages mapped by our VMM engine.

The reference package is the copy build_ref.py installed under oracle/_ref/pkg (never committed)."""
import importlib
import os
import sys
from unittest.mock import MagicMock

import pytest

import oracle
from kvcached_b200 import trace as T
from kvcached_b200 import vmm_ops as our_vmm_ops

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "oracle", "_ref", "pkg")
PAGE = 2 << 20


@pytest.fixture()
def reference_manager_module(monkeypatch):
    if not os.path.isfile(os.path.join(REF_PKG, "kvcached", "kv_cache_manager.py")):
        pytest.skip("oracle/_ref not built on this box")
    monkeypatch.setenv("KVCACHED_PAGE_PREALLOC_ENABLED", "false")
    monkeypatch.setenv("KVCACHED_CONTIGUOUS_LAYOUT", "true")
    monkeypatch.setenv("KVCACHED_IPC_NAME", f"kvb_dropin_{os.getpid()}")
    monkeypatch.syspath_prepend(REF_PKG)
    for k in [k for k in sys.modules if k == "kvcached" or k.startswith("kvcached.")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.setitem(sys.modules, "posix_ipc", MagicMock())
    import kvcached  # the reference package (pure Python part)
    monkeypatch.setitem(sys.modules, "kvcached.vmm_ops", our_vmm_ops)     # <- the drop-in
    monkeypatch.setattr(kvcached, "vmm_ops", our_vmm_ops, raising=False)
    mod = importlib.import_module("kvcached.kv_cache_manager")
    assert mod.PageAllocator is our_vmm_ops.PageAllocator and mod.InternalPage is our_vmm_ops.InternalPage
    yield mod
    for k in [k for k in sys.modules if k == "kvcached" or k.startswith("kvcached.")]:
        sys.modules.pop(k, None)


def test_unmodified_reference_manager_runs_on_our_extension(reference_manager_module):
    import torch
    torch.cuda.init()
    mod = reference_manager_module
    cfg = T.make_config("llama8b", 2048, num_layers=4)
    our_vmm_ops.init_kvcached("cuda:0", PAGE, True)
    try:
        per_layer = cfg["num_blocks"] * cfg["block_size"] * cfg["cell_size"] * 2
        (flat,) = our_vmm_ops.create_kv_tensors(per_layer, 2, "cuda:0", cfg["num_layers"])
        mgr = mod.KVCacheManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], cfg["num_layers"])
        assert mgr._post_init_done.wait(timeout=15.0)
        # the reference's own GPU test flow (tests/test_kvcache_manager.py:88-114, :164-194)
        a0 = mgr.available_size()
        ids = mgr.alloc(256)
        assert ids is not None and len(ids) == 256 and mgr.available_size() == a0 - 256
        assert our_vmm_ops.mapped_pages() == [0, 1, 2, 3]
        mgr.free(ids)
        assert mgr.available_size() == a0
        assert mgr.alloc(a0 + 1) is None
        mgr.trim()
        assert mgr.page_allocator.get_num_reserved_pages() == 0 and our_vmm_ops.mapped_pages() == []
        assert mgr.try_to_reserve(512) and len(mgr.reserved_blocks) == 512
        mgr.free_reserved()
        assert mgr.available_size() == a0
        mgr.trim()
        del mgr
        # fresh state for the seeded trace: the reference's Python bookkeeping over our page allocator == the oracle
        our_vmm_ops.shutdown_kvcached()
        our_vmm_ops.init_kvcached("cuda:0", PAGE, True)
        (flat,) = our_vmm_ops.create_kv_tensors(per_layer, 2, "cuda:0", cfg["num_layers"])
        mgr = mod.KVCacheManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], cfg["num_layers"])
        assert mgr._post_init_done.wait(timeout=15.0)
        # clear() included: the unmodified reference clear() calls page_allocator._stop_prealloc_thread(timeout=...),
        # which our PageAllocator binds (the reference's own extension does not, torch_bindings.cpp:212-213)
        trace = T.random_trace(cfg, 77, n_ops=220, extra=False, p_clear=0.03)
        assert sum(1 for o in trace["ops"] if o[0] == "clear") >= 3
        want = oracle.replay(trace)
        results = []
        for i, (op, w) in enumerate(zip(trace["ops"], want["records"])):
            r = oracle.apply_op(mgr, op, results)
            results.append(r)
            assert r == w["r"], f"op {i} {op[:2]}: {r} != {w['r']}"
            assert mgr.available_size() == w["avail"]
            assert list(mgr.avail_pages.keys()) == w["ap"]
            assert our_vmm_ops.mapped_pages() == w["mapped"]
        # data goes through the pages the reference manager asked us to map
        kv = flat.view(torch.int16).view(cfg["num_blocks"], cfg["num_layers"], 2, cfg["block_size"], -1)
        ids = mgr.alloc(130)
        kv[ids[0], :, :, 0, 0] = 11
        kv[ids[-1], :, :, 0, 0] = 22
        torch.cuda.synchronize()
        assert int(kv[ids[0], 1, 1, 0, 0]) == 11 and int(kv[ids[-1], 2, 0, 0, 0]) == 22
    finally:
        our_vmm_ops.shutdown_kvcached()


@pytest.mark.parametrize("engine", ["vllm", "sglang"])
def test_unmodified_reference_integration_interfaces_on_our_stack(reference_manager_module, engine, monkeypatch):
    """The reference's own integration entry points — init_kvcached / alloc_kv_cache / get_kv_cache_manager of
    kvcached/integration/{vllm,sglang}/interfaces.py, unmodified — on our extension AND our KVCacheManager (the
    one-line patch of INTEGRATION.md section 2): tokens written by our fused scatter kernel are read back through the
    per-layer views the reference builds (vllm/interfaces.py:291-298, sglang/interfaces.py:168-173), and the SGLang
    slot allocator expands block ids to exactly the slots those views address."""
    import numpy as np
    import torch
    import kvcached_b200
    from kvcached_b200 import kv_ops
    torch.cuda.init()
    itf = importlib.import_module(f"kvcached.integration.{engine}.interfaces")
    monkeypatch.setattr(itf, "KVCacheManager", kvcached_b200.KVCacheManager)       # <- section 2 of INTEGRATION.md
    L, bs, H, D = 32, 16, 8, 128
    cell = H * D * 2
    itf.init_kvcached(device="cuda:0")
    try:
        if engine == "vllm":
            views = itf.alloc_kv_cache((2, 4096, bs, H, D), bs, torch.float16, "cuda:0", L)
            assert len(views) == L and tuple(views[0].shape[:1]) == (2,) and tuple(views[0].shape[2:]) == (bs, H, D)
            mgr = itf.get_kv_cache_manager(4096, bs, cell, L)
            lay = kv_ops.layout_vllm_contiguous(views[0].data_ptr(), L, bs, cell, 2)
        else:
            k_views, v_views = itf.alloc_kv_cache((4096 * bs, H, D), torch.float16, "cuda:0", L, page_size=bs)
            assert len(k_views) == L == len(v_views) and tuple(k_views[0].shape[1:]) == (H, D)
            mgr = itf.get_kv_cache_manager(4096, bs, cell, L)                       # reserves the null block, like SGLang
            assert mgr.null_block == [0]
            lay = kv_ops.layout_sglang_contiguous(k_views[0].data_ptr(), L, cell, 2, block_size=bs)
        assert isinstance(mgr, kvcached_b200.KVCacheManager) and mgr._post_init_done.wait(timeout=30.0)
        ids = np.asarray(mgr.alloc(40), dtype=np.int64)
        T_tok = 300
        slots_np = np.random.default_rng(1).permutation((ids[:, None] * bs + np.arange(bs)).reshape(-1))[:T_tok]
        slots = torch.from_numpy(slots_np).cuda()
        k = torch.empty((L, T_tok, cell // 2), dtype=torch.int16, device="cuda").random_()
        v = torch.empty_like(k).random_()
        kv_ops.kv_scatter(lay, slots, k, v)
        torch.cuda.synchronize()
        blk, tok = slots // bs, slots % bs
        for layer in (0, 1, L // 2, L - 1):
            if engine == "vllm":
                got_k = views[layer][0][blk, tok]          # (T, H, D) through the reference's permuted view
                got_v = views[layer][1][blk, tok]
            else:
                got_k, got_v = k_views[layer][slots], v_views[layer][slots]
            assert torch.equal(got_k.contiguous().view(torch.int16).view(T_tok, -1), k[layer])
            assert torch.equal(got_v.contiguous().view(torch.int16).view(T_tok, -1), v[layer])
        if engine == "sglang":
            from kvcached_b200.sglang_alloc import ElasticPagedTokenToKVPoolAllocator

            class _Pool:                                    # what SGLang's memory pool exposes to its allocator
                kvcached_allocator = mgr
            al = ElasticPagedTokenToKVPoolAllocator(4096 * bs, bs, torch.float16, "cuda:0", _Pool())
            pre = torch.tensor([0, 0], dtype=torch.int64, device="cuda")
            seq = torch.tensor([40, 17], dtype=torch.int64, device="cuda")
            last = torch.tensor([-1, -1], dtype=torch.int64, device="cuda")
            out = al.alloc_extend(pre, pre.cpu(), seq, seq.cpu(), last, 57)
            assert out is not None and out.numel() == 57
            o = out.cpu().numpy()
            assert np.array_equal(o[:40] % bs, np.arange(40) % bs) and np.array_equal(o[40:] % bs, np.arange(17) % bs)
            k_views[3][out] = 7.0                           # writable through the reference's view = pages are mapped
            torch.cuda.synchronize()
            assert float(k_views[3][out[5]].flatten()[0]) == 7.0
            al.free(out)
        mgr.free(ids.tolist())
        mgr.close()
    finally:
        itf.shutdown_kvcached()


def test_unmodified_reference_sglang_mla_interface(reference_manager_module, monkeypatch):
    """BASELINE configs[3] (DeepSeek-V3 MLA latent KV, SGLang): the reference's `alloc_kv_cache(..., attention_type="MLA")`
    (sglang/interfaces.py:139-143: one buffer per layer, views of [tokens, L, 1, 576]) over our extension, tokens written by
    the fused kernel with one KV buffer per layer, slots from the device-side paged allocator."""
    import numpy as np
    import torch
    import kvcached_b200
    from kvcached_b200 import kv_ops
    from kvcached_b200.sglang_alloc import ElasticPagedTokenToKVPoolAllocator
    torch.cuda.init()
    itf = importlib.import_module("kvcached.integration.sglang.interfaces")
    monkeypatch.setattr(itf, "KVCacheManager", kvcached_b200.KVCacheManager)
    L, ps, dim = 61, 64, 576
    cell = dim * 2
    itf.init_kvcached(device="cuda:0")
    try:
        bufs = itf.alloc_kv_cache((1400 * ps, 1, dim), torch.bfloat16, "cuda:0", L, page_size=ps, attention_type="MLA")
        assert len(bufs) == L and tuple(bufs[0].shape[1:]) == (1, dim)
        mgr = itf.get_kv_cache_manager(1400, ps, cell, L, reserve_null_block=True, num_kv_buffers=1)
        assert mgr._post_init_done.wait(timeout=30.0) and mgr.null_block == [0]
        assert mgr.stats().blocks_per_page == (2 << 20) // (ps * cell) == 28        # blocks straddle page edges
        al = ElasticPagedTokenToKVPoolAllocator(1400 * ps, ps, torch.bfloat16, "cuda:0", mgr)
        pre = torch.zeros(3, dtype=torch.int64, device="cuda")
        seq = torch.tensor([300, 64, 1000], dtype=torch.int64, device="cuda")
        last = torch.full((3,), -1, dtype=torch.int64, device="cuda")
        slots = al.alloc_extend(pre, None, seq, None, last, 1364)
        assert slots is not None and slots.numel() == 1364 and len(set(slots.cpu().tolist())) == 1364
        lay = kv_ops.layout_sglang_contiguous(bufs[0].data_ptr(), L, cell, 1, block_size=ps)
        k = torch.empty((L, 1364, cell // 2), dtype=torch.int16, device="cuda").random_()
        kv_ops.kv_scatter(lay, slots, k)
        torch.cuda.synchronize()
        for layer in (0, 7, L - 1):
            got = bufs[layer][slots].contiguous().view(torch.int16).view(1364, -1)
            assert torch.equal(got, k[layer])
        back = torch.zeros_like(k)
        kv_ops.kv_gather(lay, slots, back)
        torch.cuda.synchronize()
        assert torch.equal(back, k)
        al.free(slots)
        assert mgr.available_size() > 0
        mgr.close()
    finally:
        itf.shutdown_kvcached()


@pytest.mark.parametrize("contiguous", [True, False])
def test_unmodified_reference_mamba_states_hybrid_model(reference_manager_module, monkeypatch, contiguous):
    """Hybrid models (attention + mamba layers): the reference's `alloc_mamba_states`
    (sglang/interfaces.py:178-401, unmodified) builds the conv / temporal state views of an elastic MambaPool over its
    OWN KV-cache group (one mamba slot = one kvcached block of the padded super-cell, one buffer per layer,
    `unified_pool` in the per-layer layout), next to the attention KV pool of group 0.  Both groups run on our
    extension and our KVCacheManager: slots allocated by the manager are backed in every mamba layer, states written
    through the reference's views read back, a freed-and-trimmed slot reads the zero page again, and the two groups'
    page sets stay independent (ElasticMambaPool, sglang/patches.py:812-1000, does exactly this with group ids >= 1000)."""
    import math
    from types import SimpleNamespace
    import torch
    import kvcached_b200
    torch.cuda.init()
    monkeypatch.setenv("KVCACHED_CONTIGUOUS_LAYOUT", "true" if contiguous else "false")
    for k in [k for k in sys.modules if k.startswith("kvcached.integration") or k == "kvcached.utils"]:
        monkeypatch.delitem(sys.modules, k)
    itf = importlib.import_module("kvcached.integration.sglang.interfaces")
    monkeypatch.setattr(itf, "KVCacheManager", kvcached_b200.KVCacheManager)
    itf.init_kvcached(device="cuda:0")
    try:
        assert our_vmm_ops_contiguous() == contiguous
        # ---- group 0: the attention layers' KV pool
        L_attn, bs, H, D = 4, 16, 8, 128
        k_views, v_views = itf.alloc_kv_cache((1024 * bs, H, D), torch.float16, "cuda:0", L_attn, page_size=bs)
        kv_mgr = itf.get_kv_cache_manager(1024, bs, H * D * 2, L_attn)
        assert kv_mgr._post_init_done.wait(timeout=30.0)
        # ---- group 1000: the mamba layers' state pool (Mamba2-like shapes)
        L_mamba, slots = 6, 96
        params = SimpleNamespace(shape=SimpleNamespace(conv=[(3, 1536)], temporal=(24, 64, 128)),
                                 dtype=SimpleNamespace(conv=torch.bfloat16, temporal=torch.float32))
        conv, temporal, info = itf.alloc_mamba_states(num_slots=slots, num_mamba_layers=L_mamba, cache_params=params,
                                                      device="cuda:0", group_id=1000)
        cell = info["cell_size"]
        raw = 3 * 1536 * 2 + 24 * 64 * 128 * 4
        assert cell >= raw and PAGE % cell == 0 and info["is_contiguous"] == contiguous
        m_mgr = kvcached_b200.KVCacheManager(slots, 1, cell, L_mamba, num_kv_buffers=1, group_id=1000,
                                             reserve_null_block=False)
        assert m_mgr._post_init_done.wait(timeout=30.0)
        spp = PAGE // cell                                   # slots per page
        got = m_mgr.alloc(spp + 2)                           # two pages of slots
        assert got == list(range(spp + 2))
        # (group 0 holds SGLang's null block: its page 0 is mapped and stays so)
        assert our_vmm_ops.mapped_pages(1000) == [0, 1] and our_vmm_ops.mapped_pages(0) == [0]

        def conv_at(layer, slot):
            return conv[0][layer, slot] if contiguous else conv[0][layer][slot]

        def temp_at(layer, slot):
            return temporal[layer, slot] if contiguous else temporal[layer][slot]

        for layer in (0, L_mamba - 1):
            for s in (0, spp - 1, spp + 1):
                conv_at(layer, s).fill_(float(layer + s))
                temp_at(layer, s).fill_(float(100 * layer + s))
        torch.cuda.synchronize()
        for layer in (0, L_mamba - 1):
            for s in (0, spp - 1, spp + 1):
                assert bool((conv_at(layer, s) == float(layer + s)).all())
                assert bool((temp_at(layer, s) == float(100 * layer + s)).all())
        assert tuple(conv_at(0, 0).shape) == (3, 1536) and tuple(temp_at(0, 0).shape) == (24, 64, 128)
        # the attention pool maps its own pages, independently
        ids = kv_mgr.alloc(70)
        assert len(our_vmm_ops.mapped_pages(0)) >= 2 and our_vmm_ops.mapped_pages(1000) == [0, 1]
        k_views[1][ids[0] * bs] = 3.0
        torch.cuda.synchronize()
        assert float(k_views[1][ids[0] * bs].flatten()[0]) == 3.0
        # a slot beyond what is mapped reads the zero page; freed + trimmed slots go back to it
        assert float(temp_at(0, 2 * spp + 3).flatten()[0]) == 0.0
        m_mgr.free(got)
        m_mgr.trim()
        torch.cuda.synchronize()
        assert our_vmm_ops.mapped_pages(1000) == []
        assert float(temp_at(0, 0).flatten()[0]) == 0.0 and float(conv_at(L_mamba - 1, spp + 1).flatten()[0]) == 0.0
        kv_mgr.free(ids)
        m_mgr.close()
        kv_mgr.close()
    finally:
        itf.shutdown_kvcached()


def our_vmm_ops_contiguous() -> bool:
    from kvcached_b200 import _lib
    return bool(_lib.load().kvb_contiguous_layout())
