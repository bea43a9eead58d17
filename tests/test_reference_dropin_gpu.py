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
        trace = T.random_trace(cfg, 77, n_ops=150, extra=False)
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
