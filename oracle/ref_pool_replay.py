#!/usr/bin/env python
"""Run a prefix-cache pinning scenario through the UNMODIFIED reference ElasticBlockPool
(kvcached/integration/vllm/patches.py:308-614) sitting on the UNMODIFIED reference KVCacheManager
("cpu" device: bookkeeping exact, map/unmap no-ops).

TEST INFRASTRUCTURE ONLY (golden generation; with "timing": true in the input, the reference side of bench.py's
T-mla-pin leg).  Own process, oracle/_ref/pkg on PYTHONPATH:

    python oracle/ref_pool_replay.py scenario.json out.json

The pool class is injected into a stand-in block_pool module exactly the way the reference's own
tests/test_prefix_cache.py:96-131 does it (vLLM's BlockPool base class and KVCacheBlock are the only things
the nested class takes from vLLM); the manager underneath is the real one, not the test's mock.
"""
from __future__ import annotations

import json
import os
import sys
import types
from unittest import mock


class _BlockPool:      # stand-in for vllm.v1.core.block_pool.BlockPool (the pool overrides every method it uses)
    pass


class _KVCacheBlock:   # stand-in for vllm.v1.core.kv_cache_utils.KVCacheBlock: block_id, ref_cnt, is_null
    def __init__(self, block_id: int, ref_cnt: int = 0):
        self.block_id = block_id
        self.ref_cnt = ref_cnt
        self.is_null = False


def run(fix: dict) -> dict:
    cfg = fix["config"]
    os.environ["KVCACHED_PAGE_PREALLOC_ENABLED"] = "false"
    os.environ["KVCACHED_CONTIGUOUS_LAYOUT"] = "true"
    os.environ["KVCACHED_PAGE_SIZE_MB"] = str(cfg.get("page_size", 2 << 20) >> 20)
    os.environ.setdefault("KVCACHED_IPC_NAME", f"kvb_refpool_{os.getpid()}")
    sys.modules.setdefault("posix_ipc", mock.MagicMock())

    import torch  # noqa: F401
    from kvcached import vmm_ops
    from kvcached.kv_cache_manager import KVCacheManager
    from kvcached.integration.vllm.patches import ElasticBlockPoolPatch

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from kvcached_b200.pool_trace import drive_pool   # pure Python, no extension needed

    P = cfg.get("page_size", 2 << 20)
    nkv = cfg.get("num_kv_buffers", 2)
    vmm_ops.init_kvcached("cpu", P, True)
    vmm_ops.create_kv_tensors(2 * P * nkv, 2, "cpu", 1, num_kv_buffers=nkv)
    mgr = KVCacheManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], cfg["num_layers"], num_kv_buffers=nkv)
    mgr._post_init_done.wait(timeout=15.0)

    mod = types.ModuleType("block_pool_standin")
    mod.BlockPool = _BlockPool
    mod.KVCacheBlock = _KVCacheBlock
    with mock.patch("kvcached.integration.vllm.interfaces.get_kv_cache_manager", return_value=mgr):
        ElasticBlockPoolPatch().inject_elastic_block_pool(mod)
        pool = mod.ElasticBlockPool(num_gpu_blocks=cfg["num_blocks"], block_size=cfg["block_size"],
                                    cell_size=cfg["cell_size"], num_layers=cfg["num_layers"],
                                    enable_caching=cfg.get("enable_caching", True), num_kv_buffers=nkv,
                                    max_cached_blocks=cfg.get("max_cached_blocks", 1000))

    def snap(p):
        pa = mgr.page_allocator
        refs = [(b.block_id, b.ref_cnt) for b in p.kv_block_pool if b.ref_cnt]
        return {
            "free": p.get_num_free_blocks(),
            "evictable": list(p._evictable_blocks.keys()),
            "ncached": len(p._cached_blocks),
            "nkeyed": len(p._block_id_to_key),
            "nref": len(refs),
            "refsum": sum((i + 1) * c for i, c in refs) % (1 << 61),
            "avail": mgr.available_size(),
            "pages": [pa.get_num_free_pages(), pa.get_num_inuse_pages(), pa.get_num_reserved_pages()],
            "ap": list(mgr.avail_pages.keys()),
        }

    if fix.get("timing"):      # bench.py's T-mla-pin leg: the scenario once, no snapshots, wall time
        import time
        t0 = time.perf_counter()
        recs = drive_pool(pool, fix["scenario"], lambda p: {})
        out = {"elapsed_s": time.perf_counter() - t0, "n_events": len(recs),
               "n_errors": sum(1 for r in recs if r["r"] == "ERR:ValueError"),
               "free_at_end": pool.get_num_free_blocks(), "evictable_at_end": len(pool._evictable_blocks)}
    else:
        out = {"null_block": pool.null_block.block_id, "init": snap(pool),
               "records": drive_pool(pool, fix["scenario"], snap)}
    vmm_ops.shutdown_kvcached()
    return out


if __name__ == "__main__":
    with open(sys.argv[1]) as f:
        fix = json.load(f)
    out = run(fix)
    with open(sys.argv[2], "w") as f:
        json.dump(out, f)
