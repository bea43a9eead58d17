#!/usr/bin/env python
"""Replay an alloc/free trace through the UNMODIFIED reference KVCacheManager.

TEST INFRASTRUCTURE ONLY (golden generation + the `--impl reference` bench arm).
Runs in its own process with oracle/_ref/pkg on PYTHONPATH (see build_ref.ref_env):

    python oracle/ref_replay.py trace.json out.json [--device cpu|cuda:0] [--time]

Trace format (shared with oracle/oracle.py and kvcached_b200/trace.py):
  {"config": {...}, "ops": [op, ...]}   op =
    ["alloc", n]            -> list[int] | None      kv_cache_manager.py:246
    ["free_ids", [ids]]     explicit ids              kv_cache_manager.py:307
    ["free_cat", [k, ...]]  concat of the results of earlier ops k (allocation order)
    ["free_slice", k, a, b] result_of(k)[a:b]
    ["resize", bytes]       -> bool                   kv_cache_manager.py:381
    ["trim"]                                          kv_cache_manager.py:404
    ["try_reserve", n]      -> bool                   kv_cache_manager.py:363
    ["free_reserved"]                                 kv_cache_manager.py:375
    ["avail"]               -> int                    kv_cache_manager.py:412
    ["clear"]                                         kv_cache_manager.py:444  (see the note in run())
    ["step"]                scheduler-step boundary (no-op for the reference)
    ["mark"]                records a timestamp (out["marks"]): section boundaries of bench.py's T-micro leg

After every op a snapshot is recorded: available_size, page counters, the
avail_pages / full_pages keys in dict order, reserved_blocks, in_shrink and the
sorted mapped-page set (observed by routing map/unmap through the reference's own
broadcast-callback hook, csrc/page_allocator.cpp:633-635, :664-668).
"""
from __future__ import annotations

import json
import os
import sys
import time
from unittest.mock import MagicMock


def _ids(r):
    return r if isinstance(r, list) else []


def run(trace: dict, device: str = "cpu", timing_only: bool = False, repeats: int = 1) -> dict:
    cfg = trace["config"]
    os.environ["KVCACHED_PAGE_PREALLOC_ENABLED"] = "true" if cfg.get("prealloc", False) else "false"
    os.environ["KVCACHED_CONTIGUOUS_LAYOUT"] = "true" if cfg.get("contiguous", True) else "false"
    if "max_reserved_pages" in cfg:
        os.environ["KVCACHED_MAX_RESERVED_PAGES"] = str(cfg["max_reserved_pages"])
    if "min_reserved_pages" in cfg:
        os.environ["KVCACHED_MIN_RESERVED_PAGES"] = str(cfg["min_reserved_pages"])
    os.environ["KVCACHED_PAGE_SIZE_MB"] = str(cfg.get("page_size", 2 << 20) >> 20)
    os.environ.setdefault("KVCACHED_IPC_NAME", f"kvb_refreplay_{os.getpid()}")
    sys.modules.setdefault("posix_ipc", MagicMock())

    import torch  # noqa: F401  (the reference requires torch before vmm_ops)
    from kvcached import vmm_ops
    from kvcached.kv_cache_manager import KVCacheManager

    P = cfg.get("page_size", 2 << 20)
    L = cfg["num_layers"]
    nkv = cfg.get("num_kv_buffers", 2)
    contiguous = cfg.get("contiguous", True)
    B = cfg["block_size"] * cfg["cell_size"]
    vmm_ops.init_kvcached(device, P, contiguous)
    t_create0 = time.perf_counter()
    if device == "cpu":
        # CPUPage::map only records the address (csrc/page.cpp:34-37): the virtual
        # tensor can stay tiny whatever the manager's size is.
        # One layer on the tensor side keeps the compound page (P*nkv) a divisor of both the
        # fixed base address and every offset pid*P*L*nkv the allocator emits, whatever L is
        # (csrc/ftensor.cpp:101,161 assert that alignment).  Bookkeeping still uses the real L.
        per_layer = 2 * P * nkv
        vmm_ops.create_kv_tensors(per_layer, 2, device, 1, num_kv_buffers=nkv)
    else:
        per_layer = ((cfg["num_blocks"] * B + P - 1) // P) * P * nkv
        vmm_ops.create_kv_tensors(per_layer, 2, device, L, num_kv_buffers=nkv)

    create_s = time.perf_counter() - t_create0      # create_kv_tensors: VA reservation + zero-page init (ftensor.cpp:160-176)
    mgr = KVCacheManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], L,
                         reserve_null_block=cfg.get("reserve_null_block", False),
                         num_kv_buffers=nkv)
    mgr._post_init_done.wait(timeout=15.0)
    # KVCacheManager.clear() calls page_allocator._stop_prealloc_thread(timeout=...) (kv_cache_manager.py:454), a name
    # csrc/torch_bindings.cpp:212-213 does not bind (only stop_prealloc_thread() is) — the reference's own clear()
    # raises AttributeError in this snapshot.  The ONE addition made here is that missing alias onto the reference's
    # own bound method; every line of clear() that follows runs unmodified, which is what pins clear() parity.
    pa_cls = type(mgr.page_allocator)
    if not hasattr(pa_cls, "_stop_prealloc_thread"):
        pa_cls._stop_prealloc_thread = lambda self, timeout=None: self.stop_prealloc_thread()

    mapped: set[int] = set()
    per_page = P * L * nkv if contiguous else P
    if not timing_only:
        pa = mgr.page_allocator

        def rec_map(world_size, offsets):
            for o in offsets:
                mapped.add(o // per_page)
            vmm_ops.map_to_kv_tensors(offsets)

        def rec_unmap(world_size, offsets):
            for o in offsets:
                mapped.discard(o // per_page)
            vmm_ops.unmap_from_kv_tensors(offsets)

        pa.set_should_use_worker_ipc_callback(lambda: True)
        pa.set_broadcast_map_callback(rec_map)
        pa.set_broadcast_unmap_callback(rec_unmap)
        if mgr.null_block is not None:      # page 0 was mapped before the hook existed
            mapped.add(0)

    def snap():
        pa = mgr.page_allocator
        return {
            "avail": mgr.available_size(),
            "pages": [pa.get_num_free_pages(), pa.get_num_inuse_pages(),
                      pa.get_num_reserved_pages(), pa.get_num_total_pages()],
            "nab": mgr.num_avail_blocks,
            "ap": list(mgr.avail_pages.keys()),
            "fp": sorted(mgr.full_pages.keys()),
            "rb": list(mgr.reserved_blocks),
            "shrink": bool(mgr.in_shrink),
            "mapped": sorted(mapped),
        }

    elapsed_all = []
    for _rep in range(max(1, repeats)):
        results, records, n_calls, n_blocks, elapsed = _one_pass(trace, mgr, snap, timing_only)
        elapsed_all.append(elapsed)
    out = {"elapsed_s": elapsed_all[-1], "elapsed_all_s": elapsed_all, "n_calls": n_calls, "n_blocks": n_blocks,
           "null_block": mgr.null_block, "nproc": os.cpu_count(), "device": device,
           "prealloc": bool(cfg.get("prealloc", False)), "marks": MARKS, "create_kv_tensors_s": create_s}
    if not timing_only:
        out["records"] = records
    else:
        out["final"] = {"avail": mgr.available_size()}
    vmm_ops.shutdown_kvcached()
    return out


MARKS: list = []


def _one_pass(trace, mgr, snap, timing_only):
    results: list = []
    records: list = []
    n_calls = 0
    n_blocks = 0
    t0 = time.perf_counter()
    for op in trace["ops"]:
        kind = op[0]
        r = None
        if kind == "alloc":
            try:
                r = mgr.alloc(op[1])
            except RuntimeError:          # "No free pages left", csrc/page_allocator.cpp:200-202
                r = "ERR:RuntimeError"
            n_calls += 1
            n_blocks += len(r) if isinstance(r, list) else 0
        elif kind in ("free_ids", "free_cat", "free_slice"):
            if kind == "free_ids":
                ids = list(op[1])
            elif kind == "free_cat":
                ids = [b for k in op[1] for b in _ids(results[k])]
            else:
                ids = list(_ids(results[op[1]])[op[2]:op[3]])
            mgr.free(ids)
            n_calls += 1
            n_blocks += len(ids)
        elif kind == "resize":
            try:
                r = bool(mgr.resize(op[1]))
            except AssertionError:        # kv_cache_manager.py:394 (reserved blocks present)
                r = "ERR:AssertionError"
        elif kind == "trim":
            mgr.trim()
        elif kind == "try_reserve":
            try:
                r = bool(mgr.try_to_reserve(op[1]))
            except RuntimeError:
                r = "ERR:RuntimeError"
        elif kind == "free_reserved":
            mgr.free_reserved()
        elif kind == "avail":
            r = mgr.available_size()
        elif kind == "clear":
            mgr.clear()
        elif kind == "step":
            pass
        elif kind == "mark":              # timing aid of bench.py's T-micro leg: a timestamp between trace sections
            MARKS.append(time.perf_counter())
        else:
            raise ValueError(f"unknown op {op}")
        results.append(r)
        if not timing_only:
            records.append({"r": r, **snap()})
    elapsed = time.perf_counter() - t0
    return results, records, n_calls, n_blocks, elapsed


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    device = "cpu"
    if "--device" in sys.argv:
        device = sys.argv[sys.argv.index("--device") + 1]
        args.remove(device)
    with open(args[0]) as f:
        trace = json.load(f)
    repeats = 1
    if "--repeat" in sys.argv:
        r = sys.argv[sys.argv.index("--repeat") + 1]
        args.remove(r)
        repeats = int(r)
    out = run(trace, device=device, timing_only="--time" in sys.argv, repeats=repeats)
    with open(args[1], "w") as f:
        json.dump(out, f)
