#!/usr/bin/env python
"""Generate tests/golden/*.json.gz by replaying seeded traces through the UNMODIFIED
reference (oracle/_ref, built from /root/reference by oracle/build_ref.py).

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The GPU box never runs this; it only reads the committed fixtures.

Every fixture = {"trace": {...}, "records": [...], "null_block": ...}: for every op the
reference's return value plus a state snapshot (see oracle/ref_replay.py).  Prealloc is
off (the reference's prealloc thread is nondeterministic) and physical memory
is shimmed to 180 GiB free so available_size() is decided by virtual state alone.
"""
from __future__ import annotations

import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import build_ref  # noqa: E402
from kvcached_b200 import trace as T  # noqa: E402

MiB = 1 << 20


def cases() -> dict:
    c = {}
    small = T.make_config("llama8b", 2048, num_layers=4)
    c["appendix_b"] = {"config": small, "ops": [
        ["alloc", 200], ["free_slice", 0, 0, 64], ["free_slice", 0, 70, 100], ["alloc", 3], ["alloc", 100],
        ["free_ids", [130, 199, 65, 150]], ["alloc", 5], ["resize", 10 * 2 * MiB], ["resize", 2 * 2 * MiB],
        ["alloc", 1], ["free_cat", [4]], ["free_cat", [3]], ["free_cat", [6]], ["trim"]]}
    nb = T.make_config("llama8b", 2048, num_layers=4, reserve_null_block=True)
    c["null_block_basic"] = {"config": nb, "ops": [
        ["avail"], ["alloc", 5], ["alloc", 100], ["alloc", 3], ["free_ids", list(range(16, 86))], ["avail"],
        ["alloc", 4], ["free_ids", [1, 2, 3, 4, 5] + list(range(6, 16)) + list(range(86, 109)) + [16, 17, 18, 19]],
        ["avail"]]}
    c["reserved_cap_hash_order"] = {"config": nb, "ops": [
        ["alloc", 300], ["alloc", 896], ["free_cat", [1]], ["alloc", 192], ["avail"], ["trim"], ["alloc", 70],
        ["free_cat", [0]], ["free_cat", [3, 6]], ["avail"]]}
    for seed in range(4):
        c[f"llama8b_rand{seed}"] = T.random_trace(small, seed, n_ops=320)
        c[f"llama8b_null_rand{seed}"] = T.random_trace(nb, 100 + seed, n_ops=320)
    c["llama8b_rand_plain"] = T.random_trace(small, 7, n_ops=400, extra=False)
    c["llama8b_bigpool_rand"] = T.random_trace(T.make_config("llama8b", 65536), 11, n_ops=400, max_alloc=3000)
    for seed in range(2):
        c[f"qwen7b_rand{seed}"] = T.random_trace(T.make_config("qwen7b", 4096), 200 + seed, n_ops=300, max_alloc=400)
        c[f"mla_p1_rand{seed}"] = T.random_trace(T.make_config("mla_p1", 30000, reserve_null_block=True),
                                                 300 + seed, n_ops=300, max_alloc=4000)
        c[f"mla_p64_rand{seed}"] = T.random_trace(T.make_config("mla_p64", 700), 400 + seed, n_ops=300, max_alloc=60)
        c[f"tp8_rand{seed}"] = T.random_trace(T.make_config("llama70b_tp8", 8192), 500 + seed, n_ops=300,
                                              max_alloc=1500)
    c["fixture_rand"] = T.random_trace(T.make_config("ref_fixture", 65536), 600, n_ops=300, max_alloc=2000)
    # fewer pages than KVCACHED_MAX_RESERVED_PAGES (min(N, 10) rule, page_allocator.cpp:117-118)
    c["tiny_pool_rand"] = T.random_trace(T.make_config("llama8b", 192, num_layers=2), 700, n_ops=300, max_alloc=60)
    c["one_page_rand"] = T.random_trace(T.make_config("llama8b", 64, num_layers=2), 701, n_ops=200, max_alloc=30)
    # a block that fills more than half a page: pages hold 1 block, some hold none after alignment
    c["fat_block_rand"] = T.random_trace(
        dict(T.make_config("llama8b", 40), block_size=16, cell_size=(3 * MiB // 2) // 16), 702, n_ops=200,
        max_alloc=6)
    # decode traces (scheduler-step structure)
    c["decode_small"] = T.decode_trace(T.make_config("llama8b", 4096, num_layers=4), n_requests=96, seed=0)
    c["decode_null_tok16"] = T.decode_trace(T.make_config("llama8b", 4096, num_layers=4, reserve_null_block=True),
                                            n_requests=128, seed=1, tokens_per_step=16)
    c["decode_mla"] = T.decode_trace(T.make_config("mla_p64", 2000, reserve_null_block=True), n_requests=64, seed=2,
                                     tokens_per_step=8)
    # KVCacheManager.clear() (kv_cache_manager.py:444-489) between random traffic; oracle/ref_replay.py binds the one
    # missing private name the reference's clear() trips over, the method body runs unmodified
    c["clear_basic"] = {"config": nb, "ops": [
        ["alloc", 300], ["alloc", 70], ["free_slice", 0, 10, 200], ["try_reserve", 20], ["clear"], ["avail"],
        ["alloc", 130], ["resize", 6 * 2 * MiB], ["clear"], ["alloc", 5], ["free_cat", [9]], ["clear"], ["clear"],
        ["avail"]]}
    c["clear_rand0"] = T.random_trace(small, 800, n_ops=320, p_clear=0.03)
    c["clear_null_rand0"] = T.random_trace(nb, 801, n_ops=320, p_clear=0.03)
    c["clear_null_rand1"] = T.random_trace(nb, 802, n_ops=320, p_clear=0.05)
    c["clear_mla_p1"] = T.random_trace(T.make_config("mla_p1", 30000, reserve_null_block=True), 803, n_ops=250,
                                       max_alloc=4000, p_clear=0.03)
    c["clear_tp8"] = T.random_trace(T.make_config("llama70b_tp8", 8192), 804, n_ops=250, max_alloc=1500, p_clear=0.03)
    # micro pairs (bench_alloc pattern)
    c["micro_churn_k16"] = T.micro_trace(small, 16, 40, churn=True)
    c["micro_pinned_k4"] = T.micro_trace(small, 4, 40, churn=False)
    return c


def run_reference(trace: dict) -> dict:
    env = build_ref.ref_env(cpu_only=True)
    with tempfile.TemporaryDirectory() as d:
        tin, tout = os.path.join(d, "t.json"), os.path.join(d, "o.json")
        with open(tin, "w") as f:
            json.dump(trace, f)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_replay.py"), tin, tout],
                           env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference replay failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}")
        with open(tout) as f:
            return json.load(f)


def compact(records: list) -> list:
    """Store results in full, snapshots as compact tuples."""
    return [[r["r"], r["avail"], r["pages"], r["nab"], r["ap"], r["fp"], r["rb"], int(r["shrink"]), r["mapped"]]
            for r in records]


def main() -> None:
    if not build_ref.build():
        sys.exit("reference not available: cannot generate goldens")
    only = set(sys.argv[1:])
    for name, tr in cases().items():
        if only and name not in only:
            continue
        tr = json.loads(json.dumps(tr))        # normalise numpy ints
        tr.pop("meta", None)
        out = run_reference(tr)
        fix = {"trace": tr, "records": compact(out["records"]), "null_block": out["null_block"],
               "fields": ["r", "avail", "pages[free,inuse,reserved,total]", "num_avail_blocks", "avail_pages",
                          "full_pages(sorted)", "reserved_blocks", "in_shrink", "mapped(sorted)"],
               "generator": "tests/golden/make_golden.py via oracle/ref_replay.py (reference kvcached 0.1.5, cpu device)"}
        path = os.path.join(HERE, name + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(fix, separators=(",", ":")).encode())
        n_none = sum(1 for r in out["records"] if r["r"] is None and False)
        print(f"{name:28s} ops={len(tr['ops']):5d}  {os.path.getsize(path)/1024:7.1f} KiB")


if __name__ == "__main__":
    main()
