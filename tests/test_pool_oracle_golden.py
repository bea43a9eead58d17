"""The prefix-cache pinning restatement (oracle/pool_oracle.py over oracle/kv_oracle.cpp) against the fixtures the
UNMODIFIED reference ElasticBlockPool produced (tests/golden/make_golden_pool.py): after every event the returned
block ids, get_num_free_blocks, the LRU order, key counts, ref counts, available_size and avail-page order."""
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import pool_oracle  # noqa: E402
from kvcached_b200 import pool_trace as PT  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
POOL_FIELDS = ["r", "free", "evictable", "ncached", "nkeyed", "nref", "refsum", "avail", "pages", "ap"]


def pool_golden_names():
    return sorted(f[:-len(".json.gz")] for f in os.listdir(GOLDEN) if f.startswith("pool_") and f.endswith(".json.gz"))


def load_pool_golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rb") as f:
        return json.loads(f.read().decode())


def oracle_pool_from(fix):
    cfg = fix["config"]
    mcfg = {"num_blocks": cfg["num_blocks"], "block_size": cfg["block_size"], "cell_size": cfg["cell_size"],
            "num_layers": cfg["num_layers"], "num_kv_buffers": cfg["num_kv_buffers"], "page_size": 2 << 20,
            "reserve_null_block": False}
    mgr = oracle.manager_from_config(mcfg)
    return pool_oracle.OraclePool(mgr, cfg["num_blocks"], cfg.get("enable_caching", True), cfg["max_cached_blocks"])


def compare_records(got, want, name):
    assert len(got) == len(want)
    for i, (g, w) in enumerate(zip(got, want)):
        bad = [k for k in POOL_FIELDS if g.get(k) != w.get(k)]
        assert not bad, f"{name}: event {i} {w['ev']}: fields {bad} differ: got { {k: g.get(k) for k in bad} } want { {k: w.get(k) for k in bad} }"


@pytest.mark.parametrize("name", pool_golden_names())
def test_pool_oracle_matches_reference_fixture(name):
    fix = load_pool_golden(name)
    pool = oracle_pool_from(fix)
    assert pool.null_block.block_id == fix["null_block"]
    assert {k: pool.snap()[k] for k in POOL_FIELDS[1:]} == {k: fix["init"][k] for k in POOL_FIELDS[1:]}
    got = PT.drive_pool(pool, fix["scenario"], lambda p: p.snap())
    compare_records(got, fix["records"], name)


def test_fixtures_cover_the_rules():
    """hits, on-demand eviction, cap eviction, refusals, explicit evictions and resets all occur somewhere."""
    seen = {"hit": 0, "refused": 0, "evict": 0, "reset": 0, "cap": 0, "demand": 0}
    for name in pool_golden_names():
        fix = load_pool_golden(name)
        cap = fix["config"]["max_cached_blocks"]
        prev = len(fix["init"]["evictable"])
        for r in fix["records"]:
            if isinstance(r["r"], dict) and r["r"].get("hit"):
                seen["hit"] += 1
            if r["r"] == "ERR:ValueError":
                seen["refused"] += 1
            seen[r["ev"][0]] = seen.get(r["ev"][0], 0) + 1
            n = len(r["evictable"])
            if r["ev"][0] == "finish" and cap >= 0 and n == cap and prev + r["r"].get("freed", 0) > cap if isinstance(r["r"], dict) else False:
                seen["cap"] += 1
            if r["ev"][0] in ("arrive", "decode") and n < prev and isinstance(r["r"], dict) and not r["r"].get("hit"):
                seen["demand"] += 1
            prev = n
    assert all(seen[k] > 0 for k in ("hit", "refused", "evict", "reset", "demand")), seen
