#!/usr/bin/env python
"""Generate tests/golden/pool_*.json.gz: prefix-cache pinning scenarios answered by the UNMODIFIED reference
(ElasticBlockPool over KVCacheManager, "cpu" device) through oracle/ref_pool_replay.py.  Needs /root/reference
(via oracle/_ref).  Run here, commit the fixtures; the GPU box only reads them.

    python tests/golden/make_golden_pool.py
"""
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
from kvcached_b200 import pool_trace as PT  # noqa: E402

LLAMA = {"block_size": 16, "cell_size": 2048, "num_layers": 32, "num_kv_buffers": 2}      # 64 blocks / page
MLA64 = {"block_size": 64, "cell_size": 1152, "num_layers": 61, "num_kv_buffers": 1}      # 28 blocks / page, straddling
MLA1 = {"block_size": 1, "cell_size": 1152, "num_layers": 61, "num_kv_buffers": 1}        # 1820 blocks / page

CASES = {
    # name: (config, scenario kwargs)
    "pool_llama_cap60": (dict(LLAMA, num_blocks=2048, max_cached_blocks=60), dict(seed=1, n_requests=48)),
    "pool_llama_unlimited": (dict(LLAMA, num_blocks=2048, max_cached_blocks=-1), dict(seed=2, n_requests=64, p_reset=0.0)),
    "pool_llama_cap0": (dict(LLAMA, num_blocks=1024, max_cached_blocks=0), dict(seed=3, n_requests=32)),
    "pool_llama_tight": (dict(LLAMA, num_blocks=320, max_cached_blocks=-1),            # evict-on-demand in get_new_blocks
                         dict(seed=4, n_requests=96, concurrency=6, p_reset=0.0, p_evict=0.01)),
    "pool_llama_tight2": (dict(LLAMA, num_blocks=256, max_cached_blocks=1000),
                          dict(seed=5, n_requests=96, concurrency=8, p_reset=0.005)),
    "pool_llama_refuse": (dict(LLAMA, num_blocks=100, max_cached_blocks=-1),              # get_new_blocks raises ValueError
                          dict(seed=9, n_requests=64, concurrency=10, p_reset=0.0, p_evict=0.02)),
    "pool_mla64_cap250": (dict(MLA64, num_blocks=1400, max_cached_blocks=250),            # C4: cap 16 000 tokens / 64
                          dict(seed=6, n_requests=64, block_size=64)),
    "pool_mla1_tokens": (dict(MLA1, num_blocks=20000, max_cached_blocks=1000),            # SGLang page_size 1 geometry
                         dict(seed=7, n_requests=40, prefix_blocks=(16, 200), suffix_blocks=(0, 64), decode_blocks=(0, 40),
                              block_size=1)),
    "pool_two_groups": (dict(LLAMA, num_blocks=1536, max_cached_blocks=120),            # hybrid attention: two KV-cache groups
                        dict(seed=10, n_requests=48, concurrency=6, groups=[0, 1])),
    "pool_nocache": (dict(LLAMA, num_blocks=1024, max_cached_blocks=1000, enable_caching=False), dict(seed=8, n_requests=32)),
}


def main():
    assert build_ref.build(), "oracle/_ref cannot be built (is /root/reference present?)"
    env = build_ref.ref_env(cpu_only=not os.path.exists("/dev/nvidiactl"))
    env["KVCACHED_LOG_LEVEL"] = "ERROR"
    for name, (cfg, kw) in CASES.items():
        kw = dict(kw)
        groups = kw.pop("groups", None)
        fix = {"name": name, "config": cfg, "scenario": PT.make_scenario(**kw)}
        if groups:
            fix["scenario"]["groups"] = groups
        with tempfile.TemporaryDirectory() as d:
            with open(os.path.join(d, "s.json"), "w") as f:
                json.dump(fix, f)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_pool_replay.py"), os.path.join(d, "s.json"),
                                os.path.join(d, "o.json")], env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-3000:]
            with open(os.path.join(d, "o.json")) as f:
                out = json.load(f)
        fix.update(out)
        raw = json.dumps(fix, separators=(",", ":")).encode()
        with gzip.GzipFile(os.path.join(HERE, name + ".json.gz"), "wb", mtime=0) as f:
            f.write(raw)
        recs = out["records"]
        hits = sum(1 for x in recs if isinstance(x["r"], dict) and x["r"].get("hit"))
        errs = sum(1 for x in recs if isinstance(x["r"], str) and x["r"].startswith("ERR"))
        print(f"{name}: {len(recs)} events, {hits} arrivals with hits, {errs} refused, max evictable "
              f"{max(len(x['evictable']) for x in recs)}, {len(raw) >> 10} KiB raw")


if __name__ == "__main__":
    main()
