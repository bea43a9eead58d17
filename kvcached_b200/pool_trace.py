"""Prefix-cache pinning scenarios (T-mla-pin) and the driver that runs one against ANY
ElasticBlockPool implementation through the pool's public method set only — the reference's
(kvcached/integration/vllm/patches.py:308-614), the oracle restatement, or kvcached_b200.block_pool.

The driver calls the pool the way vLLM's KV-cache manager does for one KV-cache group:
  arrival   longest cached prefix via get_cached_block(hash, [gid]) per block, touch(hits),
            get_new_blocks(missing), cache_full_blocks(request, blocks, num_cached, num_full, block_size, gid)
  decode    get_new_blocks(1) every block_size tokens; newly full blocks are hashed with cache_full_blocks
  finish    free_blocks(reversed(blocks))     (tail first, so a prefix outlives its suffixes in the LRU)
plus the rarely used entry points evict_blocks(set) and reset_prefix_cache().

A scenario is pure data (JSON-able), so the committed fixtures carry it next to the reference's answers.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np


class Request:
    """What cache_full_blocks reads from vLLM's Request: block_hashes (patches.py:436-439)."""

    def __init__(self, rid: int, block_hashes: List[bytes]):
        self.request_id = rid
        self.block_hashes = block_hashes


def make_scenario(seed: int, n_requests: int = 64, n_prefixes: int = 20, zipf_a: float = 1.1,
                  prefix_blocks=(2, 24), suffix_blocks=(0, 12), decode_blocks=(0, 6), concurrency: int = 8,
                  block_size: int = 16, group_id: int = 0, p_evict: float = 0.03, p_reset: float = 0.01,
                  p_str_hash: float = 0.0) -> dict:
    """20 shared prefixes x Zipf(1.1) reuse (SURVEY §8(d) T-mla-pin); hashes are 32-byte digests."""
    rng = np.random.default_rng(seed)
    pref = []
    for p in range(n_prefixes):
        nb = int(rng.integers(prefix_blocks[0], prefix_blocks[1] + 1))
        pref.append([rng.bytes(32).hex() for _ in range(nb)])
    w = 1.0 / np.arange(1, n_prefixes + 1) ** zipf_a
    w /= w.sum()
    reqs = []
    for r in range(n_requests):
        p = int(rng.choice(n_prefixes, p=w))
        cut = int(rng.integers(1, len(pref[p]) + 1))                 # shares a prefix of the prefix
        ns = int(rng.integers(suffix_blocks[0], suffix_blocks[1] + 1))
        hashes = pref[p][:cut] + [rng.bytes(32).hex() for _ in range(ns)]
        nd = int(rng.integers(decode_blocks[0], decode_blocks[1] + 1))
        # the last prompt block is usually partial: it is allocated but not hashed
        partial = bool(rng.integers(0, 2))
        reqs.append({"hashes": hashes, "partial": partial, "decode": nd,
                     "decode_hashes": [rng.bytes(32).hex() for _ in range(nd)]})
    # event list: arrivals interleaved with decode steps and finishes, bounded concurrency
    events = []
    running: List[int] = []
    left = {}
    nxt = 0
    while nxt < n_requests or running:
        can_arrive = nxt < n_requests and len(running) < concurrency
        u = rng.random()
        if can_arrive and (not running or u < 0.45):
            events.append(["arrive", nxt])
            running.append(nxt)
            left[nxt] = reqs[nxt]["decode"]
            nxt += 1
        else:
            r = running[int(rng.integers(0, len(running)))]
            if left[r] > 0:
                events.append(["decode", r])
                left[r] -= 1
            else:
                events.append(["finish", r])
                running.remove(r)
        v = rng.random()
        if v < p_evict and nxt > 0:
            events.append(["evict", int(rng.integers(0, nxt)), int(rng.integers(1, 6))])
        elif v < p_evict + p_reset:
            events.append(["reset"])
    return {"seed": seed, "block_size": block_size, "group_id": group_id, "requests": reqs, "events": events,
            "str_hash": p_str_hash > 0}


def drive_pool_groups(pool, scenario: dict, snap: Callable[[object], dict]) -> List[dict]:
    """The same flow for a model with several KV-cache groups (hybrid attention: full + sliding window): every request
    holds one block list per group, a prefix hit needs ALL groups (get_cached_block(hash, [g0, g1]) is all-or-nothing,
    patches.py:379-394), touch() gets the tuple-of-lists form (:523-529), keys carry the group id (:260-274)."""
    groups = scenario["groups"]
    bs = scenario["block_size"]
    reqs = scenario["requests"]
    blocks: Dict[int, list] = {}
    n_hashed: Dict[int, int] = {}
    robj: Dict[int, Request] = {}
    records = []
    for ev in scenario["events"]:
        kind, res = ev[0], None
        if kind == "arrive":
            r = ev[1]
            q = reqs[r]
            hashes = [bytes.fromhex(h) for h in q["hashes"]]
            req = Request(r, hashes + [bytes.fromhex(h) for h in q["decode_hashes"]])
            hits = [[] for _ in groups]
            for h in hashes:
                b = pool.get_cached_block(h, groups)
                if b is None:
                    break
                for g in range(len(groups)):
                    hits[g].append(b[g])
            n_hit = len(hits[0])
            n_prompt = len(hashes) + (1 if q["partial"] else 0)
            pool.touch(tuple(hits))
            per_group, ok = [], True
            for g in range(len(groups)):
                try:
                    per_group.append(hits[g] + pool.get_new_blocks(n_prompt - n_hit))
                except ValueError:
                    ok = False
                    break
            if not ok:
                for g, bl in enumerate(per_group):
                    pool.free_blocks(list(reversed(bl)))
                for g in range(len(per_group), len(groups)):
                    pool.free_blocks(list(reversed(hits[g])))
                res = "ERR:ValueError"
                blocks[r] = None
            else:
                for g, gid in enumerate(groups):
                    pool.cache_full_blocks(req, per_group[g], n_hit, len(hashes), bs, gid)
                blocks[r], n_hashed[r], robj[r] = per_group, len(hashes), req
                res = {"hit": n_hit, "ids": [[b.block_id for b in bl] for bl in per_group]}
        elif kind == "decode":
            r = ev[1]
            if blocks.get(r) is None:
                res = "skipped"
            else:
                try:
                    new_ids = []
                    for g, gid in enumerate(groups):
                        nb = pool.get_new_blocks(1)
                        blocks[r][g].extend(nb)
                        new_ids.append(nb[0].block_id)
                    full = min(len(blocks[r][0]) - 1, len(robj[r].block_hashes))
                    if full > n_hashed[r]:
                        for g, gid in enumerate(groups):
                            pool.cache_full_blocks(robj[r], blocks[r][g], n_hashed[r], full, bs, gid)
                        n_hashed[r] = full
                    res = {"ids": new_ids}
                except ValueError:
                    res = "ERR:ValueError"
        elif kind == "finish":
            r = ev[1]
            if blocks.get(r) is None:
                res = "skipped"
            else:
                for bl in blocks[r]:
                    pool.free_blocks(list(reversed(bl)))
                res = {"freed": sum(len(bl) for bl in blocks[r])}
                blocks[r] = None
        elif kind == "evict":
            r, k = ev[1], ev[2]
            bl = blocks.get(r)
            ids = [b.block_id for b in bl[0][:k]] if bl else list(range(1, 1 + k))
            pool.evict_blocks(set(ids))
            res = {"ids": ids}
        elif kind == "reset":
            res = bool(pool.reset_prefix_cache())
        else:
            raise ValueError(f"unknown event {ev}")
        records.append({"ev": ev, "r": res, **snap(pool)})
    return records


def drive_pool(pool, scenario: dict, snap: Callable[[object], dict], null_ok: bool = True) -> List[dict]:
    """Runs the scenario; returns one record per event: {"ev": [...], "r": result, **snap(pool)}."""
    if scenario.get("groups"):
        return drive_pool_groups(pool, scenario, snap)
    gid = scenario["group_id"]
    bs = scenario["block_size"]
    reqs = scenario["requests"]
    blocks: Dict[int, list] = {}
    n_hashed: Dict[int, int] = {}
    robj: Dict[int, Request] = {}
    records = []
    for ev in scenario["events"]:
        kind = ev[0]
        res = None
        if kind == "arrive":
            r = ev[1]
            q = reqs[r]
            hashes = [bytes.fromhex(h) for h in q["hashes"]]
            req = Request(r, hashes + [bytes.fromhex(h) for h in q["decode_hashes"]])
            hits = []
            for h in hashes:
                b = pool.get_cached_block(h, [gid])
                if b is None:
                    break
                hits.append(b[0] if isinstance(b, list) else b)
            n_prompt = len(hashes) + (1 if q["partial"] else 0)
            pool.touch(hits)
            try:
                new = pool.get_new_blocks(n_prompt - len(hits))
            except ValueError:
                pool.free_blocks(list(reversed(hits)))
                res = "ERR:ValueError"
                records.append({"ev": ev, "r": res, **snap(pool)})
                blocks[r] = None
                continue
            bl = hits + new
            pool.cache_full_blocks(req, bl, len(hits), len(hashes), bs, gid)
            blocks[r] = bl
            n_hashed[r] = len(hashes)
            robj[r] = req
            res = {"hit": len(hits), "ids": [b.block_id for b in bl]}
        elif kind == "decode":
            r = ev[1]
            if blocks.get(r) is None:
                res = "skipped"
            else:
                q = reqs[r]
                try:
                    nb = pool.get_new_blocks(1)
                    blocks[r].extend(nb)
                    # the block that just filled up (the old tail) gets its hash, when the request still has one
                    full = min(len(blocks[r]) - 1, len(robj[r].block_hashes))
                    if full > n_hashed[r]:
                        pool.cache_full_blocks(robj[r], blocks[r], n_hashed[r], full, bs, gid)
                        n_hashed[r] = full
                    res = {"ids": [b.block_id for b in nb]}
                except ValueError:
                    res = "ERR:ValueError"
        elif kind == "finish":
            r = ev[1]
            if blocks.get(r) is None:
                res = "skipped"
            else:
                pool.free_blocks(list(reversed(blocks[r])))
                res = {"freed": len(blocks[r])}
                blocks[r] = None
        elif kind == "evict":
            r, k = ev[1], ev[2]
            bl = blocks.get(r)
            ids = [b.block_id for b in bl[:k]] if bl else list(range(1, 1 + k))
            # evict_blocks takes a set: iterate it the way CPython does so that every implementation sees one order
            pool.evict_blocks(set(ids))
            res = {"ids": ids}
        elif kind == "reset":
            res = bool(pool.reset_prefix_cache())
        else:
            raise ValueError(f"unknown event {ev}")
        records.append({"ev": ev, "r": res, **snap(pool)})
    return records
