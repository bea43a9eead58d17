"""Synthetic alloc/free traces for the elastic KV-cache path.

Pure numpy, seeded, implementation-independent: a trace never names block ids it
could only know by running an allocator.  Frees refer to earlier ops by index, so
the same trace replays through the reference (oracle/ref_replay.py), the CPU
oracle (oracle/oracle.py) and the device-resident manager.

Op vocabulary (JSON-friendly lists):
  ["alloc", n] ["free_ids", [ids]] ["free_cat", [op, ...]] ["free_slice", op, a, b]
  ["resize", bytes] ["trim"] ["try_reserve", n] ["free_reserved"] ["avail"] ["clear"] ["step"]
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

PAGE = 2 << 20

# KV geometries of BASELINE.json's configs; bf16, block 16.
LAYOUTS: Dict[str, dict] = {
    # C1/C2  Llama-3.1-8B: 32 layers, 8 KV heads, d=128
    "llama8b": dict(num_layers=32, block_size=16, cell_size=8 * 128 * 2, num_kv_buffers=2),
    # C3     Qwen2.5-7B: 28 layers, 4 KV heads, d=128
    "qwen7b": dict(num_layers=28, block_size=16, cell_size=4 * 128 * 2, num_kv_buffers=2),
    # C4     DeepSeek-V3 MLA latent (512+64) bf16, SGLang page_size 1 and 64
    "mla_p1": dict(num_layers=61, block_size=1, cell_size=576 * 2, num_kv_buffers=1),
    "mla_p64": dict(num_layers=61, block_size=64, cell_size=576 * 2, num_kv_buffers=1),
    # C5     Llama-3.3-70B: 80 layers, 8 KV heads sharded over TP ranks
    "llama70b_tp1": dict(num_layers=80, block_size=16, cell_size=8 * 128 * 2, num_kv_buffers=2),
    "llama70b_tp2": dict(num_layers=80, block_size=16, cell_size=4 * 128 * 2, num_kv_buffers=2),
    "llama70b_tp4": dict(num_layers=80, block_size=16, cell_size=2 * 128 * 2, num_kv_buffers=2),
    "llama70b_tp8": dict(num_layers=80, block_size=16, cell_size=1 * 128 * 2, num_kv_buffers=2),
    # tests/test_kvcache_manager.py fixture of the reference: 16 layers, 8 heads x 64, fp16
    "ref_fixture": dict(num_layers=16, block_size=16, cell_size=1024, num_kv_buffers=2),
}


def make_config(layout: str, num_blocks: int, **over) -> dict:
    cfg = dict(LAYOUTS[layout])
    cfg.update(num_blocks=int(num_blocks), page_size=PAGE, reserve_null_block=False,
               prealloc=False, contiguous=True, layout=layout)
    cfg.update(over)
    return cfg


def blocks_per_page(cfg: dict) -> int:
    return cfg.get("page_size", PAGE) // (cfg["block_size"] * cfg["cell_size"])


# --------------------------------------------------------------------------- random mix
def random_trace(cfg: dict, seed: int, n_ops: int = 300, max_alloc: int = 200,
                 p_free: float = 0.45, extra: bool = True, p_clear: float = 0.0) -> dict:
    """Random allocs, partial/whole/concatenated frees and (optionally) resize / trim /
    reserve events.  Sizes are chosen against a block-count model so most allocs succeed
    and some deliberately over-ask (-> None).  p_clear > 0 also emits KVCacheManager.clear()
    (kv_cache_manager.py:444-489): everything is freed, ids held from before are never freed again.
    (The extra random draw happens only then, so traces with p_clear = 0 are what they always were.)"""
    rng = np.random.default_rng(seed)
    ops: List[list] = []
    live: List[list] = []          # [op_index, lo, hi] un-freed ranges, assuming the alloc succeeded
    used = 1 if cfg.get("reserve_null_block") else 0
    cap = cfg["num_blocks"]
    for _ in range(n_ops):
        if p_clear > 0 and rng.random() < p_clear:
            ops.append(["clear"])
            live = []
            used = 1 if cfg.get("reserve_null_block") else 0
            continue
        u = rng.random()
        if extra and u < 0.03:
            ops.append(["avail"])
        elif extra and u < 0.05:
            ops.append(["trim"])
        elif extra and u < 0.07:
            n = int(rng.integers(1, 40))
            ops.append(["try_reserve", n])
        elif extra and u < 0.09:
            ops.append(["free_reserved"])
        elif extra and u < 0.11:
            pages = cap * cfg["block_size"] * cfg["cell_size"] // cfg.get("page_size", PAGE)
            # never above the initial size: the virtual tensors are only that large, and growing past
            # the high-water mark after a shrink makes the reference hand out duplicate page ids
            # (csrc/page_allocator.cpp:343-346 appends ids from num_total_pages, not from the mark)
            newp = int(rng.integers(max(1, pages // 4), pages + 1))
            ops.append(["resize", newp * cfg.get("page_size", PAGE)])
        elif live and u < 0.11 + p_free:
            mode = rng.random()
            if mode < 0.4:                      # free one whole allocation (what is left of it)
                k = int(rng.integers(len(live)))
                op_i, lo, hi = live.pop(k)
                ops.append(["free_slice", op_i, lo, hi])
                used -= hi - lo
            elif mode < 0.8:                    # free a random slice of one allocation
                k = int(rng.integers(len(live)))
                op_i, lo, hi = live[k]
                a = int(rng.integers(lo, hi))
                b = int(rng.integers(a + 1, hi + 1))
                ops.append(["free_slice", op_i, a, b])
                used -= b - a
                rest = [r for r in ([op_i, lo, a], [op_i, b, hi]) if r[2] > r[1]]
                live[k:k + 1] = rest
            else:                               # free several whole allocations in one call
                whole = [r for r in live if r[1] == 0 and r[2] == _alloc_len(ops, r[0])]
                if len(whole) >= 2:
                    m = int(rng.integers(2, min(len(whole), 6) + 1))
                    pick = [whole[i] for i in rng.choice(len(whole), size=m, replace=False)]
                    ops.append(["free_cat", [r[0] for r in pick]])
                    for r in pick:
                        live.remove(r)
                        used -= r[2] - r[1]
                else:
                    ops.append(["avail"])
        else:
            if rng.random() < 0.04:
                n = int(cap + rng.integers(1, 50))      # certainly too many -> None
                ops.append(["alloc", n])
            else:
                n = int(min(max_alloc, max(1, rng.geometric(1.0 / 24))))
                ops.append(["alloc", n])
                if used + n <= cap * 0.9:       # the model only tracks allocs that surely fit
                    live.append([len(ops) - 1, 0, n])
                    used += n
                else:
                    # may or may not succeed; free whatever it returned right away
                    ops.append(["free_cat", [len(ops) - 1]])
    return {"config": cfg, "ops": ops}


def _alloc_len(ops, i) -> int:
    return ops[i][1] if ops[i][0] == "alloc" else -1


# --------------------------------------------------------------------------- micro
def micro_trace(cfg: dict, k: int, iters: int, churn: bool) -> dict:
    """alloc(k); free(h) pairs as in the reference's benchmarks/bench_alloc/bench_alloc.py:57-62.
    churn=True: the pool is empty between pairs (every pair takes and releases a page);
    churn=False: one pinned block keeps the page alive."""
    ops: List[list] = []
    if not churn:
        ops.append(["alloc", 1])
    for _ in range(iters):
        ops.append(["alloc", k])
        ops.append(["free_cat", [len(ops) - 1]])
        ops.append(["step"])
    return {"config": cfg, "ops": ops}


# --------------------------------------------------------------------------- decode trace
def decode_trace(cfg: dict, n_requests: int = 4096, seed: int = 0, target_peak: float = 0.65,
                 prompt_mu: float = 6.5, prompt_sigma: float = 0.8, out_mu: float = 5.5,
                 out_sigma: float = 0.7, tokens_per_step: int = 1) -> dict:
    """T-decode: `n_requests` requests, prompt ~ LogNormal(6.5, 0.8) clipped to
    [16, 8192] tokens, output ~ LogNormal(5.5, 0.7) clipped to [16, 2048]; Poisson arrivals on a
    triangular ramp 0 -> peak -> 0 sized so the pool climbs to ~`target_peak` of its blocks and
    drains to zero.  One scheduler step = `tokens_per_step` generated tokens per running request.

    Per step, in order: free(all blocks of every request that finished in the previous step, in
    allocation order), one alloc(1) for every running request whose next token crosses a block
    boundary (request-id order), then alloc(ceil(prompt/block)) for each admitted arrival.
    An arrival waits in a FIFO while the block-count model says it would not fit.
    """
    rng = np.random.default_rng(seed)
    bs = cfg["block_size"]
    cap = cfg["num_blocks"] - (1 if cfg.get("reserve_null_block") else 0)
    prompt = np.clip(np.rint(rng.lognormal(prompt_mu, prompt_sigma, n_requests)), 16, 8192).astype(np.int64)
    out = np.clip(np.rint(rng.lognormal(out_mu, out_sigma, n_requests)), 16, 2048).astype(np.int64)

    life = out.mean() / tokens_per_step                         # steps a request stays
    live_blocks = ((prompt + out / 2.0) / bs).mean()            # average blocks it holds
    lam_peak = max(target_peak * cap / (live_blocks * life), 1e-3)
    window = max(int(math.ceil(2.0 * n_requests / lam_peak)), 4)
    # arrival step of each request: inverse CDF of the triangular density on [0, window]
    u = np.sort(rng.random(n_requests))
    half = window / 2.0
    t_arr = np.where(u < 0.5, half * np.sqrt(2 * u), window - half * np.sqrt(2 * (1 - u)))
    t_arr = np.floor(t_arr).astype(np.int64)

    ops: List[list] = []
    alloc_ops: List[List[int]] = [[] for _ in range(n_requests)]   # op indices holding each request's blocks
    tokens = np.zeros(n_requests, dtype=np.int64)
    have = np.zeros(n_requests, dtype=np.int64)
    gen = np.zeros(n_requests, dtype=np.int64)
    running: List[int] = []
    waiting: List[int] = []
    finished_prev: List[int] = []
    used = 0
    peak_used = 0
    nxt = 0
    step = 0
    n_alloc = n_free = 0
    while nxt < n_requests or running or waiting or finished_prev:
        for r in finished_prev:                                    # (1) frees
            ops.append(["free_cat", alloc_ops[r]])
            used -= int(have[r])
            n_free += 1
        finished_prev = []
        still: List[int] = []
        for r in running:                                          # (2) decode growth
            for _ in range(tokens_per_step):
                if gen[r] >= out[r]:
                    break
                tokens[r] += 1
                gen[r] += 1
                need = -(-int(tokens[r]) // bs)
                if need > have[r]:
                    if used + 1 > cap:                             # pool exhausted: stall this request
                        tokens[r] -= 1
                        gen[r] -= 1
                        break
                    ops.append(["alloc", 1])
                    alloc_ops[r].append(len(ops) - 1)
                    have[r] += 1
                    used += 1
                    n_alloc += 1
            if gen[r] >= out[r]:
                finished_prev.append(r)
            else:
                still.append(r)
        running = still
        while nxt < n_requests and t_arr[nxt] <= step:             # (3) arrivals
            waiting.append(nxt)
            nxt += 1
        while waiting:
            r = waiting[0]
            nb = -(-int(prompt[r]) // bs)
            headroom = len(running) + 8                            # keep room for the next decode step
            if used + nb + headroom > cap:
                break
            waiting.pop(0)
            ops.append(["alloc", nb])
            alloc_ops[r].append(len(ops) - 1)
            tokens[r] = prompt[r]
            have[r] = nb
            used += nb
            n_alloc += 1
            running.append(r)
        peak_used = max(peak_used, used)
        ops.append(["step"])
        step += 1
        if step > 10_000_000:
            raise RuntimeError("decode_trace did not terminate")
    meta = dict(n_requests=n_requests, seed=seed, steps=step, n_alloc=n_alloc, n_free=n_free,
                n_ops=n_alloc + n_free, peak_blocks=int(peak_used), peak_frac=peak_used / max(cap, 1),
                total_blocks=int(sum(len(a) for a in alloc_ops)),
                tokens_per_step=tokens_per_step)
    return {"config": cfg, "ops": ops, "meta": meta}


def split_steps(ops: List[list]) -> List[List[int]]:
    """Indices of the ops of each scheduler step (the ["step"] markers themselves excluded)."""
    steps: List[List[int]] = []
    cur: List[int] = []
    for i, op in enumerate(ops):
        if op[0] == "step":
            steps.append(cur)
            cur = []
        else:
            cur.append(i)
    if cur:
        steps.append(cur)
    return steps
