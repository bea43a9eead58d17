"""Parity of the device-resident manager (sm_100a step kernel, through the C ABI) with the UNMODIFIED
reference on every committed fixture, and with the CPU oracle on fresh seeded traces.

Bit-exact bar: returned block ids, available_size, page counters, avail/full page sets in dict order,
reserved blocks, shrink flag and the mapped-page set after EVERY request.  Bookkeeping-only managers
(attach_vmm=False) are used here so that pools larger than the box's memory can be replayed; the VMM
side is covered by test_vmm_gpu.py.
"""
import numpy as np
import pytest

import oracle
from golden_util import diff_record, golden_names, load_golden
from kvcached_b200 import replay as R
from kvcached_b200 import trace as T

pytestmark = pytest.mark.gpu

BIG = 1 << 40


def _mgr(cfg):
    m = R.manager_from_config(cfg, attach_vmm=False)
    m.set_phys_free_pages(BIG)
    return m


@pytest.mark.parametrize("name", golden_names())
def test_per_op_matches_reference_fixture(name):
    fix = load_golden(name)
    m = _mgr(fix["trace"]["config"])
    assert m.null_block == fix["null_block"]
    results = []
    for i, (op, want) in enumerate(zip(fix["trace"]["ops"], fix["records"])):
        r = R.apply_op(m, op, results)
        results.append(r)
        got = {"r": r, **m.snapshot()}
        bad = diff_record(got, want)
        assert not bad, (f"{name}: op {i} {op[:2]} differs in {bad}: got { {k: got[k] for k in bad} } "
                         f"want { {k: want[k] for k in bad} }")
    used, inuse, empties, mism, _ = m.audit()
    assert mism == 0
    m.close()


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("mode", ["resident", "resident_x4", "host_steps"])
def test_batched_matches_reference_fixture(name, mode):
    """Whole scheduler steps per launch must give what the sequential reference calls gave."""
    fix = load_golden(name)
    trace = fix["trace"]
    m = _mgr(trace["config"])
    if mode.startswith("resident"):
        out = R.replay_resident(m, trace, steps_per_launch=4 if mode.endswith("x4") else 1)
        results = out["results"]
    else:
        ct = R.compile_trace(trace)
        res_all, held = R.replay_host_steps(m, R.HostStepPlan(ct))
        results = R.results_from_host_steps(ct, res_all, held)
    for i, (op, want) in enumerate(zip(trace["ops"], fix["records"])):
        if op[0] == "avail" and mode.startswith("resident"):
            pass   # same value: the kernel evaluates available_size() in sequence
        assert results[i] == want["r"], f"{name}/{mode}: op {i} {op[:2]}: got {results[i]} want {want['r']}"
    got = m.snapshot()
    bad = [k for k in diff_record({"r": None, **got}, {**fix["records"][-1], "r": None})]
    assert not bad, f"{name}/{mode}: final state differs in {bad}: got { {k: got[k] for k in bad} } want { {k: fix['records'][-1][k] for k in bad} }"
    assert m.audit()[3] == 0
    m.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("layout,num_blocks,kw", [
    ("llama8b", 8192, {}),
    ("llama8b", 8192, {"reserve_null_block": True}),
    ("mla_p1", 40000, {"reserve_null_block": True}),
    ("mla_p64", 1500, {}),
    ("llama70b_tp8", 20000, {}),
    ("qwen7b", 6000, {}),
])
def test_fresh_random_traces_vs_oracle(layout, num_blocks, kw, seed):
    """New traces the fixtures do not hold: the oracle (itself pinned to the reference by
    test_oracle_golden.py) is the checker."""
    cfg = T.make_config(layout, num_blocks, **kw)
    trace = T.random_trace(cfg, 9000 + seed, n_ops=400, max_alloc=max(8, num_blocks // 12))
    want = oracle.replay(trace)
    m = _mgr(cfg)
    results = []
    for i, (op, w) in enumerate(zip(trace["ops"], want["records"])):
        r = R.apply_op(m, op, results)
        results.append(r)
        got = {"r": r, **m.snapshot()}
        bad = diff_record(got, w)
        assert not bad, f"op {i} {op[:2]} differs in {bad}: got { {k: got[k] for k in bad} } want { {k: w[k] for k in bad} }"
    m.close()
    m2 = _mgr(cfg)
    out = R.replay_resident(m2, trace)
    assert out["results"] == want["results"]
    m2.close()


def test_decode_trace_full_size_vs_oracle():
    """BASELINE config C2 at full size (91 520 blocks, 4096 requests): ids of every request identical to
    the oracle's, final state identical, conservation of available_size."""
    cfg = T.make_config("llama8b", 91520)
    trace = T.decode_trace(cfg, 4096, seed=0)
    want = oracle.replay(trace, snapshots=False)
    m = _mgr(cfg)
    a0 = m.available_size()
    out = R.replay_resident(m, trace)
    assert out["results"] == want["results"]
    assert m.available_size() == a0          # everything allocated was freed
    s = m.stats()
    assert s.num_inuse_pages == 0 and s.num_avail_blocks == 0
    assert m.audit()[:4] == (0, 0, 0, 0)
    m.close()
    # the same through per-step host buffers
    m2 = _mgr(cfg)
    ct = R.compile_trace(trace)
    res_all, held = R.replay_host_steps(m2, R.HostStepPlan(ct))
    assert R.results_from_host_steps(ct, res_all, held) == want["results"]
    m2.close()


def test_events_match_oracle_order():
    cfg = T.make_config("llama8b", 4096, num_layers=4)
    trace = T.random_trace(cfg, 4242, n_ops=300)
    om = oracle.manager_from_config(cfg)
    m = _mgr(cfg)
    ro, rm = [], []
    for op in trace["ops"]:
        ro.append(oracle.apply_op(om, op, ro))
        rm.append(R.apply_op(m, op, rm))
        assert m.drain_events() == om.drain_events()
    m.close()


def test_group_indices_by_page_device_vs_libstdcxx():
    cfg = T.make_config("llama8b", 65536)
    m = _mgr(cfg)
    rng = np.random.default_rng(5)
    cases = [[5, 70, 6, 200, 130, 71, 700, 1, 579], [p * 64 for p in range(20)],
             [p * 64 + j for p in range(20) for j in range(64)]]
    for n in (1, 2, 17, 64, 500, 4096, 20000):
        cases.append(rng.choice(65536, size=n, replace=False).tolist())
        base = int(rng.integers(0, 65536 - n))
        cases.append(list(range(base, base + n)))
    for ids in cases:
        assert m.group_indices_by_page(ids) == oracle.group_indices_by_page(ids, 32768), f"n={len(ids)}"
    m.close()
    # 33..256 distinct pages (the multi-key warp replay of the hash order), few and many ids per page, small and
    # large reserve() hints, every rehash count
    m = _mgr(cfg)
    for D in (33, 34, 47, 48, 64, 97, 98, 128, 199, 200, 255, 256, 257, 300):
        for per_page in (1, 2, 7):
            pages = rng.choice(1024, size=D, replace=False)
            ids = (pages[:, None] * 64 + rng.integers(0, 64, size=(D, 1)) + np.arange(per_page)[None, :] * 0).reshape(-1)
            if per_page > 1:
                ids = np.concatenate([pages * 64 + j for j in range(per_page)])
                rng.shuffle(ids)
            ids = ids.tolist()
            assert m.group_indices_by_page(ids) == oracle.group_indices_by_page(ids, 32768), f"D={D} per_page={per_page}"
    # clustered keys (many collisions per bucket) and a request-like id list: one long run + scattered singles
    for D in (40, 100, 250):
        ids = [int(p) * 64 * 23 % 65536 for p in range(D)]
        ids = list(dict.fromkeys(ids))
        assert m.group_indices_by_page(ids) == oracle.group_indices_by_page(ids, 32768), f"clustered D={D}"
        ids = list(range(5000, 5000 + 41)) + [int(x) * 64 + 3 for x in rng.choice(900, size=D, replace=False)]
        assert m.group_indices_by_page(ids) == oracle.group_indices_by_page(ids, 32768), f"request-like D={D}"
    m.close()
    # MLA geometry: blocks straddle page edges, 1820 per page
    cfg = T.make_config("mla_p1", 60000)
    m = _mgr(cfg)
    for n in (3, 100, 5000, 30000):
        ids = rng.choice(60000, size=n, replace=False).tolist()
        assert m.group_indices_by_page(ids) == oracle.group_indices_by_page(ids, 1152)
    m.close()


def test_step_api_equals_sequential_calls():
    cfg = T.make_config("llama8b", 8192)
    a, b = _mgr(cfg), _mgr(cfg)
    rng = np.random.default_rng(3)
    held_a, held_b = [], []
    for _ in range(60):
        k_free = int(rng.integers(0, min(4, len(held_a)) + 1))
        frees_idx = sorted(rng.choice(len(held_a), size=k_free, replace=False).tolist(), reverse=True) if k_free else []
        fa = [held_a.pop(i) for i in frees_idx]
        fb = [held_b.pop(i) for i in frees_idx]
        needs = rng.integers(1, 90, size=int(rng.integers(1, 12))).tolist()
        got = a.step(fa, needs)
        for f in fb:
            b.free(f)
        want = [b.alloc(n) for n in needs]
        assert [None if g is None else g.tolist() for g in got] == want
        held_a += [g for g in got if g is not None]
        held_b += [w for w in want if w is not None]
        assert a.available_size() == b.available_size()
    assert a.snapshot() == b.snapshot()
    a.close(); b.close()


def test_over_allocation_returns_none_and_config_error():
    from kvcached_b200 import KVCachedConfigError, KVCacheManager
    m = _mgr(T.make_config("llama8b", 256, num_layers=2))
    assert m.alloc(257) is None
    x = m.alloc(256)
    assert x == list(range(256)) and m.available_size() == 0 and m.alloc(1) is None
    m.free(x)
    assert m.available_size() == 256
    m.close()
    with pytest.raises(KVCachedConfigError):
        KVCacheManager(16, 16, (4 << 20) // 16, 2, attach_vmm=False)


@pytest.mark.parametrize("seed", [0, 1])
def test_free_of_many_pages_per_call_vs_oracle(seed):
    """free() calls of <= 512 ids that name 33 .. 256 distinct pages: the shared-memory free path with the block-wide
    replay of libstdc++'s unordered_map order (ht_order_prefix + ht_order_block).  The avail_pages order after every
    call IS that order, so the snapshot comparison pins it; ids per page vary so that some pages empty out."""
    cfg = T.make_config("llama8b", 64 * 300)          # 300 pages of 64 blocks
    rng = np.random.default_rng(seed)
    om = oracle.manager_from_config(cfg)
    m = _mgr(cfg)
    held = om.alloc(64 * 280)
    assert m.alloc(64 * 280) == held
    held = np.array(held, dtype=np.int64)
    alive = np.ones(held.size, dtype=bool)
    for it, D in enumerate([33, 34, 47, 48, 64, 97, 100, 128, 199, 200, 255, 256, 40, 250]):
        pages = rng.choice(280, size=D, replace=False)
        ids = []
        for p in pages:
            cand = np.nonzero(alive[p * 64:(p + 1) * 64])[0]
            if cand.size == 0:
                continue
            k = min(cand.size, int(rng.integers(1, 3)) if D > 200 else int(rng.integers(1, 6)))
            take = rng.choice(cand, size=k, replace=False) + p * 64
            ids.extend(int(x) for x in take)
        ids = [int(held[i]) for i in rng.permutation(ids)][:512]
        for i in ids:
            alive[np.nonzero(held == i)[0][0]] = False
        om.free(ids)
        m.free(ids)
        got, want = m.snapshot(), om.snapshot()
        bad = diff_record({"r": None, **got}, {"r": None, **want})
        assert not bad, f"call {it} (D={D}, n={len(ids)}) differs in {bad}: got { {k: got[k] for k in bad} } want { {k: want[k] for k in bad} }"
        # and the allocation that follows consumes the pages in that order
        n = int(rng.integers(1, 40))
        a, b = m.alloc(n), om.alloc(n)
        assert a == b
        for i in a:
            j = np.nonzero(held == i)[0]
            if j.size:
                alive[j[0]] = True
            else:
                held = np.append(held, i)
                alive = np.append(alive, True)
    assert m.audit()[3] == 0
    m.close()
