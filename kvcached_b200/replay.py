"""Trace replay through the device-resident manager (trace vocabulary: kvcached_b200/trace.py).

Three ways to run the same trace, all giving the reference's results:
  * replay_per_op      one launch per request — the reference's call pattern (alloc/free one at a time);
  * replay_host_steps  one launch per scheduler step through KVCacheManager.run_ops with HOST buffers:
                       freed ids come from the host's own copy of earlier results (what an engine's
                       scheduler holds), allocated ids are copied back every step;
  * replay_resident    the whole request list is uploaded once; frees name earlier requests by index and
                       the ids never leave HBM (kvb_mgr_trace_*).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import (OP_ALLOC, OP_AVAIL, OP_CLEAR, OP_FREE_IDS, OP_FREE_REF, OP_FREE_RESERVED, OP_RESIZE, OP_TRIM,
                   OP_TRY_RESERVE, ST_ASSERT, ST_FALSE, ST_NO_PAGES, ST_NONE, ST_OK, KvbError)
from .kv_cache_manager import OP_DTYPE, REF_DTYPE, RES_DTYPE, KVCacheManager

INT_MAX = 2 ** 31 - 1


def manager_from_config(cfg: dict, **kw) -> KVCacheManager:
    return KVCacheManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], cfg["num_layers"],
                          reserve_null_block=cfg.get("reserve_null_block", False),
                          num_kv_buffers=cfg.get("num_kv_buffers", 2),
                          page_size=cfg.get("page_size", 2 << 20),
                          min_reserved_pages=cfg.get("min_reserved_pages", 5),
                          max_reserved_pages=cfg.get("max_reserved_pages", 10), **kw)


# ----------------------------------------------------------------------------------------- per op
def _ids(r):
    return r if isinstance(r, list) else []


def apply_op(mgr: KVCacheManager, op, results):
    kind = op[0]
    if kind == "alloc":
        try:
            return mgr.alloc(op[1])
        except RuntimeError as e:
            if isinstance(e, KvbError):
                raise
            return "ERR:RuntimeError"
    if kind == "free_ids":
        mgr.free(list(op[1]))
    elif kind == "free_cat":
        mgr.free([b for k in op[1] for b in _ids(results[k])])
    elif kind == "free_slice":
        mgr.free(list(_ids(results[op[1]])[op[2]:op[3]]))
    elif kind == "resize":
        try:
            return bool(mgr.resize(op[1]))
        except AssertionError:
            return "ERR:AssertionError"
    elif kind == "trim":
        mgr.trim()
    elif kind == "try_reserve":
        try:
            return bool(mgr.try_to_reserve(op[1]))
        except RuntimeError as e:
            if isinstance(e, KvbError):
                raise
            return "ERR:RuntimeError"
    elif kind == "free_reserved":
        mgr.free_reserved()
    elif kind == "avail":
        return mgr.available_size()
    elif kind == "clear":
        mgr.clear()
    elif kind == "step":
        pass
    else:
        raise ValueError(f"unknown op {op}")
    return None


def replay_per_op(mgr: KVCacheManager, trace: dict, snapshots: bool = True) -> dict:
    results, records = [], []
    for op in trace["ops"]:
        r = apply_op(mgr, op, results)
        results.append(r)
        if snapshots:
            records.append({"r": r, **mgr.snapshot()})
    return {"records": records, "results": results, "null_block": mgr.null_block}


# ----------------------------------------------------------------------------------------- compiled
@dataclass
class CompiledTrace:
    ops: np.ndarray            # OP_DTYPE, one per request (the ["step"] markers are dropped)
    refs: np.ndarray           # REF_DTYPE, absolute request indices
    ids: np.ndarray            # int64 pool of FREE_IDS requests
    step_offsets: np.ndarray   # int64 [n_steps + 1] request index where each scheduler step starts
    req_of_op: List[int]       # trace op index -> request index (-1 for markers)
    op_of_req: List[int]       # request index -> trace op index (-1 for synthesised requests)
    alloc_blocks: int          # sum of all alloc sizes (id arena capacity)
    n_calls: int               # request-level alloc + free calls (the metric's "op")
    reserve_null_block: bool


def compile_trace(trace: dict) -> CompiledTrace:
    ops, refs, ids, steps = [], [], [], [0]
    req_of_op: List[int] = []
    op_of_req: List[int] = []
    null = bool(trace["config"].get("reserve_null_block", False))
    n_calls = 0
    alloc_blocks = 0

    def push(kind, n=0, arg=0, src=-1):
        ops.append((kind, n, arg))
        op_of_req.append(src)
        return len(ops) - 1

    for i, op in enumerate(trace["ops"]):
        k = op[0]
        if k == "step":
            req_of_op.append(-1)
            if steps[-1] != len(ops):
                steps.append(len(ops))
            continue
        if k == "alloc":
            req_of_op.append(push(OP_ALLOC, op[1], 0, i))
            alloc_blocks += max(int(op[1]), 0)
            n_calls += 1
        elif k == "free_ids":
            req_of_op.append(push(OP_FREE_IDS, len(op[1]), len(ids), i))
            ids.extend(op[1])
            n_calls += 1
        elif k == "free_cat":
            base = len(refs)
            for t in op[1]:
                refs.append((req_of_op[t], 0, INT_MAX))
            req_of_op.append(push(OP_FREE_REF, len(op[1]), base, i))
            n_calls += 1
        elif k == "free_slice":
            base = len(refs)
            refs.append((req_of_op[op[1]], op[2], op[3]))
            req_of_op.append(push(OP_FREE_REF, 1, base, i))
            n_calls += 1
        elif k == "resize":
            req_of_op.append(push(OP_RESIZE, 0, op[1], i))
        elif k == "trim":
            req_of_op.append(push(OP_TRIM, 0, 0, i))
        elif k == "try_reserve":
            req_of_op.append(push(OP_TRY_RESERVE, op[1], 0, i))
        elif k == "free_reserved":
            req_of_op.append(push(OP_FREE_RESERVED, 0, 0, i))
        elif k == "avail":
            req_of_op.append(push(OP_AVAIL, 0, 0, i))
        elif k == "clear":
            req_of_op.append(push(OP_CLEAR, 0, 0, i))
            if null:
                push(OP_ALLOC, 1, 0, -1)     # _reserve_null_block (kv_cache_manager.py:486)
                alloc_blocks += 1
        else:
            raise ValueError(f"unknown op {op}")
    if steps[-1] != len(ops):
        steps.append(len(ops))
    return CompiledTrace(np.array(ops, dtype=OP_DTYPE).reshape(-1), np.array(refs, dtype=REF_DTYPE).reshape(-1),
                         np.array(ids, dtype=np.int64), np.array(steps, dtype=np.int64), req_of_op, op_of_req,
                         alloc_blocks, n_calls, null)


def _decode_result(kind: int, st: int, ids: Optional[np.ndarray], avail: int):
    """kvb_res -> the value the reference call returns (same encoding as oracle.apply_op)."""
    if kind == OP_ALLOC:
        if st == ST_OK:
            return ids.tolist()
        if st == ST_NO_PAGES:
            return "ERR:RuntimeError"
        return None
    if kind == OP_RESIZE:
        return "ERR:AssertionError" if st == ST_ASSERT else st == ST_OK
    if kind == OP_TRY_RESERVE:
        return "ERR:RuntimeError" if st == ST_NO_PAGES else st == ST_OK
    if kind == OP_AVAIL:
        return int(avail)
    return None


def results_from_arrays(ct: CompiledTrace, res: np.ndarray, out: np.ndarray) -> list:
    """Per TRACE op results (markers -> None), for comparison with golden records."""
    outl = []
    for i, q in enumerate(ct.req_of_op):
        if q < 0:
            outl.append(None)
            continue
        r = res[q]
        kind = int(ct.ops["kind"][q])
        ids = out[r["off"]:r["off"] + r["n"]] if kind == OP_ALLOC and r["status"] == ST_OK else None
        outl.append(_decode_result(kind, int(r["status"]), ids, int(r["avail"])))
    return outl


# ----------------------------------------------------------------------------------------- resident
def upload(mgr: KVCacheManager, ct: CompiledTrace) -> None:
    mgr._wait_post_init()
    rc = mgr._lib.kvb_mgr_trace_upload(
        mgr._h, ct.ops.ctypes.data_as(C.POINTER(_lib.Op)), ct.ops.shape[0],
        ct.ids.ctypes.data_as(_lib.p_i64) if ct.ids.size else None, ct.ids.size,
        ct.refs.ctypes.data_as(C.POINTER(_lib.Ref)) if ct.refs.shape[0] else None, ct.refs.shape[0],
        ct.step_offsets.ctypes.data_as(_lib.p_i64), ct.step_offsets.size - 1, ct.alloc_blocks + 16)
    _lib.check(rc, "trace_upload")


def run_resident(mgr: KVCacheManager, ct: CompiledTrace, step_begin: int = 0, step_end: Optional[int] = None,
                 steps_per_launch: int = 1):
    """Returns (device_ms, launches)."""
    if step_end is None:
        step_end = ct.step_offsets.size - 1
    ms = C.c_float(0)
    nl = C.c_int64(0)
    _lib.check(mgr._lib.kvb_mgr_trace_run(mgr._h, step_begin, step_end, steps_per_launch, C.byref(ms), C.byref(nl)),
               "trace_run")
    return float(ms.value), int(nl.value)


def fetch_resident(mgr: KVCacheManager, ct: CompiledTrace):
    res = np.empty(ct.ops.shape[0], dtype=RES_DTYPE)
    out = np.empty(ct.alloc_blocks + 16, dtype=np.int64)
    n_out = C.c_int64(0)
    _lib.check(mgr._lib.kvb_mgr_trace_results(mgr._h, res.ctypes.data_as(C.POINTER(_lib.Res)), res.shape[0],
                                              out.ctypes.data_as(_lib.p_i64), out.size, C.byref(n_out)),
               "trace_results")
    return res, out[:n_out.value]


def replay_resident(mgr: KVCacheManager, trace: dict, steps_per_launch: int = 1) -> dict:
    ct = compile_trace(trace)
    upload(mgr, ct)
    mgr._lib.kvb_mgr_trace_reset_cursor(mgr._h)
    ms, launches = run_resident(mgr, ct, steps_per_launch=steps_per_launch)
    res, out = fetch_resident(mgr, ct)
    return {"results": results_from_arrays(ct, res, out), "device_ms": ms, "launches": launches, "res": res,
            "out": out, "ct": ct}


# ----------------------------------------------------------------------------------------- host steps
class HostStepPlan:
    """Per-launch request templates for replay_host_steps, prepared once outside any timed region:
    request kinds/sizes and, for every free, which slices of earlier results it concatenates.  The
    run-time side is vectorised numpy (the host keeps allocated ids in one arena + start/len per request,
    the way an engine keeps block tables), so the Python driver costs a few numpy calls per launch."""

    def __init__(self, ct: CompiledTrace):
        self.ct = ct
        self.steps = []
        # A launch can only free ids the host already holds: cut a step wherever a free names a
        # request of the same step (decode traces never do; the random fixtures do).
        bounds = []
        so = ct.step_offsets
        for s in range(so.size - 1):
            b, e = int(so[s]), int(so[s + 1])
            start = b
            for q in range(b, e):
                if int(ct.ops["kind"][q]) == OP_FREE_REF:
                    a, n = int(ct.ops["arg"][q]), int(ct.ops["n"][q])
                    if n and int(ct.refs["op"][a:a + n].max()) >= start:
                        bounds.append((start, q))
                        start = q
            bounds.append((start, e))
        for b, e in bounds:
            if e <= b:
                continue
            ops = ct.ops[b:e].copy()
            free_local, ref_q, ref_lo, ref_hi, owner_start = [], [], [], [], []
            fixed = []          # (position in the free list, ids) for FREE_IDS taken from the trace itself
            for j in range(e - b):
                k = int(ops["kind"][j])
                if k == OP_FREE_REF:
                    a, n = int(ops["arg"][j]), int(ops["n"][j])
                    if n == 0:
                        ops["kind"][j] = OP_FREE_IDS
                        ops["n"][j] = 0
                        continue
                    owner_start.append(len(ref_q))
                    free_local.append(j)
                    ref_q.extend(int(x) for x in ct.refs["op"][a:a + n])
                    ref_lo.extend(int(x) for x in ct.refs["lo"][a:a + n])
                    ref_hi.extend(int(x) for x in ct.refs["hi"][a:a + n])
                    ops["kind"][j] = OP_FREE_IDS
                elif k == OP_FREE_IDS:
                    a, n = int(ops["arg"][j]), int(ops["n"][j])
                    fixed.append((j, ct.ids[a:a + n].copy()))
            alloc_local = np.nonzero(ops["kind"] == OP_ALLOC)[0]
            self.steps.append(dict(
                b=b, ops=ops, free_local=np.array(free_local, dtype=np.int64),
                ref_q=np.array(ref_q, dtype=np.int64), ref_lo=np.array(ref_lo, dtype=np.int64),
                ref_hi=np.array(ref_hi, dtype=np.int64), owner_start=np.array(owner_start, dtype=np.int64),
                fixed=fixed, alloc_local=alloc_local, alloc_global=alloc_local + b,
                need=int(ops["n"][alloc_local].sum())))


class HeldIds:
    """Host copy of every request's allocated ids: one arena + (start, len) per request."""

    def __init__(self, n_req: int, capacity: int):
        self.arena = np.empty(max(capacity, 1), dtype=np.int64)
        self.start = np.zeros(n_req, dtype=np.int64)
        self.len = np.zeros(n_req, dtype=np.int64)
        self.ok = np.zeros(n_req, dtype=bool)
        self.cur = 0

    def __getitem__(self, q: int):
        return self.arena[self.start[q]:self.start[q] + self.len[q]] if self.ok[q] else None


def replay_host_steps(mgr: KVCacheManager, plan: HostStepPlan, keep_results: bool = True, after_step=None):
    """One kvb_mgr_step per scheduler step with host buffers in and out.  Returns (res_all, held) where
    held[req] is the host copy of request req's ids (None for failed / non-alloc requests)."""
    ct = plan.ct
    n = ct.ops.shape[0]
    held = HeldIds(n, ct.alloc_blocks + 16)
    arena, h_start, h_len = held.arena, held.start, held.len
    res_all = np.empty(n, dtype=RES_DTYPE) if keep_results else None
    for st in plan.steps:
        ops = st["ops"]
        ids = None
        pieces = []
        pos = 0
        if st["ref_q"].size:
            q = st["ref_q"]
            ln = h_len[q]
            lo = np.minimum(st["ref_lo"], ln)
            cnt = np.minimum(st["ref_hi"], ln) - lo
            np.maximum(cnt, 0, out=cnt)
            csum = np.cumsum(cnt)
            total = int(csum[-1])
            if total:
                idx = np.repeat(h_start[q] + lo - (csum - cnt), cnt) + np.arange(total, dtype=np.int64)
                pieces.append(arena[idx])
            per_free = np.add.reduceat(cnt, st["owner_start"])
            # reduceat quirk: an empty trailing group cannot occur (every owner has >= 1 ref)
            ops["n"][st["free_local"]] = per_free
            ops["arg"][st["free_local"]] = np.cumsum(per_free) - per_free
            pos = total
        for j, arr in st["fixed"]:
            ops["arg"][j] = pos
            ops["n"][j] = arr.size
            pieces.append(arr)
            pos += arr.size
        if pos:
            ids = pieces[0] if len(pieces) == 1 else np.concatenate(pieces)
        res, out = mgr.run_ops(ops, ids, None, st["need"])
        b = st["b"]
        if keep_results:
            res_all[b:b + ops.shape[0]] = res
        al = st["alloc_local"]
        if al.size:
            r = res[al]
            ok = r["status"] == ST_OK
            ag = st["alloc_global"]
            cur = held.cur
            if out.size:
                arena[cur:cur + out.size] = out
            h_start[ag] = cur + r["off"]
            h_len[ag] = np.where(ok, r["n"], 0)
            held.ok[ag] = ok
            held.cur = cur + out.size
        if after_step is not None:
            after_step()
    return res_all, held


def results_from_host_steps(ct: CompiledTrace, res_all: np.ndarray, held) -> list:
    outl = []
    for i, q in enumerate(ct.req_of_op):
        if q < 0:
            outl.append(None)
            continue
        r = res_all[q]
        kind = int(ct.ops["kind"][q])
        outl.append(_decode_result(kind, int(r["status"]), held[q], int(r["avail"])))
    return outl
