#!/usr/bin/env python
"""bench.py — the elastic KV-cache hot path on B200: KV block alloc+free ops/s on the 4096-request
decode trace (BASELINE.json configs[1]), page map+unmap latency against the raw driver, and fused KV
scatter GB/s against the measured HBM peak.  One JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A bench "step" = one full replay of the trace (T-decode-4096: 89 218 request-level alloc/free calls in
3 475 scheduler steps, pool grows to 61 % of 91 520 blocks and drains to zero).

  value   whole-job ops/s with the request list resident in HBM (frees name earlier requests, ids never
          leave the device), one kernel launch per scheduler step, timed with CUDA events on the
          launching stream.  Bookkeeping kernels only.
  e2e     the same trace through KVCacheManager.run_ops with HOST buffers every scheduler step: freed
          ids copied host->device, allocated ids device->host, and the kernel's page map/unmap list
          applied to the real virtual KV tensors (cuMemMap loop) before the call returns.
  --impl reference   the UNMODIFIED reference KVCacheManager (oracle/_ref, built from /root/reference by
          oracle/build_ref.py) replaying the same trace on this box's host cores against cuda:0,
          i.e. with its own cuMemCreate/cuMemMap work, prealloc thread on (shipping default).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kv_block_alloc_free_ops_per_s"
UNIT = "ops/s"
NUM_BLOCKS = 91520          # Llama-3.1-8B pool on a 180 GB B200
NUM_BLOCKS_TP = 65536       # Llama-3.3-70B TP pool (BASELINE configs[4]): 80 layers, 8 KV heads sharded over ranks
N_REQUESTS = 4096
PAGE = 2 << 20
SCATTER_T = 65536           # tokens per fused all-layer scatter launch (16 GiB of traffic > L2)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE kvb_kv_copy_kernel<scatter> launch at exactly this shape
# (T = 65 536 tokens, 32 layers, cell 2048 B), ncu --set full, profiles/r2_kv_copy_cell2048.csv (round 1 measured
# 17.136 GB with the previous kernel: profiles/r1_kv_copy_T65536_full.csv)
SCATTER_TRAFFIC_NCU = 8_593_535_000 + 8_547_023_000


def workload_config(world: int) -> dict:
    if world > 1:
        return {"workload": f"T-decode-4096 under TP={world}: Llama-3.3-70B KV layout (80 layers, 8 KV heads sharded "
                            f"{8 // world} per rank, d=128, bf16, block 16, 2 MiB pages, contiguous layout), 4096 requests, "
                            "seed 0; rank 0 runs the block manager, its step kernel pushes each step's page list to the other "
                            "ranks through a shared-memory mailbox (--tp-transport nccl: one NCCL broadcast per step)",
                "num_blocks": NUM_BLOCKS_TP, "requests": N_REQUESTS, "tp": world,
                "l2": "scatter inputs exceed L2; a 256 MiB buffer is written between trace replays"}
    return {"workload": "T-decode-4096: Llama-3.1-8B KV layout (32 layers, 8 KV heads, d=128, bf16, block 16, "
                        "2 MiB pages, contiguous layout), 4096 requests, prompt~LogNormal(6.5,0.8), "
                        "output~LogNormal(5.5,0.7), seed 0",
            "num_blocks": NUM_BLOCKS, "requests": N_REQUESTS, "tp": world,
            "l2": "scatter inputs (16 GiB/launch) exceed L2; a 256 MiB buffer is written between trace replays"}


def build_trace(world: int = 1):
    from kvcached_b200 import trace as T
    if world > 1:
        cfg = T.make_config(f"llama70b_tp{world}", NUM_BLOCKS_TP)
    else:
        cfg = T.make_config("llama8b", NUM_BLOCKS)
    tr = T.decode_trace(cfg, N_REQUESTS, seed=0)
    return cfg, tr


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.p = None
        self.rows = []
        self.gpu = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------- reference arm
REF_CORES = "0,1"     # BASELINE.md 3.3 pins the reference's scheduler thread; its prealloc thread gets the second core


def run_reference(trace: dict, repeats: int, device: str = "cuda:0", prealloc: bool = True, timeout: int = 1500,
                  contiguous: bool = True):
    """Replays through the unmodified reference (oracle/_ref) in a subprocess; returns its JSON or raises."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    import shutil
    if not build_ref.built():
        if not build_ref.build():
            raise RuntimeError("oracle/_ref is not built and /root/reference is not present")
    tr = json.loads(json.dumps({"config": dict(trace["config"], prealloc=prealloc, contiguous=contiguous),
                                "ops": trace["ops"]}))
    env = build_ref.ref_env(cpu_only=(device == "cpu" and not os.path.exists("/dev/nvidiactl")))
    env["KVCACHED_LOG_LEVEL"] = "ERROR"
    with tempfile.TemporaryDirectory() as d:
        tin, tout = os.path.join(d, "t.json"), os.path.join(d, "o.json")
        with open(tin, "w") as f:
            json.dump(tr, f)
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_replay.py"), tin, tout, "--time",
               "--device", device, "--repeat", str(repeats)]
        pinned = False
        if shutil.which("taskset") and (os.cpu_count() or 1) >= 2:
            cmd = ["taskset", "-c", REF_CORES] + cmd
            pinned = True
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0 or not os.path.isfile(tout):
            raise RuntimeError(f"reference replay failed (rc={r.returncode}): {r.stderr[-1500:]}")
        with open(tout) as f:
            out = json.load(f)
        out["pinned"] = f"taskset -c {REF_CORES}" if pinned else "not pinned (no taskset)"
        return out


def oracle_port_ops_per_s(trace: dict) -> dict:
    """Fallback CPU baseline: the C++ oracle port, single thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    t0 = time.perf_counter()
    oracle.replay(trace, snapshots=False)
    dt = time.perf_counter() - t0
    n = sum(1 for o in trace["ops"] if o[0] in ("alloc", "free_ids", "free_cat", "free_slice"))
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"one replay of the full trace through oracle/kv_oracle.cpp ({n} calls, {dt:.3f} s, incl. Python glue)"}


REF_SAMPLE_DIV = {1: 4}     # N -> the reference replays 1 / div of the workload per step (default 8)


def reference_sample_trace(world: int):
    """The reference's step is a BOUNDED SAMPLE of the workload: the same generator and distributions with 1 / div of
    the requests on a pool 1 / div the size, i.e. the same page events and the same list lengths per call, so ops/s is
    comparable with the full trace while K + W steps end within minutes on the slowest boxes of the pool (a full replay
    through the reference takes 9 - 45 s here, the driver asks for 25 of them).  N = 1: 1024 requests, contiguous
    layout.  N > 1: 512 requests; the reference cannot create the 70B geometry's contiguous tensors (ftensor.cpp:161
    asserts that its fixed base address is a multiple of the 320 MiB compound page), so it runs its per-layer layout —
    160 driver map loops per page event."""
    from kvcached_b200 import trace as T
    div = REF_SAMPLE_DIV.get(world, 8)
    if world == 1:
        cfg = T.make_config("llama8b", NUM_BLOCKS // div)
    else:
        cfg = T.make_config(f"llama70b_tp{world}", NUM_BLOCKS_TP // div)
    return cfg, T.decode_trace(cfg, N_REQUESTS // div, seed=0)


def reference_sample_note(world: int) -> str:
    div = REF_SAMPLE_DIV.get(world, 8)
    nb = (NUM_BLOCKS if world == 1 else NUM_BLOCKS_TP) // div
    return (f"replays of a bounded sample of the workload ({N_REQUESTS // div} requests, pool {nb} blocks: same generator, "
            "distributions and page events per call as the full trace)")


def reference_main(args, rank: int, world: int):
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    cfg, tr = reference_sample_trace(world)
    sample = reference_sample_note(world)
    if world == 1:
        attempts = [("cuda:0", True, True, "cuda:0, contiguous layout, prealloc thread on (shipping default)")]
    else:
        attempts = [("cuda:0", True, False, "cuda:0, per-layer layout (the contiguous 70B geometry trips the reference's "
                                            "own alignment assert, ftensor.cpp:161), prealloc thread on")]
    attempts.append(("cpu", False, True, "cpu device (bookkeeping only: map/unmap are no-ops)"))
    n_calls = tr["meta"]["n_ops"]
    out = dev = None
    for device, prealloc, contiguous, what in attempts:
        try:
            out = run_reference(tr, repeats=K + W, device=device, prealloc=prealloc, contiguous=contiguous)
            dev = what
            break
        except Exception as e:
            sys.stderr.write(f"[bench] reference on {what} failed ({e}); trying the next configuration\n")
    if out is None:
        raise RuntimeError("the reference could not replay the trace in any configuration")
    el = out["elapsed_all_s"][W:]
    tot = sum(el)
    v = n_calls * len(el) / tot
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * tot / len(el), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": workload_config(world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 2, "kind": "reference",
                             "sample": f"{len(el)} {sample} ({n_calls} calls each) through the unmodified reference "
                                       f"KVCacheManager, {dev}; single scheduler thread + its prealloc thread, "
                                       f"{out.get('pinned')}, host has {os.cpu_count()} cores"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- extra legs (rank 0)
MICRO_KS = ((1, 3000), (4, 3000), (16, 3000), (64, 2000), (256, 1000))


def micro_alloc_free_leg(cfg: dict, local_rank: int, with_reference: bool) -> dict:
    from kvcached_b200 import replay as R
    out = {"unit": "us per alloc(k)+free pair", "api": "KVCacheManager.alloc(k) -> list, KVCacheManager.free(list)"}
    mm = R.manager_from_config(cfg, attach_vmm=False, device_index=local_rank)
    mm.set_phys_free_pages(1 << 40)
    for variant in ("churn", "no_churn"):
        pinned = mm.alloc(1) if variant == "no_churn" else None
        for kk, iters in MICRO_KS:
            for _ in range(100):
                mm.free(mm.alloc(kk))
            mm.sync()
            t0 = time.perf_counter()
            for _ in range(iters):
                mm.free(mm.alloc(kk))
            mm.sync()
            out[f"{variant}_k{kk}"] = (time.perf_counter() - t0) / iters * 1e6
        if pinned:
            mm.free(pinned)
    mm.close()
    if with_reference:
        try:
            ops, sections = [], []
            for variant in ("churn", "no_churn"):
                if variant == "no_churn":
                    ops.append(["alloc", 1])
                for kk, iters in MICRO_KS:
                    for _ in range(100):
                        ops.append(["alloc", kk])
                        ops.append(["free_cat", [len(ops) - 1]])
                    ops.append(["mark"])
                    for _ in range(iters):
                        ops.append(["alloc", kk])
                        ops.append(["free_cat", [len(ops) - 1]])
                    ops.append(["mark"])
                    sections.append((f"{variant}_k{kk}", iters))
            ref = {}
            for dev, key in (("cuda:0", "reference_cuda0"), ("cpu", "reference_bookkeeping_only")):
                o = run_reference({"config": cfg, "ops": ops}, repeats=1, device=dev, prealloc=(dev != "cpu"), timeout=600)
                mk = o["marks"]
                ref[key] = {name: (mk[2 * i + 1] - mk[2 * i]) / iters * 1e6 for i, (name, iters) in enumerate(sections)}
                ref[key]["pinned"] = o.get("pinned")
            out.update(ref)
            out["ours_vs_reference_cuda0"] = {name: out[name] / ref["reference_cuda0"][name] for name, _ in sections}
        except Exception as e:
            out["reference_error"] = str(e)[:300]
    return out


def per_call_e2e_leg(mgr, tr: dict, n_calls: int, vmm_ops) -> dict:
    """The decode trace, one alloc()/free() call per request with Python lists in and out (replay.replay_per_op)."""
    from kvcached_b200 import replay as R
    times = []
    for it in range(3):
        vmm_ops.settle()
        t0 = time.perf_counter()
        R.replay_per_op(mgr, tr, snapshots=False)
        mgr.sync()
        vmm_ops.fence()
        times.append(time.perf_counter() - t0)
    assert mgr.stats().num_inuse_pages == 0
    dt = sum(times[1:]) / len(times[1:])
    return {"value": n_calls / dt, "unit": UNIT, "ms_per_replay": dt * 1e3, "replays_s": times,
            "api": "KVCacheManager.alloc(n) / free(ids) once per request (lists), attach_vmm=True, default policy; "
                   "free() does not wait for its kernel, alloc() does; fence inside the timed region"}


def mla_pin_leg(local_rank: int, with_reference: bool) -> dict:
    """ElasticBlockPool (prefix-cache pinning on the device, kvcached_b200/block_pool.py) vs the unmodified reference
    pool over the unmodified reference manager, same scenario, bookkeeping only on both sides (the reference cannot
    create the 61-layer contiguous tensors: ftensor.cpp:161 asserts that its base address is a multiple of the 122 MiB
    compound page)."""
    from kvcached_b200 import pool_trace as PT
    from kvcached_b200 import trace as T
    from kvcached_b200.block_pool import ElasticBlockPool
    from kvcached_b200.kv_cache_manager import KVCacheManager
    out = {"unit": "pool events/s (arrive / decode / finish / evict / reset of the vLLM-style flow)"}
    for name, layout, bs, blocks, reqs, pre, suf, dec in (
            ("page64", "mla_p64", 64, 28 * 400, 1500, (2, 24), (0, 12), (0, 6)),
            ("page1", "mla_p1", 1, 1820 * 40, 400, (64, 768), (0, 256), (0, 64))):
        cfg = T.make_config(layout, blocks)
        cap = 16000 // bs
        sc = PT.make_scenario(seed=0, n_requests=reqs, n_prefixes=20, zipf_a=1.1, prefix_blocks=pre, suffix_blocks=suf,
                              decode_blocks=dec, concurrency=32, block_size=bs)
        pm = KVCacheManager(blocks, bs, cfg["cell_size"], cfg["num_layers"], num_kv_buffers=1, attach_vmm=False,
                            max_cached_blocks=cap, device_index=local_rank)
        pm.set_phys_free_pages(1 << 40)
        pool = ElasticBlockPool(blocks, bs, cfg["cell_size"], cfg["num_layers"], True, num_kv_buffers=1,
                                max_cached_blocks=cap, kv_cache_manager=pm)
        t0 = time.perf_counter()
        recs = PT.drive_pool(pool, sc, lambda p: {})
        dt = time.perf_counter() - t0
        row = {"events": len(recs), "ours_events_per_s": len(recs) / dt, "ours_s": dt, "block_size": bs, "cap_blocks": cap,
               "refused": sum(1 for r in recs if r["r"] == "ERR:ValueError"),
               "evictable_at_end": len(pool.evictable_ids())}
        pm.close()
        if with_reference:
            try:
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import build_ref
                env = build_ref.ref_env(cpu_only=not os.path.exists("/dev/nvidiactl"))
                env["KVCACHED_LOG_LEVEL"] = "ERROR"
                fix = {"config": dict(cfg, enable_caching=True, max_cached_blocks=cap), "scenario": sc, "timing": True}
                with tempfile.TemporaryDirectory() as d:
                    tin, tout = os.path.join(d, "s.json"), os.path.join(d, "o.json")
                    with open(tin, "w") as f:
                        json.dump(fix, f)
                    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_pool_replay.py"), tin, tout],
                                       env=env, capture_output=True, text=True, timeout=600)
                    if r.returncode != 0:
                        raise RuntimeError(r.stderr[-400:])
                    with open(tout) as f:
                        o = json.load(f)
                row["reference_events_per_s"] = o["n_events"] / o["elapsed_s"]
                row["reference_s"] = o["elapsed_s"]
                row["same_outcome"] = (o["n_events"] == len(recs) and o["n_errors"] == row["refused"]
                                       and o["evictable_at_end"] == row["evictable_at_end"])
                row["ours_vs_reference"] = row["ours_events_per_s"] / row["reference_events_per_s"]
            except Exception as e:
                row["reference_error"] = str(e)[:300]
        out[name] = row
    return out


# --------------------------------------------------------------------------------------- init time and the per-layer layout
def init_and_layout_leg(local_rank: int, cfg: dict, ref_create_s) -> dict:
    """create_kv_tensors wall time (VA reservation + zero-page init, ftensor.cpp:160-176) for the contiguous and the
    per-layer layout, eager and background; and page map / unmap of the per-layer layout (64 driver units per page)
    against the raw driver loop.  The caller has shut the main group down."""
    import numpy as np
    from kvcached_b200 import _lib, vmm_ops
    lib = _lib.load()
    L = cfg["num_layers"]
    per_layer_full = cfg["num_blocks"] * cfg["block_size"] * cfg["cell_size"] * 2
    out = {"reference_contiguous_create_s": ref_create_s}

    def create(contiguous: bool, per_layer: int, background: bool) -> dict:
        vmm_ops.zero_init_mode(1 if background else 0)
        vmm_ops.init_kvcached(f"cuda:{local_rank}", PAGE, contiguous)
        t0 = time.perf_counter()
        vmm_ops.create_kv_tensors_raw(per_layer, L, 2)
        t_create = time.perf_counter() - t0
        row = {"create_s": t_create, "units": vmm_ops.init_progress()["units"]}
        if background:
            # the first pages are usable long before the init is through
            ids = np.arange(40, 44, dtype=np.int64)
            t0 = time.perf_counter()
            _lib.check(lib.kvb_map_pages(0, ids.ctypes.data_as(_lib.p_i64), ids.size), "map_pages")
            row["first_4_pages_mapped_after_s"] = t_create + time.perf_counter() - t0
            t0 = time.perf_counter()
            vmm_ops.wait_init()
            row["fully_backed_after_s"] = t_create + time.perf_counter() - t0 + row["first_4_pages_mapped_after_s"] - t_create
            _lib.check(lib.kvb_unmap_pages(0, ids.ctypes.data_as(_lib.p_i64), ids.size), "unmap_pages")
        row["us_per_unit"] = vmm_ops.init_progress()["init_ms"] * 1e3 / max(row["units"], 1)
        return row

    try:
        first = create(True, per_layer_full, False)
        vmm_ops.shutdown_kvcached()
        out["contiguous_eager"] = create(True, per_layer_full, False)
        out["contiguous_eager"]["create_s_first_run"] = first["create_s"]   # the driver has slow phases: both runs shown
        vmm_ops.shutdown_kvcached()
        out["contiguous_background"] = create(True, per_layer_full, True)
        vmm_ops.shutdown_kvcached()
        # per-layer: 64 units per page; an eighth of the pool keeps the eager leg at a few seconds (the cost is per unit)
        pl = per_layer_full // 8 // (2 * PAGE) * (2 * PAGE)
        out["per_layer_eager_eighth_pool"] = create(False, pl, False)
        out["per_layer_eager_full_pool_extrapolated_s"] = out["per_layer_eager_eighth_pool"]["create_s"] * per_layer_full / pl
        # page map / unmap in the per-layer layout vs the raw loop (same 64 driver units per page)
        mu = {}
        for n in (1, 4):
            iters = 10 if n == 1 else 6
            lib.kvb_vmm_pool_config(0, 64)
            _lib.check(lib.kvb_vmm_pool_prefill(0, n + 2), "pool_prefill")
            ids = np.arange(10, 10 + n, dtype=np.int64)
            p = ids.ctypes.data_as(_lib.p_i64)
            mm, uu = [], []
            for i in range(iters + 2):
                t0 = time.perf_counter()
                rc_m = lib.kvb_map_pages(0, p, n)
                t1 = time.perf_counter()
                rc_u = lib.kvb_unmap_pages(0, p, n)
                t2 = time.perf_counter()
                assert rc_m == 1 and rc_u == 1, _lib.last_error()
                if i >= 2:
                    mm.append((t1 - t0) * 1e6)
                    uu.append((t2 - t1) * 1e6)
            lib.kvb_vmm_pool_config(0, 16)
            a, b = np.zeros(iters), np.zeros(iters)
            _lib.check(lib.kvb_vmm_raw_control(0, n, iters, 0, a.ctypes.data_as(_lib.p_f64), b.ctypes.data_as(_lib.p_f64)), "raw")
            mu[f"n{n}"] = {"map_p50_us": pctl(mm, 0.5), "unmap_p50_us": pctl(uu, 0.5), "raw_map_p50_us": pctl(a.tolist(), 0.5),
                           "raw_unmap_p50_us": pctl(b.tolist(), 0.5),
                           "map_plus_unmap_vs_raw": (pctl(mm, 0.5) + pctl(uu, 0.5)) / (pctl(a.tolist(), 0.5) + pctl(b.tolist(), 0.5)),
                           "driver_units_per_page": 2 * L}
        out["per_layer_map_unmap"] = mu
        vmm_ops.shutdown_kvcached()
        out["per_layer_background_eighth_pool"] = create(False, pl, True)
        vmm_ops.shutdown_kvcached()
    except Exception as e:
        out["error"] = str(e)[:500]
        try:
            vmm_ops.shutdown_kvcached()
        except Exception:
            pass
    vmm_ops.zero_init_mode(-1)
    return out


# --------------------------------------------------------------------------------------- T-two-model (C3)
TWO_MODELS = {"A": ("llama8b", NUM_BLOCKS), "B": ("qwen7b", 192000)}     # Llama-3.1-8B + Qwen2.5-7B (BASELINE configs[2])


def _wait_file(path: str, timeout: float = 300.0) -> None:
    t_end = time.time() + timeout
    while not os.path.exists(path):
        if time.time() > t_end:
            raise TimeoutError(f"timed out waiting for {path}")
        time.sleep(0.0005)


def two_model_worker(role: str, d: str, local_rank: int) -> None:
    """One engine PROCESS of the co-location scenario (examples/01_simple_two_models of the reference): its own virtual
    KV tensors, manager, VMM thread and usage segment; the only coupling with the other process is the device's free
    memory as cudaMemGetInfo reports it (page_allocator.cpp:442-455) — nothing is injected."""
    import numpy as np
    import torch
    from kvcached_b200 import replay as R
    from kvcached_b200 import trace as T
    from kvcached_b200 import vmm_ops
    from kvcached_b200.kv_cache_manager import shm_read
    torch.cuda.set_device(local_rank)
    torch.cuda.init()
    layout, blocks = TWO_MODELS[role]
    cfg = T.make_config(layout, blocks)
    vmm_ops.init_kvcached(f"cuda:{local_rank}", PAGE, True)
    per_layer = cfg["num_blocks"] * cfg["block_size"] * cfg["cell_size"] * 2
    vmm_ops.create_kv_tensors_raw(per_layer, cfg["num_layers"], 2)
    ipc = f"kvb_two_{os.getpid()}_{role}"
    mgr = R.manager_from_config(cfg, device_index=local_rank, ipc_name=ipc, publish_shm=True)
    tr = T.decode_trace(cfg, 2048, seed=1 if role == "A" else 2, target_peak=0.35)
    ct = R.compile_trace(tr)
    plan = R.HostStepPlan(ct)
    out = {"role": role, "layout": layout, "num_blocks": blocks, "ipc_name": mgr.shm_name(), "calls": tr["meta"]["n_ops"],
           "page_mib": vmm_ops.vmm_stats()["page_bytes"] >> 20}
    R.replay_host_steps(mgr, plan, keep_results=False)          # warm-up replay, alone-ish
    vmm_ops.fence()
    open(os.path.join(d, f"ready_{role}"), "w").close()
    _wait_file(os.path.join(d, "go1"))
    # ---- phase 1: both models replay their decode trace at the same time
    times, nones = [], 0
    for _ in range(3):
        t0 = time.perf_counter()
        res_all, _held = R.replay_host_steps(mgr, plan, keep_results=True)
        vmm_ops.fence()
        times.append(time.perf_counter() - t0)
        nones += int((res_all["status"] == -1).sum())
    out["corun_ops_per_s"] = tr["meta"]["n_ops"] * len(times) / sum(times)
    out["corun_replays_s"] = times
    out["corun_allocs_refused"] = nones
    assert mgr.stats().num_inuse_pages == 0
    # ---- kvctl limit: the parent writes a smaller total into this model's usage segment; the next alloc applies it
    open(os.path.join(d, f"phase1_{role}"), "w").close()
    _wait_file(os.path.join(d, f"limited_{role}"))
    total0 = mgr.stats().num_total_pages
    ids = mgr.alloc(1)
    mgr.free(ids)
    out["limit"] = {"total_pages_before": int(total0), "total_pages_after": int(mgr.stats().num_total_pages),
                    "segment": shm_read(mgr.shm_name())}
    open(os.path.join(d, f"limit_seen_{role}"), "w").close()
    _wait_file(os.path.join(d, f"unlimited_{role}"))
    ids = mgr.alloc(1)
    mgr.free(ids)
    out["limit"]["total_pages_after_raising_to_three_quarters"] = int(mgr.stats().num_total_pages)
    # ---- phase 2: memory freed by A becomes allocatable by B
    bpp = mgr.stats().blocks_per_page
    if role == "A":
        # 80 % of what this process can get right now (B, the parent bench process and both handle pools hold the rest)
        want_pages = int(0.8 * mgr.available_size()) // bpp
        ids = mgr.alloc_np(want_pages * bpp)
        assert ids is not None and want_pages > 16, (want_pages, mgr.available_size())
        vmm_ops.fence()
        out["handover"] = {"held_pages": want_pages, "held_gib": want_pages * vmm_ops.vmm_stats()["page_bytes"] / 2**30}
        open(os.path.join(d, "A_holds"), "w").close()
        _wait_file(os.path.join(d, "B_polling"))
        time.sleep(0.3)
        t0 = time.time()
        mgr.free(ids)
        mgr.trim()                   # settle point: every page goes back to the device before trim returns
        vmm_ops.fence()
        t1 = time.time()
        out["handover"].update(t_free_start=t0, t_free_done=t1, free_trim_fence_ms=(t1 - t0) * 1e3)
        with open(os.path.join(d, "A_freed"), "w") as f:
            json.dump({"t0": t0, "t1": t1, "pages": want_pages}, f)
    else:
        _wait_file(os.path.join(d, "A_holds"))
        a_before = mgr.available_size()
        open(os.path.join(d, "B_polling"), "w").close()
        series = []
        t_end = time.time() + 60
        while time.time() < t_end:
            series.append((time.time(), mgr.available_size()))
            if os.path.exists(os.path.join(d, "A_freed")) and len(series) > 50 and series[-1][1] == series[-50][1] \
                    and series[-1][1] > a_before:
                break
            time.sleep(0.0002)
        time.sleep(0.05)
        with open(os.path.join(d, "A_freed")) as f:
            af = json.load(f)
        a_after = mgr.available_size()
        gained = a_after - a_before
        t_half = next((t for t, a in series if a >= a_before + 0.5 * gained), None)
        t_90 = next((t for t, a in series if a >= a_before + 0.9 * gained), None)
        # ... and B can really take it: allocate (and map) 90 % of what appeared
        n_take = int(0.9 * gained)
        t0 = time.perf_counter()
        ids = mgr.alloc_np(n_take) if n_take > 0 else None
        vmm_ops.fence()
        t_alloc = time.perf_counter() - t0
        out["handover"] = {"available_blocks_before": int(a_before), "available_blocks_after": int(a_after),
                           "gained_gib": gained / bpp * vmm_ops.vmm_stats()["page_bytes"] / 2**30,
                           "ms_from_A_free_start_to_half_visible": None if t_half is None else (t_half - af["t0"]) * 1e3,
                           "ms_from_A_free_start_to_90pct_visible": None if t_90 is None else (t_90 - af["t0"]) * 1e3,
                           "ms_from_A_free_done_to_90pct_visible": None if t_90 is None else (t_90 - af["t1"]) * 1e3,
                           "B_alloc_and_map_of_90pct_ok": ids is not None, "B_alloc_and_map_ms": t_alloc * 1e3,
                           "polls": len(series)}
        if ids is not None:
            mgr.free(ids)
            mgr.trim()
    with open(os.path.join(d, f"result_{role}.json"), "w") as f:
        json.dump(out, f)
    mgr.close()
    vmm_ops.shutdown_kvcached()


def two_model_leg(local_rank: int) -> dict:
    """T-two-model: two engine processes on ONE GPU, live cudaMemGetInfo, `kvctl limit` through the usage segment, and
    the time for memory freed by one model to become allocatable by the other (default policy: retention off)."""
    from kvcached_b200.kv_cache_manager import shm_read, shm_set_limit
    d = tempfile.mkdtemp(prefix="kvb_two_")
    procs = {}
    errs = {}
    try:
        for role in ("A", "B"):
            errs[role] = open(os.path.join(d, f"err_{role}.log"), "w")
            procs[role] = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--two-model-worker", role, d,
                                            str(local_rank)], stdout=errs[role], stderr=subprocess.STDOUT)

        def alive_wait(path, timeout=300.0):
            t_end = time.time() + timeout
            while not os.path.exists(path):
                for r, p in procs.items():
                    if p.poll() not in (None, 0):
                        raise RuntimeError(f"worker {r} exited {p.returncode}: " + open(os.path.join(d, f"err_{r}.log")).read()[-1500:])
                if time.time() > t_end:
                    raise TimeoutError(path)
                time.sleep(0.01)

        for role in ("A", "B"):
            alive_wait(os.path.join(d, f"ready_{role}"))
        open(os.path.join(d, "go1"), "w").close()
        for role in ("A", "B"):
            alive_wait(os.path.join(d, f"phase1_{role}"))
        # kvctl limit (kvcached/cli/utils.py:116-141 writes total_size into the segment): a quarter of the pool, then back
        import glob
        for role in ("A", "B"):
            seg = [os.path.basename(x) for x in glob.glob(f"/dev/shm/kvb_two_{procs[role].pid}_{role}*")]
            name = seg[0]
            total = shm_read(name)[0]
            shm_set_limit(name, total // 4)
            open(os.path.join(d, f"limited_{role}"), "w").close()
            alive_wait(os.path.join(d, f"limit_seen_{role}"))
            # back up to 3/4: a total EQUAL to the size the pool was created with means "no change" to the reference's
            # watcher (it compares with the constructor's mem_size_per_layer_, page_allocator.cpp:772-774) and to ours
            shm_set_limit(name, total // 4 * 3)
            open(os.path.join(d, f"unlimited_{role}"), "w").close()
        for role in ("A", "B"):
            alive_wait(os.path.join(d, f"result_{role}.json"), timeout=300)
        out = {}
        for role in ("A", "B"):
            with open(os.path.join(d, f"result_{role}.json")) as f:
                out[role] = json.load(f)
        for p in procs.values():
            p.wait(timeout=60)
        out["coupling"] = "two processes, one GPU; each reads the device's free memory live (VMM thread sample, <= 5 ms old)"
        return out
    except Exception as e:
        return {"error": str(e)[:1500]}
    finally:
        for p in procs.values():
            if p.poll() is None:
                p.kill()
        for f in errs.values():
            f.close()


# --------------------------------------------------------------------------------------- our arm
def pctl(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


def ours_main(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch
    K, W = args.steps, max(args.warmup, 1)   # default 3; fewer only for profiling runs, never for reported numbers
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    cfg, tr = build_trace(world)
    n_calls = tr["meta"]["n_ops"]

    # ---- CPU baseline first (rank 0, N=1 only): the reference needs the GPU memory we are about to take
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _scfg, str_ = reference_sample_trace(1)
            s_calls = str_["meta"]["n_ops"]
            out = run_reference(str_, repeats=1 + args.ref_replays, device="cuda:0", prealloc=True, timeout=900)
            el = out["elapsed_all_s"][1:]
            cpu_baseline = {"value": s_calls * len(el) / sum(el), "unit": UNIT, "cores": 2, "kind": "reference",
                            "replays_s": el, "create_kv_tensors_s": out.get("create_kv_tensors_s"),
                            "sample": f"1 warm-up + {len(el)} timed {reference_sample_note(1)} ({s_calls} calls each) "
                                      f"through the unmodified reference KVCacheManager on cuda:0 (real cuMemMap work, "
                                      f"prealloc thread on = shipping default), {out.get('pinned')}; host has "
                                      f"{os.cpu_count()} cores, the scheduler path is single-threaded"}
            try:   # prealloc off = the deterministic parity mode
                o3 = run_reference(str_, repeats=1, device="cuda:0", prealloc=False, timeout=600)
                cpu_baseline["prealloc_off_value"] = s_calls / o3["elapsed_all_s"][-1]
            except Exception as e:
                cpu_baseline["prealloc_off_value"] = f"failed: {e}"[:200]
            try:   # "cpu" device: map/unmap are no-ops, i.e. the reference's bookkeeping alone
                o2 = run_reference(str_, repeats=2, device="cpu", prealloc=False, timeout=600)
                cpu_baseline["bookkeeping_only_value"] = s_calls / o2["elapsed_all_s"][-1]
            except Exception:
                pass
        except Exception as e:
            sys.stderr.write(f"[bench] reference baseline unavailable ({e}); timing the oracle port\n")
            cpu_baseline = oracle_port_ops_per_s(tr)

    from kvcached_b200 import _lib, kv_ops, vmm_ops
    from kvcached_b200 import replay as R
    from kvcached_b200.kv_cache_manager import KVCacheManager
    lib = _lib.load()
    launches0 = lib.kvb_total_launches()
    torch.cuda.init()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ct = R.compile_trace(tr)
    n_sched_steps = ct.step_offsets.size - 1

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()

    # ================= value: resident trace, bookkeeping kernels, CUDA-event timed =================
    m = R.manager_from_config(cfg, attach_vmm=False, device_index=local_rank)
    m.set_phys_free_pages(1 << 40)
    R.upload(m, ct)
    dev_ms = []
    for it in range(W + K):
        lib.kvb_mgr_trace_reset_cursor(m._h)
        flush_buf.fill_(it & 0xff)
        barrier()
        ms, nl = R.run_resident(m, ct, steps_per_launch=1)
        barrier()
        if it >= W:
            dev_ms.append(ms)
    assert m.stats().num_inuse_pages == 0, "trace must drain the pool"
    resident_launches = nl
    page_events_per_replay = m.stats().events / (W + K)
    m.close()
    t_res = torch.tensor([sum(dev_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
    res_ms_total = float(t_res.item())
    value = n_calls * K / (res_ms_total * 1e-3)

    # ================= e2e: host buffers every scheduler step + real VMM ================================
    # KVCacheManager.run_ops per scheduler step with HOST buffers; maps in place (on every TP rank) when a step returns,
    # kvb_vmm_fence() inside the timed region.  Two legs:
    #   e2e (headline)      the library's DEFAULT policy: retention off — EVERY page event of the bookkeeping is applied
    #                       to the device inside the timed region (what the reference physically does per free,
    #                       page_allocator.cpp:264-310, :648-686); after the fence the physical mapping equals the
    #                       bookkeeping's (+ <= 5 pages mapped ahead).  Unmapped handles wait in the bounded pool.
    #   e2e.with_retention  opt-in KVB200_VMM_RETAIN_MS=2000: unmapped pages keep their handle for a while (<= 10 % of
    #                       the device), most events cost no driver call; reported with its physical overshoot.
    vmm_ops.init_kvcached(f"cuda:{local_rank}", PAGE, True)
    per_layer = cfg["num_blocks"] * cfg["block_size"] * cfg["cell_size"] * 2
    ranges = vmm_ops.create_kv_tensors_raw(per_layer, cfg["num_layers"], 2)
    # attach_vmm=True (kvb_init was called).  Under TP only rank 0's manager is driven; the others follow its pages.
    mgr = R.manager_from_config(cfg, device_index=local_rank)
    plan = R.HostStepPlan(ct)
    tp = None
    TP = None
    if world > 1:
        from kvcached_b200 import tp as TP
        if args.tp_transport == "nccl":
            tp = TP.TpPageSync(mgr if rank == 0 else None, group_id=0, rank=rank, world=world, device=dev)
        else:   # the step kernel pushes page lists into a shared ring; followers' threads apply them (kvb_tp.h)
            tp = TP.TpMailbox(TP.mailbox_name(), rank, world, mgr=mgr if rank == 0 else None)
        barrier()

    def tp_fence():
        """Every rank has applied every page list rank 0 produced so far (outside per-step paths)."""
        if tp is None:
            return
        if args.tp_transport == "nccl":
            tp.sync()
            return
        if rank == 0:
            tp.fence()
        barrier()                # rank 0 is done publishing
        if rank != 0:
            tp.fence()

    # The e2e legs replay the whole trace through the driver (0.3 - 7 s per replay on this pool): at most 6 timed
    # replays after at most 3 warm-up ones, whatever K and W ask of the device-timed value.
    Ke, We = min(K, 6), min(W, 3)

    def e2e_leg(strict: bool) -> dict:
        K, W = Ke, We
        vmm_ops.set_retention(0.0 if strict else 2000.0)
        pooled_peak = 0
        vmm_ops.settle()
        barrier()
        times, over_fence, over_peak, ret_peak = [], [], [], []
        vt0 = b0 = None
        for it in range(W + K):
            if it == W:
                s0 = mgr.stats()
                b0 = (s0.h2d_bytes, s0.d2h_bytes)
                vt0 = vmm_ops.vmm_timing()
            flush_buf.fill_(it & 0xff)
            vmm_ops.footprint()          # reset the peaks
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                R.replay_host_steps(mgr, plan, keep_results=False)
            else:
                TP.replay_host_steps_tp(tp, mgr, plan, rank)
                tp_fence()               # every rank has applied everything rank 0 published
            vmm_ops.fence()              # every queued unmap is applied inside the timed region
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            fp = vmm_ops.footprint()
            pooled_peak = max(pooled_peak, fp["pooled_pages"])
            if it >= W:
                times.append(dt)
                over_fence.append(fp["physical_pages"] - fp["wanted_pages"])
                over_peak.append(fp["peak_overshoot_pages"])
                ret_peak.append(fp["peak_retained_pages"])
        s1 = mgr.stats()
        vt1 = vmm_ops.vmm_timing()
        t = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
        page_gib = vmm_ops.vmm_stats()["page_bytes"] / 2**30
        leg = {"value": n_calls * K / total, "unit": UNIT, "ms_per_step": 1e3 * total / K,
               "h2d_bytes_per_step": (s1.h2d_bytes - b0[0]) / K, "d2h_bytes_per_step": (s1.d2h_bytes - b0[1]) / K,
               "page_transitions_per_replay": (vt1["unmap_calls"] - vt0["unmap_calls"]) / K,
               "driver_ms_per_replay": sum(vt1[k] - vt0[k] for k in ("unmap_ms", "map_ms", "setaccess_ms", "create_ms",
                                                                      "release_ms")) / K,
               "physical_overshoot_pages_after_fence_max": max(over_fence),
               "physical_overshoot_pages_peak_max": max(over_peak),
               "physical_overshoot_pages_peak_mean": sum(over_peak) / len(over_peak),
               "physical_overshoot_gib_peak_max": max(over_peak) * page_gib,
               "retained_pages_peak_max": max(ret_peak),
               "pooled_unmapped_handles_pages_after_fence_max": pooled_peak, "pool_cap_pages": fp["pool_cap_pages"],
               "steps": K, "warmup": W}
        # the PHYSICAL mapping (everything mapped, not only what is wanted) against the bookkeeping
        mapped_now = vmm_ops.mapped_pages()
        phys_now = vmm_ops.vmm_stats()["mapped_pages"]
        if rank == 0:
            assert sorted(mapped_now) == mgr.mapped_pages(), "VMM mapping must equal the bookkeeping's mapped set"
            if strict:
                assert phys_now - len(mapped_now) <= 5, \
                    f"strict mode: {phys_now} pages physically mapped, bookkeeping wants {len(mapped_now)}"
        if dist is not None:   # every rank must hold exactly rank 0's mapped page set
            h = torch.tensor([hash(tuple(mapped_now)) & 0x7fffffffffff, len(mapped_now)], dtype=torch.int64, device=dev)
            hs = [torch.zeros_like(h) for _ in range(world)]
            dist.all_gather(hs, h)
            assert all(torch.equal(x, hs[0]) for x in hs), "ranks disagree on the mapped page set"
            ph = torch.tensor([phys_now - len(mapped_now)], dtype=torch.int64, device=dev)
            dist.all_reduce(ph, op=dist.ReduceOp.MAX)
            leg["physical_overshoot_pages_after_fence_max_over_ranks"] = int(ph.item())
            if strict:
                assert int(ph.item()) <= 5, "strict mode: a rank holds pages the bookkeeping does not"
        return leg

    e2e_retain = e2e_leg(strict=False)
    e2e_strict = e2e_leg(strict=True)     # the default policy; last, so the state the later legs see is the default one
    e2e_value = e2e_strict["value"]
    vst = vmm_ops.vmm_stats()
    est = vmm_ops.engine_stats()
    tp_stats = tp.stats() if (tp is not None and args.tp_transport != "nccl") else None

    # ================= page map / unmap latency vs the raw driver loop ======================================
    map_unmap = {}
    if rank == 0:
        mgr.trim()
    tp_fence()             # rank 0's trim reaches the other ranks
    vmm_ops.settle()       # every rank hands retained pages back: the leg below creates its own handles
    barrier()
    if rank == 0:
        total_pages = int(mgr.stats().num_total_pages)
        page_bytes = vst["page_bytes"]
        free_b, _tot = torch.cuda.mem_get_info()
        for n in (1, 8, 64, 512):
            # n pages through the engine + n raw-control handles must fit the pool and the device
            if n > total_pages or 2 * n * page_bytes > 0.7 * free_b:
                map_unmap[f"n{n}"] = {"skipped": f"{n} pages of {page_bytes >> 20} MiB twice do not fit "
                                                 f"(pool {total_pages} pages, {free_b >> 30} GiB free)"}
                continue
            iters = 30 if n < 64 else (8 if n < 512 else 3)
            lib.kvb_vmm_pool_config(0, max(640, n + 16))
            _lib.check(lib.kvb_vmm_pool_prefill(0, n + 8), "pool_prefill")
            first = min(100, total_pages - n)
            ids = np.arange(first, first + n, dtype=np.int64)
            p = ids.ctypes.data_as(_lib.p_i64)
            mm, uu = [], []
            for i in range(iters + 3):
                t0 = time.perf_counter()
                rc_m = lib.kvb_map_pages(0, p, n)
                t1 = time.perf_counter()
                rc_u = lib.kvb_unmap_pages(0, p, n)
                t2 = time.perf_counter()
                assert rc_m == 1 and rc_u == 1, f"map/unmap of {n} pages failed ({rc_m}, {rc_u}): {_lib.last_error()}"
                if i >= 3:
                    mm.append((t1 - t0) * 1e6)
                    uu.append((t2 - t1) * 1e6)
            lib.kvb_vmm_pool_config(0, 16)      # the control loop creates its own handles
            a = np.zeros(iters)
            b = np.zeros(iters)
            _lib.check(lib.kvb_vmm_raw_control(0, n, iters, 1, a.ctypes.data_as(_lib.p_f64), b.ctypes.data_as(_lib.p_f64)),
                       "raw_control")
            map_unmap[f"n{n}"] = {"map_p50_us": pctl(mm, 0.5), "map_p99_us": pctl(mm, 0.99),
                                  "unmap_p50_us": pctl(uu, 0.5), "unmap_p99_us": pctl(uu, 0.99),
                                  "raw_map_p50_us": pctl(a.tolist(), 0.5), "raw_unmap_p50_us": pctl(b.tolist(), 0.5),
                                  "map_plus_unmap_vs_raw": (pctl(mm, 0.5) + pctl(uu, 0.5)) /
                                                           (pctl(a.tolist(), 0.5) + pctl(b.tolist(), 0.5))}
        map_unmap["page_bytes"] = page_bytes
        map_unmap["note"] = (f"one compound page = all {cfg['num_layers']} layers' K and V pages "
                             f"({page_bytes >> 20} MiB); raw = the same cuMemUnmap+cuMemMap+cuMemSetAccess calls on "
                             "pre-created handles, same process")
        lib.kvb_vmm_pool_config(0, 16)
    barrier()
    vmm_ops.settle()

    # ================= fused KV scatter / gather through the mapped range ======================================
    L, bs, cell = cfg["num_layers"], cfg["block_size"], cfg["cell_size"]
    Tn = SCATTER_T
    if world == 1:
        ids = mgr.alloc_np(Tn // bs)
        assert ids is not None
    else:                  # rank 0 allocates; pages reach the other ranks by broadcast, ids too
        ids_t = torch.zeros(Tn // bs, dtype=torch.int64, device=dev)
        if rank == 0:
            ids0 = mgr.alloc_np(Tn // bs)
            assert ids0 is not None
            ids_t.copy_(torch.from_numpy(ids0))
        tp_fence()
        dist.broadcast(ids_t, src=0)
        ids = ids_t.cpu().numpy()
    rng = np.random.default_rng(rank)
    slots = torch.from_numpy(rng.permutation((ids[:, None] * bs + np.arange(bs)[None, :]).reshape(-1))).to(dev)
    k = torch.empty((L, Tn, cell // 2), dtype=torch.int16, device=dev).random_()
    v = torch.empty_like(k).random_()
    lay = kv_ops.layout_vllm_contiguous(ranges[0][0], L, bs, cell, 2)
    alg_bytes = Tn * L * 4 * cell + 8 * Tn
    kv = {}
    for name, fn in (("scatter", kv_ops.kv_scatter), ("gather", kv_ops.kv_gather)):
        for _ in range(3):
            fn(lay, slots, k, v)
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for e0, e1 in evs:
            e0.record()
            fn(lay, slots, k, v)
            e1.record()
        barrier()
        ms = [e0.elapsed_time(e1) for e0, e1 in evs]
        t = torch.tensor([statistics.mean(ms)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kv[name] = {"ms": float(t.item()), "gbs_per_gpu": alg_bytes / float(t.item()) / 1e6,
                    "gbs_total": world * alg_bytes / float(t.item()) / 1e6}
    # end-to-end scatter from pinned host memory (smaller batch: PCIe bound): device staging buffers allocated once,
    # one untimed pass (the first touches lazily allocated pinned pages), median of 5
    Th = 2048
    kh = torch.empty((L, Th, cell // 2), dtype=torch.int16).random_().pin_memory()
    vh = torch.empty((L, Th, cell // 2), dtype=torch.int16).random_().pin_memory()
    sh = slots[:Th].cpu().pin_memory()
    kd, vd, sd = torch.empty_like(kh, device=dev), torch.empty_like(vh, device=dev), torch.empty_like(sh, device=dev)
    hs_ms = []
    for i in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kd.copy_(kh, non_blocking=True)
        vd.copy_(vh, non_blocking=True)
        sd.copy_(sh, non_blocking=True)
        kv_ops.kv_scatter(lay, sd, kd, vd)
        chk = int(sd[0].item())     # device->host read closes the step
        if i:
            hs_ms.append((time.perf_counter() - t0) * 1e3)
    dt = statistics.median(hs_ms) / 1e3
    h2d_b = 2 * kh.numel() * 2 + Th * 8
    kv["scatter_from_pinned_host"] = {"tokens": Th, "ms": dt * 1e3, "ms_min_max": [min(hs_ms), max(hs_ms)],
                                      "pcie_h2d_gbs": h2d_b / dt / 1e9, "gbs": (Th * L * 4 * cell) / dt / 1e9,
                                      "h2d_bytes": h2d_b, "d2h_bytes": 8,
                                      "note": "bound by the host->device copy (PCIe), not by the scatter kernel"}
    del kd, vd, sd
    # every rank is done with the pages before rank 0 gives them back: without this barrier a follower that is still in
    # its loop above scatters into pages its own VMM thread is switching to the zero page (unmapped for the ~250 us of a
    # transition) — an illegal address on the followers, seen once at N=4 (session 23)
    barrier()
    if rank == 0:
        mgr.free(ids)
    tp_fence()
    del k, v
    torch.cuda.empty_cache()

    # ================= the per-rank scatter at the TP shard shapes, on this GPU (ranks do not interact) ==========
    # Llama-3.3-70B: 80 layers, 8 KV heads sharded over TP ranks -> cell 2048 / 1024 / 512 / 256 B (BASELINE configs[4])
    shard = None
    init_leg = no_zero = None
    if rank == 0 and world == 1:
        shard = {}
        mgr.close()
        vmm_ops.shutdown_kvcached()
        for tpn, cell_s in ((1, 2048), (2, 1024), (4, 512), (8, 256)):
            Ls, Ts = 80, 65536
            blocks_s = Ts // bs + 64
            vmm_ops.init_kvcached(f"cuda:{local_rank}", PAGE, True)
            per_layer_s = -(-blocks_s * bs * cell_s * 2 // (2 * PAGE)) * 2 * PAGE
            (rng_s,) = vmm_ops.create_kv_tensors_raw(per_layer_s, Ls, 2)
            ms_ = KVCacheManager(blocks_s, bs, cell_s, Ls)
            ids_s = ms_.alloc_np(Ts // bs)
            slots_s = torch.from_numpy(np.random.default_rng(tpn).permutation(
                (ids_s[:, None] * bs + np.arange(bs)[None, :]).reshape(-1))).to(dev)
            ks = torch.empty((Ls, Ts, cell_s // 2), dtype=torch.int16, device=dev).random_()
            vs = torch.empty_like(ks).random_()
            lay_s = kv_ops.layout_vllm_contiguous(rng_s[0], Ls, bs, cell_s, 2)
            alg_s = Ts * Ls * 4 * cell_s + 8 * Ts
            row = {}
            for name, fn in (("scatter", kv_ops.kv_scatter), ("gather", kv_ops.kv_gather)):
                for _ in range(3):
                    fn(lay_s, slots_s, ks, vs)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn(lay_s, slots_s, ks, vs)
                e1.record()
                torch.cuda.synchronize()
                row[name + "_gbs_per_gpu"] = alg_s / (e0.elapsed_time(e1) / 10) / 1e6
            row["cell_bytes"] = cell_s
            row["frac_of_hbm_peak_scatter"] = None
            shard[f"tp{tpn}"] = row
            del ks, vs
            ms_.close()
            vmm_ops.shutdown_kvcached()
            torch.cuda.empty_cache()
        # e2e once more WITHOUT the zero page (opt-in kvb_vmm_zero_init_mode(2)): half the sleeping driver calls per page
        # transition, unmapped pages fault instead of reading zeros
        try:
            vmm_ops.zero_init_mode(2)
            vmm_ops.init_kvcached(f"cuda:{local_rank}", PAGE, True)
            vmm_ops.create_kv_tensors_raw(per_layer, cfg["num_layers"], 2)
            mgr_nz = R.manager_from_config(cfg, device_index=local_rank)
            tnz = []
            vt0 = vmm_ops.vmm_timing()
            Knz = min(K, 4)
            for it in range(1 + Knz):
                t0 = time.perf_counter()
                R.replay_host_steps(mgr_nz, plan, keep_results=False)
                vmm_ops.fence()
                tnz.append(time.perf_counter() - t0)
                if it == 0:
                    vt0 = vmm_ops.vmm_timing()
            vt1 = vmm_ops.vmm_timing()
            no_zero = {"value": n_calls * Knz / sum(tnz[1:]), "unit": UNIT, "ms_per_step": 1e3 * sum(tnz[1:]) / Knz,
                       "steps": Knz,
                       "driver_ms_per_replay": sum(vt1[k] - vt0[k] for k in ("unmap_ms", "map_ms", "setaccess_ms", "create_ms",
                                                                             "release_ms")) / Knz,
                       "policy": "as e2e, but no zero page behind unmapped pages (they fault if touched)"}
            mgr_nz.close()
        except Exception as e:
            no_zero = {"error": str(e)[:300]}
        finally:
            vmm_ops.shutdown_kvcached()
            vmm_ops.zero_init_mode(-1)
        init_leg = init_and_layout_leg(local_rank, cfg, (cpu_baseline or {}).get("create_kv_tensors_s"))
        # the rest of the bench needs the original group again
        vmm_ops.init_kvcached(f"cuda:{local_rank}", PAGE, True)
        vmm_ops.create_kv_tensors_raw(per_layer, cfg["num_layers"], 2)
        mgr = R.manager_from_config(cfg, device_index=local_rank)

    # ================= bitmap sweep on a stress size (2^30 blocks = 128 MiB bitmap streams HBM) ===
    sweep = None
    micro = per_call = mla_pin = None
    if rank == 0:
        o4 = np.zeros(4)
        o5 = np.zeros(4)
        rc = lib.kvb_audit_stress(local_rank, 1 << 30, 64, 20, o4.ctypes.data_as(_lib.p_f64))
        rc2 = lib.kvb_audit_stress(local_rank, 1 << 31, 64, 20, o5.ctypes.data_as(_lib.p_f64))
        if rc == 0:
            sweep = {"kernel": "kvb_audit_kernel (uint4 bitmap sweep + popc, 2^30 blocks, 64 blocks/page)",
                     "ms": float(o4[0]), "algorithmic_bytes_per_launch": int(o4[1]),
                     "gbs": float(o4[1]) / float(o4[0]) / 1e6, "used_blocks_counted": int(o4[2]),
                     "note": "the pool's own bitmap is 11 KB (L2-resident, launch-bound): audit of the live pool "
                             f"takes {mgr.audit()[4] * 1e3:.1f} us"}
            if rc2 == 0:
                sweep["at_2^31_blocks"] = {"ms": float(o5[0]), "algorithmic_bytes_per_launch": int(o5[1]),
                                           "gbs": float(o5[1]) / float(o5[0]) / 1e6}
            sweep["frac_of_hbm_peak"] = sweep["gbs"] / float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) \
                if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else sweep["gbs"] / 6650.0
        else:
            sweep = {"error": _lib.last_error()}
        # ---- T-micro (benchmarks/bench_alloc/bench_alloc.py:57-62): alloc(k); free(h) pairs through the single-request
        # API the unmodified integration patches call, page-churn (pool empty between pairs) and no-churn (one pinned
        # block keeps the page), the unmodified reference timed beside it in the same run on the same sections.
        micro = micro_alloc_free_leg(cfg, local_rank, with_reference=not args.no_cpu_baseline)
        # ---- T-decode-4096 through the PER-CALL API (one alloc()/free() per request, what unmodified patches.py does),
        # real VMM, default policy; kvb_vmm_fence inside the timed region
        per_call = per_call_e2e_leg(mgr, tr, n_calls, vmm_ops)
        # ---- T-mla-pin (BASELINE configs[3]): DeepSeek-V3 MLA latent KV (61 layers, cell 1152 B, one buffer), prefix-cache
        # pinning with 20 shared prefixes x Zipf(1.1), cap 16 000 tokens, page size 64 and 1
        mla_pin = mla_pin_leg(local_rank, with_reference=not args.no_cpu_baseline)
    clk = clocks.stop()

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    if shard:
        for row in shard.values():
            row["frac_of_hbm_peak_scatter"] = row["scatter_gbs_per_gpu"] / peak
    roofline = {"kernel": f"kvb_kv_copy_kernel<scatter> (fused {L}-layer paged KV scatter, cell {cell} B, T={Tn} tokens)",
                "bound": "hbm", "achieved": kv["scatter"]["gbs_per_gpu"], "peak": peak, "unit": "GB/s",
                "frac": kv["scatter"]["gbs_per_gpu"] / peak,
                "traffic": SCATTER_TRAFFIC_NCU if (world == 1 and Tn == SCATTER_T) else None,
                "traffic_source": "ncu --set full dram bytes of one launch at this shape (profiles/r2_kv_copy_cell2048.csv)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "step_kernel": {"us_per_scheduler_step": 1e3 * res_ms_total / K / n_sched_steps,
                                "note": "kvb_decode_kernel / kvb_step_kernel: one latency-bound CTA per scheduler step over "
                                        "L2-resident state (11 KB bitmap, 366 KB rings at this size), launched with "
                                        "programmatic stream serialization inside a CUDA graph: no HBM roofline applies"}}
    total_launches = lib.kvb_total_launches() - launches0

    two_model = None
    if rank == 0 and world == 1 and not args.no_two_model:
        # the two engine processes need the device: this process lets go of its own mappings first
        mgr.trim()
        vmm_ops.settle()
        lib.kvb_vmm_pool_config(0, 0)          # ... and of the unmapped handles its pool keeps
        del flush_buf
        torch.cuda.empty_cache()
        two_model = two_model_leg(local_rank)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": res_ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int64", "data": "synthetic", "config": workload_config(world), "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": e2e_strict["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": e2e_strict["d2h_bytes_per_step"],
                        "ms_per_step": e2e_strict["ms_per_step"], "scheduler_steps": len(plan.steps),
                        "policy": "library default (retention off): every page event of the bookkeeping is applied to the "
                                  "device inside the timed region; map-ahead 5 pages; unmapped handles pooled, <= 10 % "
                                  "of the device, released after 2 s without a map",
                        "page_events_per_replay": page_events_per_replay,
                        "page_transitions_per_replay": e2e_strict["page_transitions_per_replay"],
                        "driver_ms_per_replay": e2e_strict["driver_ms_per_replay"],
                        "physical_overshoot_pages_after_fence_max": e2e_strict["physical_overshoot_pages_after_fence_max"],
                        "strict": e2e_strict,
                        "with_retention": dict(e2e_retain, policy=(
                            "opt-in KVB200_VMM_RETAIN_MS=2000: pages the bookkeeping unmapped keep their handle until the "
                            "allocator is quiet for 2 s, never more than 10 % of the device (KVB200_RETAIN_MAX_FRACTION), "
                            "leaving free the utilisation headroom + 2 % of the device, or 20 % while anybody else's "
                            "footprint grows")),
                        "vmm": {"handles_created": vst["handles_created"], "map_calls": vst["map_calls"],
                                "map_ahead_hits": est["prefetch_hits"], "retention_hits": est["retention_hits"],
                                "page_bytes": vst["page_bytes"]},
                        "no_zero_page": no_zero,
                        "tp": tp_stats, "tp_transport": (args.tp_transport if world > 1 else None),
                        "api": "KVCacheManager.run_ops (kvb_mgr_step) per scheduler step, attach_vmm=True: maps are in "
                               "place (on every TP rank) when a step returns, unmaps drain on the VMM thread, "
                               "kvb_vmm_fence() closes the timed region"},
                "gpu_launches": int(total_launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "map_unmap": map_unmap, "kv_scatter": kv, "kv_scatter_tp_shards_on_one_gpu": shard,
                "bitmap_sweep": sweep, "micro_alloc_free": micro, "e2e_per_call_api": per_call, "mla_pin": mla_pin,
                "two_model": two_model, "init_and_layout": init_leg,
                "ops_per_replay": n_calls, "scheduler_steps": int(n_sched_steps),
                "resident_launches_per_replay": int(resident_launches)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
    if tp is not None and hasattr(tp, "close"):
        tp.close()
    mgr.close()
    vmm_ops.shutdown_kvcached()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-two-model", action="store_true")
    ap.add_argument("--two-model-worker", nargs=3, metavar=("ROLE", "DIR", "LOCAL_RANK"), help=argparse.SUPPRESS)
    ap.add_argument("--ref-replays", type=int, default=3, help="timed replays of the cpu_baseline leg (N=1)")
    ap.add_argument("--tp-transport", default="mailbox", choices=["mailbox", "nccl"],
                    help="N>1: how rank 0's page lists reach the other ranks (mailbox = kernel push into a shared ring; "
                         "nccl = one broadcast per scheduler step, the checker)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.two_model_worker:
        role, d, lr = args.two_model_worker
        two_model_worker(role, d, int(lr))
        return
    try:
        if args.impl == "reference":
            reference_main(args, rank, world)
        else:
            ours_main(args, rank, world, local_rank)
    except BaseException as e:
        if isinstance(e, SystemExit) and not e.code:
            raise
        # the launcher keeps only its own wrapper's tail: say what failed, on which rank, before going down
        import traceback
        sys.stderr.write(f"[bench] rank {rank} of {world} failed:\n{traceback.format_exc()}\n")
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
