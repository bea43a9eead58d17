#!/usr/bin/env python
"""Quick B200 probe: raw VMM driver costs, per-call latency of the manager, scatter bandwidth."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvcached_b200 import _lib, kv_ops, vmm_ops  # noqa: E402
from kvcached_b200 import replay as R  # noqa: E402
from kvcached_b200 import trace as T  # noqa: E402
from kvcached_b200.kv_cache_manager import KVCacheManager  # noqa: E402

out = {}
lib = _lib.load()
PAGE = 2 << 20
import torch  # noqa: E402

torch.cuda.init()

def raw(contig, L, n, iters):
    vmm_ops.init_kvcached("cuda:0", PAGE, contig)
    per_layer = 64 * PAGE * 2
    vmm_ops.create_kv_tensors_raw(per_layer, L, 2)
    a = np.zeros(iters); b = np.zeros(iters)
    rc = lib.kvb_vmm_raw_control(0, n, iters, 1, a.ctypes.data_as(_lib.p_f64), b.ctypes.data_as(_lib.p_f64))
    assert rc == 0, _lib.last_error()
    # our own path on the same pages
    ids = np.arange(n, dtype=np.int64)
    lib.kvb_vmm_pool_config(0, 4 * n)
    lib.kvb_vmm_pool_prefill(0, n)
    mm, uu = [], []
    for _ in range(iters):
        t0 = time.perf_counter(); lib.kvb_map_pages(0, ids.ctypes.data_as(_lib.p_i64), n)
        t1 = time.perf_counter(); lib.kvb_unmap_pages(0, ids.ctypes.data_as(_lib.p_i64), n)
        t2 = time.perf_counter()
        mm.append((t1 - t0) * 1e6); uu.append((t2 - t1) * 1e6)
    # cold: create handles inside map (pool empty)
    lib.kvb_vmm_pool_config(0, 0)
    t0 = time.perf_counter(); lib.kvb_map_pages(0, ids.ctypes.data_as(_lib.p_i64), n); t1 = time.perf_counter()
    lib.kvb_unmap_pages(0, ids.ctypes.data_as(_lib.p_i64), n); t2 = time.perf_counter()
    vmm_ops.shutdown_kvcached()
    return dict(n=n, raw_map_us_p50=float(np.median(a)), raw_unmap_us_p50=float(np.median(b)),
                ours_map_us_p50=float(np.median(mm)), ours_unmap_us_p50=float(np.median(uu)),
                cold_map_us=(t1 - t0) * 1e6, cold_unmap_release_us=(t2 - t1) * 1e6)

out["vmm_compound_128MiB_n1"] = raw(True, 32, 1, 30)
out["vmm_compound_128MiB_n8"] = raw(True, 32, 8, 10)
out["vmm_2MiB_perlayer_L4_n8"] = raw(False, 4, 8, 10)

# cudaMemGetInfo cost
t0 = time.perf_counter()
for _ in range(2000):
    torch.cuda.mem_get_info()
out["mem_get_info_us"] = (time.perf_counter() - t0) / 2000 * 1e6

# manager latencies (bookkeeping only)
cfg = T.make_config("llama8b", 91520)
m = R.manager_from_config(cfg, attach_vmm=False)
pin = m.alloc(1)
for k in (1, 16, 256):
    for _ in range(50):
        m.free(m.alloc_np(k))
    t0 = time.perf_counter()
    for _ in range(2000):
        m.free(m.alloc_np(k))
    out[f"alloc{k}_free_pair_us"] = (time.perf_counter() - t0) / 2000 * 1e6
t0 = time.perf_counter()
for _ in range(5000):
    m.available_size()
out["available_size_us"] = (time.perf_counter() - t0) / 5000 * 1e6
# live phys query
m.set_phys_free_pages(-1)
t0 = time.perf_counter()
for _ in range(2000):
    m.available_size()
out["available_size_live_us"] = (time.perf_counter() - t0) / 2000 * 1e6
m.set_phys_free_pages(1 << 40)
needs = [1] * 20 + [40]
held = []
t0 = time.perf_counter()
for _ in range(2000):
    got = m.step(held[:2], needs)
    held = [g for g in got]
out["step_21allocs_2frees_us"] = (time.perf_counter() - t0) / 2000 * 1e6
m.close()

# decode trace, resident
tr = T.decode_trace(cfg, 4096, seed=0)
ct = R.compile_trace(tr)
m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
R.upload(m, ct)
for spl in (1, 8, 100000):
    lib.kvb_mgr_trace_reset_cursor(m._h)
    t0 = time.perf_counter(); ms, nl = R.run_resident(m, ct, steps_per_launch=spl); wall = time.perf_counter() - t0
    out[f"decode_resident_spl{spl}"] = dict(device_ms=ms, wall_ms=wall * 1e3, launches=nl, ops=ct.n_calls,
                                            mops_per_s=ct.n_calls / (ms * 1e-3) / 1e6)
m.close()
m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
plan = R.HostStepPlan(ct)
t0 = time.perf_counter(); R.replay_host_steps(m, plan, keep_results=False); wall = time.perf_counter() - t0
out["decode_host_steps"] = dict(wall_ms=wall * 1e3, ops=ct.n_calls, kops_per_s=ct.n_calls / wall / 1e3,
                                steps=len(plan.steps))
m.close()

# scatter bandwidth, llama8b layout, all 32 layers fused
vmm_ops.init_kvcached("cuda:0", PAGE, True)
L, bs, cell = 32, 16, 2048
blocks = 8192
per_layer = blocks * bs * cell * 2
(rng_,) = vmm_ops.create_kv_tensors_raw(per_layer, L, 2)
mg = KVCacheManager(blocks, bs, cell, L)
for Tn in (4096, 65536):
    ids = np.asarray(mg.alloc(Tn // bs), dtype=np.int64)
    slots = torch.from_numpy(np.random.default_rng(0).permutation((ids[:, None] * bs + np.arange(bs)).reshape(-1))).cuda()
    k = torch.empty((L, Tn, cell // 2), dtype=torch.int16, device="cuda").random_()
    v = torch.empty_like(k).random_()
    lay = kv_ops.layout_vllm_contiguous(rng_[0], L, bs, cell, 2)
    for name, fn in (("scatter", kv_ops.kv_scatter), ("gather", kv_ops.kv_gather)):
        for _ in range(3):
            fn(lay, slots, k, v)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(lay, slots, k, v)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byts = Tn * L * 4 * cell + 8 * Tn
        out[f"{name}_T{Tn}"] = dict(ms=ms, gbs=byts / ms / 1e6)
    mg.free(ids)
    del k, v
# torch copy for reference
a = torch.empty(1 << 30, dtype=torch.int16, device="cuda"); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): b.copy_(a)
e1.record(); torch.cuda.synchronize()
out["torch_copy_gbs"] = 2 * a.numel() * 2 / (e0.elapsed_time(e1) / 10) / 1e6
mg.close(); vmm_ops.shutdown_kvcached()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
