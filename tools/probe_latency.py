#!/usr/bin/env python
"""Where does a single alloc(k) / free(ids) call spend its time?  ctypes overhead, the C call alone, the Python method."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvcached_b200 import _lib
from kvcached_b200 import replay as R, trace as T
lib = _lib.load()
out = {}
def bench(f, n=5000, warm=200):
    for _ in range(warm): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
out["ctypes_noop_us"] = bench(lambda: lib.kvb_abi_version(), 20000)
cfg = T.make_config("llama8b", 91520)
m = R.manager_from_config(cfg, attach_vmm=False)
m.set_phys_free_pages(1 << 40)
pin = m.alloc(1)
buf = np.empty(4096, dtype=np.int64)
for k in (1, 16, 256):
    ids = m.alloc_np(k)
    m.free(ids)
    def pair_c():
        n = lib.kvb_mgr_alloc(m._h, k, buf.ctypes.data, buf.size)
        lib.kvb_mgr_free(m._h, buf.ctypes.data, n)
    def alloc_only():
        n = lib.kvb_mgr_alloc(m._h, k, buf.ctypes.data, buf.size)
    out[f"k{k}_pair_c_us"] = bench(pair_c, 3000)
    def pair_py():
        m.free(m.alloc(k))
    out[f"k{k}_pair_py_list_us"] = bench(pair_py, 3000)
    def pair_np():
        m.free(m.alloc_np(k))
    out[f"k{k}_pair_py_np_us"] = bench(pair_np, 3000)
# split: alloc alone / free alone (C level), k = 16
got = []
t0 = time.perf_counter()
for i in range(2000):
    n = lib.kvb_mgr_alloc(m._h, 16, buf.ctypes.data + 0, 16)
    got.append(buf[:16].copy())
t1 = time.perf_counter()
for g in got:
    lib.kvb_mgr_free(m._h, g.ctypes.data, 16)
t2a = time.perf_counter()
lib.kvb_mgr_sync(m._h)
t2 = time.perf_counter()
out["k16_free_host_side_us"] = (t2a - t1) / 2000 * 1e6
out["k16_alloc_only_us"] = (t1 - t0) / 2000 * 1e6
out["k16_free_only_us"] = (t2 - t1) / 2000 * 1e6
out["available_size_us"] = bench(lambda: m.available_size(), 5000)
m.close()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_latency.json", "w"), indent=1)
