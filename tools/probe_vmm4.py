#!/usr/bin/env python
"""Is a driver VMM call slower right after kernel activity?  map/unmap n pages with and without a
kernel launch + sync in between; per-call attribution from kvb_vmm_timing."""
import os
os.environ.setdefault("KVB200_LIB", "test")   # the driver probes live in libkvb200_test.so
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcached_b200 import _lib, vmm_ops
lib = _lib.load()
torch.cuda.init()
PAGE = 2 << 20
vmm_ops.init_kvcached("cuda:0", PAGE, True)
vmm_ops.create_kv_tensors_raw(64 * PAGE * 2, 32, 2)   # 64 compound pages of 128 MiB
lib.kvb_vmm_pool_config(0, 64); lib.kvb_vmm_pool_prefill(0, 40)
x = torch.zeros(1024, device="cuda")
def cyc(n, kernel, iters=30, stream_kernel=False):
    ids = np.arange(n, dtype=np.int64); p = ids.ctypes.data_as(_lib.p_i64)
    a = vmm_ops.vmm_timing(); mm = []; uu = []
    for i in range(iters + 3):
        if kernel: x.add_(1); torch.cuda.synchronize()
        t0 = time.perf_counter(); lib.kvb_map_pages(0, p, n); t1 = time.perf_counter()
        if kernel: x.add_(1); torch.cuda.synchronize()
        t2 = time.perf_counter(); lib.kvb_unmap_pages(0, p, n); t3 = time.perf_counter()
        if i >= 3: mm.append((t1 - t0) * 1e6); uu.append((t3 - t2) * 1e6)
    b = vmm_ops.vmm_timing()
    per = {k: round((b[k + "_ms"] - a[k + "_ms"]) * 1e3 / max(1, b[k + "_calls"] - a[k + "_calls"]), 1) for k in ("unmap", "map", "setaccess")}
    print(f"n={n} kernel_between={kernel}: map p50 {np.median(mm):.0f} us, unmap p50 {np.median(uu):.0f} us, per-call us {per}", flush=True)
for n in (1, 8):
    cyc(n, False); cyc(n, True); cyc(n, False)
vmm_ops.shutdown_kvcached()
