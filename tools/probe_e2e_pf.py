#!/usr/bin/env python
"""e2e replays of the bench trace alternating the map-ahead depth (same process, same box phase): seconds per replay
and the VMM thread's driver time."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kvcached_b200 import _lib, vmm_ops, replay as R, trace as T
torch.cuda.init()
PAGE = 2 << 20
cfg = T.make_config("llama8b", 91520)
ct = R.compile_trace(T.decode_trace(cfg, 4096, seed=0))
plan = R.HostStepPlan(ct)
vmm_ops.init_kvcached("cuda:0", PAGE, True)
vmm_ops.create_kv_tensors_raw(cfg["num_blocks"] * 16 * 2048 * 2, 32, 2)
m = R.manager_from_config(cfg)
depths = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,5").split(",")]   # negative: retention off, depth = -x-1
res = {d: [] for d in depths}
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    for d in depths:
        vmm_ops.set_prefetch(d if d >= 0 else -d - 1)
        vmm_ops.set_retention(2000 if d >= 0 else 0)
        a = vmm_ops.vmm_timing(); e0 = vmm_ops.engine_stats()
        t0 = time.perf_counter()
        R.replay_host_steps(m, plan, keep_results=False)
        t1 = time.perf_counter()
        vmm_ops.fence()
        t2 = time.perf_counter()
        b = vmm_ops.vmm_timing(); e1 = vmm_ops.engine_stats()
        drv = sum(b[k] - a[k] for k in ("unmap_ms", "map_ms", "setaccess_ms", "create_ms"))
        calls = b["unmap_calls"] - a["unmap_calls"]
        res[d].append(t2 - t0)
        print(f"it {it} depth {d}: replay {t1-t0:.3f} s + fence {t2-t1:.3f} s = {t2-t0:.3f} s; driver {drv/1e3:.3f} s in {calls} transitions "
              f"({drv/max(calls,1)*1e3:.0f} us each); ahead hits {e1['prefetch_hits']-e0['prefetch_hits']} retention hits {e1['retention_hits']-e0['retention_hits']} "
              f"retained {e1['retained']} free GiB {torch.cuda.mem_get_info()[0]/2**30:.1f}", flush=True)
for d in depths:
    print(f"depth {d}: median {np.median(res[d]):.3f} s  min {min(res[d]):.3f} s  -> {ct.n_calls/np.median(res[d]):.0f} ops/s (best {ct.n_calls/min(res[d]):.0f})")
m.close(); vmm_ops.shutdown_kvcached()
