#!/usr/bin/env python
"""One e2e replay of the decode trace with driver-call time attribution."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcached_b200 import _lib, vmm_ops, replay as R, trace as T
torch.cuda.init()
PAGE = 2 << 20
cfg = T.make_config("llama8b", 91520)
tr = T.decode_trace(cfg, 4096, seed=0)
ct = R.compile_trace(tr)
plan = R.HostStepPlan(ct)
pool = int(sys.argv[1]) if len(sys.argv) > 1 else 16
vmm_ops.init_kvcached("cuda:0", PAGE, True)
t0 = time.perf_counter()
vmm_ops.create_kv_tensors_raw(cfg["num_blocks"] * 16 * 2048 * 2, 32, 2)
print("create_kv_tensors s", time.perf_counter() - t0)
_lib.load().kvb_vmm_pool_config(0, pool)
m = R.manager_from_config(cfg)
for it in range(3):
    a = vmm_ops.vmm_timing()
    t0 = time.perf_counter()
    R.replay_host_steps(m, plan, keep_results=False)
    dt = time.perf_counter() - t0
    b = vmm_ops.vmm_timing()
    d = {k: round(b[k] - a[k], 1) for k in b}
    print(f"replay {it}: {dt:.2f} s  ops/s {ct.n_calls/dt:.0f}  vmm {json.dumps(d)}", flush=True)
print(vmm_ops.vmm_stats())
m.close(); vmm_ops.shutdown_kvcached()
