#!/usr/bin/env python
"""Is the driver time of an e2e replay the driver's, or ours?  Alternates, on one box within one minute, (a) e2e replays of
the bench trace through KVCacheManager.run_ops with the VMM thread's per-call-kind driver time and (b) tools/vmm_floor
(the same page events against a private reservation in a process that does nothing else)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from kvcached_b200 import vmm_ops, replay as R, trace as T
torch.cuda.init()
PAGE = 2 << 20
cfg = T.make_config("llama8b", 91520)
ct = R.compile_trace(T.decode_trace(cfg, 4096, seed=0))
plan = R.HostStepPlan(ct)
vmm_ops.init_kvcached("cuda:0", PAGE, True)
vmm_ops.create_kv_tensors_raw(cfg["num_blocks"] * 16 * 2048 * 2, 32, 2)
m = R.manager_from_config(cfg)
KEYS = ("unmap_ms", "map_ms", "setaccess_ms", "create_ms", "release_ms")
def e2e(tag):
    a = vmm_ops.vmm_timing()
    t0 = time.perf_counter()
    R.replay_host_steps(m, plan, keep_results=False)
    t1 = time.perf_counter()
    vmm_ops.fence()
    t2 = time.perf_counter()
    b = vmm_ops.vmm_timing()
    calls = b["unmap_calls"] - a["unmap_calls"]
    d = {k: b[k] - a[k] for k in KEYS}
    extra = {k: b[k] - a[k] for k in b if k.endswith("_calls")}
    print(f"e2e {tag}: scheduler loop {t1-t0:.3f} s + fence {t2-t1:.3f} s; {calls} transitions; driver ms " +
          " ".join(f"{k[:-3]}={v:.0f}" for k, v in d.items()) + f" total={sum(d.values()):.0f} ({sum(d.values())/max(calls,1)*1e3:.0f} us each); calls {extra}", flush=True)
def floor(variant):
    r = subprocess.run([os.path.join(ROOT, "tools", "vmm_floor"), os.path.join(ROOT, "tools", "events_decode4096.txt"), variant, "1", "128"],
                       capture_output=True, text=True, timeout=300)
    for ln in r.stdout.splitlines():
        if "replay" in ln or "Unmap" in ln or "SetAccess" in ln or "Create" in ln:
            print("   floor", variant, ln.strip(), flush=True)
from kvcached_b200 import _lib
lib = _lib.load()
e2e("warm-up")
if len(sys.argv) > 1 and sys.argv[1] == "pool":
    # does the handle churn (cuMemRelease + cuMemCreate beyond the 10 % pool) slow the OTHER calls down?
    for it in range(3):
        lib.kvb_vmm_pool_config(0, 142)
        e2e(f"pool cap 142 pages (default 10 %) #{it}")
        lib.kvb_vmm_pool_config(0, 1430)
        e2e(f"pool cap 1430 pages, filling #{it}")
        e2e(f"pool cap 1430 pages, no churn #{it}")
        floor("zero")
else:
    for it in range(3):
        e2e(f"strict #{it}")
        floor("zero")
        floor("zero_ondemand")
m.close(); vmm_ops.shutdown_kvcached()
