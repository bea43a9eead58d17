#!/usr/bin/env python
"""Small deterministic workload for ncu: one resident replay of a 512-request decode trace in a single
launch (step kernel) and one fused 32-layer scatter + gather of 16384 tokens."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcached_b200 import _lib, kv_ops, vmm_ops
from kvcached_b200 import replay as R, trace as T
from kvcached_b200.kv_cache_manager import KVCacheManager
which = sys.argv[1] if len(sys.argv) > 1 else "both"
torch.cuda.init()
if which in ("step", "both"):
    nreq = int(os.environ.get("NCU_REQ", "512"))
    cfg = T.make_config("llama8b", 91520 if nreq >= 4096 else 16384)
    tr = T.decode_trace(cfg, nreq, seed=0)
    ct = R.compile_trace(tr)
    m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
    R.upload(m, ct)
    spl = int(os.environ.get("NCU_SPL", str(1 << 20)))     # steps per launch: 1 = one launch per scheduler step
    for _ in range(2):
        _lib.load().kvb_mgr_trace_reset_cursor(m._h)
        ms, nl = R.run_resident(m, ct, steps_per_launch=spl)
    print("step kernel: ops", ct.n_calls, "steps", ct.step_offsets.size - 1, "launches", nl, "ms", ms)
    m.close()
if which == "single":
    # the single-request path: alloc(16) (flagged record) / free (publishes nothing) pairs
    cfg = T.make_config("llama8b", 91520)
    m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
    pin = m.alloc(1)
    for _ in range(300):
        m.free(m.alloc(16))
    m.sync()
    print("single-request pairs done")
    m.close()
if which in ("kv", "both"):
    PAGE = 2 << 20
    vmm_ops.init_kvcached("cuda:0", PAGE, True)
    L, bs, cell = int(os.environ.get("NCU_L", "32")), 16, int(os.environ.get("NCU_CELL", "2048"))
    blocks = max(4096, int(os.environ.get("NCU_T", "16384")) // 16 + 64)
    (rng_,) = vmm_ops.create_kv_tensors_raw(blocks * bs * cell * 2, L, 2)
    mg = KVCacheManager(blocks, bs, cell, L)
    Tn = int(os.environ.get("NCU_T", "16384"))
    ids = np.asarray(mg.alloc(Tn // bs), dtype=np.int64)
    slots = torch.from_numpy(np.random.default_rng(0).permutation((ids[:, None] * bs + np.arange(bs)).reshape(-1))).cuda()
    k = torch.empty((L, Tn, cell // 2), dtype=torch.int16, device="cuda").random_()
    v = torch.empty_like(k).random_()
    lay = kv_ops.layout_vllm_contiguous(rng_[0], L, bs, cell, 2)
    for _ in range(3):
        kv_ops.kv_scatter(lay, slots, k, v)
        kv_ops.kv_gather(lay, slots, k, v)
    torch.cuda.synchronize()
    print("kv done")
    mg.close(); vmm_ops.shutdown_kvcached()
