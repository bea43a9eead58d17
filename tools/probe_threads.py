#!/usr/bin/env python
"""Resident replay of the bench trace: device ms per replay for the loaded library build (KVB200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcached_b200 import _lib, replay as R, trace as T
torch.cuda.init()
lib = _lib.load()
cfg = T.make_config("llama8b", 91520)
ct = R.compile_trace(T.decode_trace(cfg, 4096, seed=0))
m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
R.upload(m, ct)
for spl in (1, 1, 100000, 100000):
    lib.kvb_mgr_trace_reset_cursor(m._h)
    ms, nl = R.run_resident(m, ct, steps_per_launch=spl)
    print(os.environ.get("KVB200_LIB", "default"), "spl", spl, "ms", round(ms, 2), "launches", nl, flush=True)
import numpy as np
prof = np.zeros(32, dtype=np.int64)
lib.kvb_mgr_profile(m._h, prof.ctypes.data_as(_lib.p_i64), 0)
if prof.any():
    names = ["pre-alloc/idle", "alloc fast", "alloc general", "free: stage ids", "F0 classify", "F1+F2 distinct pages", "F3 dict order + ranks",
             "F4 page updates", "F5 free_pages", "loop->publish", "publish", "F3 warp0 own work", "w32: phase head (next_bkt, mod)", "w32: match_any", "w32: suffix scan", "w32: permute"]
    tot = prof[:16].sum()
    for i, n in enumerate(names):
        if prof[16 + i]:
            print(f"  {n:24s} {prof[i]/tot*100:5.1f}%  {prof[i]/prof[16+i]:8.0f} cycles/mark  marks {prof[16+i]}")
assert m.stats().num_inuse_pages == 0
m.close()
