#!/usr/bin/env python
"""Cycles per phase of the step kernel (-DKVB_PROFILE build: KVB200_LIB=kvcached_b200/libkvb200_prof.so) for
(a) single alloc(k) / free(ids) launches, (b) the resident replay of the bench trace."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("KVB200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kvcached_b200", "libkvb200_prof.so"))
import numpy as np
import torch
from kvcached_b200 import _lib, replay as R, trace as T
torch.cuda.init()
lib = _lib.load()
NAMES = ["pre-alloc/idle", "alloc run, warp path", "alloc general", "free: stage ids", "F0 classify", "F1+F2 distinct pages", "F3 dict order + ranks",
         "F4 page updates", "F5 free_pages", "loop->publish", "publish", "F3 warp0 own work", "alloc run, block-wide (w32: phase head in free_core)", "free: request loop up to the free case", "F0a: early loads issued, table cleared", "F0b: hash insert"]
def show(m, title, launches):
    prof = np.zeros(32, dtype=np.int64)
    lib.kvb_mgr_profile(m._h, prof.ctypes.data_as(_lib.p_i64), 1)
    tot = prof[:11].sum() + prof[12:16].sum()
    print(f"== {title}: {tot/launches:.0f} cycles per launch in marked phases ({tot/launches/1.9:.0f} ns at 1.9 GHz)")
    for i, n in enumerate(NAMES):
        if prof[16 + i]:
            print(f"  {n:26s} {prof[i]/max(tot,1)*100:5.1f}%  {prof[i]/prof[16+i]:8.0f} cycles/mark  marks {prof[16+i]}")
cfg = T.make_config("llama8b", 91520)
m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
pin = m.alloc(1)
for k in (1, 16, 256):
    for _ in range(100): m.free(m.alloc_np(k))
    m.sync()
    prof = np.zeros(32, dtype=np.int64); lib.kvb_mgr_profile(m._h, prof.ctypes.data_as(_lib.p_i64), 1)
    t0 = time.perf_counter()
    for _ in range(1000): m.free(m.alloc_np(k))
    m.sync()
    dt = (time.perf_counter() - t0) / 1000 * 1e6
    show(m, f"alloc({k}) + free pairs, {dt:.1f} us per pair", 2000)
m.close()
m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
ct = R.compile_trace(T.decode_trace(cfg, 4096, seed=0))
R.upload(m, ct)
for spl in (1, 1):
    lib.kvb_mgr_trace_reset_cursor(m._h)
    prof = np.zeros(32, dtype=np.int64); lib.kvb_mgr_profile(m._h, prof.ctypes.data_as(_lib.p_i64), 1)
    ms, nl = R.run_resident(m, ct, steps_per_launch=spl)
    show(m, f"resident replay, {ms:.1f} ms, {nl} launches", nl)
m.close()
