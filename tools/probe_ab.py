#!/usr/bin/env python
"""Same-box A/B of kernel tuning switches (KVB200_TUNE bits, kvb_internal.h LaunchArgs.tune) on the resident replay of
the bench trace: one manager per variant, alternating, CUDA-event times of the product build.
    python tools/probe_ab.py [tune values ...]      default: 0 2 4 6 1"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcached_b200 import _lib, replay as R, trace as T
torch.cuda.init()
lib = _lib.load()
NAMES = {0: "default (PDL, early ref gather, kvb_decode_kernel)", 1: "one-warp dict-order replay",
         2: "no programmatic dependent launch", 4: "no early ref gather", 6: "no PDL, no early gather",
         8: "general kernel for every step (no kvb_decode_kernel)"}
variants = [int(x) for x in sys.argv[1:]] or [0, 2, 4, 6, 1]
cfg = T.make_config("llama8b", 91520)
ct = R.compile_trace(T.decode_trace(cfg, 4096, seed=0))
def mk(tune):
    os.environ["KVB200_TUNE"] = str(tune)
    m = R.manager_from_config(cfg, attach_vmm=False); m.set_phys_free_pages(1 << 40)
    R.upload(m, ct)
    return m
ms = {t: mk(t) for t in variants}
res = {t: [] for t in variants}
nl = 0
for it in range(8):
    for tune in variants:
        m = ms[tune]
        lib.kvb_mgr_trace_reset_cursor(m._h)
        t, nl = R.run_resident(m, ct, steps_per_launch=1)
        if it >= 2:
            res[tune].append(t)
for tune in variants:
    med = statistics.median(res[tune])
    print(f"tune={tune} {NAMES.get(tune, '')}: median {med:.2f} ms per replay ({med * 1e3 / nl:.2f} us per launch), min {min(res[tune]):.2f}, {nl} launches")
