#!/usr/bin/env python
import os
os.environ.setdefault("KVB200_LIB", "test")   # the driver probes live in libkvb200_test.so
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvcached_b200 import _lib
lib = _lib.load()
out = {}
for unit_mb in (2, 128):
    for M in (16, 256, 1430):
        for mode in (0, 1, 2):
            o = np.zeros(3)
            rc = lib.kvb_vmm_probe2(0, unit_mb << 20, M, mode, 4, o.ctypes.data_as(_lib.p_f64))
            assert rc == 0, _lib.last_error()
            out[f"unit{unit_mb}MiB_M{M}_mode{mode}"] = [round(float(x), 1) for x in o]
            print(f"unit{unit_mb}MiB M={M} mode={mode}: map_cycle={o[0]:.0f}us unmap_cycle={o[1]:.0f}us init={o[2]:.1f}ms", flush=True)
json.dump(out, open("gpurun_out/probe_vmm2.json", "w"), indent=1)
