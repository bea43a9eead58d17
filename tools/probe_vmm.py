#!/usr/bin/env python
"""Per-call CUDA VMM driver costs on this box (kvb_vmm_probe)."""
import os
os.environ.setdefault("KVB200_LIB", "test")   # the driver probes live in libkvb200_test.so
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvcached_b200 import _lib
lib = _lib.load()
keys = ["create", "map", "setaccess_unit", "setaccess_range_once", "unmap_unit", "unmap_range_once", "release", "remap_cycle_unit"]
out = {}
for unit_mb, n, iters in [(2, 1, 20), (2, 8, 10), (2, 64, 5), (128, 1, 10), (128, 8, 5), (128, 32, 3), (640, 4, 3)]:
    o = np.zeros(8)
    rc = lib.kvb_vmm_probe(0, unit_mb << 20, n, iters, o.ctypes.data_as(_lib.p_f64))
    assert rc == 0, _lib.last_error()
    out[f"unit{unit_mb}MiB_n{n}"] = {k: round(float(v), 1) for k, v in zip(keys, o)}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_vmm.json", "w"), indent=1)
