#!/usr/bin/env python
import os
os.environ.setdefault("KVB200_LIB", "test")   # the driver probes live in libkvb200_test.so
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvcached_b200 import _lib
lib = _lib.load()
for unit_mb, M in [(128, 0), (128, 100), (128, 400), (128, 800), (128, 0), (2, 800), (2, 8000)]:
    o = np.zeros(7)
    rc = lib.kvb_vmm_probe3(0, unit_mb << 20, M, 3, o.ctypes.data_as(_lib.p_f64))
    assert rc == 0, _lib.last_error()
    print(f"unit{unit_mb}MiB M={M}: same {o[0]:.0f}/{o[1]:.0f}  separate {o[2]:.0f}/{o[3]:.0f}  adjacent({int(o[6])}) {o[4]:.0f}/{o[5]:.0f} us", flush=True)
