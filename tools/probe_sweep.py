#!/usr/bin/env python
"""Bitmap sweep (kvb_audit_kernel) on stress sizes: ms and algorithmic GB/s."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvcached_b200 import _lib
lib = _lib.load()
for nb, bpp in [(1 << 30, 64), (1 << 30, 32), (1 << 30, 128), (1 << 31, 64), (1 << 27, 64), (1 << 30, 512)]:
    o = np.zeros(4)
    rc = lib.kvb_audit_stress(0, nb, bpp, 20, o.ctypes.data_as(_lib.p_f64))
    print(f"blocks=2^{int(np.log2(nb))} bpp={bpp}: rc={rc} {o[0]*1e3:.1f} us, {o[1]/2**20:.0f} MiB, {o[1]/o[0]/1e6:.0f} GB/s, used {int(o[2])} (expect {int(o[3])}) {_lib.last_error() if rc else ''}", flush=True)
