"""In-tree build of libkvb200.so (sm_100a only).

    python -m kvcached_b200.build [--force]

nvcc cross-compiles without a GPU; the built library sits next to the package so that it travels
with the repo snapshot to the GPU box.  cudart is linked statically and the CUDA driver API is
resolved at run time, so the library loads (and exports every symbol of include/kvb200.h) on a
driver-less box too.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkvb200.so")
SOURCES = ["kvb_kernels.cu", "kvb_manager.cpp", "kvb_vmm.cpp", "kvb_capi.cpp", "kvb_shm.cpp"]
HEADERS = ["kvb_internal.h", "kvb_vmm.h", "kvb_shm.h", os.path.join("..", "..", "include", "kvb200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
         "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=default,-Wall,-Wno-unused-function",
         "-x", "cu", "-cudart", "static", "-shared"]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB + ".tmp", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout[-4000:] + "\n" + r.stderr[-8000:])
    if verbose:
        print(r.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
