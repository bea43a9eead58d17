"""In-tree build of libkvb200.so (sm_100a only) and of libkvb200_test.so, its test-only twin.

    python -m kvcached_b200.build [--force] [-v]

nvcc cross-compiles without a GPU; the built libraries sit next to the package so that they travel
with the repo snapshot to the GPU box.  cudart is linked statically and the CUDA driver API is
resolved at run time, so the libraries load (and export every symbol of include/kvb200.h) on a
driver-less box too.

libkvb200.so       the product: include/kvb200.h, nothing else.
libkvb200_test.so  the same objects, with kvb_vmm.cpp / kvb_capi.cpp compiled -DKVB_TEST_HOOKS: adds the mock
                   VMM driver and the driver-cost probes of include/kvb200_test.h (tests/, tools/ only).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# KVB200_BUILD_SUFFIX=_prof KVB200_NVCC_FLAGS=-DKVB_PROFILE builds libkvb200_prof.so beside the product (tuning builds)
SUFFIX = os.environ.get("KVB200_BUILD_SUFFIX", "")
OBJ = os.path.join(HERE, "_obj" + SUFFIX)
LIB = os.path.join(HERE, f"libkvb200{SUFFIX}.so")
LIB_TEST = os.path.join(HERE, f"libkvb200{SUFFIX}_test.so")
SOURCES = ["kvb_kernels.cu", "kvb_manager.cpp", "kvb_vmm.cpp", "kvb_capi.cpp", "kvb_shm.cpp", "kvb_tp.cpp"]
HOOKED = ["kvb_vmm.cpp", "kvb_capi.cpp"]            # recompiled with -DKVB_TEST_HOOKS for the test library
HEADERS = ["kvb_internal.h", "kvb_vmm.h", "kvb_shm.h", "kvb_tp.h", "kvb_testhooks_mock.inc", "kvb_testhooks_probe.inc",
           os.path.join("..", "..", "include", "kvb200.h"), os.path.join("..", "..", "include", "kvb200_test.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
          "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=default,-Wall,-Wno-unused-function", "-x", "cu"]
LFLAGS = ["-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-shared"]


def _stale(target: str) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.isfile(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose=False):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout[-4000:] + "\n" + r.stderr[-8000:])
    if verbose and r.stderr.strip():
        print(r.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale(LIB) and not _stale(LIB_TEST):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + [f for f in os.environ.get("KVB200_NVCC_FLAGS", "").split() if f]
    jobs = [(s, os.path.join(OBJ, s + ".o"), []) for s in SOURCES]
    jobs += [(s, os.path.join(OBJ, s + ".hooks.o"), ["-DKVB_TEST_HOOKS"]) for s in HOOKED]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        futs = [ex.submit(_run, [NVCC] + CFLAGS + extra + d + ["-c", os.path.join(CSRC, s), "-o", o],
                          verbose and s.endswith(".cu")) for s, o, d in jobs]
        for f in futs:
            f.result()
    prod = [os.path.join(OBJ, s + ".o") for s in SOURCES]
    test = [os.path.join(OBJ, s + (".hooks.o" if s in HOOKED else ".o")) for s in SOURCES]
    for objs, lib in ((prod, LIB), (test, LIB_TEST)):
        _run([NVCC] + LFLAGS + objs + ["-o", lib + ".tmp", "-lpthread", "-ldl", "-lrt"])
        os.replace(lib + ".tmp", lib)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
