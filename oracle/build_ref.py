#!/usr/bin/env python
"""Build the UNMODIFIED reference (kvcached 0.1.5) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under kvcached_b200/ imports this.

What it does (recipe verified in SURVEY.md Appendix C):
  1. compiles /root/reference/csrc/*.cpp where they lie (no source is copied)
     into oracle/_ref/pkg/kvcached/vmm_ops.so with the system g++;
  2. installs the reference's pure-Python package next to it
     (oracle/_ref/pkg/kvcached/, the equivalent of `pip install --target`);
  3. builds the two helper objects the CPU-only container needs to *load* the
     extension without a driver: a libcuda.so.1 symlink to the toolkit stub and
     a cudaMemGetInfo shim (LD_PRELOAD) that reports 180 GiB free / 192 GiB.

oracle/_ref/ is git-ignored, NOT gpurun-ignored: the built files travel to the
GPU box, where the reference runs unshimmed on cuda:0 as the CPU baseline.
/root/reference itself is only needed here, at build time.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KVB_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")

SHIM_C = r"""
#include <stddef.h>
/* CPU-only container: the reference's available_size() calls cudaMemGetInfo
 * (csrc/page_allocator.cpp:442-444) and aborts without a driver. */
int cudaMemGetInfo(size_t *f, size_t *t) { *t = 192ull << 30; *f = 180ull << 30; return 0; }
int cudaDeviceSynchronize(void) { return 0; }
"""


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF, "csrc"))


def built() -> bool:
    return os.path.isfile(os.path.join(OUT, "pkg", "kvcached", "vmm_ops.so"))


def _make_writable(root: str) -> None:
    for d, _, files in os.walk(root):
        os.chmod(d, 0o755)
        for f in files:
            os.chmod(os.path.join(d, f), 0o644)


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    if built() and not force:
        return True
    if not have_reference():
        return False
    os.makedirs(os.path.join(OUT, "build"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "stub"), exist_ok=True)
    pkg = os.path.join(OUT, "pkg")

    # (2) python package (install step; sources are never committed)
    if os.path.isdir(pkg):
        _make_writable(pkg)
        shutil.rmtree(pkg)
    shutil.copytree(os.path.join(REF, "kvcached"), os.path.join(pkg, "kvcached"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.so"))
    _make_writable(pkg)   # /root/reference is mounted read-only; the copy must not be

    # (1) extension, compiled from the sources in place
    env = dict(os.environ, CXX="/usr/bin/g++", CC="/usr/bin/gcc")
    code = f"""
import glob
from torch.utils.cpp_extension import load
try:
    load(name='vmm_ops', sources=sorted(glob.glob({REF!r} + '/csrc/*.cpp')),
         extra_include_paths=[{REF!r} + '/csrc/inc', '/usr/local/cuda/include'],
         extra_cflags=['-std=c++17', '-DKVCACHED_USE_CUDA', '-O2'],
         extra_ldflags=['-L/usr/local/cuda/lib64/stubs', '-lcuda',
                        '-L/usr/local/cuda/lib64', '-lcudart'],
         build_directory={os.path.join(OUT, 'build')!r}, with_cuda=False, verbose=False)
except ImportError as e:   # libcuda.so.1 missing at import time on a CPU box: expected
    print('import after build failed (expected without a driver):', e)
"""
    subprocess.run([sys.executable, "-c", code], check=True, env=env)
    so = os.path.join(OUT, "build", "vmm_ops.so")
    if not os.path.isfile(so):
        raise RuntimeError("reference extension did not build")
    shutil.copy2(so, os.path.join(pkg, "kvcached", "vmm_ops.so"))

    # (3) loader helpers for the driver-less container
    stub = os.path.join(OUT, "stub", "libcuda.so.1")
    if not os.path.exists(stub):
        os.symlink("/usr/local/cuda/lib64/stubs/libcuda.so", stub)
    shim_c = os.path.join(OUT, "build", "shim.c")
    with open(shim_c, "w") as f:
        f.write(SHIM_C)
    subprocess.run(["/usr/bin/gcc", "-shared", "-fPIC", "-O2", "-o",
                    os.path.join(OUT, "shim.so"), shim_c], check=True)
    return True


def ref_env(cpu_only: bool) -> dict:
    """Environment for a subprocess that imports the reference from oracle/_ref."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(OUT, "pkg") + os.pathsep + env.get("PYTHONPATH", "")
    env.setdefault("KVCACHED_LOG_LEVEL", "ERROR")
    if cpu_only:
        env["LD_PRELOAD"] = os.path.join(OUT, "shim.so")
        env["LD_LIBRARY_PATH"] = os.path.join(OUT, "stub") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "reference not present, nothing built")
