"""The usage segment's wire format (csrc/inc/mem_info_tracker.hpp:25-36): /dev/shm/<name>, three int64
{total_size, used_size, prealloc_size}, flock-protected — read and written here the way kvctl / kvtop do it
(kvcached/cli/utils.py:29-153), and cross-checked with the reference's own Python helpers when oracle/_ref is built."""
import fcntl
import mmap
import os
import subprocess
import sys

import numpy as np
import pytest

from kvcached_b200 import kv_cache_manager as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kvctl_read(name):
    """RwLockedShm(RLOCK) + MemInfoStruct.from_buffer, restated (cli/utils.py:36-52, :56-99)."""
    with open("/dev/shm/" + name, "r+b") as f:
        fcntl.flock(f, fcntl.LOCK_SH)
        mm = mmap.mmap(f.fileno(), 24, access=mmap.ACCESS_READ)
        v = tuple(int(x) for x in np.frombuffer(mm, dtype="<i8", count=3))
        mm.close()
        fcntl.flock(f, fcntl.LOCK_UN)
    return v


def test_set_limit_and_read_round_trip():
    name = f"kvb200_shm_test_{os.getpid()}"
    try:
        assert K.shm_read(name) is None                       # FileNotFoundError -> None (cli/utils.py:110-113)
        assert not K.shm_set_limit(name, 1 << 30)             # update of a missing segment fails, nothing created
        assert not os.path.exists("/dev/shm/" + name)
        assert K.shm_set_limit(name, 10_000_000, create=True)
        assert os.path.getsize("/dev/shm/" + name) == 24
        assert K.shm_read(name) == (10_000_000, 0, 0) == kvctl_read(name)
        with open("/dev/shm/" + name, "r+b") as f:            # a writer charging 100 bytes the reference's way
            fcntl.flock(f, fcntl.LOCK_EX)
            mm = mmap.mmap(f.fileno(), 24)
            a = np.frombuffer(mm, dtype="<i8", count=3)
            a[1] += 100
            del a
            mm.close()
            fcntl.flock(f, fcntl.LOCK_UN)
        assert K.shm_read(name) == (10_000_000, 100, 0)
        assert K.shm_set_limit(name, 5_000_000)               # only total_size changes (cli/utils.py:124-131)
        assert K.shm_read(name) == (5_000_000, 100, 0)
    finally:
        if os.path.exists("/dev/shm/" + name):
            os.unlink("/dev/shm/" + name)


def test_reference_cli_helpers_read_our_segment():
    """get_kv_cache_limit / update_kv_cache_limit of the unmodified reference against a segment we wrote."""
    ref_pkg = os.path.join(ROOT, "oracle", "_ref", "pkg")
    if not os.path.isfile(os.path.join(ref_pkg, "kvcached", "cli", "utils.py")):
        pytest.skip("oracle/_ref not built")
    name = f"kvb200_shm_ref_{os.getpid()}"
    code = f"""
import sys
from unittest.mock import MagicMock
sys.modules.setdefault("posix_ipc", MagicMock())
sys.path.insert(0, {ref_pkg!r})
from kvcached.cli.utils import get_kv_cache_limit, update_kv_cache_limit
m = get_kv_cache_limit({name!r})
print(m.total_size, m.used_size, m.prealloc_size)
update_kv_cache_limit({name!r}, 777)
"""
    try:
        assert K.shm_set_limit(name, 4242, create=True)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-2000:]
        assert r.stdout.splitlines()[0].split() == ["4242", "0", "0"]
        assert K.shm_read(name) == (777, 0, 0)
    finally:
        if os.path.exists("/dev/shm/" + name):
            os.unlink("/dev/shm/" + name)
