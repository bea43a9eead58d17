"""Host logic of the VMM engine (kvcached_b200/csrc/kvb_vmm.cpp: wanted / hinted / physical state, cancellation,
map-ahead, retention, steal order, reserve, settle) against a mock CUDA driver that checks the legality of every call
(kvb_vmm_mock_driver, include/kvb200_test.h, exported by libkvb200_test.so only).  No GPU needed.  Each case runs in its own process: installing the mock is process-wide."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case", ["sync", "cancel_and_retention", "steal_and_reserve", "map_ahead", "random_per_layer", "oom", "layouts",
                                  "quiet_and_contended_reserve", "hints_without_memory", "retention_cap", "background_init", "no_zero_page"])
def test_vmm_engine_against_mock_driver(case):
    env = dict(os.environ)
    for k in ("KVB200_VMM_ASYNC", "KVB200_VMM_PREFETCH", "KVB200_VMM_RETAIN_MS", "KVB200_RETAIN_FREE_FRACTION"):
        env.pop(k, None)
    env["KVB200_LIB"] = "test"     # libkvb200_test.so: the product library carries no mock driver
    r = subprocess.run([sys.executable, os.path.join(HERE, "vmm_engine_child.py"), case], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and f"OK {case}" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
