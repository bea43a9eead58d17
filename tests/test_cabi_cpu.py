"""No GPU needed: the C-ABI library loads on a driver-less box and exports every symbol include/kvb200.h
declares; host-side helpers (page arithmetic, trace compilation, config errors) behave like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from kvcached_b200 import _lib
from kvcached_b200 import replay as R
from kvcached_b200 import trace as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "kvb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(kvb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n not in ("kvb_mgr",)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) > 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kvb200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in kvcached_b200/_lib.py"
    assert lib.kvb_abi_version() == 1
    assert lib.kvb_total_launches() == 0 or lib.kvb_total_launches() > 0


def test_library_has_no_link_time_driver_dependency():
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out, out


def test_page_arithmetic_matches_reference_known_answers():
    lib = _lib.load()
    s, e = C.c_int64(), C.c_int64()
    for pid, bms, want in [(0, 32768, (0, 64)), (3, 32768, (192, 256)), (0, 18432, (0, 113)), (1, 18432, (114, 227))]:
        lib.kvb_page_block_range(pid, 2 << 20, bms, C.byref(s), C.byref(e))
        assert (s.value, e.value) == want      # SURVEY Appendix B, InternalPage facts
    assert lib.kvb_page_num_blocks(2 << 20, 18432) == 113
    assert lib.kvb_page_id_of_block(113, 18432, 2 << 20) == 0
    assert lib.kvb_page_id_of_block(114, 18432, 2 << 20) == 1


def test_manager_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kvcached_b200 import KVCacheManager
    with pytest.raises(Exception):
        m = KVCacheManager(1024, 16, 2048, 4, attach_vmm=False)
        m.alloc(1)      # no CPU fallback: without the CUDA device nothing is computed


def test_block_larger_than_page_is_a_config_error():
    from kvcached_b200 import KVCachedConfigError, KVCacheManager
    with pytest.raises(KVCachedConfigError):
        KVCacheManager(16, 16, (4 << 20) // 16, 2, attach_vmm=False)


def test_init_rejects_page_sizes_that_are_not_multiples_of_2mb():
    lib = _lib.load()
    assert lib.kvb_init(0, 3 << 20, 1) < 0
    assert b"multiple of 2MB" in lib.kvb_last_error()
    assert lib.kvb_is_initialized() == 0


def test_trace_compilation_and_host_plan_shapes():
    cfg = T.make_config("llama8b", 4096, num_layers=4)
    tr = T.decode_trace(cfg, 96, seed=0)
    ct = R.compile_trace(tr)
    assert ct.n_calls == tr["meta"]["n_ops"]
    assert ct.step_offsets[0] == 0 and ct.step_offsets[-1] == ct.ops.shape[0]
    assert (np.diff(ct.step_offsets) > 0).all()
    plan = R.HostStepPlan(ct)
    assert sum(st["ops"].shape[0] for st in plan.steps) == ct.ops.shape[0]
    # every free of a launch only names requests of earlier launches
    for st in plan.steps:
        if st["ref_q"].size:
            assert st["ref_q"].max() < st["b"]
    # the random fixtures free same-step results: the plan must cut there
    tr2 = T.random_trace(cfg, 5, n_ops=200)
    plan2 = R.HostStepPlan(R.compile_trace(tr2))
    for st in plan2.steps:
        if st["ref_q"].size:
            assert st["ref_q"].max() < st["b"]


def test_decode_trace_is_balanced_in_the_oracle():
    cfg = T.make_config("llama8b", 8192)
    tr = T.decode_trace(cfg, 256, seed=3)
    m = oracle.manager_from_config(cfg)
    a0 = m.available_size()
    res = []
    for op in tr["ops"]:
        res.append(oracle.apply_op(m, op, res))
    assert m.available_size() == a0 and m.scalars()["inuse_pages"] == 0
