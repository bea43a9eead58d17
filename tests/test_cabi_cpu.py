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


def _header_functions(name="kvb200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
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


def test_product_library_carries_no_test_hooks():
    """Mock driver and driver probes live in libkvb200_test.so (include/kvb200_test.h), not in the shipped library."""
    lib = C.CDLL(_lib.LIB_PATH)
    hooks = _header_functions("kvb200_test.h")
    assert sorted(hooks) == sorted(_lib.TEST_SIGNATURES), hooks
    for name in hooks:
        assert not hasattr(lib, name), f"{name} (test hook) is exported by the product library"
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "mock" not in syms.lower() and "probe" not in syms.lower()
    tlib = C.CDLL(_lib.TEST_LIB_PATH)
    for name in hooks + _header_functions():
        assert hasattr(tlib, name), f"{name} missing from libkvb200_test.so"


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


def test_request_kinds_and_status_codes_match_the_header():
    """Every KVB_OP_* / KVB_ST_* value of include/kvb200.h has the same value in the ctypes binding."""
    import re
    from kvcached_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "kvb200.h")).read()
    ops = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define KVB_OP_(\w+) (\d+)", hdr)}
    assert len(ops) >= 24
    py = {"ALLOC": _lib.OP_ALLOC, "FREE_IDS": _lib.OP_FREE_IDS, "FREE_REF": _lib.OP_FREE_REF, "RESIZE": _lib.OP_RESIZE,
          "TRIM": _lib.OP_TRIM, "TRY_RESERVE": _lib.OP_TRY_RESERVE, "FREE_RESERVED": _lib.OP_FREE_RESERVED,
          "AVAIL": _lib.OP_AVAIL, "CLEAR": _lib.OP_CLEAR, "ALLOC_PAGE": _lib.OP_ALLOC_PAGE, "FREE_PAGES": _lib.OP_FREE_PAGES,
          "PA_RESIZE": _lib.OP_PA_RESIZE, "RESET_ORDER": _lib.OP_RESET_ORDER, "PIN_NEW": _lib.OP_PIN_NEW,
          "PIN_TOUCH": _lib.OP_PIN_TOUCH, "PIN_MARK": _lib.OP_PIN_MARK, "PIN_FREE": _lib.OP_PIN_FREE,
          "PIN_EVICT": _lib.OP_PIN_EVICT, "PIN_RESET": _lib.OP_PIN_RESET}
    for k, v in py.items():
        assert ops[k] == v, k
    assert sorted(ops.values()) == list(range(1, len(ops) + 1)), "request kinds are dense"
    st = {m.group(1): int(m.group(2).strip("()")) for m in re.finditer(r"#define KVB_ST_(\w+) (\(?-?\d+\)?)", hdr)}
    assert (st["OK"], st["FALSE"], st["NONE"], st["NO_PAGES"], st["ASSERT"]) == \
           (_lib.ST_OK, _lib.ST_FALSE, _lib.ST_NONE, _lib.ST_NO_PAGES, _lib.ST_ASSERT)
    import ctypes as C
    assert C.sizeof(_lib.Res) == 32 and C.sizeof(_lib.Op) == 16 and C.sizeof(_lib.Ref) == 12
