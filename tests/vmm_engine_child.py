"""Child process of tests/test_vmm_engine_cpu.py: drives the VMM engine (kvb_vmm.cpp) against the mock driver.
Runs in its own process because installing the mock is process-wide and irreversible."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kvcached_b200 import _lib, vmm_ops  # noqa: E402

PAGE = 2 << 20
lib = _lib.load()


def mock_stats():
    o = np.zeros(10, dtype=np.int64)
    lib.kvb_vmm_mock_stats(o.ctypes.data_as(_lib.p_i64))
    keys = ["creates", "releases", "maps", "unmaps", "access", "violations", "handles", "mappings", "_", "handle_bytes"]
    return dict(zip(keys, o.tolist()))


def events(pages, unmap):
    a = np.asarray(pages, dtype=np.int64) * 2 + (1 if unmap else 0)
    assert lib.kvb_vmm_enqueue_page_events(0, a.ctypes.data_as(_lib.p_i64), a.size) == 1, _lib.last_error()


def hints(pages):
    a = np.asarray(pages, dtype=np.int64)
    assert lib.kvb_vmm_set_prefetch_hints(0, a.ctypes.data_as(_lib.p_i64), a.size) == 0


def wait_quiet(pred, seconds=20):
    t_end = time.time() + seconds
    while time.time() < t_end:
        if pred():
            return True
        time.sleep(0.005)
    return False


def no_violations():
    st = mock_stats()
    assert st["violations"] == 0, lib.kvb_vmm_mock_last_violation()
    return st


def setup(total_gib=64, latency_us=0, n_pages=64, layers=4, contiguous=True):
    assert lib.kvb_vmm_mock_driver(int(total_gib * (1 << 30)), latency_us) == 0
    vmm_ops.init_kvcached("cuda:0", PAGE, contiguous)
    unit = PAGE * layers * 2
    rngs = vmm_ops.create_kv_tensors_raw(n_pages * PAGE * 2, layers, 2)
    st = no_violations()
    # zero page: one handle, mapped over every unit, access granted
    units = n_pages * (1 if contiguous else layers * 2)
    assert st["creates"] == 1 and st["maps"] == units and st["access"] == units and st["mappings"] == units
    vmm_ops.set_retain_max_fraction(1.0)     # the cases below model tiny devices; case_retention_cap tests the cap
    lib.kvb_vmm_pool_config(0, n_pages)      # ... and so does the handle pool's default cap (10 % of the device)
    return unit, rngs


def case_sync():
    """KVB200_VMM_ASYNC=0 semantics: every call settles; each transition is unmap + map + set-access on one unit."""
    unit, _ = setup()
    vmm_ops.set_async(False)
    vmm_ops.set_prefetch(0)
    base = mock_stats()
    events([3, 4, 5], False)
    vmm_ops.fence()
    st = no_violations()
    assert vmm_ops.mapped_pages() == [3, 4, 5]
    assert (st["unmaps"] - base["unmaps"], st["maps"] - base["maps"], st["access"] - base["access"], st["creates"] - base["creates"]) == (3, 3, 3, 3)
    events([4], True)
    vmm_ops.fence()
    st2 = no_violations()
    assert vmm_ops.mapped_pages() == [3, 5] and vmm_ops.vmm_stats()["pooled_handles"] == 1
    assert st2["unmaps"] - st["unmaps"] == 1 and st2["maps"] - st["maps"] == 1        # back on the zero page
    events([4], False)                                                                  # pooled handle reused
    events([4], False)                                                                  # double map tolerated
    events([9], True)                                                                   # unmap of an unmapped page tolerated
    vmm_ops.fence()
    st3 = no_violations()
    assert st3["creates"] == st2["creates"] and vmm_ops.mapped_pages() == [3, 4, 5]
    vmm_ops.shutdown_kvcached()
    st4 = no_violations()
    assert st4["handles"] == 0 and st4["mappings"] == 0, st4                            # nothing leaks


def case_cancel_and_retention():
    """Opposite events cancel; retained pages cost nothing when mapped again; settle / time-out give them back."""
    unit, _ = setup(latency_us=300)
    vmm_ops.set_prefetch(0)
    vmm_ops.set_retention(60_000)
    events(range(8), False)
    vmm_ops.wait_maps()
    assert vmm_ops.mapped_pages() == list(range(8))
    a = no_violations()
    events(range(8), True)
    vmm_ops.fence()
    b = no_violations()
    assert b["unmaps"] == a["unmaps"] and vmm_ops.engine_stats()["retained"] == 8 and vmm_ops.mapped_pages() == []
    events(range(8), False)
    vmm_ops.wait_maps()
    c = no_violations()
    assert (c["unmaps"], c["maps"], c["access"], c["creates"]) == (b["unmaps"], b["maps"], b["access"], b["creates"])
    assert vmm_ops.engine_stats()["retention_hits"] == 8 and vmm_ops.mapped_pages() == list(range(8))
    # a burst of map / unmap pairs on fresh pages: the unmap that meets its own pending map cancels it
    ev = []
    for p in range(20, 40):
        ev += [p * 2, p * 2 + 1]
    arr = np.asarray(ev, dtype=np.int64)
    assert lib.kvb_vmm_enqueue_page_events(0, arr.ctypes.data_as(_lib.p_i64), arr.size) == 1
    vmm_ops.fence()
    d = no_violations()
    assert vmm_ops.mapped_pages() == list(range(8))
    assert d["creates"] - c["creates"] <= 2, "pending maps were cancelled, at most the one in flight went through"
    # settle: every retained page goes back to the zero page
    events(range(8), True)
    vmm_ops.settle()
    e = no_violations()
    assert vmm_ops.engine_stats()["retained"] == 0 and vmm_ops.vmm_stats()["mapped_pages"] == 0
    assert e["mappings"] == 64 and all(True for _ in [0])
    # time-out while the allocator is quiet
    vmm_ops.set_retention(100)
    events([1, 2], False)
    vmm_ops.wait_maps()
    events([1, 2], True)
    assert wait_quiet(lambda: vmm_ops.engine_stats()["retained"] == 0 and vmm_ops.vmm_stats()["mapped_pages"] == 0)
    no_violations()
    vmm_ops.shutdown_kvcached()
    f = no_violations()
    assert f["handles"] == 0 and f["mappings"] == 0


def case_retention_cap():
    """Retained pages never hold more than the cap (a fraction of the device): the page unmapped at the cap goes back
    to the zero page at once, and the footprint counters say what is physically held beyond the bookkeeping."""
    layers = 4
    unit = PAGE * layers * 2
    setup(total_gib=100 * unit / (1 << 30), n_pages=200, layers=layers)
    vmm_ops.set_prefetch(0)
    vmm_ops.set_retention(600_000)
    vmm_ops.set_retain_free_fraction(0.0, 0.0)
    vmm_ops.set_retain_max_fraction(0.10)                   # 10 of 100 units
    events(range(40), False)
    vmm_ops.wait_maps()
    fp = vmm_ops.footprint()
    assert fp["physical_pages"] == 40 and fp["wanted_pages"] == 40 and fp["retain_cap_pages"] == 10
    events(range(40), True)
    vmm_ops.fence()
    st = no_violations()
    fp = vmm_ops.footprint()
    assert vmm_ops.engine_stats()["retained"] == 10 and fp["retained_pages"] == 10
    assert fp["physical_pages"] == 10 and fp["wanted_pages"] == 0 and fp["peak_overshoot_pages"] >= 10
    assert st["mappings"] == 200                            # 190 units on the zero page + 10 retained
    # the FIRST ten pages freed are the ones kept (the FIFO free list hands them out first)
    events(range(10), False)
    vmm_ops.wait_maps()
    assert vmm_ops.engine_stats()["retention_hits"] == 10
    # lowering the cap hands pages back; strict mode = cap irrelevant, nothing retained
    events(range(10), True)
    vmm_ops.set_retain_max_fraction(0.03)
    vmm_ops.fence()
    assert vmm_ops.engine_stats()["retained"] == 3
    vmm_ops.set_retention(0)
    vmm_ops.fence()
    fp = vmm_ops.footprint()
    assert fp["physical_pages"] == 0 and fp["retained_pages"] == 0
    vmm_ops.shutdown_kvcached()
    e = no_violations()
    assert e["handles"] == 0 and e["mappings"] == 0


def case_background_init():
    """kvb_vmm_zero_init_mode(1): create backs unit 0 only, the VMM thread the rest; pages mapped meanwhile go straight to
    their handle (no unmap of an unbacked unit), wait_init() leaves every unit backed, and a group torn down half way
    through its init unmaps only what it backed."""
    assert lib.kvb_vmm_mock_driver(64 << 30, 300) == 0          # 300 us per sleeping driver call: init takes a while
    vmm_ops.zero_init_mode(1)
    vmm_ops.init_kvcached("cuda:0", PAGE, False)                # per-layer layout: 4 layers x K,V = 8 units per page
    t0 = time.time()
    vmm_ops.create_kv_tensors_raw(200 * PAGE * 2, 4, 2)
    t_create = time.time() - t0
    st = no_violations()
    assert st["maps"] < 200 and t_create < 0.2, (st, t_create)  # 1600 units eagerly would take ~0.5 s here
    events([150, 151, 3], False)                                # long before the init gets there
    vmm_ops.wait_maps()
    st = no_violations()
    assert vmm_ops.mapped_pages() == [3, 150, 151]
    vmm_ops.wait_init()
    pr = vmm_ops.init_progress()
    assert pr["units_done"] == pr["units"] == 1600
    st = no_violations()
    assert st["mappings"] == 1600 and st["handles"] == 1 + 3 * 8
    events([150, 151, 3], True)
    vmm_ops.fence()
    st = no_violations()
    assert st["mappings"] == 1600 and vmm_ops.mapped_pages() == []
    vmm_ops.shutdown_kvcached()
    e = no_violations()
    assert e["handles"] == 0 and e["mappings"] == 0
    # torn down while the init is still running
    vmm_ops.init_kvcached("cuda:0", PAGE, False)
    vmm_ops.create_kv_tensors_raw(200 * PAGE * 2, 4, 2)
    events([7], False)
    vmm_ops.wait_maps()
    assert vmm_ops.init_progress()["units_done"] < 1600
    vmm_ops.shutdown_kvcached()
    e = no_violations()
    assert e["handles"] == 0 and e["mappings"] == 0
    vmm_ops.zero_init_mode(0)


def case_no_zero_page():
    """kvb_vmm_zero_init_mode(2): nothing is mapped at create; a map is cuMemMap + cuMemSetAccess, an unmap cuMemUnmap
    alone — two sleeping driver calls per map/unmap pair instead of four."""
    assert lib.kvb_vmm_mock_driver(64 << 30, 0) == 0
    vmm_ops.zero_init_mode(2)
    vmm_ops.init_kvcached("cuda:0", PAGE, True)
    vmm_ops.create_kv_tensors_raw(64 * PAGE * 2, 4, 2)
    vmm_ops.set_prefetch(0)
    st = no_violations()
    assert st["maps"] == 0 and st["mappings"] == 0 and st["creates"] == 1
    events([3, 4, 5], False)
    vmm_ops.fence()
    st = no_violations()
    assert st["maps"] == 3 and st["access"] == 3 and st["unmaps"] == 0 and st["mappings"] == 3
    events([4], True)
    vmm_ops.fence()
    st = no_violations()
    assert st["unmaps"] == 1 and st["maps"] == 3 and st["mappings"] == 2 and vmm_ops.mapped_pages() == [3, 5]
    events([4], False)
    vmm_ops.fence()
    assert no_violations()["mappings"] == 3
    vmm_ops.shutdown_kvcached()
    e = no_violations()
    assert e["handles"] == 0 and e["mappings"] == 0
    vmm_ops.zero_init_mode(0)


def case_steal_and_reserve():
    """A device of 40 units (+ the zero page): maps beyond the reserve take the page retained LAST; when somebody else
    takes memory, retained pages are handed back until the reserve is whole."""
    layers = 4
    unit = PAGE * layers * 2
    setup(total_gib=41 * unit / (1 << 30), n_pages=200, layers=layers)
    vmm_ops.set_prefetch(0)
    vmm_ops.set_retention(600_000)
    vmm_ops.set_retain_free_fraction(0.20, 0.15)           # reserve 0.20 both while quiet (headroom 0.05 + 0.15) and contended:
                                                           # 8.2 of 41 units stay free, a create is allowed while free - 1 >= 8.2,
                                                           # i.e. up to 32 handles including the zero page
    events(range(20), False)
    vmm_ops.wait_maps()
    events(range(20), True)                                 # 20 retained
    events(range(20, 30), False)                            # 10 more fit above the reserve (zero + 30 = 31 handles)
    vmm_ops.wait_maps()
    a = no_violations()
    assert a["creates"] == 1 + 30 and vmm_ops.engine_stats()["retained"] == 20
    events(range(30, 34), False)                            # one more handle fits; then the pages retained LAST are taken
    vmm_ops.wait_maps()
    b = no_violations()
    assert b["creates"] == a["creates"] + 1, "one create up to the reserve, three stolen handles"
    assert vmm_ops.engine_stats()["retained"] == 17
    # FIFO free list: pages come back in the order they were freed; the oldest retained ones are still there
    events(range(0, 10), False)
    vmm_ops.wait_maps()
    c = no_violations()
    assert vmm_ops.engine_stats()["retention_hits"] == 10 and c["creates"] == b["creates"] and c["unmaps"] == b["unmaps"]
    # somebody else takes 20 % of the device: the VMM thread hands retained pages back
    lib.kvb_vmm_mock_set_other_bytes(int(8 * unit))
    assert wait_quiet(lambda: vmm_ops.engine_stats()["retained"] == 0 or mock_stats()["handle_bytes"] <= (41 - 8 - 8.2) * unit + unit)
    d = no_violations()
    assert d["releases"] > c["releases"]
    assert sorted(vmm_ops.mapped_pages()) == list(range(0, 10)) + list(range(20, 34))
    vmm_ops.shutdown_kvcached()
    e = no_violations()
    assert e["handles"] == 0 and e["mappings"] == 0


def case_quiet_and_contended_reserve():
    """While nobody else moves, retained pages may fill the device up to headroom + quiet margin; when somebody else's
    footprint grows, the larger reserve applies and the VMM thread hands retained pages back until it is whole."""
    layers = 4
    unit = PAGE * layers * 2
    setup(total_gib=41 * unit / (1 << 30), n_pages=200, layers=layers)
    vmm_ops.set_prefetch(0)
    vmm_ops.set_retention(600_000)
    vmm_ops.set_retain_free_fraction(0.50, 0.0)            # contended: half the device free; quiet: the 5 % headroom
    events(range(30), False)
    vmm_ops.wait_maps()
    events(range(30), True)
    vmm_ops.fence()
    time.sleep(0.3)                                         # several polls: nobody else moves, nothing is handed back
    a = no_violations()
    assert vmm_ops.engine_stats()["retained"] == 30 and a["releases"] == 0
    lib.kvb_vmm_mock_set_other_bytes(int(3 * unit))         # somebody takes three pages' worth
    assert wait_quiet(lambda: mock_stats()["handle_bytes"] <= (41 - 3 - 20.5) * unit)
    b = no_violations()
    assert b["releases"] >= 13 and vmm_ops.engine_stats()["retained"] <= 17
    # the pages retained FIRST survive (the FIFO free list needs them first)
    events(range(0, 8), False)
    vmm_ops.wait_maps()
    assert vmm_ops.engine_stats()["retention_hits"] == 8
    vmm_ops.shutdown_kvcached()
    end = no_violations()
    assert end["handles"] == 0 and end["mappings"] == 0


def case_map_ahead():
    """Hinted pages are mapped in idle time, wanted maps go first, a moved window unmaps (retains) what left it."""
    setup(latency_us=200)
    vmm_ops.set_retention(0)
    vmm_ops.set_prefetch(4)
    hints([10, 11, 12, 13, 14, 15])                        # only `depth` of them are taken
    assert wait_quiet(lambda: vmm_ops.engine_stats()["need_hinted"] == 0 and len(vmm_ops.prefetched_pages()) == 4)
    assert vmm_ops.prefetched_pages() == [10, 11, 12, 13] and vmm_ops.mapped_pages() == []
    a = no_violations()
    events([10, 11], False)                                # already in place
    vmm_ops.wait_maps()
    b = no_violations()
    assert (b["maps"], b["unmaps"], b["access"]) == (a["maps"], a["unmaps"], a["access"])
    assert vmm_ops.engine_stats()["prefetch_hits"] == 2 and vmm_ops.mapped_pages() == [10, 11]
    hints([14, 15, 16, 17])                                # 12, 13 left the window unwanted: back to the zero page
    assert wait_quiet(lambda: vmm_ops.prefetched_pages() == [14, 15, 16, 17] and vmm_ops.pending_unmaps() == 0)
    vmm_ops.fence()
    c = no_violations()
    assert vmm_ops.vmm_stats()["mapped_pages"] == 6 and vmm_ops.mapped_pages() == [10, 11]
    vmm_ops.settle()
    hints([])
    assert wait_quiet(lambda: vmm_ops.vmm_stats()["mapped_pages"] == 2)
    vmm_ops.shutdown_kvcached()
    d = no_violations()
    assert d["handles"] == 0 and d["mappings"] == 0


def case_random_per_layer():
    """Per-layer layout (2 * layers handles per page), random event storm with latency: the engine always settles to
    the wanted state and never makes an illegal driver call."""
    layers = 3
    setup(latency_us=50, n_pages=24, layers=layers, contiguous=False)
    vmm_ops.set_prefetch(3)
    vmm_ops.set_retention(40)
    rng = np.random.default_rng(11)
    want = np.zeros(24, dtype=bool)
    for rnd in range(150):
        pages = rng.integers(0, 24, size=int(rng.integers(1, 6)))
        ev = []
        for p in pages:
            ev.append(int(p) * 2 + (1 if want[p] else 0))
            want[p] = ~want[p]
        arr = np.asarray(ev, dtype=np.int64)
        assert lib.kvb_vmm_enqueue_page_events(0, arr.ctypes.data_as(_lib.p_i64), arr.size) == 1
        if rnd % 3 == 0:
            hints(rng.choice(24, size=3, replace=False))
        if rnd % 5 == 0:
            vmm_ops.wait_maps()
            st = no_violations()
        if rnd % 37 == 0:
            vmm_ops.settle()
    hints([])
    vmm_ops.settle()
    st = no_violations()
    assert vmm_ops.mapped_pages() == np.nonzero(want)[0].tolist()
    assert vmm_ops.vmm_stats()["mapped_pages"] == int(want.sum())
    assert st["handles"] == 1 + st["creates"] - 1 - st["releases"]
    vmm_ops.shutdown_kvcached()
    end = no_violations()
    assert end["handles"] == 0 and end["mappings"] == 0


def case_layouts():
    """Handles per page id by layout (allocator.cpp:161-209): contiguous = 1 compound page of L * nKV * P bytes, per-layer =
    2 * L pages (K at off, V at off + size / 2), unified pool = L pages (K and V interleaved inside the page)."""
    L = 5
    for contiguous, unified, per_page, unit in ((True, False, 1, PAGE * L * 2), (False, False, 2 * L, PAGE), (False, True, L, PAGE)):
        assert lib.kvb_vmm_mock_driver(64 << 30, 0) == 0
        vmm_ops.init_kvcached("cuda:0", PAGE, contiguous)
        rngs = vmm_ops.create_kv_tensors_raw(16 * PAGE * 2, L, 2, 0, unified)
        assert len(rngs) == (1 if contiguous else L)
        n_pages = 16 if not unified else 32               # unified: the whole per-layer range is one pool of pages
        vmm_ops.set_async(False)
        vmm_ops.set_prefetch(0)
        a = no_violations()
        events([2, 7], False)
        vmm_ops.fence()
        b = no_violations()
        assert b["creates"] - a["creates"] == 2 * per_page and b["unmaps"] - a["unmaps"] == 2 * per_page
        assert b["handle_bytes"] - a["handle_bytes"] == 2 * per_page * unit
        assert vmm_ops.mapped_pages() == [2, 7]
        bad = np.asarray([n_pages * 2], dtype=np.int64)   # a page beyond the range is refused, nothing touched
        assert lib.kvb_vmm_enqueue_page_events(0, bad.ctypes.data_as(_lib.p_i64), 1) < 0
        events([2, 7], True)
        vmm_ops.fence()
        assert vmm_ops.mapped_pages() == [] and vmm_ops.vmm_stats()["pooled_handles"] == 2 * per_page
        vmm_ops.shutdown_kvcached()
        end = no_violations()
        assert end["handles"] == 0 and end["mappings"] == 0


def case_hints_without_memory():
    """Map-ahead hints when there is no memory for them and nothing retained to take: the hints are dropped (the VMM
    thread goes back to sleep instead of retrying forever) and wanted maps still surface the driver's refusal."""
    layers = 4
    unit = PAGE * layers * 2
    setup(total_gib=4 * unit / (1 << 30), n_pages=64, layers=layers)     # zero page + 3 units
    vmm_ops.set_prefetch(4)
    vmm_ops.set_retention(0)
    events([0, 1, 2], False)
    vmm_ops.wait_maps()
    hints([10, 11, 12, 13])
    assert wait_quiet(lambda: vmm_ops.engine_stats()["need_hinted"] == 0)
    import resource
    c0 = resource.getrusage(resource.RUSAGE_SELF)
    time.sleep(0.5)
    c1 = resource.getrusage(resource.RUSAGE_SELF)
    assert (c1.ru_utime + c1.ru_stime) - (c0.ru_utime + c0.ru_stime) < 0.2, "the VMM thread must not spin"
    assert vmm_ops.prefetched_pages() == [] and vmm_ops.mapped_pages() == [0, 1, 2]
    no_violations()
    vmm_ops.shutdown_kvcached()
    end = no_violations()
    assert end["handles"] == 0 and end["mappings"] == 0


def case_oom():
    """A device too small for what is wanted, nothing retained to take: the driver's refusal surfaces as an error of the
    call that waited for the map, the page is recorded as not mapped, nothing leaks."""
    layers = 4
    unit = PAGE * layers * 2
    setup(total_gib=5 * unit / (1 << 30), n_pages=64, layers=layers)     # zero page + 4 units
    vmm_ops.set_prefetch(0)
    events(range(10), False)
    rc = lib.kvb_vmm_wait_maps(0)
    assert rc < 0 and "out of memory" in _lib.last_error(), (rc, _lib.last_error())
    lib.kvb_vmm_fence(0)
    got = vmm_ops.mapped_pages()
    assert len(got) == 4 and set(got) <= set(range(10))
    st = no_violations()
    assert st["creates"] == 5
    events(got, True)
    events([20], False)                                  # a retained handle makes room
    vmm_ops.wait_maps()
    assert vmm_ops.mapped_pages() == [20]
    vmm_ops.shutdown_kvcached()
    end = no_violations()
    assert end["handles"] == 0 and end["mappings"] == 0


if __name__ == "__main__":
    globals()["case_" + sys.argv[1]]()
    print("OK", sys.argv[1])
