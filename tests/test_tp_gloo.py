"""N>1 host logic on CPU (world_size 2 and 3, no GPU): both TP transports of kvcached_b200/tp.py through the C ABI.

Rank 0 produces page events with the oracle (test infrastructure); every other rank applies what reaches it to its
own VMM engine — the real kvb_vmm.cpp engine on the mock driver of libkvb200_test.so — and its physically mapped
page set must equal the oracle's bookkeeping after every checked step.

  * collective: kvb_tp_pack -> torch.distributed.broadcast (gloo) -> kvb_tp_apply_host, including a page list longer
    than one buffer (clear() of a pool of more than 1100 pages: the chunked path);
  * mailbox: kvb_tp_publish -> shared-memory ring -> follower thread (kvb_tp.cpp), no collective at all, including
    lists longer than one ring slot and more records than the ring has slots (back-pressure).
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAGE = 2 << 20


def _setup_paths():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["KVB200_LIB"] = "test"          # mock VMM driver: libkvb200_test.so
    os.environ["KVB200_VMM_RETAIN_MS"] = "0"   # strict: the physical mapping follows the bookkeeping


def _mock_group(n_pages, layers=2):
    from kvcached_b200 import _lib, vmm_ops
    lib = _lib.load()
    assert lib.kvb_vmm_mock_driver(64 << 30, 0) == 0
    vmm_ops.init_kvcached("cuda:0", PAGE, True)
    vmm_ops.create_kv_tensors_raw(n_pages * PAGE * 2, layers, 2)
    vmm_ops.set_prefetch(0)
    return lib, vmm_ops


def _steps_with_a_big_clear(n_pages):
    """Page-event lists per scheduler step: a decode trace through the oracle, then every page mapped and a clear()
    that unmaps more than 1100 pages in one list."""
    import oracle
    from kvcached_b200 import trace as T
    cfg = T.make_config("llama70b_tp2", n_pages * 128)      # 128 blocks per page
    tr = T.decode_trace(cfg, 120, seed=1)
    steps = T.split_steps(tr["ops"])
    om = oracle.manager_from_config(cfg)
    results, out = [], []
    for idxs in steps:
        for i in idxs:
            while len(results) < i:
                results.append(None)
            results.append(oracle.apply_op(om, tr["ops"][i], results))
        out.append((om.drain_events(), om.mapped_pages()))
    ids = om.alloc(n_pages * 128 - 200)                     # nearly the whole pool: > 1100 pages mapped
    out.append((om.drain_events(), om.mapped_pages()))
    assert len(out[-1][1]) > 1100
    om.free(ids)
    om.trim()
    out.append((om.drain_events(), om.mapped_pages()))
    assert len(out[-1][0]) > 1100 and out[-1][1] == []
    return out


def _collective_worker(rank, world, port, q):
    _setup_paths()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from kvcached_b200 import _lib, tp
        n_pages = 1300
        lib, vmm_ops = _mock_group(n_pages)
        words = int(lib.kvb_tp_buffer_words())
        steps = _steps_with_a_big_clear(n_pages) if rank == 0 else None
        n_steps = torch.tensor([len(steps) if rank == 0 else 0])
        dist.broadcast(n_steps, src=0)
        n_bufs = n_events = 0
        for s in range(int(n_steps)):
            chunks = tp.pack_events(s, steps[s][0], hints=[1, 2, 3]) if rank == 0 else None
            k = 0
            while True:
                buf = chunks[k] if rank == 0 else torch.zeros(words, dtype=torch.int64)
                dist.broadcast(buf, src=0)
                n_bufs += 1
                seq, n = C.c_int64(0), C.c_int64(0)
                more = lib.kvb_tp_apply_host(0, C.cast(buf.data_ptr(), _lib.p_i64), C.byref(seq), C.byref(n))
                assert more >= 0, _lib.last_error()
                assert seq.value == s
                n_events += n.value
                k += 1
                if not more:
                    break
            if rank == 0:
                assert k == len(chunks)
            want = torch.tensor(steps[s][1] + [-1] * (n_pages - len(steps[s][1])), dtype=torch.int64) if rank == 0 \
                else torch.zeros(n_pages, dtype=torch.int64)
            dist.broadcast(want, src=0)
            vmm_ops.fence()
            assert vmm_ops.mapped_pages() == [p for p in want.tolist() if p >= 0], f"rank {rank} step {s}"
        o = __import__("numpy").zeros(10, dtype="int64")
        lib.kvb_vmm_mock_stats(o.ctypes.data_as(_lib.p_i64))
        assert o[5] == 0, lib.kvb_vmm_mock_last_violation()
        q.put((rank, n_bufs, n_events))
    finally:
        dist.destroy_process_group()


def _run(target, world, extra=()):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=target, args=(r, world, port, q) + tuple(extra)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    return sorted(q.get(timeout=5) for _ in range(world))


def test_collective_path_reaches_every_rank_identically_including_chunked_lists():
    out = _run(_collective_worker, 2)
    assert out[0][1] == out[1][1] and out[0][2] == out[1][2] and out[0][2] > 2200
    assert out[0][1] > 120 + 2 + 2      # the two big lists took more than one buffer each


def _mailbox_worker(rank, world, port, q, name):
    _setup_paths()
    try:
        from kvcached_b200 import tp
        n_pages = 1300
        lib, vmm_ops = _mock_group(n_pages)
        mb = tp.TpMailbox(name, rank, world, external_producer=(rank == 0))
        if rank == 0:
            steps = _steps_with_a_big_clear(n_pages)
            for s, (ev, mapped) in enumerate(steps):
                mb.publish(ev, hints=[5, 6], seq=s)
            # more records than the ring has slots: the producer must wait for the followers (back-pressure)
            for k in range(1500):
                mb.publish([(k % 7, k & 1)], seq=10_000 + k)
            mb.publish([(p, 1) for p in range(7)], seq=20_000)
            mb.fence()
            st = mb.stats()
            assert st["published"] == st["applied"] and st["relay_records"] == st["published"] and st["device_direct"] == 0
            assert st["published"] > len(steps) + 1500 + 8      # the big lists were cut into several records
            q.put((rank, st["published"], sum(len(e) for e, _ in steps) + 1507, [m for _, m in steps][-2][:5]))
            open(f"/dev/shm/{name}.done0", "w").close()      # everything is published and acknowledged
            # keep the segment alive until the followers have checked
            import time
            t_end = time.time() + 60
            while time.time() < t_end and not all(os.path.exists(f"/dev/shm/{name}.done{r}") for r in range(world)):
                time.sleep(0.01)
        else:
            import time
            t_end = time.time() + 120
            while time.time() < t_end and not os.path.exists(f"/dev/shm/{name}.done0"):
                time.sleep(0.01)
            mb.fence()
            vmm_ops.fence()
            st = mb.stats()
            assert vmm_ops.mapped_pages() == [], vmm_ops.mapped_pages()[:10]
            import numpy as np
            from kvcached_b200 import _lib
            o = np.zeros(10, dtype=np.int64)
            lib.kvb_vmm_mock_stats(o.ctypes.data_as(_lib.p_i64))
            assert o[5] == 0, lib.kvb_vmm_mock_last_violation()
            q.put((rank, st["applied"], st["events_applied"], []))
            open(f"/dev/shm/{name}.done{rank}", "w").close()
        mb.close()
    finally:
        if rank != 0:
            pass


@pytest.mark.parametrize("world", [2, 3])
def test_mailbox_path_without_a_collective(world):
    name = f"kvb200_tp_test_{os.getpid()}_{world}"
    try:
        out = _run(_mailbox_worker, world, extra=(name,))
    finally:
        for r in range(world):
            try:
                os.unlink(f"/dev/shm/{name}.done{r}")
            except OSError:
                pass
    pub, n_events = out[0][1], out[0][2]
    for r in range(1, world):
        assert out[r][1] == pub            # every follower applied every record
        assert out[r][2] == n_events       # ... and every event


def test_wire_format_round_trip():
    _setup_paths()
    from kvcached_b200 import tp
    ev = [(5, 0), (6, 0), (5, 1), (1429, 0)]
    (buf,) = tp.pack_events(77, ev)
    assert tp.unpack_events(buf) == (77, ev) and int(buf[-1]) == 0
    (empty,) = tp.pack_events(3, [])
    assert tp.unpack_events(empty) == (3, [])
    chunks = tp.pack_events(9, [(p, 0) for p in range(2500)])
    assert len(chunks) == 3 and [int(c[-1]) for c in chunks] == [1, 1, 0]
    got = [e for c in chunks for e in tp.unpack_events(c)[1]]
    assert got == [(p, 0) for p in range(2500)]
