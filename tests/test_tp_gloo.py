"""N>1 host logic on CPU (gloo, world_size 2): the page-event wire format and the one-broadcast-per-step
protocol of kvcached_b200/tp.py.  Rank 0 produces page events with the oracle (test infrastructure), every
rank applies them to its own mapped-page set; the sets must be identical on all ranks after every step and
equal to the oracle's."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from kvcached_b200 import _lib, tp
        from kvcached_b200 import trace as T
        words = int(_lib.load().kvb_tp_buffer_words())
        cfg = T.make_config("llama70b_tp2", 8192)
        tr = T.decode_trace(cfg, 200, seed=1)
        steps = T.split_steps(tr["ops"])
        om = oracle.manager_from_config(cfg) if rank == 0 else None
        results = []
        mapped = set()
        n_events = 0
        for seq, idxs in enumerate(steps):
            if rank == 0:
                for i in idxs:
                    while len(results) < i:
                        results.append(None)
                    results.append(oracle.apply_op(om, tr["ops"][i], results))
                buf = tp.pack_events(seq, om.drain_events(), words)
            else:
                buf = torch.zeros(words, dtype=torch.int64)
            dist.broadcast(buf, src=0)                      # the ONE collective of a scheduler step
            got_seq, events = tp.unpack_events(buf)
            assert got_seq == seq
            for page, unmap in events:
                (mapped.discard if unmap else mapped.add)(page)
            n_events += len(events)
            if seq % 50 == 0 or seq == len(steps) - 1:      # rank-consistency check (debug mode of SURVEY 8(e))
                h = torch.tensor([hash(tuple(sorted(mapped))) & 0x7fffffffffff, len(mapped)], dtype=torch.int64)
                hs = [torch.zeros_like(h) for _ in range(world)]
                dist.all_gather(hs, h)
                assert all(torch.equal(x, hs[0]) for x in hs)
                if rank == 0:
                    assert sorted(mapped) == om.mapped_pages()
        q.put((rank, n_events, len(mapped)))
    finally:
        dist.destroy_process_group()


def test_page_events_reach_every_rank_identically():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    out = sorted(q.get(timeout=5) for _ in range(2))
    assert out[0][1] == out[1][1] and out[0][1] > 0     # both ranks applied the same number of events
    assert out[0][2] == out[1][2]


def test_wire_format_round_trip():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from kvcached_b200 import _lib, tp
    words = int(_lib.load().kvb_tp_buffer_words())
    ev = [(5, 0), (6, 0), (5, 1), (1429, 0)]
    buf = tp.pack_events(77, ev, words)
    assert buf.numel() == words and tp.unpack_events(buf) == (77, ev)
    assert tp.unpack_events(tp.pack_events(3, [], words)) == (3, [])
