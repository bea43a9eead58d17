"""The tensor-parallel page path through the C ABI on real devices: rank 0's step KERNEL pushes every step's page
list into the shared ring (kvb_tp_init, device-direct), the follower rank's thread maps the same pages into its own
virtual KV tensors, and after every scheduler step

    follower's PHYSICAL mapping == rank 0's physical mapping == the oracle's mapped set   (strict retention)

Two processes, like two TP workers.  With one GPU both ranks live on cuda:0 (each process has its own VA ranges and
physical handles — exactly what two ranks on two GPUs have); with >= 2 GPUs the follower runs on cuda:1 and the
collective path (kvb_mgr_tp_export -> NCCL broadcast -> kvb_tp_apply) is checked against the same expectations.
Replaces what kvcached/tp_ipc_util.py:96-145, :250-267 does with sockets."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAGE = 2 << 20


def _paths():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["KVB200_VMM_RETAIN_MS"] = "0"       # strict: physical mapping == bookkeeping at every fence
    os.environ["KVB200_TP_TIMEOUT_MS"] = "120000"


def _cfg(big):
    from kvcached_b200 import trace as T
    # 2 layers -> 8 MiB compound pages; big: 1200 pages (9.4 GiB when all mapped) for the > 1100-page clear()
    return T.make_config("llama70b_tp2", (1200 if big else 96) * 128, num_layers=2)


def _open_group(cfg, dev):
    from kvcached_b200 import vmm_ops
    torch.cuda.set_device(dev)
    torch.cuda.init()
    vmm_ops.init_kvcached(f"cuda:{dev}", PAGE, True)
    per_layer = cfg["num_blocks"] * cfg["block_size"] * cfg["cell_size"] * 2
    ranges = vmm_ops.create_kv_tensors_raw(per_layer, cfg["num_layers"], 2)
    return vmm_ops, ranges


def _rank0(name, world, dev, big, relay, to_f, from_f, result):
    try:
        _paths()
        import oracle
        from kvcached_b200 import kv_ops, tp
        from kvcached_b200 import replay as R
        from kvcached_b200 import trace as T
        cfg = _cfg(big)
        vmm_ops, ranges = _open_group(cfg, dev)
        m = R.manager_from_config(cfg, device_index=dev)
        m.set_phys_free_pages(1 << 40)
        om = oracle.manager_from_config(cfg)
        pre = m.alloc(3)                              # pages mapped BEFORE the mailbox exists must reach the followers
        assert pre == om.alloc(3)
        mb = tp.TpMailbox(name, 0, world, mgr=m, host_relay=relay)
        trace = T.decode_trace(cfg, 60 if big else 90, seed=3)
        steps = T.split_steps(trace["ops"])
        res_o, res_m = [], []
        checks = 0

        def check(tag):
            nonlocal checks
            want = om.mapped_pages()
            vmm_ops.fence()
            assert sorted(vmm_ops.mapped_pages()) == want, f"rank 0 {tag}"
            phys = vmm_ops.vmm_stats()["mapped_pages"]      # everything physically mapped: wanted + mapped ahead (<= 5)
            assert len(want) <= phys <= len(want) + 5, f"rank 0 holds pages the bookkeeping does not: {tag}"
            mb.fence()
            for q in to_f:
                q.put(("check", tag, want))
            for q in from_f:
                ok, msg = q.get(timeout=120)
                assert ok, msg
            checks += 1

        for s, idxs in enumerate(steps):
            for i in idxs:
                while len(res_o) < i:
                    res_o.append(None)
                    res_m.append(None)
                res_o.append(oracle.apply_op(om, trace["ops"][i], res_o))
                res_m.append(R.apply_op(m, trace["ops"][i], res_m))
                assert res_m[-1] == res_o[-1]
            if s % 7 == 0 or s == len(steps) - 1:
                check(f"step {s}")
        # one launch with many requests (the batched path), then trim: settle record
        frees = []
        ids = m.step(frees, [5, 700, 33])
        for n, got in zip([5, 700, 33], ids):
            assert got.tolist() == om.alloc(n)
        check("batched step")
        if big:
            n = cfg["num_blocks"] - 128 * 20
            a = m.alloc(n)
            assert a == om.alloc(n) and len(om.mapped_pages()) > 1100
            check("pool nearly full")
            m.clear()                                  # > 1100 unmaps in ONE list: chunked host relay
            om.clear()
            check("clear of > 1100 pages")
        else:
            m.trim(); om.trim()
            check("trim")
        # the follower writes/reads through ITS mapping of the pages rank 0 allocates now
        ids = m.alloc(40)
        assert ids == om.alloc(40)
        mb.fence()
        for q in to_f:
            q.put(("data", ids, None))
        for q in from_f:
            ok, msg = q.get(timeout=120)
            assert ok, msg
        st = mb.stats()
        if not relay:
            assert st["device_direct"] == 1, "cudaHostRegister of the mailbox failed: kernel push not exercised"
            assert st["device_records"] > 0
        assert st["published"] == st["applied"]
        for q in to_f:
            q.put(("stop", None, None))
        result.put(("rank0", checks, st))
        m.close()
        mb.close()
        vmm_ops.shutdown_kvcached()
    except BaseException as e:   # surface in the parent
        import traceback
        result.put(("rank0-error", traceback.format_exc(), None))
        for q in to_f:
            q.put(("stop", None, None))
        raise


def _follower(name, rank, world, dev, big, inq, outq, result):
    try:
        _paths()
        from kvcached_b200 import kv_ops, tp
        cfg = _cfg(big)
        vmm_ops, ranges = _open_group(cfg, dev)
        mb = tp.TpMailbox(name, rank, world)
        n_checks = 0
        while True:
            kind, a, b = inq.get(timeout=300)
            if kind == "stop":
                break
            try:
                if kind == "check":
                    mb.fence()
                    vmm_ops.fence()
                    got = sorted(vmm_ops.mapped_pages())
                    assert got == b, f"rank {rank} {a}: mapped {got[:8]}.. ({len(got)}) != {b[:8]}.. ({len(b)})"
                    phys = vmm_ops.vmm_stats()["mapped_pages"]
                    assert len(b) <= phys <= len(b) + 5, f"rank {rank} {a}: physical footprint {phys} vs {len(b)}"
                    n_checks += 1
                else:                                   # data through the follower's own physical pages
                    ids = np.asarray(a, dtype=np.int64)
                    L, bs, cell = cfg["num_layers"], cfg["block_size"], cfg["cell_size"]
                    slots = torch.from_numpy((ids[:, None] * bs + np.arange(bs)).reshape(-1)).cuda()
                    k = torch.empty((L, slots.numel(), cell // 2), dtype=torch.int16, device="cuda").random_()
                    v = torch.empty_like(k).random_()
                    lay = kv_ops.layout_vllm_contiguous(ranges[0][0], L, bs, cell, 2)
                    kv_ops.kv_scatter(lay, slots, k, v)
                    k2, v2 = torch.zeros_like(k), torch.zeros_like(v)
                    kv_ops.kv_gather(lay, slots, k2, v2)
                    torch.cuda.synchronize()
                    assert torch.equal(k, k2) and torch.equal(v, v2) and int(k.abs().sum()) != 0
                outq.put((True, ""))
            except BaseException as e:
                outq.put((False, repr(e)))
                raise
        result.put((f"rank{rank}", n_checks, mb.stats()))
        mb.close()
        vmm_ops.shutdown_kvcached()
    except BaseException:
        import traceback
        result.put((f"rank{rank}-error", traceback.format_exc(), None))
        raise


def _run(world, big, relay, devs):
    ctx = mp.get_context("spawn")
    name = f"kvb200_tp_gpu_{os.getpid()}_{world}{int(big)}{int(relay)}"
    to_f = [ctx.Queue() for _ in range(world - 1)]
    from_f = [ctx.Queue() for _ in range(world - 1)]
    result = ctx.Queue()
    procs = [ctx.Process(target=_rank0, args=(name, world, devs[0], big, relay, to_f, from_f, result))]
    procs += [ctx.Process(target=_follower, args=(name, r, world, devs[r], big, to_f[r - 1], from_f[r - 1], result))
              for r in range(1, world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    for p in procs:
        if p.is_alive():
            p.kill()
    out = []
    while not result.empty():
        out.append(result.get())
    errs = [o for o in out if o[0].endswith("error")]
    assert not errs and all(p.exitcode == 0 for p in procs), (errs, [p.exitcode for p in procs])
    got = dict((o[0], o) for o in out)
    assert got["rank0"][1] >= 5
    for r in range(1, world):
        assert got[f"rank{r}"][1] == got["rank0"][1]          # every check ran on every rank
        assert got[f"rank{r}"][2]["applied"] == got["rank0"][2]["published"]
    return got


def test_kernel_pushes_page_lists_to_a_follower_process_same_gpu():
    got = _run(2, big=False, relay=False, devs=[0, 0])
    assert got["rank0"][2]["device_records"] > 10


def test_clear_of_more_than_1100_pages_world2():
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~20 GiB per rank")
    got = _run(2, big=True, relay=False, devs=[0, 0])
    assert got["rank0"][2]["relay_records"] >= 5               # the > 256-event lists went through the chunked relay


def test_host_relay_only_mode():
    _run(2, big=False, relay=True, devs=[0, 0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_three_ranks_on_separate_gpus():
    n = min(torch.cuda.device_count(), 3)
    _run(n, big=False, relay=False, devs=list(range(n)))


# ---- the collective path (NCCL) against the same expectations -----------------------------------------------------
def _nccl_worker(rank, world, port, result):
    try:
        _paths()
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import oracle
        from kvcached_b200 import replay as R
        from kvcached_b200 import tp
        from kvcached_b200 import trace as T
        cfg = _cfg(False)
        vmm_ops, _ = _open_group(cfg, rank)
        m = om = None
        if rank == 0:
            m = R.manager_from_config(cfg, device_index=0)
            m.set_phys_free_pages(1 << 40)
            om = oracle.manager_from_config(cfg)
        sync = tp.TpPageSync(m, 0, rank, world, torch.device("cuda", rank))
        trace = T.decode_trace(cfg, 60, seed=5)
        steps = T.split_steps(trace["ops"])
        res_o, res_m = [], []
        for s, idxs in enumerate(steps):
            if rank == 0:
                for i in idxs:
                    while len(res_o) < i:
                        res_o.append(None)
                        res_m.append(None)
                    res_o.append(oracle.apply_op(om, trace["ops"][i], res_o))
                    res_m.append(R.apply_op(m, trace["ops"][i], res_m))
            sync.sync()
            if s % 10 == 0 or s == len(steps) - 1:
                want = torch.full((cfg["num_blocks"] // 128 + 1,), -1, dtype=torch.int64, device="cuda")
                if rank == 0:
                    w = om.mapped_pages()
                    want[:len(w)] = torch.tensor(w, dtype=torch.int64)
                dist.broadcast(want, src=0)
                vmm_ops.fence()
                assert sorted(vmm_ops.mapped_pages()) == [p for p in want.tolist() if p >= 0], f"rank {rank} step {s}"
        # a list longer than one buffer: everything mapped, then clear()
        if rank == 0:
            n = cfg["num_blocks"] - 128
            assert m.alloc(n) == om.alloc(n)
        sync.sync()
        if rank == 0:
            m.clear(); om.clear()
        sync.sync()
        vmm_ops.fence()
        assert vmm_ops.mapped_pages() == []
        result.put((rank, sync.n_syncs, sync.n_events))
        dist.barrier()
        if m is not None:
            m.close()
        vmm_ops.shutdown_kvcached()
        dist.destroy_process_group()
    except BaseException:
        import traceback
        result.put((rank, "error", traceback.format_exc()))
        raise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_collective_path_nccl_two_ranks():
    ctx = mp.get_context("spawn")
    result = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, result)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    for p in procs:
        if p.is_alive():
            p.kill()
    out = sorted(result.get() for _ in range(2))
    assert all(p.exitcode == 0 for p in procs) and all(o[1] != "error" for o in out), out
    assert out[0][1] == out[1][1] and out[1][2] > 0
