"""BASELINE configs[2]: two models (Llama-3.1-8B and Qwen2.5-7B KV geometries) co-located on one B200, each with its own
manager, virtual KV tensors (group 0 / group 1) and usage segment, coupled only through free physical memory and kvctl
limits — as in the reference (examples/01_simple_two_models, 02_memory_control: two engine processes, live
cudaMemGetInfo in available_size(), `kvctl limit`).  The free-memory series is made deterministic the way SURVEY §8(d)
T-two-model says: a shared physical budget, from which each manager's phys-free page count is derived after every step
and fed to BOTH the device managers and the two oracles.  Checked after every scheduler step: every alloc result,
available_size, page counters, the physically mapped page set of each group, the usage segments, and that data
written by one model survives the other model's map / unmap traffic."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402

from kvcached_b200 import kv_cache_manager as K  # noqa: E402
from kvcached_b200 import trace as T  # noqa: E402
from kvcached_b200 import vmm_ops  # noqa: E402

pytestmark = pytest.mark.gpu
P = 2 << 20


def test_two_colocated_models_share_one_gpu():
    import torch
    torch.cuda.init()
    vmm_ops.init_kvcached("cuda:0", P, True)
    L = 4                                                     # layers scaled down; cell sizes are the real ones
    cfgs = [T.make_config("llama8b", 64 * 40, num_layers=L), T.make_config("qwen7b", 128 * 48, num_layers=L)]
    names = [f"kvb200_two_{os.getpid()}_a", f"kvb200_two_{os.getpid()}_b"]
    mgrs, orcs, flats, per_page = [], [], [], []
    try:
        for g, cfg in enumerate(cfgs):
            per_layer = cfg["num_blocks"] * cfg["block_size"] * cfg["cell_size"] * 2
            (flat,) = vmm_ops.create_kv_tensors(per_layer, 2, "cuda:0", L, 2, g)
            flats.append(flat.view(torch.int16))
            mgrs.append(K.KVCacheManager(cfg["num_blocks"], cfg["block_size"], cfg["cell_size"], L, group_id=g,
                                         page_size=P, ipc_name=names[g]))
            orcs.append(oracle.manager_from_config(cfg))
            per_page.append(L * P * 2)
        traces = [T.decode_trace(cfgs[0], 160, seed=11, target_peak=0.7, out_mu=4.0, prompt_mu=5.5),
                  T.decode_trace(cfgs[1], 200, seed=12, target_peak=0.7, out_mu=4.2, prompt_mu=5.2)]
        steps = [T.split_steps(t["ops"]) for t in traces]
        # anti-phase: model B starts when model A is near its peak
        shift = len(steps[0]) // 2
        n_steps = max(len(steps[0]), shift + len(steps[1]))
        budget = 26 * per_page[0]                             # physical bytes both models may map together (< A's + B's peak)
        results = [[None] * len(t["ops"]) for t in traces]
        want = [[None] * len(t["ops"]) for t in traces]

        def set_budgets():
            mapped = [len(o.mapped_pages()) * per_page[g] for g, o in enumerate(orcs)]
            for g in range(2):
                free_pages = max(budget - sum(mapped), 0) // per_page[g]
                mgrs[g].set_phys_free_pages(free_pages)
                orcs[g].set_phys_free_pages(free_pages)

        def kv_view(g):
            cfg = cfgs[g]
            blk_elems = cfg["block_size"] * cfg["cell_size"] // 2
            return flats[g].view(-1)[: cfg["num_blocks"] * L * 2 * blk_elems].view(cfg["num_blocks"], L, 2, blk_elems)

        held = [{}, {}]                                       # per model: alloc op index -> (first block, value written)
        none_seen = 0
        for s in range(n_steps):
            set_budgets()
            for g in range(2):
                k = s - (shift if g else 0)
                if k < 0 or k >= len(steps[g]):
                    continue
                for i in steps[g][k]:
                    op = traces[g]["ops"][i]
                    results[g][i] = oracle.apply_op(mgrs[g], op, results[g])
                    want[g][i] = oracle.apply_op(orcs[g], op, want[g])
                    assert results[g][i] == want[g][i], f"model {g} step {k} op {i} {op[:2]}"
                    if op[0] == "alloc":
                        if results[g][i] is None:
                            none_seen += 1
                        elif results[g][i]:
                            b, val = results[g][i][0], (1000 * g + i) % 30000
                            kv_view(g)[b, :, :, 0] = val      # a value only this allocation writes
                            held[g][i] = (b, val)
                    elif op[0] == "free_cat":
                        for q in op[1]:
                            held[g].pop(q, None)
                    elif op[0].startswith("free"):
                        held[g].clear()
                    set_budgets()
            if s == n_steps // 3:                             # kvctl limit on model A: half of its pool
                assert K.shm_set_limit(names[0], cfgs[0]["num_blocks"] * 16 * 2048 * L * 2 // 2)
                orcs[0].set_resize_target(cfgs[0]["num_blocks"] * 16 * 2048 // 2)
            for g in range(2):
                so, st = orcs[g].scalars(), mgrs[g].stats()
                assert (st.num_free_pages, st.num_total_pages, st.num_reserved_pages, st.in_shrink) == \
                       (so["free_pages"], so["total_pages"], so["reserved_pages"], so["in_shrink"]), f"model {g} step {s}"
                assert mgrs[g].available_size() == orcs[g].available_size()
                assert vmm_ops.mapped_pages(g) == orcs[g].mapped_pages() == mgrs[g].mapped_pages()
                t, used, pre = K.shm_read(names[g])
                assert used == st.num_inuse_pages * per_page[g] and pre == st.num_reserved_pages * per_page[g]
            if s % 16 == 0:                                   # data survives the other model's map / unmap traffic
                torch.cuda.synchronize()
                for g in range(2):
                    kv = kv_view(g)
                    for i, (b, val) in list(held[g].items())[:64]:
                        assert int(kv[b, 0, 0, 0]) == val and int(kv[b, L - 1, 1, 0]) == val, f"model {g} block {b} lost its data"
        assert none_seen > 0, "the shared budget must refuse some allocations, else the coupling is not exercised"
        for g in range(2):
            assert mgrs[g].audit()[3] == 0
    finally:
        for m in mgrs:
            m.close()
        vmm_ops.shutdown_kvcached()
        for n in names:
            assert not os.path.exists("/dev/shm/" + n)
