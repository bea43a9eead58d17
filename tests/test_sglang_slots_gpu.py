"""alloc_extend / alloc_decode / free of the SGLang paged allocator as ONE device launch each, against the oracle
block manager + the restated SGLang slot arithmetic (oracle/sglang_slots.py), and against the property that makes the
result usable at all: token t of a sequence lives at block(t // page) * page + t % page of that sequence's own blocks."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
import sglang_slots as SS  # noqa: E402

from kvcached_b200.kv_cache_manager import KVCacheManager  # noqa: E402
from kvcached_b200.sglang_alloc import ElasticPagedTokenToKVPoolAllocator  # noqa: E402

pytestmark = pytest.mark.gpu
P = 2 << 20


def _snap_equal(m, o):
    so, st = o.scalars(), m.stats()
    assert (st.num_free_pages, st.num_avail_blocks, st.num_reserved_pages) == (so["free_pages"], so["num_avail_blocks"], so["reserved_pages"])
    assert m.available_size() == o.available_size()
    assert m.avail_pages() == o.avail_pages()


@pytest.mark.parametrize("page,cell,blocks,seed", [(16, 2048, 4096, 0), (64, 1152, 700, 1), (4, 512, 20000, 2), (16, 2048, 96, 3)])
def test_extend_decode_free_match_oracle(page, cell, blocks, seed):
    import torch
    L, nkv = 4, (1 if cell == 1152 else 2)
    m = KVCacheManager(blocks, page, cell, L, num_kv_buffers=nkv, page_size=P, attach_vmm=False)
    m.set_phys_free_pages(1 << 40)
    o = oracle.manager_from_config(dict(num_blocks=blocks, block_size=page, cell_size=cell, num_layers=L, num_kv_buffers=nkv,
                                        page_size=P, reserve_null_block=False))
    o.set_phys_free_pages(1 << 40)
    al = ElasticPagedTokenToKVPoolAllocator(blocks * page, page, torch.bfloat16, "cuda:0", m)
    rng = np.random.default_rng(seed)
    seqs = {}            # id -> list of slots (token order)
    nxt = 0
    dev = lambda a: torch.from_numpy(np.asarray(a, dtype=np.int64)).cuda()   # noqa: E731
    refused = 0
    try:
        for rnd in range(60):
            kind = rng.choice(["extend", "decode", "free"], p=[0.4, 0.4, 0.2]) if seqs else "extend"
            if kind == "extend":
                ids = []
                for _ in range(int(rng.integers(1, 9))):
                    if seqs and rng.random() < 0.5:
                        c = int(rng.choice(list(seqs)))
                        if c not in ids:
                            ids.append(c)
                    else:
                        seqs[nxt] = []
                        ids.append(nxt)
                        nxt += 1
                pre = np.array([len(seqs[i]) for i in ids])
                ext = rng.integers(1, 5 * page, size=len(ids))
                seq = pre + ext
                last = np.array([seqs[i][-1] if seqs[i] else -1 for i in ids])
                n_new = SS.num_new_pages(seq, page, pre)
                want_ids = o.alloc(n_new) if n_new > 0 else []
                got = al.alloc_extend(dev(pre), None, dev(seq), None, dev(last), int(ext.sum()))
                if want_ids is None:
                    assert got is None
                    refused += 1
                    for i in ids:
                        if not seqs[i]:
                            del seqs[i]
                    _snap_equal(m, o)
                    continue
                got = got.cpu().numpy()
                assert np.array_equal(got, SS.extend_indices(pre, seq, last, want_ids, page))
                k = 0
                for i, e in zip(ids, ext.tolist()):
                    seqs[i].extend(got[k:k + e].tolist())
                    k += e
            elif kind == "decode":
                ids = list(rng.choice(list(seqs), size=min(len(seqs), int(rng.integers(1, 12))), replace=False))
                ids = [int(i) for i in ids if seqs[int(i)]]
                if not ids:
                    continue
                seq = np.array([len(seqs[i]) + 1 for i in ids])
                last = np.array([seqs[i][-1] for i in ids])
                n_new = SS.num_new_pages(seq, page, decode=True)
                want_ids = o.alloc(n_new) if n_new > 0 else []
                got = al.alloc_decode(dev(seq), None, dev(last))
                if want_ids is None:
                    assert got is None
                    refused += 1
                    _snap_equal(m, o)
                    continue
                got = got.cpu().numpy()
                assert np.array_equal(got, SS.decode_indices(seq, last, want_ids, page))
                for i, sl in zip(ids, got.tolist()):
                    seqs[i].append(sl)
            else:
                ids = [int(i) for i in rng.choice(list(seqs), size=min(len(seqs), int(rng.integers(1, 4))), replace=False)]
                slots = np.concatenate([np.asarray(seqs[i], dtype=np.int64) for i in ids]) if any(seqs[i] for i in ids) else np.zeros(0, np.int64)
                for i in ids:
                    del seqs[i]
                if slots.size == 0:
                    continue
                rng.shuffle(slots)
                o.free(SS.freed_blocks(slots, page))
                al.free(dev(slots))
            _snap_equal(m, o)
            # the property: every sequence's tokens are laid out page by page inside its own blocks
            owner = {}
            for i, sl in seqs.items():
                a = np.asarray(sl, dtype=np.int64)
                if a.size == 0:
                    continue
                assert np.array_equal(a % page, np.arange(a.size) % page)
                blk = a // page
                assert np.all(blk[: a.size // page * page].reshape(-1, page) == blk[: a.size // page * page].reshape(-1, page)[:, :1])
                for b in np.unique(blk).tolist():
                    assert owner.setdefault(b, i) == i, "two sequences share a block"
        if blocks <= 128:
            assert refused > 0
        assert m.audit()[3] == 0
    finally:
        m.close()


def test_token_level_allocator_keeps_ids_on_the_device():
    """ElasticTokenToKVPoolAllocator (page_size 1): alloc -> device tensor, free(device tensor) in the tensor's order —
    ids, free-list order and every later allocation equal the oracle fed with the same lists."""
    import torch
    from kvcached_b200.sglang_alloc import ElasticTokenToKVPoolAllocator
    blocks, cell, L = 30000, 1152, 4                       # MLA token slots: 1820 per page, straddling page edges
    m = KVCacheManager(blocks, 1, cell, L, num_kv_buffers=1, page_size=P, attach_vmm=False)
    m.set_phys_free_pages(1 << 40)
    o = oracle.manager_from_config(dict(num_blocks=blocks, block_size=1, cell_size=cell, num_layers=L, num_kv_buffers=1,
                                        page_size=P, reserve_null_block=False))
    o.set_phys_free_pages(1 << 40)
    al = ElasticTokenToKVPoolAllocator(blocks, torch.bfloat16, "cuda:0", m)
    rng = np.random.default_rng(9)
    held = []
    try:
        for rnd in range(40):
            if held and rng.random() < 0.45:
                i = int(rng.integers(0, len(held)))
                t = held.pop(i)
                perm = torch.from_numpy(rng.permutation(t.numel())).cuda()
                t = t[perm]                                 # SGLang frees in whatever order its request pool hands back
                o.free(t.cpu().tolist())
                al.free(t)
            else:
                n = int(rng.integers(1, 4000))
                want = o.alloc(n)
                got = al.alloc(n)
                if want is None:
                    assert got is None
                else:
                    assert got.device.type == "cuda" and got.cpu().tolist() == want
                    held.append(got)
            _snap_equal(m, o)
        al.free_group_begin()
        for t in held[:2]:
            al.free(t)
        assert m.available_size() == o.available_size()     # nothing freed yet
        al.free_group_end()
        if held[:2]:
            o.free(torch.cat(held[:2]).cpu().tolist())
        _snap_equal(m, o)
        assert m.audit()[3] == 0
    finally:
        m.close()
