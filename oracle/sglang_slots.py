"""CPU restatement of the slot arithmetic behind ElasticPagedTokenToKVPoolAllocator.alloc_extend / alloc_decode /
free (kvcached/integration/sglang/patches.py:188-288).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the index arithmetic lives in SGLang (a dependency of the reference:
`get_num_new_pages` in sglang/srt/utils and the Triton kernels `alloc_extend_kernel` / `alloc_decode_kernel` in
sglang/srt/mem_cache/allocator.py), which is neither under /root/reference nor installed in this image.  What is
restated here is SGLang's published three-part rule, anchored on the reference's call sites (argument order and the
`if num_new_pages > 0` guard, patches.py:208-236, :249-272) and checked by properties in tests/test_sglang_slots_gpu.py
(every sequence's token t sits at slot block(t // page) * page + t % page of ITS OWN blocks).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def num_new_pages(seq_lens: np.ndarray, page: int, prefix_lens: Optional[np.ndarray] = None, decode: bool = False) -> int:
    if decode:
        prefix_lens = seq_lens - 1
    after = (seq_lens + page - 1) // page
    before = (prefix_lens + page - 1) // page
    return int((after - before).sum())


def extend_indices(prefix_lens, seq_lens, last_loc, new_pages: List[int], page: int) -> np.ndarray:
    """Slot indices of every new token, sequences back to back; `new_pages` are consumed in batch order."""
    out: List[int] = []
    k = 0
    for pre, seq, last in zip(prefix_lens.tolist(), seq_lens.tolist(), last_loc.tolist()):
        n_new = (seq + page - 1) // page - (pre + page - 1) // page
        mine = new_pages[k:k + n_new]
        k += n_new
        full_before = (pre + page - 1) // page * page
        part1 = min(seq, full_before) - pre                    # rest of the old partial page
        out.extend(last + 1 + t for t in range(part1))
        if pre + part1 == seq:
            continue
        part2 = seq // page * page - full_before               # whole new pages
        out.extend(mine[t // page] * page + t % page for t in range(part2))
        if pre + part1 + part2 == seq:
            continue
        part3 = seq - seq // page * page                       # new partial tail page
        out.extend(mine[n_new - 1] * page + t for t in range(part3))
    return np.asarray(out, dtype=np.int64)


def decode_indices(seq_lens, last_loc, new_pages: List[int], page: int) -> np.ndarray:
    out = []
    k = 0
    for seq, last in zip(seq_lens.tolist(), last_loc.tolist()):
        if (seq + page - 1) // page == (seq - 1 + page - 1) // page:
            out.append(last + 1)
        else:
            out.append(new_pages[k] * page)
            k += 1
    return np.asarray(out, dtype=np.int64)


def freed_blocks(free_index: np.ndarray, page: int) -> List[int]:
    """torch.unique(free_index // page_size): sorted ascending (patches.py:279)."""
    return np.unique(free_index // page).tolist()
