#!/usr/bin/env python
"""Fused KV scatter / gather GB/s for the TP cell sizes (Llama-3.3-70B shards): cell 2048 / 1024 / 512 / 256 bytes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcached_b200 import kv_ops, vmm_ops
from kvcached_b200.kv_cache_manager import KVCacheManager
PAGE = 2 << 20
torch.cuda.init()
for cell in (2048, 1024, 512, 256):
    L, bs, T = 80, 16, 65536
    blocks = T // bs + 64
    vmm_ops.init_kvcached("cuda:0", PAGE, True)
    per_layer = -(-blocks * bs * cell * 2 // (2 * PAGE)) * 2 * PAGE
    (rng_,) = vmm_ops.create_kv_tensors_raw(per_layer, L, 2)
    m = KVCacheManager(blocks, bs, cell, L)
    ids = m.alloc_np(T // bs)
    slots = torch.from_numpy(np.random.default_rng(0).permutation((ids[:, None] * bs + np.arange(bs)).reshape(-1))).cuda()
    k = torch.empty((L, T, cell // 2), dtype=torch.int16, device="cuda").random_()
    v = torch.empty_like(k).random_()
    lay = kv_ops.layout_vllm_contiguous(rng_[0], L, bs, cell, 2)
    alg = T * L * 4 * cell + 8 * T
    out = []
    for name, fn in (("scatter", kv_ops.kv_scatter), ("gather", kv_ops.kv_gather)):
        for _ in range(3):
            fn(lay, slots, k, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(lay, slots, k, v)
        e1.record(); torch.cuda.synchronize()
        out.append(f"{name} {alg / (e0.elapsed_time(e1) / 10) / 1e6:.0f} GB/s")
    print(f"cell {cell} B: " + ", ".join(out), flush=True)
    del k, v
    m.close(); vmm_ops.shutdown_kvcached()
    torch.cuda.empty_cache()
