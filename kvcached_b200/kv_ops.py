"""Fused paged KV scatter / gather through the mapped virtual range (kvb_kv_scatter / kvb_kv_gather).

The reference has no such kernel (the engines own the KV write); it defines the address layouts
(kvcached/integration/vllm/interfaces.py:88-314, integration/sglang/interfaces.py:66-175).  The four
builders below restate those layouts as (base, strides) so ONE kernel serves all of them
(SURVEY.md Appendix D.2).

Source / destination of the dense side: tensors of shape [num_layers, T, cell_bytes] viewed as bytes
(any dtype with the same byte layout, e.g. [L, T, H, D] bf16), K and V separately.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _lib
from ._lib import KvLayout


def layout_vllm_contiguous(base: int, num_layers: int, block_size: int, cell_bytes: int, num_kv: int = 2) -> KvLayout:
    """[num_blocks, L, nKV, block, H*D]: base + (((b*L + l)*nKV + kv)*block + t)*cell  (vllm/interfaces.py:291-298)"""
    blk = block_size * cell_bytes
    return KvLayout(base, num_layers * num_kv * blk, cell_bytes, num_kv * blk, blk, cell_bytes, block_size,
                    num_layers, num_kv, 0)


def layout_sglang_contiguous(base: int, num_layers: int, cell_bytes: int, num_kv: int = 2,
                             block_size: int = 1) -> KvLayout:
    """token-major [tokens, L, nKV, H*D]: base + ((s*L + l)*nKV + kv)*cell  (sglang/interfaces.py:168-173)"""
    tok = num_layers * num_kv * cell_bytes
    return KvLayout(base, block_size * tok, tok, num_kv * cell_bytes, cell_bytes, cell_bytes, block_size,
                    num_layers, num_kv, 0)


def layout_per_layer(base_layer0: int, layer_stride: int, per_layer_bytes: int, num_layers: int, block_size: int,
                     cell_bytes: int, num_kv: int = 2) -> KvLayout:
    """per-layer ranges at a constant stride: base_l + kv*(per_layer_bytes/2) + s*cell
    (vllm/interfaces.py:281-290, sglang/interfaces.py:156-166).  Ranges reserved back to back by
    kvb_create_kv_tensors have layer_stride == per_layer_bytes."""
    return KvLayout(base_layer0, block_size * cell_bytes, cell_bytes, layer_stride, per_layer_bytes // 2, cell_bytes,
                    block_size, num_layers, num_kv, 0)


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _stream(stream) -> Optional[int]:
    if stream is None:
        import torch
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def kv_scatter(lay: KvLayout, slots, k_src, v_src=None, layer_begin: int = 0, layer_end: Optional[int] = None,
               stream=None) -> None:
    """Write token K/V (device tensors [layers, T, cell]) to their slots (int64 device tensor [T])."""
    if layer_end is None:
        layer_end = lay.num_layers
    _lib.check(_lib.load().kvb_kv_scatter(C.byref(lay), slots.data_ptr(), slots.numel(), _ptr(k_src), _ptr(v_src),
                                          layer_begin, layer_end, _stream(stream)), "kv_scatter")


def kv_gather(lay: KvLayout, slots, k_dst, v_dst=None, layer_begin: int = 0, layer_end: Optional[int] = None,
              stream=None) -> None:
    if layer_end is None:
        layer_end = lay.num_layers
    _lib.check(_lib.load().kvb_kv_gather(C.byref(lay), slots.data_ptr(), slots.numel(), _ptr(k_dst), _ptr(v_dst),
                                         layer_begin, layer_end, _stream(stream)), "kv_gather")
