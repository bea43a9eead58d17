"""Rank-consistent page mapping under tensor parallelism: ONE small NCCL broadcast per scheduler step.

Replaces kvcached/tp_ipc_util.py (per-rank AF_UNIX listener threads, pickle, one connection per rank
per operation, :96-267).  Rank 0 owns the block manager; after every step the page map/unmap list its
kernel produced is packed as [seq, n, e0..e(n-1)] into a fixed int64 device buffer
(kvb_mgr_tp_export), broadcast to the TP group over NVLink (torch.distributed, backend nccl; gloo on
CPU for the tests), and every other rank runs the same batched cuMemMap loop on its own head shard
(kvb_tp_apply).  Block ids are identical on all ranks by construction — only pages travel.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _lib


class TpPageSync:
    def __init__(self, mgr, group_id: int, rank: int, world: int, device, process_group=None, src: int = 0):
        import torch
        self.mgr = mgr
        self.group_id = group_id
        self.rank = rank
        self.world = world
        self.src = src
        self.pg = process_group
        self.lib = _lib.load()
        self.words = int(self.lib.kvb_tp_buffer_words())
        self.buf = torch.zeros(self.words, dtype=torch.int64, device=device)
        self.last_seq = -1
        self.n_syncs = 0
        self.n_events = 0
        self._is_cuda = self.buf.is_cuda

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream) if self._is_cuda else None

    def sync(self) -> int:
        """Called by every rank once per scheduler step, after rank `src` ran the step.  Returns the
        number of page events applied on this rank."""
        import torch.distributed as dist
        if self.rank == self.src:
            _lib.check(self.lib.kvb_mgr_tp_export(self.mgr._h, self.buf.data_ptr(), self._stream()), "tp_export")
        dist.broadcast(self.buf, src=self.src, group=self.pg)
        self.n_syncs += 1
        if self.rank == self.src:
            return 0          # rank src applied its own events inside kvb_mgr_step
        seq, n = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.kvb_tp_apply(self.group_id, self.buf.data_ptr(), self._stream(), C.byref(seq),
                                         C.byref(n)), "tp_apply")
        self.last_seq = seq.value
        self.n_events += n.value
        return n.value


def pack_events(seq: int, events, words: int):
    """Host-side packing with the same wire format (used by the gloo/CPU tests)."""
    import torch
    buf = torch.zeros(words, dtype=torch.int64)
    buf[0] = seq
    buf[1] = len(events)
    for i, (page, unmap) in enumerate(events):
        buf[2 + i] = page * 2 + unmap
    return buf


def unpack_events(buf):
    n = int(buf[1])
    return int(buf[0]), [(int(e) >> 1, int(e) & 1) for e in buf[2:2 + n].tolist()]


def replay_host_steps_tp(tp: TpPageSync, mgr, plan, rank: int) -> None:
    """replay.replay_host_steps under TP: rank 0 runs every scheduler step, all ranks sync pages after it."""
    from .replay import replay_host_steps
    if rank == tp.src:
        replay_host_steps(mgr, plan, keep_results=False, after_step=tp.sync)
    else:
        for _ in plan.steps:
            tp.sync()
