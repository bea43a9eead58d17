"""Rank-consistent page mapping under tensor parallelism.

Replaces kvcached/tp_ipc_util.py (per-rank AF_UNIX listener threads, pickle, one connection per rank per
operation, :96-267).  Rank 0 owns the block manager; block ids are identical on all ranks by construction —
only page lists travel.  Two transports over the same C ABI:

  * TpMailbox (default): rank 0's step KERNEL writes each step's page list into a shared-memory ring that is
    mapped pinned host memory for the GPU and plain memory for every rank's follower thread (kvb_tp_init,
    csrc/kvb_tp.{h,cpp}).  No collective, no Python on the followers' path, steps without page events cost
    nothing; kvb_mgr_step on rank 0 returns once a step's maps are in place on every rank.
  * TpPageSync: the page list as one fixed-size int64 buffer per chunk, broadcast with torch.distributed
    (NCCL over NVLink; gloo on CPU) — the checker of the mailbox, and the transport for groups spanning nodes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _lib


class TpMailbox:
    """kvb_tp_init / kvb_tp_fence / kvb_tp_stats / kvb_tp_shutdown for one KV-cache group.

    rank 0 passes its KVCacheManager; other ranks pass mgr=None and must have created their KV tensors
    (init_kvcached + create_kv_tensors) first — their follower thread maps into them."""

    def __init__(self, name: str, rank: int, world: int, mgr=None, group_id: int = 0, host_relay: bool = False,
                 external_producer: bool = False):
        self.lib = _lib.load()
        self.name, self.rank, self.world, self.group_id = name, rank, world, group_id
        h = None
        if rank == 0 and not external_producer:
            if mgr is None:
                raise ValueError("rank 0 needs its KVCacheManager")
            mgr._wait_post_init()
            h = mgr._h
        elif mgr is not None:
            raise ValueError("only rank 0 drives a KVCacheManager")
        self._mgr = mgr
        flags = (1 if host_relay else 0) | (2 if (external_producer and rank == 0) else 0)
        _lib.check(self.lib.kvb_tp_init(name.encode(), rank, world, group_id, h, flags), "kvb_tp_init")
        self._open = True

    def fence(self) -> None:
        """rank 0: every rank applied everything published; other ranks: this rank did."""
        _lib.check(self.lib.kvb_tp_fence(self.group_id), "kvb_tp_fence")

    def publish(self, events, hints=(), seq: int = 0, settle: bool = False) -> None:
        """Producer side for page lists computed elsewhere: events = [(page, is_unmap), ...]."""
        ev = np.array([p * 2 + u for p, u in events], dtype=np.int64)
        hi = np.array(list(hints), dtype=np.int64)
        _lib.check(self.lib.kvb_tp_publish(self.group_id, seq, ev.ctypes.data_as(_lib.p_i64) if ev.size else None, ev.size,
                                           hi.ctypes.data_as(_lib.p_i64) if hi.size else None, hi.size,
                                           1 if settle else 0), "kvb_tp_publish")

    def stats(self) -> dict:
        o = np.zeros(8, dtype=np.int64)
        _lib.check(self.lib.kvb_tp_stats(self.group_id, o.ctypes.data_as(_lib.p_i64)), "kvb_tp_stats")
        keys = ("published", "applied", "events_applied", "device_direct", "device_records", "relay_records",
                "ack_waits", "ack_wait_us")
        return {k: int(v) for k, v in zip(keys, o)}

    def close(self) -> None:
        if self._open:
            if self._mgr is not None and getattr(self._mgr, "_h", None):
                self.lib.kvb_mgr_tp_detach(self._mgr._h)
            self.lib.kvb_tp_shutdown(self.group_id)
            self._open = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mailbox_name(tag: Optional[str] = None) -> str:
    """A name all ranks of one job agree on without talking: the rendezvous port torchrun gave them."""
    return tag or f"kvb200_tp_{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'job')}"


class TpPageSync:
    """The collective transport: one torch.distributed broadcast per chunk of rank 0's pending page list."""

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

    def _ptr(self):
        return C.cast(self.buf.data_ptr(), _lib.p_i64)

    def sync(self) -> int:
        """Called by every rank once per scheduler step, after rank `src` ran the step.  Returns the number of
        page events applied on this rank.  A list longer than one buffer travels in chunks: every rank loops until
        the buffer's `more` word is 0, so no rank can leave (or raise) ahead of a collective the others are in."""
        import torch.distributed as dist
        applied = 0
        while True:
            err = None
            if self.rank == self.src:
                rc = (self.lib.kvb_mgr_tp_export(self.mgr._h, self.buf.data_ptr(), self._stream()) if self._is_cuda
                      else self.lib.kvb_mgr_tp_export_host(self.mgr._h, self._ptr()))
                if rc < 0:            # never before the collective: ship an empty, final buffer and raise after it
                    err = _lib.last_error()
                    self.buf.zero_()
            dist.broadcast(self.buf, src=self.src, group=self.pg)
            self.n_syncs += 1
            if err is not None:
                raise _lib.KvbError(f"tp_export failed: {err}")
            if self.rank == self.src:
                more = rc         # rank src applied its own events inside kvb_mgr_step
            else:
                seq, n = C.c_int64(0), C.c_int64(0)
                more = (self.lib.kvb_tp_apply(self.group_id, self.buf.data_ptr(), self._stream(), C.byref(seq), C.byref(n))
                        if self._is_cuda else self.lib.kvb_tp_apply_host(self.group_id, self._ptr(), C.byref(seq), C.byref(n)))
                _lib.check(more, "tp_apply")
                self.last_seq = seq.value
                self.n_events += n.value
                applied += n.value
            if not more:
                return applied


def pack_events(seq: int, events, words: Optional[int] = None, hints=()):
    """All chunks of one page list in the collective path's wire format (kvb_tp_pack): a list of int64 tensors."""
    import torch
    lib = _lib.load()
    words = int(lib.kvb_tp_buffer_words())
    ev = np.array([p * 2 + u for p, u in events], dtype=np.int64)
    hi = np.array(list(hints), dtype=np.int64)
    cur = C.c_int64(0)
    out = []
    while True:
        buf = np.zeros(words, dtype=np.int64)
        more = _lib.check(lib.kvb_tp_pack(seq, ev.ctypes.data_as(_lib.p_i64) if ev.size else None, ev.size, C.byref(cur),
                                          hi.ctypes.data_as(_lib.p_i64) if hi.size else None, hi.size,
                                          buf.ctypes.data_as(_lib.p_i64)), "tp_pack")
        out.append(torch.from_numpy(buf))
        if not more:
            return out


def unpack_events(buf):
    n = int(buf[1])
    return int(buf[0]), [(int(e) >> 1, int(e) & 1) for e in buf[2:2 + n].tolist()]


def replay_host_steps_tp(tp, mgr, plan, rank: int) -> None:
    """replay.replay_host_steps under TP.  Mailbox: rank 0 replays, the kernel pushes, followers' threads apply —
    no per-step call on any other rank.  Collective: all ranks sync pages after every step."""
    from .replay import replay_host_steps
    if isinstance(tp, TpMailbox):
        if rank == 0:
            replay_host_steps(mgr, plan, keep_results=False)
        return
    if rank == tp.src:
        replay_host_steps(mgr, plan, keep_results=False, after_step=tp.sync)
    else:
        for _ in plan.steps:
            tp.sync()
