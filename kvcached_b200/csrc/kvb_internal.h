// kvb_internal.h — structures shared by the sm_100a bookkeeping kernels and their host launcher.
// Not part of the ABI (include/kvb200.h is).
#pragma once

#include <cstdint>

#include "../../include/kvb200.h"

namespace kvb {

// Scalars of the device-resident manager state (int64 each).  They restate the scalar members of
// the reference's PageAllocator (csrc/inc/page_allocator.hpp:138-148) and KVCacheManager
// (kvcached/kv_cache_manager.py:171-179).
enum Sc : int {
  SC_NFREE = 0,     // num_free_pages_
  SC_NTOTAL,        // num_total_pages_
  SC_NAB,           // num_avail_blocks
  SC_INSHRINK,      // in_shrink
  SC_TARGET,        // target_num_blocks (-1 = None)
  SC_RB_HEAD,       // reserved_blocks ring: head index
  SC_NRB,           // len(reserved_blocks)
  SC_FL_HEAD,       // free_page_list_ ring
  SC_FL_SIZE,
  SC_RL_HEAD,       // reserved_page_list_ ring
  SC_RL_SIZE,
  SC_CL_HEAD,       // reclaimed_page_list_ ring
  SC_CL_SIZE,
  SC_AP_TAIL,       // avail_pages insertion-order stack: one past the last entry
  SC_NAVAILPG,      // len(avail_pages)
  SC_NFULLPG,       // len(full_pages)
  SC_NMAPPED,       // pages with a physical mapping (in use + reserved)
  SC_HAS_NULL,      // null block held
  SC_OUT_CURSOR,    // next free position of the id output arena (resident traces)
  SC_NEV,           // entries in the page event log
  SC_BAD_FREES,     // ids ignored by free(): not allocated / duplicate / out of range
  SC_UNKNOWN_FREES, // pages named by free() that are not in use (the reference logs and skips)
  SC_PHYS,          // physical page budget of the running launch
  SC_LAUNCHES,
  SC_LRU_HEAD,      // evictable LRU: first ring index that may be live
  SC_LRU_TAIL,      // one past the last entry
  SC_LRU_COUNT,     // len(_evictable_blocks)
  SC_COUNT = 32
};

enum : int { PG_NONE = 0, PG_AVAIL = 1, PG_FULL = 2 };

// Everything the kernels need to find the state; passed by value (constant bank).
struct DevState {
  int64_t P;            // page size in bytes
  int64_t B;            // block_mem_size = block_size * cell_size
  int32_t bpp;          // P / B  (InternalPage::get_num_blocks)
  int32_t exact;        // B divides P: page p holds exactly blocks [p*bpp, (p+1)*bpp)
  int64_t num_blocks;
  int32_t max_pages;    // pages of the virtual range created at init (capacity of every page array)
  int32_t q_mask;       // page deque capacity - 1 (power of two)
  int32_t rb_mask;      // reserved-block ring capacity - 1
  int32_t ap_cap;       // avail stack capacity
  int32_t min_reserved, max_reserved;
  int32_t ev_cap;
  int32_t n_primes;
  int64_t *sc;          // [SC_COUNT]
  int32_t *fl, *rl, *cl;        // page deques (ring buffers)
  int32_t *ap;                  // avail_pages in insertion order, -1 = tombstone
  int32_t *pg_state, *pg_cnt, *pg_head, *pg_pos, *pg_mapped;
  int32_t *pg_first, *pg_slot;  // scratch of free(): first index naming the page, its slot
  uint32_t *ring;               // [max_pages][bpp] per-page FIFO of free block offsets
  uint32_t *bitmap;             // bit per block: 1 = allocated
  int64_t *rb;                  // reserved_blocks ring
  int64_t *fids;                // scratch: ids of the running free()
  int32_t *fpage;               // scratch: their pages
  int32_t *g_dist, *g_ord, *g_cnt, *g_next, *g_bkt, *g_ptf;  // large-call fallbacks of the smem tables
  int32_t g_bkt_cap;
  int64_t f_cap;                // capacity of fids / fpage
  // prefix-cache pinning (ElasticBlockPool, integration/vllm/patches.py:308-614)
  int32_t *pin_ref;             // [num_blocks] KVCacheBlock.ref_cnt
  uint32_t *pin_keyed;          // bit per block: _block_id_to_key has the block
  int32_t *lru;                 // _evictable_blocks in insertion order, -1 = tombstone
  int32_t *lru_pos;             // [num_blocks] ring index or -1
  int32_t *pin_last;            // [num_blocks] scratch: last list index naming the block (-1 idle)
  uint32_t *tmp_bits;           // bit per block, all zero between requests (FREE_SLOTS: blocks named by the slots)
  int32_t *pin_in;              // [f_cap] scratch: the running PIN request's ids as int32 (-1 = not a block)
  int32_t lru_cap;
  int32_t max_cached;           // -1 unlimited, 0 evict on free, > 0 cap
  int64_t *ev;                  // page event log: page*2 + is_unmap
  const uint64_t *primes;       // libstdc++ __prime_list
  const uint16_t *grow_tab;     // [kGrowTab] bucket count after the rehash of a FULL table of nb buckets (nb < kGrowTab)
  long long *prof;              // -DKVB_PROFILE builds: [0..15] cycles per phase, [16..31] marks (else null)
};

constexpr int kGrowTab = 1024;

// What one launch publishes for the host (written last; pinned host or device memory).
struct Summary {
  int64_t seq;
  int64_t ops_done;       // requests completed by this launch
  int64_t out_used;       // ids written to the output buffer by this launch (step mode) / arena cursor
  int64_t n_events;       // events appended by this launch
  int64_t ev_begin;       // index of this launch's first event in the device log
  int64_t sc[SC_COUNT];
  int64_t ev_inline_n;
  int64_t ev_inline[256];
  int64_t next_free_n;    // how many entries of next_free are valid
  int64_t next_free[16];  // the pages alloc_page() will MAP next: the front of free_page_list (map-ahead hints)
  int64_t tp_published;   // 1 = this launch wrote its page list into the TP mailbox itself (kvb_tp.h)
};

struct LaunchArgs {
  const kvb_op *ops;      // request array
  int32_t op_begin, op_end;
  const int64_t *ids;     // FREE_IDS pool
  const kvb_ref *refs;    // FREE_REF pool
  kvb_res *res;           // results, indexed like ops
  int64_t *out_ids;       // id output buffer / arena
  int64_t out_cap;
  int64_t phys_free;      // get_avail_physical_pages() at launch time
  Summary *summary;
  int64_t seq;
  int32_t arena_mode;     // 0: output cursor and event log restart at 0; 1: both continue (resident traces);
                          // 2: cursor continues, event log restarts (relaunch of a partially done step)
  int32_t group_only;     // 1: requests are FREE_IDS probes for kvb_group_indices_by_page (state untouched)
  int64_t *grp_pages, *grp_counts, *grp_ids;   // outputs of group_only
  // SGLang paged allocator requests (KVB_OP_ALLOC_EXTEND / ALLOC_DECODE / FREE_SLOTS): device arrays
  const int64_t *slot_pre, *slot_seq, *slot_last;   // [slot_bs]; FREE_SLOTS: slot_pre = free_index[n]
  int64_t *slot_out;                                // token slot indices
  int64_t slot_tokens;                              // capacity of slot_out
  int32_t slot_bs, slot_ps;
  int64_t resize_target;  // per-layer size the shm limit asks for (-1: none), page_allocator.cpp:764-782
  kvb_res *host_res;      // zero-copy step: pinned host mirrors the kernel fills itself (else null)
  int64_t *host_out;
  // tensor-parallel mailbox (kvb_tp.h): device view of the shared ring, null = nothing to push
  int64_t *tp_ring;       // slot r % slots at tp_ring + (r & tp_mask) * tp_slot_words
  int64_t *tp_published;  // header word: records published so far
  int64_t tp_pub;         // record number this launch may write
  int32_t tp_mask, tp_slot_words, tp_flags, tp_pad;
  // single-request launches (kvb_mgr_alloc / kvb_mgr_free): the request — and up to kInlineIds ids of a free — ride in
  // the kernel parameters, so the kernel starts without a PCIe read of the host's request buffer
  int32_t inline_on;      // 0 = requests at `ops`; 1 = inline_op with ids in inline_ids; 2 = inline_op with ids at `ids`
  int32_t out_on_host;    // res / out_ids / summary ARE pinned host memory (single-request path): fence before seq
  kvb_op inline_op;
  int64_t inline_ids[256];
  // keep_events: the page event log continues from the previous launch (it published nothing: a free() whose summary
  // is left to the next launch) and this launch reports the log from its start.
  int32_t keep_events;
  int32_t safe_before;    // resident traces: results of requests below this index come from launches older than this
                          // launch's predecessor in the stream / graph (0 = unknown: nothing may be read early)
  int32_t tune;           // tuning switches (KVB200_TUNE): bit 0 = many-page frees replay the dict order on one warp,
                          // bit 1 = launches without programmatic stream serialization, bit 2 = no early ref gather,
                          // bit 3 = always the general kernel (no kvb_decode_kernel)
  // Flagged publish of a single-request launch: every 64-bit word of `ll` carries 32 bits of payload and the launch's
  // 32-bit flag, so the host validates each word on its own and the kernel needs NO system-scope fence between
  // payload and sequence word (8-byte stores are single-copy atomic over PCIe).  Layout: LlWord below.
  uint32_t ll_flag;
  uint64_t *ll;
};
constexpr int kInlineIds = 256;
// 32-bit payload slots of the flagged record (64-bit values take two: lo, hi)
enum LlWord : int {
  LL_STATUS = 0, LL_N, LL_AVAIL, LL_AUX = LL_AVAIL + 2, LL_NEV = LL_AUX + 2, LL_NHINT, LL_HINT,
  LL_SC = LL_HINT + 16, LL_EV = LL_SC + 2 * SC_COUNT, LL_IDS = LL_EV + 2 * 16
};
constexpr int kLlEvents = 16;    // page events carried inline (more: the host reads the device log)

#ifndef KVB_STEP_THREADS
#define KVB_STEP_THREADS 512
#endif
constexpr int kStepThreads = KVB_STEP_THREADS;   // 256, 384 or 512 (>= OPS_CHUNK, multiple of 32)

size_t step_kernel_smem_bytes();
// launches (host side, defined in kvb_kernels.cu)
// alloc_free_only: every request of the launch is ALLOC / FREE_IDS / FREE_REF / FREE_RESERVED / AVAIL -> kvb_decode_kernel
int launch_step(const DevState &st, const LaunchArgs &a, void *stream, bool alloc_free_only = false);
int launch_one(const DevState &st, const LaunchArgs &a, void *stream);   // single ALLOC / FREE_IDS / AVAIL request (inline_op)
#define KVB_ST_RETRY (-100)   /* kvb_one_kernel: the request needs the general kernel; nothing was touched */
int launch_step_prepare();   // per-device function attributes (must not happen inside a stream capture)
int launch_init_state(const DevState &st, int64_t num_total_pages, void *stream);
int launch_audit(const DevState &st, int64_t *d_out4, void *stream, bool zero_counters = true);
int launch_dump_avail(const DevState &st, int64_t *d_out, int64_t *d_n, void *stream);
int launch_kv_copy(const kvb_kv_layout &lay, const int64_t *slots, int64_t T, void *k, void *v, int32_t l0,
                   int32_t l1, int gather, void *stream);

}  // namespace kvb
