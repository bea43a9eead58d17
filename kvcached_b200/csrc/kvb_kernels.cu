// kvb_kernels.cu — hand-written sm_100a kernels of the elastic KV-cache hot path.
//
//   kvb_step_kernel   one CTA replays a batch of alloc / free / resize / ... requests against the
//                     device-resident block<->page state with exactly the results N sequential calls
//                     of the reference KVCacheManager would give (kvcached/kv_cache_manager.py:246-506,
//                     csrc/page_allocator.cpp:40-100, :161-460, :471-498).
//   kvb_audit_kernel  grid-wide coalesced uint4 sweep of the used-block bitmap (+popc) that recounts
//                     every page and flags fully-free pages.
//   kvb_kv_copy_kernel fused paged KV scatter / gather through the mapped virtual range, 128-bit
//                     accesses, all layers and K/V in one launch.
//
// State design (DevState, kvb_internal.h).  The reference keeps, on the host, a std::vector FIFO per
// page, two Python dicts (avail_pages in insertion order, full_pages) and three std::deque page lists.
// Here:  per page a ring of free block offsets + count + head (the FIFO), a used-block bitmap (1 bit
// per block; membership, double-free rejection, popc audits), page state words, the avail_pages
// insertion order as an append-only stack with tombstones (popitem() = last live entry), and the
// three page lists as ring buffers.  Order, not just set membership, is what makes results bit-exact
// (SURVEY.md Appendix A), so every structure that decides an order is carried explicitly.
//
// Why one CTA: requests of a step are sequentially dependent by definition (request i sees the state
// request i-1 left).  Parallelism is taken INSIDE a request or a run of requests:
//   * a run of consecutive alloc() requests consumes one deterministic stream of free blocks
//     (reserved blocks, then avail pages newest-first, then fresh pages); thread 0 walks pages and
//     cuts the stream into segments, all warps then emit ids / bitmap bits / ring inits in parallel;
//   * free(ids): all threads classify ids (page arithmetic, bitmap test-and-clear), find the distinct
//     pages in first-appearance order (atomicMin + __ballot_sync/__popc compaction), one lane replays
//     libstdc++'s unordered_map insertion policy to get the reference's dict order while a second warp
//     computes per-page FIFO ranks with __match_any_sync, then all threads apply per-page updates
//     (empty-page detection, avail stack re-insertion) with ballot/popc prefix sums.
#include <cuda_runtime.h>

#include <climits>
#include <cstddef>
#include <cstdlib>
#include <cstdint>

#include "kvb_internal.h"

namespace kvb {

namespace {

// Phase timing of the step kernel (tuning aid): build with -DKVB_PROFILE, read with kvb_mgr_profile().
// Thread 0 charges the cycles since the previous mark to `slot`.  Compiled out of the product build.
#ifdef KVB_PROFILE
#define KVB_PROF(slot)                                                                 \
  do {                                                                                 \
    if (threadIdx.x == 0) {                                                            \
      long long t_ = clock64();                                                        \
      c.prof[slot] += t_ - s.prof_t;                                                   \
      c.prof[16 + (slot)] += 1;                                                        \
      s.prof_t = t_;                                                                   \
    }                                                                                  \
  } while (0)
#else
#define KVB_PROF(slot) do { } while (0)
#endif

constexpr int NT = kStepThreads;
constexpr int NW = NT / 32;
constexpr int OPS_CHUNK = 256;   // requests staged in shared memory at a time
constexpr int SEG_CAP = 1024;    // stream segments per generation round
constexpr int FRESH_CAP = 512;   // fresh pages per generation round
constexpr int APW = 256;         // prefetched window of the avail stack
constexpr int NEWPG_W = 64;      // prefetched window of upcoming fresh page ids
constexpr int HT_D_CAP = 2048;   // distinct pages per free() handled in shared memory
constexpr int HT_NB_CAP = 4099;  // hash buckets handled in shared memory
constexpr int EV_INLINE = 256;
// call sites of the one-barrier scans: each owns its row of Smem::wsite
enum : int { WS_GATHER = 0, WS_FS_D, WS_FS_F4, WS_AF0, WS_AF1, WS_AF2, WS_AF3, WS_COUNT };
constexpr int EARLY_K = 4;        // leading free() requests of a resident step whose ids are gathered before the wait
constexpr int PRIMES_SM = 128;   // __prime_list[0..127] reaches 2^17; larger bucket counts read the global table

enum : int { SEG_RB = 0, SEG_RING = 1, SEG_FRESH = 2 };

struct Seg {
  int32_t out_rel;   // offset of the segment's first id, relative to the round's output base
  int32_t page;      // page id (RING/FRESH)
  int32_t base;      // RB: ring index into rb; RING: ring position; FRESH: block offset inside the page
  int32_t len_kind;  // len | kind << 28
};

// Resumable state of the alloc-run walk (thread 0 only).
struct AllocSim {
  int32_t op;          // next request of the run
  int32_t req_active;  // a request is half done (round ran out of segment slots)
  int64_t req_rem;     // blocks the active request still needs
  int64_t req_off;     // output offset of the active request's first id
  int64_t req_n;
  int32_t ap_top;      // avail stack: one past the entry under inspection
  int32_t ap_tail0;    // stack tail when the run started (window origin)
  int32_t cur_p;       // page being consumed (-1 none)
  int32_t cur_cnt;     // free blocks left in it
  int32_t cur_head;    // ring position / fresh offset of its next free block
  int32_t cur_fresh;   // 1: page taken in this run, ring not materialised yet
  int32_t cur_idx;     // its stack index (-1: not on the stack yet)
  int32_t newpg_j;     // fresh pages taken so far in this run
  int32_t rl_size0;    // reserved_page_list size when the run started
  int32_t nseg, nfresh;
  int32_t done;
  int32_t failed;      // request hit "No free pages left"
  int64_t out_pos;     // stream position (absolute in the output buffer)
  int64_t round_base;  // output position where this round started
};

struct WarpPlan {
  int32_t T, take_rb, take_av, rem2, nlive, F, from_rl, last_used, tail, pad;
  int64_t cursor, A0, nev0, rbh;
};

struct Smem {
  int64_t sc[SC_COUNT];
  int64_t acc[8];
  kvb_op ops[OPS_CHUNK];
  Seg segs[SEG_CAP];
  int2 fresh[FRESH_CAP];
  int32_t apw[APW], apcnt[APW], aphead[APW];
  int32_t newpg[NEWPG_W];
  int32_t cpage[APW], ccnt[APW], chead[APW], csidx[APW], ccum[APW];   // live window entries, newest first
  uint32_t primes32[PRIMES_SM];                                        // head of libstdc++'s __prime_list
  uint16_t grow16[kGrowTab];                                           // DevState::grow_tab; MUST follow primes32 (next_bkt_grow)
  int32_t ht_key[HT_D_CAP], ht_next[HT_D_CAP], ht_ord[HT_D_CAP], ht_cnt[HT_D_CAP];
  int32_t ht_bkt[HT_NB_CAP + 1];
  int32_t wtot[NW], wtot2[NW];
  int32_t wsite[WS_COUNT][NW];   // per-call-site buffers of the one-barrier scans (block_*_1b)
  int32_t misc[16];
  int64_t tmp64[NT];
  int32_t early_ids[EARLY_K][NT];   // ids of the launch's leading FREE_REF requests, gathered before pdl_enter()
  int32_t early_n[EARLY_K];         // their id counts, -1 = gather it after the wait
  uint32_t namask[OPS_CHUNK / 32];  // staged chunk: bit = this request is not an ALLOC (or lies past the chunk's end)
  WarpPlan wplan;                   // alloc_run_warp: warp 0's plan for the CTA
  int32_t wroff[128];               // ... and the id-stream offset of each request of the run
  AllocSim sim;
  long long prof_t;
};

static_assert(offsetof(Smem, grow16) == offsetof(Smem, primes32) + PRIMES_SM * sizeof(uint32_t), "grow16 follows primes32");

enum : int { MI_D = 0, MI_NB0, MI_E, MI_USE_G, MI_STOP, MI_BASE_A, MI_BASE_B, MI_N, MI_BAD, MI_UNK, MI_K };

// ------------------------------------------------------------------------------------ page arithmetic
// InternalPage::get_block_range / get_num_blocks, csrc/page_allocator.cpp:87-100
__device__ __forceinline__ int64_t pg_start(const DevState &c, int p) {
  return c.exact ? (int64_t)p * c.bpp : ((int64_t)p * c.P + c.B - 1) / c.B;
}
__device__ __forceinline__ int pg_nblk(const DevState &c, int p) {
  if (c.exact) return c.bpp;
  int64_t s = ((int64_t)p * c.P + c.B - 1) / c.B;
  int64_t e = ((int64_t)(p + 1) * c.P) / c.B;
  return (int)(e - s);
}
// PageAllocator::get_page_id, csrc/page_allocator.cpp:457-460
__device__ __forceinline__ int64_t blk_page(const DevState &c, int64_t id) {
  if (c.exact) return (id < INT_MAX) ? (int64_t)((int32_t)id / c.bpp) : id / c.bpp;
  return (id * c.B) / c.P;
}

// KVCacheManager.available_size, kv_cache_manager.py:412-423
__device__ __forceinline__ int64_t avail_size(const DevState &c, const int64_t *sc) {
  int64_t a = sc[SC_NAB] + sc[SC_NRB];
  if (sc[SC_INSHRINK]) return a;
  int64_t phys = (sc[SC_PHYS] > 0 ? sc[SC_PHYS] : 0) + sc[SC_RL_SIZE];
  int64_t fp = sc[SC_NFREE] < phys ? sc[SC_NFREE] : phys;
  return a + fp * c.bpp;
}
// KVCacheManager._get_num_alloced_blocks, kv_cache_manager.py:493-506
__device__ __forceinline__ int64_t alloced_blocks(const DevState &c, const int64_t *sc) {
  return sc[SC_NFULLPG] * c.bpp + sc[SC_NAVAILPG] * c.bpp - sc[SC_NAB] + sc[SC_NRB];
}

// ------------------------------------------------------------------------------------ thread-0 helpers
__device__ __forceinline__ void ev_push(const DevState &c, int64_t *sc, int page, int unmap) {
  int64_t n = sc[SC_NEV];
  if (n < c.ev_cap) c.ev[n] = (int64_t)page * 2 + unmap;
  sc[SC_NEV] = n + 1;
  c.pg_mapped[page] = unmap ? 0 : 1;
  sc[SC_NMAPPED] += unmap ? -1 : 1;
  sc[SC_PHYS] += unmap ? 1 : -1;
}
__device__ __forceinline__ void dq_push_back(int32_t *q, int mask, int64_t *sc, int ih, int is, int v) {
  q[(int)((sc[ih] + sc[is]) & mask)] = v;
  sc[is] += 1;
}
__device__ __forceinline__ int dq_pop_front(int32_t *q, int mask, int64_t *sc, int ih, int is) {
  int v = q[(int)(sc[ih] & mask)];
  sc[ih] = (sc[ih] + 1) & mask;
  sc[is] -= 1;
  return v;
}
__device__ __forceinline__ int dq_pop_back(int32_t *q, int mask, int64_t *sc, int ih, int is) {
  sc[is] -= 1;
  return q[(int)((sc[ih] + sc[is]) & mask)];
}

// PageAllocator::resize, csrc/page_allocator.cpp:312-401 (thread 0; rare)
__device__ __noinline__ int pa_resize(const DevState &c, int64_t *sc, int64_t new_mem_size, int *capacity_err) {
  int64_t new_pages = new_mem_size / c.P;
  int64_t total = sc[SC_NTOTAL];
  if (new_pages < total - sc[SC_NFREE]) return 0;
  if (new_pages == total) return 1;
  if (new_pages > total) {
    int64_t expand = new_pages - total;
    int64_t reuse = sc[SC_CL_SIZE] < expand ? sc[SC_CL_SIZE] : expand;
    if (expand - reuse > 0) {
      // The reference would append ids total..new-1 here, which alias live pages after a shrink
      // (page_allocator.cpp:343-346) and lie beyond the virtual range created at init.
      if (capacity_err) *capacity_err = 1;
      return 0;
    }
    for (int64_t i = 0; i < reuse; ++i) {
      int p = dq_pop_front(c.cl, c.q_mask, sc, SC_CL_HEAD, SC_CL_SIZE);
      dq_push_back(c.fl, c.q_mask, sc, SC_FL_HEAD, SC_FL_SIZE, p);
    }
    sc[SC_NFREE] += reuse;
    sc[SC_NTOTAL] = new_pages;
    return 1;
  }
  int64_t reclaim = total - new_pages;
  if (sc[SC_FL_SIZE] < reclaim) {
    if (sc[SC_RL_SIZE] == 0) return 0;
    while (sc[SC_RL_SIZE] > 0) {   // reserved pages are unmapped and appended to the free list
      int p = dq_pop_front(c.rl, c.q_mask, sc, SC_RL_HEAD, SC_RL_SIZE);
      ev_push(c, sc, p, 1);
      dq_push_back(c.fl, c.q_mask, sc, SC_FL_HEAD, SC_FL_SIZE, p);
    }
    if (sc[SC_FL_SIZE] < reclaim) return 0;
  }
  for (int64_t i = 0; i < reclaim; ++i) {
    int p = dq_pop_back(c.fl, c.q_mask, sc, SC_FL_HEAD, SC_FL_SIZE);
    dq_push_back(c.cl, c.q_mask, sc, SC_CL_HEAD, SC_CL_SIZE, p);
  }
  sc[SC_NFREE] -= reclaim;
  sc[SC_NTOTAL] = new_pages;
  return 1;
}

// KVCacheManager.resize, kv_cache_manager.py:381-401 (thread 0): returns the request status
__device__ int mgr_resize(const DevState &c, int64_t *sc, int64_t new_mem_size) {
  int cap_err = 0;
  if (pa_resize(c, sc, new_mem_size, &cap_err)) {
    if (sc[SC_INSHRINK]) { sc[SC_INSHRINK] = 0; sc[SC_TARGET] = -1; }
    return KVB_ST_OK;
  }
  if (cap_err) return KVB_ST_CAPACITY;
  if (sc[SC_NRB] > 0) return KVB_ST_ASSERT;
  sc[SC_INSHRINK] = 1;
  sc[SC_TARGET] = new_mem_size / c.B;
  return KVB_ST_FALSE;
}

// PageAllocator::trim, csrc/page_allocator.cpp:403-427 (thread 0)
__device__ void pa_trim(const DevState &c, int64_t *sc) {
  while (sc[SC_RL_SIZE] > 0) {
    int p = dq_pop_front(c.rl, c.q_mask, sc, SC_RL_HEAD, SC_RL_SIZE);
    ev_push(c, sc, p, 1);
    dq_push_back(c.fl, c.q_mask, sc, SC_FL_HEAD, SC_FL_SIZE, p);
  }
}

// ------------------------------------------------------------------------------------ block-wide scans
// Ordered exclusive prefix of a per-thread flag over the CTA: __ballot_sync + __popc inside a warp,
// warp totals through shared memory.  Returns the exclusive rank; *total receives the CTA total.
// Second stage of both scans: lane i holds warp i's total, two warp reductions (REDUX.SUM) give the sum over the warps
// below mine and the CTA total.  The kernel is issue-bound (one CTA = four warps per scheduler, every warp runs the
// uniform code): a 16-iteration loop over wtot cost ~80 issue slots per warp and call, this costs ~8
// (profiles/r2_step_instruction_counts.md).
__device__ __forceinline__ void warp_totals(const int32_t *wtot, int lane, int w, int *base, int *tot) {
  static_assert(NW <= 32, "one lane per warp total");
  const int t = lane < NW ? wtot[lane] : 0;
  *tot = __reduce_add_sync(0xffffffffu, t);
  *base = __reduce_add_sync(0xffffffffu, lane < w ? t : 0);
}
__device__ __forceinline__ int block_rank(bool flag, int32_t *wtot, int *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned m = __ballot_sync(0xffffffffu, flag);
  int r = __popc(m & ((1u << lane) - 1u));
  if (lane == 0) wtot[w] = __popc(m);
  __syncthreads();
  int base, tot;
  warp_totals(wtot, lane, w, &base, &tot);
  __syncthreads();
  *total = tot;
  return base + r;
}
// Ordered exclusive prefix sum of a per-thread int over the CTA (warp shuffles + shared memory).
__device__ __forceinline__ int block_exscan(int v, int32_t *wtot, int *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) wtot[w] = x;
  __syncthreads();
  int base, tot;
  warp_totals(wtot, lane, w, &base, &tot);
  __syncthreads();
  *total = tot;
  return base + x - v;
}

// One-barrier forms.  The trailing barrier of block_rank / block_exscan only protects `wtot` against the NEXT scan that
// writes it; a scan whose buffer belongs to one call site alone does not need it, provided every thread passes some
// other block barrier between reading the totals here and the next execution of the same site (true of every site
// below: each is followed by barriers of its own function before that function can run again).
__device__ __forceinline__ int block_rank_1b(bool flag, int32_t *site, int *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) site[w] = __popc(m);
  __syncthreads();
  int base, tot;
  warp_totals(site, lane, w, &base, &tot);
  *total = tot;
  return base + __popc(m & ((1u << lane) - 1u));
}
__device__ __forceinline__ int block_exscan_1b(int v, int32_t *site, int *total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) site[w] = x;
  __syncthreads();
  int base, tot;
  warp_totals(site, lane, w, &base, &tot);
  *total = tot;
  return base + x - v;
}
// Two ranks in one pass: both counts are <= NT < 2^16, so the warp totals travel packed in one word.
__device__ __forceinline__ void block_rank2_1b(bool fa, bool fb, int32_t *site, int *ra, int *ta, int *rb, int *tb) {
  static_assert(NT < (1 << 16), "packed totals");
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
  if (lane == 0) site[w] = __popc(ma) | (__popc(mb) << 16);
  __syncthreads();
  int base, tot;
  warp_totals(site, lane, w, &base, &tot);
  const unsigned lt = (1u << lane) - 1u;
  *ra = (base & 0xffff) + __popc(ma & lt);
  *ta = tot & 0xffff;
  *rb = (int)((unsigned)base >> 16) + __popc(mb & lt);
  *tb = (int)((unsigned)tot >> 16);
}
// alloc_run_fast's three scans (exclusive sum, rank, exclusive sum) and its vote in one pass: rows 0..3 of `site`.
__device__ __forceinline__ void block_scan3_1b(int v0, bool f1, int v2, bool bad, int32_t (*site)[NW], int *x0, int *t0,
                                               int *x1, int *t1, int *x2, int *t2, bool *any_bad) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int a = v0, b = v2;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int ya = __shfl_up_sync(0xffffffffu, a, d), yb = __shfl_up_sync(0xffffffffu, b, d);
    if (lane >= d) { a += ya; b += yb; }
  }
  const unsigned m = __ballot_sync(0xffffffffu, f1);
  const bool wbad = __any_sync(0xffffffffu, bad);
  if (lane == 31) { site[0][w] = a; site[2][w] = b; }
  if (lane == 0) { site[1][w] = __popc(m); site[3][w] = wbad ? 1 : 0; }
  __syncthreads();
  int base, tot;
  warp_totals(site[0], lane, w, &base, &tot);
  *x0 = base + a - v0; *t0 = tot;
  warp_totals(site[1], lane, w, &base, &tot);
  *x1 = base + __popc(m & ((1u << lane) - 1u)); *t1 = tot;
  warp_totals(site[2], lane, w, &base, &tot);
  *x2 = base + b - v2; *t2 = tot;
  warp_totals(site[3], lane, w, &base, &tot);
  *any_bad = tot != 0;
}

// ------------------------------------------------------------------------------------ libstdc++ policy
// std::__detail::_Prime_rehash_policy::_M_next_bkt (libstdc++ 13, hashtable_c++0x.cc) with the
// library's own __prime_list uploaded at init.
__device__ int next_bkt(const DevState &c, const uint32_t *primes32, int64_t n) {
  // __fast_bkt[] = {2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13}, one nibble per entry
  if (n < 14) return n == 0 ? 1 : (int)((0x22235577BBBBDDull >> (4 * (13 - n))) & 15);
  if ((uint64_t)n <= primes32[PRIMES_SM - 1]) {
    int lo = 6, hi = PRIMES_SM - 1;   // lower_bound over [__prime_list + 6, ...)
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (primes32[mid] < (uint32_t)n) lo = mid + 1; else hi = mid;
    }
    return (int)primes32[lo];
  }
  int lo = 6, hi = c.n_primes - 1;   // lower_bound over [__prime_list + 6, last)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (c.primes[mid] < (uint64_t)n) lo = mid + 1; else hi = mid;
  }
  return (int)c.primes[lo];
}

// The same for the bucket counts a free() of at most NT ids can reach (n <= __prime_list[PRIMES_SM - 1], 32-bit
// arithmetic, branch-free lower_bound over the shared copy of the table): ~30 instructions instead of ~170, and every
// warp of the CTA runs it.
__device__ __forceinline__ int next_bkt_small(const uint32_t *primes32, int n) {
  if (n < 14) return n <= 0 ? 1 : (int)((0x22235577BBBBDDull >> (4 * (13 - n))) & 15);
  int lo = 6;
#pragma unroll
  for (int len = PRIMES_SM - 6; len > 1;) {       // first index in [lo, lo + len) whose prime is >= n
    const int half = len >> 1;
    if (primes32[lo + half - 1] < (uint32_t)n) lo += half;
    len -= half;
  }
  return (int)primes32[lo];
}
// _M_need_rehash's request when key number placed + 1 arrives at a table of nb buckets.  The replay only asks when the
// table is full (placed == nb): that answer sits in the table that follows primes32 in shared memory (Smem::grow16).
__device__ __forceinline__ int next_bkt_grow(const uint32_t *primes32, int placed, int nb) {
  if (placed == nb && nb < kGrowTab) return reinterpret_cast<const uint16_t *>(primes32 + PRIMES_SM)[nb];
  const int want = placed + 2 > 2 * nb ? placed + 2 : 2 * nb;
  return next_bkt_small(primes32, want);
}

// Replays std::unordered_map<int64_t, ...>::operator[] for keys[0..D) (distinct, first-appearance
// order) after reserve(n_ids / bpp + 1) and returns the container's iteration order as slot indices:
// _M_insert_bucket_begin / _M_rehash_aux of libstdc++'s _Hashtable with unique keys, identity hash,
// max_load_factor 1 (csrc/page_allocator.cpp:471-498 builds exactly this map; pybind11 walks it).
// bkt[b] holds the node BEFORE bucket b's first node: -1 empty, -2 the before-begin sentinel.
// Sequential (one lane); used when a call names more than 32 distinct pages.
__device__ __forceinline__ void ht_order_seq(const DevState &c, const uint32_t *primes32, const int32_t *keys, int D,
                                             int nb, int32_t *next, int32_t *bkt, int32_t *ord) {
  int head = -1;
  for (int d = 0; d < D; ++d) {
    unsigned k = (unsigned)keys[d];
    if (d + 1 > nb) {   // _M_need_rehash: n_elt + 1 > _M_next_resize (= nb at load factor 1)
      int64_t want = (int64_t)d + 2;
      if (want < 2 * (int64_t)nb) want = 2 * (int64_t)nb;
      unsigned nb2 = (unsigned)next_bkt(c, primes32, want);
      for (unsigned b = 0; b < nb2; ++b) bkt[b] = -1;
      int p = head, bbegin = 0;
      head = -1;
      while (p >= 0) {   // _M_rehash_aux(unique keys)
        int nx = next[p];
        int b = (int)((unsigned)keys[p] % nb2);
        int prev = bkt[b];
        if (prev == -1) {
          next[p] = head;
          if (head >= 0) bkt[bbegin] = p;
          head = p;
          bkt[b] = -2;
          bbegin = b;
        } else if (prev == -2) {
          next[p] = head;
          head = p;
        } else {
          next[p] = next[prev];
          next[prev] = p;
        }
        p = nx;
      }
      nb = (int)nb2;
    }
    int b = (int)(k % (unsigned)nb);
    int prev = bkt[b];
    if (prev == -1) {            // empty bucket: new node becomes the global head
      next[d] = head;
      if (head >= 0) bkt[(unsigned)keys[head] % (unsigned)nb] = d;
      head = d;
      bkt[b] = -2;
    } else if (prev == -2) {     // bucket starts at the list head: insert after before-begin
      next[d] = head;
      head = d;
    } else {                     // insert after the bucket's before-node
      next[d] = next[prev];
      next[prev] = d;
    }
  }
  int i = 0;
  for (int p = head; p >= 0; p = next[p]) ord[i++] = p;
}

// The same order for D <= 32 without any pointer chasing, one warp, keys in registers.
// Between two rehashes the container obeys one rule: a key joins the FRONT of its bucket's run, and a key
// of a new bucket becomes the list head.  So a phase (all keys placed under one bucket count nb, in
// sequence S) ends as: bucket runs ordered by LAST-first-appearance first, and inside a run the later key
// first — i.e. S sorted by (first position of my bucket in S, my position in S), both descending.  A rehash
// re-places the current list front to back under the new nb, which is the same rule applied to S = current
// order; keys inserted afterwards simply extend S.  Lane l holds the key at position l of S.
// Keys of the runs whose leader (first key of a bucket in S) sits at a lane above g.  Leaders hold their run's size
// (<= 32): six votes slice the sizes by bit, every lane sums the bits above its own g — independent instructions,
// where a suffix scan over the lanes is a chain of five dependent shuffles plus a sixth to fetch the leader's value.
__device__ __forceinline__ int keys_of_later_runs(bool leader, int size, int g) {
  const unsigned above = ~((2u << g) - 1u);       // g = 31: no lane above
  int r = 0;
#pragma unroll
  for (int b = 0; b < 6; ++b) r += __popc(__ballot_sync(0xffffffffu, leader && ((size >> b) & 1)) & above) << b;
  return r;
}

#ifdef KVB_PROFILE
#define KVB_PROF_W(slot) do { if (threadIdx.x == 0) { long long t_ = clock64(); c.prof[slot] += t_ - *pt; c.prof[16 + (slot)] += 1; *pt = t_; } } while (0)
#else
#define KVB_PROF_W(slot) do { } while (0)
#endif
__device__ __forceinline__ void ht_order_warp(const DevState &c, const uint32_t *primes32, const int32_t *keys,
                                              int D, int nb, int32_t *ord, int32_t *scratch64, long long *pt = nullptr) {
  const int lane = threadIdx.x & 31;
  const unsigned gt = ~((2u << lane) - 1u);   // lanes above mine
  // cur_slot / cur_key = slot index and key of the element at S position `lane` (valid for lane < placed)
  int cur_slot = lane;
  unsigned cur_key = lane < D ? (unsigned)keys[lane] : 0u;
  int placed = 0;   // keys already in the container
  if (nb == 2 && D > 2) {
    // Two keys under two buckets always end as [key 1, key 0] — same bucket: the later key joins the front of the run;
    // different buckets: the later key becomes the list head — so the first phase of the commonest case (a free() of
    // fewer than 2 * blocks-per-page ids reserves 2 buckets) is a swap of lanes 0 and 1.
    placed = 2;
    if (lane < 2) { cur_slot = 1 - lane; cur_key = (unsigned)keys[1 - lane]; }
    nb = next_bkt_grow(primes32, 2, 2);
  }
  while (true) {
    // keys inserted under this bucket count: until the container holds nb keys (the next insert rehashes)
    const int upto = D < nb ? D : nb;
    // S = [current order of `placed` keys] ++ [slots placed .. upto-1 in insertion order]
    if (lane >= placed && lane < upto) { cur_slot = lane; cur_key = (unsigned)keys[lane]; }
    const bool in = lane < upto;
    const unsigned b = in ? cur_key % (unsigned)nb : 0xffffffffu - lane;
    KVB_PROF_W(12);
    const unsigned grp = __match_any_sync(0xffffffffu, b);
    KVB_PROF_W(13);
    const int gfirst = __ffs(grp) - 1;                  // first position of my bucket in S
    const int later_in_grp = __popc(grp & gt);
    // keys whose bucket first appears after mine come before me: suffix sum of run sizes over run leaders
    const int before = keys_of_later_runs(in && gfirst == lane, __popc(grp), gfirst);
    const int newpos = before + later_in_grp;
    KVB_PROF_W(14);
    // permute through shared memory: position p takes the element whose newpos == p
    if (in) { scratch64[newpos] = cur_slot; scratch64[32 + newpos] = (int)cur_key; }
    __syncwarp();
    if (in) { cur_slot = scratch64[lane]; cur_key = (unsigned)scratch64[32 + lane]; }
    __syncwarp();
    placed = upto;
    KVB_PROF_W(15);
    if (placed >= D) break;
    nb = next_bkt_grow(primes32, placed, nb);           // _M_need_rehash -> _M_next_bkt(max(n+1+1, 2*nb)); nb < D <= 32 here
  }
  if (lane < D) ord[lane] = cur_slot;
}

// The same rule for 32 < D <= 32 * HT_WK keys, still one warp: lane l holds positions l, 32 + l, 64 + l, ... of S
// in registers; per phase the bucket of every key gets its first position (atomicMin into a shared table), each
// key the number of later keys of its bucket (chunks walked from the back: __match_any_sync inside a chunk plus
// the bucket's running count), run leaders a suffix sum of run sizes, and S is permuted through shared memory.
// A request that decoded for a few hundred tokens names 40-130 pages: the sequential replay costs ~500 cycles a
// key there, this ~40.  tab_first / tab_cnt hold >= the final bucket count entries; perm >= 2 * 32 * HT_WK words.
constexpr int HT_WK = 8;
__device__ __forceinline__ void ht_order_warp_multi(const DevState &c, const uint32_t *primes32, const int32_t *keys,
                                                    int D, int nb, int32_t *ord, int32_t *tab_first, int32_t *tab_cnt,
                                                    int64_t *perm, int placed0 = 0, int slot0 = 0, unsigned key0 = 0u) {
  const int lane = threadIdx.x & 31;
  const unsigned gt = ~((2u << lane) - 1u);   // lanes above mine
  int32_t *suf_at = reinterpret_cast<int32_t *>(perm + 32 * HT_WK);
  int slot[HT_WK];
  unsigned key[HT_WK];
#pragma unroll
  for (int k = 0; k < HT_WK; ++k) {
    const int p = 32 * k + lane;
    slot[k] = p;
    key[k] = p < D ? (unsigned)keys[p] : 0u;
  }
  // hand-over from the single-key phases (ht_order_hybrid): positions below placed0 (<= 32) hold the container's
  // current order, `nb` is the bucket count of the phase about to run
  if (lane < placed0) { slot[0] = slot0; key[0] = key0; }
  while (true) {
    const int upto = D < nb ? D : nb;        // keys in the container when this bucket count ends
    int b[HT_WK], later[HT_WK];
#pragma unroll
    for (int k = 0; k < HT_WK; ++k) {
      const int p = 32 * k + lane;
      b[k] = p < upto ? (int)(key[k] % (unsigned)nb) : -1;
      if (b[k] >= 0) { tab_first[b[k]] = INT_MAX; tab_cnt[b[k]] = 0; }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < HT_WK; ++k)
      if (b[k] >= 0) atomicMin(&tab_first[b[k]], 32 * k + lane);
    __syncwarp();
    // keys of my bucket at later positions: chunks from the back
#pragma unroll
    for (int k = HT_WK - 1; k >= 0; --k) {
      later[k] = 0;
      if (32 * k >= upto) continue;          // uniform
      const int bb = b[k] >= 0 ? b[k] : -1 - lane;
      const unsigned m = __match_any_sync(0xffffffffu, bb);
      if (b[k] >= 0) later[k] = tab_cnt[b[k]] + __popc(m & gt);
      __syncwarp();
      if (b[k] >= 0 && (m & gt) == 0) tab_cnt[b[k]] += __popc(m);   // the group's highest lane
      __syncwarp();
    }
    // inclusive suffix sum, over positions, of the run sizes held by run leaders (first key of a bucket)
    int carry = 0;
#pragma unroll
    for (int k = HT_WK - 1; k >= 0; --k) {
      if (32 * k >= upto) continue;          // uniform
      const int p = 32 * k + lane;
      int v = (b[k] >= 0 && tab_first[b[k]] == p) ? tab_cnt[b[k]] : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_down_sync(0xffffffffu, v, d);
        if (lane + d < 32) v += y;
      }
      if (b[k] >= 0) suf_at[p] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 0);
    }
    __syncwarp();
    // new position = keys of buckets that first appear after mine + later keys of my bucket
#pragma unroll
    for (int k = 0; k < HT_WK; ++k)
      if (b[k] >= 0) {
        const int newpos = suf_at[tab_first[b[k]]] - tab_cnt[b[k]] + later[k];
        perm[newpos] = ((int64_t)slot[k] << 32) | key[k];
      }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < HT_WK; ++k)
      if (b[k] >= 0) {
        const int64_t e = perm[32 * k + lane];
        slot[k] = (int)(e >> 32);
        key[k] = (unsigned)e;
      }
    __syncwarp();
    if (upto >= D) break;
    nb = next_bkt_grow(primes32, upto, nb);   // _M_need_rehash -> _M_next_bkt(max(n + 1 + 1, 2 * nb)); nb < D <= 256 here
  }
#pragma unroll
  for (int k = 0; k < HT_WK; ++k)
    if (32 * k + lane < D) ord[32 * k + lane] = slot[k];
}

// 32 < D <= 32 * HT_WK: the first rehash phases hold at most 32 keys (bucket counts 2, 5, 11, 23, ...) and run in the
// one-key-per-lane form of ht_order_warp (registers and votes only, ~4x cheaper per phase than the multi-key form);
// the phase that first exceeds 32 keys, and the ones after it, continue in ht_order_warp_multi.
// The single-key phases of a replay with D > 32 keys: runs rehash phases in the one-key-per-lane form while they hold
// at most 32 keys.  On return *nb is the bucket count of the NEXT phase, *placed the keys placed so far (<= 32), and
// lane l < *placed holds the element at position l of the container's order (cur_slot / cur_key).
__device__ __forceinline__ void ht_order_prefix(const DevState &c, const uint32_t *primes32, const int32_t *keys, int D,
                                                int *nb_io, int *placed_out, int *slot_out, unsigned *key_out,
                                                int32_t *scratch64) {
  const int lane = threadIdx.x & 31;
  const unsigned gt = ~((2u << lane) - 1u);
  int nb = *nb_io;
  int cur_slot = lane;
  unsigned cur_key = lane < D ? (unsigned)keys[lane] : 0u;
  int placed = 0;
  if (nb == 2 && D > 2) {          // first phase = swap of lanes 0 and 1 (see ht_order_warp)
    placed = 2;
    if (lane < 2) { cur_slot = 1 - lane; cur_key = (unsigned)keys[1 - lane]; }
    nb = next_bkt_grow(primes32, 2, 2);
  }
  while ((D < nb ? D : nb) <= 32) {
    const int upto = D < nb ? D : nb;
    if (lane >= placed && lane < upto) { cur_slot = lane; cur_key = (unsigned)keys[lane]; }
    const bool in = lane < upto;
    const unsigned b = in ? cur_key % (unsigned)nb : 0xffffffffu - lane;
    const unsigned grp = __match_any_sync(0xffffffffu, b);
    const int gfirst = __ffs(grp) - 1;
    const int later_in_grp = __popc(grp & gt);
    const int before = keys_of_later_runs(in && gfirst == lane, __popc(grp), gfirst);
    const int newpos = before + later_in_grp;
    if (in) { scratch64[newpos] = cur_slot; scratch64[32 + newpos] = (int)cur_key; }
    __syncwarp();
    if (in) { cur_slot = scratch64[lane]; cur_key = (unsigned)scratch64[32 + lane]; }
    __syncwarp();
    placed = upto;
    if (placed >= D) break;                                  // (cannot happen for D > 32; kept for safety)
    nb = next_bkt_grow(primes32, placed, nb);                // placed <= 32, nb <= 32 here
  }
  *nb_io = nb;
  *placed_out = placed;
  *slot_out = cur_slot;
  *key_out = cur_key;
}

// 32 < D <= 32 * HT_WK on ONE warp: the first rehash phases hold at most 32 keys (bucket counts 2, 5, 11, 23, ...) and run
// in the one-key-per-lane form (registers and votes only); the phase that first exceeds 32 keys, and the ones after
// it, continue in ht_order_warp_multi.
__device__ __forceinline__ void ht_order_hybrid(const DevState &c, const uint32_t *primes32, const int32_t *keys, int D,
                                                int nb, int32_t *ord, int32_t *tab_first, int32_t *tab_cnt,
                                                int64_t *perm) {
  const int lane = threadIdx.x & 31;
  int placed, cur_slot;
  unsigned cur_key;
  ht_order_prefix(c, primes32, keys, D, &nb, &placed, &cur_slot, &cur_key, reinterpret_cast<int32_t *>(perm));
  if (placed >= D) {
    if (lane < D) ord[lane] = cur_slot;
    return;
  }
  ht_order_warp_multi(c, primes32, keys, D, nb, ord, tab_first, tab_cnt, perm, placed, cur_slot, cur_key);
}

// The later phases BLOCK-wide (all NT threads call; D <= 256): thread t holds the element at position t of the
// container's order, warp c = t / 32 is chunk c.  Per phase, with S = current order ++ newly inserted keys:
//   first position of my bucket in S         atomicMin into a shared table (one atomic per bucket and chunk);
//   keys of my bucket at later positions     inside my chunk: __match_any_sync; in later chunks: the per-(bucket,
//                                            chunk) counts are PACKED into one shared 64-bit word per bucket (8 chunks x
//                                            8 bits), so one read gives them all — no chunk-by-chunk walk;
//   keys of buckets that first appear later  suffix sum, over positions, of the bucket sizes held by bucket leaders;
//   new position = the two sums; the permutation goes through shared memory in place.
// Four block barriers per phase whatever the key count; the one-warp form walks its 8 chunks twice per phase.
__device__ void ht_order_block(const DevState &c, Smem &s, const int32_t *keys, int D, int nb, int placed, int32_t *ord) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const unsigned gt = ~((2u << lane) - 1u), lt = (1u << lane) - 1u;
  int32_t *tab_first = s.ht_bkt;                  // [nb <= 1024]
  int32_t *suf = s.ht_bkt + 1024;                 // [256]
  int32_t *cur_slot = s.ht_bkt + 1280;            // [256] element at position t
  int32_t *cur_key = s.ht_bkt + 1536;
  unsigned long long *packed = reinterpret_cast<unsigned long long *>(s.tmp64);   // [nb <= NT]
  while (true) {
    const int upto = D < nb ? D : nb;
    const bool in = t < upto;
    int slot = 0;
    unsigned key = 0;
    if (in) {
      if (t >= placed) { slot = t; key = (unsigned)keys[t]; }
      else { slot = cur_slot[t]; key = (unsigned)cur_key[t]; }
    }
    const int b = in ? (int)(key % (unsigned)nb) : -1;
    if (in) { tab_first[b] = INT_MAX; packed[b] = 0ull; }
    __syncthreads();
    const unsigned m = __match_any_sync(0xffffffffu, in ? b : -1 - lane);
    const int intra_later = __popc(m & gt);
    if (in && (m & lt) == 0) {                     // lowest lane of my bucket inside this chunk
      atomicMin(&tab_first[b], t);
      atomicAdd(&packed[b], (unsigned long long)__popc(m) << (8 * w));
    }
    __syncthreads();
    int gfirst = 0, cnt = 0, later = intra_later;
    if (in) {
      gfirst = tab_first[b];
      const unsigned long long pk = packed[b];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ck = (int)((pk >> (8 * k)) & 0xffull);
        cnt += ck;
        if (k > w) later += ck;
      }
    }
    int v = (in && gfirst == t) ? cnt : 0;         // inclusive suffix sum over positions
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_down_sync(0xffffffffu, v, d);
      if (lane + d < 32) v += y;
    }
    if (lane == 0) s.wtot[w] = v;
    if (in) suf[t] = v;                            // suffix inside my chunk; the chunks above are added by the reader
    __syncthreads();
    int above = (lane < 8 && lane < NW) ? s.wtot[lane] : 0;   // lane i: keys held by chunks i, i + 1, ... (D <= 256: 8 chunks)
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) above += __shfl_down_sync(0xffffffffu, above, d);   // lanes >= 8 hold 0
    above = __shfl_sync(0xffffffffu, above, ((in ? gfirst : 0) >> 5) + 1);           // chunks above my bucket's first key
    if (in) {
      const int newpos = suf[gfirst] + above - cnt + later;
      cur_slot[newpos] = slot;
      cur_key[newpos] = (int)key;
    }
    __syncthreads();
    placed = upto;
    if (upto >= D) break;
    nb = next_bkt_grow(s.primes32, upto, nb);      // _M_need_rehash -> _M_next_bkt(max(n + 1 + 1, 2 * nb)); nb < D <= 256
  }
  if (t < D) ord[t] = cur_slot[t];
  __syncthreads();
}

// bucket count the table ends with (decides shared vs global tables before the replay)
__device__ int ht_final_buckets(const DevState &c, const uint32_t *primes32, int D, int nb0) {
  int nb = nb0;
  while (D > nb) {
    int64_t want = (int64_t)nb + 2;
    if (want < 2 * (int64_t)nb) want = 2 * (int64_t)nb;
    nb = next_bkt(c, primes32, want);
  }
  return nb;
}

// ==================================================================================== alloc run
// Thread 0: walk requests [sim.op, run_end) and cut the free-block stream into segments
// (kv_cache_manager.py:250-304 for every request, in order).
__device__ void alloc_walk(const DevState &c, Smem &s, const LaunchArgs &a, int chunk_base, int run_end,
                           int to_reserved) {
  AllocSim &m = s.sim;
  int64_t *sc = s.sc;
  m.nseg = 0;
  m.nfresh = 0;
  m.round_base = m.out_pos;
  while (true) {
    if (!m.req_active) {
      if (m.op >= run_end) { m.done = 1; break; }
      if (m.nseg >= SEG_CAP - 2) goto round_full;
      const kvb_op &o = s.ops[m.op - chunk_base];
      int64_t need = o.n;
      kvb_res r{};
      if (avail_size(c, sc) < need || need < 0) {   // -> None, state unchanged
        r.status = to_reserved ? KVB_ST_FALSE : KVB_ST_NONE;
        r.n = 0;
        r.off = m.out_pos;
        r.avail = avail_size(c, sc);
        a.res[m.op] = r;
        m.op++;
        continue;
      }
      if (!to_reserved && m.out_pos + need > a.out_cap) {   // output buffer exhausted: stop before this request
        s.misc[MI_STOP] = 1;
        m.done = 1;
        break;
      }
      m.req_active = 1;
      m.req_rem = need;
      m.req_off = m.out_pos;
      m.req_n = need;
      m.failed = 0;
      // reserved blocks first (:272-277)
      int64_t take = sc[SC_NRB] < need ? sc[SC_NRB] : need;
      if (take > 0) {
        Seg g;
        g.out_rel = (int32_t)(m.out_pos - m.round_base);
        g.page = -1;
        g.base = (int32_t)sc[SC_RB_HEAD];
        g.len_kind = (int32_t)take | (SEG_RB << 28);
        s.segs[m.nseg++] = g;
        sc[SC_RB_HEAD] = (sc[SC_RB_HEAD] + take) & c.rb_mask;
        sc[SC_NRB] -= take;
        m.out_pos += take;
        m.req_rem -= take;
      }
    }
    // page loop (:279-302)
    while (m.req_rem > 0) {
      if (m.nseg >= SEG_CAP - 1 || m.nfresh >= FRESH_CAP) goto round_full;
      if (m.cur_p < 0) {
        // avail_pages.popitem(): last live entry of the stack
        int i = m.ap_top - 1;
        int found = -1;
        while (i >= 0) {
          int k = m.ap_tail0 - 1 - i;
          int e = (k < APW) ? s.apw[k] : c.ap[i];
          if (e >= 0) { found = e; break; }
          --i;
        }
        if (found >= 0) {
          int k = m.ap_tail0 - 1 - i;
          m.cur_p = found;
          m.cur_cnt = (k < APW) ? s.apcnt[k] : c.pg_cnt[found];
          m.cur_head = (k < APW) ? s.aphead[k] : c.pg_head[found];
          m.cur_fresh = 0;
          m.cur_idx = i;
          m.ap_top = i + 1;
        } else {
          m.ap_top = 0;
          // PageAllocator::alloc_page (page_allocator.cpp:161-237): reserved list first, then free list
          int j = m.newpg_j;
          int p;
          int from_reserved = sc[SC_RL_SIZE] > 0;
          if (!from_reserved && sc[SC_FL_SIZE] <= 0) {   // "No free pages left"
            m.failed = 1;
            break;
          }
          if (from_reserved) {
            p = (j < NEWPG_W) ? s.newpg[j] : c.rl[(int)(sc[SC_RL_HEAD] & c.q_mask)];
            sc[SC_RL_HEAD] = (sc[SC_RL_HEAD] + 1) & c.q_mask;
            sc[SC_RL_SIZE] -= 1;
          } else {
            p = (j < NEWPG_W) ? s.newpg[j] : c.fl[(int)(sc[SC_FL_HEAD] & c.q_mask)];
            sc[SC_FL_HEAD] = (sc[SC_FL_HEAD] + 1) & c.q_mask;
            sc[SC_FL_SIZE] -= 1;
            ev_push(c, sc, p, 0);
          }
          m.newpg_j = j + 1;
          sc[SC_NFREE] -= 1;
          int nb = pg_nblk(c, p);
          c.pg_head[p] = 0;
          if (nb <= 0) {   // page holds no whole block: parked in full_pages (:287-289)
            c.pg_state[p] = PG_FULL;
            c.pg_cnt[p] = 0;
            sc[SC_NFULLPG] += 1;
            continue;
          }
          sc[SC_NAB] += nb;
          s.fresh[m.nfresh++] = make_int2(p, nb);
          m.cur_p = p;
          m.cur_cnt = nb;
          m.cur_head = 0;
          m.cur_fresh = 1;
          m.cur_idx = -1;
        }
      }
      int t = (int)(m.cur_cnt < m.req_rem ? m.cur_cnt : m.req_rem);
      Seg g;
      g.out_rel = (int32_t)(m.out_pos - m.round_base);
      g.page = m.cur_p;
      g.base = m.cur_head;
      g.len_kind = t | ((m.cur_fresh ? SEG_FRESH : SEG_RING) << 28);
      s.segs[m.nseg++] = g;
      m.cur_cnt -= t;
      m.cur_head += t;
      if (!m.cur_fresh && m.cur_head >= c.bpp) m.cur_head -= c.bpp;
      sc[SC_NAB] -= t;
      m.out_pos += t;
      m.req_rem -= t;
      if (m.cur_cnt == 0) {   // page.full() -> full_pages
        int p = m.cur_p;
        c.pg_state[p] = PG_FULL;
        c.pg_cnt[p] = 0;
        c.pg_head[p] = m.cur_fresh ? 0 : m.cur_head;
        sc[SC_NFULLPG] += 1;
        if (m.cur_idx >= 0) { sc[SC_NAVAILPG] -= 1; m.ap_top = m.cur_idx; }
        m.cur_p = -1;
      }
    }
    {
      // request finished (or failed)
      kvb_res r{};
      r.status = m.failed ? KVB_ST_NO_PAGES : KVB_ST_OK;
      r.n = m.failed ? 0 : (int32_t)m.req_n;
      r.off = m.req_off;
      r.avail = avail_size(c, sc);
      a.res[m.op] = r;
      m.req_active = 0;
      m.req_rem = 0;
      m.op++;
    }
  }
round_full:
  if (m.done) {
    // run over: the page under consumption goes back to avail_pages as the newest entry
    if (m.cur_p >= 0) {
      int p = m.cur_p;
      c.pg_cnt[p] = m.cur_cnt;
      if (m.cur_fresh) {
        // ring of a fresh page is materialised as the identity; its head is the consumed prefix
        c.pg_head[p] = m.cur_head;
        c.pg_state[p] = PG_AVAIL;
        c.ap[m.ap_top] = p;
        c.pg_pos[p] = m.ap_top;
        m.ap_top += 1;
        sc[SC_NAVAILPG] += 1;
      } else {
        c.pg_head[p] = m.cur_head;
      }
      m.cur_p = -1;
    }
    sc[SC_AP_TAIL] = m.ap_top;
  }
}

// All threads: materialise the round — fresh-page rings, output ids, bitmap bits.
__device__ void alloc_emit(const DevState &c, Smem &s, const LaunchArgs &a, int64_t *dst, int64_t dst_mask) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const AllocSim &m = s.sim;
  // fresh pages: ring[p][t] = t (InternalPage::init, page_allocator.cpp:44-53)
  for (int f = w; f < m.nfresh; f += NW) {
    int2 pg = s.fresh[f];
    uint32_t *r = c.ring + (size_t)pg.x * c.bpp;
    for (int t = lane; t < pg.y; t += 32) r[t] = (uint32_t)t;
  }
  for (int g = w; g < m.nseg; g += NW) {
    Seg sg = s.segs[g];
    int kind = sg.len_kind >> 28, len = sg.len_kind & 0x0fffffff;
    int64_t o = m.round_base + sg.out_rel;
    if (kind == SEG_RB) {
      for (int t = lane; t < len; t += 32) {
        int64_t id = c.rb[(sg.base + t) & c.rb_mask];
        if (dst_mask) dst[(o + t) & dst_mask] = id; else dst[o + t] = id;
      }
    } else {
      int64_t start = pg_start(c, sg.page);
      const uint32_t *r = c.ring + (size_t)sg.page * c.bpp;
      for (int t = lane; t < len; t += 32) {
        int64_t id;
        if (kind == SEG_FRESH) {
          id = start + sg.base + t;
        } else {
          int pos = sg.base + t;
          if (pos >= c.bpp) pos -= c.bpp;
          id = start + r[pos];
        }
        if (dst_mask) dst[(o + t) & dst_mask] = id; else dst[o + t] = id;
        atomicOr(&c.bitmap[id >> 5], 1u << (id & 31));
      }
    }
  }
}

// Fast path of an alloc run: every request succeeds and the free-block stream is a pure function of the
// state at the start of the run, so nothing has to be walked in sequence:
//   * request i takes stream positions [off_i, off_i + n_i), off = exclusive prefix sum of the sizes;
//   * the stream is [reserved_blocks][avail pages newest first, each page's FIFO][fresh pages, whole];
//   * in an exact geometry (block size divides page size) available_size() drops by exactly one per
//     block taken, across page boundaries too, so request i sees available_size() = A0 - off_i and the
//     run cannot fail iff A0 >= total.
// Preconditions (else the caller takes the sequential walk, which is the general restatement):
// exact geometry, not shrinking, <= NT requests, A0 >= total, the avail pages needed sit in the
// prefetched window, at most NEWPG_W fresh pages.  Returns 1 if the run was handled (uniform).
__device__ int alloc_run_fast(const DevState &c, Smem &s, const LaunchArgs &a, int chunk_base, int op0, int run_end) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int64_t *sc = s.sc;
  const int R = run_end - op0;
  if (!c.exact || sc[SC_INSHRINK] || R > NT) return 0;
  const int tail = (int)sc[SC_AP_TAIL];
  const int rls = (int)sc[SC_RL_SIZE], fls = (int)sc[SC_FL_SIZE];
  // ---- sizes -> offsets; prefetch the top of the avail stack and the upcoming fresh pages ----
  int n_i = 0;
  if (tid < R) n_i = s.ops[op0 - chunk_base + tid].n;
  int e = -1, ecnt = 0, ehead = 0;
  if (tid < APW) {
    int i = tail - 1 - tid;
    e = (i >= 0) ? c.ap[i] : -1;
    if (e >= 0) { ecnt = c.pg_cnt[e]; ehead = c.pg_head[e]; }
  }
  if (tid >= NT - NEWPG_W) {
    int j = tid - (NT - NEWPG_W);
    int v = -1;
    if (j < rls) v = c.rl[(int)((sc[SC_RL_HEAD] + j) & c.q_mask)];
    else if (j - rls < fls) v = c.fl[(int)((sc[SC_FL_HEAD] + (j - rls)) & c.q_mask)];
    s.newpg[j] = v;
  }
  // one pass, one barrier: request offsets, position among live entries (newest first), cumulative free blocks, and
  // the vote on odd sizes (they take the general walk)
  bool neg;
  int T, off_i, nlive, li, sum_av, cum;
  {
    const bool odd = n_i < 0 || n_i > (1 << 21);
    block_scan3_1b(odd ? 0 : n_i, e >= 0, e >= 0 ? ecnt : 0, odd, &s.wsite[WS_AF0], &off_i, &T, &li, &nlive, &cum, &sum_av, &neg);
  }
  if (e >= 0) {
    s.cpage[li] = e;
    s.ccnt[li] = ecnt;
    s.chead[li] = ehead;
    s.csidx[li] = tail - 1 - tid;
    s.ccum[li] = cum;
  }
  const int64_t A0 = avail_size(c, sc);
  const int64_t cursor = sc[SC_OUT_CURSOR];
  const int nrb = (int)sc[SC_NRB];
  const int take_rb = nrb < T ? nrb : T;
  const int rem1 = T - take_rb;
  const int take_av = rem1 < sum_av ? rem1 : sum_av;
  const int rem2 = rem1 - take_av;                           // blocks that come from fresh pages
  const int F = (rem2 + c.bpp - 1) / c.bpp;
  const bool ok = !neg && A0 >= T && cursor + T <= a.out_cap && (rem2 == 0 || tail <= APW) && F <= NEWPG_W &&
                  F <= rls + fls;
  __syncthreads();                                           // compact window visible (or nothing touched yet)
  if (!ok) return 0;
  const int from_rl = F < rls ? F : rls;                      // alloc_page(): reserved list first, then free list
  const int from_fl = F - from_rl;
  const int last_used = F ? rem2 - (F - 1) * c.bpp : 0;       // blocks taken from the last fresh page
  const int64_t nev0 = sc[SC_NEV];
  // ---- emit ids (all warps): warp-strided over stream positions ----
  const int64_t rbh = sc[SC_RB_HEAD];
  for (int j = tid; j < T; j += NT) {
    int64_t id;
    if (j < take_rb) {
      id = c.rb[(int)((rbh + j) & c.rb_mask)];
    } else {
      int j1 = j - take_rb;
      if (j1 < take_av) {
        int lo = 0, hi = nlive - 1;                           // last live entry with ccum <= j1
        while (lo < hi) {
          int mid = (lo + hi + 1) >> 1;
          if (s.ccum[mid] <= j1) lo = mid; else hi = mid - 1;
        }
        int pg = s.cpage[lo];
        int pos = s.chead[lo] + (j1 - s.ccum[lo]);
        if (pos >= c.bpp) pos -= c.bpp;
        id = (int64_t)pg * c.bpp + c.ring[(size_t)pg * c.bpp + pos];
      } else {
        int j2 = j1 - take_av;
        int q = j2 / c.bpp;
        id = (int64_t)s.newpg[q] * c.bpp + (j2 - q * c.bpp);
      }
      atomicOr(&c.bitmap[id >> 5], 1u << (id & 31));
    }
    a.out_ids[cursor + j] = id;
  }
  // ---- per-request results ----
  if (tid < R) {
    kvb_res r{};
    r.status = KVB_ST_OK;
    r.n = n_i;
    r.off = cursor + off_i;
    r.avail = A0 - off_i - n_i;
    a.res[op0 + tid] = r;
  }
  // ---- page state: consumed / partly consumed avail pages ----
  int n_full_av = 0;   // uniform: live entries entirely consumed
  {
    int lo = 0, hi = nlive;                                   // first live entry NOT entirely consumed
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (s.ccum[mid] + s.ccnt[mid] <= take_av) lo = mid + 1; else hi = mid;
    }
    n_full_av = lo;
  }
  if (tid < nlive) {
    int pg = s.cpage[tid], cn = s.ccnt[tid], cm = s.ccum[tid];
    if (tid < n_full_av) {
      int h = s.chead[tid] + cn;
      c.pg_state[pg] = PG_FULL;
      c.pg_cnt[pg] = 0;
      c.pg_head[pg] = h >= c.bpp ? h - c.bpp : h;
    } else if (cm < take_av) {                                 // the one page left partly consumed
      int used = take_av - cm;
      int h = s.chead[tid] + used;
      c.pg_cnt[pg] = cn - used;
      c.pg_head[pg] = h >= c.bpp ? h - c.bpp : h;
    }
  }
  // ---- fresh pages: one warp per page ----
  for (int q = w; q < F; q += NW) {
    int pg = s.newpg[q];
    bool partial = (q == F - 1) && last_used < c.bpp;
    if (lane == 0) {
      c.pg_state[pg] = partial ? PG_AVAIL : PG_FULL;
      c.pg_cnt[pg] = partial ? c.bpp - last_used : 0;
      c.pg_head[pg] = partial ? last_used : 0;
      if (q >= from_rl) {                                      // came off the free list: needs a mapping
        int64_t k = nev0 + (q - from_rl);
        if (k < c.ev_cap) c.ev[k] = (int64_t)pg * 2;
        c.pg_mapped[pg] = 1;
      }
      if (partial) { c.ap[0] = pg; c.pg_pos[pg] = 0; }         // the stack is empty when fresh pages are taken
    }
    if (partial) {                                             // InternalPage::init: FIFO = identity
      uint32_t *r = c.ring + (size_t)pg * c.bpp;
      for (int t = lane; t < c.bpp; t += 32) r[t] = (uint32_t)t;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const bool fresh_partial = F > 0 && last_used < c.bpp;
    sc[SC_RB_HEAD] = (rbh + take_rb) & c.rb_mask;
    sc[SC_NRB] -= take_rb;
    sc[SC_NAB] += (int64_t)F * c.bpp - rem2 - take_av;
    sc[SC_NAVAILPG] += (fresh_partial ? 1 : 0) - n_full_av;
    sc[SC_NFULLPG] += n_full_av + F - (fresh_partial ? 1 : 0);
    sc[SC_NFREE] -= F;
    sc[SC_RL_HEAD] = (sc[SC_RL_HEAD] + from_rl) & c.q_mask;
    sc[SC_RL_SIZE] -= from_rl;
    sc[SC_FL_HEAD] = (sc[SC_FL_HEAD] + from_fl) & c.q_mask;
    sc[SC_FL_SIZE] -= from_fl;
    sc[SC_NEV] += from_fl;
    sc[SC_NMAPPED] += from_fl;
    sc[SC_PHYS] -= from_fl;
    sc[SC_OUT_CURSOR] = cursor + T;
    int new_tail;
    if (F > 0) new_tail = fresh_partial ? 1 : 0;
    else if (n_full_av < nlive) new_tail = s.csidx[n_full_av] + 1;   // newest surviving entry
    else new_tail = tail > APW ? tail - APW : 0;                     // whole window consumed (or empty)
    sc[SC_AP_TAIL] = new_tail;
  }
  __syncthreads();
  return 1;
}

// alloc_run_fast planned at WARP scope: runs of at most 128 requests whose blocks come from the reserved blocks, the 32
// newest entries of the avail stack and at most 32 fresh pages — a decode step's "one block each" run, a single
// alloc(n).  Same stream rule, same results.  Warp 0 PLANS with shuffles and votes (the block-wide form pays three
// block scans for it) and leaves the plan in shared memory; after ONE barrier the whole CTA carries it out: all NT
// threads emit ids and results, warp 1 retires the consumed avail pages, the warps share the fresh pages, lane 0
// of the last warp updates the counters.  (The first version also emitted on warp 0: T / 32 dependent passes over the id stream, 64
// of them for alloc(2048).)  Returns 1 if the run was handled, 0 if it needs the wider windows of alloc_run_fast
// (nothing touched).  Uniform.
__device__ int alloc_run_warp(const DevState &c, Smem &s, const LaunchArgs &a, int chunk_base, int op0, int run_end) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int64_t *sc = s.sc;
  const int R = run_end - op0;
  constexpr int RQ = 4;     // requests per lane: runs of up to 128 requests
  if (!c.exact || sc[SC_INSHRINK] || R > 32 * RQ) return 0;
  WarpPlan &pl = s.wplan;
  if (tid < 32) {
    const unsigned FULL = 0xffffffffu;
    const int tail = (int)sc[SC_AP_TAIL];
    const int rls = (int)sc[SC_RL_SIZE], fls = (int)sc[SC_FL_SIZE];
    int n_q[RQ];              // lane l holds requests [RQ * l, RQ * l + RQ) of the run
    int n_i = 0;
    bool odd = false;
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const int idx = lane * RQ + q;
      n_q[q] = idx < R ? s.ops[op0 - chunk_base + idx].n : 0;
      odd |= n_q[q] < 0 || n_q[q] > (1 << 21);
      n_i += odd ? 0 : n_q[q];
    }
    // the 32 newest stack entries and the next 32 pages alloc_page() would hand out
    int e = -1, ecnt = 0, ehead = 0;
    {
      const int i = tail - 1 - lane;
      e = (i >= 0) ? c.ap[i] : -1;
      if (e >= 0) { ecnt = c.pg_cnt[e]; ehead = c.pg_head[e]; }
    }
    int npg = -1;
    if (lane < rls) npg = c.rl[(int)((sc[SC_RL_HEAD] + lane) & c.q_mask)];
    else if (lane - rls < fls) npg = c.fl[(int)((sc[SC_FL_HEAD] + (lane - rls)) & c.q_mask)];
    const bool neg = __any_sync(FULL, odd);
    // exclusive scans over the warp
    int off_i = n_i, cum = e >= 0 ? ecnt : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(FULL, off_i, d), z = __shfl_up_sync(FULL, cum, d);
      if (lane >= d) { off_i += y; cum += z; }
    }
    const int T = __shfl_sync(FULL, off_i, 31), sum_av = __shfl_sync(FULL, cum, 31);
    off_i -= n_i;
    cum -= (e >= 0 ? ecnt : 0);
    const unsigned livem = __ballot_sync(FULL, e >= 0);
    const int nlive = __popc(livem), li = __popc(livem & ((1u << lane) - 1u));
    const int64_t A0 = avail_size(c, sc);
    const int64_t cursor = sc[SC_OUT_CURSOR];
    const int nrb = (int)sc[SC_NRB];
    const int take_rb = nrb < T ? nrb : T;
    const int rem1 = T - take_rb;
    const int take_av = rem1 < sum_av ? rem1 : sum_av;
    const int rem2 = rem1 - take_av;                           // blocks that come from fresh pages
    const int F = (rem2 + c.bpp - 1) / c.bpp;
    const bool ok = !neg && T <= 2048 && A0 >= T && cursor + T <= a.out_cap && (rem2 == 0 || tail <= 32) && F <= 32 &&
                    F <= rls + fls;
    if (ok) {
      if (e >= 0) { s.cpage[li] = e; s.ccnt[li] = ecnt; s.chead[li] = ehead; s.csidx[li] = tail - 1 - lane; s.ccum[li] = cum; }
      s.newpg[lane] = npg;
      {
        int off_q = off_i;
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
          const int idx = lane * RQ + q;
          if (idx < R) s.wroff[idx] = off_q;
          off_q += n_q[q];
        }
      }
      if (lane == 0) {
        pl.T = T; pl.take_rb = take_rb; pl.take_av = take_av; pl.rem2 = rem2; pl.nlive = nlive; pl.F = F;
        pl.from_rl = F < rls ? F : rls;                          // alloc_page(): reserved list first, then free list
        pl.last_used = F ? rem2 - (F - 1) * c.bpp : 0;
        pl.tail = tail;
        pl.cursor = cursor; pl.A0 = A0; pl.nev0 = sc[SC_NEV]; pl.rbh = sc[SC_RB_HEAD];
      }
    }
    if (lane == 0) s.misc[MI_E] = ok ? 1 : 0;
  }
  __syncthreads();
  if (!s.misc[MI_E]) { __syncthreads(); return 0; }
  // ---- the whole CTA carries the plan out; nothing below reads sc except the counter thread at the end
  const int T = pl.T, take_rb = pl.take_rb, take_av = pl.take_av, rem2 = pl.rem2, nlive = pl.nlive, F = pl.F;
  const int from_rl = pl.from_rl, from_fl = F - from_rl, last_used = pl.last_used, tail = pl.tail;
  const int64_t cursor = pl.cursor, A0 = pl.A0, nev0 = pl.nev0, rbh = pl.rbh;
  for (int j = tid; j < T; j += NT) {
    int64_t id;
    if (j < take_rb) {
      id = c.rb[(int)((rbh + j) & c.rb_mask)];
    } else {
      const int j1 = j - take_rb;
      if (j1 < take_av) {
        int lo = 0, hi = nlive - 1;                         // last live entry with ccum <= j1
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (s.ccum[mid] <= j1) lo = mid; else hi = mid - 1;
        }
        const int pg = s.cpage[lo];
        int pos = s.chead[lo] + (j1 - s.ccum[lo]);
        if (pos >= c.bpp) pos -= c.bpp;
        id = (int64_t)pg * c.bpp + c.ring[(size_t)pg * c.bpp + pos];
      } else {
        const int j2 = j1 - take_av;
        const int q = j2 / c.bpp;
        id = (int64_t)s.newpg[q] * c.bpp + (j2 - q * c.bpp);
      }
      atomicOr(&c.bitmap[id >> 5], 1u << (id & 31));
    }
    a.out_ids[cursor + j] = id;
  }
  if (tid < R) {
    const int nq = s.ops[op0 - chunk_base + tid].n, off_q = s.wroff[tid];
    kvb_res r{};
    r.status = KVB_ST_OK;
    r.n = nq;
    r.off = cursor + off_q;
    r.avail = A0 - off_q - nq;
    a.res[op0 + tid] = r;
  }
  // consumed / partly consumed avail pages
  int n_full_av = 0;
  if (w == 1 || w == NW - 1) {                              // needed by the page-update warp and by the counter thread
    int lo = 0, hi = nlive;                                 // first live entry NOT entirely consumed
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s.ccum[mid] + s.ccnt[mid] <= take_av) lo = mid + 1; else hi = mid;
    }
    n_full_av = lo;
  }
  if (w == (NW > 1 ? 1 : 0) && lane < nlive) {
    const int pg = s.cpage[lane], cn = s.ccnt[lane], cm = s.ccum[lane];
    if (lane < n_full_av) {
      const int hh = s.chead[lane] + cn;
      c.pg_state[pg] = PG_FULL;
      c.pg_cnt[pg] = 0;
      c.pg_head[pg] = hh >= c.bpp ? hh - c.bpp : hh;
    } else if (cm < take_av) {                               // the one page left partly consumed
      const int used = take_av - cm;
      const int hh = s.chead[lane] + used;
      c.pg_cnt[pg] = cn - used;
      c.pg_head[pg] = hh >= c.bpp ? hh - c.bpp : hh;
    }
  }
  // fresh pages: one warp each
  for (int q = w; q < F; q += NW) {
    const int pg = s.newpg[q];
    const bool partial = (q == F - 1) && last_used < c.bpp;
    if (lane == 0) {
      c.pg_state[pg] = partial ? PG_AVAIL : PG_FULL;
      c.pg_cnt[pg] = partial ? c.bpp - last_used : 0;
      c.pg_head[pg] = partial ? last_used : 0;
      if (q >= from_rl) {                                    // came off the free list: needs a mapping
        const int64_t k = nev0 + (q - from_rl);
        if (k < c.ev_cap) c.ev[k] = (int64_t)pg * 2;
        c.pg_mapped[pg] = 1;
      }
      if (partial) { c.ap[0] = pg; c.pg_pos[pg] = 0; }       // the stack is empty when fresh pages are taken
    }
    if (partial) {                                           // InternalPage::init: FIFO = identity
      uint32_t *rr = c.ring + (size_t)pg * c.bpp;
      for (int t = lane; t < c.bpp; t += 32) rr[t] = (uint32_t)t;
    }
  }
  if (tid == NT - 32) {        // the counters: lane 0 of the last warp, which has no ids to emit unless T > NT - 32
    const bool fresh_partial = F > 0 && last_used < c.bpp;
    sc[SC_RB_HEAD] = (rbh + take_rb) & c.rb_mask;
    sc[SC_NRB] -= take_rb;
    sc[SC_NAB] += (int64_t)F * c.bpp - rem2 - take_av;
    sc[SC_NAVAILPG] += (fresh_partial ? 1 : 0) - n_full_av;
    sc[SC_NFULLPG] += n_full_av + F - (fresh_partial ? 1 : 0);
    sc[SC_NFREE] -= F;
    sc[SC_RL_HEAD] = (sc[SC_RL_HEAD] + from_rl) & c.q_mask;
    sc[SC_RL_SIZE] -= from_rl;
    sc[SC_FL_HEAD] = (sc[SC_FL_HEAD] + from_fl) & c.q_mask;
    sc[SC_FL_SIZE] -= from_fl;
    sc[SC_NEV] += from_fl;
    sc[SC_NMAPPED] += from_fl;
    sc[SC_PHYS] -= from_fl;
    sc[SC_OUT_CURSOR] = cursor + T;
    int new_tail;
    if (F > 0) new_tail = fresh_partial ? 1 : 0;
    else if (n_full_av < nlive) new_tail = s.csidx[n_full_av] + 1;   // newest surviving entry
    else new_tail = tail > 32 ? tail - 32 : 0;                       // whole window consumed (or empty)
    sc[SC_AP_TAIL] = new_tail;
  }
  __syncthreads();
  return 1;
}


// A run of consecutive ALLOC requests [op0, run_end) (all inside the staged chunk), or one
// TRY_RESERVE request (to_reserved = 1, ids appended to reserved_blocks instead of the output).
__device__ __noinline__ void alloc_run(const DevState &c, Smem &s, const LaunchArgs &a, int chunk_base, int op0, int run_end,
                          int to_reserved) {
  const int tid = threadIdx.x;
  int64_t *sc = s.sc;
  // prefetch the top of the avail stack and the next fresh page ids
  {
    int tail = (int)sc[SC_AP_TAIL];
    if (tid < APW) {
      int i = tail - 1 - tid;
      int e = (i >= 0) ? c.ap[i] : -1;
      s.apw[tid] = e;
      if (e >= 0) { s.apcnt[tid] = c.pg_cnt[e]; s.aphead[tid] = c.pg_head[e]; }
    }
    if (tid >= NT - NEWPG_W) {
      int j = tid - (NT - NEWPG_W);
      int rls = (int)sc[SC_RL_SIZE], fls = (int)sc[SC_FL_SIZE];
      int v = -1;
      if (j < rls) v = c.rl[(int)((sc[SC_RL_HEAD] + j) & c.q_mask)];
      else if (j - rls < fls) v = c.fl[(int)((sc[SC_FL_HEAD] + (j - rls)) & c.q_mask)];
      s.newpg[j] = v;
    }
    if (tid == 0) {
      AllocSim &m = s.sim;
      m.op = op0;
      m.req_active = 0;
      m.req_rem = 0;
      m.ap_top = tail;
      m.ap_tail0 = tail;
      m.cur_p = -1;
      m.newpg_j = 0;
      m.done = 0;
      m.failed = 0;
      if (to_reserved) {
        // ids go to the back of the reserved_blocks ring
        m.out_pos = sc[SC_RB_HEAD] + sc[SC_NRB];
      } else {
        m.out_pos = sc[SC_OUT_CURSOR];
      }
    }
  }
  __syncthreads();
  // NOTE on newpg: entry j of the window is the j-th page alloc_page() will hand out in this run
  // (all of the reserved list, then the free list): nothing pushes to either list during a run.
  while (true) {
    if (tid == 0) alloc_walk(c, s, a, chunk_base, run_end, to_reserved);
    __syncthreads();
    if (to_reserved) alloc_emit(c, s, a, c.rb, (int64_t)c.rb_mask);
    else alloc_emit(c, s, a, a.out_ids, 0);
    int done = s.sim.done;
    __syncthreads();
    if (done) break;
  }
  if (tid == 0) {
    if (to_reserved) {
      // try_to_reserve: reserved_blocks.extend(result) (:371); alloc() consumed the old front first
      const kvb_res &r = a.res[op0];
      if (r.status == KVB_ST_OK) sc[SC_NRB] += r.n;
    } else {
      sc[SC_OUT_CURSOR] = s.sim.out_pos;
    }
  }
  __syncthreads();
}

// ==================================================================================== free
// Compact the avail stack (drop tombstones, keep order) when it has grown past max_pages.
__device__ void ap_compact(const DevState &c, Smem &s) {
  int tail = (int)s.sc[SC_AP_TAIL];
  if (tail <= c.max_pages) return;   // uniform
  int base = 0;
  for (int i0 = 0; i0 < tail; i0 += NT) {
    int i = i0 + threadIdx.x;
    int e = (i < tail) ? c.ap[i] : -1;
    int tot;
    int r = block_rank(e >= 0, s.wtot, &tot);
    // all reads of this chunk happened before block_rank's barriers; writes land at or below i
    if (e >= 0) { c.ap[base + r] = e; c.pg_pos[e] = base + r; }
    base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) s.sc[SC_AP_TAIL] = base;
  __syncthreads();
}

// free(ids) for the n ids already gathered in c.fids (kv_cache_manager.py:307-360).
// `n_decl` is len(indices) as the reference sees it (sizes the unordered_map reserve()).
// group_only: no state change; writes the grouping (pages, counts, ids) for kvb_group_indices_by_page.
__device__ void free_core(const DevState &c, Smem &s, const LaunchArgs &a, int n, int group_only) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int64_t *sc = s.sc;
  if (n <= 0) return;
  if (!group_only) ap_compact(c, s);
  if (tid == 0) { s.misc[MI_BAD] = 0; s.misc[MI_UNK] = 0; }
  for (int i = tid; i < HT_D_CAP; i += NT) s.ht_cnt[i] = 0;
  __syncthreads();

  // ---- F0: classify ids: page arithmetic + bitmap test-and-clear (a block must be allocated) -----
  int bad = 0;
  for (int j = tid; j < n; j += NT) {
    int64_t id = c.fids[j];
    int32_t pg = INT_MIN;
    if (id >= 0 && id < c.num_blocks) {
      int64_t p64 = blk_page(c, id);
      if (p64 <= c.max_pages) {   // page max_pages can be named (tail of a non-page-aligned pool); it is never in use
        int p = (int)p64;
        bool ok = true;
        if (!c.exact && !group_only) {   // a block straddling the page edge is never handed out
          int64_t st = pg_start(c, p);
          ok = id >= st && id < st + pg_nblk(c, p);
        }
        if (ok && !group_only) {
          uint32_t bit = 1u << (id & 31);
          uint32_t old = atomicAnd(&c.bitmap[id >> 5], ~bit);
          ok = (old & bit) != 0;
        }
        pg = ok ? p : ~p;   // ~p: page takes part in the dict order, the id is ignored
      }
    }
    if (pg < 0) bad++;
    c.fpage[j] = pg;
  }
  if (bad) atomicAdd(&s.misc[MI_BAD], bad);
  __syncthreads();
  KVB_PROF(4);

  // ---- F1: first index naming each page (only run heads can be first) ---------------------------
  for (int j = tid; j < n; j += NT) {
    int32_t pg = c.fpage[j];
    if (pg == INT_MIN) continue;
    int p = pg < 0 ? ~pg : pg;
    bool head = true;
    if (j > 0) {
      int32_t q = c.fpage[j - 1];
      if (q != INT_MIN && (q < 0 ? ~q : q) == p) head = false;
    }
    if (head) atomicMin(&c.pg_first[p], j);
  }
  __syncthreads();

  // ---- F2: distinct pages in first-appearance order (ballot/popc compaction) -------------------
  int D = 0;
  for (int j0 = 0; j0 < n; j0 += NT) {
    int j = j0 + tid;
    int p = -1;
    bool first = false;
    if (j < n) {
      int32_t pg = c.fpage[j];
      if (pg != INT_MIN) {
        p = pg < 0 ? ~pg : pg;
        first = (c.pg_first[p] == j);
      }
    }
    int tot;
    int r = block_rank(first, s.wtot, &tot);
    if (first) {
      int d = D + r;
      c.pg_slot[p] = d;
      if (d < HT_D_CAP) s.ht_key[d] = p;
      c.g_dist[d] = p;
      c.g_cnt[d] = 0;
    }
    D += tot;
  }
  __syncthreads();
  KVB_PROF(5);

  // ---- F3: dict order (one lane) || per-page FIFO ranks (one warp) --------------------------------
  // groups.reserve(n / bpp + 1): csrc/page_allocator.cpp:477
  int nb0 = next_bkt(c, s.primes32, (int64_t)n / c.bpp + 1);
  const int nb_final = D > 32 ? ht_final_buckets(c, s.primes32, D, nb0) : nb0;
  const bool use_g = (D > HT_D_CAP) || (D > 32 && nb_final > HT_NB_CAP);
  const bool multi = D > 32 && D <= 32 * HT_WK && nb_final <= HT_D_CAP;   // tab_cnt = ht_next holds HT_D_CAP entries
  int32_t *keys = use_g ? c.g_dist : s.ht_key;
  int32_t *ord = use_g ? c.g_ord : s.ht_ord;
  int32_t *cnt = use_g ? c.g_cnt : s.ht_cnt;
  if (w == 0) {
    if (D <= 32) {
#ifdef KVB_PROFILE
      ht_order_warp(c, s.primes32, s.ht_key, D, nb0, s.ht_ord, s.ht_next, &s.prof_t);
#else
      ht_order_warp(c, s.primes32, s.ht_key, D, nb0, s.ht_ord, s.ht_next);   // registers + warp votes
#endif
    } else if (multi) {
      ht_order_hybrid(c, s.primes32, s.ht_key, D, nb0, s.ht_ord, s.ht_bkt, s.ht_next, s.tmp64);
    } else if (!use_g) {
      for (int b = lane; b < nb0; b += 32) s.ht_bkt[b] = -1;
      __syncwarp();
      if (lane == 0) ht_order_seq(c, s.primes32, s.ht_key, D, nb0, s.ht_next, s.ht_bkt, s.ht_ord);
    } else {
      for (int b = lane; b < nb0; b += 32) c.g_bkt[b] = -1;
      __syncwarp();
      if (lane == 0) ht_order_seq(c, s.primes32, c.g_dist, D, nb0, c.g_next, c.g_bkt, c.g_ord);
    }
    KVB_PROF(11);
  } else if (w == 1) {
    for (int j0 = 0; j0 < n; j0 += 32) {
      int j = j0 + lane;
      int32_t pg = (j < n) ? c.fpage[j] : INT_MIN;
      int slot = (pg >= 0) ? c.pg_slot[pg] : -1 - lane;   // distinct negatives: no grouping
      unsigned mask = __match_any_sync(0xffffffffu, slot);
      int r = __popc(mask & ((1u << lane) - 1u));
      int basecnt = 0;
      if (slot >= 0) basecnt = cnt[slot];
      __syncwarp();
      if (slot >= 0 && r == 0) cnt[slot] = basecnt + __popc(mask);
      __syncwarp();
      if (slot >= 0) {
        int rank = basecnt + r;
        if (group_only) {
          c.fpage[j] = rank;   // reuse as rank store; the page is recoverable from the id
        } else {
          int p = pg;
          int pos = c.pg_head[p] + c.pg_cnt[p] + rank;   // InternalPage::free_batch appends (:69-73)
          pos %= c.bpp;
          c.ring[(size_t)p * c.bpp + pos] = (uint32_t)(c.fids[j] - pg_start(c, p));
        }
      } else if (group_only && j < n) {
        c.fpage[j] = -1;
      }
    }
  }
  __syncthreads();
  KVB_PROF(6);

  if (group_only) {
    // pages in dict order, their id counts, ids regrouped by page
    int run = 0;
    for (int d0 = 0; d0 < D; d0 += NT) {
      int d = d0 + tid;
      int slot = (d < D) ? ord[d] : 0;
      int cn = (d < D) ? cnt[slot] : 0;
      int tot;
      int ex = block_exscan(cn, s.wtot, &tot);
      if (d < D) {
        a.grp_pages[d] = keys[slot];
        a.grp_counts[d] = cn;
        c.g_ptf[slot] = run + ex;   // output offset of the slot's first id
      }
      run += tot;
    }
    __syncthreads();
    for (int j = tid; j < n; j += NT) {
      int rank = c.fpage[j];
      if (rank < 0) continue;
      int64_t id = c.fids[j];
      int p = (int)blk_page(c, id);
      a.grp_ids[c.g_ptf[c.pg_slot[p]] + rank] = id;
    }
    __syncthreads();
    for (int d = tid; d < D; d += NT) c.pg_first[keys[d]] = INT_MAX;
    if (tid == 0) s.misc[MI_D] = D;
    __syncthreads();
    return;
  }

  // ---- F4: per-page update in dict order (:322-349) -----------------------------------------------
  if (tid == 0) { s.acc[0] = 0; s.acc[1] = 0; s.acc[2] = 0; s.acc[3] = 0; }
  __syncthreads();
  int tail = (int)sc[SC_AP_TAIL];
  int E = 0;
  for (int d0 = 0; d0 < D; d0 += NT) {
    int d = d0 + tid;
    bool live = false, empty = false;
    int p = -1, add = 0, st = PG_NONE, newc = 0;
    if (d < D) {
      int slot = ord[d];
      p = keys[slot];
      add = cnt[slot];
      st = c.pg_state[p];
      c.pg_first[p] = INT_MAX;   // reset scratch
      if (st == PG_NONE) {
        atomicAdd(&s.misc[MI_UNK], 1);   // page not in avail_pages / full_pages: logged and skipped
      } else if (add > 0) {
        live = true;
        newc = c.pg_cnt[p] + add;
        empty = (newc == pg_nblk(c, p));   // InternalPage::empty (:75-77)
      }
    }
    int totS, totE;
    int rs = block_rank(live && !empty, s.wtot, &totS);
    int re = block_rank(live && empty, s.wtot2, &totE);
    int d_nab = 0, d_av = 0, d_fu = 0;
    if (live) {
      if (st == PG_AVAIL) c.ap[c.pg_pos[p]] = -1;   // avail_pages.pop(page_id)
      long long dnab = add;
      if (empty) {
        c.g_ptf[E + re] = p;
        c.pg_state[p] = PG_NONE;
        c.pg_cnt[p] = 0;
        dnab -= newc;
      } else {
        c.ap[tail + rs] = p;            // avail_pages[page_id] = page (newest)
        c.pg_pos[p] = tail + rs;
        c.pg_state[p] = PG_AVAIL;
        c.pg_cnt[p] = newc;
      }
      d_nab = (int)dnab;
      d_av = (empty ? 0 : 1) - (st == PG_AVAIL ? 1 : 0);
      d_fu = -(st == PG_FULL ? 1 : 0);
    }
    // counter deltas: one warp reduction + one shared atomic per warp instead of three contended atomics per page
    {
      const int s_nab = __reduce_add_sync(0xffffffffu, d_nab);
      const int s_av = __reduce_add_sync(0xffffffffu, d_av);
      const int s_fu = __reduce_add_sync(0xffffffffu, d_fu);
      if (lane == 0) {
        if (s_nab) atomicAdd((unsigned long long *)&s.acc[0], (unsigned long long)(long long)s_nab);
        if (s_av) atomicAdd((unsigned long long *)&s.acc[1], (unsigned long long)(long long)s_av);
        if (s_fu) atomicAdd((unsigned long long *)&s.acc[2], (unsigned long long)(long long)s_fu);
      }
    }
    tail += totS;
    E += totE;
  }
  __syncthreads();
  KVB_PROF(7);

  // ---- F5: PageAllocator::free_pages(pages_to_free) (page_allocator.cpp:264-310) --------------------
  int rl_size = (int)sc[SC_RL_SIZE], fl_size = (int)sc[SC_FL_SIZE];
  int room = c.max_reserved - rl_size;
  int k = room > 0 ? (room < E ? room : E) : 0;
  int64_t nev0 = sc[SC_NEV];
  for (int i = tid; i < E; i += NT) {
    int p = c.g_ptf[i];
    if (i < k) {
      c.rl[(int)((sc[SC_RL_HEAD] + rl_size + i) & c.q_mask)] = p;   // stays mapped
    } else {
      int u = i - k;
      c.fl[(int)((sc[SC_FL_HEAD] + fl_size + u) & c.q_mask)] = p;
      if (nev0 + u < c.ev_cap) c.ev[nev0 + u] = (int64_t)p * 2 + 1;
      c.pg_mapped[p] = 0;
    }
  }
  __syncthreads();
  if (tid == 0) {
    sc[SC_NAB] += s.acc[0];
    sc[SC_NAVAILPG] += s.acc[1];
    sc[SC_NFULLPG] += s.acc[2];
    sc[SC_AP_TAIL] = tail;
    sc[SC_NFREE] += E;
    sc[SC_RL_SIZE] += k;
    sc[SC_FL_SIZE] += E - k;
    sc[SC_NEV] += E - k;
    sc[SC_NMAPPED] -= E - k;
    sc[SC_PHYS] += E - k;
    sc[SC_BAD_FREES] += s.misc[MI_BAD];
    sc[SC_UNKNOWN_FREES] += s.misc[MI_UNK];
    // pending shrink completes once usage is under the target (:354-360); resize result ignored
    if (sc[SC_INSHRINK] && alloced_blocks(c, sc) <= sc[SC_TARGET]) {
      pa_resize(c, sc, sc[SC_TARGET] * c.B, nullptr);
      sc[SC_INSHRINK] = 0;
      sc[SC_TARGET] = -1;
    }
  }
  __syncthreads();
  KVB_PROF(8);
}

// free(ids) for calls of at most NT ids — one id per thread, the call's scratch in SHARED memory.
// free_core keeps ids, pages, first-appearance and slot tables in global scratch arrays: every phase boundary is an
// L2 round trip plus a block barrier (the phase profile charged 12 k cycles to a 16-id free).  Here the only global
// round trips on the critical path are the bitmap test-and-clear and one parallel load of the touched pages' state:
//   1. page of my id; distinct pages and their first index through a 1 024-entry shared hash table (atomicCAS claim +
//      atomicMin of the thread index) -> slots in first-appearance order (ballot/popc rank), page state -> shared;
//   2. warp 0 replays the unordered_map order of the slots (ht_order_warp / ht_order_hybrid) while warp 1 ranks the
//      ids inside their pages (__match_any_sync) and appends them to the page FIFOs;
//   3. per page, in dict order, the same update rule as free_core's F4 with every input in shared memory; F5 as there.
// Results are free_core's, bit for bit.  Returns 0 (uniform, NO state touched) when the call names more than
// HS_D_CAP distinct pages: the caller then takes free_core.
constexpr int HS_D_CAP = 32 * HT_WK;
__device__ int free_small(const DevState &c, Smem &s, const LaunchArgs &a, int n, int64_t my_id) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int64_t *sc = s.sc;
  int32_t *Hkey = s.ht_bkt, *Hval = s.ht_bkt + 1024, *Hslot = s.ht_bkt + 2048;
  int32_t *scr = reinterpret_cast<int32_t *>(s.segs);        // alloc-run scratch, idle during a free
  int32_t *sl = scr, *pst = scr + 512, *pcnt = scr + 1024, *phead = scr + 1536, *ppos = scr + 2048;
  int64_t *ids64 = reinterpret_cast<int64_t *>(scr + 2560);
  int32_t *efree = scr + 3584;                               // pages_to_free, dict order (<= HS_D_CAP entries)
  static_assert(sizeof(s.segs) >= (3584 + 512) * 4 && HT_NB_CAP + 1 >= 3072 && NT <= 512 && HS_D_CAP <= 512, "free_small scratch");
  if (n <= 0) return 1;
  ap_compact(c, s);
  for (int i = tid; i < 1024; i += NT) { Hkey[i] = -1; Hval[i] = INT_MAX; }
  if (tid == 0) { s.misc[MI_BAD] = 0; s.misc[MI_UNK] = 0; s.acc[0] = 0; s.acc[1] = 0; s.acc[2] = 0; s.acc[3] = 0; }
  bool part = false;      // my id names a page that takes part in the dict order
  int p = -1;
  if (tid < n && my_id >= 0 && my_id < c.num_blocks) {
    const int64_t p64 = blk_page(c, my_id);
    if (p64 <= c.max_pages) { p = (int)p64; part = true; }
  }
  // Issued now, consumed after the classification below: the bitmap test-and-clear (a block must be allocated) and the
  // state of my id's page travel to L2 and back while the CTA builds its hash table.  (A call that turns out to name
  // too many pages for this path puts its bits back, so "nothing touched" still holds for the caller.)
  bool ok = part;
  if (ok && !c.exact) {   // a block straddling the page edge is never handed out
    const int64_t st0 = pg_start(c, p);
    ok = my_id >= st0 && my_id < st0 + pg_nblk(c, p);
  }
  const uint32_t bit = ok ? 1u << (my_id & 31) : 0u;
  uint32_t old_word = 0;
  if (ok) old_word = atomicAnd(&c.bitmap[my_id >> 5], ~bit);
  int e_st = 0, e_cnt = 0, e_head = 0, e_pos = 0;
  if (part) { e_st = c.pg_state[p]; e_cnt = c.pg_cnt[p]; e_head = c.pg_head[p]; e_pos = c.pg_pos[p]; }
  __syncthreads();
  KVB_PROF(14);
  unsigned h = 0;
  if (part) {
    h = ((unsigned)p * 0x9E3779B1u) >> 22;
    while (true) {
      const int prev = atomicCAS(&Hkey[h], -1, p);
      if (prev == -1 || prev == p) break;
      h = (h + 1) & 1023u;
    }
    atomicMin(&Hval[h], tid);
  }
  __syncthreads();
  KVB_PROF(15);
  const bool first = part && Hval[h] == tid;
  int D;
  const int r = block_rank_1b(first, s.wsite[WS_FS_D], &D);
  if (D > HS_D_CAP) {
    if (ok && (old_word & bit)) atomicOr(&c.bitmap[my_id >> 5], bit);
    __syncthreads();
    return 0;
  }
  KVB_PROF(4);
  // ---- from here on the state changes
  ok = ok && (old_word & bit) != 0;
  if (first) {
    Hslot[h] = r;
    s.ht_key[r] = p;
    s.ht_cnt[r] = 0;
    pst[r] = e_st;
    pcnt[r] = e_cnt;
    phead[r] = e_head;
    ppos[r] = e_pos;
  }
  {
    const unsigned bm = __ballot_sync(0xffffffffu, tid < n && !ok);
    if (lane == 0 && bm) atomicAdd(&s.misc[MI_BAD], __popc(bm));
  }
  __syncthreads();
  if (tid < n) { sl[tid] = ok ? Hslot[h] : -1; ids64[tid] = my_id; }
  __syncthreads();
  KVB_PROF(5);
  // ---- dict order (warp 0) || FIFO ranks and ring appends (warp 1); groups.reserve(n / bpp + 1), page_allocator.cpp:477
  const int nb0 = next_bkt_small(s.primes32, n / c.bpp + 1);      // n <= NT
  // block-wide later phases: bucket counts stay <= 503 (D <= 256 and a rehash at most doubles a count below D), which
  // is what the packed per-bucket words have room for
  const bool blockwide = D > 32 && nb0 <= 512 && !(a.tune & 1);
  if (w == 0) {
    if (D <= 32) {
      ht_order_warp(c, s.primes32, s.ht_key, D, nb0, s.ht_ord, s.ht_next);
    } else if (!blockwide) {
      ht_order_hybrid(c, s.primes32, s.ht_key, D, nb0, s.ht_ord, s.ht_bkt, s.ht_next, s.tmp64);
    } else {
      // the phases that hold <= 32 keys here, on one warp; the rest block-wide after the barrier below
      int nb1 = nb0, placed1, cs;
      unsigned ck;
      ht_order_prefix(c, s.primes32, s.ht_key, D, &nb1, &placed1, &cs, &ck, s.ht_next);
      if (lane < placed1) { s.ht_bkt[1280 + lane] = cs; s.ht_bkt[1536 + lane] = (int)ck; }
      if (lane == 0) { s.misc[MI_NB0] = nb1; s.misc[MI_K] = placed1; }
    }
    KVB_PROF(11);
  } else if (w == 1) {
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      const int slot = (j < n) ? sl[j] : -1;
      const unsigned mask = __match_any_sync(0xffffffffu, slot >= 0 ? slot : -1 - lane);
      const int rk = __popc(mask & ((1u << lane) - 1u));
      const int basecnt = slot >= 0 ? s.ht_cnt[slot] : 0;
      __syncwarp();
      if (slot >= 0 && rk == 0) s.ht_cnt[slot] = basecnt + __popc(mask);
      __syncwarp();
      if (slot >= 0) {
        const int pp = s.ht_key[slot];
        int pos = phead[slot] + pcnt[slot] + basecnt + rk;   // InternalPage::free_batch appends (:69-73)
        pos %= c.bpp;
        c.ring[(size_t)pp * c.bpp + pos] = (uint32_t)(ids64[j] - pg_start(c, pp));
      }
    }
  }
  __syncthreads();
  if (blockwide) ht_order_block(c, s, s.ht_key, D, s.misc[MI_NB0], s.misc[MI_K], s.ht_ord);
  KVB_PROF(6);
  // ---- per-page update in dict order (:322-349): free_core's F4 with shared-memory inputs
  int tail = (int)sc[SC_AP_TAIL];
  int E = 0;
  {
    bool live = false, empty = false;
    int pp = -1, add = 0, st = PG_NONE, newc = 0, slot = 0;
    if (tid < D) {
      slot = s.ht_ord[tid];
      pp = s.ht_key[slot];
      add = s.ht_cnt[slot];
      st = pst[slot];
      if (st == PG_NONE) {
        atomicAdd(&s.misc[MI_UNK], 1);   // page not in avail_pages / full_pages: logged and skipped
      } else if (add > 0) {
        live = true;
        newc = pcnt[slot] + add;
        empty = (newc == pg_nblk(c, pp));   // InternalPage::empty (:75-77)
      }
    }
    int totS, totE, rs, re;
    block_rank2_1b(live && !empty, live && empty, s.wsite[WS_FS_F4], &rs, &totS, &re, &totE);
    int d_nab = 0, d_av = 0, d_fu = 0;
    if (live) {
      if (st == PG_AVAIL) c.ap[ppos[slot]] = -1;   // avail_pages.pop(page_id)
      long long dnab = add;
      if (empty) {
        efree[re] = pp;
        c.pg_state[pp] = PG_NONE;
        c.pg_cnt[pp] = 0;
        dnab -= newc;
      } else {
        c.ap[tail + rs] = pp;            // avail_pages[page_id] = page (newest)
        c.pg_pos[pp] = tail + rs;
        c.pg_state[pp] = PG_AVAIL;
        c.pg_cnt[pp] = newc;
      }
      d_nab = (int)dnab;
      d_av = (empty ? 0 : 1) - (st == PG_AVAIL ? 1 : 0);
      d_fu = -(st == PG_FULL ? 1 : 0);
    }
    const int s_nab = __reduce_add_sync(0xffffffffu, d_nab);
    const int s_av = __reduce_add_sync(0xffffffffu, d_av);
    const int s_fu = __reduce_add_sync(0xffffffffu, d_fu);
    if (lane == 0) {
      if (s_nab) atomicAdd((unsigned long long *)&s.acc[0], (unsigned long long)(long long)s_nab);
      if (s_av) atomicAdd((unsigned long long *)&s.acc[1], (unsigned long long)(long long)s_av);
      if (s_fu) atomicAdd((unsigned long long *)&s.acc[2], (unsigned long long)(long long)s_fu);
    }
    tail += totS;
    E = totE;
  }
  __syncthreads();
  KVB_PROF(7);
  // ---- PageAllocator::free_pages(pages_to_free) (page_allocator.cpp:264-310): free_core's F5
  const int rl_size = (int)sc[SC_RL_SIZE], fl_size = (int)sc[SC_FL_SIZE];
  const int room = c.max_reserved - rl_size;
  const int k = room > 0 ? (room < E ? room : E) : 0;
  const int64_t nev0 = sc[SC_NEV];
  for (int i = tid; i < E; i += NT) {
    const int pp = efree[i];
    if (i < k) {
      c.rl[(int)((sc[SC_RL_HEAD] + rl_size + i) & c.q_mask)] = pp;   // stays mapped
    } else {
      const int u = i - k;
      c.fl[(int)((sc[SC_FL_HEAD] + fl_size + u) & c.q_mask)] = pp;
      if (nev0 + u < c.ev_cap) c.ev[nev0 + u] = (int64_t)pp * 2 + 1;
      c.pg_mapped[pp] = 0;
    }
  }
  __syncthreads();
  if (tid == 0) {
    sc[SC_NAB] += s.acc[0];
    sc[SC_NAVAILPG] += s.acc[1];
    sc[SC_NFULLPG] += s.acc[2];
    sc[SC_AP_TAIL] = tail;
    sc[SC_NFREE] += E;
    sc[SC_RL_SIZE] += k;
    sc[SC_FL_SIZE] += E - k;
    sc[SC_NEV] += E - k;
    sc[SC_NMAPPED] -= E - k;
    sc[SC_PHYS] += E - k;
    sc[SC_BAD_FREES] += s.misc[MI_BAD];
    sc[SC_UNKNOWN_FREES] += s.misc[MI_UNK];
    if (sc[SC_INSHRINK] && alloced_blocks(c, sc) <= sc[SC_TARGET]) {   // pending shrink (:354-360)
      pa_resize(c, sc, sc[SC_TARGET] * c.B, nullptr);
      sc[SC_INSHRINK] = 0;
      sc[SC_TARGET] = -1;
    }
  }
  __syncthreads();
  KVB_PROF(8);
  return 1;
}

// Gather the ids of a FREE_REF request into c.fids: concat of result(op)[lo:hi] of earlier requests.
// Returns the id count (uniform).
__device__ int gather_refs(const DevState &c, Smem &s, const LaunchArgs &a, int64_t ref_off, int nrefs) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int total = 0;
  int64_t *r_src = s.tmp64;     // staged (source offset, length, destination) of one chunk of refs
  int32_t *r_len = s.ht_next;   // borrowed shared arrays (free_core has not started)
  int32_t *r_dst = s.ht_ord;
  for (int r0 = 0; r0 < nrefs; r0 += NT) {
    int r = r0 + tid;
    int len = 0;
    int64_t src = 0;
    if (r < nrefs) {
      kvb_ref rf = a.refs[ref_off + r];
      kvb_res rs = a.res[rf.op];
      int n = rs.status == KVB_ST_OK ? rs.n : 0;
      int lo = rf.lo < 0 ? (rf.lo + n < 0 ? 0 : rf.lo + n) : (rf.lo > n ? n : rf.lo);
      int hi = rf.hi < 0 ? (rf.hi + n < 0 ? 0 : rf.hi + n) : (rf.hi > n ? n : rf.hi);
      len = hi > lo ? hi - lo : 0;
      src = rs.off + lo;
    }
    int tot;
    int ex = block_exscan(len, s.wtot, &tot);
    r_src[tid] = src;
    r_len[tid] = len;
    r_dst[tid] = total + ex;
    __syncthreads();
    int nr = nrefs - r0 < NT ? nrefs - r0 : NT;
    for (int q = w; q < nr; q += NW) {   // one warp copies one ref's ids
      int64_t so = r_src[q];
      int ln = r_len[q], dp = r_dst[q];
      for (int t = lane; t < ln; t += 32)
        if (dp + t < c.f_cap) c.fids[dp + t] = a.out_ids[so + t];
    }
    total += tot;
    __syncthreads();
  }
  return total;
}

// The same gather for calls that fit one id per thread (at most NT refs naming at most NT ids): thread t takes the
// t-th id of the concatenation into a REGISTER (binary search of its ref in the shared offset table) — no staging
// pass through c.fids.  Returns the id count, or -1 (uniform, nothing written) when the call does not fit.
// EARLY = the call is made before pdl_enter(), while the predecessor launch may still be running: results of requests
// below `safe_before` come from launches that had completed before this one could start (a launch lets its successor
// in only after its own wait), anything at or above it makes the call return -1.  Those loads bypass L1 (ld.cg): a
// line that also holds a result the predecessor is still writing must not stay cached on this SM.
template <bool EARLY>
__device__ int gather_refs_reg(const DevState &c, Smem &s, const LaunchArgs &a, int64_t ref_off, int nrefs, int64_t *my_id,
                               int safe_before = 0) {
  const int tid = threadIdx.x;
  *my_id = -1;
  if (nrefs > NT) return -1;
  int64_t *r_src = s.tmp64;
  int32_t *r_dst = s.ht_ord;
  int len = 0;
  int64_t src = 0;
  bool late = false;
  if (tid < nrefs) {
    kvb_ref rf = a.refs[ref_off + tid];
    kvb_res rs;
    if (EARLY) {
      late = rf.op >= safe_before;
      const int4 *q = reinterpret_cast<const int4 *>(a.res + (late ? 0 : rf.op));
      static_assert(sizeof(kvb_res) == 32, "kvb_res is read as two 16-byte words");
      const int4 lo4 = __ldcg(q), hi4 = __ldcg(q + 1);
      memcpy(&rs, &lo4, 16);
      memcpy(reinterpret_cast<char *>(&rs) + 16, &hi4, 16);
    } else {
      rs = a.res[rf.op];
    }
    int n = rs.status == KVB_ST_OK ? rs.n : 0;
    int lo = rf.lo < 0 ? (rf.lo + n < 0 ? 0 : rf.lo + n) : (rf.lo > n ? n : rf.lo);
    int hi = rf.hi < 0 ? (rf.hi + n < 0 ? 0 : rf.hi + n) : (rf.hi > n ? n : rf.hi);
    len = hi > lo ? hi - lo : 0;
    src = rs.off + lo;
  }
  if (EARLY && __syncthreads_or(late)) return -1;
  int tot;
  const int ex = block_exscan_1b(len, s.wsite[WS_GATHER], &tot);
  if (tot > NT) { __syncthreads(); return -1; }
  r_src[tid] = src;
  r_dst[tid] = tid < nrefs ? ex : INT_MAX;     // offsets are non-decreasing; empty refs share their successor's offset
  __syncthreads();
  if (tid < tot) {
    int lo = 0, hi = nrefs - 1;                // last ref whose first id is at or before position tid
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (r_dst[mid] <= tid) lo = mid; else hi = mid - 1;
    }
    const int64_t *src_id = a.out_ids + r_src[lo] + (tid - r_dst[lo]);
    *my_id = EARLY ? __ldcg(src_id) : *src_id;
  }
  __syncthreads();
  return tot;
}

// ==================================================================================== clear
// KVCacheManager.clear (kv_cache_manager.py:444-489; semantics SURVEY A.7): everything released,
// reserved pages trimmed, free list sorted ascending, counters zeroed.  The null block is
// re-reserved by the caller through the alloc path.
__device__ void clear_all(const DevState &c, Smem &s, int order_only) {
  const int tid = threadIdx.x;
  int64_t *sc = s.sc;
  // membership of the new free list = old free list (+ reserved list + every in-use page for clear()).
  // pg_slot is borrowed as the membership flag.  order_only = PageAllocator::reset_free_page_order.
  for (int p = tid; p < c.max_pages; p += NT) c.pg_slot[p] = (!order_only && c.pg_state[p] != PG_NONE) ? 1 : 0;
  __syncthreads();
  int fls = (int)sc[SC_FL_SIZE], rls = order_only ? 0 : (int)sc[SC_RL_SIZE];
  for (int i = tid; i < fls; i += NT) c.pg_slot[c.fl[(int)((sc[SC_FL_HEAD] + i) & c.q_mask)]] = 1;
  for (int i = tid; i < rls; i += NT) c.pg_slot[c.rl[(int)((sc[SC_RL_HEAD] + i) & c.q_mask)]] = 1;
  __syncthreads();
  // unmap events for every mapped page (in-use pages + the trimmed reserve), ascending
  int nev = (int)sc[SC_NEV];
  int base = 0, evn = 0;
  for (int p0 = 0; p0 < c.max_pages; p0 += NT) {
    int p = p0 + tid;
    bool in = (p < c.max_pages) && c.pg_slot[p];
    bool mp = !order_only && (p < c.max_pages) && c.pg_mapped[p];
    int tot, totm;
    int r = block_rank(in, s.wtot, &tot);
    int rm = block_rank(mp, s.wtot2, &totm);
    if (in) c.fl[base + r] = p;   // reset_free_page_order: ascending (page_allocator.cpp:703-709)
    if (mp) {
      if (nev + evn + rm < c.ev_cap) c.ev[nev + evn + rm] = (int64_t)p * 2 + 1;
      c.pg_mapped[p] = 0;
    }
    if (!order_only && p < c.max_pages) {
      c.pg_state[p] = PG_NONE;
      c.pg_cnt[p] = 0;
      c.pg_head[p] = 0;
    }
    base += tot;
    evn += totm;
  }
  __syncthreads();
  if (order_only) {
    if (tid == 0) { sc[SC_FL_HEAD] = 0; sc[SC_FL_SIZE] = base; }
    __syncthreads();
    return;
  }
  int64_t nwords = (c.num_blocks + 31) / 32;
  for (int64_t i = tid; i < nwords; i += NT) c.bitmap[i] = 0;
  if (tid == 0) {
    sc[SC_FL_HEAD] = 0;
    sc[SC_FL_SIZE] = base;
    sc[SC_RL_HEAD] = 0;
    sc[SC_RL_SIZE] = 0;
    sc[SC_NFREE] = base;
    sc[SC_NAB] = 0;
    sc[SC_INSHRINK] = 0;
    sc[SC_TARGET] = -1;
    sc[SC_RB_HEAD] = 0;
    sc[SC_NRB] = 0;
    sc[SC_AP_TAIL] = 0;
    sc[SC_NAVAILPG] = 0;
    sc[SC_NFULLPG] = 0;
    sc[SC_NEV] += evn;
    sc[SC_PHYS] += evn;
    sc[SC_NMAPPED] -= evn;
    sc[SC_HAS_NULL] = 0;
  }
  __syncthreads();
}


// ==================================================================================== prefix-cache pinning
// ElasticBlockPool (kvcached/integration/vllm/patches.py:308-614) on the device: per-block ref_cnt, a bit per
// block for "has a cache key" (_block_id_to_key), and _evictable_blocks — an insertion-ordered dict used as an
// LRU — as an append-only ring with tombstones plus each block's ring position (the avail_pages idiom).

// Compact the LRU ring (drop tombstones, keep order) when appending `room` more entries would overflow it.
__device__ void lru_compact(const DevState &c, Smem &s, int room) {
  int64_t *sc = s.sc;
  const int head = (int)sc[SC_LRU_HEAD], tail = (int)sc[SC_LRU_TAIL];
  if (tail + room <= c.lru_cap) return;   // uniform
  int base = 0;
  for (int i0 = head; i0 < tail; i0 += NT) {
    int i = i0 + threadIdx.x;
    int e = (i < tail) ? c.lru[i] : -1;
    int tot;
    int r = block_rank(e >= 0, s.wtot, &tot);
    if (e >= 0) { c.lru[base + r] = e; c.lru_pos[e] = base + r; }   // lands at or below the chunk just read
    base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) { sc[SC_LRU_HEAD] = 0; sc[SC_LRU_TAIL] = base; }
  __syncthreads();
}

// _evict_blocks_from_pool(k) (patches.py:467-481): pop the k oldest evictable blocks and drop their keys.
// Their ids go to c.fids[0..k) — the argument of the ONE kv_cache_manager.free() that follows — and to
// dst[0..k) for the caller.  Requires k <= SC_LRU_COUNT.
__device__ void lru_pop_oldest(const DevState &c, Smem &s, int k, int64_t *dst) {
  int64_t *sc = s.sc;
  int head = (int)sc[SC_LRU_HEAD];
  const int tail = (int)sc[SC_LRU_TAIL];
  int got = 0;
  while (got < k && head < tail) {   // uniform
    const int i = head + threadIdx.x;
    const int e = (i < tail) ? c.lru[i] : -1;
    int tot;
    const int r = block_rank(e >= 0, s.wtot, &tot);
    if (e >= 0 && got + r < k) {
      c.fids[got + r] = e;
      if (dst) dst[got + r] = e;
      c.lru[i] = -1;
      c.lru_pos[e] = -1;
      atomicAnd(&c.pin_keyed[e >> 5], ~(1u << (e & 31)));
    }
    if (got + tot <= k) {            // every live entry of the chunk was taken
      head = head + NT < tail ? head + NT : tail;
      got += tot;
    } else {                         // the entry of rank k-got-1 is the last one taken
      if (e >= 0 && got + r == k - 1) s.misc[MI_K] = i + 1;
      __syncthreads();
      head = s.misc[MI_K];
      got = k;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    sc[SC_LRU_HEAD] = head;
    sc[SC_LRU_COUNT] -= got;
    if (sc[SC_LRU_COUNT] == 0) { sc[SC_LRU_HEAD] = 0; sc[SC_LRU_TAIL] = 0; }
  }
  __syncthreads();
}

// CTA-wide sum of a per-thread count (uniform result)
__device__ __forceinline__ int block_sum(int v, int32_t *wtot) {
  int tot;
  block_exscan(v, wtot, &tot);
  return tot;
}

// Stage the id list of a PIN request as int32 (-1 = not a block id) out of the caller's id pool.
__device__ __forceinline__ void pin_stage(const DevState &c, const LaunchArgs &a, int64_t arg, int n, int32_t *dst) {
  for (int j = threadIdx.x; j < n; j += NT) {
    int64_t id = a.ids[arg + j];
    dst[j] = (id >= 0 && id < c.num_blocks) ? (int32_t)id : -1;
  }
}

// touch(blocks) (patches.py:519-533): ref_cnt += 1 per occurrence; the block leaves the evictable pool.
__device__ void pin_touch(const DevState &c, Smem &s, const LaunchArgs &a, int64_t arg, int n) {
  int removed = 0;
  for (int j = threadIdx.x; j < n; j += NT) {
    int64_t id = a.ids[arg + j];
    if (id < 0 || id >= c.num_blocks) continue;
    atomicAdd(&c.pin_ref[id], 1);
    int pos = atomicExch(&c.lru_pos[id], -1);
    if (pos >= 0) { c.lru[pos] = -1; removed++; }
  }
  int tot = block_sum(removed, s.wtot);
  if (threadIdx.x == 0) {
    s.sc[SC_LRU_COUNT] -= tot;
    if (s.sc[SC_LRU_COUNT] == 0) { s.sc[SC_LRU_HEAD] = 0; s.sc[SC_LRU_TAIL] = 0; }
  }
  __syncthreads();
}

// free_blocks(ordered_blocks) with caching on (patches.py:546-568).  Sequential semantics: ref_cnt -= 1 per
// occurrence; the occurrence that reaches 0 (= the LAST occurrence of that block in the list) either appends
// the block to the LRU (keyed) or to the list handed to kv_cache_manager.free (not keyed), both in list order;
// then the LRU is cut back to max_cached_blocks (oldest first, a second free call).
// Returns the number of blocks evicted by the cap; their ids are written to dst.
__device__ int pin_free(const DevState &c, Smem &s, const LaunchArgs &a, int64_t arg, int n, int64_t *dst) {
  const int tid = threadIdx.x;
  int64_t *sc = s.sc;
  lru_compact(c, s, n);
  pin_stage(c, a, arg, n, c.pin_in);
  __syncthreads();
  for (int j = tid; j < n; j += NT) {
    int id = c.pin_in[j];
    if (id >= 0) { atomicMax(&c.pin_last[id], j); atomicSub(&c.pin_ref[id], 1); }
  }
  __syncthreads();
  const int tail = (int)sc[SC_LRU_TAIL];
  int nk = 0, nu = 0;
  for (int j0 = 0; j0 < n; j0 += NT) {
    int j = j0 + tid;
    int id = (j < n) ? c.pin_in[j] : -1;
    bool last = id >= 0 && c.pin_last[id] == j;
    bool zero = last && c.pin_ref[id] == 0;
    bool keyed = zero && ((c.pin_keyed[id >> 5] >> (id & 31)) & 1u);
    int tk, tu;
    int rk = block_rank(zero && keyed, s.wtot, &tk);
    int ru = block_rank(zero && !keyed, s.wtot2, &tu);
    if (zero && keyed) { c.lru[tail + nk + rk] = id; c.lru_pos[id] = tail + nk + rk; }
    if (zero && !keyed) c.fids[nu + ru] = id;
    if (last) c.pin_last[id] = -1;
    nk += tk;
    nu += tu;
  }
  if (tid == 0) { sc[SC_LRU_TAIL] = tail + nk; sc[SC_LRU_COUNT] += nk; }
  __syncthreads();
  if (nu > 0) free_core(c, s, a, nu, 0);
  int k = 0;
  if (c.max_cached >= 0 && sc[SC_LRU_COUNT] > c.max_cached) {
    k = (int)(sc[SC_LRU_COUNT] - c.max_cached);
    __syncthreads();
    lru_pop_oldest(c, s, k, dst);
    free_core(c, s, a, k, 0);
  }
  return k;
}

// evict_blocks(block_ids) (patches.py:570-588), ids in the caller's set-iteration order: keys are dropped;
// blocks sitting in the LRU are removed from it and freed (one free call, list order).
__device__ void pin_evict(const DevState &c, Smem &s, const LaunchArgs &a, int64_t arg, int n) {
  const int tid = threadIdx.x;
  pin_stage(c, a, arg, n, c.pin_in);
  __syncthreads();
  int nf = 0;
  for (int j0 = 0; j0 < n; j0 += NT) {
    int j = j0 + tid;
    int id = (j < n) ? c.pin_in[j] : -1;
    bool inl = false;
    if (id >= 0) {
      atomicAnd(&c.pin_keyed[id >> 5], ~(1u << (id & 31)));
      int pos = c.lru_pos[id];
      if (pos >= 0) { c.lru[pos] = -1; c.lru_pos[id] = -1; inl = true; }
    }
    int tot;
    int r = block_rank(inl, s.wtot, &tot);
    if (inl) c.fids[nf + r] = id;
    nf += tot;
  }
  if (tid == 0) {
    s.sc[SC_LRU_COUNT] -= nf;
    if (s.sc[SC_LRU_COUNT] == 0) { s.sc[SC_LRU_HEAD] = 0; s.sc[SC_LRU_TAIL] = 0; }
  }
  __syncthreads();
  if (nf > 0) free_core(c, s, a, nf, 0);
}

// reset_prefix_cache() (patches.py:590-603): every evictable block is freed (LRU order, one call), every key dropped.
__device__ void pin_reset(const DevState &c, Smem &s, const LaunchArgs &a) {
  const int tid = threadIdx.x;
  int64_t *sc = s.sc;
  const int head = (int)sc[SC_LRU_HEAD], tail = (int)sc[SC_LRU_TAIL];
  int nf = 0;
  for (int i0 = head; i0 < tail; i0 += NT) {
    int i = i0 + tid;
    int e = (i < tail) ? c.lru[i] : -1;
    int tot;
    int r = block_rank(e >= 0, s.wtot, &tot);
    if (e >= 0) { c.fids[nf + r] = e; c.lru_pos[e] = -1; }
    nf += tot;
  }
  const int64_t nwords = (c.num_blocks + 31) / 32;
  for (int64_t i = tid; i < nwords; i += NT) c.pin_keyed[i] = 0;
  if (tid == 0) { sc[SC_LRU_HEAD] = 0; sc[SC_LRU_TAIL] = 0; sc[SC_LRU_COUNT] = 0; }
  __syncthreads();
  if (nf > 0) free_core(c, s, a, nf, 0);
}


// ==================================================================================== SGLang slot expansion
// get_num_new_pages + the index arithmetic of SGLang's alloc_extend_kernel / alloc_decode_kernel
// (kvcached/integration/sglang/patches.py:188-279 calls them; SGLang, sglang/srt/mem_cache/allocator.py).
// Pass 1: per sequence new pages = ceil(seq/ps) - ceil(pre/ps) and extend length seq - pre, exclusive-scanned
// over the batch (first page / first output position of every sequence).  Returns {pages, tokens} (uniform).
__device__ void slots_scan(const DevState &c, Smem &s, const LaunchArgs &a, int decode, int64_t *n_pages, int64_t *n_tok) {
  const int bs = a.slot_bs, ps = a.slot_ps;
  int64_t carry_p = 0, carry_t = 0;
  for (int i0 = 0; i0 < bs; i0 += NT) {
    const int i = i0 + threadIdx.x;
    int np = 0, el = 0;
    if (i < bs) {
      const int64_t seq = a.slot_seq[i];
      const int64_t pre = decode ? seq - 1 : a.slot_pre[i];
      np = (int)((seq + ps - 1) / ps - (pre + ps - 1) / ps);
      el = (int)(seq - pre);
      if (np < 0) np = 0;
      if (el < 0) el = 0;
    }
    int tp, tt;
    const int ep = block_exscan(np, s.wtot, &tp);
    const int et = block_exscan(el, s.wtot2, &tt);
    if (i < bs) { c.fpage[i] = (int32_t)(carry_p + ep); c.fids[i] = carry_t + et; }
    carry_p += tp;
    carry_t += tt;
  }
  *n_pages = carry_p;
  *n_tok = carry_t;
}

// Pass 2: one warp per sequence writes its slot indices; `ids` = the block ids just allocated.
__device__ void slots_expand(const DevState &c, const LaunchArgs &a, int decode, const int64_t *ids) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t ps = a.slot_ps;
  for (int i = w; i < a.slot_bs; i += NW) {
    const int64_t seq = a.slot_seq[i];
    const int64_t pre = decode ? seq - 1 : a.slot_pre[i];
    const int64_t last = a.slot_last[i];
    const int64_t pages_before = (pre + ps - 1) / ps, pages_after = (seq + ps - 1) / ps;
    const int64_t *pg = ids + c.fpage[i];
    if (decode) {
      if (lane == 0) a.slot_out[i] = (pages_after == pages_before) ? last + 1 : pg[0] * ps;
      continue;
    }
    int64_t *out = a.slot_out + c.fids[i];
    const int64_t full_before = pages_before * ps;
    const int64_t part1 = (seq < full_before ? seq : full_before) - pre;          // rest of the old partial page
    for (int64_t t = lane; t < part1; t += 32) out[t] = last + 1 + t;
    if (pre + part1 == seq) continue;
    const int64_t part2 = seq / ps * ps - full_before;                            // whole new pages
    for (int64_t t = lane; t < part2; t += 32) out[part1 + t] = pg[t / ps] * ps + t % ps;
    if (pre + part1 + part2 == seq) continue;
    const int64_t part3 = seq - seq / ps * ps;                                    // new partial tail page
    const int64_t tailpg = pg[pages_after - pages_before - 1];
    for (int64_t t = lane; t < part3; t += 32) out[part1 + part2 + t] = tailpg * ps + t;
  }
}

// free(free_index): blocks = sorted unique(free_index // page_size) -> c.fids; returns their count (uniform).
__device__ int slots_unique_blocks(const DevState &c, Smem &s, const LaunchArgs &a, int64_t n) {
  const int tid = threadIdx.x;
  const int64_t ps = a.slot_ps;
  for (int64_t j = tid; j < n; j += NT) {
    int64_t slot = a.slot_pre[j];
    if (slot < 0) continue;
    int64_t b = slot / ps;
    if (b < c.num_blocks) atomicOr(&c.tmp_bits[b >> 5], 1u << (b & 31));
  }
  __syncthreads();
  const int64_t nwords = (c.num_blocks + 31) / 32;
  int base = 0;
  for (int64_t w0 = 0; w0 < nwords; w0 += NT) {
    const int64_t wi = w0 + tid;
    unsigned wd = (wi < nwords) ? c.tmp_bits[wi] : 0u;
    int tot;
    int pos = base + block_exscan(__popc(wd), s.wtot, &tot);
    if (wd) c.tmp_bits[wi] = 0;   // scratch goes back to all-zero
    while (wd) {
      int b = __ffs(wd) - 1;
      wd &= wd - 1;
      c.fids[pos++] = (wi << 5) + b;
    }
    base += tot;
  }
  __syncthreads();
  return base;
}

}  // namespace

// Programmatic dependent launch (sm_90+): a launch made with cudaLaunchAttributeProgrammaticStreamSerialization becomes
// resident while its predecessor in the stream / graph is still running and parks here until that grid has completed
// and its memory is visible; it then lets ITS successor start the same way.  The launch gap between the kernels of
// consecutive scheduler steps (CTA scheduling, parameter and shared-memory set-up) disappears behind the running kernel.
// Without the attribute both instructions do nothing.  Nothing of the device state may be read before pdl_enter().
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ==================================================================================== the step kernel
// Flagged publish of a single-request launch: payload and flag travel in the same 8-byte store, so no fence orders
// them (LlWord in kvb_internal.h).  res and ids were written to device scratch by the request; the caller's block
// barrier made them visible.
__device__ __forceinline__ void publish_flagged(const DevState &c, Smem &s, const LaunchArgs &a, int64_t ev_begin,
                                                int64_t out_begin) {
  const int tid = threadIdx.x;
  const int64_t *sc = s.sc;
  const kvb_res r = a.res[a.op_begin];
  const int64_t nev = sc[SC_NEV] - ev_begin;
  const int nids = (r.status == KVB_ST_OK && r.n > 0 && sc[SC_OUT_CURSOR] > out_begin) ? (int)(sc[SC_OUT_CURSOR] - out_begin) : 0;
  const int nevin = nev < kLlEvents ? (int)nev : kLlEvents;
  const int total = LL_IDS + nids;
  const uint64_t flag = (uint64_t)a.ll_flag << 32;
  const int64_t fls = sc[SC_FL_SIZE];
  for (int wd = tid; wd < total; wd += NT) {
    uint32_t v = 0;
    if (wd == LL_STATUS) v = (uint32_t)r.status;
    else if (wd == LL_N) v = (uint32_t)r.n;
    else if (wd < LL_AUX) v = (uint32_t)((uint64_t)r.avail >> (32 * (wd - LL_AVAIL)));
    else if (wd < LL_NEV) v = (uint32_t)((uint64_t)r.aux >> (32 * (wd - LL_AUX)));
    else if (wd == LL_NEV) v = (uint32_t)nev;
    else if (wd == LL_NHINT) v = (uint32_t)(fls < 16 ? fls : 16);
    else if (wd < LL_SC) {
      const int h = wd - LL_HINT;
      v = h < fls ? (uint32_t)c.fl[(int)((sc[SC_FL_HEAD] + h) & c.q_mask)] : 0xffffffffu;
    } else if (wd < LL_EV) {
      const int q = wd - LL_SC;
      v = (uint32_t)((uint64_t)sc[q >> 1] >> (32 * (q & 1)));
    } else if (wd < LL_IDS) {
      const int q = wd - LL_EV;
      v = (q >> 1) < nevin ? (uint32_t)((uint64_t)c.ev[ev_begin + (q >> 1)] >> (32 * (q & 1))) : 0u;
    } else {
      v = (uint32_t)a.out_ids[out_begin + (wd - LL_IDS)];     // block ids are < 2^31
    }
    a.ll[wd] = flag | v;
  }
}

// ==================================================================================== the step kernel
// Every thread of the CTA calls this when a chunk of requests is staged (kind = its request's kind, -1 past the end):
// one vote per warp leaves, per 32 requests, the mask of those that are NOT an ALLOC — the end of an alloc run is then a
// find-first-set for every thread on its own, no barrier and no shared atomics at the start of each run.
__device__ __forceinline__ void stage_run_masks(Smem &s, int kind) {
  static_assert(NT >= OPS_CHUNK, "one thread per staged request");
  const unsigned na = __ballot_sync(0xffffffffu, kind != KVB_OP_ALLOC);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < OPS_CHUNK / 32) s.namask[w] = na;
}

// COMPACT = the kernel of a scheduler step that holds nothing but alloc(n) / free(ids) requests (kvb_decode_kernel): the
// twenty other request kinds are not compiled in.  A launch executes a few thousand instructions exactly once, so
// what it pays for is instruction FETCH (profiles/r2_step_single_rules.csv) — a kernel a third of the size with the
// hot paths side by side fetches less.  The host picks the kernel per launch (kvb_manager.cpp launch()).
template <bool COMPACT>
__device__ __forceinline__ void step_body(const DevState &c, const LaunchArgs &a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &s = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  int64_t *sc = s.sc;

  // Before pdl_enter(): only what no kernel writes — launch parameters, the prime table, the caller's request arrays.
  // The first chunk of requests is staged and the ref lists of its FREE_REF requests are pulled into L2 (the trace is
  // read once per replay, i.e. from DRAM) while the predecessor kernel is still running.
  {
    const int n_first = (a.op_end - a.op_begin) < OPS_CHUNK ? (a.op_end - a.op_begin) : OPS_CHUNK;
    int kind0 = -1;
    if (tid < n_first) {
      const kvb_op o = a.inline_on ? a.inline_op : a.ops[a.op_begin + tid];
      s.ops[tid] = o;
      kind0 = o.kind;
      if (o.kind == KVB_OP_FREE_REF && a.refs && o.n > 0) {
        const char *r0 = reinterpret_cast<const char *>(a.refs + o.arg);
        const int bytes = (o.n < 1024 ? o.n : 1024) * (int)sizeof(kvb_ref);
        for (int b = 0; b < bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + b));
      }
    }
    stage_run_masks(s, kind0);
  }
  if (tid < 16) s.misc[tid] = 0;
  if (tid >= 64 && tid < 64 + PRIMES_SM) s.primes32[tid - 64] = (uint32_t)c.primes[tid - 64];
  for (int i = tid; i < kGrowTab / 2; i += NT) reinterpret_cast<uint32_t *>(s.grow16)[i] = reinterpret_cast<const uint32_t *>(c.grow_tab)[i];
  // A resident trace opens most steps with the free() of a request that finished one step earlier; its ids are
  // results of launches older than the predecessor (a.safe_before, set by kvb_mgr_trace_run) unless its last block
  // was allocated in the step just before — so the whole gather (refs -> results -> ids in registers) can run here too.
  // A scheduler step lists its frees first (trace.py), so this covers up to EARLY_K of them.
  if (tid < EARLY_K) s.early_n[tid] = -1;
  if (a.safe_before > 0 && !a.inline_on && c.num_blocks < (1ll << 31)) {
    for (int k = 0; k < EARLY_K && a.op_begin + k < a.op_end; ++k) {
      const kvb_op ok = a.ops[a.op_begin + k];
      if (ok.kind != KVB_OP_FREE_REF) break;
      int64_t id = -1;
      const int nk = gather_refs_reg<true>(c, s, a, ok.arg, ok.n, &id, a.safe_before);   // ends with a block barrier
      s.early_ids[k][tid] = (int32_t)id;                                                // block ids are < 2^31, -1 = none
      if (tid == 0) s.early_n[k] = nk;
    }
  }
  pdl_enter();
  if (tid < SC_COUNT) {      // the launch's own adjustments ride on the load: one barrier instead of two
    int64_t v = c.sc[tid];
    if (tid == SC_PHYS) v = a.phys_free;
    if (tid == SC_LAUNCHES) v += 1;
    if (tid == SC_OUT_CURSOR && a.arena_mode == 0) v = 0;
    if (tid == SC_NEV && a.arena_mode != 1 && !a.keep_events) v = 0;
    sc[tid] = v;
  }
  __syncthreads();
  const int64_t ev_begin = a.keep_events ? 0 : sc[SC_NEV];
  const int64_t out_begin = sc[SC_OUT_CURSOR];
#ifdef KVB_PROFILE
  if (tid == 0) s.prof_t = clock64();
#endif

  int op = a.op_begin;
  while (op < a.op_end && !s.misc[MI_STOP]) {
    const int chunk_base = op;
    const int chunk_n = (a.op_end - op) < OPS_CHUNK ? (a.op_end - op) : OPS_CHUNK;
    if (chunk_base != a.op_begin) {      // the first chunk was staged before pdl_enter()
      int kind_q = -1;
      if (tid < chunk_n) { const kvb_op oq = a.inline_on ? a.inline_op : a.ops[chunk_base + tid]; s.ops[tid] = oq; kind_q = oq.kind; }
      stage_run_masks(s, kind_q);
      __syncthreads();
    }
    while (op < chunk_base + chunk_n && !s.misc[MI_STOP]) {
      const kvb_op o = s.ops[op - chunk_base];
      if (a.arena_mode != 1 && sc[SC_NEV] + 2 * (int64_t)c.max_pages + 8 > c.ev_cap) {   // event log nearly full: hand back to the host
        if (tid == 0) s.misc[MI_STOP] = 1;
        __syncthreads();
        break;
      }
      if (!COMPACT && a.group_only) {
        for (int j = tid; j < o.n; j += NT) c.fids[j] = a.ids[o.arg + j];
        __syncthreads();
        free_core(c, s, a, o.n, 1);
        if (tid == 0) { kvb_res r{}; r.status = KVB_ST_OK; r.n = s.misc[MI_D]; r.off = 0; r.avail = 0; a.res[op] = r; }
        op++;
        continue;
      }
      if (!COMPACT && a.resize_target > 0 && (o.kind == KVB_OP_ALLOC || o.kind == KVB_OP_TRY_RESERVE || o.kind == KVB_OP_PIN_NEW || o.kind == KVB_OP_ALLOC_DEV)) {
        // _alloc() first applies a limit change published through the shm segment (kv_cache_manager.py:258-260,
        // page_allocator.cpp:764-782); the result is ignored there too.  Repeating it for the later requests of
        // a run changes nothing, so once per alloc-type request (= once per run) is exact.
        if (tid == 0) mgr_resize(c, sc, a.resize_target);
        __syncthreads();
      }
      // (COMPACT: the host never sends another kind; one that arrives anyway is answered like an unknown kind)
      switch (COMPACT && !(o.kind == KVB_OP_ALLOC || o.kind == KVB_OP_FREE_IDS || o.kind == KVB_OP_FREE_REF ||
                           o.kind == KVB_OP_FREE_RESERVED || o.kind == KVB_OP_AVAIL) ? -1 : o.kind) {
        case KVB_OP_ALLOC: {
          // end of the run of consecutive ALLOC requests inside the staged chunk: first "not an ALLOC" bit after op
          int e = chunk_base + chunk_n;
          {
            const int r = op + 1 - chunk_base;
            for (int wd = r >> 5; wd < OPS_CHUNK / 32; ++wd) {
              unsigned mk = s.namask[wd];
              if (wd == (r >> 5)) mk &= 0xffffffffu << (r & 31);
              if (mk) { e = chunk_base + wd * 32 + __ffs(mk) - 1; break; }
            }
          }
          if (e - op > NT) e = op + NT;
          KVB_PROF(0);
          if (alloc_run_warp(c, s, a, chunk_base, op, e)) {
            op = e;
            KVB_PROF(1);
          } else if (alloc_run_fast(c, s, a, chunk_base, op, e)) {
            op = e;
            KVB_PROF(12);
          } else {
            alloc_run(c, s, a, chunk_base, op, e, 0);
            op = s.sim.op;   // == e unless the output buffer ran out
            KVB_PROF(2);
          }
          break;
        }
        case KVB_OP_TRY_RESERVE: {
          if constexpr (!COMPACT) {
            alloc_run(c, s, a, chunk_base, op, op + 1, 1);
            op++;
            break;
          }
        }
        case KVB_OP_FREE_IDS:
        case KVB_OP_FREE_REF:
        case KVB_OP_FREE_RESERVED: {
          int n = 0;
          bool done = false;
          KVB_PROF(13);
          if (o.kind == KVB_OP_FREE_IDS) {
            n = o.n;
            if (n <= NT && n <= c.f_cap) {     // one id per thread, straight from the request: no staging pass
              int64_t my = -1;
              if (tid < n) my = (a.inline_on == 1) ? a.inline_ids[tid] : a.ids[o.arg + tid];
              KVB_PROF(3);
              done = free_small(c, s, a, n, my) != 0;
            }
            if (!done) {
              if (a.inline_on == 1) { for (int j = tid; j < n; j += NT) c.fids[j] = a.inline_ids[j]; }
              else { for (int j = tid; j < n; j += NT) c.fids[j] = a.ids[o.arg + j]; }
            }
          } else if (o.kind == KVB_OP_FREE_REF) {
            int64_t my = -1;
            const int ek = op - a.op_begin;
            if (ek < EARLY_K && s.early_n[ek] >= 0) { n = s.early_n[ek]; my = s.early_ids[ek][tid]; }
            else n = gather_refs_reg<false>(c, s, a, o.arg, o.n, &my);
            if (n >= 0 && n <= c.f_cap) {
              KVB_PROF(3);
              done = free_small(c, s, a, n, my) != 0;
            }
            if (!done) n = gather_refs(c, s, a, o.arg, o.n);
          } else {
            n = (int)sc[SC_NRB];
            for (int j = tid; j < n; j += NT) c.fids[j] = c.rb[(int)((sc[SC_RB_HEAD] + j) & c.rb_mask)];
          }
          if (!done) {           // ids were staged in c.fids (uniform; a call free_small served ended with a barrier)
            __syncthreads();
            KVB_PROF(3);
          }
          const bool too_many = n > c.f_cap;   // more ids than blocks exist: cannot be a valid call
          if (!done && !too_many && n <= NT) done = free_small(c, s, a, n, tid < n ? c.fids[tid] : -1) != 0;
          if (!done && !too_many) free_core(c, s, a, n, 0);
          if (tid == 0) {
            if (o.kind == KVB_OP_FREE_RESERVED) { sc[SC_NRB] = 0; sc[SC_RB_HEAD] = 0; }
            kvb_res r{}; r.status = too_many ? KVB_ERR_OVERFLOW : KVB_ST_OK; r.n = n; r.off = 0; r.avail = avail_size(c, sc);
            a.res[op] = r;
          }
          op++;                  // (the barrier that closes every request follows the switch)
          break;
        }
        case KVB_OP_RESIZE: {   // KVCacheManager.resize, kv_cache_manager.py:381-401
          if constexpr (!COMPACT) {
            if (tid == 0) {
              kvb_res r{}; r.n = 0; r.off = 0;
              r.status = mgr_resize(c, sc, o.arg);
              r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_PA_RESIZE: {
          if constexpr (!COMPACT) {
            if (tid == 0) {
              kvb_res r{}; r.n = 0; r.off = 0;
              int cap_err = 0;
              r.status = pa_resize(c, sc, o.arg, &cap_err) ? KVB_ST_OK : (cap_err ? KVB_ST_CAPACITY : KVB_ST_FALSE);
              r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_TRIM: {
          if constexpr (!COMPACT) {
            if (tid == 0) {
              pa_trim(c, sc);
              kvb_res r{}; r.status = KVB_ST_OK; r.n = 0; r.off = 0; r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_AVAIL: {
          if (tid == 0) {
            kvb_res r{}; r.status = KVB_ST_OK; r.n = 0; r.off = 0; r.avail = avail_size(c, sc);
            a.res[op] = r;
          }
          op++;
          break;
        }
        case KVB_OP_CLEAR: {
          if constexpr (!COMPACT) {
            // free_reserved(); free_pages(all); trim(); reset_free_page_order(); counters = 0
            clear_all(c, s, 0);
            if (tid == 0) {
              kvb_res r{}; r.status = KVB_ST_OK; r.n = 0; r.off = 0; r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_RESET_ORDER: {
          if constexpr (!COMPACT) {
            clear_all(c, s, 1);
            if (tid == 0) {
              kvb_res r{}; r.status = KVB_ST_OK; r.n = 0; r.off = 0; r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_ALLOC_PAGE: {   // PageAllocator::alloc_page, page_allocator.cpp:161-237
          if constexpr (!COMPACT) {
            if (tid == 0) {
              kvb_res r{}; r.n = 0; r.off = -1;
              if (sc[SC_RL_SIZE] > 0) {
                r.off = dq_pop_front(c.rl, c.q_mask, sc, SC_RL_HEAD, SC_RL_SIZE);
                sc[SC_NFREE] -= 1;
                r.status = KVB_ST_OK;
              } else if (sc[SC_FL_SIZE] > 0) {
                int p = dq_pop_front(c.fl, c.q_mask, sc, SC_FL_HEAD, SC_FL_SIZE);
                sc[SC_NFREE] -= 1;
                ev_push(c, sc, p, 0);
                r.off = p;
                r.status = KVB_ST_OK;
              } else {
                r.status = KVB_ST_NO_PAGES;
              }
              if (r.status == KVB_ST_OK) {   // the page is held by the caller: account it as full
                c.pg_state[(int)r.off] = PG_FULL;
                c.pg_cnt[(int)r.off] = 0;
                sc[SC_NFULLPG] += 1;
              }
              r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_FREE_PAGES: {   // PageAllocator::free_pages, page_allocator.cpp:264-310
          if constexpr (!COMPACT) {
            if (tid == 0) {
              int done = 0;
              for (int i = 0; i < o.n; ++i) {
                int64_t p64 = a.ids[o.arg + i];
                if (p64 < 0 || p64 >= c.max_pages) continue;
                int p = (int)p64;
                if (c.pg_state[p] == PG_NONE) continue;
                sc[SC_NFULLPG] -= (c.pg_state[p] == PG_FULL);
                sc[SC_NAVAILPG] -= (c.pg_state[p] == PG_AVAIL);
                c.pg_state[p] = PG_NONE;
                sc[SC_NFREE] += 1;
                if (sc[SC_RL_SIZE] < c.max_reserved) {
                  dq_push_back(c.rl, c.q_mask, sc, SC_RL_HEAD, SC_RL_SIZE, p);
                } else {
                  ev_push(c, sc, p, 1);
                  dq_push_back(c.fl, c.q_mask, sc, SC_FL_HEAD, SC_FL_SIZE, p);
                }
                done++;
              }
              kvb_res r{}; r.status = KVB_ST_OK; r.n = done; r.off = 0; r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_PIN_NEW: {   // get_new_blocks(n), patches.py:483-517
          if constexpr (!COMPACT) {
            const int n = o.n;
            int64_t evicted = 0;
            int st = KVB_ST_NONE;
            if (n >= 0 && (int64_t)n <= avail_size(c, sc) + sc[SC_LRU_COUNT]) {
              for (int it = 0; it < 2; ++it) {
                const int64_t A = avail_size(c, sc);
                if (A < n && sc[SC_LRU_COUNT] > 0) {
                  const int k = (int)((n - A) < sc[SC_LRU_COUNT] ? (n - A) : sc[SC_LRU_COUNT]);
                  __syncthreads();
                  lru_pop_oldest(c, s, k, a.out_ids + sc[SC_OUT_CURSOR]);
                  if (tid == 0) sc[SC_OUT_CURSOR] += k;
                  __syncthreads();
                  free_core(c, s, a, k, 0);
                  evicted += k;
                }
                if (!alloc_run_fast(c, s, a, chunk_base, op, op + 1)) alloc_run(c, s, a, chunk_base, op, op + 1, 0);
                __syncthreads();
                st = a.res[op].status;
                if (st != KVB_ST_NONE) break;   // ids, or RuntimeError("No free pages left") which the pool does not catch
              }
              if (st == KVB_ST_OK) {
                const int64_t off = a.res[op].off;
                for (int t = tid; t < n; t += NT) c.pin_ref[a.out_ids[off + t]] = 1;
              }
            } else if (tid == 0) {   // "Cannot get n free blocks from the pool": ValueError, nothing changed
              kvb_res r{}; r.status = KVB_ST_NONE; r.n = 0; r.off = sc[SC_OUT_CURSOR]; r.avail = avail_size(c, sc); r.aux = 0;
              a.res[op] = r;
            }
            __syncthreads();
            if (tid == 0) a.res[op].aux = evicted;
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_PIN_TOUCH:
        case KVB_OP_PIN_MARK:
        case KVB_OP_PIN_FREE:
        case KVB_OP_PIN_EVICT:
        case KVB_OP_PIN_RESET: {
          if constexpr (!COMPACT) {
            int64_t aux = 0;
            const int64_t off0 = sc[SC_OUT_CURSOR];
            const bool too_many = o.n > c.f_cap;
            if (too_many) {
            } else if (o.kind == KVB_OP_PIN_TOUCH) {
              pin_touch(c, s, a, o.arg, o.n);
            } else if (o.kind == KVB_OP_PIN_MARK) {   // cache_full_blocks: _block_id_to_key[block_id] = key (patches.py:463-465)
              for (int j = tid; j < o.n; j += NT) {
                int64_t id = a.ids[o.arg + j];
                if (id >= 0 && id < c.num_blocks) atomicOr(&c.pin_keyed[id >> 5], 1u << (id & 31));
              }
            } else if (o.kind == KVB_OP_PIN_FREE) {
              aux = pin_free(c, s, a, o.arg, o.n, a.out_ids + off0);
            } else if (o.kind == KVB_OP_PIN_EVICT) {
              pin_evict(c, s, a, o.arg, o.n);
            } else {
              pin_reset(c, s, a);
            }
            __syncthreads();
            if (tid == 0) {
              sc[SC_OUT_CURSOR] = off0 + aux;
              kvb_res r{}; r.status = too_many ? KVB_ERR_OVERFLOW : KVB_ST_OK; r.n = o.n; r.off = off0; r.avail = avail_size(c, sc); r.aux = aux;
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_ALLOC_EXTEND:
        case KVB_OP_ALLOC_DECODE: {
          if constexpr (!COMPACT) {
            const int decode = o.kind == KVB_OP_ALLOC_DECODE;
            int64_t npg, ntok;
            slots_scan(c, s, a, decode, &npg, &ntok);
            int st = KVB_ST_OK;
            if ((decode ? (int64_t)a.slot_bs : ntok) > a.slot_tokens || npg > INT_MAX) {
              st = KVB_ERR_OVERFLOW;                 // caller's index buffer is too small
            } else if (npg > 0) {                    // the reference skips alloc() when no page is needed (:217, :258)
              if (tid == 0) {
                s.ops[op - chunk_base].n = (int32_t)npg;
                if (a.resize_target > 0) mgr_resize(c, sc, a.resize_target);   // alloc()'s limit check, kv_cache_manager.py:258-260
              }
              __syncthreads();
              if (!alloc_run_fast(c, s, a, chunk_base, op, op + 1)) alloc_run(c, s, a, chunk_base, op, op + 1, 0);
              __syncthreads();
              st = a.res[op].status;
            }
            if (st == KVB_ST_OK) slots_expand(c, a, decode, a.out_ids + (npg > 0 ? a.res[op].off : 0));
            __syncthreads();
            if (tid == 0) {
              kvb_res r{};
              if (npg > 0 && st != KVB_ERR_OVERFLOW) r = a.res[op];
              else { r.off = sc[SC_OUT_CURSOR]; r.avail = avail_size(c, sc); }
              r.status = st;
              r.n = st == KVB_ST_OK ? (int32_t)npg : 0;
              r.aux = decode ? a.slot_bs : ntok;    // slot indices written
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_ALLOC_DEV: {   // alloc(n); ids also land in the caller's device array
          if constexpr (!COMPACT) {
            if (!alloc_run_fast(c, s, a, chunk_base, op, op + 1)) alloc_run(c, s, a, chunk_base, op, op + 1, 0);
            __syncthreads();
            if (a.res[op].status == KVB_ST_OK) {
              const int64_t off = a.res[op].off;
              for (int t = tid; t < o.n; t += NT) a.slot_out[t] = a.out_ids[off + t];
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_FREE_DEV: {    // free(ids), ids taken in order from the caller's device array
          if constexpr (!COMPACT) {
            const int n = (int)o.arg;
            const bool too_many = n > c.f_cap;
            if (!too_many) {
              bool done = false;
              if (n <= NT) done = free_small(c, s, a, n, tid < n ? a.slot_pre[tid] : -1) != 0;
              if (!done) {
                for (int j = tid; j < n; j += NT) c.fids[j] = a.slot_pre[j];
                __syncthreads();
                free_core(c, s, a, n, 0);
              }
            }
            if (tid == 0) {
              kvb_res r{}; r.status = too_many ? KVB_ERR_OVERFLOW : KVB_ST_OK; r.n = n; r.off = 0; r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        case KVB_OP_FREE_SLOTS: {
          if constexpr (!COMPACT) {
            const int nb = slots_unique_blocks(c, s, a, o.arg);
            if (nb > 0) free_core(c, s, a, nb, 0);
            if (tid == 0) {
              kvb_res r{}; r.status = KVB_ST_OK; r.n = nb; r.off = 0; r.avail = avail_size(c, sc);
              a.res[op] = r;
            }
            __syncthreads();
            op++;
            break;
          }
        }
        default: {
          if (tid == 0) { kvb_res r{}; r.status = KVB_ERR_INVALID; r.n = 0; r.off = 0; r.avail = 0; a.res[op] = r; }
          op++;
          break;
        }
      }
      __syncthreads();
    }
    __syncthreads();
  }

  // publish
  __syncthreads();
  KVB_PROF(9);
  if (tid < SC_COUNT) c.sc[tid] = sc[tid];
  if (a.ll) publish_flagged(c, s, a, ev_begin, out_begin);
  Summary *sm = a.summary;
  if (sm) {
    int64_t nev = sc[SC_NEV] - ev_begin;
    int64_t nin = nev < EV_INLINE ? nev : EV_INLINE;
    for (int i = tid; i < nin; i += NT) sm->ev_inline[i] = c.ev[ev_begin + i];
    if (tid < SC_COUNT) sm->sc[tid] = sc[tid];
    if (tid < 16) {   // map-ahead hints: reserved pages are mapped already, so the next mappings are the free list's front
      const int64_t fls = sc[SC_FL_SIZE];
      sm->next_free[tid] = tid < fls ? c.fl[(int)((sc[SC_FL_HEAD] + tid) & c.q_mask)] : -1;
      if (tid == 0) sm->next_free_n = fls < 16 ? fls : 16;
    }
    // Tensor parallel: this launch's page list goes straight into the mailbox every rank of the node polls (kvb_tp.h)
    // — plain stores into mapped host memory, a system-scope fence, then the slot's sequence word.  Lists longer
    // than a slot are left to the host relay (clear / trim of a big pool); steps without events publish nothing.
    const bool tp_push = a.tp_ring && nev > 0 && nev <= EV_INLINE;
    if (tp_push) {
      int64_t *slot = a.tp_ring + (size_t)(a.tp_pub & a.tp_mask) * a.tp_slot_words;
      for (int i = tid; i < (int)nev; i += NT) slot[8 + 16 + i] = c.ev[ev_begin + i];
      if (tid < 16) {
        const int64_t fls = sc[SC_FL_SIZE];
        slot[8 + tid] = tid < fls ? c.fl[(int)((sc[SC_FL_HEAD] + tid) & c.q_mask)] : -1;
      }
    }
    const bool to_host = a.host_res != nullptr || a.out_on_host;
    if (a.host_res) {
      // zero-copy step: results go straight to the caller's pinned host buffers over PCIe
      const kvb_res *dr = a.res + a.op_begin;
      kvb_res *hr = a.host_res + a.op_begin;
      for (int i = tid; i < op - a.op_begin; i += NT) hr[i] = dr[i];
      const int64_t o0 = out_begin, o1 = sc[SC_OUT_CURSOR];
      for (int64_t i = o0 + tid; i < o1; i += NT) a.host_out[i] = a.out_ids[i];
    }
    if (tid == 0) {
      if (tp_push) {
        int64_t *slot = a.tp_ring + (size_t)(a.tp_pub & a.tp_mask) * a.tp_slot_words;
        const int64_t fls = sc[SC_FL_SIZE];
        slot[1] = a.seq;
        slot[2] = nev;
        slot[3] = 0;                                   // no continuation
        slot[4] = (a.tp_flags & 1) ? 0 : (fls < 16 ? fls : 16);
        slot[5] = a.tp_flags;
      }
      sm->tp_published = tp_push ? 1 : 0;
      sm->ops_done = op - a.op_begin;
      sm->out_used = sc[SC_OUT_CURSOR] - out_begin;
      sm->n_events = nev;
      sm->ev_begin = ev_begin;
      sm->ev_inline_n = nin;
    }
    // ONE system-scope fence per thread after everything it wrote for the host (results, ids, summary, mailbox
    // slot), then the barrier: when thread 0 passes it, all of that is visible system-wide, and the sequence words
    // the host / the TP followers poll can follow without a second fence.
    if (to_host || tp_push) __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      if (tp_push) {
        int64_t *slot = a.tp_ring + (size_t)(a.tp_pub & a.tp_mask) * a.tp_slot_words;
        *reinterpret_cast<volatile int64_t *>(slot) = a.tp_pub + 1;   // followers poll this word
        *reinterpret_cast<volatile int64_t *>(a.tp_published) = a.tp_pub + 1;
      }
      *reinterpret_cast<volatile int64_t *>(&sm->seq) = a.seq;        // written last: the host polls it
    }
  }
  KVB_PROF(10);
}

// __grid_constant__: the device functions take the parameter structs by reference; without it every thread first
// copies them (2.7 KB) to its local memory.
__global__ void __launch_bounds__(NT, 1) kvb_step_kernel(const __grid_constant__ DevState c, const __grid_constant__ LaunchArgs a) {
  step_body<false>(c, a);
}
__global__ void __launch_bounds__(NT, 1) kvb_decode_kernel(const __grid_constant__ DevState c, const __grid_constant__ LaunchArgs a) {
  step_body<true>(c, a);
}

// ==================================================================================== single requests
// kvb_mgr_alloc / kvb_mgr_free launch THIS kernel: the same device functions (alloc_run_warp, free_small, the flagged
// publish) without the other twenty request kinds around them.  kvb_step_kernel is 1.3 MB of SASS and a launch
// executes a few thousand instructions of it exactly once: ncu charged 66 % of the issue stalls of a single-request
// launch to instruction fetch (profiles/r2_step_single_rules.csv).  Requests this kernel cannot serve with those two
// paths (shrink mode, blocks straddling page edges, windows too small, > 256 distinct pages) are answered with
// KVB_ST_RETRY — nothing touched — and the host sends them through kvb_step_kernel.
__global__ void __launch_bounds__(NT, 1) kvb_one_kernel(const __grid_constant__ DevState c, const __grid_constant__ LaunchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &s = *reinterpret_cast<Smem *>(smem_raw);
  const int tid = threadIdx.x;
  int64_t *sc = s.sc;
  if (tid < 16) s.misc[tid] = 0;
  if (tid >= 64 && tid < 64 + PRIMES_SM) s.primes32[tid - 64] = (uint32_t)c.primes[tid - 64];
  for (int i = tid; i < kGrowTab / 2; i += NT) reinterpret_cast<uint32_t *>(s.grow16)[i] = reinterpret_cast<const uint32_t *>(c.grow_tab)[i];
  if (tid == 0) s.ops[0] = a.inline_op;
  pdl_enter();                           // nothing of the device state is read above this line
  if (tid < SC_COUNT) {
    int64_t v = c.sc[tid];
    if (tid == SC_PHYS) v = a.phys_free;
    if (tid == SC_LAUNCHES) v += 1;
    if (tid == SC_OUT_CURSOR) v = 0;
    if (tid == SC_NEV && !a.keep_events) v = 0;
    sc[tid] = v;
  }
  __syncthreads();
  const int64_t ev_begin = a.keep_events ? 0 : sc[SC_NEV];
#ifdef KVB_PROFILE
  if (tid == 0) s.prof_t = clock64();
#endif
  const kvb_op o = s.ops[0];
  int retry = 0;
  if (o.kind == KVB_OP_ALLOC) {
    if (a.resize_target > 0) {            // _alloc() first applies a limit change (kv_cache_manager.py:258-260)
      if (tid == 0) mgr_resize(c, sc, a.resize_target);
      __syncthreads();
    }
    retry = !alloc_run_warp(c, s, a, 0, 0, 1);
  } else if (o.kind == KVB_OP_FREE_IDS) {
    int64_t my = -1;
    if (tid < o.n) my = (a.inline_on == 1) ? a.inline_ids[tid] : a.ids[tid];
    retry = (o.n > NT) || !free_small(c, s, a, o.n, my);
    if (!retry && tid == 0) {
      kvb_res r{}; r.status = KVB_ST_OK; r.n = o.n; r.off = 0; r.avail = avail_size(c, sc);
      a.res[0] = r;
    }
  } else if (tid == 0) {                  // KVB_OP_AVAIL: the flush of frees that published nothing
    kvb_res r{}; r.status = KVB_ST_OK; r.n = 0; r.off = 0; r.avail = avail_size(c, sc);
    a.res[0] = r;
  }
  if (retry && tid == 0) {
    kvb_res r{}; r.status = KVB_ST_RETRY; r.n = 0; r.off = 0; r.avail = avail_size(c, sc);
    a.res[0] = r;
  }
  __syncthreads();
  KVB_PROF(9);
  if (tid < SC_COUNT) c.sc[tid] = sc[tid];
  if (a.ll) publish_flagged(c, s, a, ev_begin, 0);
  KVB_PROF(10);
}

// ==================================================================================== init
// Initial state: free_page_list = 0..N-1 (page_allocator.cpp:121-124), everything else empty.
__global__ void kvb_init_state_kernel(DevState c, int64_t num_total_pages) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = i; p <= c.max_pages; p += stride) {
    c.pg_state[p] = PG_NONE;
    c.pg_cnt[p] = 0;
    c.pg_head[p] = 0;
    c.pg_pos[p] = -1;
    c.pg_mapped[p] = 0;
    c.pg_first[p] = INT_MAX;
    c.pg_slot[p] = 0;
    if (p < num_total_pages) c.fl[p] = (int32_t)p;
  }
  const int64_t nwords = (c.num_blocks + 31) / 32;
  for (int64_t wd = i; wd < nwords + 4; wd += stride) { c.bitmap[wd] = 0; c.pin_keyed[wd] = 0; c.tmp_bits[wd] = 0; }
  for (int64_t b = i; b < c.num_blocks; b += stride) { c.pin_ref[b] = 0; c.lru_pos[b] = -1; c.pin_last[b] = -1; }
  if (i < SC_COUNT) {
    int64_t v = 0;
    if (i == SC_NFREE || i == SC_NTOTAL || i == SC_FL_SIZE) v = num_total_pages;
    if (i == SC_TARGET) v = -1;
    c.sc[i] = v;
  }
}

// streaming 128-bit accesses (no L1 allocation): shared by the audit sweep and the KV scatter / gather
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// ==================================================================================== audit sweep
// Recount every page from the used-block bitmap with coalesced 128-bit loads + popc and compare with
// the per-page counters; also counts in-use pages whose blocks are all free ("reclaimable").
// out4 = {used blocks, in-use pages, fully-free in-use pages, mismatching pages}
//   mode 0  bpp in {32, 64, 128}: a thread's uint4 (128 blocks) covers 128/bpp whole pages
//   mode 1  bpp = 128 * g, g a power of two <= 32: g adjacent lanes hold one page, shuffle-reduce
//   mode 2  anything else (blocks straddling page edges, odd sizes): one warp per page, masked words
__device__ __forceinline__ void audit_page(const DevState &c, int64_t p, unsigned used_p, unsigned long long &used,
                                           unsigned long long &inuse, unsigned long long &empties,
                                           unsigned long long &bad) {
  int st = c.pg_state[p];
  used += used_p;
  if (st != PG_NONE) {
    inuse++;
    if (used_p == 0) empties++;
    if ((int)used_p != pg_nblk(c, (int)p) - c.pg_cnt[p]) bad++;
  } else if (used_p != 0) {
    bad++;
  }
}

// Streaming path of the audit (bpp in {32, 64, 128}: a uint4 of bitmap = PPV whole pages): UNR independent 128-bit
// bitmap loads plus the matching state / count vectors are issued before anything is consumed.  HBM-bound at
// scale: 16 B of bitmap + 8 B of counters per page.
template <int PPV>
__device__ __forceinline__ void audit_stream(const DevState &c, int64_t gtid, int64_t gthreads, unsigned long long &used,
                                             unsigned long long &inuse, unsigned long long &empties,
                                             unsigned long long &bad) {
  const int64_t nvec = ((int64_t)c.max_pages + PPV - 1) / PPV;
  const uint4 *b4 = reinterpret_cast<const uint4 *>(c.bitmap);
  constexpr int UNR = 4;
  for (int64_t q0 = gtid; q0 < nvec; q0 += gthreads * UNR) {
    uint4 v[UNR];
    int st[UNR][PPV], cn[UNR][PPV];
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int64_t q = q0 + (int64_t)k * gthreads;
      v[k] = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int x = 0; x < PPV; ++x) { st[k][x] = PG_NONE; cn[k][x] = 0; }
      if (q < nvec) {
        v[k] = ld_stream(b4 + q);
        const int64_t p = q * PPV;
        if (PPV == 4 && p + 4 <= c.max_pages) {
          int4 a = *reinterpret_cast<const int4 *>(c.pg_state + p), b = *reinterpret_cast<const int4 *>(c.pg_cnt + p);
          st[k][0] = a.x; st[k][1 % PPV] = a.y; st[k][2 % PPV] = a.z; st[k][3 % PPV] = a.w;
          cn[k][0] = b.x; cn[k][1 % PPV] = b.y; cn[k][2 % PPV] = b.z; cn[k][3 % PPV] = b.w;
        } else if (PPV == 2 && p + 2 <= c.max_pages) {
          int2 a = *reinterpret_cast<const int2 *>(c.pg_state + p), b = *reinterpret_cast<const int2 *>(c.pg_cnt + p);
          st[k][0] = a.x; st[k][1 % PPV] = a.y;
          cn[k][0] = b.x; cn[k][1 % PPV] = b.y;
        } else {
#pragma unroll
          for (int x = 0; x < PPV; ++x)
            if (p + x < c.max_pages) { st[k][x] = c.pg_state[p + x]; cn[k][x] = c.pg_cnt[p + x]; }
        }
      }
    }
    unsigned u32 = 0, in32 = 0, em32 = 0, bad32 = 0;
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int64_t q = q0 + (int64_t)k * gthreads;
      if (q >= nvec) continue;
      unsigned up[PPV];
      if (PPV == 4) { up[0] = __popc(v[k].x); up[1 % PPV] = __popc(v[k].y); up[2 % PPV] = __popc(v[k].z); up[3 % PPV] = __popc(v[k].w); }
      else if (PPV == 2) { up[0] = __popc(v[k].x) + __popc(v[k].y); up[1 % PPV] = __popc(v[k].z) + __popc(v[k].w); }
      else { up[0] = __popc(v[k].x) + __popc(v[k].y) + __popc(v[k].z) + __popc(v[k].w); }
#pragma unroll
      for (int x = 0; x < PPV; ++x) {
        if (q * PPV + x < c.max_pages) {
          u32 += up[x];
          if (st[k][x] != PG_NONE) {
            in32++;
            em32 += (up[x] == 0);
            bad32 += ((int)up[x] != c.bpp - cn[k][x]);
          } else {
            bad32 += (up[x] != 0);
          }
        }
      }
    }
    used += u32; inuse += in32; empties += em32; bad += bad32;
  }
}

__global__ void __launch_bounds__(256) kvb_audit_kernel(DevState c, int64_t *out4, int mode) {
  const int lane = threadIdx.x & 31;
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
  unsigned long long used = 0, inuse = 0, empties = 0, bad = 0;
  if (mode == 0) {
    if (c.bpp == 32) audit_stream<4>(c, gtid, gthreads, used, inuse, empties, bad);
    else if (c.bpp == 64) audit_stream<2>(c, gtid, gthreads, used, inuse, empties, bad);
    else audit_stream<1>(c, gtid, gthreads, used, inuse, empties, bad);
  } else if (mode == 1) {
    // bpp = 128 * g: g adjacent lanes hold one page; UNR independent loads in flight, shuffle-reduce each
    const int g = c.bpp / 128;
    const int gsh = __ffs(g) - 1;             // g is a power of two: no 64-bit divisions in the loop
    const int64_t nvec = (int64_t)c.max_pages * g;
    const uint4 *b4 = reinterpret_cast<const uint4 *>(c.bitmap);
    constexpr int UNR = 4;
    const int64_t span = (nvec + gthreads * UNR - 1) / (gthreads * UNR);
    for (int64_t it = 0; it < span; ++it) {   // uniform trip count: shuffles need the whole warp
      unsigned u[UNR];
      int64_t q[UNR];
      int st[UNR], cn[UNR];
#pragma unroll
      for (int k = 0; k < UNR; ++k) {          // every load of the trip is issued before anything is consumed
        q[k] = (it * UNR + k) * gthreads + gtid;
        uint4 v = make_uint4(0, 0, 0, 0);
        st[k] = PG_NONE;
        cn[k] = 0;
        if (q[k] < nvec) {
          v = ld_stream(b4 + q[k]);
          if ((q[k] & (g - 1)) == 0) { st[k] = c.pg_state[q[k] >> gsh]; cn[k] = c.pg_cnt[q[k] >> gsh]; }
        }
        u[k] = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
      }
      unsigned u32 = 0, in32 = 0, em32 = 0, bad32 = 0;
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        for (int d = 1; d < g; d <<= 1) u[k] += __shfl_xor_sync(0xffffffffu, u[k], d);
        if (q[k] < nvec && (q[k] & (g - 1)) == 0) {
          u32 += u[k];
          if (st[k] != PG_NONE) {
            in32++;
            em32 += (u[k] == 0);
            bad32 += ((int)u[k] != c.bpp - cn[k]);
          } else {
            bad32 += (u[k] != 0);
          }
        }
      }
      used += u32; inuse += in32; empties += em32; bad += bad32;
    }
  } else {
    const int64_t warp = gtid >> 5, nwarps = gthreads >> 5;
    for (int64_t p = warp; p < c.max_pages; p += nwarps) {
      int64_t start = pg_start(c, (int)p);
      int64_t end = start + pg_nblk(c, (int)p);
      unsigned u = 0;
      for (int64_t wd = (start >> 5) + lane; wd <= ((end - 1) >> 5) && end > start; wd += 32) {
        unsigned x = c.bitmap[wd];
        int64_t lo = wd << 5;
        if (lo < start) x &= ~0u << (start - lo);
        if (lo + 32 > end) x &= ~0u >> (lo + 32 - end);
        u += __popc(x);
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) u += __shfl_xor_sync(0xffffffffu, u, d);
      if (lane == 0) audit_page(c, p, u, used, inuse, empties, bad);
    }
  }
  // warp-level reduction, one atomic per warp and counter
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    used += __shfl_xor_sync(0xffffffffu, used, d);
    inuse += __shfl_xor_sync(0xffffffffu, inuse, d);
    empties += __shfl_xor_sync(0xffffffffu, empties, d);
    bad += __shfl_xor_sync(0xffffffffu, bad, d);
  }
  if (lane == 0) {
    if (used) atomicAdd((unsigned long long *)&out4[0], used);
    if (inuse) atomicAdd((unsigned long long *)&out4[1], inuse);
    if (empties) atomicAdd((unsigned long long *)&out4[2], empties);
    if (bad) atomicAdd((unsigned long long *)&out4[3], bad);
  }
}

// Dump avail_pages in insertion order (tests): skips tombstones.
__global__ void kvb_dump_avail_kernel(DevState c, int64_t *out, int64_t *n_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int tail = (int)c.sc[SC_AP_TAIL];
    int64_t n = 0;
    for (int i = 0; i < tail; ++i) {
      int e = c.ap[i];
      if (e >= 0) out[n++] = e;
    }
    *n_out = n;
  }
}

// ==================================================================================== KV scatter / gather
// Moves token K / V cells between the dense buffers [layer][T][cell] (K and V separately) and the paged range
//   base + block*stride_block + in_block*stride_token + layer*stride_layer + kv*stride_kv
// (kvcached/integration/vllm/interfaces.py:281-298, integration/sglang/interfaces.py:132-173).  HBM-bound:
// 2*cell read + 2*cell written per token and layer, nothing is touched twice.
//   * `g` lanes (a power of two chosen by the launcher) serve one cell, a warp serves 32/g cells side by side;
//   * every lane keeps V*U independent 128-bit loads in flight before its first store: V vectors of each of U cells
//     (2 KiB cell: V=4, U=1 as two passes; 512 B: V=2, U=2; 256 B: V=1, U=4) — with one cell per pass a 256-byte
//     cell leaves a lane a single load in flight and the kernel waits on latency instead of streaming;
//   * loads bypass L1 (no reuse), stores are whole 128-byte lines;
//   * a warp owns a CONTIGUOUS run of the flattened (kv, layer, token) space, token fastest: its dense-side
//     traffic is one contiguous stream, and (kv, layer, token) advance incrementally — one 64-bit division per
//     warp instead of three per cell;
//   * grid = persistent CTAs, SM count x CTAs resident at once.
// A negative slot is padding and moves nothing.
template <int GATHER, int V, int U>
__global__ void __launch_bounds__(256) kvb_kv_copy_kernel(kvb_kv_layout lay, const int64_t *__restrict__ slots,
                                                          int64_t T, unsigned char *kbuf, unsigned char *vbuf,
                                                          int32_t l0, int32_t l1, int32_t g, int32_t bs_shift, int32_t chunked) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & (g - 1);            // lane inside its group
  const int sub = lane / g;                 // which of the warp's cells this group serves
  const int cpw = 32 / g;                   // cells per warp and pass
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nl = l1 - l0;
  const int64_t items = T * nl * lay.num_kv;
  const int per_iter = cpw * U;
  // chunked: a warp owns one contiguous run; strided: pass p of warp w covers cells [(p*nwarps + w) * per_iter, +per_iter)
  // so that neighbouring warps stream neighbouring addresses at the same time
  int64_t begin, end, step;
  if (chunked) {
    int64_t chunk = (items + nwarps - 1) / nwarps;
    chunk = (chunk + per_iter - 1) / per_iter * per_iter;
    begin = warp * chunk;
    end = begin + chunk < items ? begin + chunk : items;
    step = per_iter;
  } else {
    begin = warp * per_iter;
    end = items;
    step = nwarps * per_iter;
  }
  if (begin >= items) return;
  const int vec = (int)(lay.cell_bytes >> 4);   // 16-byte vectors per cell
  int64_t lk = begin / T;                       // (kv * nl + layer) of the pass's first cell
  int64_t tok = begin - lk * T;
  for (int64_t base = begin; base < end; base += step) {
    const uint4 *src[U];
    uint4 *dst[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int off = k * cpw + sub;
      src[k] = nullptr;
      dst[k] = nullptr;
      if (base + off < end) {   // (strided: end = items)
        int64_t t = tok + off, q = lk;
        if (t >= T) { const int64_t w = t / T; q += w; t -= w * T; }
        const int64_t slot = slots[t];
        if (slot >= 0) {
          const int kv = (int)(q / nl);
          const int layer = (int)(q - (int64_t)kv * nl);
          int64_t blk;
          int inb;
          if (bs_shift >= 0) { blk = slot >> bs_shift; inb = (int)(slot & ((1 << bs_shift) - 1)); }
          else { blk = slot / lay.block_size; inb = (int)(slot - blk * lay.block_size); }
          unsigned char *paged = reinterpret_cast<unsigned char *>(lay.base) + blk * lay.stride_block +
                                 inb * lay.stride_token + (int64_t)(l0 + layer) * lay.stride_layer + kv * lay.stride_kv;
          unsigned char *dense = (kv ? vbuf : kbuf) + ((int64_t)layer * T + t) * lay.cell_bytes;
          src[k] = reinterpret_cast<const uint4 *>(GATHER ? paged : dense);
          dst[k] = reinterpret_cast<uint4 *>(GATHER ? dense : paged);
        }
      }
    }
    for (int v0 = gl; v0 < vec; v0 += g * V) {
      uint4 a[U][V];
#pragma unroll
      for (int k = 0; k < U; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (src[k] && v0 + j * g < vec) a[k][j] = ld_stream(src[k] + v0 + j * g);
#pragma unroll
      for (int k = 0; k < U; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (src[k] && v0 + j * g < vec) st_stream(dst[k] + v0 + j * g, a[k][j]);
    }
    tok += step;
    if (tok >= T) { const int64_t w = tok / T; lk += w; tok -= w * T; }
  }
}

// ==================================================================================== host launchers
size_t step_kernel_smem_bytes() { return sizeof(Smem); }

// Everything cached here is per DEVICE (managers may live on several GPUs of one process): SM count, resident
// grid sizes, and the step kernel's dynamic shared-memory attribute, which belongs to the device's context.
constexpr int kMaxDev = 64;
static int cur_dev() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return (dev >= 0 && dev < kMaxDev) ? dev : 0;
}
static int g_sm_count[kMaxDev];
static int sm_count() {
  int dev = cur_dev();
  if (!g_sm_count[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count[dev] = n > 0 ? n : 148;
  }
  return g_sm_count[dev];
}

// Persistent grid = SMs x CTAs of this kernel that are resident at once (registers decide: 3 for the audit, 6 / 8
// for scatter / gather): every CTA runs the same number of grid-stride trips, no partial last wave.
template <typename K>
static int resident_grid(K kernel, int threads, int *cache_by_dev) {
  int *cache = cache_by_dev + cur_dev();
  if (!*cache) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm <= 0) {
      cudaGetLastError();
      per_sm = 4;
    }
    *cache = per_sm * sm_count();
  }
  return *cache;
}

// One CTA of NT threads; with programmatic stream serialization unless KVB200_PDL=0 (pdl_enter() above).
static int launch_bookkeeping(void (*kernel)(const DevState, const LaunchArgs), const DevState &st, const LaunchArgs &a, void *stream) {
  static const bool pdl = []() { const char *e = getenv("KVB200_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = sizeof(Smem);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && !(a.tune & 2)) ? 1 : 0;     // KVB200_TUNE bit 1: same-process A/B of the attribute
  return (int)cudaLaunchKernelEx(&cfg, kernel, st, a);
}

int launch_step_prepare() {
  static bool attr_set[kMaxDev];
  const int dev = cur_dev();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kvb_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(kvb_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  return 0;
}

int launch_step(const DevState &st, const LaunchArgs &a, void *stream, bool alloc_free_only) {
  int rc = launch_step_prepare();
  if (rc) return rc;
  static const bool compact_on = []() { const char *e = getenv("KVB200_DECODE_KERNEL"); return !(e && e[0] == '0'); }();
  const bool compact = alloc_free_only && compact_on && !(a.tune & 8) && !a.group_only && a.resize_target <= 0;
  return launch_bookkeeping(compact ? kvb_decode_kernel : kvb_step_kernel, st, a, stream);
}

int launch_one(const DevState &st, const LaunchArgs &a, void *stream) {
  static bool attr_set[kMaxDev];
  const int dev = cur_dev();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kvb_one_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  return launch_bookkeeping(kvb_one_kernel, st, a, stream);
}

int launch_init_state(const DevState &st, int64_t num_total_pages, void *stream) {
  kvb_init_state_kernel<<<sm_count(), 256, 0, (cudaStream_t)stream>>>(st, num_total_pages);
  return (int)cudaGetLastError();
}

int launch_audit(const DevState &st, int64_t *d_out4, void *stream, bool zero_counters) {
  if (zero_counters) {
    cudaError_t e = cudaMemsetAsync(d_out4, 0, 4 * sizeof(int64_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
  }
  int mode = 2;
  if (st.exact && (st.bpp == 32 || st.bpp == 64 || st.bpp == 128)) mode = 0;
  else if (st.exact && st.bpp % 128 == 0 && st.bpp / 128 <= 32 && ((st.bpp / 128) & (st.bpp / 128 - 1)) == 0) mode = 1;
  static int grid_audit[kMaxDev];
  kvb_audit_kernel<<<resident_grid(kvb_audit_kernel, 256, grid_audit), 256, 0, (cudaStream_t)stream>>>(st, d_out4, mode);
  return (int)cudaGetLastError();
}

int launch_dump_avail(const DevState &st, int64_t *d_out, int64_t *d_n, void *stream) {
  kvb_dump_avail_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(st, d_out, d_n);
  return (int)cudaGetLastError();
}

template <int GATHER, int V, int U>
static int launch_kv_copy_t(const kvb_kv_layout &lay, const int64_t *slots, int64_t T, void *k, void *v, int32_t l0,
                            int32_t l1, int g, int bs_shift, int chunked, int64_t cells, void *stream) {
  static int grid_cache[kMaxDev];
  int64_t grid = resident_grid(kvb_kv_copy_kernel<GATHER, V, U>, 256, grid_cache);
  const int64_t want = (cells * g + 255) / 256;   // one group of g lanes per cell
  if (want < grid) grid = want;
  kvb_kv_copy_kernel<GATHER, V, U><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      lay, slots, T, (unsigned char *)k, (unsigned char *)v, l0, l1, g, bs_shift, chunked);
  return (int)cudaGetLastError();
}

int launch_kv_copy(const kvb_kv_layout &lay, const int64_t *slots, int64_t T, void *k, void *v, int32_t l0,
                   int32_t l1, int gather, void *stream) {
  if (T <= 0 || l1 <= l0) return 0;
  const int64_t cells = T * (int64_t)(l1 - l0) * lay.num_kv;
  const int vec = (int)(lay.cell_bytes >> 4);
  // lanes per cell: the largest power of two that leaves every lane 4 vectors, but never fewer than 8 lanes (gather)
  // / 16 lanes (scatter) as long as a lane keeps 2 vectors: a store instruction then writes whole 128-byte lines of a
  // cell instead of half lines.
  const int min_lanes = gather ? 8 : 16;
  int g = 32;
  while (g > 1 && g * 4 > vec) g >>= 1;
  while (g < min_lanes && g * 4 <= vec) g <<= 1;   // doubling keeps >= 2 vectors per lane
  if (g > vec) g = vec >= 1 ? (1 << (31 - __builtin_clz((unsigned)vec))) : 1;   // never more lanes than vectors
  // vectors per lane and cell, and cells per lane group and pass: V * U loads in flight per lane (KVB200_COPY_MLP: 4 or 8;
  // KVB200_COPY_ORDER: chunked or strided — tuning switches, the defaults are the measured best)
  static int mlp = 0, chunked = 0;
  if (!mlp) {
    const char *e = getenv("KVB200_COPY_MLP");
    const char *o = getenv("KVB200_COPY_ORDER");
    // measured on B200 (tools/probe_scatter.py, 80 layers, T = 65 536, scatter GB/s at cell 2048 / 1024 / 512 / 256 B):
    //   4 loads in flight, strided 6128 / 6038 / 5655 / 5515     chunked 6132 / 6051 / 5683 / 5578
    //   8 loads in flight, strided 6165 / 5940 / 5584 / 5245     chunked 6215 / 6095 / 6045 / 5835   <- default
    chunked = (o && o[0] == 's') ? 0 : 1;
    mlp = (e && atoi(e) == 4) ? 4 : 8;
  }
  const int per_lane = (vec + g - 1) / g;
  const int V = per_lane >= 4 ? 4 : (per_lane >= 2 ? 2 : 1);
  int bs_shift = -1;
  if (lay.block_size > 0 && (lay.block_size & (lay.block_size - 1)) == 0) bs_shift = __builtin_ctz((unsigned)lay.block_size);
#define KVB_COPY(GA, VV, UU) launch_kv_copy_t<GA, VV, UU>(lay, slots, T, k, v, l0, l1, g, bs_shift, chunked, cells, stream)
#define KVB_COPY_V(GA)                                                                  \
  (mlp == 8 ? (V == 4 ? KVB_COPY(GA, 4, 2) : V == 2 ? KVB_COPY(GA, 2, 4) : KVB_COPY(GA, 1, 8)) \
            : (V == 4 ? KVB_COPY(GA, 4, 1) : V == 2 ? KVB_COPY(GA, 2, 2) : KVB_COPY(GA, 1, 4)))
  return gather ? KVB_COPY_V(1) : KVB_COPY_V(0);
#undef KVB_COPY_V
#undef KVB_COPY
}

}  // namespace kvb
