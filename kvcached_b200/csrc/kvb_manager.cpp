// kvb_manager.cpp — host side of the device-resident block manager: owns the device state, packs
// request batches, launches the sm_100a step kernel, reads results back and hands the kernel's
// page map/unmap list to the VMM engine.  No bookkeeping decisions are taken here: every block id,
// page id and counter comes from the kernel (kvb_kernels.cu).  There is no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "kvb_internal.h"
#include "kvb_shm.h"
#include "kvb_tp.h"
#include "kvb_vmm.h"

using namespace kvb;

// libstdc++'s own bucket-count table (src/shared/hashtable-aux.cc, exported as
// _ZNSt8__detail12__prime_listE): uploaded to the device so that the kernel's replay of
// std::unordered_map's rehash policy uses the very numbers the reference's container uses.
namespace std { namespace __detail { extern const unsigned long __prime_list[]; } }

#define CUDA_TRY(call)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      set_error("%s failed: %s at %s:%d", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return KVB_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

namespace kvb { std::atomic<int64_t> g_total_launches{0}; }

constexpr int kFastSlots = 4;
constexpr int kFastOut = 2048;             // ids one fast alloc can return
constexpr int kFastIn = 8192;              // ids one fast free can take beyond the inline parameter block
struct FastSlot {
  kvb::Summary sum;                       // summary-carrying launches (a free of more than kInlineIds ids)
  kvb_res res;
  int64_t in_ids[kFastIn];
  uint64_t ll[kvb::LL_IDS + kFastOut];    // flagged record of a single-request launch (kvb_internal.h: LlWord)
};

namespace {
constexpr int kNumPrimes = 256 + 48;   // entries of libstdc++'s __prime_list before the sentinel (64-bit)

int64_t env_i64(const char *name, int64_t dflt) {
  const char *e = getenv(name);
  if (!e || !*e) return dflt;
  return atoll(e);
}
double env_f64(const char *name, double dflt) {
  const char *e = getenv(name);
  if (!e || !*e) return dflt;
  return atof(e);
}
int64_t pow2_ge(int64_t v) {
  int64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

struct kvb_mgr {
  kvb_mgr_config cfg{};
  int dev = 0;
  cudaStream_t stream = nullptr;
  DevState st{};
  unsigned char *d_block = nullptr;   // one allocation carved into the state arrays
  size_t d_block_bytes = 0;
  std::mutex mu;

  int64_t P = 0, B = 0, bpp = 0, mem_size = 0, max_pages = 0, f_cap = 0;
  double gpu_util = 0.95;

  // step I/O (pinned host <-> device mirrors)
  unsigned char *h_in = nullptr, *d_in = nullptr;
  size_t in_cap = 0;
  unsigned char *h_out = nullptr, *d_out = nullptr;
  size_t out_cap_bytes = 0;
  Summary last{};
  int64_t seq = 0;
  int64_t launches = 0;
  int64_t phys_override = -1;
  int64_t phys_cached = -1;        // last live get_avail_physical_pages() (cudaMemGetInfo costs ~100 us here)
  double phys_time_us = 0;
  double phys_refresh_us = 1000;
  int has_null = 0;
  int64_t total_events = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;   // bytes the step path copied over PCIe (the e2e accounting)
  bool zero_copy = true;   // step I/O: the kernel reads/writes the pinned host buffers itself (KVB200_STEP_IO=copy disables)
  std::vector<int64_t> pending_events;
  int64_t *d_scratch = nullptr;   // small device scratch (audit counters, dump sizes)
  int64_t *d_dump = nullptr;
  int64_t dump_cap = 0;

  // resident trace
  kvb_op *t_ops = nullptr;
  int64_t *t_ids = nullptr;
  kvb_ref *t_refs = nullptr;
  kvb_res *t_res = nullptr;
  int64_t *t_out = nullptr;
  int64_t t_out_cap = 0, t_nops = 0;
  Summary *t_sum = nullptr;   // device
  std::vector<int64_t> t_steps;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // resident replay as a CUDA graph: the launch sequence of a step range is captured once and replayed (one kernel
  // node per scheduler step; what changes between replays lives in device memory, not in the launch arguments)
  cudaGraphExec_t t_graph = nullptr;
  std::vector<int32_t> t_other;            // resident trace: prefix count of requests that are not alloc / free
  int64_t t_graph_key[5] = {-1, -1, -1, -1, -1};   // step_begin, step_end, steps_per_launch, phys_free, resize_target
  int64_t t_graph_launches = 0;
  int64_t phys_capture = -1;                       // >= 0: value every launch of a capture uses
  std::unique_ptr<ShmUsage> shm;   // usage publishing + limit (mem_info_tracker.hpp); null = off
  int64_t shm_publishes = 0;
  // single-request fast path: a ring of pinned slots; kvb_mgr_free launches and returns, its summary is collected by
  // the next call (or the VMM thread's idle hook)
  struct FastSlot *fast = nullptr;         // pinned host, kFastSlots entries
  struct Pending { int64_t seq; int slot; };
  std::deque<Pending> pending;             // launched, summary not yet consumed (oldest first)
  kvb_res *d_fast_res = nullptr;           // device scratch of single-request launches
  int64_t *d_fast_out = nullptr;
  bool log_open = false;                   // silent launches (frees that published nothing) since the last publishing one
  int silent_count = 0;
  int fast_next = 0;
  bool fast_on = true;                     // KVB200_FAST_PATH=0 disables
  bool async_free = true;                  // KVB200_ASYNC_FREE=0: kvb_mgr_free waits for its kernel
  int64_t n_fast = 0, n_async_free = 0;
  int tune = 0;                            // KVB200_TUNE: kernel tuning switches (LaunchArgs.tune)
  int safe_before_next = 0;                // LaunchArgs.safe_before of the next launch (kvb_mgr_trace_run sets it per launch)
  bool in_hook = false;                    // collect_pending runs on the VMM thread (idle hook)
  bool tp_on = false;              // kvb_tp_init attached this manager (rank 0) to the group's TP mailbox
  int tp_flags_next = 0;           // flags of the record the next launch may push (kTpFlagSettle)
  double tp_timeout_ms = 60000.0;
};

namespace {

int ensure_io(kvb_mgr *m, size_t in_bytes, size_t out_bytes) {
  if (in_bytes > m->in_cap) {
    size_t cap = std::max<size_t>(align_up(in_bytes * 2, 4096), 1 << 16);
    if (m->h_in) cudaFreeHost(m->h_in);
    if (m->d_in) cudaFree(m->d_in);
    m->h_in = m->d_in = nullptr;
    CUDA_TRY(cudaHostAlloc((void **)&m->h_in, cap, cudaHostAllocDefault));
    CUDA_TRY(cudaMalloc((void **)&m->d_in, cap));
    m->in_cap = cap;
  }
  if (out_bytes > m->out_cap_bytes) {
    size_t cap = std::max<size_t>(align_up(out_bytes * 2, 4096), 1 << 16);
    if (m->h_out) cudaFreeHost(m->h_out);
    if (m->d_out) cudaFree(m->d_out);
    m->h_out = m->d_out = nullptr;
    CUDA_TRY(cudaHostAlloc((void **)&m->h_out, cap, cudaHostAllocDefault));
    CUDA_TRY(cudaMalloc((void **)&m->d_out, cap));
    m->out_cap_bytes = cap;
  }
  return 0;
}

// PageAllocator::get_avail_physical_pages, csrc/page_allocator.cpp:442-455
int64_t phys_free_pages(kvb_mgr *m) {
  if (m->phys_capture >= 0) return m->phys_capture;
  if (m->phys_override >= 0) return m->phys_override;
  // live like the reference, but at most once per refresh interval; our own map/unmap activity
  // invalidates the cached value (apply_events), other tenants' activity is seen within the interval.
  double now = now_us();
  if (m->phys_cached >= 0 && now - m->phys_time_us < m->phys_refresh_us) return m->phys_cached;
  m->phys_time_us = now;
  size_t fr = 0, tot = 0;
  // with live tensors the VMM thread samples device memory; the scheduler thread stays out of the driver
  VmmGroup *gq = m->cfg.attach_vmm ? vmm_group(m->cfg.group_id, false) : nullptr;
  if (!(gq && gq->cached_meminfo(&fr, &tot))) {
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
  }
  double headroom = (double)tot * (1.0 - m->gpu_util);
  double avail = (double)fr - headroom;
  size_t page_all = (size_t)m->P * (size_t)m->cfg.num_layers * (size_t)m->cfg.num_kv_buffers;
  if (m->cfg.attach_vmm) {
    VmmGroup *g = vmm_group(m->cfg.group_id, false);
    if (g) {
      std::lock_guard<std::mutex> lk(g->mu);
      if (avail < 0 && g->pooled_pages() > 0) {
        // somebody else needs the memory: the handle pool shrinks until the headroom is restored
        int64_t give = (int64_t)((-avail + (double)page_all - 1) / (double)page_all);
        avail += (double)g->pool_release(give) * (double)page_all;
      }
      if (avail < 0) avail = 0;
      int64_t pages = (int64_t)((size_t)avail / page_all) + g->pooled_pages();   // pooled handles are memory we hold
      m->phys_cached = pages;
      return pages;
    }
  }
  if (avail < 0) avail = 0;   // the reference's size_t subtraction wraps here; we clamp
  int64_t pages = (int64_t)((size_t)avail / page_all);
  m->phys_cached = pages;
  return pages;
}

int apply_events(kvb_mgr *m, const int64_t *ev, int64_t n, bool settle = false) {
  if (n <= 0) return 0;
  m->total_events += n;
  if (m->pending_events.size() > (1u << 22)) m->pending_events.clear();
  m->pending_events.insert(m->pending_events.end(), ev, ev + n);
  if (!m->cfg.attach_vmm) return 0;
  VmmGroup *g = vmm_group(m->cfg.group_id, false);
  if (!g || !g->created()) {
    set_error("page events produced but the KV tensors of group %ld are not created", (long)m->cfg.group_id);
    return KVB_ERR_STATE;
  }
  // Maps must be in place before the step's block ids are handed out; unmaps drain behind the scheduler
  // on the group's VMM thread (KVB200_VMM_ASYNC=0: everything settles here).  With async_sched the VMM
  // thread synchronises the device before each unmap (page_allocator.cpp:671-673).
  std::lock_guard<std::mutex> lk(g->mu);
  g->set_sync_before_unmap(m->cfg.async_sched != 0);
  int rc = g->enqueue_events(ev, n);
  if (rc >= 0 && !(m->in_hook && g->async()))   // the VMM thread's own hook only queues (it cannot wait for itself)
    rc = (g->async() && !settle) ? g->wait_maps() : (settle ? g->settle() : g->fence());
  m->phys_cached = -1;   // physical memory just changed hands
  return rc < 0 ? rc : 0;
}

// The page map / unmap list of one launch: to the other TP ranks first (so that they work while this rank does),
// then to this rank's own VMM engine; when the list contains a map, wait until every rank has it in place — the ids
// the step hands out must be backed everywhere before they are used (the reference waits for every worker's reply,
// tp_ipc_util.py:160-170).
int deliver_events(kvb_mgr *m, const Summary *hs, bool settle) {
  const int64_t n = hs->n_events;
  if (n <= 0) return 0;
  std::vector<int64_t> big;
  const int64_t *ev = hs->ev_inline;
  if (n > hs->ev_inline_n) {
    big.resize((size_t)n);
    CUDA_TRY(cudaMemcpy(big.data(), m->st.ev + hs->ev_begin, (size_t)n * 8, cudaMemcpyDeviceToHost));
    ev = big.data();
  }
  TpMailbox *tp = m->tp_on ? tp_mailbox(m->cfg.group_id) : nullptr;
  bool has_map = false;
  if (tp) {
    for (int64_t i = 0; i < n && !has_map; ++i) has_map = !(ev[i] & 1);
    if (hs->tp_published) {
      tp->note_device_record();
    } else {
      int rc = tp->publish(hs->seq, ev, n, hs->next_free, settle ? 0 : hs->next_free_n, settle ? kTpFlagSettle : 0);
      if (rc) return rc;
    }
  }
  int rc = apply_events(m, ev, n, settle);
  if (rc) return rc;
  if (tp && (has_map || settle)) rc = tp->wait_applied(m->tp_timeout_ms);
  return rc;
}

// Requests kvb_decode_kernel serves (the rest of the request kinds need kvb_step_kernel).
static inline bool alloc_free_kind(int kind) {
  return kind == KVB_OP_ALLOC || kind == KVB_OP_FREE_IDS || kind == KVB_OP_FREE_REF || kind == KVB_OP_FREE_RESERVED ||
         kind == KVB_OP_AVAIL;
}

// One launch over requests [op_begin, op_end) of a device-resident request array.
int launch(kvb_mgr *m, const kvb_op *d_ops, int op_begin, int op_end, const int64_t *d_ids, const kvb_ref *d_refs,
           kvb_res *d_res, int64_t *d_outids, int64_t out_cap, Summary *d_sum, int arena_mode, int group_only,
           int64_t *grp_pages = nullptr, int64_t *grp_counts = nullptr, int64_t *grp_ids = nullptr,
           kvb_res *host_res = nullptr, int64_t *host_out = nullptr, const LaunchArgs *slots = nullptr,
           bool alloc_free_only = false) {
  LaunchArgs a{};
  if (slots) {
    a.slot_pre = slots->slot_pre; a.slot_seq = slots->slot_seq; a.slot_last = slots->slot_last;
    a.slot_out = slots->slot_out; a.slot_tokens = slots->slot_tokens; a.slot_bs = slots->slot_bs; a.slot_ps = slots->slot_ps;
  }
  a.ops = d_ops;
  a.op_begin = op_begin;
  a.op_end = op_end;
  a.ids = d_ids;
  a.refs = d_refs;
  a.res = d_res;
  a.out_ids = d_outids;
  a.out_cap = out_cap;
  a.phys_free = phys_free_pages(m);
  a.summary = d_sum;
  a.seq = ++m->seq;
  a.arena_mode = arena_mode;
  a.group_only = group_only;
  a.grp_pages = grp_pages;
  a.grp_counts = grp_counts;
  a.grp_ids = grp_ids;
  a.host_res = host_res;
  a.host_out = host_out;
  a.resize_target = (m->shm && !group_only) ? m->shm->resize_target(m->mem_size, m->cfg.num_layers, m->cfg.num_kv_buffers) : -1;
  a.tune = m->tune;
  // The kernel stages its requests BEFORE it waits for its predecessor (programmatic dependent launch): allowed only
  // where the request array is final at launch time — the resident trace, or pinned host memory the host has filled.
  // A request array that an async copy on the stream is still filling takes the plain, fully serialised launch.
  if (group_only || (arena_mode != 1 && !m->zero_copy)) a.tune |= 2;
  a.safe_before = (arena_mode == 1 && !(a.tune & 4)) ? m->safe_before_next : 0;
  m->safe_before_next = 0;
  if (m->tp_on && !group_only && arena_mode != 1 && d_sum) {
    // tensor parallel: the kernel pushes this launch's page list into the mailbox itself (kvb_tp.h); the ring
    // must have a free slot before the launch, the host learns from the summary whether it was used
    TpMailbox *tp = tp_mailbox(m->cfg.group_id);
    if (tp && tp->device_ring()) {
      int rc_room = tp->wait_room(1, m->tp_timeout_ms);
      if (rc_room) return rc_room;
      a.tp_ring = tp->device_ring();
      a.tp_published = tp->device_published();
      a.tp_pub = tp->pub();
      a.tp_mask = kTpSlots - 1;
      a.tp_slot_words = kTpSlotWords;
      a.tp_flags = m->tp_flags_next;
    }
  }
  int rc = launch_step(m->st, a, m->stream, alloc_free_only);
  if (rc != 0) {
    set_error("step kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    return KVB_ERR_CUDA;
  }
  m->launches++;
  g_total_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}


// What every completed launch owes the host side: the summary becomes the stats cache, its page list goes to the
// other TP ranks and the VMM engine, usage is published, map-ahead hints are refreshed.
int after_launch(kvb_mgr *m, const Summary *hs, bool settle) {
  m->last = *hs;
  int rc = deliver_events(m, hs, settle);
  if (rc) return rc;
  if (m->shm) {   // PageAllocator::update_memory_usage (page_allocator.cpp:688-701), once per launch
    const int64_t per_page = m->cfg.num_layers * m->P * m->cfg.num_kv_buffers;
    const int64_t *sc = m->last.sc;
    m->shm_publishes += m->shm->publish((sc[SC_NTOTAL] - sc[SC_NFREE]) * per_page, sc[SC_RL_SIZE] * per_page);
  }
  if (m->cfg.attach_vmm) {
    VmmGroup *g = vmm_group(m->cfg.group_id, false);
    if (g && g->created()) {
      std::lock_guard<std::mutex> lk(g->mu);
      if (settle) {   // memory goes back on purpose: nothing stays mapped ahead, and earlier queued unmaps settle too
        g->set_prefetch_hints(nullptr, 0);
        rc = g->settle();
        if (rc) return rc;
      } else {        // map-ahead hints: the pages the free list will hand out next
        g->set_prefetch_hints(hs->next_free, hs->next_free_n);
      }
    }
  }
  return 0;
}

// Wait (spinning on the pinned sequence word the kernel writes last) for launch `want`.
int wait_seq(kvb_mgr *m, volatile int64_t *vseq, int64_t want) {
  uint64_t spins = 0;
  while (*vseq != want) {
    if ((++spins & 0x3ff) == 0) {   // every ~1k polls make sure the kernel is still alive
      cudaError_t q = cudaStreamQuery(m->stream);
      if (q == cudaSuccess) {
        if (*vseq == want) break;
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        if (*vseq != want) { set_error("step kernel finished without publishing its summary"); return KVB_ERR_STATE; }
        break;
      }
      if (q != cudaErrorNotReady) { set_error("step kernel failed: %s", cudaGetErrorString(q)); return KVB_ERR_CUDA; }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return 0;
}

int fast_op(kvb_mgr *m, int kind, int32_t n, const int64_t *ids, bool wait, kvb_res *res, int64_t *out, int64_t cap);

// Launches that returned before their kernel finished: summary-carrying ones are consumed oldest first; frees that
// published nothing (log_open) are flushed by one tiny publishing request whose record carries their page events.
// block = false (the VMM thread's idle hook): summary-carrying launches only if they have landed.
int collect_pending(kvb_mgr *m, bool block = true) {
  while (!m->pending.empty()) {
    kvb_mgr::Pending p = m->pending.front();
    FastSlot *sl = m->fast + p.slot;
    volatile int64_t *vseq = &sl->sum.seq;
    if (!block && *vseq != p.seq) return 0;
    int rc = wait_seq(m, vseq, p.seq);
    m->pending.pop_front();
    if (rc) return rc;
    m->d2h_bytes += (int64_t)(40 * 8 + sl->sum.ev_inline_n * 8 + sizeof(kvb_res) + sl->sum.out_used * 8);
    rc = after_launch(m, &sl->sum, false);
    if (rc) return rc;
  }
  if (m->log_open) {
    kvb_res r{};
    int rc = fast_op(m, KVB_OP_AVAIL, 0, nullptr, true, &r, nullptr, 0);
    if (rc) return rc;
  }
  return 0;
}

int fast_init(kvb_mgr *m) {
  if (m->fast) return 0;
  CUDA_TRY(cudaHostAlloc((void **)&m->fast, sizeof(FastSlot) * kFastSlots, cudaHostAllocDefault));
  memset(m->fast, 0, sizeof(FastSlot) * kFastSlots);
  CUDA_TRY(cudaMalloc((void **)&m->d_fast_res, sizeof(kvb_res) * 2));
  CUDA_TRY(cudaMalloc((void **)&m->d_fast_out, (size_t)kFastOut * 8));
  return 0;
}

// One request, in the kernel parameters.
//   wait:                 the record comes back as flagged 8-byte words (payload + launch flag in one store, no
//                         system fence in the kernel); the host validates every word.
//   !wait, n <= kInlineIds: a free() that publishes NOTHING — its page events stay in the device log, which the
//                         next publishing launch reports from its start (keep_events).
//   !wait, larger:        summary-carrying launch, waited for (its ids travel through the slot's pinned buffer).
// Returns <0, or 0 with *res filled (wait) / untouched (no wait).
int fast_op(kvb_mgr *m, int kind, int32_t n, const int64_t *ids, bool wait, kvb_res *res, int64_t *out, int64_t cap) {
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  int rc = fast_init(m);
  if (rc) return rc;
  const bool big_free = kind == KVB_OP_FREE_IDS && n > kInlineIds;
  if (big_free) wait = true;
  // the event log of unpublished frees must never come near its capacity
  if (!wait && m->log_open && (int64_t)m->silent_count * (2 * m->max_pages + 8) > m->st.ev_cap / 2) wait = true;
  const int slot = m->fast_next;
  m->fast_next = (m->fast_next + 1) % kFastSlots;
  FastSlot *sl = m->fast + slot;
  LaunchArgs a{};
  a.inline_op = kvb_op{kind, n, 0};
  a.inline_on = 1;
  if (kind == KVB_OP_FREE_IDS) {
    if (!big_free) {
      memcpy(a.inline_ids, ids, (size_t)n * 8);
    } else {
      a.inline_on = 2;
      memcpy(sl->in_ids, ids, (size_t)n * 8);
      a.ids = sl->in_ids;   // pinned host memory: the kernel reads it over PCIe
    }
    m->h2d_bytes += (int64_t)n * 8;
  }
  m->h2d_bytes += (int64_t)sizeof(kvb_op);
  a.op_begin = 0;
  a.op_end = 1;
  a.res = m->d_fast_res;
  a.out_ids = m->d_fast_out;
  a.out_cap = kFastOut;
  a.phys_free = phys_free_pages(m);
  a.seq = ++m->seq;
  a.arena_mode = 0;
  a.keep_events = m->log_open ? 1 : 0;
  a.tune = m->tune;
  a.resize_target = m->shm ? m->shm->resize_target(m->mem_size, m->cfg.num_layers, m->cfg.num_kv_buffers) : -1;
  const uint32_t flag = (uint32_t)(a.seq & 0x7fffffff) | 0x80000000u;
  if (big_free) {
    a.summary = &sl->sum;
    a.host_res = &sl->res;
    reinterpret_cast<volatile int64_t &>(sl->sum.seq) = 0;
  } else if (wait) {
    a.ll = sl->ll;
    a.ll_flag = flag;
  }
  // the compact single-request kernel; a free whose ids do not fit the parameters goes through the general one
  int lrc = big_free ? launch_step(m->st, a, m->stream, true) : launch_one(m->st, a, m->stream);
  if (lrc != 0) { set_error("step kernel launch failed: %s", cudaGetErrorString((cudaError_t)lrc)); return KVB_ERR_CUDA; }
  m->launches++;
  m->n_fast++;
  g_total_launches.fetch_add(1, std::memory_order_relaxed);
  if (!wait) {          // silent: nothing to wait for, nothing to read
    m->log_open = true;
    m->silent_count++;
    m->n_async_free++;
    return 0;
  }
  m->log_open = false;  // this launch reports the whole log
  m->silent_count = 0;
  if (big_free) {
    rc = wait_seq(m, &sl->sum.seq, a.seq);
    if (rc) return rc;
    m->d2h_bytes += (int64_t)(40 * 8 + sl->sum.ev_inline_n * 8 + sizeof(kvb_res));
    rc = after_launch(m, &sl->sum, false);
    if (rc) return rc;
    *res = sl->res;
    return 0;
  }
  // ---- flagged record: word 0 first (with the liveness checks), then every word on its own
  volatile uint64_t *ll = sl->ll;
  {
    uint64_t spins = 0;
    while ((uint32_t)(ll[0] >> 32) != flag) {
      if ((++spins & 0x3ff) == 0) {
        cudaError_t q = cudaStreamQuery(m->stream);
        if (q == cudaSuccess) {
          if ((uint32_t)(ll[0] >> 32) == flag) break;
          CUDA_TRY(cudaStreamSynchronize(m->stream));
          if ((uint32_t)(ll[0] >> 32) != flag) { set_error("step kernel finished without publishing its record"); return KVB_ERR_STATE; }
          break;
        }
        if (q != cudaErrorNotReady) { set_error("step kernel failed: %s", cudaGetErrorString(q)); return KVB_ERR_CUDA; }
      }
    }
  }
  auto word = [&](int i) -> uint32_t {
    uint64_t v = ll[i];
    uint64_t spins = 0;
    while ((uint32_t)(v >> 32) != flag) {
      if ((++spins & 0xfffff) == 0 && cudaStreamQuery(m->stream) != cudaErrorNotReady) {
        v = ll[i];
        if ((uint32_t)(v >> 32) == flag) break;
        return 0xdeadbeefu;   // kernel gone, word never arrived: caught by the status check below
      }
      v = ll[i];
    }
    return (uint32_t)v;
  };
  auto word64 = [&](int i) -> int64_t { return (int64_t)((uint64_t)word(i) | ((uint64_t)word(i + 1) << 32)); };
  Summary hs{};
  hs.seq = a.seq;
  hs.ops_done = 1;
  kvb_res r{};
  r.status = (int32_t)word(LL_STATUS);
  r.n = (int32_t)word(LL_N);
  r.off = 0;
  r.avail = word64(LL_AVAIL);
  r.aux = word64(LL_AUX);
  hs.n_events = (int64_t)word(LL_NEV);
  hs.ev_begin = 0;
  hs.ev_inline_n = std::min<int64_t>(hs.n_events, kLlEvents);
  for (int i = 0; i < hs.ev_inline_n; ++i) hs.ev_inline[i] = word64(LL_EV + 2 * i);
  hs.next_free_n = (int64_t)word(LL_NHINT);
  for (int i = 0; i < 16; ++i) { uint32_t h = word(LL_HINT + i); hs.next_free[i] = h == 0xffffffffu ? -1 : (int64_t)h; }
  for (int i = 0; i < SC_COUNT; ++i) hs.sc[i] = word64(LL_SC + 2 * i);
  const int nids = (r.status == KVB_ST_OK && r.n > 0 && kind == KVB_OP_ALLOC) ? r.n : 0;
  hs.out_used = nids;
  m->d2h_bytes += (int64_t)(LL_IDS + nids) * 8;
  if (nids > 0) {
    if (!out || nids > cap) { set_error("alloc: caller buffer holds %ld ids, %d returned", (long)cap, nids); return KVB_ERR_OVERFLOW; }
    for (int i = 0; i < nids; ++i) out[i] = (int64_t)word(LL_IDS + i);
  }
  if (hs.n_events > hs.ev_inline_n) CUDA_TRY(cudaStreamSynchronize(m->stream));   // the device log is read below
  rc = after_launch(m, &hs, false);
  if (rc) return rc;
  *res = r;
  return 0;
}

bool fast_ok(const kvb_mgr *m) { return m->fast_on && !m->tp_on; }

}  // namespace

// ============================================================================================ create
extern "C" kvb_mgr *kvb_mgr_create(const kvb_mgr_config *cfg) {
  if (!cfg) { set_error("null config"); return nullptr; }
  auto *m = new kvb_mgr();
  m->cfg = *cfg;
  kvb_mgr_config &c = m->cfg;
  if (c.page_size <= 0) c.page_size = vmm_initialized() ? vmm_page_size() : (2 << 20);
  if (c.contiguous_layout < 0) c.contiguous_layout = vmm_initialized() ? (int)vmm_contiguous() : 1;
  if (c.num_kv_buffers <= 0) c.num_kv_buffers = 2;
  if (c.min_reserved_pages < 0) c.min_reserved_pages = env_i64("KVCACHED_MIN_RESERVED_PAGES", 5);
  if (c.max_reserved_pages < 0) c.max_reserved_pages = env_i64("KVCACHED_MAX_RESERVED_PAGES", 10);
  m->phys_refresh_us = env_f64("KVB200_MEMINFO_REFRESH_US", 1000);
  { const char *io = getenv("KVB200_STEP_IO"); m->zero_copy = !(io && strcmp(io, "copy") == 0); }
  m->gpu_util = c.gpu_utilization > 0 ? c.gpu_utilization : env_f64("KVCACHED_GPU_UTILIZATION", 0.95);
  m->P = c.page_size;
  m->B = c.block_size * c.cell_size;
  if (c.num_blocks <= 0 || c.block_size <= 0 || c.cell_size <= 0 || c.num_layers <= 0) {
    set_error("KVCacheManager: num_blocks, block_size, cell_size, num_layers must be positive");
    delete m;
    return nullptr;
  }
  if (m->B > m->P) {   // KVCachedConfigError, kv_cache_manager.py:104-116
    set_error("Block memory size (%ld bytes = block_size %ld x cell_size %ld) exceeds page size (%ld bytes). "
              "Set KVCACHED_PAGE_SIZE_MB to at least %ld.",
              (long)m->B, (long)c.block_size, (long)c.cell_size, (long)m->P,
              (long)(((m->B + (2 << 20) - 1) / (2 << 20)) * 2));
    delete m;
    return nullptr;
  }
  m->bpp = m->P / m->B;
  m->mem_size = c.num_blocks * m->B;            // kv_cache_manager.py:118
  m->max_pages = m->mem_size / m->P;            // page_allocator.cpp:115-116
  if (m->max_pages <= 0 || m->max_pages > (1 << 28) || c.num_blocks > (int64_t)1 << 31) {
    set_error("KVCacheManager: pool of %ld pages / %ld blocks is outside the supported range", (long)m->max_pages,
              (long)c.num_blocks);
    delete m;
    return nullptr;
  }
  m->dev = c.device_index >= 0 ? c.device_index : (vmm_initialized() ? vmm_device() : 0);
  auto fail = [&](const char *what) -> kvb_mgr * {
    if (what) set_error("%s: %s", what, cudaGetErrorString(cudaGetLastError()));
    kvb_mgr_destroy(m);
    return nullptr;
  };
  if (cudaSetDevice(m->dev) != cudaSuccess) return fail("cudaSetDevice");
  if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate");
  if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess) return fail("cudaEventCreate");

  // ---- carve the device state --------------------------------------------------------------------
  DevState &s = m->st;
  s.P = m->P;
  s.B = m->B;
  s.bpp = (int32_t)m->bpp;
  s.exact = (m->P % m->B == 0) ? 1 : 0;
  s.num_blocks = c.num_blocks;
  s.max_pages = (int32_t)m->max_pages;
  int64_t qcap = pow2_ge(m->max_pages + 2);
  s.q_mask = (int32_t)(qcap - 1);
  int64_t rbcap = pow2_ge(2 * c.num_blocks + 4);
  s.rb_mask = (int32_t)(rbcap - 1);
  s.ap_cap = (int32_t)(2 * m->max_pages + 8);
  s.min_reserved = (int32_t)std::min<int64_t>(m->max_pages, c.min_reserved_pages);   // page_allocator.cpp:117-118
  s.max_reserved = (int32_t)std::min<int64_t>(m->max_pages, c.max_reserved_pages);
  s.ev_cap = (int32_t)std::max<int64_t>(1 << 20, 4 * m->max_pages + 64);
  s.n_primes = kNumPrimes;
  m->f_cap = c.num_blocks + 1024;
  s.f_cap = m->f_cap;
  int64_t bkt_cap = 3 * std::max<int64_t>(2 * m->max_pages, m->f_cap / m->bpp + 1) + 64;
  s.g_bkt_cap = (int32_t)bkt_cap;
  int64_t nwords = (c.num_blocks + 31) / 32;

  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  size_t o_sc = carve(SC_COUNT * 8), o_fl = carve(qcap * 4), o_rl = carve(qcap * 4), o_cl = carve(qcap * 4);
  size_t o_ap = carve((size_t)s.ap_cap * 4);
  size_t o_pg[7];
  const size_t cap_pages = (size_t)m->max_pages + 1;   // ids < num_blocks can name page max_pages (never in use)
  for (auto &o : o_pg) o = carve(cap_pages * 4);
  size_t o_ring = carve((size_t)m->max_pages * m->bpp * 4);
  size_t o_bm = carve((size_t)(nwords + 8) * 4);
  size_t o_rb = carve((size_t)rbcap * 8);
  size_t o_fids = carve((size_t)m->f_cap * 8), o_fpage = carve((size_t)m->f_cap * 4);
  size_t o_g[5];
  for (auto &o : o_g) o = carve(cap_pages * 4);
  size_t o_bkt = carve((size_t)bkt_cap * 4);
  s.lru_cap = (int32_t)(2 * c.num_blocks + 2048);
  s.max_cached = (int32_t)std::max<int64_t>(-1, std::min<int64_t>(c.max_cached_blocks, INT32_MAX));
  size_t o_pref = carve((size_t)c.num_blocks * 4), o_pkey = carve((size_t)(nwords + 8) * 4);
  size_t o_lru = carve((size_t)s.lru_cap * 4), o_lpos = carve((size_t)c.num_blocks * 4);
  size_t o_plast = carve((size_t)c.num_blocks * 4), o_pin = carve((size_t)m->f_cap * 4);
  size_t o_tbits = carve((size_t)(nwords + 8) * 4);
  size_t o_ev = carve((size_t)s.ev_cap * 8);
  size_t o_pr = carve((size_t)kNumPrimes * 8);
  size_t o_gt = carve((size_t)kGrowTab * 2);
  m->d_block_bytes = off;
  if (cudaMalloc((void **)&m->d_block, off) != cudaSuccess) return fail("cudaMalloc(state)");
  unsigned char *b = m->d_block;
  s.sc = (int64_t *)(b + o_sc);
  s.fl = (int32_t *)(b + o_fl);
  s.rl = (int32_t *)(b + o_rl);
  s.cl = (int32_t *)(b + o_cl);
  s.ap = (int32_t *)(b + o_ap);
  s.pg_state = (int32_t *)(b + o_pg[0]);
  s.pg_cnt = (int32_t *)(b + o_pg[1]);
  s.pg_head = (int32_t *)(b + o_pg[2]);
  s.pg_pos = (int32_t *)(b + o_pg[3]);
  s.pg_mapped = (int32_t *)(b + o_pg[4]);
  s.pg_first = (int32_t *)(b + o_pg[5]);
  s.pg_slot = (int32_t *)(b + o_pg[6]);
  s.ring = (uint32_t *)(b + o_ring);
  s.bitmap = (uint32_t *)(b + o_bm);
  s.rb = (int64_t *)(b + o_rb);
  s.fids = (int64_t *)(b + o_fids);
  s.fpage = (int32_t *)(b + o_fpage);
  s.g_dist = (int32_t *)(b + o_g[0]);
  s.g_ord = (int32_t *)(b + o_g[1]);
  s.g_cnt = (int32_t *)(b + o_g[2]);
  s.g_next = (int32_t *)(b + o_g[3]);
  s.g_ptf = (int32_t *)(b + o_g[4]);
  s.g_bkt = (int32_t *)(b + o_bkt);
  s.pin_ref = (int32_t *)(b + o_pref);
  s.pin_keyed = (uint32_t *)(b + o_pkey);
  s.lru = (int32_t *)(b + o_lru);
  s.lru_pos = (int32_t *)(b + o_lpos);
  s.pin_last = (int32_t *)(b + o_plast);
  s.pin_in = (int32_t *)(b + o_pin);
  s.tmp_bits = (uint32_t *)(b + o_tbits);
  s.ev = (int64_t *)(b + o_ev);
  s.primes = (const uint64_t *)(b + o_pr);
  static_assert(sizeof(unsigned long) == sizeof(uint64_t), "LP64 expected");
  if (cudaMemcpy((void *)s.primes, std::__detail::__prime_list, kNumPrimes * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    return fail("cudaMemcpy(primes)");
  {
    // What libstdc++ itself answers when a table of nb buckets holding nb keys takes one more (_M_need_rehash:
    // _M_next_bkt(max(n_elt + n_ins, growth_factor * n_bkt)) at load factor 1) — asked of the real policy object, so
    // the kernels' rehash chain 2 -> 5 -> 11 -> 23 -> 47 -> ... is one table read per phase instead of a binary search.
    std::vector<uint16_t> gt((size_t)kGrowTab);
    for (int nb = 0; nb < kGrowTab; ++nb) {
      std::__detail::_Prime_rehash_policy pol;
      const size_t want = std::max<size_t>((size_t)nb + 2, (size_t)2 * nb);
      gt[(size_t)nb] = (uint16_t)pol._M_next_bkt(want);
    }
    s.grow_tab = (const uint16_t *)(b + o_gt);
    if (cudaMemcpy((void *)s.grow_tab, gt.data(), gt.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
      return fail("cudaMemcpy(grow table)");
  }
  if (cudaMalloc((void **)&m->d_scratch, 64 * 8) != cudaSuccess) return fail("cudaMalloc(scratch)");
  if (cudaMemset(m->d_scratch, 0, 64 * 8) != cudaSuccess) return fail("cudaMemset(scratch)");
#ifdef KVB_PROFILE
  s.prof = (long long *)(m->d_scratch + 32);   // upper half of the scratch block
#endif
  if (cudaMalloc((void **)&m->t_sum, sizeof(Summary)) != cudaSuccess) return fail("cudaMalloc(summary)");

  int rc = launch_init_state(s, m->max_pages, m->stream);
  if (rc != 0 || cudaStreamSynchronize(m->stream) != cudaSuccess) return fail("init kernel");
  // summary cache = initial state
  memset(&m->last, 0, sizeof(m->last));
  m->last.sc[SC_NFREE] = m->last.sc[SC_NTOTAL] = m->last.sc[SC_FL_SIZE] = m->max_pages;
  m->last.sc[SC_TARGET] = -1;

  // usage publishing (PageAllocator ctor, page_allocator.cpp:126-130): on for managers that drive real tensors,
  // or when asked for; total = mem_size * num_layers * num_kv_buffers (:120)
  {
    int want = c.publish_shm;
    if (want < 0) { const char *e = getenv("KVB200_SHM"); want = e && *e ? atoi(e) : (c.attach_vmm ? 1 : 0); }
    if (want) {
      m->shm.reset(new ShmUsage(m->mem_size * c.num_layers * c.num_kv_buffers, c.group_id, c.ipc_name[0] ? c.ipc_name : nullptr));
      if (!m->shm->ok()) m->shm.reset();   // like the reference: logged, not fatal (mem_info_tracker.hpp:218-222)
    }
  }

  m->tune = (int)env_i64("KVB200_TUNE", 0);
  m->fast_on = env_i64("KVB200_FAST_PATH", 1) != 0;
  m->async_free = env_i64("KVB200_ASYNC_FREE", 1) != 0;
  if (c.attach_vmm && vmm_initialized()) {
    // a free() that returned before its kernel finished leaves its page list to the next manager call; when the
    // engine goes idle instead, the VMM thread picks it up on its next timer wake-up (<= 50 ms)
    VmmGroup *g = vmm_group(c.group_id, false);
    if (g) g->set_idle_hook([m] {
      if ((m->pending.empty() && !m->log_open) || !m->mu.try_lock()) return;
      m->in_hook = true;
      collect_pending(m, false);
      m->in_hook = false;
      m->mu.unlock();
    });
  }

  // _post_init: reserve the null block (kv_cache_manager.py:233-243)
  if (c.reserve_null_block) {
    int64_t id = -1;
    int64_t r = kvb_mgr_alloc(m, 1, &id, 1);
    if (r != 1 || id != 0) {
      set_error("Failed to reserve null block at index 0 (got rc=%ld id=%ld)", (long)r, (long)id);
      kvb_mgr_destroy(m);
      return nullptr;
    }
    m->has_null = 1;
  }
  return m;
}

extern "C" int kvb_mgr_destroy(kvb_mgr *m) {
  if (!m) return 0;
  cudaSetDevice(m->dev);
  if (m->cfg.attach_vmm && vmm_initialized()) {
    VmmGroup *g = vmm_group(m->cfg.group_id, false);
    if (g) g->set_idle_hook(nullptr);   // returns once a running hook has finished
  }
  {
    std::lock_guard<std::mutex> lk(m->mu);
    collect_pending(m);
  }
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->cfg.attach_vmm && vmm_initialized()) {
    VmmGroup *g = vmm_group(m->cfg.group_id, false);
    if (g && g->created()) g->fence();
  }
  if (m->d_block) cudaFree(m->d_block);
  if (m->fast) cudaFreeHost(m->fast);
  if (m->d_fast_res) cudaFree(m->d_fast_res);
  if (m->d_fast_out) cudaFree(m->d_fast_out);
  if (m->h_in) cudaFreeHost(m->h_in);
  if (m->d_in) cudaFree(m->d_in);
  if (m->h_out) cudaFreeHost(m->h_out);
  if (m->d_out) cudaFree(m->d_out);
  if (m->d_scratch) cudaFree(m->d_scratch);
  if (m->d_dump) cudaFree(m->d_dump);
  if (m->t_ops) cudaFree(m->t_ops);
  if (m->t_ids) cudaFree(m->t_ids);
  if (m->t_refs) cudaFree(m->t_refs);
  if (m->t_res) cudaFree(m->t_res);
  if (m->t_out) cudaFree(m->t_out);
  if (m->t_sum) cudaFree(m->t_sum);
  if (m->t_graph) cudaGraphExecDestroy(m->t_graph);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->stream) cudaStreamDestroy(m->stream);
  cudaGetLastError();
  delete m;
  return 0;
}

// ============================================================================================ step
// Layout of the step buffers:
//   in : [ops n_ops*16][ids n_ids*8][refs n_refs*12, padded]
//   out: [Summary][res n_ops*24][ids out_cap*8]
static int64_t step_locked(kvb_mgr *m, const kvb_op *ops, int64_t n_ops, const int64_t *ids, int64_t n_ids,
                           const kvb_ref *refs, int64_t n_refs, kvb_res *res, int64_t *out_ids, int64_t out_cap) {
  if (n_ops <= 0) return 0;
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  int64_t need_out = 0;
  for (int64_t i = 0; i < n_ops; ++i) {
    if (ops[i].kind == KVB_OP_ALLOC && ops[i].n > 0) need_out += ops[i].n;
    if (ops[i].kind == KVB_OP_PIN_NEW && ops[i].n > 0) need_out += 2 * (int64_t)ops[i].n;   // evicted ids + new ids
    if (ops[i].kind == KVB_OP_PIN_FREE && ops[i].n > 0) need_out += ops[i].n;               // ids evicted by the cap
    const bool pin_list = ops[i].kind == KVB_OP_PIN_TOUCH || ops[i].kind == KVB_OP_PIN_MARK ||
                          ops[i].kind == KVB_OP_PIN_FREE || ops[i].kind == KVB_OP_PIN_EVICT;
    if (ops[i].kind == KVB_OP_FREE_IDS || ops[i].kind == KVB_OP_FREE_PAGES || pin_list) {
      if (ops[i].n < 0 || ops[i].arg < 0 || ops[i].arg + ops[i].n > n_ids) { set_error("free request %ld indexes outside the id pool", (long)i); return KVB_ERR_INVALID; }
      if (ops[i].n > m->f_cap) { set_error("free of %d ids exceeds the pool's block count", ops[i].n); return KVB_ERR_INVALID; }
    }
    if (ops[i].kind == KVB_OP_FREE_REF) {
      if (ops[i].n < 0 || ops[i].arg < 0 || ops[i].arg + ops[i].n > n_refs) { set_error("free request %ld indexes outside the ref pool", (long)i); return KVB_ERR_INVALID; }
      for (int64_t r = ops[i].arg; r < ops[i].arg + ops[i].n; ++r)
        if (refs[r].op < 0 || refs[r].op >= i) { set_error("free request %ld refers to request %d that does not precede it", (long)i, refs[r].op); return KVB_ERR_INVALID; }
    }
  }
  // trim / clear / resize hand memory back on purpose: their unmaps settle before the call returns
  bool settle = false, af_only = true;
  for (int64_t i = 0; i < n_ops; ++i) {
    settle |= ops[i].kind == KVB_OP_TRIM || ops[i].kind == KVB_OP_CLEAR || ops[i].kind == KVB_OP_RESIZE ||
              ops[i].kind == KVB_OP_PA_RESIZE;
    af_only &= alloc_free_kind(ops[i].kind);
  }
  if (need_out > out_cap) { set_error("out_ids capacity %ld < %ld ids requested (alloc n, get_new_blocks 2n, free_blocks n)", (long)out_cap, (long)need_out); return KVB_ERR_OVERFLOW; }
  size_t b_ops = (size_t)n_ops * sizeof(kvb_op), b_ids = (size_t)n_ids * 8, b_refs = align_up((size_t)n_refs * sizeof(kvb_ref), 8);
  size_t in_bytes = b_ops + b_ids + b_refs;
  size_t b_res = (size_t)n_ops * sizeof(kvb_res);
  size_t o_res = align_up(sizeof(Summary), 16), o_ids = o_res + align_up(b_res, 16);
  size_t out_bytes = o_ids + (size_t)need_out * 8;
  int rc = ensure_io(m, in_bytes, out_bytes);
  if (rc) return rc;
  memcpy(m->h_in, ops, b_ops);
  if (n_ids) memcpy(m->h_in + b_ops, ids, b_ids);
  if (n_refs) memcpy(m->h_in + b_ops + b_ids, refs, (size_t)n_refs * sizeof(kvb_ref));
  const bool zc = m->zero_copy;
  // copy mode: one H2D copy in, one D2H copy out.  zero-copy mode: the kernel reads the pinned input
  // buffer over PCIe itself and writes results + summary straight into the pinned output buffer; the
  // host polls the summary's sequence number instead of synchronising the stream.
  const unsigned char *in_base = zc ? m->h_in : m->d_in;
  if (!zc) CUDA_TRY(cudaMemcpyAsync(m->d_in, m->h_in, in_bytes, cudaMemcpyHostToDevice, m->stream));
  m->h2d_bytes += (int64_t)in_bytes;
  const kvb_op *d_ops = (const kvb_op *)in_base;
  const int64_t *d_ids = (const int64_t *)(in_base + b_ops);
  const kvb_ref *d_refs = (const kvb_ref *)(in_base + b_ops + b_ids);
  Summary *d_sum = zc ? (Summary *)m->h_out : (Summary *)m->d_out;
  kvb_res *d_res = (kvb_res *)(m->d_out + o_res);
  int64_t *d_outids = (int64_t *)(m->d_out + o_ids);
  kvb_res *h_res = (kvb_res *)(m->h_out + o_res);
  int64_t *h_outids = (int64_t *)(m->h_out + o_ids);

  int64_t done = 0;
  bool first = true;
  m->tp_flags_next = settle ? kTpFlagSettle : 0;
  while (done < n_ops) {
    volatile int64_t *vseq = &((Summary *)m->h_out)->seq;
    if (zc) *vseq = 0;
    rc = launch(m, d_ops, (int)done, (int)n_ops, d_ids, d_refs, d_res, d_outids, need_out, d_sum, first ? 0 : 2, 0,
                nullptr, nullptr, nullptr, zc ? h_res : nullptr, zc ? h_outids : nullptr, nullptr, af_only);
    if (rc) return rc;
    if (zc) {
      const int64_t want = m->seq;
      uint64_t spins = 0;
      while (*vseq != want) {
        if ((++spins & 0x3ff) == 0) {   // every ~1k polls make sure the kernel is still alive
          cudaError_t q = cudaStreamQuery(m->stream);
          if (q == cudaSuccess) {
            if (*vseq == want) break;
            CUDA_TRY(cudaStreamSynchronize(m->stream));
            if (*vseq != want) { set_error("step kernel finished without publishing its summary"); return KVB_ERR_STATE; }
            break;
          }
          if (q != cudaErrorNotReady) { set_error("step kernel failed: %s", cudaGetErrorString(q)); return KVB_ERR_CUDA; }
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
    } else {
      CUDA_TRY(cudaMemcpyAsync(m->h_out, m->d_out, out_bytes, cudaMemcpyDeviceToHost, m->stream));
      CUDA_TRY(cudaStreamSynchronize(m->stream));
    }
    const Summary *hs = (const Summary *)m->h_out;
    m->last = *hs;
    m->d2h_bytes += zc ? (int64_t)(40 * 8 + hs->ev_inline_n * 8 + hs->ops_done * (int64_t)sizeof(kvb_res) + hs->out_used * 8)
                       : (int64_t)out_bytes;
    // page map/unmap list of this launch -> the other TP ranks and this rank's VMM engine
    rc = deliver_events(m, hs, settle);
    if (rc) return rc;
    if (hs->ops_done <= 0) { set_error("step kernel made no progress at request %ld", (long)done); return KVB_ERR_STATE; }
    done += hs->ops_done;
    first = false;
  }
  if (m->shm) {   // PageAllocator::update_memory_usage (page_allocator.cpp:688-701), once per step
    const int64_t per_page = m->cfg.num_layers * m->P * m->cfg.num_kv_buffers;
    const int64_t *sc = m->last.sc;
    m->shm_publishes += m->shm->publish((sc[SC_NTOTAL] - sc[SC_NFREE]) * per_page, sc[SC_RL_SIZE] * per_page);
  }
  if (m->cfg.attach_vmm) {
    VmmGroup *g = vmm_group(m->cfg.group_id, false);
    if (g && g->created()) {
      std::lock_guard<std::mutex> lk(g->mu);
      const Summary *hs2 = (const Summary *)m->h_out;
      if (settle) {   // memory goes back on purpose: nothing stays mapped ahead, and earlier queued unmaps settle too
        g->set_prefetch_hints(nullptr, 0);
        rc = g->settle();
        if (rc) return rc;
      } else {        // map-ahead hints: the pages the free list will hand out next
        g->set_prefetch_hints(hs2->next_free, hs2->next_free_n);
      }
    }
  }
  memcpy(res, m->h_out + o_res, b_res);
  const Summary *hs = (const Summary *)m->h_out;
  int64_t used = hs->sc[SC_OUT_CURSOR];
  if (used > 0) memcpy(out_ids, m->h_out + o_ids, (size_t)used * 8);
  return used;
}

extern "C" int64_t kvb_mgr_step(kvb_mgr *m, const kvb_op *ops, int64_t n_ops, const int64_t *ids, int64_t n_ids,
                                const kvb_ref *refs, int64_t n_refs, kvb_res *res, int64_t *out_ids,
                                int64_t out_cap) {
  if (!m || (n_ops > 0 && (!ops || !res))) { set_error("kvb_mgr_step: null argument"); return KVB_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(m->mu);
  return step_locked(m, ops, n_ops, ids, n_ids, refs, n_refs, res, out_ids, out_cap);
}

// ---- single-request wrappers ---------------------------------------------------------------------------
static int one_op(kvb_mgr *m, int kind, int32_t n, int64_t arg, const int64_t *ids, int64_t n_ids, kvb_res *r,
                  int64_t *out, int64_t cap) {
  kvb_op o{kind, n, arg};
  std::lock_guard<std::mutex> lk(m->mu);
  int64_t rc = step_locked(m, &o, 1, ids, n_ids, nullptr, 0, r, out, cap);
  return rc < 0 ? (int)rc : 0;
}

extern "C" int64_t kvb_mgr_alloc(kvb_mgr *m, int64_t n, int64_t *out, int64_t cap) {
  if (!m || n < 0 || n > INT32_MAX) { set_error("kvb_mgr_alloc: bad argument"); return KVB_ERR_INVALID - 100; }
  kvb_res r{};
  int rc;
  if (n <= kFastOut && fast_ok(m)) {   // one request in the kernel parameters, result straight into a pinned slot
    std::lock_guard<std::mutex> lk(m->mu);
    rc = fast_op(m, KVB_OP_ALLOC, (int32_t)n, nullptr, true, &r, out, cap);
    if (rc == 0 && r.status == KVB_ST_RETRY) {   // not a case for the compact kernel: nothing was touched
      kvb_op o{KVB_OP_ALLOC, (int32_t)n, 0};
      int64_t rs = step_locked(m, &o, 1, nullptr, 0, nullptr, 0, &r, out, cap);
      rc = rs < 0 ? (int)rs : 0;
    }
  } else {
    rc = one_op(m, KVB_OP_ALLOC, (int32_t)n, 0, nullptr, 0, &r, out, cap);
  }
  if (rc) return rc - 100;
  if (r.status == KVB_ST_OK) return r.n;
  return r.status;
}
extern "C" int kvb_mgr_free(kvb_mgr *m, const int64_t *ids, int64_t n) {
  if (!m || n < 0 || (n > 0 && !ids)) { set_error("kvb_mgr_free: bad argument"); return KVB_ERR_INVALID; }
  if (n == 0) return 0;   // kv_cache_manager.py:310-311
  kvb_res r{};
  if (n <= kFastIn && n <= m->f_cap && fast_ok(m)) {
    // free() returns nothing, so it does not wait for its kernel (stream order keeps the sequence); the launch's
    // summary — counters, page unmaps — is consumed by the next call on this manager, or by the VMM thread when idle
    std::lock_guard<std::mutex> lk(m->mu);
    int rc = fast_op(m, KVB_OP_FREE_IDS, (int32_t)n, ids, !m->async_free, &r, nullptr, 0);
    if (rc == 0 && r.status == KVB_ST_RETRY) {   // (only a waiting free can be told to retry)
      kvb_op o{KVB_OP_FREE_IDS, (int32_t)n, 0};
      int64_t rs = step_locked(m, &o, 1, ids, n, nullptr, 0, &r, nullptr, 0);
      rc = rs < 0 ? (int)rs : 0;
    }
    return rc;
  }
  return one_op(m, KVB_OP_FREE_IDS, (int32_t)n, 0, ids, n, &r, nullptr, 0);
}
extern "C" int kvb_mgr_resize(kvb_mgr *m, int64_t new_mem_size) {
  if (!m || new_mem_size <= 0) { set_error("new_mem_size must be positive"); return KVB_ERR_INVALID; }
  kvb_res r{};
  int rc = one_op(m, KVB_OP_RESIZE, 0, new_mem_size, nullptr, 0, &r, nullptr, 0);
  return rc ? rc - 100 : r.status;
}
extern "C" int kvb_mgr_trim(kvb_mgr *m) {
  if (!m) return KVB_ERR_INVALID;
  kvb_res r{};
  return one_op(m, KVB_OP_TRIM, 0, 0, nullptr, 0, &r, nullptr, 0);
}
extern "C" int kvb_mgr_try_to_reserve(kvb_mgr *m, int64_t n) {
  if (!m || n < 0 || n > INT32_MAX) return KVB_ERR_INVALID - 100;
  kvb_res r{};
  int rc = one_op(m, KVB_OP_TRY_RESERVE, (int32_t)n, 0, nullptr, 0, &r, nullptr, 0);
  return rc ? rc - 100 : r.status;
}
extern "C" int kvb_mgr_free_reserved(kvb_mgr *m) {
  if (!m) return KVB_ERR_INVALID;
  kvb_res r{};
  return one_op(m, KVB_OP_FREE_RESERVED, 0, 0, nullptr, 0, &r, nullptr, 0);
}
extern "C" int kvb_mgr_clear(kvb_mgr *m) {
  if (!m) return KVB_ERR_INVALID;
  kvb_op ops[2] = {{KVB_OP_CLEAR, 0, 0}, {KVB_OP_ALLOC, 1, 0}};
  kvb_res r[2]{};
  int64_t id = -1;
  std::lock_guard<std::mutex> lk(m->mu);
  m->has_null = 0;
  int64_t rc = step_locked(m, ops, m->cfg.reserve_null_block ? 2 : 1, nullptr, 0, nullptr, 0, r, &id, 1);
  if (rc < 0) return (int)rc;
  if (m->cfg.reserve_null_block) {
    if (r[1].status != KVB_ST_OK || id != 0) { set_error("Failed to reserve null block at index 0"); return KVB_ST_NULL_BLOCK; }
    m->has_null = 1;
  }
  return 0;
}

extern "C" int64_t kvb_mgr_available_size(kvb_mgr *m) {
  if (!m) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  const int64_t *sc = m->last.sc;
  int64_t a = sc[SC_NAB] + sc[SC_NRB];
  if (sc[SC_INSHRINK]) return a;
  ensure_device(m->dev);
  int64_t phys = phys_free_pages(m) + sc[SC_RL_SIZE];
  return a + std::min(sc[SC_NFREE], phys) * m->bpp;
}

extern "C" int kvb_mgr_get_stats(kvb_mgr *m, kvb_mgr_stats *o) {
  if (!m || !o) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  const int64_t *sc = m->last.sc;
  o->num_free_pages = sc[SC_NFREE];
  o->num_total_pages = sc[SC_NTOTAL];
  o->num_inuse_pages = sc[SC_NTOTAL] - sc[SC_NFREE];
  o->num_reserved_pages = sc[SC_RL_SIZE];
  o->num_avail_blocks = sc[SC_NAB];
  o->num_reserved_blocks = sc[SC_NRB];
  o->in_shrink = sc[SC_INSHRINK];
  o->target_num_blocks = sc[SC_TARGET];
  o->num_avail_pages = sc[SC_NAVAILPG];
  o->num_full_pages = sc[SC_NFULLPG];
  o->num_mapped_pages = sc[SC_NMAPPED];
  o->has_null_block = m->has_null;
  o->blocks_per_page = m->bpp;
  o->block_mem_size = m->B;
  o->page_size = m->P;
  o->mem_size = m->mem_size;
  o->bad_frees = sc[SC_BAD_FREES];
  o->unknown_page_frees = sc[SC_UNKNOWN_FREES];
  o->launches = m->launches;
  o->events = m->total_events;
  o->h2d_bytes = m->h2d_bytes;
  o->d2h_bytes = m->d2h_bytes;
  o->num_evictable_blocks = sc[SC_LRU_COUNT];
  return 0;
}

// Everything launched so far has completed and its page list has reached the VMM engine (kvb_mgr_free returns before
// its kernel does).
extern "C" int kvb_mgr_sync(kvb_mgr *m) {
  if (!m) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  return collect_pending(m);
}

extern "C" int kvb_mgr_set_phys_free_pages(kvb_mgr *m, int64_t n) {
  if (!m) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  m->phys_override = n;
  return 0;
}

// ============================================================================================ SGLang slot requests
// One request whose inputs / outputs are the caller's DEVICE arrays; the host only learns status and counts.
static int slots_locked(kvb_mgr *m, int kind, const LaunchArgs &sl, int64_t arg, int64_t arena_cap, void *stream, kvb_res *res,
                        int32_t op_n = 0) {
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  size_t o_res = align_up(sizeof(Summary), 16), o_ids = o_res + 64;
  int rc = collect_pending(m);
  if (rc) return rc;
  rc = ensure_io(m, sizeof(kvb_op), o_ids + (size_t)arena_cap * 8);
  if (rc) return rc;
  kvb_op op{kind, op_n, arg};
  memcpy(m->h_in, &op, sizeof(op));
  m->tp_flags_next = 0;
  // order our launch after whatever produced the arrays on the caller's stream
  CUDA_TRY(cudaEventRecord(m->ev0, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev0, 0));
  volatile int64_t *vseq = &((Summary *)m->h_out)->seq;
  *vseq = 0;
  m->h2d_bytes += (int64_t)sizeof(kvb_op);
  rc = launch(m, (const kvb_op *)m->h_in, 0, 1, nullptr, nullptr, (kvb_res *)(m->d_out + o_res), (int64_t *)(m->d_out + o_ids),
              arena_cap, (Summary *)m->h_out, 0, 0, nullptr, nullptr, nullptr, (kvb_res *)(m->h_out + o_res),
              (int64_t *)(m->h_out + o_ids), &sl);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(m->stream));   // the index array is complete when the call returns (host-ordered for the caller)
  if (*vseq != m->seq) { set_error("slot request: kernel finished without publishing its summary"); return KVB_ERR_STATE; }
  const Summary *hs = (const Summary *)m->h_out;
  m->last = *hs;
  m->d2h_bytes += (int64_t)(40 * 8 + hs->ev_inline_n * 8 + sizeof(kvb_res) + hs->out_used * 8);
  rc = deliver_events(m, hs, false);
  if (rc) return rc;
  if (m->shm) {
    const int64_t per_page = m->cfg.num_layers * m->P * m->cfg.num_kv_buffers;
    const int64_t *sc = m->last.sc;
    m->shm_publishes += m->shm->publish((sc[SC_NTOTAL] - sc[SC_NFREE]) * per_page, sc[SC_RL_SIZE] * per_page);
  }
  if (m->cfg.attach_vmm) {
    VmmGroup *g = vmm_group(m->cfg.group_id, false);
    if (g && g->created()) {
      std::lock_guard<std::mutex> lk(g->mu);
      g->set_prefetch_hints(hs->next_free, hs->next_free_n);
    }
  }
  memcpy(res, m->h_out + o_res, sizeof(kvb_res));
  return 0;
}

static int alloc_slots(kvb_mgr *m, int kind, int64_t bs, uint64_t pre, uint64_t seq, uint64_t last, int64_t tokens,
                       uint64_t out, void *stream, int64_t *num_new_pages) {
  if (!m || bs <= 0 || bs > m->f_cap || !seq || !last || !out || (kind == KVB_OP_ALLOC_EXTEND && (!pre || tokens < 0))) {
    set_error("alloc_extend / alloc_decode: bad argument");
    return KVB_ERR_INVALID;
  }
  LaunchArgs sl{};
  sl.slot_pre = (const int64_t *)(uintptr_t)pre;
  sl.slot_seq = (const int64_t *)(uintptr_t)seq;
  sl.slot_last = (const int64_t *)(uintptr_t)last;
  sl.slot_out = (int64_t *)(uintptr_t)out;
  sl.slot_tokens = kind == KVB_OP_ALLOC_DECODE ? bs : tokens;
  sl.slot_bs = (int32_t)bs;
  sl.slot_ps = (int32_t)m->cfg.block_size;
  // new pages <= extend tokens / page_size + one partial page per sequence
  int64_t arena = (kind == KVB_OP_ALLOC_DECODE ? 0 : tokens / m->cfg.block_size) + bs + 8;
  kvb_res r{};
  std::lock_guard<std::mutex> lk(m->mu);
  int rc = slots_locked(m, kind, sl, 0, arena, stream, &r);
  if (rc) return rc;
  if (num_new_pages) *num_new_pages = r.n;
  if (r.status == KVB_ST_OK) return 1;
  if (r.status == KVB_ERR_OVERFLOW) { set_error("alloc_extend: the sequences need %ld slot indices, the buffer holds %ld", (long)r.aux, (long)sl.slot_tokens); return KVB_ERR_OVERFLOW; }
  return r.status;
}
extern "C" int kvb_mgr_alloc_extend(kvb_mgr *m, int64_t bs, uint64_t prefix_lens, uint64_t seq_lens, uint64_t last_loc,
                                    int64_t extend_num_tokens, uint64_t out_indices, void *stream, int64_t *num_new_pages) {
  return alloc_slots(m, KVB_OP_ALLOC_EXTEND, bs, prefix_lens, seq_lens, last_loc, extend_num_tokens, out_indices, stream, num_new_pages);
}
extern "C" int kvb_mgr_alloc_decode(kvb_mgr *m, int64_t bs, uint64_t seq_lens, uint64_t last_loc, uint64_t out_indices,
                                    void *stream, int64_t *num_new_pages) {
  return alloc_slots(m, KVB_OP_ALLOC_DECODE, bs, 0, seq_lens, last_loc, bs, out_indices, stream, num_new_pages);
}
extern "C" int64_t kvb_mgr_free_slots(kvb_mgr *m, uint64_t free_index, int64_t n, void *stream) {
  if (!m || n < 0 || (n > 0 && !free_index)) { set_error("free_slots: bad argument"); return KVB_ERR_INVALID; }
  if (n == 0) return 0;   // patches.py:283-284
  LaunchArgs sl{};
  sl.slot_pre = (const int64_t *)(uintptr_t)free_index;
  sl.slot_ps = (int32_t)m->cfg.block_size;
  kvb_res r{};
  std::lock_guard<std::mutex> lk(m->mu);
  int rc = slots_locked(m, KVB_OP_FREE_SLOTS, sl, n, 8, stream, &r);
  if (rc) return rc;
  return r.n;
}

extern "C" int64_t kvb_mgr_alloc_device(kvb_mgr *m, int64_t n, uint64_t out_ids_dev, void *stream) {
  if (!m || n < 0 || n > INT32_MAX || (n > 0 && !out_ids_dev)) { set_error("alloc_device: bad argument"); return KVB_ERR_INVALID - 100; }
  LaunchArgs sl{};
  sl.slot_out = (int64_t *)(uintptr_t)out_ids_dev;
  sl.slot_tokens = n;
  sl.slot_ps = (int32_t)m->cfg.block_size;
  kvb_res r{};
  std::lock_guard<std::mutex> lk(m->mu);
  int rc = slots_locked(m, KVB_OP_ALLOC_DEV, sl, 0, n + 8, stream, &r, (int32_t)n);
  if (rc) return rc - 100;
  return r.status == KVB_ST_OK ? r.n : r.status;
}
extern "C" int kvb_mgr_free_device(kvb_mgr *m, uint64_t ids_dev, int64_t n, void *stream) {
  if (!m || n < 0 || (n > 0 && !ids_dev)) { set_error("free_device: bad argument"); return KVB_ERR_INVALID; }
  if (n == 0) return 0;
  LaunchArgs sl{};
  sl.slot_pre = (const int64_t *)(uintptr_t)ids_dev;
  sl.slot_ps = (int32_t)m->cfg.block_size;
  kvb_res r{};
  std::lock_guard<std::mutex> lk(m->mu);
  int rc = slots_locked(m, KVB_OP_FREE_DEV, sl, n, 8, stream, &r);
  if (rc) return rc;
  return r.status == KVB_ST_OK ? 0 : r.status;
}

// ============================================================================================ resident traces
extern "C" int kvb_mgr_trace_upload(kvb_mgr *m, const kvb_op *ops, int64_t n_ops, const int64_t *ids, int64_t n_ids,
                                    const kvb_ref *refs, int64_t n_refs, const int64_t *step_offsets, int64_t n_steps,
                                    int64_t out_capacity) {
  if (!m || !ops || n_ops <= 0 || !step_offsets || n_steps <= 0) { set_error("trace_upload: bad argument"); return KVB_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  for (void *p : {(void *)m->t_ops, (void *)m->t_ids, (void *)m->t_refs, (void *)m->t_res, (void *)m->t_out})
    if (p) cudaFree(p);
  m->t_ops = nullptr; m->t_ids = nullptr; m->t_refs = nullptr; m->t_res = nullptr; m->t_out = nullptr;
  CUDA_TRY(cudaMalloc((void **)&m->t_ops, (size_t)n_ops * sizeof(kvb_op)));
  CUDA_TRY(cudaMalloc((void **)&m->t_ids, std::max<size_t>((size_t)n_ids * 8, 8)));
  CUDA_TRY(cudaMalloc((void **)&m->t_refs, std::max<size_t>((size_t)n_refs * sizeof(kvb_ref), 16)));
  CUDA_TRY(cudaMalloc((void **)&m->t_res, (size_t)n_ops * sizeof(kvb_res)));
  CUDA_TRY(cudaMalloc((void **)&m->t_out, std::max<size_t>((size_t)out_capacity * 8, 8)));
  CUDA_TRY(cudaMemcpy(m->t_ops, ops, (size_t)n_ops * sizeof(kvb_op), cudaMemcpyHostToDevice));
  if (n_ids) CUDA_TRY(cudaMemcpy(m->t_ids, ids, (size_t)n_ids * 8, cudaMemcpyHostToDevice));
  if (n_refs) CUDA_TRY(cudaMemcpy(m->t_refs, refs, (size_t)n_refs * sizeof(kvb_ref), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(m->t_res, 0, (size_t)n_ops * sizeof(kvb_res)));
  if (m->t_graph) { cudaGraphExecDestroy(m->t_graph); m->t_graph = nullptr; }
  m->t_out_cap = out_capacity;
  m->t_nops = n_ops;
  m->t_steps.assign(step_offsets, step_offsets + n_steps + 1);
  m->t_other.assign((size_t)n_ops + 1, 0);     // prefix count of requests kvb_decode_kernel does not serve
  for (int64_t i = 0; i < n_ops; ++i) m->t_other[(size_t)i + 1] = m->t_other[(size_t)i] + (alloc_free_kind(ops[i].kind) ? 0 : 1);
  if (m->t_steps.front() != 0 || m->t_steps.back() != n_ops) { set_error("trace_upload: step_offsets must span [0, n_ops]"); return KVB_ERR_INVALID; }
  return 0;
}

extern "C" int kvb_mgr_trace_reset_cursor(kvb_mgr *m) {
  if (!m) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  CUDA_TRY(cudaMemsetAsync(m->st.sc + SC_OUT_CURSOR, 0, 8, m->stream));
  CUDA_TRY(cudaMemsetAsync(m->st.sc + SC_NEV, 0, 8, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  return 0;
}

extern "C" int kvb_mgr_trace_run(kvb_mgr *m, int64_t step_begin, int64_t step_end, int steps_per_launch,
                                 float *device_ms, int64_t *n_launches) {
  if (!m || !m->t_ops) { set_error("trace_run: no trace uploaded"); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(m->mu);
  int64_t n_steps = (int64_t)m->t_steps.size() - 1;
  if (step_begin < 0 || step_end > n_steps || step_begin > step_end) { set_error("trace_run: bad step range"); return KVB_ERR_INVALID; }
  if (steps_per_launch <= 0) steps_per_launch = 1;
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  int64_t launches0 = m->launches;
  auto run_range = [&]() -> int {
    // Results of requests below the PREVIOUS launch's first request are complete when a launch becomes resident (its
    // predecessor lets it in only after having waited for its own predecessor); the first launch of the range follows
    // a memset / a synchronised stream, so everything before it is complete.
    int64_t prev_ob = -1;
    // The summary (counters, event count) is the range's closing report: only its last launch writes one — the earlier
    // ones would be overwritten unread.  Per-request results and ids are written by every launch.
    int64_t last_s = -1;
    for (int64_t s = step_begin; s < step_end; s += steps_per_launch) {
      int64_t e = std::min<int64_t>(s + steps_per_launch, step_end);
      if (m->t_steps[(size_t)e] != m->t_steps[(size_t)s]) last_s = s;
    }
    for (int64_t s = step_begin; s < step_end; s += steps_per_launch) {
      int64_t e = std::min<int64_t>(s + steps_per_launch, step_end);
      int64_t ob = m->t_steps[(size_t)s], oe = m->t_steps[(size_t)e];
      if (oe == ob) continue;
      m->safe_before_next = (int)(prev_ob < 0 ? ob : prev_ob);
      const bool af_only = m->t_other[(size_t)oe] == m->t_other[(size_t)ob];
      int rc = launch(m, m->t_ops, (int)ob, (int)oe, m->t_ids, m->t_refs, m->t_res, m->t_out, m->t_out_cap,
                      s == last_s ? m->t_sum : nullptr, 1, 0,
                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, af_only);
      if (rc) return rc;
      prev_ob = ob;
    }
    return 0;
  };
  // Launch-bound loop -> CUDA graph (KVB200_TRACE_GRAPH=0 disables): captured once per (range, physical budget, limit)
  const bool use_graph = env_i64("KVB200_TRACE_GRAPH", 1) != 0 && (step_end - step_begin) / steps_per_launch >= 8;
  const int64_t phys_now = phys_free_pages(m);
  const int64_t rt_now = m->shm ? m->shm->resize_target(m->mem_size, m->cfg.num_layers, m->cfg.num_kv_buffers) : -1;
  const int64_t key[5] = {step_begin, step_end, steps_per_launch, phys_now, rt_now};
  bool have_graph = use_graph && m->t_graph && memcmp(key, m->t_graph_key, sizeof(key)) == 0;
  if (use_graph && !have_graph) {
    if (m->t_graph) { cudaGraphExecDestroy(m->t_graph); m->t_graph = nullptr; }
    launch_step_prepare();
    cudaGraph_t g = nullptr;
    m->phys_capture = phys_now;
    const int64_t l0 = m->launches;
    if (cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int rc = run_range();
      cudaError_t ce = cudaStreamEndCapture(m->stream, &g);
      if (rc == 0 && ce == cudaSuccess && g && cudaGraphInstantiate(&m->t_graph, g, 0) == cudaSuccess) {
        memcpy(m->t_graph_key, key, sizeof(key));
        m->t_graph_launches = m->launches - l0;
        have_graph = true;
      } else {
        m->t_graph = nullptr;
      }
      if (g) cudaGraphDestroy(g);
    }
    cudaGetLastError();
    m->phys_capture = -1;
    g_total_launches.fetch_sub(m->launches - l0, std::memory_order_relaxed);   // captured, not executed
    m->launches = l0;
  }
  // events of the whole range accumulate in the device log
  CUDA_TRY(cudaMemsetAsync(m->st.sc + SC_NEV, 0, 8, m->stream));
  CUDA_TRY(cudaEventRecord(m->ev0, m->stream));
  if (have_graph) {
    CUDA_TRY(cudaGraphLaunch(m->t_graph, m->stream));
    m->launches += m->t_graph_launches;
    g_total_launches.fetch_add(m->t_graph_launches, std::memory_order_relaxed);
  } else {
    int rc = run_range();
    if (rc) return rc;
  }
  CUDA_TRY(cudaEventRecord(m->ev1, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (device_ms) CUDA_TRY(cudaEventElapsedTime(device_ms, m->ev0, m->ev1));
  if (n_launches) *n_launches = m->launches - launches0;
  if (m->launches == launches0) return 0;
  CUDA_TRY(cudaMemcpy(&m->last, m->t_sum, sizeof(Summary), cudaMemcpyDeviceToHost));
  int64_t nev = m->last.sc[SC_NEV];
  if (nev > m->st.ev_cap) { set_error("trace_run: page event log overflow (%ld > %d)", (long)nev, m->st.ev_cap); return KVB_ERR_OVERFLOW; }
  if (m->last.sc[SC_OUT_CURSOR] > m->t_out_cap) { set_error("trace_run: id arena overflow"); return KVB_ERR_OVERFLOW; }
  if (nev > 0) {
    std::vector<int64_t> ev((size_t)nev);
    CUDA_TRY(cudaMemcpy(ev.data(), m->st.ev, (size_t)nev * 8, cudaMemcpyDeviceToHost));
    int rc = apply_events(m, ev.data(), nev);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int kvb_mgr_trace_results(kvb_mgr *m, kvb_res *res, int64_t n_ops, int64_t *out_ids, int64_t out_cap,
                                     int64_t *n_out) {
  if (!m || !m->t_ops) { set_error("trace_results: no trace uploaded"); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  if (n_ops > m->t_nops) n_ops = m->t_nops;
  if (res && n_ops > 0) CUDA_TRY(cudaMemcpy(res, m->t_res, (size_t)n_ops * sizeof(kvb_res), cudaMemcpyDeviceToHost));
  int64_t cur = 0;
  CUDA_TRY(cudaMemcpy(&cur, m->st.sc + SC_OUT_CURSOR, 8, cudaMemcpyDeviceToHost));
  if (n_out) *n_out = cur;
  if (out_ids) {
    if (cur > out_cap) { set_error("trace_results: out_cap %ld < %ld", (long)out_cap, (long)cur); return KVB_ERR_OVERFLOW; }
    if (cur > 0) CUDA_TRY(cudaMemcpy(out_ids, m->t_out, (size_t)cur * 8, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ============================================================================================ inspection
static int ensure_dump(kvb_mgr *m, int64_t n) {
  if (n <= m->dump_cap) return 0;
  if (m->d_dump) cudaFree(m->d_dump);
  m->d_dump = nullptr;
  CUDA_TRY(cudaMalloc((void **)&m->d_dump, (size_t)n * 8));
  m->dump_cap = n;
  return 0;
}

extern "C" int64_t kvb_mgr_dump_list(kvb_mgr *m, int which, int64_t *out, int64_t cap) {
  if (!m || !out) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  int64_t sc[SC_COUNT];
  CUDA_TRY(cudaMemcpy(sc, m->st.sc, sizeof(sc), cudaMemcpyDeviceToHost));
  std::vector<int64_t> v;
  if (which >= 0 && which <= 2) {
    const int32_t *q = which == 0 ? m->st.fl : which == 1 ? m->st.rl : m->st.cl;
    int64_t head = sc[which == 0 ? SC_FL_HEAD : which == 1 ? SC_RL_HEAD : SC_CL_HEAD];
    int64_t size = sc[which == 0 ? SC_FL_SIZE : which == 1 ? SC_RL_SIZE : SC_CL_SIZE];
    std::vector<int32_t> h((size_t)m->st.q_mask + 1);
    CUDA_TRY(cudaMemcpy(h.data(), q, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < size; ++i) v.push_back(h[(size_t)((head + i) & m->st.q_mask)]);
  } else if (which == 3) {
    int rc = ensure_dump(m, m->max_pages + 1);
    if (rc) return rc;
    if (launch_dump_avail(m->st, m->d_dump, m->d_scratch, m->stream) != 0) { set_error("dump kernel launch failed"); return KVB_ERR_CUDA; }
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    int64_t n = 0;
    CUDA_TRY(cudaMemcpy(&n, m->d_scratch, 8, cudaMemcpyDeviceToHost));
    v.resize((size_t)n);
    if (n) CUDA_TRY(cudaMemcpy(v.data(), m->d_dump, (size_t)n * 8, cudaMemcpyDeviceToHost));
  } else if (which == 4 || which == 6) {
    std::vector<int32_t> h((size_t)m->max_pages);
    CUDA_TRY(cudaMemcpy(h.data(), which == 4 ? m->st.pg_state : m->st.pg_mapped, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < m->max_pages; ++p)
      if (which == 4 ? h[(size_t)p] == PG_FULL : h[(size_t)p] != 0) v.push_back(p);
  } else if (which == 5) {
    int64_t head = sc[SC_RB_HEAD], size = sc[SC_NRB];
    std::vector<int64_t> h((size_t)size);
    for (int64_t i = 0; i < size; ++i)
      CUDA_TRY(cudaMemcpy(&h[(size_t)i], m->st.rb + ((head + i) & m->st.rb_mask), 8, cudaMemcpyDeviceToHost));
    v = h;
  } else if (which == 7) {   // _evictable_blocks in LRU order
    int64_t head = sc[SC_LRU_HEAD], tail = sc[SC_LRU_TAIL];
    std::vector<int32_t> h((size_t)std::max<int64_t>(tail - head, 0));
    if (!h.empty()) CUDA_TRY(cudaMemcpy(h.data(), m->st.lru + head, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int32_t e : h) if (e >= 0) v.push_back(e);
  } else if (which == 8) {   // blocks that hold a cache key
    int64_t nwords = (m->cfg.num_blocks + 31) / 32;
    std::vector<uint32_t> h((size_t)nwords);
    CUDA_TRY(cudaMemcpy(h.data(), m->st.pin_keyed, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < m->cfg.num_blocks; ++b) if ((h[(size_t)(b >> 5)] >> (b & 31)) & 1u) v.push_back(b);
  } else if (which == 9) {   // (block, ref_cnt) of every referenced block
    std::vector<int32_t> h((size_t)m->cfg.num_blocks);
    CUDA_TRY(cudaMemcpy(h.data(), m->st.pin_ref, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < m->cfg.num_blocks; ++b) if (h[(size_t)b]) { v.push_back(b); v.push_back(h[(size_t)b]); }
  } else {
    set_error("dump_list: unknown list %d", which);
    return KVB_ERR_INVALID;
  }
  if ((int64_t)v.size() > cap) return KVB_ERR_OVERFLOW;
  if (!v.empty()) memcpy(out, v.data(), v.size() * 8);
  return (int64_t)v.size();
}

extern "C" int64_t kvb_mgr_page_free_blocks(kvb_mgr *m, int64_t p, int64_t *out, int64_t cap) {
  if (!m || p < 0 || p >= m->max_pages) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  int32_t st = 0, cnt = 0, head = 0;
  CUDA_TRY(cudaMemcpy(&st, m->st.pg_state + p, 4, cudaMemcpyDeviceToHost));
  if (st == PG_NONE) return -1;
  CUDA_TRY(cudaMemcpy(&cnt, m->st.pg_cnt + p, 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(&head, m->st.pg_head + p, 4, cudaMemcpyDeviceToHost));
  if (cnt > cap) return KVB_ERR_OVERFLOW;
  std::vector<uint32_t> ring((size_t)m->bpp);
  CUDA_TRY(cudaMemcpy(ring.data(), m->st.ring + (size_t)p * m->bpp, ring.size() * 4, cudaMemcpyDeviceToHost));
  int64_t start = (p * m->P + m->B - 1) / m->B;
  for (int i = 0; i < cnt; ++i) out[i] = start + ring[(size_t)((head + i) % m->bpp)];
  return cnt;
}

extern "C" int64_t kvb_mgr_drain_events(kvb_mgr *m, int64_t *out, int64_t cap) {
  if (!m) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  int64_t n = (int64_t)m->pending_events.size();
  if (n > cap) return KVB_ERR_OVERFLOW;
  if (n) memcpy(out, m->pending_events.data(), (size_t)n * 8);
  m->pending_events.clear();
  return n;
}

extern "C" int kvb_mgr_audit(kvb_mgr *m, int64_t *out4, float *device_ms) {
  if (!m || !out4) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  CUDA_TRY(cudaEventRecord(m->ev0, m->stream));
  g_total_launches.fetch_add(1, std::memory_order_relaxed);
  if (launch_audit(m->st, m->d_scratch + 8, m->stream) != 0) { set_error("audit kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())); return KVB_ERR_CUDA; }
  CUDA_TRY(cudaEventRecord(m->ev1, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (device_ms) CUDA_TRY(cudaEventElapsedTime(device_ms, m->ev0, m->ev1));
  CUDA_TRY(cudaMemcpy(out4, m->d_scratch + 8, 32, cudaMemcpyDeviceToHost));
  return 0;
}

// The audit sweep on a bitmap large enough to stream from HBM (the pool's own bitmap is 11 KB and L2-resident):
// num_blocks bits + per-page state/count words, pattern = every third page fully used, every seventh half used.
// out = {avg kernel ms, algorithmic bytes per launch (bitmap + 8 B per page), used blocks counted, used blocks expected}
extern "C" int kvb_audit_stress(int device_index, int64_t num_blocks, int32_t blocks_per_page, int32_t iters, double *out4) {
  if (num_blocks <= 0 || blocks_per_page <= 0 || blocks_per_page % 32 || num_blocks % blocks_per_page || iters <= 0 || !out4) {
    set_error("audit_stress: num_blocks must be a positive multiple of blocks_per_page, itself a multiple of 32");
    return KVB_ERR_INVALID;
  }
  if (ensure_device(device_index < 0 ? 0 : device_index)) { set_error("cudaSetDevice failed"); return KVB_ERR_CUDA; }
  const int64_t pages = num_blocks / blocks_per_page, nwords = num_blocks / 32, wpp = blocks_per_page / 32;
  if (pages > INT32_MAX) { set_error("audit_stress: too many pages"); return KVB_ERR_INVALID; }
  DevState st{};
  st.P = 2 << 20;
  st.B = st.P / blocks_per_page;
  st.bpp = blocks_per_page;
  st.exact = 1;
  st.num_blocks = num_blocks;
  st.max_pages = (int32_t)pages;
  int64_t *d_out = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&] {
    if (st.bitmap) cudaFree(st.bitmap);
    if (st.pg_state) cudaFree(st.pg_state);
    if (st.pg_cnt) cudaFree(st.pg_cnt);
    if (d_out) cudaFree(d_out);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define ST_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e_)); cleanup(); return KVB_ERR_CUDA; } } while (0)
  ST_TRY(cudaMalloc((void **)&st.bitmap, (size_t)(nwords + 8) * 4));
  ST_TRY(cudaMalloc((void **)&st.pg_state, (size_t)pages * 4));
  ST_TRY(cudaMalloc((void **)&st.pg_cnt, (size_t)pages * 4));
  ST_TRY(cudaMalloc((void **)&d_out, 64));
  ST_TRY(cudaEventCreate(&e0));
  ST_TRY(cudaEventCreate(&e1));
  {
    // host-built pattern, uploaded in slabs
    const int64_t slab_pages = 1 << 16;
    std::vector<uint32_t> bm((size_t)(slab_pages * wpp));
    std::vector<int32_t> ps((size_t)slab_pages), pc((size_t)slab_pages);
    for (int64_t p0 = 0; p0 < pages; p0 += slab_pages) {
      int64_t n = std::min(slab_pages, pages - p0);
      for (int64_t i = 0; i < n; ++i) {
        int64_t p = p0 + i;
        int used = (p % 3 == 0) ? blocks_per_page : (p % 7 == 0 ? blocks_per_page / 2 : 0);
        for (int64_t w = 0; w < wpp; ++w) {
          int64_t left = used - w * 32;
          bm[(size_t)(i * wpp + w)] = left >= 32 ? 0xffffffffu : (left > 0 ? ((1u << left) - 1u) : 0u);
        }
        ps[(size_t)i] = used == blocks_per_page ? PG_FULL : (used ? PG_AVAIL : PG_NONE);
        pc[(size_t)i] = used ? blocks_per_page - used : 0;
      }
      ST_TRY(cudaMemcpy(st.bitmap + p0 * wpp, bm.data(), (size_t)(n * wpp) * 4, cudaMemcpyHostToDevice));
      ST_TRY(cudaMemcpy(st.pg_state + p0, ps.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
      ST_TRY(cudaMemcpy(st.pg_cnt + p0, pc.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    }
  }
  int64_t expect = 0;
  for (int64_t p = 0; p < pages; ++p) expect += (p % 3 == 0) ? blocks_per_page : (p % 7 == 0 ? blocks_per_page / 2 : 0);
  for (int w = 0; w < 3; ++w)
    if (launch_audit(st, d_out, nullptr) != 0) { set_error("audit launch failed"); cleanup(); return KVB_ERR_CUDA; }
  ST_TRY(cudaDeviceSynchronize());
  float total_ms = 0;
  for (int it = 0; it < iters; ++it) {   // the kernel alone: its counter memset sits outside the event pair
    ST_TRY(cudaMemsetAsync(d_out, 0, 32, nullptr));
    ST_TRY(cudaEventRecord(e0, nullptr));
    if (launch_audit(st, d_out, nullptr, false) != 0) { set_error("audit launch failed"); cleanup(); return KVB_ERR_CUDA; }
    ST_TRY(cudaEventRecord(e1, nullptr));
    ST_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    ST_TRY(cudaEventElapsedTime(&ms, e0, e1));
    total_ms += ms;
    g_total_launches.fetch_add(1, std::memory_order_relaxed);
  }
  int64_t h[4] = {0, 0, 0, 0};
  ST_TRY(cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost));
#undef ST_TRY
  out4[0] = total_ms / iters;
  out4[1] = (double)(nwords * 4 + pages * 8);
  out4[2] = (double)h[0];
  out4[3] = (double)expect;
  int bad = (h[0] != expect || h[3] != 0) ? 1 : 0;
  cleanup();
  if (bad) { set_error("audit_stress: counted %ld used blocks (expected %ld), %ld mismatching pages", (long)h[0], (long)expect, (long)h[3]); return KVB_ERR_STATE; }
  return 0;
}

extern "C" int kvb_mgr_bitmap(kvb_mgr *m, uint64_t *dev_ptr, int64_t *n_words32) {
  if (!m) return KVB_ERR_INVALID;
  if (dev_ptr) *dev_ptr = (uint64_t)(uintptr_t)m->st.bitmap;
  if (n_words32) *n_words32 = (m->cfg.num_blocks + 31) / 32;
  return 0;
}

// group_indices_by_page through the same device code path free() uses (state untouched)
extern "C" int64_t kvb_group_indices_by_page(kvb_mgr *m, const int64_t *ids, int64_t n, int64_t *pages_out,
                                             int64_t *counts_out, int64_t *grouped_out) {
  if (!m || n < 0 || (n > 0 && (!ids || !pages_out || !counts_out || !grouped_out))) return KVB_ERR_INVALID;
  if (n == 0) return 0;
  if (n > m->f_cap) { set_error("group_indices_by_page: %ld ids exceed the pool's block count", (long)n); return KVB_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  kvb_op op{KVB_OP_FREE_IDS, (int32_t)n, 0};
  size_t b_ops = sizeof(kvb_op), b_ids = (size_t)n * 8;
  size_t o_res = align_up(sizeof(Summary), 16), o_pg = o_res + 64, o_ct = o_pg + b_ids, o_gr = o_ct + b_ids;
  int rc = ensure_io(m, b_ops + b_ids, o_gr + b_ids);
  if (rc) return rc;
  memcpy(m->h_in, &op, b_ops);
  memcpy(m->h_in + b_ops, ids, b_ids);
  CUDA_TRY(cudaMemcpyAsync(m->d_in, m->h_in, b_ops + b_ids, cudaMemcpyHostToDevice, m->stream));
  rc = launch(m, (const kvb_op *)m->d_in, 0, 1, (const int64_t *)(m->d_in + b_ops), nullptr, (kvb_res *)(m->d_out + o_res),
              nullptr, 0, (Summary *)m->d_out, 2, 1, (int64_t *)(m->d_out + o_pg), (int64_t *)(m->d_out + o_ct),
              (int64_t *)(m->d_out + o_gr));
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(m->h_out, m->d_out, o_gr + b_ids, cudaMemcpyDeviceToHost, m->stream));
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  const kvb_res *r = (const kvb_res *)(m->h_out + o_res);
  int64_t D = r->n;
  memcpy(pages_out, m->h_out + o_pg, (size_t)D * 8);
  memcpy(counts_out, m->h_out + o_ct, (size_t)D * 8);
  memcpy(grouped_out, m->h_out + o_gr, b_ids);
  return D;
}

// ============================================================================================ tensor parallel
// (1) The mailbox (kvb_tp.h): rank 0's kernel pushes each step's page list into a shared ring every rank polls.
extern "C" int kvb_tp_init(const char *name, int rank, int world_size, int64_t group_id, kvb_mgr *m, int flags) {
  if (rank == 0 && !m && (flags & 2)) {
    // external producer: page lists come through kvb_tp_publish (a host that runs its bookkeeping elsewhere)
    const double t_ms = env_f64("KVB200_TP_TIMEOUT_MS", 60000.0);
    int rc0 = tp_open(name, 0, world_size, group_id, false, t_ms);
    if (!rc0) rc0 = tp_mailbox(group_id)->wait_attached(t_ms);
    return rc0;
  }
  if (rank == 0 && !m) { set_error("kvb_tp_init: rank 0 needs its block manager"); return KVB_ERR_INVALID; }
  if (rank != 0 && m) { set_error("kvb_tp_init: only rank 0 drives a block manager"); return KVB_ERR_INVALID; }
  const double timeout_ms = env_f64("KVB200_TP_TIMEOUT_MS", 60000.0);
  const bool direct = !(flags & 1) && env_i64("KVB200_TP_DEVICE_DIRECT", 1) != 0;
  if (rank != 0) return tp_open(name, rank, world_size, group_id, false, timeout_ms);
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  int rc = collect_pending(m);
  if (rc) return rc;
  rc = tp_open(name, 0, world_size, group_id, direct, timeout_ms);
  if (rc) return rc;
  TpMailbox *tp = tp_mailbox(group_id);
  m->tp_timeout_ms = timeout_ms;
  // every page this manager mapped before the mailbox existed (the null block's page) opens the ring
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  std::vector<int32_t> mapped((size_t)m->max_pages);
  CUDA_TRY(cudaMemcpy(mapped.data(), m->st.pg_mapped, mapped.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<int64_t> ev;
  for (int64_t p = 0; p < m->max_pages; ++p) if (mapped[(size_t)p]) ev.push_back(p * 2);
  rc = tp->publish(m->seq, ev.data(), (int64_t)ev.size(), m->last.next_free, m->last.next_free_n, 0);
  if (!rc) rc = tp->wait_attached(timeout_ms);
  if (!rc) rc = tp->wait_applied(timeout_ms);
  if (rc) { tp_close(group_id); return rc; }
  m->tp_on = true;
  return 0;
}

// rank 0: every rank that attached has applied everything published; other ranks: this rank has.
extern "C" int kvb_tp_fence(int64_t group_id) {
  TpMailbox *tp = tp_mailbox(group_id);
  if (!tp) { set_error("kvb_tp_fence: group %ld has no mailbox", (long)group_id); return KVB_ERR_STATE; }
  const double timeout_ms = env_f64("KVB200_TP_TIMEOUT_MS", 60000.0);
  return tp->rank() == 0 ? tp->wait_applied(timeout_ms) : tp->wait_caught_up(timeout_ms);
}

extern "C" int kvb_tp_publish(int64_t group_id, int64_t seq, const int64_t *events, int64_t n, const int64_t *hints,
                              int64_t n_hints, int flags) {
  TpMailbox *tp = tp_mailbox(group_id);
  if (!tp) { set_error("kvb_tp_publish: group %ld has no mailbox", (long)group_id); return KVB_ERR_STATE; }
  if (n < 0 || (n > 0 && !events) || n_hints < 0 || (n_hints > 0 && !hints)) return KVB_ERR_INVALID;
  int rc = tp->publish(seq, events, n, hints, n_hints, flags);
  if (rc) return rc;
  bool has_map = (flags & kTpFlagSettle) != 0;
  for (int64_t i = 0; i < n && !has_map; ++i) has_map = !(events[i] & 1);
  return has_map ? tp->wait_applied(env_f64("KVB200_TP_TIMEOUT_MS", 60000.0)) : 0;
}

extern "C" int kvb_tp_stats(int64_t group_id, int64_t *out8) {
  TpMailbox *tp = tp_mailbox(group_id);
  if (!tp || !out8) return KVB_ERR_STATE;
  tp->stats(out8);
  return 0;
}

extern "C" int kvb_tp_shutdown(int64_t group_id) { return tp_close(group_id); }

extern "C" int kvb_mgr_tp_detach(kvb_mgr *m) {
  if (!m) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  m->tp_on = false;
  return 0;
}

// (2) The collective path, kept as the checker of (1) and for multi-node groups: the page list as a fixed-size
// int64 buffer [seq, n, e0 .. e(n-1), n_hints, hints ..., (last word) more] that the host side broadcasts with ONE
// NCCL broadcast per chunk (torch.distributed).  Lists of any length travel in chunks: `more` = 1 says another
// buffer of the same step follows (the reference ships page lists of any length, tp_ipc_util.py:173-267).
static constexpr int64_t kTpWords = 1024;
static constexpr int64_t kTpMaxEvents = kTpWords - 4 - 16;   // seq, n, n_hints, more + 16 hints
extern "C" int64_t kvb_tp_buffer_words(void) { return kTpWords; }

// packs the next chunk of `events[*cursor .. n)`; returns 1 if more chunks follow, 0 if this was the last
extern "C" int kvb_tp_pack(int64_t seq, const int64_t *events, int64_t n, int64_t *cursor, const int64_t *hints,
                           int64_t n_hints, int64_t *buf) {
  if (!buf || !cursor || n < 0 || *cursor < 0 || *cursor > n || (n > 0 && !events)) return KVB_ERR_INVALID;
  const int64_t k = std::min<int64_t>(n - *cursor, kTpMaxEvents);
  const bool last = *cursor + k >= n;
  int64_t nh = last ? std::max<int64_t>(0, std::min<int64_t>(n_hints, 16)) : 0;
  memset(buf, 0, (size_t)kTpWords * 8);
  buf[0] = seq;
  buf[1] = k;
  if (k) memcpy(buf + 2, events + *cursor, (size_t)k * 8);
  buf[2 + k] = nh;
  for (int64_t i = 0; i < nh; ++i) buf[3 + k + i] = hints[i];
  buf[kTpWords - 1] = last ? 0 : 1;
  *cursor += k;
  return last ? 0 : 1;
}

// host buffer -> this rank's VMM engine; returns 1 if another chunk of the same step follows, 0 if complete
extern "C" int kvb_tp_apply_host(int64_t group_id, const int64_t *buf, int64_t *seq_out, int64_t *n_out) {
  if (!buf) return KVB_ERR_INVALID;
  VmmGroup *g = vmm_group(group_id, false);
  if (!g || !g->created()) { set_error("tp_apply: KV tensors of group %ld are not created", (long)group_id); return KVB_ERR_STATE; }
  const int64_t n = buf[1], more = buf[kTpWords - 1];
  if (seq_out) *seq_out = buf[0];
  if (n_out) *n_out = n;
  if (n < 0 || n > kTpMaxEvents || (more != 0 && more != 1)) { set_error("tp_apply: corrupt header n=%ld more=%ld", (long)n, (long)more); return KVB_ERR_INVALID; }
  const int64_t *ev = buf + 2;
  std::lock_guard<std::mutex> lk(g->mu);
  int rc = 0;
  if (n > 0) {
    rc = g->enqueue_events(ev, n);
    // maps wait only when the step's list is complete: chunks of one list are one batch for the engine
    if (rc >= 0 && !more) rc = g->async() ? g->wait_maps() : g->fence();
  } else if (!more) {
    rc = g->async() ? g->wait_maps() : 0;
  }
  if (rc < 0) return rc;
  const int64_t nh = ev[(size_t)n];
  if (!more && nh >= 0 && nh <= 16) g->set_prefetch_hints(ev + n + 1, nh);
  return (int)more;
}

// rank 0: next chunk of the manager's pending page events -> device buffer `dst`.  Returns 1 while more chunks
// remain (call again after the broadcast), 0 when the list is drained.
extern "C" int kvb_mgr_tp_export(kvb_mgr *m, uint64_t dst, void *stream) {
  if (!m || !dst) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) { set_error("cudaSetDevice(%d) failed", m->dev); return KVB_ERR_CUDA; }
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  int rc = ensure_io(m, (size_t)kTpWords * 8, 0);
  if (rc) return rc;
  int64_t cursor = 0;
  const int64_t n = (int64_t)m->pending_events.size();
  int more = kvb_tp_pack(m->seq, m->pending_events.data(), n, &cursor, m->last.next_free, m->last.next_free_n, (int64_t *)m->h_in);
  if (more < 0) return more;
  m->pending_events.erase(m->pending_events.begin(), m->pending_events.begin() + cursor);
  CUDA_TRY(cudaMemcpyAsync((void *)(uintptr_t)dst, m->h_in, (size_t)kTpWords * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return more;
}
// same, into a HOST buffer (gloo / CPU groups)
extern "C" int kvb_mgr_tp_export_host(kvb_mgr *m, int64_t *dst) {
  if (!m || !dst) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  { int rc0 = collect_pending(m); if (rc0) return rc0; }
  int64_t cursor = 0;
  const int64_t n = (int64_t)m->pending_events.size();
  int more = kvb_tp_pack(m->seq, m->pending_events.data(), n, &cursor, m->last.next_free, m->last.next_free_n, dst);
  if (more < 0) return more;
  m->pending_events.erase(m->pending_events.begin(), m->pending_events.begin() + cursor);
  return more;
}

// every other rank: device buffer -> pinned host (ONE copy + one stream synchronisation) -> kvb_tp_apply_host
extern "C" int kvb_tp_apply(int64_t group_id, uint64_t src, void *stream, int64_t *seq_out, int64_t *n_out) {
  if (!src) return KVB_ERR_INVALID;
  static thread_local int64_t *pinned = nullptr;
  if (!pinned && cudaHostAlloc((void **)&pinned, (size_t)kTpWords * 8, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    pinned = nullptr;
  }
  std::vector<int64_t> pageable;
  int64_t *buf = pinned;
  if (!buf) { pageable.resize((size_t)kTpWords); buf = pageable.data(); }
  CUDA_TRY(cudaMemcpyAsync(buf, (const void *)(uintptr_t)src, (size_t)kTpWords * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return kvb_tp_apply_host(group_id, buf, seq_out, n_out);
}

// ---- shm usage segment (kvctl / kvtop wire format) ---------------------------------------------------------
extern "C" int kvb_mgr_shm_name(kvb_mgr *m, char *out, int cap) {
  if (!m || !out || cap <= 0) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (!m->shm) { out[0] = 0; return 0; }
  snprintf(out, (size_t)cap, "%s", m->shm->name().c_str());
  return (int)m->shm->name().size();
}
extern "C" int64_t kvb_mgr_shm_publishes(kvb_mgr *m) { return m ? m->shm_publishes : KVB_ERR_INVALID; }
extern "C" int kvb_shm_read(const char *ipc_name, int64_t *out3) {
  if (!ipc_name || !out3) return KVB_ERR_INVALID;
  return shm_read(ipc_name, out3) == 0 ? 0 : KVB_ERR_STATE;
}
extern "C" int kvb_shm_set_limit(const char *ipc_name, int64_t total_bytes, int create) {
  if (!ipc_name || total_bytes < 0) return KVB_ERR_INVALID;
  return shm_set_limit(ipc_name, total_bytes, create != 0) == 0 ? 0 : KVB_ERR_STATE;
}

// -DKVB_PROFILE builds: cycles per phase [0..15] and mark counts [16..31] of the step kernel; zeroes otherwise
extern "C" int kvb_mgr_profile(kvb_mgr *m, int64_t *out32, int reset) {
  if (!m || !out32) return KVB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ensure_device(m->dev)) return KVB_ERR_CUDA;
  CUDA_TRY(cudaStreamSynchronize(m->stream));
  CUDA_TRY(cudaMemcpy(out32, m->d_scratch + 32, 32 * 8, cudaMemcpyDeviceToHost));
  if (reset) CUDA_TRY(cudaMemset(m->d_scratch + 32, 0, 32 * 8));
  return 0;
}

// kernels launched by this library since load (bookkeeping steps, audits, KV scatter/gather)
extern "C" int64_t kvb_total_launches(void) { return g_total_launches.load(); }
