// kvb_vmm.h — batched CUDA VMM engine: reserved VA ranges of every layer's K/V tensor, one shared
// zero page, pooled physical handles, and the cuMemUnmap / cuMemMap / cuMemSetAccess loop.
// Replaces csrc/allocator.cpp (FTensorAllocator), csrc/ftensor.cpp (FTensor), csrc/page.cpp (GPUPage)
// and csrc/inc/gpu_vmm.hpp of the reference.  Host-only C++.
#pragma once

#include <cuda.h>

#include <atomic>

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace kvb {

void set_error(const char *fmt, ...);
const char *get_error();

// CUDA driver entry points, resolved at first use through cudaGetDriverEntryPoint so that the
// library itself has no link-time dependency on libcuda.so.1 (it must load on driver-less boxes).
struct Driver {
  CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
  bool ok = false;
};
const Driver *driver();   // nullptr (and error set) when no driver is available

struct VmmRange {
  CUdeviceptr base = 0;
  size_t bytes = 0;
};

class VmmGroup {
 public:
  VmmGroup(int dev, size_t page_size, bool contiguous) : dev_(dev), P_(page_size), contiguous_(contiguous) {}
  ~VmmGroup();

  // FTensorAllocator::create_kv_tensors, allocator.cpp:121-154
  int create(int64_t size_per_layer, int64_t num_layers, int64_t num_kv, bool unified);
  bool created() const { return created_; }
  const std::vector<VmmRange> &ranges() const { return ranges_; }

  // page-id addressed batch ops (each runs the whole driver loop over all layers' ranges); synchronous:
  // they return when the physical state equals the requested one
  int map_pages(const int64_t *ids, int64_t n);
  int unmap_pages(const int64_t *ids, int64_t n);
  int apply_events(const int64_t *events, int64_t n);   // page*2 + is_unmap, in order; = enqueue + fence

  // ---- asynchronous engine -----------------------------------------------------------------------
  // Events only change the WANTED state of a page; one worker thread (driver VMM calls do not overlap
  // across threads on this stack, tools/vmm_stall2.cu) reconciles physical with wanted state, maps
  // before unmaps.  A map that meets its own page's pending unmap (or the reverse) cancels it: no driver
  // call at all.  wait_maps() is what a scheduler step needs before its block ids are usable; unmaps
  // drain in the background (the reference's prealloc thread hides map latency the same way,
  // page_allocator.cpp:524-617, but changes the bookkeeping; this does not).
  int enqueue_events(const int64_t *events, int64_t n);
  int wait_maps();    // every wanted mapping is physically in place
  int fence();        // every wanted mapping in place, every unwanted (and not hinted) page back on the zero page
  // Map-ahead: the pages the bookkeeping will hand out next (the front of its free page list, published by the
  // step kernel) are mapped by the VMM thread while it has nothing else to do, so the map event that follows
  // finds them in place.  Physical state only — the bookkeeping's mapped set, counters and page order do not
  // change (the reference's prealloc thread, page_allocator.cpp:524-617, moves pages to its reserved list and so
  // changes results with timing).  Priority: wanted maps > hinted maps > unmaps.  A hinted page that falls out of
  // the hint window without being wanted is unmapped again.  Off in synchronous mode and when depth is 0.
  // Retention: a page the bookkeeping unmapped keeps its physical handle mapped (its data is stale, nobody else's)
  // until the allocator has been quiet (no map event) for `ms` milliseconds; if the bookkeeping maps it again
  // before that, nothing happens at all.  It is put
  // back on the zero page and its handle released earlier when a map needs a handle that can neither come from the
  // pool nor be created within the memory headroom (the most recently retained page goes first: the free page list
  // is a FIFO, so that page is the one needed last), when the pool cap is exceeded, at settle points and under
  // memory pressure.  Idle pooled handles are released after the same time.  0 = unmap at once (and in
  // synchronous mode).  Count-based hysteresis is what the reference has (max_reserved_pages, page_allocator.cpp:264-310).
  void set_retention_ms(double ms);
  // Retained pages are a cache in memory nobody asked for yet.  While somebody else's footprint on the device is
  // growing (another process, or this process's own torch allocations) — and for 10 s after it last grew — they are
  // kept only as long as this fraction of the device stays free (default KVB200_RETAIN_FREE_FRACTION or 0.20): the
  // VMM thread hands retained pages back until the reserve is whole again.  While nobody else moves, the reserve is
  // the utilisation headroom plus KVB200_RETAIN_QUIET_MARGIN (0.02 of the device): a co-located model that wakes up
  // always sees that margin as available, and taking any of it switches the larger reserve on within one poll.
  void set_retain_free_fraction(double f, double quiet_margin = -1);
  // Retained pages never hold more than this fraction of the device (default KVB200_RETAIN_MAX_FRACTION or 0.10): a
  // page unmapped by the bookkeeping while the cap is reached goes back to the zero page at once.
  void set_retain_max_fraction(double f);
  // out8 = {pages physically mapped (wanted + mapped ahead + retained + unmap pending), pages the bookkeeping wants
  // mapped, peak of the difference since the last call, pages retained now, retention cap in pages, peak retained,
  // pages' worth of unmapped handles in the pool, pool cap in pages}
  void footprint(int64_t *out8);
  int settle();       // fence + every retained page back on the zero page (trim / clear / resize / kvctl limit)
  void set_prefetch_hints(const int64_t *pages, int64_t n);
  void set_prefetch_depth(int depth);
  int prefetch_depth() const { return prefetch_depth_; }
  int64_t prefetched_pages(int64_t *out, int64_t cap);   // hinted, physically mapped, not (yet) wanted
  void set_async(bool on);
  bool async() const { return async_; }
  void set_sync_before_unmap(bool on) { sync_before_unmap_ = on; }
  int64_t pending_unmaps();
  // Zero-page init (ftensor.cpp:160-176 does it eagerly: 0.6-2 s at start-up, bench_layout/README.md:75).  Background
  // mode: create() backs unit 0 only, the VMM thread the rest in its idle time; wait_init() = every unit is backed.
  int wait_init();
  void init_progress(int64_t *out3);   // {units backed or skipped so far, units, microseconds spent in init}
  // Called by the VMM thread on its timer wake-ups (<= 50 ms apart) with no engine lock held: the block manager hangs
  // the collection of launches that returned before their kernel finished (asynchronous frees) on it.
  // set_idle_hook returns once a hook that is running has finished.
  void set_idle_hook(std::function<void()> fn);
  // Free / total device memory as last seen by the VMM thread (it polls every few milliseconds while it works and
  // at every idle wake-up, and accounts for its own cuMemCreate / cuMemRelease in between), so that the scheduler
  // thread never enters the driver for available_size().  False if there is no recent sample.
  bool cached_meminfo(size_t *free_bytes, size_t *total_bytes);
  // reference byte offsets (page_allocator.cpp:619-631) -> page ids
  int64_t offset_to_page(int64_t off) const { return off / (int64_t)(contiguous_ ? unit_ : P_); }

  void pool_config(int64_t max_cached);
  int pool_prefill(int64_t n_pages);
  void stats(int64_t *s8);
  void engine_stats(int64_t *s8);   // {events elided, map-ahead hits, pages mapped ahead now, need_map, need_unmap, need_spec, hints, depth}
  void timing(double *out10);   // us spent in / calls of {create, release, unmap, map, setaccess}
  int64_t mapped_pages(int64_t *out, int64_t cap);
  int64_t pooled_pages();
  int64_t pool_release(int64_t n_pages);   // give pooled physical memory back to the device
  int64_t num_pages() const { return n_pages_; }
  size_t unit() const { return unit_; }
  int64_t num_layers() const { return L_; }
  int64_t num_kv() const { return nkv_; }
  int raw_control(int64_t n_pages, int64_t iters, double *map_us, double *unmap_us);

  std::mutex mu;   // serialises API callers (create, pool policy, inspection); the engine state has its own lock

 private:
  using Lock = std::unique_lock<std::mutex>;
  bool process_next(Lock &lk);              // one page transition; lock held on entry and exit, released around driver calls
  int transition(Lock &lk, int64_t p, bool to_mapped);
  void update(int64_t p, int want, int spec, int phys, int force = -1);   // -1 = keep; maintains the counters
  void set_want(int64_t p, bool v) { update(p, v, -1, -1, v ? 0 : -1); }
  void set_phys(int64_t p, bool v) { update(p, -1, -1, v, v ? -1 : 0); }
  void make_unwanted(int64_t p);            // a mapped page nobody wants or hints: retain it, or queue its unmap
  bool force_one_retained(bool newest);     // pick a retained page and queue its unmap
  void expire(double now_ms);               // retention / idle-pool time-outs (VMM thread)
  bool can_create(bool soft);               // room for one more page above the retention reserve (soft) / the utilisation headroom
  void poll_meminfo(Lock &lk);              // cudaMemGetInfo on the VMM thread (lock released around the call)
  void ensure_worker(Lock &lk);
  int drain(Lock &lk, bool maps_only);
  void worker_main();
  void init_some(Lock &lk);
  void stop_worker();
  CUdeviceptr sub_addr(int64_t p, int sub) const;

  int dev_;
  size_t P_;
  bool contiguous_;
  bool unified_ = false;
  bool created_ = false;
  int64_t L_ = 0, nkv_ = 0;
  size_t size_per_layer_ = 0;
  size_t unit_ = 0;       // bytes of one physical handle: compound page (contiguous) or page
  int n_sub_ = 1;         // handles per page id
  int64_t n_pages_ = 0;
  std::vector<VmmRange> ranges_;
  CUdeviceptr va_base_ = 0;
  size_t va_bytes_ = 0;
  CUmemGenericAllocationHandle zero_ = 0;
  bool have_zero_ = false;
  std::vector<CUmemGenericAllocationHandle> handles_;   // [n_pages * n_sub], 0 = zero page mapped
  std::vector<uint8_t> zero_ok_;  // per unit: the VA unit has a mapping (zero page or handle)
  int64_t n_units_ = 0, init_next_ = 0;
  bool bg_init_ = false, zero_page_ = true;
  double init_ms_ = 0;
  std::vector<uint8_t> mapped_;   // physical state
  std::vector<uint8_t> want_;     // state after every event received so far
  std::vector<uint8_t> spec_;     // hinted: map ahead of the event
  std::vector<uint8_t> force_;    // unwanted + mapped + must go back to the zero page now (else: retained)
  std::vector<uint32_t> ret_seq_; // validates entries of retained_
  struct Retained { int64_t p; uint32_t seq; double t_ms; };
  std::deque<Retained> retained_; // oldest first
  struct PoolEnt { CUmemGenericAllocationHandle h; double t_ms; };
  double retention_ms_ = 0, pool_idle_ms_ = 2000.0;
  int64_t n_retained_ = 0, n_retention_hits_ = 0, release_debt_ = 0, pool_floor_ = 16;
  int64_t n_want_ = 0, peak_over_ = 0, peak_retained_ = 0;
  double retain_max_frac_ = 0.10;
  int64_t retain_cap_pages() const;
  double mem_t_ms_ = -1e18, gpu_util_ = 0.95, retain_free_frac_ = 0.20, last_map_ms_ = -1e18;
  double mem_free_ = 0, mem_total_ = 0;     // bytes, as of mem_t_ms_ (adjusted by our own creates / releases)
  double quiet_margin_ = 0.02, others_base_ = -1, contended_until_ms_ = -1e18;
  double reserve_fraction(double now) const;
  std::deque<int64_t> qm_, qu_, qs_;   // pages whose wanted state may differ from the physical one (maps / unmaps / hinted maps)
  std::vector<int64_t> hints_;
  int prefetch_depth_ = 0;
  int64_t need_map_ = 0, need_unmap_ = 0, need_spec_ = 0, n_spec_mapped_ = 0, n_prefetch_hits_ = 0, inflight_ = -1;
  bool inflight_to_mapped_ = false;
  std::mutex qmu_, hook_mu_;
  std::function<void()> idle_hook_;
  std::condition_variable cv_work_, cv_done_;
  std::thread worker_;
  bool worker_running_ = false, stop_ = false, async_ = false, sync_before_unmap_ = false;
  int werr_ = 0;
  std::string werr_msg_;
  int64_t n_elided_ = 0;
  std::deque<PoolEnt> pool_;      // idle physical handles, most recently returned last
  int64_t pool_cap_pages_ = -1;   // pages' worth of handles kept after unmap; -1 = default chosen in create()
  CUmemAllocationProp prop_{};
  CUmemAccessDesc access_{};
  double t_us_[5] = {0, 0, 0, 0, 0};
  int64_t t_n_[5] = {0, 0, 0, 0, 0};
  int64_t n_mapped_ = 0, n_created_ = 0, n_released_ = 0, n_map_calls_ = 0, n_unmap_calls_ = 0;
};

extern std::atomic<int> g_zero_init_mode;
// process-global registry (FTensorAllocator::init / global_allocator / shutdown, allocator.cpp:72-119)
int vmm_init(int dev, int64_t page_size, bool contiguous);
int vmm_shutdown();
bool vmm_initialized();
int64_t vmm_page_size();
bool vmm_contiguous();
int vmm_device();
VmmGroup *vmm_group(int64_t group_id, bool create_if_missing);
// cudaSetDevice only when the calling thread is on another device (cudaSetDevice itself costs tens of
// microseconds on this stack; cudaGetDevice does not)
int ensure_device(int dev);
#ifdef KVB_TEST_HOOKS   // libkvb200_test.so only (kvb_testhooks_*.inc)
// mock driver (host-logic tests without a GPU): legality-checking stand-ins for the VMM driver calls
int vmm_mock_driver(int64_t total_bytes, int latency_us);
bool vmm_mock_installed();
void vmm_mock_set_other_bytes(int64_t bytes);
void vmm_mock_stats(int64_t *out10);
const char *vmm_mock_last_violation();
int vmm_probe(int dev, size_t unit, int64_t n, int64_t iters, double *out8);
int vmm_probe3(int dev, size_t unit, int64_t M, int64_t iters, double *out7);
int vmm_probe2(int dev, size_t unit, int64_t M, int mode, int64_t iters, double *out3);

#endif

}  // namespace kvb
