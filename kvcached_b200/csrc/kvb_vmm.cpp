// kvb_vmm.cpp — see kvb_vmm.h.
#include "kvb_vmm.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>

namespace kvb {

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  const char *lvl = getenv("KVCACHED_LOG_LEVEL");
  if (lvl && (lvl[0] == 'D' || lvl[0] == 'd')) fprintf(stderr, "[kvb200] %s\n", g_err);
}
const char *get_error() { return g_err; }

// ------------------------------------------------------------------------------------------ driver
static Driver g_drv;
static std::once_flag g_drv_once;

template <typename F>
static bool resolve(const char *name, F *out) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return false;
  *out = reinterpret_cast<F>(fn);
  return true;
}

const Driver *driver() {
  std::call_once(g_drv_once, [] {
    bool ok = true;
    ok &= resolve("cuMemAddressReserve", &g_drv.MemAddressReserve);
    ok &= resolve("cuMemAddressFree", &g_drv.MemAddressFree);
    ok &= resolve("cuMemCreate", &g_drv.MemCreate);
    ok &= resolve("cuMemRelease", &g_drv.MemRelease);
    ok &= resolve("cuMemMap", &g_drv.MemMap);
    ok &= resolve("cuMemUnmap", &g_drv.MemUnmap);
    ok &= resolve("cuMemSetAccess", &g_drv.MemSetAccess);
    ok &= resolve("cuMemGetAllocationGranularity", &g_drv.MemGetAllocationGranularity);
    ok &= resolve("cuGetErrorString", &g_drv.GetErrorString);
    g_drv.ok = ok;
  });
  if (!g_drv.ok) {
    cudaGetLastError();
    set_error("CUDA driver entry points unavailable (no driver / no GPU)");
    return nullptr;
  }
  return &g_drv;
}


#ifdef KVB_TEST_HOOKS
#include "kvb_testhooks_mock.inc"   // libkvb200_test.so only: legality-checking stand-ins for the VMM driver calls
#else
namespace mock { static constexpr bool installed = false; }
static bool mock_meminfo(size_t *, size_t *) { return false; }
#endif

static const char *cu_str(CUresult r) {
  const char *s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : "unknown CUresult";
}
#define CU_TRY(call)                                                         \
  do {                                                                       \
    CUresult r_ = (call);                                                    \
    if (r_ != CUDA_SUCCESS) {                                                \
      set_error("%s failed: %s (%d) at %s:%d", #call, cu_str(r_), (int)r_, __FILE__, __LINE__); \
      return KVB_ERR_CUDA_;                                                  \
    }                                                                        \
  } while (0)
static constexpr int KVB_ERR_CUDA_ = -2;
static constexpr int KVB_ERR_INVALID_ = -1;
static constexpr int KVB_ERR_STATE_ = -3;

// The reference reserves at 0x1f0_0000_0000 + running offset (ftensor.cpp:26-47, constants.hpp).  We
// hint at a different window so both libraries can live in one process.
static constexpr uint64_t kVaHint = 0x2f0000000000ull;
std::atomic<int> g_zero_init_mode{-1};   // -1 env KVB200_VMM_ZERO_INIT, 0 eager, 1 background
static std::atomic<size_t> g_va_off{0};

// ------------------------------------------------------------------------------------------ group
VmmGroup::~VmmGroup() {
  stop_worker();   // drains what is queued, then exits
  const Driver *d = g_drv.ok ? &g_drv : nullptr;
  if (!d) return;
  if (!mock::installed) cudaSetDevice(dev_);
  if (va_base_) {
    if (zero_page_ && (init_next_ >= n_units_ || n_units_ == 0)) {
      d->MemUnmap(va_base_, va_bytes_);      // tolerate stale mappings at teardown (ftensor.cpp:78-98)
    } else {                                  // holes (background init cut short / no zero page): only backed units
      for (int64_t u = 0; u < n_units_; ++u)
        if (zero_ok_[(size_t)u]) d->MemUnmap(sub_addr(u / n_sub_, (int)(u % n_sub_)), unit_);
    }
    d->MemAddressFree(va_base_, va_bytes_);
  }
  for (auto h : handles_)
    if (h) d->MemRelease(h);
  for (auto &e : pool_) d->MemRelease(e.h);
  if (have_zero_) d->MemRelease(zero_);
}

int VmmGroup::create(int64_t size_per_layer, int64_t num_layers, int64_t num_kv, bool unified) {
  const Driver *d = driver();
  if (!d) return KVB_ERR_CUDA_;
  if (created_) {
    if (num_layers != L_) { set_error("create_kv_tensors: num_layers changed (%ld -> %ld)", (long)L_, (long)num_layers); return KVB_ERR_INVALID_; }
    return (int)ranges_.size();
  }
  if (size_per_layer <= 0 || num_layers <= 0 || num_kv <= 0) { set_error("create_kv_tensors: bad sizes"); return KVB_ERR_INVALID_; }
  if (!mock::installed && (cudaSetDevice(dev_) != cudaSuccess || cudaFree(0) != cudaSuccess)) { set_error("cudaSetDevice(%d) failed", dev_); return KVB_ERR_CUDA_; }
  // allocator.cpp:130-136: round the per-layer size up to the page size
  size_t aligned = ((size_t)size_per_layer + P_ - 1) / P_ * P_;
  L_ = num_layers;
  nkv_ = num_kv;
  unified_ = unified;
  size_per_layer_ = aligned;

  memset(&prop_, 0, sizeof(prop_));
  prop_.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop_.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop_.location.id = dev_;
  access_.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access_.location.id = dev_;
  access_.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  size_t gran = 0;
  CU_TRY(d->MemGetAllocationGranularity(&gran, &prop_, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  if (gran == 0 || P_ % gran != 0) { set_error("page size %zu is not a multiple of the VMM granularity %zu", P_, gran); return KVB_ERR_INVALID_; }

  if (contiguous_) {
    unit_ = P_ * (size_t)num_layers * (size_t)num_kv;   // compound page, allocator.cpp:139-142
    n_sub_ = 1;
    size_t total = aligned * (size_t)num_layers;
    n_pages_ = (int64_t)(total / unit_);
    ranges_.resize(1);
    ranges_[0].bytes = total;
  } else {
    unit_ = P_;
    if (unified) {
      n_sub_ = (int)num_layers;
      n_pages_ = (int64_t)(aligned / P_);
    } else {
      if (aligned % (2 * P_) != 0) { set_error("per-layer size %zu must be a multiple of 2 * page size (allocator.cpp:46-52)", aligned); return KVB_ERR_INVALID_; }
      n_sub_ = (int)num_layers * 2;
      n_pages_ = (int64_t)(aligned / 2 / P_);
    }
    ranges_.resize((size_t)num_layers);
    for (auto &r : ranges_) r.bytes = aligned;
  }
  // reserve VA (ftensor.cpp:26-47).  ONE reservation for all layers: per-layer ranges sit back to back,
  // so the scatter/gather kernel addresses every layer with a constant stride.
  va_bytes_ = 0;
  for (auto &r : ranges_) va_bytes_ += r.bytes;
  {
    size_t off = g_va_off.fetch_add(va_bytes_);
    CU_TRY(d->MemAddressReserve(&va_base_, va_bytes_, 2u << 20, (CUdeviceptr)(kVaHint + off), 0));
    size_t o = 0;
    for (auto &r : ranges_) { r.base = va_base_ + o; o += r.bytes; }
  }
  // one zero page backs every unmapped page (ftensor.cpp:160-176).  Map per unit, grant access once
  // per range instead of once per page.
  CU_TRY(d->MemCreate(&zero_, unit_, &prop_, 0));
  have_zero_ = true;
  n_created_++;
  // Eager (default, like the reference): every unit is on the zero page when create returns — one cuMemMap + one
  // cuMemSetAccess per unit (a range call over many mappings costs MORE per unit on this driver, tools/vmm_stall2.cu).
  // Background (KVB200_VMM_ZERO_INIT=background / kvb_vmm_zero_init_mode(1)): only unit 0 is mapped here, the VMM
  // thread backs the rest in its idle time; pages the bookkeeping maps meanwhile go straight to their handle.  The
  // "stray accesses never fault" guarantee (ftensor.cpp:134) then holds from kvb_vmm_wait_init() on.
  n_units_ = n_pages_ * n_sub_;
  zero_ok_.assign((size_t)n_units_, 0);
  {
    const char *zi = getenv("KVB200_VMM_ZERO_INIT");
    int mode = g_zero_init_mode.load();
    if (mode < 0) mode = (zi && (zi[0] == 'b' || zi[0] == 'B')) ? 1 : ((zi && (zi[0] == 'n' || zi[0] == 'N')) ? 2 : 0);
    bg_init_ = mode == 1;
    // mode 2 ("none"): NO zero page — an unmapped page has no mapping at all, so a map is cuMemMap + cuMemSetAccess
    // and an unmap is cuMemUnmap alone (half the sleeping driver calls of the zero-page cycle); a stray access to an
    // unmapped page FAULTS.  For engines that never touch blocks they do not hold; the reference always keeps the
    // zero page (ftensor.cpp:134), so this is opt-in.
    zero_page_ = mode != 2;
  }
  const auto t_init0 = std::chrono::steady_clock::now();
  const int64_t eager_units = !zero_page_ ? 0 : (bg_init_ ? 1 : n_units_);
  for (int64_t u = 0; u < eager_units; ++u) {
    CUdeviceptr a = sub_addr(u / n_sub_, (int)(u % n_sub_));
    CU_TRY(d->MemMap(a, unit_, 0, zero_, 0));
    CU_TRY(d->MemSetAccess(a, unit_, &access_, 1));
    zero_ok_[(size_t)u] = 1;
  }
  init_next_ = zero_page_ ? eager_units : n_units_;
  init_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_init0).count();
  // the zero page must read as zeros
  if (zero_page_ && !mock::installed && (cudaMemset((void *)sub_addr(0, 0), 0, unit_) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)) {
    set_error("zero page memset failed: %s", cudaGetErrorString(cudaGetLastError()));
    return KVB_ERR_CUDA_;
  }
  handles_.assign((size_t)n_pages_ * n_sub_, 0);
  mapped_.assign((size_t)n_pages_, 0);
  want_.assign((size_t)n_pages_, 0);
  spec_.assign((size_t)n_pages_, 0);
  force_.assign((size_t)n_pages_, 0);
  ret_seq_.assign((size_t)n_pages_, 0);
  {
    // how long an unmapped page keeps its physical handle / an idle handle stays pooled (0 = not at all)
    const char *e = getenv("KVB200_VMM_RETAIN_MS");
    // default 0: the physical mapping follows the bookkeeping (what is unmapped there goes back to the zero page);
    // retention is an opt-in cache for boxes whose driver round trip is slow
    retention_ms_ = e && *e ? atof(e) : 0.0;
    if (retention_ms_ < 0) retention_ms_ = 0;
    const char *pi = getenv("KVB200_POOL_IDLE_MS");
    pool_idle_ms_ = pi && *pi ? atof(pi) : 2000.0;
    if (pool_idle_ms_ < 0) pool_idle_ms_ = 0;
    const char *u = getenv("KVCACHED_GPU_UTILIZATION");
    gpu_util_ = u && *u ? atof(u) : 0.95;
    const char *rf = getenv("KVB200_RETAIN_FREE_FRACTION");
    retain_free_frac_ = rf && *rf ? atof(rf) : 0.20;
    const char *qm = getenv("KVB200_RETAIN_QUIET_MARGIN");
    quiet_margin_ = qm && *qm ? atof(qm) : 0.02;
    const char *fl = getenv("KVB200_POOL_FLOOR");
    pool_floor_ = fl && *fl ? atoll(fl) : 16;
    const char *mx = getenv("KVB200_RETAIN_MAX_FRACTION");
    retain_max_frac_ = mx && *mx ? atof(mx) : 0.10;
    if (retain_max_frac_ < 0) retain_max_frac_ = 0;
    if (retain_max_frac_ > 1) retain_max_frac_ = 1;
    // the cap needs the device size before the VMM thread's first sample
    size_t fr = 0, tot = 0;
    if (mock::installed ? mock_meminfo(&fr, &tot) : (cudaMemGetInfo(&fr, &tot) == cudaSuccess)) {
      mem_free_ = (double)fr;
      mem_total_ = (double)tot;
    } else {
      cudaGetLastError();
    }
  }
  {
    // pages mapped ahead of demand; the reference keeps KVCACHED_MIN_RESERVED_PAGES (5) pre-mapped by default
    const char *e = getenv("KVB200_VMM_PREFETCH");
    int d = e && *e ? atoi(e) : 5;
    prefetch_depth_ = d < 0 ? 0 : (d > 16 ? 16 : d);
  }
  {
    // KVB200_VMM_ASYNC=0: the calling thread runs the driver loop itself and every call returns settled
    const char *e = getenv("KVB200_VMM_ASYNC");
    async_ = !(e && e[0] == '0');
  }
  if (init_next_ < n_units_) {   // background init needs the VMM thread
    if (!async_) {
      for (int64_t u = init_next_; u < n_units_; ++u) {
        CUdeviceptr a = sub_addr(u / n_sub_, (int)(u % n_sub_));
        CU_TRY(d->MemMap(a, unit_, 0, zero_, 0));
        CU_TRY(d->MemSetAccess(a, unit_, &access_, 1));
        zero_ok_[(size_t)u] = 1;
      }
      init_next_ = n_units_;
    }
  }
  if (pool_cap_pages_ < 0) {
    // Physical handles are recycled rather than released per unmap (page.cpp:10-25 creates and releases one per map):
    // the pool cap bounds idle + retained handles; idle ones go back to the device after the retention time
    // (expire()), all of them when the device runs short (phys_free_pages -> pool_release) or on pool_config.
    const char *e = getenv("KVB200_HANDLE_POOL_PAGES");
    int64_t pages = e ? atoll(e) : -1;
    if (pages < 0) {
      // unmapped handles waiting for the next map: never more than KVB200_POOL_MAX_FRACTION (0.10) of the device —
      // beyond that an unmap releases its handle (cuMemRelease) like the reference does on every unmap
      const char *pf = getenv("KVB200_POOL_MAX_FRACTION");
      double frac = pf && *pf ? atof(pf) : 0.10;
      pages = n_pages_;
      if (mem_total_ > 0 && frac >= 0 && frac < 1) pages = std::min<int64_t>(pages, std::max<int64_t>(pool_floor_, (int64_t)(mem_total_ * frac / ((double)unit_ * n_sub_))));
    }
    pool_cap_pages_ = pages;
  }
  created_ = true;
  if (init_next_ < n_units_) {
    Lock lk(qmu_);
    ensure_worker(lk);
  }
  return (int)ranges_.size();
}

CUdeviceptr VmmGroup::sub_addr(int64_t p, int sub) const {
  if (contiguous_) return ranges_[0].base + (size_t)p * unit_;
  if (unified_) return ranges_[(size_t)sub].base + (size_t)p * P_;
  int l = sub >> 1, kv = sub & 1;   // K at off, V at off + size/2 (allocator.cpp:46-52, :178-207)
  return ranges_[(size_t)l].base + (size_t)kv * (size_per_layer_ / 2) + (size_t)p * P_;
}

// ------------------------------------------------------------------------------------------ engine
// FTensor::map (ftensor.cpp:100-118) / FTensor::unmap (:120-140) for every layer / K / V range of one page.
// Driver calls run with the engine lock RELEASED; the page is marked in flight so nobody else touches it.
// One cuMemSetAccess per unit: on this driver (580.x, B200) a single call spanning several mappings costs
// more than the per-unit calls it replaces (tools/vmm_stall2.cu: R=4 -> 4x, R=16 -> 67x per call), so
// adjacent units are deliberately NOT merged.
#define CU_TIMED_L(slot, call)                                               \
  if (!rc) {                                                                 \
    auto t0_ = std::chrono::steady_clock::now();                             \
    CUresult r_ = (call);                                                    \
    tt[slot] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0_).count(); \
    tn[slot] += 1;                                                           \
    if (r_ != CUDA_SUCCESS) {                                                \
      snprintf(emsg, sizeof(emsg), "%s failed: %s (%d) at %s:%d", #call, cu_str(r_), (int)r_, __FILE__, __LINE__); \
      rc = KVB_ERR_CUDA_;                                                    \
    }                                                                        \
  }

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// One place where a page's (wanted, hinted, physical, forced) state changes, so that the counters stay exact:
//   need_map    wanted, not mapped                  need_spec      hinted, not wanted, not mapped
//   need_unmap  mapped, unwanted, unhinted, FORCED  n_retained     mapped, unwanted, unhinted, not forced
//   n_spec_mapped  hinted pages mapped ahead of their event
void VmmGroup::update(int64_t p, int want, int spec, int phys, int force) {
  bool w = want_[(size_t)p], sp = spec_[(size_t)p], ph = mapped_[(size_t)p], f = force_[(size_t)p];
  need_map_ -= (w && !ph);
  need_unmap_ -= (ph && !w && !sp && f);
  n_retained_ -= (ph && !w && !sp && !f);
  need_spec_ -= (sp && !w && !ph);
  n_spec_mapped_ -= (sp && !w && ph);
  if (want >= 0) w = want != 0;
  if (spec >= 0) sp = spec != 0;
  if (phys >= 0 && ph != (phys != 0)) { ph = phys != 0; n_mapped_ += ph ? 1 : -1; }
  if (force >= 0) f = force != 0;
  if (w || sp || !ph) f = false;
  n_want_ += (int)w - (int)want_[(size_t)p];
  want_[(size_t)p] = w;
  spec_[(size_t)p] = sp;
  mapped_[(size_t)p] = ph;
  force_[(size_t)p] = f;
  need_map_ += (w && !ph);
  need_unmap_ += (ph && !w && !sp && f);
  n_retained_ += (ph && !w && !sp && !f);
  need_spec_ += (sp && !w && !ph);
  n_spec_mapped_ += (sp && !w && ph);
  if (n_mapped_ - n_want_ > peak_over_) peak_over_ = n_mapped_ - n_want_;
  if (n_retained_ > peak_retained_) peak_retained_ = n_retained_;
}

// A mapped page that is neither wanted nor hinted any more.
void VmmGroup::make_unwanted(int64_t p) {
  if (!mapped_[(size_t)p] || want_[(size_t)p] || spec_[(size_t)p]) return;
  if (async_ && retention_ms_ > 0 && !force_[(size_t)p] && n_retained_ <= retain_cap_pages()) {
    // (the page counts as retained from the moment it became unwanted, hence <=: at the cap it goes back at once —
    // the newest page is the one the FIFO free list needs last)
    retained_.push_back(Retained{p, ++ret_seq_[(size_t)p], now_ms()});
    // the pool cap bounds idle handles, retained ones included
    while ((int64_t)pool_.size() / n_sub_ + n_retained_ > pool_cap_pages_ && force_one_retained(false)) {}
  } else {
    update(p, -1, -1, -1, 1);
    qu_.push_back(p);
  }
}

// Retained pages are a cache nobody asked for: never more than retain_max_frac_ of the device.
int64_t VmmGroup::retain_cap_pages() const {
  if (mem_total_ <= 0) return n_pages_;
  return (int64_t)(mem_total_ * retain_max_frac_ / ((double)unit_ * n_sub_));
}
void VmmGroup::set_retain_max_fraction(double f) {
  Lock lk(qmu_);
  retain_max_frac_ = f < 0 ? 0 : (f > 1 ? 1 : f);
  while (n_retained_ > retain_cap_pages() && force_one_retained(true)) {}
  if (async_ && need_unmap_ > 0) ensure_worker(lk);
}
void VmmGroup::footprint(int64_t *o) {   // out8
  Lock lk(qmu_);
  o[0] = n_mapped_;
  o[1] = n_want_;
  o[2] = peak_over_;
  o[3] = n_retained_;
  o[4] = retain_cap_pages();
  o[5] = peak_retained_;
  o[6] = n_sub_ ? (int64_t)pool_.size() / n_sub_ : 0;
  o[7] = pool_cap_pages_;
  peak_over_ = n_mapped_ - n_want_;
  peak_retained_ = n_retained_;
}

bool VmmGroup::force_one_retained(bool newest) {
  while (!retained_.empty()) {
    Retained r = newest ? retained_.back() : retained_.front();
    if (newest) retained_.pop_back(); else retained_.pop_front();
    const size_t p = (size_t)r.p;
    if (ret_seq_[p] != r.seq || !mapped_[p] || want_[p] || spec_[p] || force_[p]) continue;   // stale entry
    update(r.p, -1, -1, -1, 1);
    qu_.push_back(r.p);
    return true;
  }
  return false;
}

void VmmGroup::expire(double now) {
  while (!retained_.empty()) {
    const Retained &r = retained_.front();
    const size_t p = (size_t)r.p;
    const bool live = ret_seq_[p] == r.seq && mapped_[p] && !want_[p] && !spec_[p] && !force_[p];
    // the clock runs while the allocator is quiet: as long as pages keep being mapped, the FIFO free list will
    // come back to this one, and memory others need is handed over by the reserve rule below, not by age
    if (live && now - std::max(r.t_ms, last_map_ms_) < retention_ms_) break;
    if (live) {
      update(r.p, -1, -1, -1, 1);
      qu_.push_back(r.p);
      release_debt_ += n_sub_;          // its handle goes back to the device, not to the pool
    }
    retained_.pop_front();
  }
  // somebody else took memory: retained pages go back until the reserve is whole again (newest first, like a steal)
  if (mem_total_ > 0 && n_retained_ > 0) {
    const double reserve = reserve_fraction(now) * mem_total_;
    double short_by = reserve - mem_free_ - (double)(need_unmap_ + release_debt_ / n_sub_) * (double)unit_ * n_sub_;
    while (short_by > 0 && force_one_retained(true)) {
      release_debt_ += n_sub_;
      short_by -= (double)unit_ * n_sub_;
    }
  }
  // idle pooled handles above the floor go back to the device too (oldest first, a few per wake-up)
  for (int k = 0; k < 8 && (int64_t)pool_.size() > pool_floor_ * n_sub_ &&
                  now - std::max(pool_.front().t_ms, last_map_ms_) >= pool_idle_ms_; ++k) {
    g_drv.MemRelease(pool_.front().h);
    pool_.pop_front();
    n_released_++; mem_free_ += (double)unit_;
  }
}

// Device memory is sampled on the VMM thread only (cudaMemGetInfo is a driver round trip like the VMM calls and
// stalls like them); between samples our own creates and releases are accounted.
void VmmGroup::poll_meminfo(Lock &lk) {
  lk.unlock();
  size_t fr = 0, tot = 0;
  bool ok;
  if (mock::installed) {
    ok = mock_meminfo(&fr, &tot);
  } else {
    ensure_device(dev_);
    ok = cudaMemGetInfo(&fr, &tot) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  lk.lock();
  if (ok) {
    mem_free_ = (double)fr;
    mem_total_ = (double)tot;
    mem_t_ms_ = now_ms();
    // everybody else's footprint = used memory that is not our handles; a rise of one page over the lowest level
    // seen since the last rise means somebody is taking memory: the larger reserve applies for the next 10 s
    const double ours = (double)(n_created_ - n_released_) * (double)unit_;
    const double others = mem_total_ - mem_free_ - ours;
    if (others_base_ < 0 || others < others_base_) others_base_ = others;
    if (others > others_base_ + (double)unit_ * n_sub_) {
      contended_until_ms_ = mem_t_ms_ + 10000.0;
      others_base_ = others;
    }
  }
}

double VmmGroup::reserve_fraction(double now) const {
  const double headroom = 1.0 - gpu_util_;
  if (now < contended_until_ms_) return std::max(retain_free_frac_, headroom);
  return std::min(std::max(retain_free_frac_, headroom), headroom + quiet_margin_);
}

bool VmmGroup::cached_meminfo(size_t *free_bytes, size_t *total_bytes) {
  Lock lk(qmu_);
  // the VMM thread samples every 5 ms while it works but only every 50 ms while idle: an older sample is not served
  // (the caller then asks the driver itself, at most once per KVB200_MEMINFO_REFRESH_US), so that memory another
  // process just released is seen within milliseconds
  if (!worker_running_ || now_ms() - mem_t_ms_ > 10.0 || mem_total_ <= 0) return false;
  *free_bytes = (size_t)(mem_free_ > 0 ? mem_free_ : 0);
  *total_bytes = (size_t)mem_total_;
  return true;
}

// Headroom for one more page under KVCACHED_GPU_UTILIZATION, like the bookkeeping's own available-memory
// arithmetic (page_allocator.cpp:442-455), from the VMM thread's last sample.
bool VmmGroup::can_create(bool soft) {
  if (mem_total_ <= 0) return true;   // never sampled (synchronous mode): let the driver decide
  const double reserve = soft ? reserve_fraction(now_ms()) : (1.0 - gpu_util_);
  return mem_free_ - (double)unit_ * n_sub_ >= mem_total_ * reserve;
}

void VmmGroup::set_retain_free_fraction(double f, double quiet_margin) {
  Lock lk(qmu_);
  retain_free_frac_ = f < 0 ? 0 : (f > 1 ? 1 : f);
  if (quiet_margin >= 0) quiet_margin_ = quiet_margin > 1 ? 1 : quiet_margin;
}

int VmmGroup::transition(Lock &lk, int64_t p, bool to_mapped) {
  // handles for a map come from the pool (under the lock); cuMemCreate only when it is empty
  std::vector<CUmemGenericAllocationHandle> hs((size_t)n_sub_, 0);
  int need_create = 0;
  if (to_mapped) {
    for (int s = 0; s < n_sub_; ++s) {
      if (!pool_.empty()) { hs[(size_t)s] = pool_.back().h; pool_.pop_back(); } else need_create++;
    }
  } else {
    for (int s = 0; s < n_sub_; ++s) hs[(size_t)s] = handles_[(size_t)p * n_sub_ + s];
  }
  inflight_ = p;
  inflight_to_mapped_ = to_mapped;
  const bool sync_first = !to_mapped && sync_before_unmap_;
  lk.unlock();
  // VMM driver calls of two threads do not overlap on this stack, they collide (tools/vmm_stall2.cu: two threads
  // reach 73 remaps/s together, one reaches 1 312): groups of one process take turns.
  static std::mutex driver_turn;
  std::lock_guard<std::mutex> turn(driver_turn);
  double tt[5] = {0, 0, 0, 0, 0};
  int64_t tn[5] = {0, 0, 0, 0, 0};
  char emsg[384] = "";
  int rc = 0, created = 0;
  if (!mock::installed) ensure_device(dev_);
  // the engine may still have kernels in flight on pages about to be unmapped (page_allocator.cpp:671-673)
  if (sync_first && !mock::installed) cudaDeviceSynchronize();
  if (to_mapped) {
    for (int s = 0; s < n_sub_ && !rc; ++s)
      if (!hs[(size_t)s]) {
        CU_TIMED_L(0, g_drv.MemCreate(&hs[(size_t)s], unit_, &prop_, 0));
        if (!rc) created++; else hs[(size_t)s] = 0;
      }
  }
  int done = 0;
  for (int s = 0; s < n_sub_ && !rc; ++s) {
    CUdeviceptr a = sub_addr(p, s);
    if (zero_ok_[(size_t)p * n_sub_ + s]) { CU_TIMED_L(2, g_drv.MemUnmap(a, unit_)); }   // (background init: not backed yet)
    if (to_mapped || zero_page_) {
      CU_TIMED_L(3, g_drv.MemMap(a, unit_, 0, to_mapped ? hs[(size_t)s] : zero_, 0));
      CU_TIMED_L(4, g_drv.MemSetAccess(a, unit_, &access_, 1));
      if (!rc) zero_ok_[(size_t)p * n_sub_ + s] = 1;
    } else if (!rc) {
      zero_ok_[(size_t)p * n_sub_ + s] = 0;      // no zero page: the unit is simply unmapped now
    }
    if (!rc) done++;
  }
  if (rc && to_mapped) {
    // a map that failed half way (per-layer layout): the units already switched go back to the zero page, so that
    // the page is recorded as it really is and its handles can return to the pool unmapped
    for (int s = 0; s <= done && s < n_sub_; ++s) {
      CUdeviceptr a = sub_addr(p, s);
      g_drv.MemUnmap(a, unit_);
      if (zero_page_ && g_drv.MemMap(a, unit_, 0, zero_, 0) == CUDA_SUCCESS) g_drv.MemSetAccess(a, unit_, &access_, 1);
      else zero_ok_[(size_t)p * n_sub_ + s] = 0;
    }
  }
  lk.lock();
  inflight_ = -1;
  for (int i = 0; i < 5; ++i) { t_us_[i] += tt[i]; t_n_[i] += tn[i]; }
  n_created_ += created;
  mem_free_ -= (double)created * (double)unit_;
  if (rc) {
    if (!werr_) { werr_ = rc; werr_msg_ = emsg; }
    // leave the page as it was recorded; handles taken from the pool go back
    if (to_mapped) for (auto h : hs) if (h) pool_.push_back(PoolEnt{h, now_ms()});
    update(p, mapped_[(size_t)p], 0, -1, 0);
    return rc;
  }
  if (to_mapped) {
    for (int s = 0; s < n_sub_; ++s) handles_[(size_t)p * n_sub_ + s] = hs[(size_t)s];
  } else {
    for (int s = 0; s < n_sub_; ++s) {
      handles_[(size_t)p * n_sub_ + s] = 0;
      if (release_debt_ <= 0 && (int64_t)pool_.size() < pool_cap_pages_ * n_sub_) {
        pool_.push_back(PoolEnt{hs[(size_t)s], now_ms()});
      } else {
        if (release_debt_ > 0) release_debt_--;
        auto t0 = std::chrono::steady_clock::now();
        g_drv.MemRelease(hs[(size_t)s]);
        t_us_[1] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        t_n_[1] += 1;
        n_released_++; mem_free_ += (double)unit_;
      }
    }
  }
  set_phys(p, to_mapped);
  if (to_mapped) make_unwanted(p);   // wanted no more while it was being mapped: retain or unmap
  return 0;
}

// One page transition, maps first.  A map that finds the pool empty while unmaps are pending lets one
// unmap go first: it yields a handle without growing the footprint, and the work is due anyway.
bool VmmGroup::process_next(Lock &lk) {
  while (!qm_.empty()) {
    int64_t p = qm_.front();
    if (!(want_[(size_t)p] && !mapped_[(size_t)p])) { qm_.pop_front(); continue; }   // cancelled or done
    if ((int64_t)pool_.size() < n_sub_) {
      // no idle handle: a pending unmap yields one; else create within the headroom; else take the page
      // retained last (the FIFO free list needs it last) — only then create regardless and let the driver say no
      if (need_unmap_ > 0) break;
      if (!can_create(true) && force_one_retained(true)) break;   // (with nothing retained: create down to the headroom,
    }                                                             //  and past it only to let the driver say no)
    qm_.pop_front();
    transition(lk, p, true);
    return true;
  }
  while (need_map_ == 0 && !qs_.empty()) {   // hinted maps: only when no wanted map is waiting
    int64_t p = qs_.front();
    if (!(spec_[(size_t)p] && !want_[(size_t)p] && !mapped_[(size_t)p])) { qs_.pop_front(); continue; }
    if ((int64_t)pool_.size() < n_sub_) {
      if (need_unmap_ > 0) break;                      // an unmap yields the handle
      if (!can_create(true)) {                         // map-ahead never eats into the reserve
        if (!force_one_retained(true)) {               // no memory for it: the hint is dropped, not parked
          qs_.pop_front();
          update(p, -1, 0, -1);
          continue;
        }
        break;
      }
    }
    qs_.pop_front();
    transition(lk, p, true);
    return true;
  }
  while (!qu_.empty()) {
    int64_t p = qu_.front();
    qu_.pop_front();
    if (!(!want_[(size_t)p] && !spec_[(size_t)p] && mapped_[(size_t)p] && force_[(size_t)p])) continue;
    transition(lk, p, false);
    return true;
  }
  return false;
}

void VmmGroup::worker_main() {
  Lock lk(qmu_);
  while (true) {
    cv_work_.wait_for(lk, std::chrono::milliseconds(50),
                      [&] { return stop_ || need_map_ > 0 || need_unmap_ > 0 || need_spec_ > 0 || init_next_ < n_units_; });
    if (stop_ && need_map_ == 0 && need_unmap_ == 0) break;
    if (now_ms() - mem_t_ms_ > 5.0) poll_meminfo(lk);
    expire(now_ms());
    if (need_map_ == 0 && need_unmap_ == 0 && need_spec_ == 0 && init_next_ < n_units_ && !stop_) {
      init_some(lk);      // background zero-page init: a few units, then look for real work again
      cv_done_.notify_all();
      continue;
    }
    if (need_map_ == 0 && need_unmap_ == 0 && need_spec_ == 0) {   // timer wake-up, nothing due
      lk.unlock();
      {
        std::lock_guard<std::mutex> hk(hook_mu_);
        if (idle_hook_) idle_hook_();
      }
      lk.lock();
      continue;
    }
    bool did = process_next(lk);
    cv_done_.notify_all();
    if (!did) {
      // counters say there is work but the queues are empty: cannot happen; avoid spinning if it does
      if (stop_) break;
      need_map_ = need_unmap_ = need_spec_ = 0;
      for (int64_t p = 0; p < n_pages_; ++p) {
        bool w = want_[(size_t)p], sp = spec_[(size_t)p], ph = mapped_[(size_t)p], f = force_[(size_t)p];
        if (w && !ph) { need_map_++; qm_.push_back(p); }
        if (!w && !sp && ph && f) { need_unmap_++; qu_.push_back(p); }
        if (sp && !w && !ph) { need_spec_++; qs_.push_back(p); }
      }
    }
  }
  worker_running_ = false;
  cv_done_.notify_all();
}

// Background zero-page init (VMM thread, idle time): up to 16 units per call, lock released around the driver calls.
// Units the bookkeeping mapped meanwhile are skipped (they hold their handle; their unmap puts them on the zero page).
void VmmGroup::init_some(Lock &lk) {
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < 16 && init_next_ < n_units_ && need_map_ == 0; ++k) {
    const int64_t u = init_next_;
    if (!zero_ok_[(size_t)u]) {
      const CUdeviceptr a = sub_addr(u / n_sub_, (int)(u % n_sub_));
      lk.unlock();
      CUresult r = g_drv.MemMap(a, unit_, 0, zero_, 0);
      if (r == CUDA_SUCCESS) r = g_drv.MemSetAccess(a, unit_, &access_, 1);
      lk.lock();
      if (r != CUDA_SUCCESS) {
        if (!werr_) { werr_ = KVB_ERR_CUDA_; werr_msg_ = std::string("zero-page init failed: ") + cu_str(r); }
        init_next_ = n_units_;
        return;
      }
      zero_ok_[(size_t)u] = 1;
    }
    init_next_ = u + 1;
  }
  init_ms_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int VmmGroup::wait_init() {
  if (!created_) return 0;
  Lock lk(qmu_);
  if (init_next_ >= n_units_) return 0;
  ensure_worker(lk);
  cv_done_.wait(lk, [&] { return init_next_ >= n_units_ || werr_; });
  return werr_ ? werr_ : 0;
}

void VmmGroup::init_progress(int64_t *out3) {
  Lock lk(qmu_);
  out3[0] = init_next_;
  out3[1] = n_units_;
  out3[2] = (int64_t)(init_ms_ * 1000.0);
}

void VmmGroup::set_idle_hook(std::function<void()> fn) {
  std::lock_guard<std::mutex> hk(hook_mu_);
  idle_hook_ = std::move(fn);
}

void VmmGroup::set_async(bool on) {
  Lock lk(qmu_);
  if (on == async_) return;
  if (!on) {
    lk.unlock();
    stop_worker();
    lk.lock();
  }
  async_ = on;
}
void VmmGroup::stop_worker() {
  {
    Lock lk(qmu_);
    if (!worker_running_ && !worker_.joinable()) return;
    stop_ = true;
    cv_work_.notify_all();
  }
  if (worker_.joinable()) worker_.join();
  Lock lk(qmu_);
  stop_ = false;
}

// Inline reconciliation by the calling thread (synchronous mode, or nobody to wait for).
void VmmGroup::ensure_worker(Lock &lk) {
  if (!worker_running_) {
    if (worker_.joinable()) { lk.unlock(); worker_.join(); lk.lock(); }
    worker_running_ = true;
    worker_ = std::thread([this] { worker_main(); });
  }
  cv_work_.notify_all();
}

int VmmGroup::drain(Lock &lk, bool maps_only) {
  if (async_ && created_) {
    ensure_worker(lk);
    // a page being unmapped right now that is wanted again does not show in need_map_ until the unmap is done
    cv_done_.wait(lk, [&] {
      if (werr_) return true;
      if (need_map_ > 0) return false;
      if (maps_only) return !(inflight_ >= 0 && !inflight_to_mapped_ && want_[(size_t)inflight_]);
      // a hinted (unwanted) map in flight is nobody's business
      return need_unmap_ == 0 && (inflight_ < 0 || (inflight_to_mapped_ && !want_[(size_t)inflight_]));
    });
  } else {
    while (!werr_ && (need_map_ > 0 || (!maps_only && need_unmap_ > 0)))
      if (!process_next(lk)) break;
  }
  if (werr_) {
    int rc = werr_;
    set_error("%s", werr_msg_.c_str());
    werr_ = 0;
    werr_msg_.clear();
    return rc;
  }
  return 0;
}

int VmmGroup::enqueue_events(const int64_t *ev, int64_t n) {
  if (!created_) return 0;
  if (n <= 0) return 1;
  if (!driver()) return KVB_ERR_CUDA_;
  Lock lk(qmu_);
  for (int64_t i = 0; i < n; ++i) {
    int64_t p = ev[i] >> 1;
    bool un = ev[i] & 1;
    if (p < 0 || p >= n_pages_) { set_error("%s: page %ld out of range [0, %ld)", un ? "unmap" : "map", (long)p, (long)n_pages_); return KVB_ERR_INVALID_; }
    bool w = want_[(size_t)p];
    if (w == !un) continue;   // "already mapped" / "not mapped" is tolerated (allocator.cpp:174-177, ftensor.cpp:104-107)
    if (!un && mapped_[(size_t)p] && inflight_ != p) {
      if (spec_[(size_t)p]) n_prefetch_hits_++;          // mapped ahead: no wait
      else n_retention_hits_++;                          // still holds the handle it had: nothing to do
    }
    set_want(p, !un);
    if (!un) last_map_ms_ = now_ms();
    if (un) {
      if (mapped_[(size_t)p]) make_unwanted(p);          // retained, or its unmap is queued (hinted pages stay)
      else if (inflight_ != p) n_elided_++;              // met its own pending map: nothing to do physically
    } else {
      if (!mapped_[(size_t)p] || inflight_ == p) qm_.push_back(p);
      else n_elided_++;
    }
  }
  n_map_calls_++;
  if (async_ && (need_map_ > 0 || need_unmap_ > 0 || n_retained_ > 0)) ensure_worker(lk);
  return 1;
}

int VmmGroup::wait_maps() {
  if (!created_) return 0;
  Lock lk(qmu_);
  return drain(lk, true);
}
int VmmGroup::fence() {
  if (!created_) return 0;
  Lock lk(qmu_);
  return drain(lk, false);
}
int VmmGroup::settle() {
  if (!created_) return 0;
  Lock lk(qmu_);
  while (force_one_retained(false)) {}
  return drain(lk, false);
}

void VmmGroup::set_retention_ms(double ms) {
  Lock lk(qmu_);
  retention_ms_ = ms < 0 ? 0 : ms;
  if (retention_ms_ == 0) while (force_one_retained(false)) {}
  if (async_ && need_unmap_ > 0) ensure_worker(lk);
}

void VmmGroup::set_prefetch_depth(int depth) {
  Lock lk(qmu_);
  prefetch_depth_ = depth < 0 ? 0 : (depth > 16 ? 16 : depth);
}

// New hint window: pages that left it without being wanted go back to the zero page, new ones are queued for
// the VMM thread's idle time.  No waiting here.
void VmmGroup::set_prefetch_hints(const int64_t *pages, int64_t n) {
  if (!created_) return;
  Lock lk(qmu_);
  if (!async_ || prefetch_depth_ <= 0) n = 0;
  if (n > prefetch_depth_) n = prefetch_depth_;
  bool same = (int64_t)hints_.size() == n;
  for (int64_t i = 0; same && i < n; ++i) same = hints_[(size_t)i] == pages[i];
  if (same) return;
  std::vector<int64_t> old;
  old.swap(hints_);
  for (int64_t i = 0; i < n; ++i)
    if (pages[i] >= 0 && pages[i] < n_pages_) hints_.push_back(pages[i]);
  for (int64_t p : old) update(p, -1, 0, -1);
  for (int64_t p : hints_) {
    update(p, -1, 1, -1);
    if (!want_[(size_t)p] && !mapped_[(size_t)p]) qs_.push_back(p);
  }
  for (int64_t p : old)
    if (!spec_[(size_t)p]) make_unwanted(p);   // mapped ahead in vain: retained or unmapped (in flight: transition() does it)
  if (need_spec_ > 0 || need_unmap_ > 0) ensure_worker(lk);
}

int64_t VmmGroup::prefetched_pages(int64_t *out, int64_t cap) {
  Lock lk(qmu_);
  int64_t n = 0;
  for (int64_t p : hints_)
    if (spec_[(size_t)p] && !want_[(size_t)p] && mapped_[(size_t)p]) {
      if (n >= cap) return -5;
      out[n++] = p;
    }
  return n;
}

int64_t VmmGroup::pending_unmaps() {
  Lock lk(qmu_);
  return need_unmap_;
}

int VmmGroup::map_pages(const int64_t *ids, int64_t n) {
  if (!created_) return 0;
  std::vector<int64_t> ev((size_t)std::max<int64_t>(n, 0));
  for (int64_t i = 0; i < n; ++i) ev[(size_t)i] = ids[i] * 2;
  int rc = enqueue_events(ev.data(), n);
  if (rc < 0) return rc;
  rc = fence();
  return rc ? rc : 1;
}
int VmmGroup::unmap_pages(const int64_t *ids, int64_t n) {
  if (!created_) return 0;
  std::vector<int64_t> ev((size_t)std::max<int64_t>(n, 0));
  for (int64_t i = 0; i < n; ++i) ev[(size_t)i] = ids[i] * 2 + 1;
  int rc = enqueue_events(ev.data(), n);
  if (rc < 0) return rc;
  n_unmap_calls_++;
  {
    // the direct API keeps the reference's meaning: these pages are on the zero page when the call returns
    Lock lk(qmu_);
    for (int64_t i = 0; i < n; ++i) {
      int64_t p = ids[i];
      if (p >= 0 && p < n_pages_ && mapped_[(size_t)p] && !want_[(size_t)p] && !spec_[(size_t)p] && !force_[(size_t)p]) {
        update(p, -1, -1, -1, 1);
        qu_.push_back(p);
      }
    }
  }
  rc = fence();
  return rc ? rc : 1;
}
int VmmGroup::apply_events(const int64_t *ev, int64_t n) {
  if (!created_) return 0;
  int rc = enqueue_events(ev, n);
  if (rc < 0) return rc;
  rc = fence();
  return rc ? rc : 1;
}

void VmmGroup::pool_config(int64_t max_cached_pages) {
  Lock lk(qmu_);
  pool_cap_pages_ = max_cached_pages;
  while ((int64_t)pool_.size() / n_sub_ + n_retained_ > pool_cap_pages_ && force_one_retained(false)) {}
  while ((int64_t)pool_.size() > pool_cap_pages_ * n_sub_ && g_drv.ok) {
    g_drv.MemRelease(pool_.front().h);
    pool_.pop_front();
    n_released_++; mem_free_ += (double)unit_;
  }
  if (async_ && need_unmap_ > 0) ensure_worker(lk);
}
int VmmGroup::pool_prefill(int64_t n_pages) {
  if (!created_) { set_error("pool_prefill: tensors not created"); return KVB_ERR_STATE_; }
  if (!driver()) return KVB_ERR_CUDA_;
  int rc = fence();
  if (rc) return rc;
  cudaSetDevice(dev_);
  Lock lk(qmu_);
  int64_t want = std::min<int64_t>(n_pages, pool_cap_pages_) * n_sub_;
  while ((int64_t)pool_.size() < want) {
    CUmemGenericAllocationHandle h;
    CU_TRY(g_drv.MemCreate(&h, unit_, &prop_, 0));
    n_created_++;
    pool_.push_back(PoolEnt{h, now_ms()});
  }
  return 0;
}
// pages' worth of physical memory this group holds without a mapping that is wanted: pooled handles plus
// pages whose unmap is still queued (their handles come back to the pool)
int64_t VmmGroup::pooled_pages() {
  Lock lk(qmu_);
  return n_sub_ ? (int64_t)pool_.size() / n_sub_ + need_unmap_ + n_retained_ + n_spec_mapped_ : 0;
}
int64_t VmmGroup::pool_release(int64_t n_pages) {
  Lock lk(qmu_);
  int64_t n = 0;
  while (n < n_pages * n_sub_ && !pool_.empty() && g_drv.ok) {
    g_drv.MemRelease(pool_.front().h);
    pool_.pop_front();
    n_released_++; mem_free_ += (double)unit_;
    n++;
  }
  // still short: retained pages give their memory back too (the VMM thread unmaps them, their handles are released)
  int64_t got = n_sub_ ? n / n_sub_ : 0;
  for (int64_t k = got; k < n_pages && force_one_retained(false); ++k) { release_debt_ += n_sub_; got++; }
  if (async_ && need_unmap_ > 0) ensure_worker(lk);
  return got;
}

void VmmGroup::engine_stats(int64_t *s) {
  Lock lk(qmu_);
  s[0] = n_elided_;
  s[1] = n_prefetch_hits_;
  s[2] = n_spec_mapped_;
  s[3] = need_map_;
  s[4] = need_unmap_;
  s[5] = need_spec_;
  s[6] = n_retained_;
  s[7] = n_retention_hits_;
}

void VmmGroup::timing(double *out10) {
  Lock lk(qmu_);
  for (int i = 0; i < 5; ++i) { out10[i] = t_us_[i]; out10[5 + i] = (double)t_n_[i]; }
}

void VmmGroup::stats(int64_t *s) {
  Lock lk(qmu_);
  s[0] = n_mapped_;
  s[1] = (int64_t)pool_.size();
  s[2] = n_created_;
  s[3] = n_released_;
  s[4] = n_map_calls_;
  s[5] = n_unmap_calls_;
  s[6] = n_mapped_ * (int64_t)unit_ * n_sub_;
  s[7] = (int64_t)unit_ * n_sub_;
}
int64_t VmmGroup::mapped_pages(int64_t *out, int64_t cap) {
  if (created_) fence();   // inspection sees the settled state; pages mapped ahead of their event are not "mapped"
  Lock lk(qmu_);
  int64_t n = 0;
  for (int64_t p = 0; p < n_pages_; ++p)
    if (mapped_[(size_t)p] && want_[(size_t)p]) {
      if (n >= cap) return -5;
      out[n++] = p;
    }
  return n;
}

// Raw-driver control for the map/unmap latency metric: the same three driver calls per unit as
// map_one / unmap_one on pre-created handles over a private range, nothing else in the loop
// (cf. the reference's benchmarks/bench_vmm/bench_vmm.cpp:139-228).
int VmmGroup::raw_control(int64_t n_pages, int64_t iters, double *map_us, double *unmap_us) {
  const Driver *d = driver();
  if (!d) return KVB_ERR_CUDA_;
  if (!created_) { set_error("raw_control: tensors not created"); return KVB_ERR_STATE_; }
  fence();   // driver VMM calls of two threads do not overlap, they collide
  cudaSetDevice(dev_);
  size_t n_units = (size_t)n_pages * n_sub_;
  size_t bytes = n_units * unit_;
  CUdeviceptr base = 0;
  CU_TRY(d->MemAddressReserve(&base, bytes, 2u << 20, 0, 0));
  std::vector<CUmemGenericAllocationHandle> hs(n_units);
  for (auto &h : hs) CU_TRY(d->MemCreate(&h, unit_, &prop_, 0));
  for (size_t u = 0; u < n_units; ++u) CU_TRY(d->MemMap(base + u * unit_, unit_, 0, zero_, 0));
  CU_TRY(d->MemSetAccess(base, bytes, &access_, 1));
  using clk = std::chrono::steady_clock;
  for (int64_t it = 0; it < iters; ++it) {
    auto t0 = clk::now();
    for (size_t u = 0; u < n_units; ++u) {
      CUdeviceptr a = base + u * unit_;
      CU_TRY(d->MemUnmap(a, unit_));
      CU_TRY(d->MemMap(a, unit_, 0, hs[u], 0));
      CU_TRY(d->MemSetAccess(a, unit_, &access_, 1));
    }
    auto t1 = clk::now();
    for (size_t u = 0; u < n_units; ++u) {
      CUdeviceptr a = base + u * unit_;
      CU_TRY(d->MemUnmap(a, unit_));
      CU_TRY(d->MemMap(a, unit_, 0, zero_, 0));
      CU_TRY(d->MemSetAccess(a, unit_, &access_, 1));
    }
    auto t2 = clk::now();
    map_us[it] = std::chrono::duration<double, std::micro>(t1 - t0).count();
    unmap_us[it] = std::chrono::duration<double, std::micro>(t2 - t1).count();
  }
  d->MemUnmap(base, bytes);
  d->MemAddressFree(base, bytes);
  for (auto h : hs) d->MemRelease(h);
  return 0;
}

#ifdef KVB_TEST_HOOKS
#include "kvb_testhooks_probe.inc"  // libkvb200_test.so only: per-call driver cost probes
#endif

// ------------------------------------------------------------------------------------------ registry
struct Registry {
  std::mutex mu;
  bool inited = false;
  int dev = 0;
  int64_t page_size = 2 << 20;
  bool contiguous = true;
  std::map<int64_t, std::unique_ptr<VmmGroup>> groups;
};
static Registry &reg() {
  static Registry r;
  return r;
}

int vmm_init(int dev, int64_t page_size, bool contiguous) {
  Registry &r = reg();
  std::lock_guard<std::mutex> lk(r.mu);
  if (page_size == 0) page_size = 2 << 20;
  if (page_size < 0 || page_size % (2 << 20) != 0) {   // the reference abort()s (allocator.cpp:84-90)
    set_error("Invalid page size: %ld, must be a multiple of 2MB (2097152 bytes)", (long)page_size);
    return KVB_ERR_INVALID_;
  }
  if (dev < 0) {
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  }
  r.groups.clear();   // re-initialising drops existing allocators (allocator.cpp:75-78)
  r.inited = true;
  r.dev = dev;
  r.page_size = page_size;
  r.contiguous = contiguous;
  return 0;
}
int vmm_shutdown() {
  Registry &r = reg();
  std::lock_guard<std::mutex> lk(r.mu);
  r.groups.clear();
  r.inited = false;
  return 0;
}
bool vmm_initialized() { return reg().inited; }
int64_t vmm_page_size() { return reg().page_size; }
bool vmm_contiguous() { return reg().contiguous; }
int vmm_device() { return reg().dev; }
int ensure_device(int dev) {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == dev) return 0;
  return cudaSetDevice(dev) == cudaSuccess ? 0 : -2;
}

VmmGroup *vmm_group(int64_t group_id, bool create_if_missing) {
  Registry &r = reg();
  std::lock_guard<std::mutex> lk(r.mu);
  if (!r.inited) { set_error("kvb_init() must be called first"); return nullptr; }
  auto it = r.groups.find(group_id);
  if (it != r.groups.end()) return it->second.get();
  if (!create_if_missing) return nullptr;
  auto g = std::make_unique<VmmGroup>(r.dev, (size_t)r.page_size, r.contiguous);   // lazily, allocator.cpp:101-114
  VmmGroup *p = g.get();
  r.groups[group_id] = std::move(g);
  return p;
}

}  // namespace kvb
