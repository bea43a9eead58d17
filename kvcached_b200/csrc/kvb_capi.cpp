// kvb_capi.cpp — the extern "C" surface of include/kvb200.h that is not the block manager
// (kvb_manager.cpp): library info, the VMM registry / virtual KV tensors, page arithmetic helpers and
// the KV scatter / gather launchers.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "kvb_internal.h"
#include "kvb_vmm.h"
#ifdef KVB_TEST_HOOKS
#include "../../include/kvb200_test.h"
#endif

using namespace kvb;
namespace kvb { extern std::atomic<int64_t> g_total_launches; }

#define CUDA_TRY(call)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      set_error("%s failed: %s at %s:%d", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return KVB_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

extern "C" {

const char *kvb_last_error(void) { return get_error(); }
int kvb_abi_version(void) { return 1; }

// ------------------------------------------------------------------------------------------ VMM
int kvb_init(int device_index, int64_t page_size, int contiguous_layout) {
  return vmm_init(device_index, page_size, contiguous_layout != 0);
}
int kvb_shutdown(void) { return vmm_shutdown(); }
int kvb_is_initialized(void) { return vmm_initialized() ? 1 : 0; }
int64_t kvb_page_size(void) { return vmm_page_size(); }
int kvb_contiguous_layout(void) { return vmm_contiguous() ? 1 : 0; }

int kvb_create_kv_tensors(int64_t group_id, int64_t size_per_layer, int64_t num_layers, int64_t num_kv_buffers,
                          int unified_pool, uint64_t *out_ptrs, int64_t *out_bytes, int cap) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  int n = g->create(size_per_layer, num_layers, num_kv_buffers, unified_pool != 0);
  if (n < 0) return n;
  if (n > cap) { set_error("create_kv_tensors: %d ranges, caller buffer holds %d", n, cap); return KVB_ERR_OVERFLOW; }
  for (int i = 0; i < n; ++i) {
    out_ptrs[i] = (uint64_t)g->ranges()[(size_t)i].base;
    out_bytes[i] = (int64_t)g->ranges()[(size_t)i].bytes;
  }
  return n;
}

int kvb_kv_tensors_created(int64_t group_id) {
  VmmGroup *g = vmm_group(group_id, vmm_initialized());
  return g && g->created() ? 1 : 0;
}

static int by_offsets(int64_t group_id, const int64_t *offsets, int64_t n, bool unmap) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  if (!g->created()) { set_error("try to %s KV tensors when KV tensors are not created", unmap ? "unmap from" : "map to"); return 0; }
  std::vector<int64_t> ids((size_t)n);
  for (int64_t i = 0; i < n; ++i) ids[(size_t)i] = g->offset_to_page(offsets[i]);
  return unmap ? g->unmap_pages(ids.data(), n) : g->map_pages(ids.data(), n);
}
int kvb_map_to_kv_tensors(int64_t group_id, const int64_t *offsets, int64_t n) { return by_offsets(group_id, offsets, n, false); }
int kvb_unmap_from_kv_tensors(int64_t group_id, const int64_t *offsets, int64_t n) { return by_offsets(group_id, offsets, n, true); }

int kvb_map_pages(int64_t group_id, const int64_t *page_ids, int64_t n) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->map_pages(page_ids, n);
}
int kvb_unmap_pages(int64_t group_id, const int64_t *page_ids, int64_t n) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->unmap_pages(page_ids, n);
}
int kvb_apply_page_events(int64_t group_id, const int64_t *events, int64_t n) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->apply_events(events, n);
}
int kvb_vmm_enqueue_page_events(int64_t group_id, const int64_t *events, int64_t n) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->enqueue_events(events, n);
}
int kvb_vmm_wait_maps(int64_t group_id) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->wait_maps();
}
int kvb_vmm_fence(int64_t group_id) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->fence();
}
int kvb_vmm_settle(int64_t group_id) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->settle();
}
int kvb_vmm_set_retention_ms(int64_t group_id, double ms) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  g->set_retention_ms(ms);
  return 0;
}
int kvb_vmm_set_retain_free_fraction(int64_t group_id, double f, double quiet_margin) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  g->set_retain_free_fraction(f, quiet_margin);
  return 0;
}
int kvb_vmm_set_retain_max_fraction(int64_t group_id, double f) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  g->set_retain_max_fraction(f);
  return 0;
}
int kvb_vmm_zero_init_mode(int mode) {
  int old = g_zero_init_mode.load();
  if (mode >= -1 && mode <= 2) g_zero_init_mode.store(mode);
  return old;
}
int kvb_vmm_wait_init(int64_t group_id) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  return g->wait_init();
}
int kvb_vmm_init_progress(int64_t group_id, int64_t *out3) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g || !out3) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  g->init_progress(out3);
  return 0;
}
int kvb_vmm_footprint(int64_t group_id, int64_t *out8) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g || !out8) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  g->footprint(out8);
  return 0;
}
int kvb_vmm_set_async(int64_t group_id, int on) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  g->set_async(on != 0);
  return 0;
}
int64_t kvb_vmm_pending_unmaps(int64_t group_id) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  return g->pending_unmaps();
}
int kvb_vmm_set_prefetch_hints(int64_t group_id, const int64_t *pages, int64_t n) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  g->set_prefetch_hints(pages, n);
  return 0;
}
#ifdef KVB_TEST_HOOKS   /* include/kvb200_test.h — libkvb200_test.so only */
int kvb_vmm_mock_driver(int64_t total_bytes, int latency_us) { return vmm_mock_driver(total_bytes, latency_us); }
int kvb_vmm_mock_set_other_bytes(int64_t bytes) { vmm_mock_set_other_bytes(bytes); return 0; }
int kvb_vmm_mock_stats(int64_t *out10) { if (!out10) return KVB_ERR_INVALID; vmm_mock_stats(out10); return 0; }
const char *kvb_vmm_mock_last_violation(void) { return vmm_mock_last_violation(); }
#endif
int kvb_vmm_prefetch_config(int64_t group_id, int depth) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  g->set_prefetch_depth(depth);
  if (depth <= 0) g->set_prefetch_hints(nullptr, 0);
  return 0;
}
int64_t kvb_vmm_prefetched_pages(int64_t group_id, int64_t *out, int64_t cap) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  return g->prefetched_pages(out, cap);
}
int kvb_vmm_engine_stats(int64_t group_id, int64_t *out8) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  g->engine_stats(out8);
  return 0;
}
int kvb_vmm_pool_config(int64_t group_id, int64_t max_cached) {
  VmmGroup *g = vmm_group(group_id, true);
  if (!g) return KVB_ERR_STATE;
  std::lock_guard<std::mutex> lk(g->mu);
  g->pool_config(max_cached);
  return 0;
}
int kvb_vmm_pool_prefill(int64_t group_id, int64_t n) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->pool_prefill(n);
}
int kvb_vmm_stats(int64_t group_id, int64_t *stats8) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  g->stats(stats8);
  return 0;
}
int kvb_vmm_timing(int64_t group_id, double *out10) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  g->timing(out10);
  return 0;
}
int64_t kvb_vmm_mapped_pages(int64_t group_id, int64_t *out, int64_t cap) {
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->mapped_pages(out, cap);
}
int64_t kvb_vmm_avail_physical_pages(int64_t group_id, int64_t num_layers, int64_t num_kv_buffers,
                                     double gpu_utilization) {
  if (num_layers <= 0 || num_kv_buffers <= 0) return KVB_ERR_INVALID;
  if (gpu_utilization <= 0) {
    const char *e = getenv("KVCACHED_GPU_UTILIZATION");
    gpu_utilization = e && *e ? atof(e) : 0.95;
  }
  if (vmm_initialized()) ensure_device(vmm_device());
  size_t fr = 0, tot = 0;
  CUDA_TRY(cudaMemGetInfo(&fr, &tot));
  double avail = (double)fr - (double)tot * (1.0 - gpu_utilization);
  if (avail < 0) avail = 0;
  int64_t pages = (int64_t)((size_t)avail / (size_t)vmm_page_size() / (size_t)num_layers / (size_t)num_kv_buffers);
  VmmGroup *g = vmm_initialized() ? vmm_group(group_id, false) : nullptr;
  if (g) pages += g->pooled_pages();
  return pages;
}
int kvb_vmm_raw_control(int64_t group_id, int64_t n_pages, int64_t iters, int compound, double *map_us,
                        double *unmap_us) {
  (void)compound;
  VmmGroup *g = vmm_group(group_id, false);
  if (!g) { set_error("group %ld does not exist", (long)group_id); return KVB_ERR_STATE; }
  std::lock_guard<std::mutex> lk(g->mu);
  return g->raw_control(n_pages, iters, map_us, unmap_us);
}

#ifdef KVB_TEST_HOOKS   /* include/kvb200_test.h — libkvb200_test.so only */
int kvb_vmm_probe(int device_index, int64_t unit_bytes, int64_t n_units, int64_t iters, double *out8) {
  if (unit_bytes <= 0 || unit_bytes % (2 << 20) || n_units <= 0 || iters <= 0 || !out8) { set_error("vmm_probe: bad argument"); return KVB_ERR_INVALID; }
  return vmm_probe(device_index < 0 ? 0 : device_index, (size_t)unit_bytes, n_units, iters, out8);
}

int kvb_vmm_probe2(int device_index, int64_t unit_bytes, int64_t m_units, int mode, int64_t iters, double *out3) {
  if (unit_bytes <= 0 || unit_bytes % (2 << 20) || m_units < 16 || iters <= 0 || !out3) { set_error("vmm_probe2: bad argument"); return KVB_ERR_INVALID; }
  return vmm_probe2(device_index < 0 ? 0 : device_index, (size_t)unit_bytes, m_units, mode, iters, out3);
}

int kvb_vmm_probe3(int device_index, int64_t unit_bytes, int64_t m_units, int64_t iters, double *out7) {
  if (unit_bytes <= 0 || unit_bytes % (2 << 20) || m_units < 0 || iters <= 0 || !out7) { set_error("vmm_probe3: bad argument"); return KVB_ERR_INVALID; }
  return vmm_probe3(device_index < 0 ? 0 : device_index, (size_t)unit_bytes, m_units, iters, out7);
}
#endif

// ------------------------------------------------------------------------------------------ page helpers
// InternalPage::get_block_range / get_num_blocks / PageAllocator::get_page_id (page_allocator.cpp:87-100, :457-460)
void kvb_page_block_range(int64_t page_id, int64_t page_size, int64_t bms, int64_t *start, int64_t *end) {
  *start = (page_id * page_size + bms - 1) / bms;
  *end = ((page_id + 1) * page_size) / bms;
}
int64_t kvb_page_num_blocks(int64_t page_size, int64_t bms) { return page_size / bms; }
int64_t kvb_page_id_of_block(int64_t block_id, int64_t bms, int64_t page_size) { return block_id * bms / page_size; }

// ------------------------------------------------------------------------------------------ KV scatter / gather
static int check_layout(const kvb_kv_layout *l, int32_t l0, int32_t l1) {
  if (!l || !l->base || l->cell_bytes <= 0 || (l->cell_bytes & 15) || l->block_size <= 0 || l->num_layers <= 0 ||
      (l->num_kv != 1 && l->num_kv != 2) || l0 < 0 || l1 > l->num_layers || l0 > l1 || (l->base & 15) ||
      (l->stride_block & 15) || (l->stride_token & 15) || (l->stride_layer & 15) || (l->stride_kv & 15)) {
    set_error("kv layout: bad geometry (cell bytes and every stride must be multiples of 16, layers in range)");
    return KVB_ERR_INVALID;
  }
  return 0;
}
int kvb_kv_scatter(const kvb_kv_layout *lay, uint64_t slots_dev, int64_t T, uint64_t k_src, uint64_t v_src,
                   int32_t l0, int32_t l1, void *stream) {
  int rc = check_layout(lay, l0, l1);
  if (rc) return rc;
  if (T < 0 || !slots_dev || !k_src || (lay->num_kv == 2 && !v_src)) { set_error("kv_scatter: null buffer"); return KVB_ERR_INVALID; }
  int e = launch_kv_copy(*lay, (const int64_t *)(uintptr_t)slots_dev, T, (void *)(uintptr_t)k_src, (void *)(uintptr_t)v_src,
                         l0, l1, 0, stream);
  g_total_launches.fetch_add(1, std::memory_order_relaxed);
  if (e) { set_error("kv_scatter launch failed: %s", cudaGetErrorString((cudaError_t)e)); return KVB_ERR_CUDA; }
  return 0;
}
int kvb_kv_gather(const kvb_kv_layout *lay, uint64_t slots_dev, int64_t T, uint64_t k_dst, uint64_t v_dst,
                  int32_t l0, int32_t l1, void *stream) {
  int rc = check_layout(lay, l0, l1);
  if (rc) return rc;
  if (T < 0 || !slots_dev || !k_dst || (lay->num_kv == 2 && !v_dst)) { set_error("kv_gather: null buffer"); return KVB_ERR_INVALID; }
  int e = launch_kv_copy(*lay, (const int64_t *)(uintptr_t)slots_dev, T, (void *)(uintptr_t)k_dst, (void *)(uintptr_t)v_dst,
                         l0, l1, 1, stream);
  g_total_launches.fetch_add(1, std::memory_order_relaxed);
  if (e) { set_error("kv_gather launch failed: %s", cudaGetErrorString((cudaError_t)e)); return KVB_ERR_CUDA; }
  return 0;
}

static int kv_host(const kvb_kv_layout *lay, const int64_t *slots, int64_t T, void *k, void *v, int gather) {
  int rc = check_layout(lay, 0, lay ? lay->num_layers : 0);
  if (rc) return rc;
  if (T <= 0) return 0;
  if (!slots || !k || (lay->num_kv == 2 && !v)) { set_error("kv host copy: null buffer"); return KVB_ERR_INVALID; }
  size_t per = (size_t)lay->num_layers * (size_t)T * (size_t)lay->cell_bytes;
  int64_t *d_slots = nullptr;
  unsigned char *d_k = nullptr, *d_v = nullptr;
  cudaStream_t st = nullptr;
  auto cleanup = [&] {
    if (d_slots) cudaFree(d_slots);
    if (d_k) cudaFree(d_k);
    if (d_v) cudaFree(d_v);
  };
#define KV_TRY(call)                                                                           \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      set_error("%s failed: %s", #call, cudaGetErrorString(e_));                               \
      cleanup();                                                                               \
      return KVB_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)
  KV_TRY(cudaMalloc((void **)&d_slots, (size_t)T * 8));
  KV_TRY(cudaMalloc((void **)&d_k, per));
  if (lay->num_kv == 2) KV_TRY(cudaMalloc((void **)&d_v, per));
  KV_TRY(cudaMemcpyAsync(d_slots, slots, (size_t)T * 8, cudaMemcpyHostToDevice, st));
  if (!gather) {
    KV_TRY(cudaMemcpyAsync(d_k, k, per, cudaMemcpyHostToDevice, st));
    if (d_v) KV_TRY(cudaMemcpyAsync(d_v, v, per, cudaMemcpyHostToDevice, st));
  }
  int e = launch_kv_copy(*lay, d_slots, T, d_k, d_v, 0, lay->num_layers, gather, st);
  g_total_launches.fetch_add(1, std::memory_order_relaxed);
  if (e) { set_error("kv copy launch failed: %s", cudaGetErrorString((cudaError_t)e)); cleanup(); return KVB_ERR_CUDA; }
  if (gather) {
    KV_TRY(cudaMemcpyAsync(k, d_k, per, cudaMemcpyDeviceToHost, st));
    if (d_v) KV_TRY(cudaMemcpyAsync(v, d_v, per, cudaMemcpyDeviceToHost, st));
  }
  KV_TRY(cudaStreamSynchronize(st));
#undef KV_TRY
  cleanup();
  return 0;
}
int kvb_kv_scatter_host(const kvb_kv_layout *lay, const int64_t *slots, int64_t T, const void *k_src,
                        const void *v_src) {
  return kv_host(lay, slots, T, (void *)k_src, (void *)v_src, 0);
}
int kvb_kv_gather_host(const kvb_kv_layout *lay, const int64_t *slots, int64_t T, void *k_dst, void *v_dst) {
  return kv_host(lay, slots, T, k_dst, v_dst, 1);
}

}  // extern "C"
