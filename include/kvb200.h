/* kvb200.h — C ABI of libkvb200.so, the B200-native elastic KV-cache hot path.
 *
 * This is the drop-in boundary for the reference's `kvcached.vmm_ops` extension
 * (csrc/torch_bindings.cpp:182-258) and for the bookkeeping its Python
 * `KVCacheManager` performs on the host (kvcached/kv_cache_manager.py:58-506).
 * Plain pointers, sizes and integer status codes only: no torch types, no C++
 * types, no exceptions, no abort().  The Python package `kvcached_b200` binds
 * these with ctypes and re-creates the reference's Python names on top
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns an int status unless stated: 0 = KVB_OK, <0 = error;
 *     kvb_last_error() returns the message of the last error on the calling thread.
 *   - "insufficient space" is not an error: alloc-type entry points report it per
 *     request with KVB_ST_NONE (the reference returns None, kv_cache_manager.py:262-266).
 *   - block ids / page ids cross the ABI as int64_t, like the reference's
 *     std::vector<int64_t> / List[int].
 *   - all bookkeeping state lives in device memory; host code only launches the
 *     sm_100a kernels and runs the CUDA driver VMM loop the kernels ask for.
 *   - thread safety: one mutex per manager / per VMM group; every entry point sets
 *     the CUDA device itself, so it may be called from foreign threads.
 */
#ifndef KVB200_H_
#define KVB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVB_OK 0
#define KVB_ERR_INVALID (-1)   /* bad argument / bad configuration            */
#define KVB_ERR_CUDA (-2)      /* a CUDA runtime or driver call failed         */
#define KVB_ERR_STATE (-3)     /* not initialised / tensors not created / ...  */
#define KVB_ERR_NOMEM (-4)     /* host or device allocation failed             */
#define KVB_ERR_OVERFLOW (-5)  /* caller buffer too small                      */

/* per-request status written by the bookkeeping kernel */
#define KVB_ST_OK 1
#define KVB_ST_FALSE 0          /* resize()/try_to_reserve() returned False    */
#define KVB_ST_NONE (-1)        /* alloc() -> None: available_size() < need    */
#define KVB_ST_NO_PAGES (-2)    /* reference raises RuntimeError("No free pages left"), page_allocator.cpp:200-207 */
#define KVB_ST_ASSERT (-3)      /* reference AssertionError: reserved blocks present at resize, kv_cache_manager.py:394 */
#define KVB_ST_NULL_BLOCK (-4)  /* null block could not be reserved at index 0, kv_cache_manager.py:239-243 */
#define KVB_ST_CAPACITY (-5)    /* resize beyond the virtual range created at init (the reference would alias page ids) */

const char *kvb_last_error(void);
/* ABI version of the loaded library (bumped on any signature change). */
int kvb_abi_version(void);
/* kernels this library launched since it was loaded (bookkeeping steps, audits, KV scatter/gather) */
int64_t kvb_total_launches(void);

/* ------------------------------------------------------------------------------------------
 * 1. Virtual KV tensors and the batched VMM engine
 *    replaces csrc/allocator.cpp (FTensorAllocator), csrc/ftensor.cpp (FTensor),
 *    csrc/page.cpp (GPUPage) and the free functions of csrc/torch_bindings.cpp:20-60.
 * ------------------------------------------------------------------------------------------ */

/* init_kvcached(dev_str, page_size, contiguous_layout) — torch_bindings.cpp:20-24, allocator.cpp:72-99.
 * page_size 0 = default 2 MiB; must be a multiple of 2 MiB (the reference abort()s, we return KVB_ERR_INVALID).
 * Re-initialising drops every existing group, like the reference. */
int kvb_init(int device_index, int64_t page_size, int contiguous_layout);
/* shutdown_kvcached() — torch_bindings.cpp:26-29 */
int kvb_shutdown(void);
int kvb_is_initialized(void);
int64_t kvb_page_size(void);
int kvb_contiguous_layout(void);

/* create_kv_tensors(size, dtype_size, dev, num_layers, num_kv_buffers, group_id, unified_pool)
 * — torch_bindings.cpp:31-40, allocator.cpp:121-154, :277-296.
 * `size_per_layer` = bytes of one layer's K and V together (rounded up to the page size).
 * contiguous layout: ONE range of size_per_layer*num_layers bytes whose compound page is
 *   page_size*num_layers*num_kv_buffers; per-layer layout: num_layers ranges.
 * Writes the device base address and byte size of every range (1 or num_layers entries)
 * and returns the number of ranges (>0) or an error (<0).  The whole range is readable and
 * writable at once: unmapped pages alias one shared zero page (ftensor.cpp:160-176). */
int kvb_create_kv_tensors(int64_t group_id, int64_t size_per_layer, int64_t num_layers,
                          int64_t num_kv_buffers, int unified_pool,
                          uint64_t *out_ptrs, int64_t *out_bytes, int cap);
/* Zero-page init.  Eager (default, like ftensor.cpp:160-176): every unit of the range is on the zero page when
 * kvb_create_kv_tensors returns — one cuMemMap + cuMemSetAccess per unit, 0.3 s for the contiguous Llama-8B pool
 * here, ~17 s for its per-layer layout (91 520 units; the reference measures 0.6-2 s on GB10, bench_layout/README.md:75).
 * Background (kvb_vmm_zero_init_mode(1) or KVB200_VMM_ZERO_INIT=background, before create): create backs unit 0 only
 * and returns; the group's VMM thread backs the rest in its idle time (while the engine loads weights), pages the
 * bookkeeping maps meanwhile go straight to their handle.  Unbacked, unmapped units FAULT if touched, so the
 * reference's "stray accesses never fault" guarantee holds from kvb_vmm_wait_init() on.
 * None (mode 2, KVB200_VMM_ZERO_INIT=none): no zero page at all — a map is cuMemMap + cuMemSetAccess, an unmap is
 * cuMemUnmap alone (half the sleeping driver calls of a zero-page cycle); touching an unmapped page faults.  Opt-in,
 * for engines that never touch blocks they do not hold.
 * kvb_vmm_zero_init_mode: -1 = follow the environment, 0 eager, 1 background, 2 none; returns the previous mode.
 * kvb_vmm_init_progress out3 = {units done, units, microseconds spent in init}. */
int kvb_vmm_zero_init_mode(int mode);
int kvb_vmm_wait_init(int64_t group_id);
int kvb_vmm_init_progress(int64_t group_id, int64_t *out3);
/* kv_tensors_created(group_id) — torch_bindings.cpp:42-46; returns 1/0 */
int kvb_kv_tensors_created(int64_t group_id);

/* map_to_kv_tensors / unmap_from_kv_tensors(offsets, group_id) — torch_bindings.cpp:48-60,
 * allocator.cpp:161-257.  offsets are the reference's byte offsets
 * (page_id*page_size*(L*nKV if contiguous else 1), page_allocator.cpp:619-631).
 * One call runs the whole cuMemUnmap/cuMemMap/cuMemSetAccess loop over all layers' ranges with
 * pooled physical handles.  Mapping an already mapped page / unmapping an unmapped one is
 * tolerated (ftensor.cpp:104-107, :124-127).  Returns 1 (True), 0 (tensors not created) or <0. */
int kvb_map_to_kv_tensors(int64_t group_id, const int64_t *offsets, int64_t n);
int kvb_unmap_from_kv_tensors(int64_t group_id, const int64_t *offsets, int64_t n);
/* same, addressed by page id (what the device bookkeeping emits) */
int kvb_map_pages(int64_t group_id, const int64_t *page_ids, int64_t n);
int kvb_unmap_pages(int64_t group_id, const int64_t *page_ids, int64_t n);
/* events[i] = page_id*2 + is_unmap, applied in order (the bookkeeping kernel's event log,
 * and the payload rank 0 broadcasts to the other tensor-parallel ranks). */
int kvb_apply_page_events(int64_t group_id, const int64_t *events, int64_t n);

/* ---- asynchronous VMM engine --------------------------------------------------------------------
 * Events change the WANTED state of a page; the group's single VMM thread reconciles the physical
 * state with it, maps before unmaps; an event that meets its own page's pending opposite cancels it
 * (no driver call).  kvb_mgr_step / kvb_tp_apply enqueue the kernel's list and wait for the maps only;
 * kvb_vmm_fence waits for everything (unmaps included).  Replaces what the reference's prealloc
 * thread does for map latency (page_allocator.cpp:524-617) without touching the bookkeeping.
 * KVB200_VMM_ASYNC=0 (or kvb_vmm_set_async(group, 0)): no thread, every call settles before returning. */
int kvb_vmm_enqueue_page_events(int64_t group_id, const int64_t *events, int64_t n);
int kvb_vmm_wait_maps(int64_t group_id);
int kvb_vmm_fence(int64_t group_id);
int kvb_vmm_set_async(int64_t group_id, int on);
/* Retention (OPT-IN; default KVB200_VMM_RETAIN_MS or 0 = off: the physical mapping follows the bookkeeping, a page
 * unmapped there goes back to the zero page).  With ms > 0 a page the bookkeeping unmapped keeps its physical handle
 * mapped until the allocator has been quiet (no map event) for `ms` milliseconds; mapped again before that it costs
 * nothing.  Retained pages never hold more than KVB200_RETAIN_MAX_FRACTION (0.10) of the device, and go back earlier
 * when a map needs a handle that can neither come from the pool nor be created within the KVCACHED_GPU_UTILIZATION
 * headroom (the page retained last goes first: the free page list is a FIFO), under memory pressure and at settle
 * points (trim / clear / resize / kvctl limit, kvb_vmm_settle).
 * Independently of retention, UNMAPPED handles wait in a pool for the next map instead of cuMemRelease + cuMemCreate
 * per transition (page.cpp:10-25): at most KVB200_POOL_MAX_FRACTION (0.10) of the device, released after
 * KVB200_POOL_IDLE_MS (2000) without a map down to KVB200_POOL_FLOOR (16) handles.  The reference's hysteresis is
 * count-based: max_reserved_pages stay mapped (page_allocator.cpp:264-310) — mirrored exactly in the bookkeeping. */
int kvb_vmm_set_retention_ms(int64_t group_id, double ms);   /* default: KVB200_VMM_RETAIN_MS or 0 = retention off */
/* The reserve retained pages must leave free.  While somebody else's footprint on the device is growing (another
 * process, or this process's own allocations) and for 10 s after it last grew: `fraction` of the device (default
 * KVB200_RETAIN_FREE_FRACTION or 0.20) — the VMM thread hands retained pages back (newest first) until it is whole.
 * While nobody else moves: the KVCACHED_GPU_UTILIZATION headroom plus `quiet_margin` (default
 * KVB200_RETAIN_QUIET_MARGIN or 0.02 of the device; < 0 = keep), which a co-located model that wakes up sees as
 * available at once; taking any of it switches the larger reserve on within one poll (5-50 ms). */
int kvb_vmm_set_retain_free_fraction(int64_t group_id, double fraction, double quiet_margin);
/* Retained pages never hold more than `fraction` of the device (default KVB200_RETAIN_MAX_FRACTION or 0.10): at the
 * cap a page the bookkeeping unmaps goes back to the zero page at once (the newest one — the FIFO free list needs it
 * last).  The reference's count-based equivalent is max_reserved_pages (page_allocator.cpp:264-310). */
int kvb_vmm_set_retain_max_fraction(int64_t group_id, double fraction);
/* Physical footprint against the bookkeeping: out8 = {pages physically mapped now (wanted + mapped ahead + retained +
 * unmap pending), pages the bookkeeping wants mapped, PEAK of the difference since the previous call, pages retained
 * now, retention cap in pages, peak retained since the previous call, pages' worth of unmapped handles in the pool,
 * pool cap in pages}.  With retention off (the default) the difference is map-ahead pages plus unmaps still queued,
 * and map-ahead pages only after kvb_vmm_fence. */
int kvb_vmm_footprint(int64_t group_id, int64_t *out8);
int kvb_vmm_settle(int64_t group_id);
int64_t kvb_vmm_pending_unmaps(int64_t group_id);
/* Map-ahead: the step kernel publishes the next pages its free list will hand out; the VMM thread maps up to `depth`
 * of them in its idle time (wanted maps > hinted maps > unmaps), so the map event that follows costs nothing on
 * the scheduler's path.  Physical state only: the bookkeeping (mapped set, counters, page order) is unchanged —
 * unlike the reference's prealloc thread (page_allocator.cpp:524-617), whose reserved list alters results with
 * timing.  depth 0 = off; default KVB200_VMM_PREFETCH or 5 (the reference's KVCACHED_MIN_RESERVED_PAGES). */
int kvb_vmm_prefetch_config(int64_t group_id, int depth);
int kvb_vmm_set_prefetch_hints(int64_t group_id, const int64_t *pages, int64_t n);   /* what kvb_mgr_step does with the kernel's list */
int64_t kvb_vmm_prefetched_pages(int64_t group_id, int64_t *out, int64_t cap);   /* hinted + mapped, not yet wanted */
/* out8 = {events elided, map-ahead hits, pages mapped ahead now, need_map, forced unmaps pending, need_hinted,
 *         pages retained now, retention hits} */
int kvb_vmm_engine_stats(int64_t group_id, int64_t *out8);

/* physical-handle pool (no cuMemCreate / cuMemRelease on the hot path; page.cpp:10-25 creates and
 * releases one handle per map).  `max_cached` handles are kept after unmap, the rest released;
 * kvb_vmm_pool_prefill creates up to n handles ahead of demand. */
int kvb_vmm_pool_config(int64_t group_id, int64_t max_cached);
int kvb_vmm_pool_prefill(int64_t group_id, int64_t n);
/* stats[0..7] = mapped pages, pooled handles, handles created, handles released, map calls,
 * unmap calls, bytes mapped, compound page bytes */
int kvb_vmm_stats(int64_t group_id, int64_t *stats8);
/* wall time attribution of the driver loop: out10 = microseconds spent in {cuMemCreate, cuMemRelease, cuMemUnmap,
 * cuMemMap, cuMemSetAccess} followed by the five call counts */
int kvb_vmm_timing(int64_t group_id, double *out10);
/* sorted ids of the currently mapped pages; returns the count or KVB_ERR_OVERFLOW */
int64_t kvb_vmm_mapped_pages(int64_t group_id, int64_t *out, int64_t cap);
/* get_avail_physical_pages() arithmetic — page_allocator.cpp:442-455:
 * max(free - total*(1-gpu_utilization), 0) / page_size / num_layers / num_kv_buffers, plus the
 * handles parked in the pool (their memory is already ours). */
int64_t kvb_vmm_avail_physical_pages(int64_t group_id, int64_t num_layers, int64_t num_kv_buffers,
                                     double gpu_utilization);
/* raw-driver control loop for the map/unmap latency metric (cf. benchmarks/bench_vmm/bench_vmm.cpp:139-228):
 * maps and unmaps `n_pages` pre-created handles `iters` times over a private range and writes
 * per-iteration microseconds.  compound=1 uses the group's compound page size, else page_size. */
int kvb_vmm_raw_control(int64_t group_id, int64_t n_pages, int64_t iters, int compound,
                        double *map_us, double *unmap_us);

/* ------------------------------------------------------------------------------------------
 * 2. Device-resident block manager
 *    replaces the host bookkeeping of kvcached/kv_cache_manager.py:58-506 and
 *    csrc/page_allocator.cpp:40-100, :103-460, :471-498, :619-709.
 * ------------------------------------------------------------------------------------------ */
typedef struct kvb_mgr kvb_mgr;

typedef struct kvb_mgr_config {
  int64_t num_blocks;          /* KVCacheManager(num_blocks, ...)           kv_cache_manager.py:60 */
  int64_t block_size;          /* tokens per block                                              */
  int64_t cell_size;           /* bytes of one token's K (or V) in one layer                    */
  int64_t num_layers;
  int64_t num_kv_buffers;      /* 2 = K+V, 1 = MLA                                              */
  int64_t page_size;           /* 0 = kvb_page_size() / 2 MiB                                   */
  int64_t group_id;
  int64_t min_reserved_pages;  /* <0 = env KVCACHED_MIN_RESERVED_PAGES or 5  page_allocator.cpp:24-37 */
  int64_t max_reserved_pages;  /* <0 = env KVCACHED_MAX_RESERVED_PAGES or 10 */
  int32_t reserve_null_block;
  int32_t async_sched;         /* synchronise the device before unmapping    page_allocator.cpp:671-673 */
  int32_t contiguous_layout;   /* <0 = kvb_contiguous_layout()                                  */
  int32_t attach_vmm;          /* 1 = apply the kernel's map/unmap events to the group's tensors;
                                  0 = bookkeeping only (events are still logged)               */
  int32_t device_index;        /* <0 = device of kvb_init, else cudaGetDevice                   */
  int32_t world_size;          /* >1: events are also exported for the TP broadcast             */
  double gpu_utilization;      /* <=0 = env KVCACHED_GPU_UTILIZATION or 0.95                    */
  int64_t max_cached_blocks;   /* prefix-cache LRU cap: -1 unlimited, 0 evict on free, >0 cap (patches.py:247-257);
                                  only the KVB_OP_PIN_* requests read it                                     */
  int32_t publish_shm;         /* usage segment for kvctl / kvtop (mem_info_tracker.hpp): 1 on, 0 off,
                                  <0 = env KVB200_SHM, else on iff attach_vmm                               */
  int32_t pad_;
  char ipc_name[64];           /* "" = KVCACHED_IPC_NAME or kvcached_engine_<pgid>, "_g<group>" appended for
                                  group != 0 (mem_info_tracker.hpp:150-166, :235-247)                       */
} kvb_mgr_config;

/* one request of a batched step */
typedef struct kvb_op {
  int32_t kind;   /* KVB_OP_* */
  int32_t n;      /* ALLOC/TRY_RESERVE: blocks wanted; FREE_IDS: number of ids; FREE_REF: number of refs */
  int64_t arg;    /* FREE_IDS: offset into the id pool; FREE_REF: offset into the ref pool (in refs);
                     RESIZE: new size in bytes */
} kvb_op;

#define KVB_OP_ALLOC 1          /* KVCacheManager.alloc          kv_cache_manager.py:246-304 */
#define KVB_OP_FREE_IDS 2       /* KVCacheManager.free           kv_cache_manager.py:307-360 */
#define KVB_OP_FREE_REF 3       /* free(concat of earlier results' slices), ids never leave the device */
#define KVB_OP_RESIZE 4         /* KVCacheManager.resize         kv_cache_manager.py:381-401 */
#define KVB_OP_TRIM 5           /* KVCacheManager.trim           kv_cache_manager.py:404-409 */
#define KVB_OP_TRY_RESERVE 6    /* KVCacheManager.try_to_reserve kv_cache_manager.py:363-372 */
#define KVB_OP_FREE_RESERVED 7  /* KVCacheManager.free_reserved  kv_cache_manager.py:375-378 */
#define KVB_OP_AVAIL 8          /* KVCacheManager.available_size kv_cache_manager.py:412-423 */
#define KVB_OP_CLEAR 9          /* KVCacheManager.clear          kv_cache_manager.py:444-489 */
#define KVB_OP_ALLOC_PAGE 10    /* PageAllocator.alloc_page      page_allocator.cpp:161-237  */
#define KVB_OP_FREE_PAGES 11    /* PageAllocator.free_pages      page_allocator.cpp:264-310  */
#define KVB_OP_PA_RESIZE 12     /* PageAllocator.resize          page_allocator.cpp:312-401  */
#define KVB_OP_RESET_ORDER 13   /* PageAllocator.reset_free_page_order page_allocator.cpp:703-709 */
/* prefix-cache pinning — ElasticBlockPool, kvcached/integration/vllm/patches.py:308-614.  Reference counts, the
 * "has a cache key" flag and the insertion-ordered LRU of evictable blocks live on the device; the hash -> block
 * dictionary stays with the caller (hashes are host data), which learns evicted ids from the result. */
#define KVB_OP_PIN_NEW 14       /* get_new_blocks(n): evict the oldest evictable blocks if short, alloc n, ref=1   :483-517 */
#define KVB_OP_PIN_TOUCH 15     /* touch(blocks): ref += 1, leave the evictable pool                                :519-533 */
#define KVB_OP_PIN_MARK 16      /* cache_full_blocks: the listed blocks received a cache key                        :396-465 */
#define KVB_OP_PIN_FREE 17      /* free_blocks(ordered): ref -= 1; at 0 keyed blocks join the LRU, others are freed;
                                   then the LRU is cut back to max_cached_blocks                                    :535-568 */
#define KVB_OP_PIN_EVICT 18     /* evict_blocks(ids): drop the keys; evictable ones are freed                       :570-588 */
#define KVB_OP_PIN_RESET 19     /* reset_prefix_cache(): free every evictable block, drop every key                 :590-603 */
/* SGLang paged allocator (kvcached/integration/sglang/patches.py:188-297), issued by kvb_mgr_alloc_extend /
 * kvb_mgr_alloc_decode / kvb_mgr_free_slots: sequence lengths and slot indices are DEVICE arrays */
#define KVB_OP_ALLOC_EXTEND 20
#define KVB_OP_ALLOC_DECODE 21
#define KVB_OP_FREE_SLOTS 22
#define KVB_OP_ALLOC_DEV 23     /* alloc(n) with the ids written to a DEVICE array (SGLang token allocator, patches.py:101-103) */
#define KVB_OP_FREE_DEV 24      /* free(ids) with the ids read, in order, from a DEVICE array            (patches.py:105-113) */

/* a slice of an earlier request's result: ids = result(op)[lo:hi] (Python slice semantics, clipped) */
typedef struct kvb_ref {
  int32_t op;  /* index of the earlier ALLOC request (absolute index in the resident trace, or index
                  inside the current step for kvb_mgr_step) */
  int32_t lo;
  int32_t hi;
} kvb_ref;

/* result of one request */
typedef struct kvb_res {
  int32_t status;  /* KVB_ST_*                                               */
  int32_t n;       /* ids produced (ALLOC) / ids consumed (FREE)              */
  int64_t off;     /* offset of the first id in the id output buffer          */
  int64_t avail;   /* available_size() right after the request                */
  int64_t aux;     /* PIN_NEW / PIN_FREE: number of evicted block ids; they sit in the id output buffer directly
                      before `off` (PIN_NEW) or at `off` (PIN_FREE) so the caller can drop their cache keys     */
} kvb_res;

/* counters of the device state as of the last completed launch */
typedef struct kvb_mgr_stats {
  int64_t num_free_pages, num_inuse_pages, num_reserved_pages, num_total_pages;
  int64_t num_avail_blocks, num_reserved_blocks, in_shrink, target_num_blocks;
  int64_t num_avail_pages, num_full_pages, num_mapped_pages, has_null_block;
  int64_t blocks_per_page, block_mem_size, page_size, mem_size;
  int64_t bad_frees, unknown_page_frees, launches, events;
  int64_t h2d_bytes, d2h_bytes;   /* bytes kvb_mgr_step copied host->device / device->host so far */
  int64_t num_evictable_blocks;   /* len(_evictable_blocks) of the prefix-cache LRU (patches.py:605-606) */
} kvb_mgr_stats;

/* KVCacheManager.__init__ + _post_init — kv_cache_manager.py:60-243.  Returns NULL on error
 * (block larger than page -> KVCachedConfigError, :104-116). */
kvb_mgr *kvb_mgr_create(const kvb_mgr_config *cfg);
int kvb_mgr_destroy(kvb_mgr *m);

/* ---- batched fast path: one scheduler step = one kernel launch ----------------------------
 * Runs `n_ops` requests with exactly the results N sequential reference calls would give.
 * `ids` is the pool FREE_IDS requests index, `refs` the pool FREE_REF requests index (refs name
 * requests of THIS call).  Results: res[i] for every request, block ids of all ALLOC requests
 * concatenated into out_ids (res[i].off/res[i].n locate them).  All buffers are HOST memory;
 * the call copies in, launches, copies out, then applies the kernel's page map/unmap list to
 * the group's virtual tensors (cuMemMap loop) before returning.  Returns the number of ids
 * written to out_ids, or <0. */
int64_t kvb_mgr_step(kvb_mgr *m, const kvb_op *ops, int64_t n_ops, const int64_t *ids, int64_t n_ids,
                     const kvb_ref *refs, int64_t n_refs, kvb_res *res, int64_t *out_ids, int64_t out_cap);

/* ---- single-request entry points (the reference's method set) ---------------------------- */
/* alloc(n): returns n (ids in out), KVB_ST_NONE, KVB_ST_NO_PAGES, or <= KVB_ERR_* - 100 */
int64_t kvb_mgr_alloc(kvb_mgr *m, int64_t n, int64_t *out, int64_t cap);
/* free(ids) returns nothing, so it does not wait for its kernel: stream order keeps the sequence, and the launch's
 * results (counters, page unmaps) are consumed by the next call on this manager — or, when the engine goes idle, by the
 * VMM thread within 50 ms.  KVB200_ASYNC_FREE=0 makes it wait.  kvb_mgr_sync waits for everything launched so far. */
int kvb_mgr_free(kvb_mgr *m, const int64_t *ids, int64_t n);
int kvb_mgr_sync(kvb_mgr *m);
int kvb_mgr_resize(kvb_mgr *m, int64_t new_mem_size);      /* 1 True, 0 False, KVB_ST_ASSERT */
int kvb_mgr_trim(kvb_mgr *m);
int kvb_mgr_try_to_reserve(kvb_mgr *m, int64_t n);         /* 1 / 0 / KVB_ST_NO_PAGES */
int kvb_mgr_free_reserved(kvb_mgr *m);
int kvb_mgr_clear(kvb_mgr *m);
/* available_size(): host arithmetic on the counters the last launch published + live physical
 * memory, no launch (kv_cache_manager.py:412-423, page_allocator.cpp:442-455). */
int64_t kvb_mgr_available_size(kvb_mgr *m);
int kvb_mgr_get_stats(kvb_mgr *m, kvb_mgr_stats *out);
/* parity/test hooks: override the physical-page budget (<0 = live cudaMemGetInfo) */
int kvb_mgr_set_phys_free_pages(kvb_mgr *m, int64_t n);

/* ---- SGLang paged allocator: allocate + expand to token slot indices on the device ---------------------------
 * ElasticPagedTokenToKVPoolAllocator.alloc_extend / alloc_decode / free
 * (kvcached/integration/sglang/patches.py:188-297).  The reference counts the new pages on the host
 * (get_num_new_pages on CPU copies of the lengths), calls kv_cache_manager.alloc, ships the ids back with
 * torch.tensor(list) and expands them with SGLang's Triton kernels alloc_extend_kernel / alloc_decode_kernel
 * (sglang/srt/mem_cache/allocator.py; SGLang is a dependency of the reference, not part of it).  Here ONE launch
 * counts the new pages from the device arrays, runs the block allocation and writes the slot indices:
 *   part 1  the rest of the sequence's last partial page:   last_loc + 1 + t
 *   part 2  whole new pages:                                 block * page_size + t % page_size
 *   part 3  the new partial tail page:                       last new block * page_size + t
 * (decode: one token per sequence: last_loc + 1, or new block * page_size when seq_len % page_size == 1).
 * All array arguments are DEVICE pointers to int64; `stream` is the stream that produced them (NULL = legacy
 * default).  Returns 1, or KVB_ST_NONE when available_size() < new pages (the reference returns None; nothing
 * is written), or <0.  *num_new_pages receives the page count.  page_size = tokens per block. */
int kvb_mgr_alloc_extend(kvb_mgr *m, int64_t bs, uint64_t prefix_lens, uint64_t seq_lens, uint64_t last_loc,
                         int64_t extend_num_tokens, uint64_t out_indices, void *stream, int64_t *num_new_pages);
int kvb_mgr_alloc_decode(kvb_mgr *m, int64_t bs, uint64_t seq_lens, uint64_t last_loc, uint64_t out_indices,
                         void *stream, int64_t *num_new_pages);
/* free(free_index): kv_cache_manager.free(torch.unique(free_index // page_size)) — the unique sorted block list is
 * built on the device (bitmap of touched blocks + ordered popc compaction).  Returns the number of blocks freed. */
int64_t kvb_mgr_free_slots(kvb_mgr *m, uint64_t free_index, int64_t n, void *stream);

/* ElasticTokenToKVPoolAllocator (page_size 1, kvcached/integration/sglang/patches.py:60-118): alloc(need_size) returns a
 * device tensor of ids, free(free_index) takes one — both without the list / torch.tensor(list) round trip of the
 * reference.  kvb_mgr_alloc_device returns n, KVB_ST_NONE or <0; kvb_mgr_free_device frees the ids in array order (the
 * order decides the per-page FIFO, so it is kept). */
int64_t kvb_mgr_alloc_device(kvb_mgr *m, int64_t n, uint64_t out_ids_dev, void *stream);
int kvb_mgr_free_device(kvb_mgr *m, uint64_t ids_dev, int64_t n, void *stream);

/* ---- resident traces: requests and results stay in HBM ------------------------------------
 * Upload a whole trace once; then run ranges of it.  FREE_REF refs name absolute request indices.
 * kvb_mgr_trace_run launches one kernel per step range [step_begin, step_end) given by
 * step_offsets (n_steps+1 request indices) without any host round trip in between; page
 * events accumulate on the device and are applied (if attach_vmm) when the call returns. */
int kvb_mgr_trace_upload(kvb_mgr *m, const kvb_op *ops, int64_t n_ops, const int64_t *ids, int64_t n_ids,
                         const kvb_ref *refs, int64_t n_refs, const int64_t *step_offsets, int64_t n_steps,
                         int64_t out_capacity);
int kvb_mgr_trace_run(kvb_mgr *m, int64_t step_begin, int64_t step_end, int steps_per_launch,
                      float *device_ms, int64_t *n_launches);
int kvb_mgr_trace_results(kvb_mgr *m, kvb_res *res, int64_t n_ops, int64_t *out_ids, int64_t out_cap,
                          int64_t *n_out);
int kvb_mgr_trace_reset_cursor(kvb_mgr *m);

/* ---- state inspection (tests, debugging, the consistency audit) ---------------------------- */
/* which: 0 free_page_list, 1 reserved_page_list, 2 reclaimed_page_list, 3 avail_pages (insertion
 * order), 4 full_pages (ascending), 5 reserved_blocks, 6 mapped pages per the bookkeeping (ascending),
 * 7 evictable blocks in LRU order (oldest first), 8 keyed blocks (ascending), 9 (block, ref_cnt) pairs of every
 * block with ref_cnt != 0 (ascending, flattened) */
int64_t kvb_mgr_dump_list(kvb_mgr *m, int which, int64_t *out, int64_t cap);
int64_t kvb_mgr_page_free_blocks(kvb_mgr *m, int64_t page_id, int64_t *out, int64_t cap);
/* drains the ordered page event log (page*2 + is_unmap) accumulated since the last drain */
int64_t kvb_mgr_drain_events(kvb_mgr *m, int64_t *out, int64_t cap);
/* full-bitmap audit kernel (coalesced uint4 sweep + popc): recounts used blocks per page from the
 * bitmap and checks them against the per-page counters.  out4 = {used blocks, in-use pages,
 * fully-free in-use pages, mismatching pages}.  device_ms (optional) = kernel time. */
int kvb_mgr_audit(kvb_mgr *m, int64_t *out4, float *device_ms);
/* the same sweep over a synthetic bitmap big enough to stream from HBM (the pool's own is L2-resident):
 * out4 = {avg kernel ms, algorithmic bytes per launch = bitmap + 8 B per page, used blocks counted, expected} */
int kvb_audit_stress(int device_index, int64_t num_blocks, int32_t blocks_per_page, int32_t iters, double *out4);
/* tuning aid: cycles per phase [0..15] and mark counts [16..31] of the step kernel in a -DKVB_PROFILE build
 * (all zero in the product build) */
int kvb_mgr_profile(kvb_mgr *m, int64_t *out32, int reset);
/* device pointer + word count of the used-block bitmap (tests, external sweeps) */
int kvb_mgr_bitmap(kvb_mgr *m, uint64_t *dev_ptr, int64_t *n_words32);

/* ---- usage segment: /dev/shm/<ipc_name>, int64 {total_size, used_size, prealloc_size}, flock-protected —
 * the reference's wire format (csrc/inc/mem_info_tracker.hpp:25-36, kvcached/cli/utils.py:29-153), so kvctl and
 * kvtop work unchanged.  The manager publishes used = in-use pages * L * P * nKV and prealloc = reserved pages *
 * L * P * nKV once per step (page_allocator.cpp:688-701) and applies a changed total_size (kvctl limit) as
 * KVCacheManager.resize(total / L / nKV) before its next alloc (kv_cache_manager.py:258-260). */
int kvb_mgr_shm_name(kvb_mgr *m, char *out, int cap);          /* length, 0 = publishing is off            */
int64_t kvb_mgr_shm_publishes(kvb_mgr *m);                     /* segment writes so far                    */
int kvb_shm_read(const char *ipc_name, int64_t *out3);         /* get_kv_cache_limit, cli/utils.py:104-113 */
int kvb_shm_set_limit(const char *ipc_name, int64_t total_bytes, int create);   /* update_kv_cache_limit :116-141 */

/* ---- tensor parallel: rank-consistent page mapping ----------------------------------------
 * Replaces kvcached/tp_ipc_util.py:96-267 (one AF_UNIX connection + pickle round trip per rank per map/unmap call).
 *
 * (1) The mailbox — no per-step collective.  kvb_tp_init(name, rank, world, group, mgr, flags):
 *       rank 0 (mgr = its block manager) creates the POSIX shared-memory ring /dev/shm/<name>[_g<group>], registers
 *         it with CUDA as mapped pinned host memory and waits for the other ranks; from then on the step KERNEL
 *         writes each step's page list (and the map-ahead hints) straight into the next ring slot — plain stores
 *         over PCIe, a system-scope fence, then the slot's sequence word.  Steps without page events publish nothing.
 *       ranks > 0 (mgr = NULL) open the ring and start one thread that polls the sequence word, feeds the list to the
 *         rank's own VMM engine (same wait-for-maps rule as kvb_mgr_step) and acknowledges in the ring.
 *     kvb_mgr_step on rank 0 returns when a step's MAPS are in place on EVERY rank (the reference waits for every
 *     worker's reply, tp_ipc_util.py:160-170); unmaps drain behind it on every rank.  Lists longer than one slot
 *     (256 events: clear / trim / resize of a big pool) and boxes where CUDA cannot register the segment
 *     (flags bit 0, or KVB200_TP_DEVICE_DIRECT=0) are relayed into the same ring by rank 0's host thread, in chunks.
 *     All ranks of a TP group live on one node (that is what the reference's AF_UNIX sockets assume too).
 *     flags bit 1 on rank 0 with mgr = NULL: external producer — the ring is fed through kvb_tp_publish only.
 *     kvb_tp_fence: rank 0 — every rank applied everything published; others — this rank did.
 *     kvb_tp_publish: the producer side for hosts that obtain page lists elsewhere (events[i] = page*2 + is_unmap).
 *     kvb_tp_stats out8 = {records published, records applied (rank 0: slowest rank), events applied by this rank,
 *       device-direct (1/0), records written by the kernel, records relayed by the host, acknowledgement waits,
 *       microseconds spent in those waits}. */
int kvb_tp_init(const char *name, int rank, int world_size, int64_t group_id, kvb_mgr *mgr, int flags);
int kvb_tp_fence(int64_t group_id);
int kvb_tp_publish(int64_t group_id, int64_t seq, const int64_t *events, int64_t n, const int64_t *hints,
                   int64_t n_hints, int flags);
int kvb_tp_stats(int64_t group_id, int64_t *out8);
int kvb_tp_shutdown(int64_t group_id);
int kvb_mgr_tp_detach(kvb_mgr *m);
/* (2) The collective path — the checker of (1), and the transport for groups that span nodes.  After a step, rank 0's
 * pending page events are packed as [seq, n, e0 .. e(n-1), n_hints, hints ..., more] into a buffer of
 * kvb_tp_buffer_words() int64 words (`more`, the last word, = 1 when the list continues in another buffer: lists of
 * any length travel in chunks); the host side broadcasts each buffer with ONE NCCL broadcast (torch.distributed) and
 * every other rank applies it.  kvb_mgr_tp_export / kvb_tp_pack return 1 while more chunks remain, 0 after the last;
 * kvb_tp_apply / kvb_tp_apply_host return the buffer's `more` (maps are waited for when the list is complete). */
int64_t kvb_tp_buffer_words(void);
int kvb_mgr_tp_export(kvb_mgr *m, uint64_t dst_device_ptr, void *stream);
int kvb_mgr_tp_export_host(kvb_mgr *m, int64_t *dst);
int kvb_tp_pack(int64_t seq, const int64_t *events, int64_t n, int64_t *cursor, const int64_t *hints, int64_t n_hints,
                int64_t *buf);
int kvb_tp_apply(int64_t group_id, uint64_t src_device_ptr, void *stream, int64_t *seq_out, int64_t *n_out);
int kvb_tp_apply_host(int64_t group_id, const int64_t *buf, int64_t *seq_out, int64_t *n_out);

/* ------------------------------------------------------------------------------------------
 * 3. Page arithmetic helpers — InternalPage statics, csrc/page_allocator.cpp:87-100, :457-498
 * ------------------------------------------------------------------------------------------ */
void kvb_page_block_range(int64_t page_id, int64_t page_size, int64_t block_mem_size, int64_t *start,
                          int64_t *end);
int64_t kvb_page_num_blocks(int64_t page_size, int64_t block_mem_size);
int64_t kvb_page_id_of_block(int64_t block_id, int64_t block_mem_size, int64_t page_size);
/* group_indices_by_page on the DEVICE (same kernel code path free() uses): writes the distinct
 * pages in the reference's dict order, the number of ids of each, and the ids regrouped by page
 * (input order inside a page).  Returns the number of distinct pages. */
int64_t kvb_group_indices_by_page(kvb_mgr *m, const int64_t *ids, int64_t n, int64_t *pages_out,
                                  int64_t *counts_out, int64_t *grouped_out);

/* ------------------------------------------------------------------------------------------
 * 4. Fused paged KV scatter / gather (sm_100a, 128-bit accesses)
 *    The reference has no such kernel; it defines the address layout the engines' kernels write
 *    through (integration/vllm/interfaces.py:281-298, integration/sglang/interfaces.py:132-173).
 * ------------------------------------------------------------------------------------------ */
typedef struct kvb_kv_layout {
  uint64_t base;           /* device address of layer 0, K                         */
  int64_t stride_block;    /* bytes between consecutive blocks                      */
  int64_t stride_token;    /* bytes between consecutive tokens inside a block       */
  int64_t stride_layer;    /* bytes between layers                                  */
  int64_t stride_kv;       /* bytes between K and V of one layer                    */
  int64_t cell_bytes;      /* contiguous bytes of one token's K (or V) in one layer */
  int32_t block_size;      /* tokens per block                                      */
  int32_t num_layers;
  int32_t num_kv;          /* 2 = K and V, 1 = MLA                                  */
  int32_t pad_;
} kvb_kv_layout;

/* src layout: [layer_end - layer_begin][T][cell_bytes] contiguous per buffer (k_src = K, v_src = V; v may be 0
 * when num_kv == 1): the dense buffers hold exactly the layers of THIS launch, first launched layer first — the
 * fused all-layer call (0, num_layers) passes [num_layers][T][cell], a per-layer call (l, l+1) passes the engine's
 * own [T, H, D] tensor of that layer.  slots[T] are token slots (block = slot / block_size); a negative slot is
 * padding and moves nothing (its dense row is left untouched by the gather).
 * Pointers are DEVICE addresses; `stream` is a cudaStream_t (NULL = legacy default stream). */
int kvb_kv_scatter(const kvb_kv_layout *lay, uint64_t slots_dev, int64_t T, uint64_t k_src, uint64_t v_src,
                   int32_t layer_begin, int32_t layer_end, void *stream);
int kvb_kv_gather(const kvb_kv_layout *lay, uint64_t slots_dev, int64_t T, uint64_t k_dst, uint64_t v_dst,
                  int32_t layer_begin, int32_t layer_end, void *stream);
/* host-buffer convenience used by the end-to-end path: copies slots and K/V from pinned or pageable
 * host memory, scatters, and (gather) copies back. */
int kvb_kv_scatter_host(const kvb_kv_layout *lay, const int64_t *slots, int64_t T, const void *k_src,
                        const void *v_src);
int kvb_kv_gather_host(const kvb_kv_layout *lay, const int64_t *slots, int64_t T, void *k_dst, void *v_dst);

#ifdef __cplusplus
}
#endif
#endif /* KVB200_H_ */
