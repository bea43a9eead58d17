// kv_oracle.cpp — CPU restatement of the reference's elastic KV-cache bookkeeping.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; nothing under kvcached_b200/
// links, imports or executes it.  Parity is PINNED: tests/test_oracle_golden.py
// checks this file against tests/golden/*.json.gz, which were produced by replaying
// the same traces through the unmodified reference (oracle/ref_replay.py).
//
// Each function cites the reference code it restates (paths relative to the
// reference root, kvcached 0.1.5).  The restatement keeps the reference's container
// choices where they decide observable order: std::deque for the page lists,
// std::vector for the per-page FIFO, and a real
// std::unordered_map<int64_t, std::vector<int64_t>> for group_indices_by_page, so
// the iteration order is libstdc++'s own rather than a model.
//
// Build: /usr/bin/g++ -O2 -std=c++17 -shared -fPIC -o oracle/libkvoracle.so oracle/kv_oracle.cpp

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <list>
#include <unordered_map>
#include <vector>

namespace {

using i64 = int64_t;

// csrc/page_allocator.cpp:40-100 (InternalPage)
struct Page {
  i64 id = 0, start = 0, end = 0;
  std::vector<i64> free_list;  // alloc from the front, free to the back
  void init(i64 P, i64 B) {    // :44-53, get_block_range :87-95
    start = (id * P + B - 1) / B;
    end = ((id + 1) * P) / B;
    free_list.clear();
    for (i64 i = start; i < end; ++i) free_list.push_back(i);
  }
  bool empty() const { return (i64)free_list.size() == end - start; }  // :75-77
  bool full() const { return free_list.empty(); }                      // :79
};

// Insertion-ordered dict with popitem() (Python dict semantics, kv_cache_manager.py:172-173).
struct OrderedPages {
  std::list<i64> order;
  std::unordered_map<i64, std::pair<std::list<i64>::iterator, Page>> m;
  bool has(i64 k) const { return m.count(k) != 0; }
  size_t size() const { return m.size(); }
  void put(Page &&p) {  // d[k] = v: a new key goes to the end, an existing key keeps its place
    i64 k = p.id;
    auto ex = m.find(k);
    if (ex != m.end()) { ex->second.second = std::move(p); return; }
    order.push_back(k);
    auto it = std::prev(order.end());
    m.emplace(k, std::make_pair(it, std::move(p)));
  }
  Page pop(i64 k) {
    auto it = m.find(k);
    Page p = std::move(it->second.second);
    order.erase(it->second.first);
    m.erase(it);
    return p;
  }
  Page popitem() { return pop(order.back()); }
  void clear() { order.clear(); m.clear(); }
};

struct Event { i64 page; int unmap; };

struct Oracle {
  // config
  i64 P, B, L, nkv, num_blocks;
  i64 min_reserved, max_reserved;
  bool prealloc, reserve_null;
  // csrc/inc/page_allocator.hpp:138-148
  i64 num_free_pages, num_total_pages;
  std::deque<i64> free_list, reserved_list, reclaimed_list;
  bool prealloc_needed = false, prealloc_running = false;
  // kvcached/kv_cache_manager.py:171-179
  i64 num_avail_blocks = 0;
  OrderedPages avail_pages, full_pages;
  std::vector<i64> reserved_blocks;
  bool in_shrink = false;
  i64 target_num_blocks = -1;
  bool has_null = false;
  // environment the reference reads live
  i64 phys_free_pages = (i64)1 << 40;   // get_avail_physical_pages, page_allocator.cpp:442-455
  i64 resize_target = -1;               // get_resize_target, :780
  // observation
  std::vector<char> mapped;
  std::vector<Event> events;
  i64 n_unknown_page_frees = 0;

  i64 bpp() const { return P / B; }  // InternalPage::get_num_blocks :97-100

  // ---- PageAllocator --------------------------------------------------------
  void map_pages(const std::vector<i64> &ids) {   // :619-646
    for (i64 p : ids) {
      if ((size_t)p >= mapped.size()) mapped.resize(p + 1, 0);
      mapped[p] = 1;
      events.push_back({p, 0});
      phys_free_pages -= 1;
    }
  }
  void unmap_pages(const std::vector<i64> &ids) {  // :648-686
    for (i64 p : ids) {
      if ((size_t)p < mapped.size()) mapped[p] = 0;
      events.push_back({p, 1});
      phys_free_pages += 1;
    }
  }
  i64 avail_physical() const { return phys_free_pages < 0 ? 0 : phys_free_pages; }

  // :161-237.  returns -1 on the reference's std::runtime_error.
  i64 alloc_page() {
    if (!reserved_list.empty()) {                  // fast path :169-190
      i64 p = reserved_list.front();
      reserved_list.pop_front();
      num_free_pages--;
      if ((i64)reserved_list.size() < min_reserved) prealloc_needed = true;
      return p;
    }
    if (!free_list.empty()) {                      // slow path :193-216
      i64 p = free_list.front();
      free_list.pop_front();
      num_free_pages--;
      map_pages({p});
      if (prealloc) prealloc_needed = true;        // :226-228
      return p;
    }
    return -1;                                     // :200-207 throws
  }
  void free_pages(const std::vector<i64> &ids) {   // :264-310
    num_free_pages += (i64)ids.size();
    i64 num_to_reserve = max_reserved - (i64)reserved_list.size();
    std::vector<i64> to_unmap;
    if (num_to_reserve > 0) {
      size_t k = std::min<size_t>((size_t)num_to_reserve, ids.size());
      reserved_list.insert(reserved_list.end(), ids.begin(), ids.begin() + k);
      to_unmap.assign(ids.begin() + k, ids.end());
      if (to_unmap.empty()) return;
    } else {
      to_unmap = ids;
    }
    unmap_pages(to_unmap);
    free_list.insert(free_list.end(), to_unmap.begin(), to_unmap.end());
  }
  bool pa_resize(i64 new_mem_size) {               // :312-401
    i64 new_num_pages = new_mem_size / P;
    if (new_num_pages < num_total_pages - num_free_pages) return false;
    if (new_num_pages == num_total_pages) return true;
    if (new_num_pages > num_total_pages) {
      i64 num_to_expand = new_num_pages - num_total_pages;
      i64 num_to_reuse = std::min<i64>((i64)reclaimed_list.size(), num_to_expand);
      for (i64 i = 0; i < num_to_reuse; ++i) {
        free_list.push_back(reclaimed_list.front());
        reclaimed_list.pop_front();
      }
      num_to_expand -= num_to_reuse;
      num_free_pages += num_to_reuse;
      if (num_to_expand > 0) {
        for (i64 i = num_total_pages; i < num_total_pages + num_to_expand; ++i) free_list.push_back(i);
        num_free_pages += num_to_expand;
      }
      num_total_pages = new_num_pages;
      return true;
    }
    i64 num_to_reclaim = num_total_pages - new_num_pages;
    if ((i64)free_list.size() < num_to_reclaim) {
      if (reserved_list.empty()) return false;
      std::vector<i64> to_unmap(reserved_list.begin(), reserved_list.end());
      reserved_list.clear();
      unmap_pages(to_unmap);
      free_list.insert(free_list.end(), to_unmap.begin(), to_unmap.end());
      if ((i64)free_list.size() < num_to_reclaim) return false;
    }
    for (i64 i = 0; i < num_to_reclaim; ++i) {
      reclaimed_list.push_back(free_list.back());
      free_list.pop_back();
    }
    num_free_pages -= num_to_reclaim;
    num_total_pages = new_num_pages;
    return true;
  }
  void pa_trim() {                                 // :403-427
    std::vector<i64> to_unmap(reserved_list.begin(), reserved_list.end());
    reserved_list.clear();
    if (to_unmap.empty()) return;
    unmap_pages(to_unmap);
    free_list.insert(free_list.end(), to_unmap.begin(), to_unmap.end());
  }
  // :536-617, made deterministic: one legal schedule of the reference's prealloc
  // thread, namely "it runs to completion right after the operation that woke it".
  void run_prealloc() {
    if (!prealloc || !prealloc_running || !prealloc_needed) return;
    prealloc_needed = false;
    i64 to_reserve = std::max<i64>(0, min_reserved - (i64)reserved_list.size());
    to_reserve = std::min({to_reserve, (i64)free_list.size(), avail_physical()});
    if (to_reserve <= 0) return;
    std::vector<i64> ids;
    for (i64 i = 0; i < to_reserve && !free_list.empty(); ++i) {
      ids.push_back(free_list.front());
      free_list.pop_front();
    }
    map_pages(ids);
    reserved_list.insert(reserved_list.end(), ids.begin(), ids.end());
  }

  // ---- KVCacheManager -------------------------------------------------------
  i64 available_size() const {                     // kv_cache_manager.py:412-423
    i64 avail = num_avail_blocks + (i64)reserved_blocks.size();
    if (in_shrink) return avail;
    i64 phys = avail_physical() + (i64)reserved_list.size();
    return avail + std::min(num_free_pages, phys) * bpp();
  }
  i64 num_alloced_blocks() const {                 // :493-506
    return (i64)full_pages.size() * bpp() + (i64)avail_pages.size() * bpp() - num_avail_blocks +
           (i64)reserved_blocks.size();
  }
  // :250-304.  rc: >=0 number of ids written; -1 None; -2 RuntimeError (alloc_page threw)
  i64 alloc(i64 need, std::vector<i64> &ret) {
    ret.clear();
    if (resize_target > 0) resize(resize_target);
    if (available_size() < need) return -1;
    i64 remaining = need;
    if (!reserved_blocks.empty()) {
      i64 m = std::min<i64>((i64)reserved_blocks.size(), remaining);
      ret.assign(reserved_blocks.begin(), reserved_blocks.begin() + m);
      reserved_blocks.erase(reserved_blocks.begin(), reserved_blocks.begin() + m);
      remaining -= m;
    }
    while (remaining > 0) {
      Page page;
      if (avail_pages.size() == 0) {
        i64 pid = alloc_page();
        if (pid < 0) return -2;
        page.id = pid;
        page.init(P, B);
        if (page.free_list.empty()) {             // :287-289 zero-block page parked
          full_pages.put(std::move(page));
          continue;
        }
        num_avail_blocks += (i64)page.free_list.size();
      } else {
        page = avail_pages.popitem();             // :292 last inserted
      }
      i64 t = std::min<i64>((i64)page.free_list.size(), remaining);
      ret.insert(ret.end(), page.free_list.begin(), page.free_list.begin() + t);   // page.alloc :55-67
      page.free_list.erase(page.free_list.begin(), page.free_list.begin() + t);
      if (page.full()) full_pages.put(std::move(page));
      else avail_pages.put(std::move(page));
      num_avail_blocks -= t;
      remaining -= t;
    }
    return (i64)ret.size();
  }
  void free(const std::vector<i64> &ids) {         // :307-360
    if (ids.empty()) return;
    // group_indices_by_page, csrc/page_allocator.cpp:471-498
    std::unordered_map<i64, std::vector<i64>> groups;
    groups.reserve(ids.size() / bpp() + 1);
    for (i64 idx : ids) groups[idx * B / P].push_back(idx);

    std::vector<i64> pages_to_free;
    for (auto &kv : groups) {                      // unordered_map iteration order
      i64 pid = kv.first;
      Page page;
      if (full_pages.has(pid)) page = full_pages.pop(pid);
      else if (avail_pages.has(pid)) page = avail_pages.pop(pid);
      else { n_unknown_page_frees++; continue; }   // :329-340 logged and skipped
      num_avail_blocks += (i64)kv.second.size();
      page.free_list.insert(page.free_list.end(), kv.second.begin(), kv.second.end());
      if (page.empty()) {
        pages_to_free.push_back(pid);
        num_avail_blocks -= (i64)page.free_list.size();
      } else {
        avail_pages.put(std::move(page));
      }
    }
    if (!pages_to_free.empty()) free_pages(pages_to_free);
    if (in_shrink && num_alloced_blocks() <= target_num_blocks) {   // :354-360
      pa_resize(target_num_blocks * B);            // result ignored
      in_shrink = false;
      target_num_blocks = -1;
    }
  }
  // :381-401.  rc 1 True, 0 False, -3 AssertionError (reserved blocks present)
  int resize(i64 new_mem_size) {
    if (pa_resize(new_mem_size)) {
      if (in_shrink) { in_shrink = false; target_num_blocks = -1; }
      return 1;
    }
    if (!reserved_blocks.empty()) return -3;
    in_shrink = true;
    target_num_blocks = new_mem_size / B;
    free_reserved();
    return 0;
  }
  int try_to_reserve(i64 need) {                   // :363-372
    if (available_size() < need) return 0;
    std::vector<i64> r;
    i64 rc = alloc(need, r);
    if (rc == -2) return -2;
    if (rc < 0) return 0;
    reserved_blocks.insert(reserved_blocks.end(), r.begin(), r.end());
    return 1;
  }
  void free_reserved() {                           // :375-378
    if (!reserved_blocks.empty()) {
      std::vector<i64> r = reserved_blocks;
      free(r);
      reserved_blocks.clear();
    }
  }
  int reserve_null_block() {                       // :233-243
    if (!reserve_null) { has_null = false; return 0; }
    std::vector<i64> r;
    i64 rc = alloc(1, r);
    if (rc != 1 || r[0] != 0) return -1;
    has_null = true;
    return 0;
  }
  // :444-489 (the reference raises AttributeError at :454; semantics per SURVEY A.7)
  int clear() {
    prealloc_running = false;
    free_reserved();
    std::vector<i64> ids(avail_pages.order.begin(), avail_pages.order.end());
    ids.insert(ids.end(), full_pages.order.begin(), full_pages.order.end());
    if (!ids.empty()) free_pages(ids);
    avail_pages.clear();
    full_pages.clear();
    pa_trim();
    std::sort(free_list.begin(), free_list.end());   // reset_free_page_order :703-709
    target_num_blocks = -1;
    in_shrink = false;
    num_avail_blocks = 0;
    int rc = reserve_null_block();
    start_prealloc();
    return rc;
  }
  void start_prealloc() {                          // :524-528, :717-733
    if (!prealloc) return;
    prealloc_running = true;
    prealloc_needed = true;
    run_prealloc();
  }
};

}  // namespace

extern "C" {

struct kvo_config {
  int64_t num_blocks, block_mem_size, page_size, num_layers, num_kv_buffers;
  int64_t min_reserved_pages, max_reserved_pages;   // env defaults 5 / 10, page_allocator.cpp:24-32
  int32_t reserve_null_block, prealloc;
};

void *kvo_create(const kvo_config *c) {
  if (c->block_mem_size > c->page_size) return nullptr;   // KVCachedConfigError, kv_cache_manager.py:104-116
  auto *o = new Oracle();
  o->P = c->page_size; o->B = c->block_mem_size; o->L = c->num_layers; o->nkv = c->num_kv_buffers;
  o->num_blocks = c->num_blocks;
  i64 mem_size = c->num_blocks * c->block_mem_size;        // :118
  o->num_free_pages = o->num_total_pages = mem_size / o->P;  // page_allocator.cpp:115-116
  o->min_reserved = std::min(o->num_free_pages, (i64)c->min_reserved_pages);   // :117-118
  o->max_reserved = std::min(o->num_free_pages, (i64)c->max_reserved_pages);
  o->prealloc = c->prealloc != 0;
  o->reserve_null = c->reserve_null_block != 0;
  for (i64 i = 0; i < o->num_free_pages; ++i) o->free_list.push_back(i);
  // _post_init, kv_cache_manager.py:190-227
  if (o->reserve_null_block() != 0) { delete o; return nullptr; }
  o->start_prealloc();
  return o;
}
void kvo_destroy(void *h) { delete (Oracle *)h; }

int64_t kvo_alloc(void *h, int64_t need, int64_t *out, int64_t cap) {
  auto *o = (Oracle *)h;
  std::vector<i64> r;
  i64 rc = o->alloc(need, r);
  o->run_prealloc();
  if (rc < 0) return rc;
  if ((i64)r.size() > cap) return -9;
  std::memcpy(out, r.data(), r.size() * sizeof(i64));
  return rc;
}
void kvo_free(void *h, const int64_t *ids, int64_t n) {
  auto *o = (Oracle *)h;
  o->free(std::vector<i64>(ids, ids + n));
  o->run_prealloc();
}
int kvo_resize(void *h, int64_t bytes) { return ((Oracle *)h)->resize(bytes); }
void kvo_trim(void *h) { ((Oracle *)h)->pa_trim(); }
int kvo_try_to_reserve(void *h, int64_t n) {
  auto *o = (Oracle *)h;
  int rc = o->try_to_reserve(n);
  o->run_prealloc();
  return rc;
}
void kvo_free_reserved(void *h) { ((Oracle *)h)->free_reserved(); }
int64_t kvo_available_size(void *h) { return ((Oracle *)h)->available_size(); }
int kvo_clear(void *h) { return ((Oracle *)h)->clear(); }
void kvo_set_phys_free_pages(void *h, int64_t n) { ((Oracle *)h)->phys_free_pages = n; }
void kvo_set_resize_target(void *h, int64_t bytes) { ((Oracle *)h)->resize_target = bytes; }

// scalars: [num_free_pages, num_inuse, num_reserved_pages, num_total_pages, num_avail_blocks,
//           in_shrink, n_avail_pages, n_full_pages, n_reserved_blocks, n_mapped, has_null, target_num_blocks]
void kvo_scalars(void *h, int64_t *s) {
  auto *o = (Oracle *)h;
  s[0] = o->num_free_pages; s[1] = o->num_total_pages - o->num_free_pages;
  s[2] = (i64)o->reserved_list.size(); s[3] = o->num_total_pages; s[4] = o->num_avail_blocks;
  s[5] = o->in_shrink; s[6] = (i64)o->avail_pages.size(); s[7] = (i64)o->full_pages.size();
  s[8] = (i64)o->reserved_blocks.size();
  i64 nm = 0; for (char c : o->mapped) nm += c;
  s[9] = nm; s[10] = o->has_null; s[11] = o->target_num_blocks;
}
static int64_t copy_out(const std::vector<i64> &v, int64_t *out, int64_t cap) {
  if ((i64)v.size() > cap) return -9;
  std::memcpy(out, v.data(), v.size() * sizeof(i64));
  return (i64)v.size();
}
int64_t kvo_avail_pages(void *h, int64_t *out, int64_t cap) {
  auto *o = (Oracle *)h;
  return copy_out(std::vector<i64>(o->avail_pages.order.begin(), o->avail_pages.order.end()), out, cap);
}
int64_t kvo_full_pages(void *h, int64_t *out, int64_t cap) {
  auto *o = (Oracle *)h;
  std::vector<i64> v(o->full_pages.order.begin(), o->full_pages.order.end());
  std::sort(v.begin(), v.end());
  return copy_out(v, out, cap);
}
int64_t kvo_reserved_blocks(void *h, int64_t *out, int64_t cap) {
  return copy_out(((Oracle *)h)->reserved_blocks, out, cap);
}
int64_t kvo_mapped_pages(void *h, int64_t *out, int64_t cap) {
  auto *o = (Oracle *)h;
  std::vector<i64> v;
  for (size_t i = 0; i < o->mapped.size(); ++i) if (o->mapped[i]) v.push_back((i64)i);
  return copy_out(v, out, cap);
}
// which: 0 free_page_list, 1 reserved_page_list, 2 reclaimed_page_list
int64_t kvo_page_list(void *h, int which, int64_t *out, int64_t cap) {
  auto *o = (Oracle *)h;
  const std::deque<i64> &d = which == 0 ? o->free_list : which == 1 ? o->reserved_list : o->reclaimed_list;
  return copy_out(std::vector<i64>(d.begin(), d.end()), out, cap);
}
int64_t kvo_page_free_blocks(void *h, int64_t pid, int64_t *out, int64_t cap) {
  auto *o = (Oracle *)h;
  if (o->avail_pages.has(pid)) return copy_out(o->avail_pages.m.at(pid).second.free_list, out, cap);
  if (o->full_pages.has(pid)) return 0;
  return -1;
}
// drains the ordered map/unmap event log: out[i] = page*2 + is_unmap
int64_t kvo_drain_events(void *h, int64_t *out, int64_t cap) {
  auto *o = (Oracle *)h;
  if ((i64)o->events.size() > cap) return -9;
  for (size_t i = 0; i < o->events.size(); ++i) out[i] = o->events[i].page * 2 + o->events[i].unmap;
  i64 n = (i64)o->events.size();
  o->events.clear();
  return n;
}

// group_indices_by_page alone (csrc/page_allocator.cpp:471-498): writes the distinct page ids in
// libstdc++ iteration order and the ids regrouped by page (input order inside a page).
int64_t kvo_group_indices_by_page(const int64_t *ids, int64_t n, int64_t block_mem_size, int64_t page_size,
                                  int64_t *pages_out, int64_t *counts_out, int64_t *grouped_out) {
  std::unordered_map<i64, std::vector<i64>> groups;
  i64 bpp = page_size / block_mem_size;
  groups.reserve(n / bpp + 1);
  for (i64 i = 0; i < n; ++i) groups[ids[i] * block_mem_size / page_size].push_back(ids[i]);
  i64 d = 0, w = 0;
  for (auto &kv : groups) {
    pages_out[d] = kv.first;
    counts_out[d] = (i64)kv.second.size();
    for (i64 b : kv.second) grouped_out[w++] = b;
    ++d;
  }
  return d;
}

// InternalPage statics, csrc/page_allocator.cpp:87-100
void kvo_block_range(int64_t page_id, int64_t page_size, int64_t bms, int64_t *start, int64_t *end) {
  *start = (page_id * page_size + bms - 1) / bms;
  *end = ((page_id + 1) * page_size) / bms;
}
int64_t kvo_num_blocks(int64_t page_size, int64_t bms) { return page_size / bms; }

}  // extern "C"
