// kvb_shm.h — memory-usage publishing for kvctl / kvtop and the limit they write back.
// Same wire format as the reference (csrc/inc/mem_info_tracker.hpp:25-36, kvcached/cli/utils.py:29-52):
// file /dev/shm/<ipc_name>[_g<group>], three little-endian int64 {total_size, used_size, prealloc_size},
// writers under flock(LOCK_EX), readers under flock(LOCK_SH).  What differs: the segment is opened and mapped
// ONCE and published once per scheduler step if it changed — the reference opens, locks, maps, unmaps and closes
// it on every page transition (mem_info_tracker.hpp:176-187: ~7 syscalls per alloc_page / free_page) — and the
// limit is read straight from the mapping each step instead of by a 100 ms polling thread
// (page_allocator.cpp:764-782).
#pragma once

#include <cstdint>
#include <string>

namespace kvb {

class ShmUsage {
 public:
  // MemInfoTracker(total, group_id, ipc_name): creates the segment and writes {total, 0, 0}
  // (mem_info_tracker.hpp:150-166, :216-226).  ok() is false if the segment could not be created.
  ShmUsage(int64_t total_mem_size, int64_t group_id, const char *ipc_name);
  ~ShmUsage();   // unlinks, like MemInfoTracker::cleanup (:228-232)
  bool ok() const { return map_ != nullptr; }
  const std::string &name() const { return name_; }
  // update_memory_usage (:171-183); skipped when nothing changed.  Returns 1 if written.
  int publish(int64_t used, int64_t prealloc);
  // check_and_get_resize_target (:187-200): total / num_layers / num_kv_buffers if it differs from
  // current_mem_size, else -1.  A plain 8-byte load from the mapping.
  int64_t resize_target(int64_t current_mem_size, int64_t num_layers, int64_t num_kv_buffers) const;

 private:
  std::string name_, path_;
  int fd_ = -1;
  volatile int64_t *map_ = nullptr;
  int64_t last_used_ = -1, last_prealloc_ = -1;
};

std::string shm_default_name(int64_t group_id);   // KVCACHED_IPC_NAME or kvcached_engine_<pgid>, + _g<group>
// kvctl-side helpers (kvcached/cli/utils.py:104-153): 0 on success, -1 if the segment does not exist
int shm_read(const char *ipc_name, int64_t out3[3]);
int shm_set_limit(const char *ipc_name, int64_t total, bool create);

}  // namespace kvb
