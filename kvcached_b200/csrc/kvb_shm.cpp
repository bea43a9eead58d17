// kvb_shm.cpp — see kvb_shm.h.
#include "kvb_shm.h"

#include <fcntl.h>
#include <sys/file.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>

namespace kvb {

static constexpr size_t kShmBytes = 3 * sizeof(int64_t);

static std::string shm_path(const std::string &name) {
  if (name.empty()) return "";
  if (name[0] == '/') return name;                     // get_ipc_path, mem_info_tracker.hpp:136-142
  return std::string("/dev/shm/") + name;
}

std::string shm_default_name(int64_t group_id) {
  std::string base;
  const char *e = getenv("KVCACHED_IPC_NAME");         // mem_info_tracker.hpp:235-241
  if (e && e[0]) {
    base = e;
  } else {
    char buf[64];
    snprintf(buf, sizeof(buf), "kvcached_engine_%d", (int)getpgrp());
    base = buf;
  }
  if (group_id != 0) base += "_g" + std::to_string(group_id);   // :155-160
  return base;
}

ShmUsage::ShmUsage(int64_t total, int64_t group_id, const char *ipc_name) {
  name_ = (ipc_name && ipc_name[0]) ? std::string(ipc_name) : shm_default_name(group_id);
  path_ = shm_path(name_);
  fd_ = ::open(path_.c_str(), O_RDWR | O_CREAT, 0666);
  if (fd_ < 0) return;
  struct stat st;
  if (fstat(fd_, &st) != 0 || (size_t)st.st_size < kShmBytes) {
    if (ftruncate(fd_, (off_t)kShmBytes) != 0) { ::close(fd_); fd_ = -1; return; }
  }
  void *p = mmap(nullptr, kShmBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
  if (p == MAP_FAILED) { ::close(fd_); fd_ = -1; return; }
  map_ = (volatile int64_t *)p;
  flock(fd_, LOCK_EX);                                 // init_kv_cache_limit: {limit, 0, 0}
  map_[0] = total;
  map_[1] = 0;
  map_[2] = 0;
  flock(fd_, LOCK_UN);
  last_used_ = last_prealloc_ = 0;
}

ShmUsage::~ShmUsage() {
  if (map_) munmap((void *)map_, kShmBytes);
  if (fd_ >= 0) ::close(fd_);
  if (!path_.empty()) ::unlink(path_.c_str());
}

int ShmUsage::publish(int64_t used, int64_t prealloc) {
  if (!map_) return 0;
  if (used == last_used_ && prealloc == last_prealloc_) return 0;
  flock(fd_, LOCK_EX);                                 // total_size is the limit writer's field: left alone
  map_[1] = used;
  map_[2] = prealloc;
  flock(fd_, LOCK_UN);
  last_used_ = used;
  last_prealloc_ = prealloc;
  return 1;
}

int64_t ShmUsage::resize_target(int64_t cur, int64_t L, int64_t nkv) const {
  if (!map_ || L <= 0 || nkv <= 0) return -1;
  int64_t total = map_[0];                             // aligned 8-byte load: atomic on this platform
  int64_t per_layer = total / L / nkv;
  return per_layer != cur ? per_layer : -1;
}

int shm_read(const char *ipc_name, int64_t out3[3]) {
  std::string path = shm_path(ipc_name ? ipc_name : "");
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) return -1;
  flock(fd, LOCK_SH);
  ssize_t n = pread(fd, out3, kShmBytes, 0);
  flock(fd, LOCK_UN);
  ::close(fd);
  return n == (ssize_t)kShmBytes ? 0 : -1;
}

int shm_set_limit(const char *ipc_name, int64_t total, bool create) {
  std::string path = shm_path(ipc_name ? ipc_name : "");
  int fd = ::open(path.c_str(), create ? (O_RDWR | O_CREAT) : O_RDWR, 0666);
  if (fd < 0) return -1;
  struct stat st;
  bool fresh = fstat(fd, &st) == 0 && (size_t)st.st_size < kShmBytes;
  if (fresh && ftruncate(fd, (off_t)kShmBytes) != 0) { ::close(fd); return -1; }
  flock(fd, LOCK_EX);
  ssize_t n = pwrite(fd, &total, sizeof(total), 0);    // update_kv_cache_limit: only total_size changes
  flock(fd, LOCK_UN);
  ::close(fd);
  return n == (ssize_t)sizeof(total) ? 0 : -1;
}

}  // namespace kvb
