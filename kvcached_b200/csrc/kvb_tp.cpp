// kvb_tp.cpp — see kvb_tp.h.
#include "kvb_tp.h"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "kvb_vmm.h"

namespace kvb {

namespace {
constexpr uint64_t kMagic = 0x6b76623230305450ull;   // "kvb200TP"
constexpr int64_t kVersion = 1;
constexpr int E_INVALID = -1, E_CUDA = -2, E_STATE = -3;

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}
// spin first (a record follows the previous one within microseconds while the scheduler is busy), then nap
struct Backoff {
  int spins = 0;
  void pause() {
    if (++spins < 4000) cpu_relax();
    else std::this_thread::sleep_for(std::chrono::microseconds(spins < 20000 ? 20 : 200));
  }
  void reset() { spins = 0; }
};
size_t segment_bytes() { return 4096 + (size_t)kTpSlots * kTpSlotWords * 8; }
std::string shm_path(const char *name, int64_t group_id) {
  std::string s = "/";
  s += name;
  if (group_id != 0) s += "_g" + std::to_string(group_id);
  return s;
}
}  // namespace

static_assert(sizeof(TpHeader) <= 4096, "mailbox header must fit its page");
static_assert((kTpSlots & (kTpSlots - 1)) == 0, "ring size must be a power of two");

TpMailbox::~TpMailbox() { close(); }

int TpMailbox::open(const char *name, int rank, int world, int64_t group_id, bool want_device_direct, double timeout_ms) {
  if (!name || !*name || rank < 0 || world < 1 || world > kTpMaxRanks || rank >= world) {
    set_error("tp_init: bad argument (name, rank %d, world %d)", rank, world);
    return E_INVALID;
  }
  name_ = shm_path(name, group_id);
  rank_ = rank;
  world_ = world;
  group_id_ = group_id;
  bytes_ = segment_bytes();
  creator_ = rank == 0;
  if (creator_) {
    shm_unlink(name_.c_str());   // a stale segment of a crashed job
    fd_ = shm_open(name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd_ < 0 || ftruncate(fd_, (off_t)bytes_) != 0) {
      set_error("tp_init: cannot create /dev/shm%s: %s", name_.c_str(), strerror(errno));
      return E_STATE;
    }
  } else {
    const double t0 = now_ms();
    while (true) {
      fd_ = shm_open(name_.c_str(), O_RDWR, 0600);
      struct stat st {};
      if (fd_ >= 0 && fstat(fd_, &st) == 0 && (size_t)st.st_size >= bytes_) break;
      if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
      if (now_ms() - t0 > timeout_ms) { set_error("tp_init: mailbox /dev/shm%s did not appear within %.0f ms", name_.c_str(), timeout_ms); return E_STATE; }
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  base_ = (unsigned char *)mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
  if (base_ == MAP_FAILED) { base_ = nullptr; set_error("tp_init: mmap failed: %s", strerror(errno)); return E_STATE; }
  hdr_ = reinterpret_cast<TpHeader *>(base_);
  ring_ = reinterpret_cast<int64_t *>(base_ + 4096);
  if (creator_) {
    memset(base_, 0, bytes_);
    hdr_->version = kVersion;
    hdr_->world = world;
    hdr_->slots = kTpSlots;
    hdr_->slot_words = kTpSlotWords;
    if (want_device_direct) {
      // the kernel writes the ring itself: the segment becomes mapped pinned host memory of this process
      if (cudaHostRegister(base_, bytes_, cudaHostRegisterMapped | cudaHostRegisterPortable) == cudaSuccess) {
        void *d = nullptr;
        if (cudaHostGetDevicePointer(&d, base_, 0) == cudaSuccess && d) {
          registered_ = true;
          d_ring_ = reinterpret_cast<int64_t *>((unsigned char *)d + 4096);
          d_published_ = reinterpret_cast<int64_t *>((unsigned char *)d + offsetof(TpHeader, published));
        } else {
          cudaGetLastError();
          cudaHostUnregister(base_);
        }
      } else {
        cudaGetLastError();
      }
    }
    hdr_->device_direct.store(registered_ ? 1 : 0);
    std::atomic_thread_fence(std::memory_order_release);
    hdr_->magic = kMagic;   // last: followers wait for it
  } else {
    const double t0 = now_ms();
    while (reinterpret_cast<volatile uint64_t &>(hdr_->magic) != kMagic) {
      if (now_ms() - t0 > timeout_ms) { set_error("tp_init: mailbox /dev/shm%s never initialised", name_.c_str()); return E_STATE; }
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (hdr_->version != kVersion || hdr_->world != world || hdr_->slot_words != kTpSlotWords) {
      set_error("tp_init: mailbox /dev/shm%s belongs to another job (version %ld, world %ld)", name_.c_str(), (long)hdr_->version, (long)hdr_->world);
      return E_STATE;
    }
    stop_ = false;
    thread_ = std::thread([this] { follower_main(); });
    const double t1 = now_ms();
    while (!hdr_->attached[rank_].load(std::memory_order_acquire) && now_ms() - t1 < timeout_ms)
      std::this_thread::sleep_for(std::chrono::microseconds(100));
  }
  return 0;
}

void TpMailbox::close() {
  if (!base_) return;
  if (creator_) {
    hdr_->stop.store(1, std::memory_order_release);
  } else {
    stop_ = true;
    if (thread_.joinable()) thread_.join();
    hdr_->attached[rank_].store(0, std::memory_order_release);
  }
  if (registered_) { cudaHostUnregister(base_); cudaGetLastError(); registered_ = false; }
  munmap(base_, bytes_);
  base_ = nullptr;
  hdr_ = nullptr;
  ring_ = d_ring_ = d_published_ = nullptr;
  if (fd_ >= 0) ::close(fd_);
  fd_ = -1;
  if (creator_) shm_unlink(name_.c_str());
}

int64_t TpMailbox::min_applied() const {
  int64_t m = INT64_MAX;
  for (int r = 1; r < world_; ++r)
    if (hdr_->attached[r].load(std::memory_order_acquire)) m = std::min(m, hdr_->applied[r].load(std::memory_order_acquire));
  return m;
}

bool TpMailbox::room(int64_t records) {
  int64_t m = min_applied();
  return m == INT64_MAX ? true : pub_ + records - m <= kTpSlots - 1;
}

int TpMailbox::wait_room(int64_t records, double timeout_ms) {
  if (room(records)) return 0;
  const double t0 = now_ms();
  Backoff b;
  while (!room(records)) {
    if (now_ms() - t0 > timeout_ms) { set_error("tp: followers did not drain the mailbox within %.0f ms", timeout_ms); return E_STATE; }
    b.pause();
  }
  return 0;
}

int TpMailbox::publish(int64_t mgr_seq, const int64_t *events, int64_t n, const int64_t *hints, int64_t nh, int flags) {
  if (!base_ || !creator_) { set_error("tp_publish: not the producer"); return E_STATE; }
  if (n < 0 || nh < 0) return E_INVALID;
  if (nh > kTpSlotHints) nh = kTpSlotHints;
  int64_t done = 0;
  do {
    const int64_t k = std::min<int64_t>(n - done, kTpSlotEvents);
    int rc = wait_room(1, 60000.0);
    if (rc) return rc;
    int64_t *slot = ring_ + (size_t)(pub_ & (kTpSlots - 1)) * kTpSlotWords;
    const bool last = done + k >= n;
    slot[TW_MGR_SEQ] = mgr_seq;
    slot[TW_NEV] = k;
    slot[TW_MORE] = last ? 0 : 1;
    slot[TW_NHINTS] = last ? nh : 0;
    slot[TW_FLAGS] = last ? flags : 0;
    if (last) for (int64_t i = 0; i < nh; ++i) slot[kTpSlotHdr + i] = hints[i];
    if (k) memcpy(slot + kTpSlotHdr + kTpSlotHints, events + done, (size_t)k * 8);
    std::atomic_thread_fence(std::memory_order_release);
    reinterpret_cast<std::atomic<int64_t> *>(slot + TW_SEQ)->store(pub_ + 1, std::memory_order_release);
    pub_ += 1;
    hdr_->published.store(pub_, std::memory_order_release);
    n_relay_records_++;
    done += k;
  } while (done < n);
  return 0;
}

int TpMailbox::wait_attached(double timeout_ms) {
  if (!base_ || !creator_) return 0;
  const double t0 = now_ms();
  while (true) {
    int n = 0;
    for (int r = 1; r < world_; ++r) n += hdr_->attached[r].load(std::memory_order_acquire) ? 1 : 0;
    if (n == world_ - 1) return 0;
    if (now_ms() - t0 > timeout_ms) { set_error("tp_init: %d of %d follower ranks attached within %.0f ms", n, world_ - 1, timeout_ms); return E_STATE; }
    std::this_thread::sleep_for(std::chrono::microseconds(200));
  }
}

int TpMailbox::wait_applied(double timeout_ms) {
  if (!base_ || !creator_) return 0;
  n_waits_++;
  const auto t0 = std::chrono::steady_clock::now();
  Backoff b;
  int rc = 0;
  while (true) {
    bool all = true;
    for (int r = 1; r < world_ && !rc; ++r) {
      if (!hdr_->attached[r].load(std::memory_order_acquire)) continue;
      if (hdr_->failed[r].load(std::memory_order_acquire)) { set_error("tp: rank %d failed to apply a page list", r); rc = E_CUDA; }
      if (hdr_->applied[r].load(std::memory_order_acquire) < pub_) all = false;
    }
    if (rc || all) break;
    if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > timeout_ms) {
      set_error("tp: followers did not acknowledge record %ld within %.0f ms", (long)pub_, timeout_ms);
      rc = E_STATE;
      break;
    }
    b.pause();
  }
  wait_ns_ += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

int TpMailbox::wait_caught_up(double timeout_ms) {
  if (!base_ || creator_) return 0;
  const double t0 = now_ms();
  Backoff b;
  while (hdr_->applied[rank_].load(std::memory_order_acquire) < hdr_->published.load(std::memory_order_acquire)) {
    if (hdr_->failed[rank_].load()) { set_error("tp: this rank failed to apply a page list"); return E_CUDA; }
    if (now_ms() - t0 > timeout_ms) { set_error("tp: rank %d did not catch up within %.0f ms", rank_, timeout_ms); return E_STATE; }
    b.pause();
  }
  return 0;
}

// Follower: one thread per group.  Record r lives in slot r % slots and is complete when its first word reads r + 1.
void TpMailbox::follower_main() {
  int64_t next = 0;   // rank 0 opens the ring with a record naming every page mapped before the mailbox existed
  hdr_->applied[rank_].store(next, std::memory_order_release);
  hdr_->attached[rank_].store(1, std::memory_order_release);
  Backoff b;
  std::vector<int64_t> pending;   // events of a chunked record
  while (!stop_.load(std::memory_order_relaxed)) {
    int64_t *slot = ring_ + (size_t)(next & (kTpSlots - 1)) * kTpSlotWords;
    if (reinterpret_cast<std::atomic<int64_t> *>(slot + TW_SEQ)->load(std::memory_order_acquire) != next + 1) {
      if (hdr_->stop.load(std::memory_order_relaxed)) break;
      b.pause();
      continue;
    }
    b.reset();
    const int64_t n = slot[TW_NEV], more = slot[TW_MORE], nh = slot[TW_NHINTS], flags = slot[TW_FLAGS];
    int rc = 0;
    if (n < 0 || n > kTpSlotEvents || nh < 0 || nh > kTpSlotHints) {
      rc = E_INVALID;
    } else {
      pending.insert(pending.end(), slot + kTpSlotHdr + kTpSlotHints, slot + kTpSlotHdr + kTpSlotHints + n);
      if (!more) {
        VmmGroup *g = vmm_group(group_id_, false);
        if (!g || !g->created()) {
          rc = E_STATE;
        } else {
          std::lock_guard<std::mutex> lk(g->mu);
          if (!pending.empty()) {
            rc = g->enqueue_events(pending.data(), (int64_t)pending.size());
            if (rc >= 0) rc = (flags & kTpFlagSettle) ? g->settle() : (g->async() ? g->wait_maps() : g->fence());
            if (rc > 0) rc = 0;
          }
          if (!rc) {
            if (flags & kTpFlagSettle) g->set_prefetch_hints(nullptr, 0);
            else g->set_prefetch_hints(slot + kTpSlotHdr, nh);
          }
        }
        n_applied_events_ += (int64_t)pending.size();
        pending.clear();
      }
    }
    if (rc < 0) hdr_->failed[rank_].store(1, std::memory_order_release);
    next += 1;
    hdr_->applied[rank_].store(next, std::memory_order_release);
  }
  hdr_->attached[rank_].store(0, std::memory_order_release);
}

void TpMailbox::stats(int64_t *o) {
  o[0] = creator_ ? pub_ : (hdr_ ? hdr_->published.load() : 0);
  o[1] = hdr_ ? (creator_ ? (min_applied() == INT64_MAX ? o[0] : min_applied()) : hdr_->applied[rank_].load()) : 0;
  o[2] = n_applied_events_.load();
  o[3] = registered_ ? 1 : (hdr_ ? hdr_->device_direct.load() : 0);
  o[4] = n_device_records_.load();
  o[5] = n_relay_records_.load();
  o[6] = n_waits_.load();
  o[7] = wait_ns_.load() / 1000;   // microseconds rank 0 waited for acknowledgements
}

// ------------------------------------------------------------------------------------------ registry
namespace {
std::mutex g_tp_mu;
std::map<int64_t, std::unique_ptr<TpMailbox>> g_tp;
}  // namespace

TpMailbox *tp_mailbox(int64_t group_id) {
  std::lock_guard<std::mutex> lk(g_tp_mu);
  auto it = g_tp.find(group_id);
  return it == g_tp.end() ? nullptr : it->second.get();
}

int tp_open(const char *name, int rank, int world, int64_t group_id, bool device_direct, double timeout_ms) {
  auto mb = std::make_unique<TpMailbox>();
  int rc = mb->open(name, rank, world, group_id, device_direct, timeout_ms);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(g_tp_mu);
  g_tp[group_id] = std::move(mb);
  return 0;
}

int tp_close(int64_t group_id) {
  std::unique_ptr<TpMailbox> mb;
  {
    std::lock_guard<std::mutex> lk(g_tp_mu);
    auto it = g_tp.find(group_id);
    if (it == g_tp.end()) return 0;
    mb = std::move(it->second);
    g_tp.erase(it);
  }
  mb.reset();
  return 0;
}

}  // namespace kvb
