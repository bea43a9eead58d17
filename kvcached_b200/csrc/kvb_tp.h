// kvb_tp.h — rank-consistent page mapping under tensor parallelism without a per-step collective.
//
// The reference ships every map / unmap list from the scheduler to each TP worker over its own AF_UNIX
// connection (connect + pickle + reply per rank per operation, kvcached/tp_ipc_util.py:96-267).  Here rank 0's
// bookkeeping kernel PUSHES the page list of a step into a mailbox every rank of the node has mapped: a ring of
// fixed-size records in a POSIX shared-memory segment that rank 0 registers with CUDA as mapped pinned host
// memory.  The kernel's publish phase writes [events, hints] into the next slot with plain stores over PCIe, a
// system-scope fence, then the slot's sequence word; every follower rank runs one thread that polls that word,
// feeds the events to its own VMM engine (same wait-for-maps rule, same map-ahead hints) and acknowledges by
// storing the record number into its `applied` word.  Rank 0 waits for the acknowledgements only when the record
// contains a MAP (ids handed out by the step must be backed on every rank before they are used — the reference
// waits for every reply, tp_ipc_util.py:160-170); steps without page events publish nothing and cost nothing.
//
// Records longer than one slot (clear / trim / resize of a big pool) and boxes where CUDA cannot register the
// segment are relayed by rank 0's host thread into the same ring, in chunks, with back-pressure — the
// consumer side is identical.
#pragma once

#include <atomic>
#include <cstdint>
#include <string>
#include <thread>

namespace kvb {

constexpr int kTpMaxRanks = 64;
constexpr int kTpSlotEvents = 256;                 // = EV_INLINE of the step kernel
constexpr int kTpSlotHdr = 8;                      // seq, mgr_seq, n_events, more, n_hints, flags, 2 spare
constexpr int kTpSlotHints = 16;
constexpr int kTpSlotWords = kTpSlotHdr + kTpSlotHints + kTpSlotEvents;   // 280 int64 = 2240 B
constexpr int kTpSlots = 1024;                     // ring capacity (power of two)
constexpr int kTpFlagSettle = 1;                   // trim / clear / resize: retained pages go back on every rank

enum TpSlotWord : int { TW_SEQ = 0, TW_MGR_SEQ, TW_NEV, TW_MORE, TW_NHINTS, TW_FLAGS, TW_SPARE0, TW_SPARE1 };

struct TpHeader {                                  // first 4096 bytes of the segment
  uint64_t magic;
  int64_t version, world, slots, slot_words;
  std::atomic<int64_t> published;                  // records written so far (device or rank 0's host)
  std::atomic<int64_t> stop;                       // rank 0 is going away
  std::atomic<int64_t> device_direct;              // 1 = the kernel writes the ring itself
  std::atomic<int64_t> applied[kTpMaxRanks];       // per rank: records whose maps are in place
  std::atomic<int64_t> attached[kTpMaxRanks];      // per rank: follower thread running
  std::atomic<int64_t> failed[kTpMaxRanks];        // per rank: a VMM call failed (rank 0 surfaces it)
};

// One mapping of the mailbox in this process (producer on rank 0, consumer elsewhere).
class TpMailbox {
 public:
  ~TpMailbox();
  // rank 0: create + (try to) register with CUDA; others: open (waits up to timeout_ms for rank 0), start the thread
  int open(const char *name, int rank, int world, int64_t group_id, bool want_device_direct, double timeout_ms);
  void close();

  // ---- producer (rank 0) ------------------------------------------------------------------------------
  int64_t *device_ring() const { return d_ring_; }            // null = host relay only
  int64_t *device_published() const { return d_published_; }
  int64_t pub() const { return pub_; }
  void note_device_record() { pub_ += 1; n_device_records_++; }
  bool room(int64_t records);                                  // ring has `records` free slots (all followers far enough)
  int wait_room(int64_t records, double timeout_ms);
  // host relay: the list is cut into slot-sized chunks; blocks while the ring is full
  int publish(int64_t mgr_seq, const int64_t *events, int64_t n, const int64_t *hints, int64_t nh, int flags);
  int wait_attached(double timeout_ms);                        // every follower rank of the world runs its thread
  // until every attached follower has applied everything published (or one of them failed)
  int wait_applied(double timeout_ms);

  // ---- consumer (rank > 0) --------------------------------------------------------------------------
  int wait_caught_up(double timeout_ms);                       // this rank applied everything published so far

  void stats(int64_t *out8);
  int rank() const { return rank_; }
  int world() const { return world_; }
  const std::string &name() const { return name_; }

 private:
  void follower_main();
  int64_t min_applied() const;

  std::string name_;
  int rank_ = 0, world_ = 1;
  int64_t group_id_ = 0;
  int fd_ = -1;
  size_t bytes_ = 0;
  unsigned char *base_ = nullptr;
  TpHeader *hdr_ = nullptr;
  int64_t *ring_ = nullptr;       // host view
  int64_t *d_ring_ = nullptr;     // device view (registered segment)
  int64_t *d_published_ = nullptr;
  bool registered_ = false, creator_ = false;
  int64_t pub_ = 0;               // producer: records published so far
  std::thread thread_;
  std::atomic<bool> stop_{false};
  // counters
  std::atomic<int64_t> n_device_records_{0}, n_relay_records_{0}, n_applied_events_{0}, n_waits_{0};
  std::atomic<int64_t> wait_ns_{0};
};

// process-global registry, keyed by group id
TpMailbox *tp_mailbox(int64_t group_id);
int tp_open(const char *name, int rank, int world, int64_t group_id, bool device_direct, double timeout_ms);
int tp_close(int64_t group_id);

}  // namespace kvb
