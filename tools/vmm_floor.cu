// vmm_floor.cu — what does the CUDA driver charge for the page map/unmap events of the bench trace?
// Replays tools/events_decode4096.txt (one line per scheduler step: n, then n events page*2+is_unmap) against
// a private VA reservation in a process that does nothing else, in several variants, and prints per-call
// statistics.  A measurement tool (not part of the product); build:
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/vmm_floor tools/vmm_floor.cu -lcuda
// usage: vmm_floor <events.txt> <variant> [reps=3] [unit_mib=128]
//   variants: zero | nozero | zero_kernel | zero_ondemand | zero_thr2 | zero_bgkernel | zero_sync | zero_multi<K>
//   (zero_multi<K>: K zero handles, page p is backed by zero handle p % K — how much of the unmap tail is the ALIAS count?)
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    CUresult r_ = (x);                                                                 \
    if (r_ != CUDA_SUCCESS) {                                                          \
      const char *s = nullptr;                                                         \
      cuGetErrorString(r_, &s);                                                        \
      fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, s ? s : "?", __FILE__, __LINE__); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }

__global__ void tiny(int *p) { if (threadIdx.x == 0) atomicAdd(p, 1); }
__global__ void spin(int *p, long long cycles) {
  long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (threadIdx.x == 0) atomicAdd(p, 1);
}

struct Stat {
  std::vector<double> v;
  void add(double x) { v.push_back(x); }
  void print(const char *name) {
    if (v.empty()) return;
    std::vector<double> s = v;
    std::sort(s.begin(), s.end());
    double sum = 0;
    for (double x : s) sum += x;
    printf("    %-10s n=%6zu mean=%8.1f p50=%8.1f p90=%8.1f p99=%9.1f max=%9.1f us  total=%8.1f ms\n", name, s.size(),
           sum / s.size(), s[s.size() / 2], s[s.size() * 9 / 10], s[s.size() * 99 / 100], s.back(), sum / 1e3);
  }
  void clear() { v.clear(); }
};

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage\n"); return 2; }
  std::string variant = argv[2];
  int reps = argc > 3 ? atoi(argv[3]) : 3;
  size_t unit = (size_t)(argc > 4 ? atoi(argv[4]) : 128) << 20;
  std::vector<std::vector<long>> steps;
  {
    FILE *f = fopen(argv[1], "r");
    if (!f) { perror("events"); return 1; }
    long n;
    while (fscanf(f, "%ld", &n) == 1) {
      std::vector<long> s((size_t)n);
      for (long i = 0; i < n; ++i) if (fscanf(f, "%ld", &s[(size_t)i]) != 1) return 1;
      steps.push_back(std::move(s));
    }
    fclose(f);
  }
  long max_page = 0, n_ev = 0;
  for (auto &s : steps) for (long e : s) { max_page = std::max(max_page, e >> 1); n_ev++; }
  const long N = 1430;   // pages of the bench pool
  if (max_page >= N) { fprintf(stderr, "page out of range\n"); return 1; }

  cudaSetDevice(0);
  cudaFree(0);
  int *flag;
  cudaMalloc(&flag, 4);
  cudaMemset(flag, 0, 4);
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = 0;
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = 0;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;

  const bool zero_design = variant.rfind("zero", 0) == 0;
  const bool ondemand = variant == "zero_ondemand";
  const bool kern = variant == "zero_kernel";
  const bool thr2 = variant == "zero_thr2";
  const bool bgk = variant == "zero_bgkernel";
  const bool syncb = variant == "zero_sync";

  auto t_init0 = clk::now();
  CUdeviceptr base = 0;
  CK(cuMemAddressReserve(&base, unit * (size_t)N, 2u << 20, 0x4f0000000000ull, 0));
  CUmemGenericAllocationHandle zero = 0;
  int nzero = 1;
  if (variant.rfind("zero_multi", 0) == 0) nzero = std::max(1, atoi(variant.c_str() + 10));
  std::vector<CUmemGenericAllocationHandle> zeros;
  if (zero_design) {
    for (int k = 0; k < nzero; ++k) { CUmemGenericAllocationHandle z; CK(cuMemCreate(&z, unit, &prop, 0)); zeros.push_back(z); }
    zero = zeros[0];
    for (long p = 0; p < N; ++p) {
      CK(cuMemMap(base + (size_t)p * unit, unit, 0, zeros[(size_t)(p % nzero)], 0));
      CK(cuMemSetAccess(base + (size_t)p * unit, unit, &acc, 1));
    }
  }
  auto t_init1 = clk::now();
  std::vector<CUmemGenericAllocationHandle> pool;
  if (!ondemand) {
    for (int i = 0; i < 880; ++i) {
      CUmemGenericAllocationHandle h;
      CK(cuMemCreate(&h, unit, &prop, 0));
      pool.push_back(h);
    }
  }
  auto t_init2 = clk::now();
  printf("variant=%s unit=%zu MiB events/replay=%ld steps=%zu  init: reserve+zero %.1f ms, pool create %.1f ms (%zu handles)\n",
         variant.c_str(), unit >> 20, n_ev, steps.size(), us(t_init0, t_init1) / 1e3, us(t_init1, t_init2) / 1e3, pool.size());
  std::vector<CUmemGenericAllocationHandle> held((size_t)N, 0);

  std::atomic<bool> stop{false};
  std::thread bg;
  if (bgk) {
    bg = std::thread([&] {
      cudaSetDevice(0);
      cudaStream_t s;
      cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
      while (!stop.load()) {
        spin<<<148, 256, 0, s>>>(flag, 200000);   // ~100 us of GPU work per launch, GPU always busy
        cudaStreamSynchronize(s);
      }
    });
  }

  Stat s_unmap, s_map, s_acc, s_create, s_step, s_unmap_zero;
  auto do_event = [&](long e, Stat &su, Stat &sm, Stat &sa, Stat &sc) {
    long p = e >> 1;
    bool un = e & 1;
    CUdeviceptr a = base + (size_t)p * unit;
    if (!un) {
      CUmemGenericAllocationHandle h;
      if (pool.empty()) {
        auto t0 = clk::now();
        CK(cuMemCreate(&h, unit, &prop, 0));
        sc.add(us(t0, clk::now()));
      } else { h = pool.back(); pool.pop_back(); }
      if (zero_design) {
        auto t0 = clk::now();
        CK(cuMemUnmap(a, unit));
        s_unmap_zero.add(us(t0, clk::now()));
      }
      auto t1 = clk::now();
      CK(cuMemMap(a, unit, 0, h, 0));
      auto t2 = clk::now();
      CK(cuMemSetAccess(a, unit, &acc, 1));
      auto t3 = clk::now();
      sm.add(us(t1, t2));
      sa.add(us(t2, t3));
      held[(size_t)p] = h;
    } else {
      auto t0 = clk::now();
      CK(cuMemUnmap(a, unit));
      auto t1 = clk::now();
      su.add(us(t0, t1));
      if (zero_design) {
        CK(cuMemMap(a, unit, 0, zeros[(size_t)(p % nzero)], 0));
        auto t2 = clk::now();
        CK(cuMemSetAccess(a, unit, &acc, 1));
        auto t3 = clk::now();
        sm.add(us(t1, t2));
        sa.add(us(t2, t3));
      }
      pool.push_back(held[(size_t)p]);
      held[(size_t)p] = 0;
    }
  };

  for (int rep = 0; rep < reps; ++rep) {
    s_unmap.clear(); s_map.clear(); s_acc.clear(); s_create.clear(); s_step.clear(); s_unmap_zero.clear();
    auto T0 = clk::now();
    for (auto &st : steps) {
      if (kern) { tiny<<<1, 32>>>(flag); cudaStreamSynchronize(0); }
      if (st.empty()) continue;
      if (syncb) cudaDeviceSynchronize();
      auto a = clk::now();
      if (thr2 && st.size() >= 2) {
        // events of one step touch distinct pages (checked below) -> two driver threads, private handle lists
        std::vector<long> A(st.begin(), st.begin() + st.size() / 2), B(st.begin() + st.size() / 2, st.end());
        // split the pool so the threads do not share it
        std::vector<CUmemGenericAllocationHandle> poolB;
        size_t needB = 0;
        for (long e : B) if (!(e & 1)) needB++;
        for (size_t i = 0; i < needB && !pool.empty(); ++i) { poolB.push_back(pool.back()); pool.pop_back(); }
        Stat bu, bm, ba, bc;
        std::thread tb([&] {
          cudaSetDevice(0);
          for (long e : B) {
            long p = e >> 1; bool un = e & 1; CUdeviceptr ad = base + (size_t)p * unit;
            auto t0 = clk::now();
            CK(cuMemUnmap(ad, unit));
            auto t1 = clk::now();
            CUmemGenericAllocationHandle h = zero;
            if (!un) { h = poolB.back(); poolB.pop_back(); held[(size_t)p] = h; } else { poolB.push_back(held[(size_t)p]); held[(size_t)p] = 0; }
            CK(cuMemMap(ad, unit, 0, h, 0));
            auto t2 = clk::now();
            CK(cuMemSetAccess(ad, unit, &acc, 1));
            auto t3 = clk::now();
            bu.add(us(t0, t1)); bm.add(us(t1, t2)); ba.add(us(t2, t3));
          }
        });
        for (long e : A) do_event(e, s_unmap, s_map, s_acc, s_create);
        tb.join();
        for (auto h : poolB) pool.push_back(h);
        for (double x : bu.v) s_unmap.add(x);
        for (double x : bm.v) s_map.add(x);
        for (double x : ba.v) s_acc.add(x);
      } else {
        for (long e : st) do_event(e, s_unmap, s_map, s_acc, s_create);
      }
      s_step.add(us(a, clk::now()));
    }
    auto T1 = clk::now();
    printf("  replay %d: %.1f ms total, %.1f us per page event\n", rep, us(T0, T1) / 1e3, us(T0, T1) / n_ev);
    s_create.print("cuMemCreate");
    s_unmap.print("cuMemUnmap");
    s_unmap_zero.print("Unmap(zero)");
    s_map.print("cuMemMap");
    s_acc.print("SetAccess");
    s_step.print("step");
    // back to the initial state: unmap what the trace left mapped (its reserved pages)
    for (long p = 0; p < N; ++p)
      if (held[(size_t)p]) {
        CUdeviceptr a = base + (size_t)p * unit;
        CK(cuMemUnmap(a, unit));
        if (zero_design) { CK(cuMemMap(a, unit, 0, zeros[(size_t)(p % nzero)], 0)); CK(cuMemSetAccess(a, unit, &acc, 1)); }
        pool.push_back(held[(size_t)p]);
        held[(size_t)p] = 0;
      }
    fflush(stdout);
  }
  if (bgk) { stop = true; bg.join(); }
  size_t fr = 0, tot = 0;
  cudaMemGetInfo(&fr, &tot);
  printf("  free %.1f GiB of %.1f GiB at exit\n", fr / 1073741824.0, tot / 1073741824.0);
  return 0;
}
