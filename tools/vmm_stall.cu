// vmm_stall.cu — where do the 100 ms / 1 s stalls of cuMemUnmap / cuMemSetAccess come from?
// Cycles zero<->handle remaps for a fixed wall time; for every call records wall and THREAD CPU time (a call that
// sleeps in the kernel has cpu << wall; one that spins has cpu ~ wall), while a watchdog thread measures its own
// scheduling gaps (CPU starvation / cgroup throttling shows there too).  Measurement tool only.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/vmm_stall tools/vmm_stall.cu -lcuda
// usage: vmm_stall <seconds> <held_handles> <unit_mib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <time.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <algorithm>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s = nullptr; cuGetErrorString(r_, &s); fprintf(stderr, "%s: %s\n", #x, s ? s : "?"); exit(1); } } while (0)
using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }
static double cpu_us() { timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3; }
static void cat(const char *path) {
  FILE *f = fopen(path, "r"); if (!f) { printf("  %s: n/a\n", path); return; }
  char b[512]; printf("  %s:", path); while (fgets(b, sizeof b, f)) { for (char *c = b; *c; ++c) if (*c == '\n') *c = ' '; printf(" %s|", b); } printf("\n"); fclose(f);
}
int main(int argc, char **argv) {
  double secs = argc > 1 ? atof(argv[1]) : 15;
  int held = argc > 2 ? atoi(argv[2]) : 64;
  size_t unit = (size_t)(argc > 3 ? atoi(argv[3]) : 128) << 20;
  cat("/sys/fs/cgroup/cpu.max"); cat("/sys/fs/cgroup/cpu.stat"); cat("/proc/loadavg");
  cudaSetDevice(0); cudaFree(0);
  CUmemAllocationProp prop{}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  CUmemAccessDesc acc{}; acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  const int K = 64;
  CUdeviceptr base; CK(cuMemAddressReserve(&base, unit * K, 2u << 20, 0, 0));
  CUmemGenericAllocationHandle zero; CK(cuMemCreate(&zero, unit, &prop, 0));
  std::vector<CUmemGenericAllocationHandle> hs((size_t)std::max(held, K));
  for (auto &h : hs) CK(cuMemCreate(&h, unit, &prop, 0));
  for (int k = 0; k < K; ++k) { CK(cuMemMap(base + k * unit, unit, 0, zero, 0)); CK(cuMemSetAccess(base + k * unit, unit, &acc, 1)); }
  std::atomic<bool> stop{false};
  std::atomic<double> gap_max{0}; std::atomic<long> gaps_1ms{0}, gaps_10ms{0};
  std::thread wd([&] {
    auto last = clk::now();
    while (!stop.load(std::memory_order_relaxed)) {
      auto now = clk::now(); double g = us(last, now); last = now;
      if (g > 1000) gaps_1ms++;
      if (g > 10000) gaps_10ms++;
      if (g > gap_max.load()) gap_max = g;
    }
  });
  struct Rec { int kind; double wall, cpu, at; };
  std::vector<Rec> recs; recs.reserve(1 << 20);
  auto T0 = clk::now();
  auto timed = [&](int kind, auto &&fn) { double c0 = cpu_us(); auto t0 = clk::now(); fn(); auto t1 = clk::now(); recs.push_back({kind, us(t0, t1), cpu_us() - c0, us(T0, t0) / 1e6}); };
  long cycles = 0;
  while (us(T0, clk::now()) < secs * 1e6) {
    for (int k = 0; k < K; ++k) {
      CUdeviceptr a = base + k * unit;
      timed(0, [&] { CK(cuMemUnmap(a, unit)); });
      timed(1, [&] { CK(cuMemMap(a, unit, 0, hs[k], 0)); });
      timed(2, [&] { CK(cuMemSetAccess(a, unit, &acc, 1)); });
    }
    for (int k = 0; k < K; ++k) {
      CUdeviceptr a = base + k * unit;
      timed(0, [&] { CK(cuMemUnmap(a, unit)); });
      timed(1, [&] { CK(cuMemMap(a, unit, 0, zero, 0)); });
      timed(2, [&] { CK(cuMemSetAccess(a, unit, &acc, 1)); });
    }
    cycles++;
  }
  stop = true; wd.join();
  const char *names[3] = {"cuMemUnmap", "cuMemMap", "SetAccess"};
  printf("held=%d handles x %zu MiB, %ld cycles of %d remaps in %.1f s; watchdog: max gap %.0f us, gaps>1ms %ld, >10ms %ld\n",
         held, unit >> 20, cycles, 2 * K, secs, gap_max.load(), gaps_1ms.load(), gaps_10ms.load());
  for (int kind = 0; kind < 3; ++kind) {
    std::vector<double> w; double tw = 0, tc = 0, slow_w = 0, slow_c = 0; long nslow = 0;
    for (auto &r : recs) if (r.kind == kind) { w.push_back(r.wall); tw += r.wall; tc += r.cpu; if (r.wall > 2000) { nslow++; slow_w += r.wall; slow_c += r.cpu; } }
    std::sort(w.begin(), w.end());
    printf("  %-10s n=%zu p50=%.1f p90=%.1f p99=%.1f max=%.1f us | wall %.1f ms cpu %.1f ms | >2ms: n=%ld wall %.1f ms cpu %.1f ms\n", names[kind], w.size(),
           w[w.size() / 2], w[w.size() * 9 / 10], w[w.size() * 99 / 100], w.back(), tw / 1e3, tc / 1e3, nslow, slow_w / 1e3, slow_c / 1e3);
  }
  printf("  slowest 12 calls (kind wall_us cpu_us at_s):");
  std::vector<Rec> s = recs; std::sort(s.begin(), s.end(), [](const Rec &a, const Rec &b) { return a.wall > b.wall; });
  for (int i = 0; i < 12 && i < (int)s.size(); ++i) printf(" [%s %.0f %.0f @%.2f]", names[s[i].kind], s[i].wall, s[i].cpu, s[i].at);
  printf("\n  per-second wall in driver (ms):");
  std::vector<double> per((size_t)secs + 2, 0.0); for (auto &r : recs) per[(size_t)r.at] += r.wall / 1e3;
  for (double x : per) printf(" %.0f", x);
  printf("\n");
  cat("/sys/fs/cgroup/cpu.stat"); cat("/proc/loadavg");
  return 0;
}
