// vmm_stall4.cu — does the CPU the calling thread runs on matter for the sleeping VMM calls?  Round-robin 0.5 s windows
// over "unpinned" and each online CPU; prints /proc/interrupts nvidia rows before and after.  Measurement tool only.
#include <cuda.h>
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s = nullptr; cuGetErrorString(r_, &s); fprintf(stderr, "%s: %s\n", #x, s ? s : "?"); exit(1); } } while (0)
using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }
static void irqs() {
  FILE *f = fopen("/proc/interrupts", "r"); if (!f) return; char b[4096]; int first = 1;
  while (fgets(b, sizeof b, f)) { if (first || strstr(b, "nvidia") || strstr(b, "vfio") || strstr(b, "MSI")) printf("%s", b); first = 0; }
  fclose(f);
}
int main(int argc, char **argv) {
  int rounds = argc > 1 ? atoi(argv[1]) : 4;
  size_t unit = 128u << 20;
  int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);
  irqs();
  cudaSetDevice(0); cudaFree(0);
  CUmemAllocationProp prop{}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  CUmemAccessDesc acc{}; acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  const int K = 64;
  CUdeviceptr base; CK(cuMemAddressReserve(&base, unit * K, 2u << 20, 0, 0));
  CUmemGenericAllocationHandle zero; CK(cuMemCreate(&zero, unit, &prop, 0));
  std::vector<CUmemGenericAllocationHandle> hs(K);
  for (auto &h : hs) CK(cuMemCreate(&h, unit, &prop, 0));
  for (int k = 0; k < K; ++k) { CK(cuMemMap(base + k * unit, unit, 0, zero, 0)); CK(cuMemSetAccess(base + k * unit, unit, &acc, 1)); }
  int NV = ncpu + 1;
  std::vector<std::vector<double>> wall(NV);
  int k = 0, pass = 0;
  for (int r = 0; r < rounds; ++r)
    for (int v = 0; v < NV; ++v) {
      cpu_set_t set; CPU_ZERO(&set);
      if (v == 0) for (int c = 0; c < ncpu; ++c) CPU_SET(c, &set); else CPU_SET(v - 1, &set);
      sched_setaffinity(0, sizeof set, &set);
      auto W0 = clk::now();
      while (us(W0, clk::now()) < 5e5) {
        CUdeviceptr a = base + (size_t)k * unit;
        auto t0 = clk::now();
        CK(cuMemUnmap(a, unit));
        CK(cuMemMap(a, unit, 0, pass ? zero : hs[k], 0));
        CK(cuMemSetAccess(a, unit, &acc, 1));
        wall[v].push_back(us(t0, clk::now()));
        if (++k == K) { k = 0; pass ^= 1; }
      }
    }
  for (int v = 0; v < NV; ++v) {
    auto w = wall[v]; std::sort(w.begin(), w.end()); double s = 0; for (double x : w) s += x;
    char name[32]; if (v == 0) snprintf(name, sizeof name, "unpinned"); else snprintf(name, sizeof name, "cpu %d", v - 1);
    printf("%-10s n=%6zu mean=%8.1f p50=%7.1f p90=%8.1f p99=%9.1f max=%9.1f us per page event\n", name, w.size(), s / w.size(), w[w.size() / 2], w[w.size() * 9 / 10], w[w.size() * 99 / 100], w.back());
  }
  irqs();
  return 0;
}
