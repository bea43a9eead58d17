// vmm_stall2.cu — do the sleeping VMM driver calls overlap across threads, and do range calls amortise?
//   threads T : T threads, each cycling zero<->handle remaps over its own 32 units; aggregate remaps/s
//   range R   : one thread; per group of R contiguous units ONE cuMemUnmap(range), R cuMemMap, ONE cuMemSetAccess(range)
// Measurement tool only.  nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/vmm_stall2 tools/vmm_stall2.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s = nullptr; cuGetErrorString(r_, &s); fprintf(stderr, "%s: %s\n", #x, s ? s : "?"); exit(1); } } while (0)
using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }
int main(int argc, char **argv) {
  const char *mode = argc > 1 ? argv[1] : "threads";
  int X = argc > 2 ? atoi(argv[2]) : 1;
  double secs = argc > 3 ? atof(argv[3]) : 8;
  size_t unit = (size_t)(argc > 4 ? atoi(argv[4]) : 128) << 20;
  cudaSetDevice(0); cudaFree(0);
  CUcontext ctx; CK(cuCtxGetCurrent(&ctx));
  CUmemAllocationProp prop{}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  CUmemAccessDesc acc{}; acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUmemGenericAllocationHandle zero; CK(cuMemCreate(&zero, unit, &prop, 0));
  if (!strcmp(mode, "threads")) {
    const int K = 32, T = X;
    CUdeviceptr base; CK(cuMemAddressReserve(&base, unit * K * T, 2u << 20, 0, 0));
    std::vector<CUmemGenericAllocationHandle> hs((size_t)K * T);
    for (auto &h : hs) CK(cuMemCreate(&h, unit, &prop, 0));
    for (int k = 0; k < K * T; ++k) { CK(cuMemMap(base + k * unit, unit, 0, zero, 0)); CK(cuMemSetAccess(base + k * unit, unit, &acc, 1)); }
    std::atomic<long> remaps{0}; std::atomic<bool> stop{false};
    std::vector<std::thread> th;
    auto T0 = clk::now();
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
      CK(cuCtxSetCurrent(ctx));
      while (!stop.load()) {
        for (int pass = 0; pass < 2; ++pass)
          for (int k = 0; k < K && !stop.load(); ++k) {
            CUdeviceptr a = base + (size_t)(t * K + k) * unit;
            CK(cuMemUnmap(a, unit));
            CK(cuMemMap(a, unit, 0, pass ? zero : hs[(size_t)t * K + k], 0));
            CK(cuMemSetAccess(a, unit, &acc, 1));
            remaps++;
          }
      }
    });
    std::this_thread::sleep_for(std::chrono::duration<double>(secs));
    stop = true; for (auto &t : th) t.join();
    double el = us(T0, clk::now()) / 1e6;
    printf("threads=%d unit=%zu MiB: %ld remaps in %.2f s -> %.0f remaps/s aggregate, %.1f us per remap per thread\n", T, unit >> 20, remaps.load(), el, remaps / el, el * 1e6 * T / remaps);
  } else {
    const int R = X, K = 64;
    CUdeviceptr base; CK(cuMemAddressReserve(&base, unit * K, 2u << 20, 0, 0));
    std::vector<CUmemGenericAllocationHandle> hs(K);
    for (auto &h : hs) CK(cuMemCreate(&h, unit, &prop, 0));
    for (int k = 0; k < K; ++k) { CK(cuMemMap(base + k * unit, unit, 0, zero, 0)); CK(cuMemSetAccess(base + k * unit, unit, &acc, 1)); }
    std::vector<double> tu, ta; long units = 0;
    auto T0 = clk::now();
    while (us(T0, clk::now()) < secs * 1e6) {
      for (int pass = 0; pass < 2; ++pass)
        for (int g = 0; g + R <= K; g += R) {
          CUdeviceptr a = base + (size_t)g * unit;
          auto t0 = clk::now();
          CK(cuMemUnmap(a, unit * R));
          auto t1 = clk::now();
          for (int r = 0; r < R; ++r) CK(cuMemMap(a + r * unit, unit, 0, pass ? zero : hs[g + r], 0));
          auto t2 = clk::now();
          CK(cuMemSetAccess(a, unit * R, &acc, 1));
          auto t3 = clk::now();
          tu.push_back(us(t0, t1)); ta.push_back(us(t2, t3)); units += R;
        }
    }
    double el = us(T0, clk::now());
    auto st = [](std::vector<double> v, const char *n) { std::sort(v.begin(), v.end()); double s = 0; for (double x : v) s += x; printf("   %s per call: n=%zu mean=%.1f p50=%.1f p90=%.1f p99=%.1f us\n", n, v.size(), s / v.size(), v[v.size() / 2], v[v.size() * 9 / 10], v[v.size() * 99 / 100]); };
    printf("range R=%d unit=%zu MiB: %.1f us per unit remap overall\n", R, unit >> 20, el / units);
    st(tu, "cuMemUnmap(range)"); st(ta, "SetAccess(range) ");
  }
  return 0;
}
