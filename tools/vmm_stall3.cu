// vmm_stall3.cu — interleaved A/B of conditions around the sleeping VMM calls (1 s windows, round-robin, so slow
// drifts of the box hit every variant equally).  Measurement tool only.
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s = nullptr; cuGetErrorString(r_, &s); fprintf(stderr, "%s: %s\n", #x, s ? s : "?"); exit(1); } } while (0)
using clk = std::chrono::steady_clock;
static double us(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); }
__global__ void tiny(int *p) { if (threadIdx.x == 0) atomicAdd(p, 1); }
__global__ void spin(int *p, long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} if (threadIdx.x == 0) atomicAdd(p, 1); }
int main(int argc, char **argv) {
  int rounds = argc > 1 ? atoi(argv[1]) : 5;
  size_t unit = 128u << 20;
  cudaSetDevice(0); cudaFree(0);
  int *flag; cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
  cudaStream_t s2; cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  CUmemAllocationProp prop{}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  CUmemAccessDesc acc{}; acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  const int K = 64;
  CUdeviceptr base; CK(cuMemAddressReserve(&base, unit * K, 2u << 20, 0, 0));
  CUmemGenericAllocationHandle zero; CK(cuMemCreate(&zero, unit, &prop, 0));
  std::vector<CUmemGenericAllocationHandle> hs(K);
  for (auto &h : hs) CK(cuMemCreate(&h, unit, &prop, 0));
  for (int k = 0; k < K; ++k) { CK(cuMemMap(base + k * unit, unit, 0, zero, 0)); CK(cuMemSetAccess(base + k * unit, unit, &acc, 1)); }
  const char *names[] = {"plain", "tiny kernel+sync before each remap", "gpu kept busy (async spin kernels)", "ctx SCHED_SPIN", "ctx SCHED_BLOCKING_SYNC", "ctx SCHED_YIELD", "nozero (map+access / unmap only)"};
  const int NV = 7;
  std::vector<std::vector<double>> wall(NV);
  std::vector<double> tot(NV, 0.0);
  int k = 0, pass = 0;
  for (int r = 0; r < rounds; ++r)
    for (int v = 0; v < NV; ++v) {
      if (v == 3) CK(cuCtxSetFlags(CU_CTX_SCHED_SPIN));
      else if (v == 4) CK(cuCtxSetFlags(CU_CTX_SCHED_BLOCKING_SYNC));
      else if (v == 5) CK(cuCtxSetFlags(CU_CTX_SCHED_YIELD));
      else CK(cuCtxSetFlags(CU_CTX_SCHED_AUTO));
      // bring every unit back to the zero page, pass 0
      auto W0 = clk::now();
      while (us(W0, clk::now()) < 1e6) {
        CUdeviceptr a = base + (size_t)k * unit;
        if (v == 1) { tiny<<<1, 32>>>(flag); cudaStreamSynchronize(0); }
        if (v == 2) spin<<<148, 128, 0, s2>>>(flag, 400000);
        auto t0 = clk::now();
        if (v == 6) {
          // nozero cycle: unit goes zero -> hole -> handle -> hole -> zero; time only the nozero-equivalent calls
          CK(cuMemUnmap(a, unit));                  // leave the zero design (untimed)
          t0 = clk::now();
          CK(cuMemMap(a, unit, 0, hs[k], 0)); CK(cuMemSetAccess(a, unit, &acc, 1));   // nozero map  (1 sleeping call)
          CK(cuMemUnmap(a, unit));                                                        // nozero unmap (1 sleeping call)
          auto t1 = clk::now();
          CK(cuMemMap(a, unit, 0, zero, 0)); CK(cuMemSetAccess(a, unit, &acc, 1));     // back (untimed)
          wall[v].push_back(us(t0, t1) / 2); tot[v] += us(t0, t1) / 2;                   // per event
        } else {
          CK(cuMemUnmap(a, unit));
          CK(cuMemMap(a, unit, 0, pass ? zero : hs[k], 0));
          CK(cuMemSetAccess(a, unit, &acc, 1));
          auto t1 = clk::now();
          wall[v].push_back(us(t0, t1)); tot[v] += us(t0, t1);
        }
        if (++k == K) { k = 0; pass ^= 1; }
      }
      if (v == 2) cudaStreamSynchronize(s2);
      if (v == 6 || pass) {   // keep state simple: finish the sweep so every unit is zero-backed again
        while (k != 0 || pass) { CUdeviceptr a = base + (size_t)k * unit; CK(cuMemUnmap(a, unit)); CK(cuMemMap(a, unit, 0, pass ? zero : hs[k], 0)); CK(cuMemSetAccess(a, unit, &acc, 1)); if (++k == K) { k = 0; pass ^= 1; } }
      }
    }
  for (int v = 0; v < NV; ++v) {
    auto w = wall[v]; std::sort(w.begin(), w.end());
    printf("%-40s n=%6zu mean=%8.1f p50=%7.1f p90=%8.1f p99=%9.1f max=%9.1f us per page event\n", names[v], w.size(), tot[v] / w.size(), w[w.size() / 2], w[w.size() * 9 / 10], w[w.size() * 99 / 100], w.back());
  }
  return 0;
}
