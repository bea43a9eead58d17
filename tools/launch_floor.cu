// launch_floor.cu — what one launch + pinned-flag round trip costs on this box, for the single-request path:
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/launch_floor tools/launch_floor.cu && tools/launch_floor
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>

struct Big { long long pad[300]; };   // ~2.4 KB of kernel parameters, like LaunchArgs with inline ids

__global__ void k_flag(volatile long long *flag, long long v) {
  if (threadIdx.x == 0) { __threadfence_system(); *flag = v; }
}
__global__ void k_flag_big(Big b, volatile long long *flag, long long v) {
  if (threadIdx.x == 0) { __threadfence_system(); *flag = v + (b.pad[7] & 0); }
}
__global__ void k_read_flag(const long long *in, volatile long long *flag, long long v) {
  extern __shared__ long long sm[];
  if (threadIdx.x == 0) { sm[0] = in[0]; }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence_system(); *flag = v + (sm[0] & 0); }
}
__global__ void k_read2_flag(const long long *in, volatile long long *flag, long long v) {
  extern __shared__ long long sm[];
  if (threadIdx.x == 0) { long long a = in[0]; sm[0] = in[8 + (a & 7)]; }   // two dependent PCIe reads
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence_system(); *flag = v + (sm[0] & 0); }
}

static double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main() {
  cudaSetDevice(0);
  cudaFree(0);
  long long *h = nullptr;
  cudaHostAlloc((void **)&h, 4096, cudaHostAllocDefault);
  memset(h, 0, 4096);
  volatile long long *flag = h + 64;
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaFuncSetAttribute(k_read_flag, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  cudaFuncSetAttribute(k_read2_flag, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  const int N = 5000;
  Big big{};
  auto run = [&](const char *name, auto launch) {
    for (int i = 0; i < 200; ++i) { *flag = 0; launch((long long)i + 1); while (*flag != i + 1) {} }
    double t0 = now_us();
    for (int i = 0; i < N; ++i) { *flag = 0; launch((long long)i + 1); while (*flag != i + 1) {} }
    double rt = (now_us() - t0) / N;
    cudaStreamSynchronize(st);
    t0 = now_us();
    for (int i = 0; i < N; ++i) launch((long long)i + 1);
    double host = (now_us() - t0) / N;
    cudaStreamSynchronize(st);
    double dev = (now_us() - t0) / N;
    printf("%-44s round trip %6.2f us   launch call %5.2f us   back-to-back %5.2f us/launch\n", name, rt, host, dev);
  };
  run("flag only, 32 threads", [&](long long v) { k_flag<<<1, 32, 0, st>>>(flag, v); });
  run("flag only, 512 threads", [&](long long v) { k_flag<<<1, 512, 0, st>>>(flag, v); });
  run("flag only, 32 threads, 2.4 KB params", [&](long long v) { k_flag_big<<<1, 32, 0, st>>>(big, flag, v); });
  run("1 PCIe read + flag, 512 thr, 0 smem", [&](long long v) { k_read_flag<<<1, 512, 64, st>>>(h, flag, v); });
  run("1 PCIe read + flag, 512 thr, 56 KB smem", [&](long long v) { k_read_flag<<<1, 512, 56 * 1024, st>>>(h, flag, v); });
  run("2 dependent PCIe reads + flag, 512 thr, 56 KB", [&](long long v) { k_read2_flag<<<1, 512, 56 * 1024, st>>>(h, flag, v); });
  // sync-based completion instead of polling a flag
  {
    double t0 = now_us();
    for (int i = 0; i < N; ++i) { k_flag<<<1, 32, 0, st>>>(flag, (long long)i + 1); cudaStreamSynchronize(st); }
    printf("%-44s round trip %6.2f us\n", "flag only + cudaStreamSynchronize", (now_us() - t0) / N);
  }
  return 0;
}
