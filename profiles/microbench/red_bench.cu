// RED.F64 throughput microbenchmark: how does red.global.add.f64 scale with the address pattern
// of the 32 lanes of one instruction?  (profiles/microbench; not part of the library)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red(double* a, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory"); }

// run = number of consecutive doubles covered by adjacent lanes (1 = fully scattered, 32 = one 256-B segment per warp)
__device__ __forceinline__ void st(double* a, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory"); }
template <int OP>
__global__ void k(double* out, size_t n, int run, int iters, uint64_t seed) {
  const int lane = threadIdx.x & 31;
  uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  uint64_t s = seed + warp * 0x9E3779B97F4A7C15ull;
  for (int it = 0; it < iters; ++it) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t h = s;
    // group id within warp: lanes [g*run, (g+1)*run) share a random base
    const int g = lane / run;
    h ^= (uint64_t)g * 0xD1B54A32D192ED03ull;
    h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 31;
    size_t base = (h % (n / 32)) * 32;  // 256-B aligned segment
    if (OP == 0) red(out + base + (lane % run), 1.0); else st(out + base + (lane % run), 1.0);
  }
}

int main() {
  const size_t n = 300ull << 20 >> 3;  // 300 MB of doubles (> L2)
  double* d; cudaMalloc(&d, n * 8); cudaMemset(d, 0, n * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 8, threads = 256, iters = 2000;
  for (int pass = 0; pass < 2; ++pass)
    for (int run : {1, 2, 3, 4, 8, 16, 32}) {
      cudaEventRecord(e0);
      k<0><<<blocks, threads>>>(d, n, run, iters, 1234 + pass);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (pass) printf("run=%2d  %.1f G red/s  (%.3f ms)\n", run, (double)blocks * threads * iters / ms / 1e6, ms);
      cudaEventRecord(e0);
      k<1><<<blocks, threads>>>(d, n, run, iters, 1234 + pass);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      if (pass) printf("run=%2d  %.1f G st/s  (%.3f ms)\n", run, (double)blocks * threads * iters / ms / 1e6, ms);
    }
  // small footprint (L2 resident): 32 MB
  const size_t n2 = 32ull << 20 >> 3;
  for (int run : {1, 4, 32}) {
    cudaEventRecord(e0);
    k<0><<<blocks, threads>>>(d, n2, run, iters, 99);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("L2-resident run=%2d  %.1f G red/s\n", run, (double)blocks * threads * iters / ms / 1e6);
  }
  return 0;
}
