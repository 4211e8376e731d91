// fp64_peak.cu -- measured FP64 peaks of the device: vector DFMA and tensor-core DMMA.8x8x4 (mma.sync m8n8k4.f64).
// The FP64 roofline denominator for the Q2-hexahedron kernels (cq_dmma.cuh) and the point-by-point kernels.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench/fp64_peak profiles/microbench/fp64_peak.cu
// prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 1.2345e300) out[0] = s;
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double d[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
  if (s == 1.2345e300) out[0] = s;
}

template <class K>
static double time_ms(K launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* out;
  cudaMalloc(&out, 8);
  const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 4096;
  constexpr int ILP = 8;
  const double ms_fma = time_ms([&] { dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double fl_fma = 2.0 * ILP * (double)iters * blocks * threads;
  const double ms_mma = time_ms([&] { dmma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double fl_mma = 2.0 * 8 * 8 * 4 * ILP * (double)iters * blocks * (threads / 32);
  printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_8x8x4_tflops\": %.2f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f, "
         "\"how\": \"%d blocks x %d threads, %d independent chains per thread, %d iterations, best of 5, CUDA events\"}\n",
         p.name, sms, fl_fma / ms_fma * 1e-9, fl_mma / ms_mma * 1e-9, ms_fma, ms_mma, blocks, threads, ILP, iters);
  return 0;
}
