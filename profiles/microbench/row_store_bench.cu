// row_store_bench.cu -- how fast can one B200 store a CSR value array row by row?  (phase 2 of the owner-computes
// path, csrc/cq_owner.cuh: 17 M dofs x 3 rows x ~190 doubles = 77.5 GB; cudaMemset does it in 10.6 ms = 7.3 TB/s.)
// Variants: warp per dof with 8-byte stores (256 B per instruction, the kernel's pattern), warp per dof with 16-byte
// stores after peeling to alignment, and a plain grid-stride 16-byte fill for reference.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a row_store_bench.cu -o row_store_bench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 8) warp_rows_8(double* v, const long long ndof, const int len) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  for (long long d = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); d < ndof; d += nw) {
    double* row = v + d * len;
    const double x = (double)d;
#pragma unroll 4
    for (int i = lane; i < len; i += 32) row[i] = (i % 3 == 0) ? x : 0.0;
  }
}

__global__ void __launch_bounds__(256, 8) warp_rows_16(double* v, const long long ndof, const int len) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  for (long long d = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); d < ndof; d += nw) {
    double* row = v + d * len;
    const double x = (double)d;
    int i0 = 0;
    if ((reinterpret_cast<unsigned long long>(row) & 15) != 0) {  // peel one double
      if (lane == 0) row[0] = x;
      i0 = 1;
    }
    const int n2 = (len - i0) / 2;
    double2* r2 = reinterpret_cast<double2*>(row + i0);
#pragma unroll 4
    for (int i = lane; i < n2; i += 32) r2[i] = make_double2(((2 * i + i0) % 3 == 0) ? x : 0.0, ((2 * i + 1 + i0) % 3 == 0) ? x : 0.0);
    if (lane == 0 && i0 + 2 * n2 < len) row[len - 1] = 0.0;
  }
}

// a block writes 8 consecutive dofs (one contiguous run) cooperatively
__global__ void __launch_bounds__(256, 8) block_rows_16(double* v, const long long ndof, const int len) {
  const long long nb = gridDim.x;
  for (long long b = blockIdx.x; b * 8 < ndof; b += nb) {
    double* base = v + b * 8 * len;
    long long n = 8LL * len;
    if ((b + 1) * 8 > ndof) n = (ndof - b * 8) * len;
    long long i0 = 0;
    if ((reinterpret_cast<unsigned long long>(base) & 15) != 0) {
      if (threadIdx.x == 0) base[0] = 1.0;
      i0 = 1;
    }
    const long long n2 = (n - i0) / 2;
    double2* r2 = reinterpret_cast<double2*>(base + i0);
    for (long long i = threadIdx.x; i < n2; i += 256) r2[i] = make_double2((double)b, 0.0);
    if (threadIdx.x == 0 && i0 + 2 * n2 < n) base[n - 1] = 0.0;
  }
}

__global__ void fill16(double2* v, const long long n2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) v[i] = make_double2(1.0, 0.0);
}

template <class F>
static float timeit(F f, int reps = 3) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const int len = 571;                    // doubles per dof (3 rows of ~190): odd, so rows alternate 16-byte alignment
  const long long ndof = 8LL * 1000 * 1000;  // 36.5 GB
  const long long n = ndof * len;
  double* v;
  if (cudaMalloc(&v, n * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const double gb = n * 8 / 1e9;
  float t;
  t = timeit([&] { cudaMemsetAsync(v, 0, n * 8); });
  printf("{\"variant\": \"cudaMemset\", \"GB\": %.1f, \"ms\": %.3f, \"TBps\": %.2f}\n", gb, t, gb / t);
  t = timeit([&] { fill16<<<sms * 16, 256>>>(reinterpret_cast<double2*>(v), n / 2); });
  printf("{\"variant\": \"grid-stride 16-byte fill\", \"ms\": %.3f, \"TBps\": %.2f}\n", t, gb / t);
  for (int bps : {8, 4, 2}) {
    t = timeit([&] { warp_rows_8<<<sms * bps, 256>>>(v, ndof, len); });
    printf("{\"variant\": \"warp per dof, 8-byte stores\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"TBps\": %.2f}\n", bps, t, gb / t);
    t = timeit([&] { warp_rows_16<<<sms * bps, 256>>>(v, ndof, len); });
    printf("{\"variant\": \"warp per dof, 16-byte stores\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"TBps\": %.2f}\n", bps, t, gb / t);
    t = timeit([&] { block_rows_16<<<sms * bps, 256>>>(v, ndof, len); });
    printf("{\"variant\": \"block per 8 dofs, 16-byte stores\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"TBps\": %.2f}\n", bps, t, gb / t);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
