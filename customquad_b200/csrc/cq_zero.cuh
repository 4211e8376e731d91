// cq_zero.cuh -- A.zeroEntries() fused into the persistent assembly kernels (cq_assemble_matrix_zero).
//
// The reference zeroes the matrix and then adds the element tensors (assemble_matrix.py:20, :103-111).  Done
// as a memset followed by the kernel, every line of `values` crosses the HBM interface three times: the memset
// writes it, the first RED reads it back (the matrix is far larger than the L2) and the eviction writes it again
// -- measured on the Q2 bench matrix (2.1 GB): 2.1 GB of the kernel's 7.1 GB of DRAM reads, plus 0.29 ms of memset.
//
// The persistent kernels deal the stages (32 entries) out statically: stage t runs on CTA t % grid in "wave"
// t / grid.  For the usual cell lists (ascending cells) the rows a wave touches move through the matrix like a
// front, so the zeroing can run just ahead of it:
//   setup (two small kernels per call):  zero_hi[k] = max over the waves <= k of (end of the last CSR row a wave
//     touches), i.e. values[0 .. zero_hi[k]) covers everything the waves 0..k scatter into (any cell order: a
//     shuffled list just makes zero_hi[0] large);
//   in the kernel (assemble_moment_split_kernel, ZW = 1):  one more warp per CTA zeroes the CTA's share of the
//     ranges of four waves at a time (plain coalesced 16-B stores), zero_ahead waves ahead of the CTA's moment
//     phase (paced through a shared-memory progress word), fences, and counts the CTA into zero_done[k] of those
//     waves; before the expansion warps scatter wave k they look at the counters of the waves up to k (loaded one
//     iteration ahead; complete long ago unless a CTA lags more than the look-ahead).  The never-touched rest
//     [zero_hi[last], nnz) is zeroed up front.
// The zeroed lines are still dirty in L2 when the REDs arrive a few microseconds later, so a line is written to
// HBM once and never read (ncu: 9.2 -> 7.04 GB per Q2 launch).  The kernel is launched cooperatively: the wait
// needs all CTAs resident.  Opt-in (cq_ctx_set_fused_zero): with the traffic gone the kernel is bound by its
// stage ring and the step time does not change -- profiles/r1/fused_zero_study.md.
#pragma once
#include "cq_kernels.cuh"

namespace cq {

// The completion counters are read with relaxed gpu-scope loads (served by the L2) and no acquire fence: what follows
// the check are only REDs -- atomics performed at the L2, issued after the branch on the loaded value resolved --
// to lines whose zero stores were performed at the L2 before the counter was incremented (fence on the zeroing
// side).  An ld.acquire here stalls every later memory instruction of the warp for the load's round trip, once per
// stage (measured: Q1 kernel 0.80 -> 1.14 ms).
CQ_DEVINL unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
CQ_DEVINL void red_relaxed_inc_u32(unsigned* p) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// One warp zeroes part `part` of `nparts` of values[lo, hi) (lo a multiple of 4, `values` 16-B aligned).
CQ_DEVINL void zero_share(double* values, const long long lo, const long long hi, const long long part, const long long nparts,
                          const int lane) {
  const long long len = hi - lo;
  if (len <= 0) return;
  const long long chunk = ((len + nparts - 1) / nparts + 3) & ~3LL;  // whole 32-B sectors
  const long long a = lo + part * chunk;
  long long b = a + chunk;
  if (b > hi) b = hi;
  for (long long i = a + 2 * lane; i < b; i += 64) {
    if (i + 1 < b) *reinterpret_cast<double2*>(values + i) = make_double2(0.0, 0.0);
    else values[i] = 0.0;
  }
}

// ... leaving values[klo, khi) alone (cq_csr.zero_keep_*: the interface rows of a peer scatter, zeroed by the caller
// before its fence -- the neighbour's kernels may already have added into them)
CQ_DEVINL void zero_share_keep(double* values, const long long lo, const long long hi, const long long klo, const long long khi,
                               const long long part, const long long nparts, const int lane) {
  if (khi <= klo || hi <= klo || lo >= khi) {
    zero_share(values, lo, hi, part, nparts, lane);
    return;
  }
  if (lo < klo) zero_share(values, lo, klo, part, nparts, lane);
  long long b = khi;
  if (b < hi) {
    if (b & 1) {  // the 16-byte stores of zero_share start at an even index
      if (part == 0 && lane == 0) values[b] = 0.0;
      ++b;
    }
    zero_share(values, b, hi, part, nparts, lane);
  }
}

// ---- setup kernels -------------------------------------------------------------------------------------------
// one block per wave: raw[wave] = max(end of the last (blocked) CSR row the wave's cells touch), rounded up to 4
template <int DUMMY = 0>
__global__ void wave_hi_kernel(const int32_t* __restrict__ cells, const int64_t num_entries, const int32_t* __restrict__ dofmap,
                               const int n, const int bs, const int64_t* __restrict__ indptr, const long long nnz, const int grid_main,
                               unsigned long long* __restrict__ raw) {
  __shared__ int smax[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t wave = blockIdx.x; wave * grid_main * 32 < num_entries; wave += gridDim.x) {
    const int64_t k0 = wave * grid_main * 32;
    const int64_t k1 = k0 + (int64_t)grid_main * 32 < num_entries ? k0 + (int64_t)grid_main * 32 : num_entries;
    int m = -1;
    for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
      const int32_t* d = dofmap + (int64_t)__ldg(cells + k) * n;
      for (int i = 0; i < n; ++i) m = max(m, __ldg(d + i));
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) smax[warp] = m;
    __syncthreads();
    if (warp == 0) {
      m = lane < nwarps ? smax[lane] : -1;
      m = __reduce_max_sync(0xffffffffu, m);
      if (lane == 0) {
        long long hi = m >= 0 ? __ldg(indptr + ((long long)m * bs + bs)) : 0;  // end of row m*bs + bs-1
        hi = (hi + 3) & ~3LL;
        raw[wave] = (unsigned long long)(hi > nnz ? nnz : hi);
      }
    }
    __syncthreads();
  }
}

// in-place prefix maximum over `waves` entries (one block)
template <int DUMMY = 0>
__global__ void wave_scan_kernel(unsigned long long* __restrict__ raw, const int64_t waves) {
  __shared__ unsigned long long part[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t per = (waves + nt - 1) / nt;
  const int64_t a = (int64_t)tid * per, b = a + per < waves ? a + per : waves;
  unsigned long long m = 0;
  for (int64_t i = a; i < b; ++i) m = max(m, raw[i]);
  part[tid] = m;
  __syncthreads();
  for (int off = 1; off < nt; off <<= 1) {  // inclusive scan (max) of the segment maxima
    const unsigned long long o = tid >= off ? part[tid - off] : 0ull;
    __syncthreads();
    part[tid] = max(part[tid], o);
    __syncthreads();
  }
  unsigned long long run = tid > 0 ? part[tid - 1] : 0ull;
  for (int64_t i = a; i < b; ++i) {
    run = max(run, raw[i]);
    raw[i] = run;
  }
}

}  // namespace cq
