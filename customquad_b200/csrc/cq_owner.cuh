// cq_owner.cuh -- owner-computes assembly of blocked (bs >= 2) Q2-hexahedron matrices (config 4:
// demo/vector_poisson.py-style vector Poisson on Q2 hexahedra, reference demo/vector_poisson.py:95,121-127).
//
// The reference adds the 81 x 81 element tensor of every cell into the PETSc matrix (assemble_matrix.py:103-111).
// On the blocked AIJ pattern a scalar coupling (i, j) owns a bs x bs block of slots of which the forms of the
// menu fill the diagonal only, so a RED scatter touches every 32-byte sector of the value array for one double in
// three: with A.zeroEntries() in front, each line crosses the HBM interface three times (memset write, read for the
// first RED, write-back) -- 270 GB per launch on the 128^3 mesh against 77.5 GB of matrix (profiles/r2/config4_dram_traffic.txt).
//
// Here every matrix row is written exactly once, by its owner:
//   phase 0  (setup, per call):  transpose of the dofmap restricted to the listed entries: for each scalar dof the
//            (list position k, local index li) pairs of the entries that touch it (count, exclusive scan, fill);
//   phase 1  the staged tensor-core kernel (cq_dmma.cuh) computes the scalar 27 x 27 block of each entry and writes it
//            to a work buffer [num_entries, 729] instead of scattering it (KParams::elem_scalar);
//   phase 2  one warp per scalar dof gathers row li of the blocks of its entries (in ascending list order: the result
//            is bit-reproducible from run to run, unlike the RED scatter) into a shared-memory row through the scatter
//            map, and stores the bs blocked rows -- diagonal slots from the row, the others zero -- with coalesced
//            256-byte stores; rows nobody touches are stored as zeros (A.zeroEntries() fused, no memset).
// Used for the call that zeroes the matrix (cq_assemble_matrix_zero: assemble_matrix(form, qr)); adding to an existing
// matrix stays with the RED scatter, which is faster there (measured at 128^3: 46.6 ms against 54 ms for a
// read-modify-write of every row by its owner).
// Traffic: 12.2 GB of blocks written and read back + 77.5 GB of matrix written once.
// Not used with peer rows (the neighbour rank's rows are not this rank's to store) -- the RED path handles those.
#pragma once
#include <cub/device/device_scan.cuh>

#include "cq_kernels.cuh"

namespace cq {

struct OwnerParams {
  int64_t num_entries;
  int64_t num_dofs;
  int n;     // scalar dofs per cell
  int npos;  // scatter-map stride
  int bs;  // 2 or 3 here
  const int32_t* __restrict__ cells;
  const int32_t* __restrict__ dofmap;
  const uint8_t* __restrict__ elem_pos;
  const int64_t* __restrict__ indptr;
  double* values;
  int* cnt;                        // [num_dofs + 1]
  unsigned int* ptr;               // [num_dofs + 1]
  int2* pairs;                     // [num_entries * n]: (k * 32 + li, cell), grouped by dof
  const double* __restrict__ elem; // [num_entries, n * n]
};

__global__ void owner_count_kernel(const OwnerParams O) {
  const int64_t total = O.num_entries * O.n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = idx / O.n;
    const int li = (int)(idx - k * O.n);
    const int32_t dof = __ldg(O.dofmap + (int64_t)__ldg(O.cells + k) * O.n + li);
    atomicAdd(O.cnt + dof, 1);
  }
}

// cnt[dof] counts down to zero again: the buffer is clean for the next call
__global__ void owner_fill_kernel(const OwnerParams O) {
  const int64_t total = O.num_entries * O.n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = idx / O.n;
    const int li = (int)(idx - k * O.n);
    const int32_t cell = __ldg(O.cells + k);
    const int32_t dof = __ldg(O.dofmap + (int64_t)cell * O.n + li);
    const int slot = atomicSub(O.cnt + dof, 1) - 1;
    O.pairs[(int64_t)O.ptr[dof] + slot] = make_int2((int)k * 32 + li, cell);  // n <= 32, num_entries * n < 2^31
  }
}

constexpr int OWNER_WARPS = 8;
constexpr int OWNER_ROW_MAX = 256;  // scatter positions are bytes: a scalar row holds at most 256 couplings

// BS: the block size as a compile-time constant (the slot -> (coupling, component) split of the store loop is a
// division per stored double: with a run-time divisor the kernel was bound by its instruction count, 23 ms for the
// 77.5 GB of config 4 instead of the ~13 ms the stores take)
template <int BS_T>  // 0: run-time block size
__global__ void __launch_bounds__(OWNER_WARPS * 32, 8) owner_rows_kernel(const OwnerParams O) {
  const int BS = BS_T > 0 ? BS_T : O.bs;
  __shared__ double buf_all[OWNER_WARPS][OWNER_ROW_MAX];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* __restrict__ buf = buf_all[wib];
  const int n = O.n, nn = O.n * O.n;
  const int64_t nwarps = (int64_t)gridDim.x * OWNER_WARPS;
  for (int64_t d = (int64_t)blockIdx.x * OWNER_WARPS + wib; d < O.num_dofs; d += nwarps) {
    const unsigned int p0 = __ldg(O.ptr + d), p1 = __ldg(O.ptr + d + 1);
    const int64_t r0 = __ldg(O.indptr + d * BS);
    const int len_b = (int)(__ldg(O.indptr + d * BS + 1) - r0);  // blocked row length (the BS rows of a dof are equally long and contiguous)
    const int len = len_b / BS;
    for (int p = lane; p < len; p += 32) buf[p] = 0.0;
    __syncwarp();
    for (unsigned int c0 = p0; c0 < p1; c0 += 32) {
      const int m = (int)(p1 - c0 < 32u ? p1 - c0 : 32u);
      int2 pr = make_int2(0x7fffffff, 0);
      if (lane < m) pr = O.pairs[c0 + lane];
      int key = pr.x;
      // ascending list position: a fixed summation order.  U entries at a time: their loads are all in flight before
      // the first add (the adds of one entry hit distinct slots; consecutive entries are ordered by the warp barrier)
      constexpr int U = 4;
      for (int it = 0; it < m; it += U) {
        double v[U];
        unsigned pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int kmin = __reduce_min_sync(0xffffffffu, key);  // 0x7fffffff once the chunk is exhausted
          const int src = __ffs(__ballot_sync(0xffffffffu, key == kmin)) - 1;
          const int32_t cl = __shfl_sync(0xffffffffu, pr.y, src);
          if (lane == src) key = 0x7fffffff;
          const int li = kmin & 31;
          v[u] = 0.0;
          pos[u] = 0;
          if (it + u < m && lane < n) {
            v[u] = __ldg(O.elem + (int64_t)(kmin >> 5) * nn + li * n + lane);
            pos[u] = __ldg(O.elem_pos + (int64_t)cl * O.npos + li * n + lane);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (it + u < m && lane < n) buf[pos[u]] += v[u];
          __syncwarp();
        }
      }
    }
    // row a of the dof (the BS rows are contiguous and equally long): slot idx = p * BS + b holds buf[p] iff b == a
#pragma unroll
    for (int a = 0; a < (BS_T > 0 ? BS_T : BS); ++a) {
      double* __restrict__ row = O.values + r0 + (int64_t)a * len_b;
#pragma unroll 4
      for (int idx = lane; idx < len_b; idx += 32) {
        const int p = idx / BS;
        row[idx] = (idx - p * BS == a) ? buf[p] : 0.0;
      }
    }
    __syncwarp();
  }
}

}  // namespace cq
