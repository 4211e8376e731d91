// cq_kernels.cuh -- kernel parameter block, group reduction, CSR / vector / scalar sinks and the
// direct-load ("v1") assembly kernel: G lanes cooperate on one (cell, rule) entry of the qr list,
// each lane accumulating the points q = r, r+G, ... of that entry in registers.
//
// Replaces (reference): the numba cell loop customquad/assemble_matrix.py:84-111 (gather of
// coords[vertices[cell]], zero-fill, kernel call, MatSetValuesLocal ADD_VALUES),
// assemble_vector.py:74-94, assemble_scalar.py:64-82.
#pragma once
#include "cq_forms.cuh"

namespace cq {

struct KParams {
  // one qr tuple in ragged form (cq_quadrature)
  int64_t num_entries;
  const int32_t* __restrict__ cells;
  const int64_t* __restrict__ offsets;
  const double* __restrict__ pts;
  const double* __restrict__ w;
  const double* __restrict__ normals;
  // mesh (cq_mesh)
  const int32_t* __restrict__ geom_dofmap;
  const double* __restrict__ coords;
  // space (cq_space)
  const int32_t* __restrict__ dofmap;
  const double* __restrict__ coeffs;
  // sinks
  const int64_t* __restrict__ indptr;
  const int32_t* __restrict__ indices;
  double* values;  // CSR values (matrix)
  long long peer_row0;                         // rows >= peer_row0 are scattered into a neighbour's matrix (cq_csr)
  const long long* __restrict__ peer_rowstart; // [num_rows - peer_row0 + 1], relative to `values`
  double* out;     // vector b / scalar m / element tensors
  int* err_flag;
  const uint8_t* __restrict__ elem_pos;  // scatter positions [num_cells, NPOS] or null
  unsigned int* work_counter;            // ctx-owned ticket counter of the staged kernels, zeroed before a launch with `dynamic`
  int dynamic;                           // 1: producers draw their stages from work_counter (in traversal order) instead of the static interleaving
  int64_t num_points;                    // length of pts / w (bulk-copy tail handling)
  long long tile_sp, tile_sg;            // traversal hint (cq_quadrature.plane_entries / plane_tiles) in stages: per plane, per tile; 0 = list order
  int stage_points;                      // points capacity of one smem stage
  int mesh_flags;                        // cq_mesh.flags
  int dbg;                               // profiling builds only (-DCQ_PROFILING, CQ_DEBUG): 1 consumers skip all work, 2 skip the scatter
  int bs;          // block size used by the matrix scatter (compile-time BS for rank < 2)
  int elem_out;    // 1: write un-scattered element tensors to out[k, :]
  double* elem_scalar;  // owner-computes path (cq_owner.cuh): the kernel writes the scalar N x N block of entry k to elem_scalar[k, :] instead of scattering it
  long long num_dofs;   // scalar dofs of the space (rows / bs)
  // fused A.zeroEntries() (cq_assemble_matrix_zero, cq_zero.cuh): the kernel itself zeroes `values` just ahead of
  // the scatter, so the REDs hit zeroed lines that are still in L2 instead of reading a memset's lines back from HBM
  int zero_req;                          // 1: values are uninitialised, zero them (in the kernel, or by a memset)
  long long zero_keep_lo, zero_keep_hi;  // ... except values[zero_keep_lo, zero_keep_hi) (cq_csr.zero_keep_*: zeroed by the caller before its fence)
  int zero_ahead;                        // waves zeroed ahead of the scatter
  long long nnz;
  const long long* __restrict__ zero_hi; // [waves] values[0 .. zero_hi[k]) covers every slot the waves <= k touch; null: no fused zeroing
  unsigned int* zero_done;               // [waves] warps that have finished zeroing their part of wave k's range
  double betaN;
  double p[CQ_MAX_CONSTS];
};

// ---------------------------------------------------------------------------------------------
// Transpose-reduce over the G lanes of a group: on return lane r (0..G-1) holds in acc[0..M/G)
// the group sums of the original entries [r*M/G, (r+1)*M/G).  M*(1-1/G) 64-bit shuffles instead
// of M*log2(G) for a butterfly.  All 32 lanes of the warp must call it.
// ---------------------------------------------------------------------------------------------
template <int M, int S>
CQ_DEVINL void transpose_reduce_step(double* acc, const int r) {
  if constexpr (S >= 1) {
    constexpr int H = M / 2;
    const bool up = (r & S) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const double send = up ? acc[i] : acc[i + H];
      const double keep = up ? acc[i + H] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
    }
    transpose_reduce_step<H, S / 2>(acc, r);
  }
}

CQ_HD constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// work-skipping experiments exist in profiling builds only; in the product CQ_DBG(P) is the constant 0
#ifdef CQ_PROFILING
#define CQ_DBG(P) ((P).dbg)
#else
#define CQ_DBG(P) 0
#endif

// position of (row, col) in a sorted-column CSR, -1 if absent (MatSetValuesLocal's row search)
CQ_DEVINL int64_t csr_find(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const int64_t row,
                           const int32_t col) {
  int64_t lo = __ldg(indptr + row);
  const int64_t end = __ldg(indptr + row + 1);
  int64_t hi = end;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(indices + mid) < col) lo = mid + 1;
    else hi = mid;
  }
  return (lo < end && __ldg(indices + lo) == col) ? lo : -1;
}

CQ_DEVINL void csr_add(const KParams& P, const int64_t row, const int32_t col, const double v) {
  const int64_t s = csr_find(P.indptr, P.indices, row, col);
  if (s >= 0) red_add_f64(P.values + s, v);
  else atomicExch(P.err_flag, 1);
}

// Traversal position -> stage index (32 consecutive list entries) under the plane / tile hint: the NS stages form
// planes of SP stages (the last one possibly partial); tile g holds the stages [g SG, (g+1) SG) of every plane; the
// traversal runs tile by tile, inside a tile plane by plane.  A bijection of [0, NS); identity without the hint.
CQ_DEVINL long long stage_of(const KParams& P, long long tp, const long long NS) {
  const long long SP = P.tile_sp, SG = P.tile_sg;
  if (SP <= 0) return tp;
  const long long NPf = NS / SP, L = NS - NPf * SP;
  for (long long g0 = 0;; g0 += SG) {
    const long long len = SG < SP - g0 ? SG : SP - g0;                       // stages of this tile inside a full plane
    const long long endL = L < g0 + len ? L : g0 + len;
    const long long lenL = endL > g0 ? endL - g0 : 0;                        // ... inside the partial last plane
    const long long cnt = NPf * len + lenL;
    if (tp < cnt || g0 + SG >= SP) {
      const long long p = tp < NPf * len ? tp / len : NPf;
      return p * SP + g0 + (tp - p * len);
    }
    tp -= cnt;
  }
}

// upper-triangular entry index -> (i, j)
template <int N>
CQ_DEVINL void tri_unpack(const int e, int& i, int& j) {
  int row = 0, start = 0;
#pragma unroll 1
  while (row + 1 < N && start + (N - row) <= e) {
    start += N - row;
    ++row;
  }
  i = row;
  j = row + (e - start);
}

// ---------------------------------------------------------------------------------------------
// Sink for one reduced element tensor.  Lane r of the group owns entries [r*B, (r+1)*B) of the
// padded accumulator (B = MPAD/G), held in acc[0..B).
// ---------------------------------------------------------------------------------------------
template <class T, int G, int MPAD>
CQ_DEVINL void sink_entries(const KParams& P, const double* acc, const int r, const int64_t k, const int32_t cell) {
  constexpr int B = MPAD / G;
  constexpr int N = T::N;
  if constexpr (T::RANK == 2) {
    const int bs = P.bs;
    const int nl = N * bs;
#pragma unroll
    for (int kk = 0; kk < B; ++kk) {
      const int e = r * B + kk;
      if (e < T::NACC) {
        int i, j;
        tri_unpack<N>(e, i, j);
        const double v = acc[kk];
        if (P.elem_out) {
          double* o = P.out + k * (int64_t)(nl * nl);
          for (int a = 0; a < bs; ++a) {
            o[(i * bs + a) * nl + j * bs + a] = v;
            o[(j * bs + a) * nl + i * bs + a] = v;
          }
        } else {
          const int32_t di = __ldg(P.dofmap + (int64_t)cell * N + i);
          const int32_t dj = __ldg(P.dofmap + (int64_t)cell * N + j);
          for (int a = 0; a < bs; ++a) {
            csr_add(P, (int64_t)di * bs + a, dj * bs + a, v);
            if (i != j) csr_add(P, (int64_t)dj * bs + a, di * bs + a, v);
          }
        }
      }
    }
  } else if constexpr (T::RANK == 1) {
#pragma unroll
    for (int kk = 0; kk < B; ++kk) {
      const int e = r * B + kk;
      if (e < T::NACC) {
        if (P.elem_out) {
          P.out[k * T::NACC + e] = acc[kk];
        } else {
          const int i = e / T::BS, a = e % T::BS;
          const int32_t di = __ldg(P.dofmap + (int64_t)cell * N + i);
          red_add_f64(P.out + (int64_t)di * T::BS + a, acc[kk]);  // b[dofs[cell, j]] += b_local[j]
        }
      }
    }
  }
}

// gather coords[vertices[cell, :]] (assemble_matrix.py:86): X[v][k], v in local vertex order
template <int DIM>
CQ_DEVINL void load_cell_vertices(const KParams& P, const int32_t cell, double (&X)[1 << DIM][DIM]) {
  constexpr int NV = 1 << DIM;
  int32_t vid[NV];
  if constexpr (DIM == 2) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(P.geom_dofmap) + cell);
    vid[0] = v.x; vid[1] = v.y; vid[2] = v.z; vid[3] = v.w;
  } else {
    const int4 v0 = __ldg(reinterpret_cast<const int4*>(P.geom_dofmap) + 2 * (int64_t)cell);
    const int4 v1 = __ldg(reinterpret_cast<const int4*>(P.geom_dofmap) + 2 * (int64_t)cell + 1);
    vid[0] = v0.x; vid[1] = v0.y; vid[2] = v0.z; vid[3] = v0.w;
    vid[4] = v1.x; vid[5] = v1.y; vid[6] = v1.z; vid[7] = v1.w;
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int k = 0; k < DIM; ++k) X[v][k] = __ldg(P.coords + 3 * (int64_t)vid[v] + k);
}

// block-wide sum of one double per thread -> one atomic per block (scalar forms)
CQ_DEVINL void block_sum_atomic(double v, double* target) {
  __shared__ double warp_sums[32];
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_sums[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double t = lane < nw ? warp_sums[lane] : 0.0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
    if (lane == 0) atomicAdd(target, t);
  }
}

// ---------------------------------------------------------------------------------------------
// v1: direct global loads.  One group of G lanes per qr-list entry; a warp handles 32/G
// consecutive entries per round; rounds are grid-strided.
// ---------------------------------------------------------------------------------------------
template <class T, int G, int THREADS>
__global__ void __launch_bounds__(THREADS) assemble_direct_kernel(const KParams P) {
  constexpr int DIM = T::DIM;
  constexpr int CPW = 32 / G;
  constexpr int MPAD = round_up(T::NACC, G);
  const int lane = threadIdx.x & 31;
  const int r = lane & (G - 1);
  const int grp = lane / G;
  const int64_t warp_id = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
  const int64_t num_warps = (int64_t)gridDim.x * (THREADS / 32);
  double msum = 0.0;

  for (int64_t base = warp_id * CPW; base < P.num_entries; base += num_warps * CPW) {
    const int64_t k = base + grp;
    const bool active = k < P.num_entries;
    int32_t cell = 0;
    int64_t q0 = 0, q1 = 0;
    if (active) {
      cell = __ldg(P.cells + k);
      q0 = __ldg(P.offsets + k);
      q1 = __ldg(P.offsets + k + 1);
    }
    double acc[MPAD];
#pragma unroll
    for (int i = 0; i < MPAD; ++i) acc[i] = 0.0;

    if (q1 > q0) {
      double X[1 << DIM][DIM];
      load_cell_vertices<DIM>(P, cell, X);
      CellGeom<DIM> geo;
      geo.init(X, T::NEED_H);
      init_affine<DIM>(geo);
      const double gamma = T::NEED_H ? P.betaN / geo.hK : 0.0;
      const double* coef = nullptr;
      double coef_reg[T::NEED_COEF && T::NCOEF <= 32 ? T::NCOEF : 1];
      if constexpr (T::NEED_COEF) {
        coef = P.coeffs + (int64_t)cell * T::NCOEF;
        if constexpr (T::NCOEF <= 32) {
#pragma unroll
          for (int i = 0; i < T::NCOEF; ++i) coef_reg[i] = __ldg(coef + i);
          coef = coef_reg;
        }
      }
      const double* fp = T::NEED_NORMAL ? P.p + 1 : P.p;
      double (&accT)[T::NACC] = *reinterpret_cast<double (*)[T::NACC]>(&acc[0]);
      for (int64_t q = q0 + r; q < q1; q += G) {
        double xi[DIM], nrm[DIM];
        if constexpr (DIM == 2) {
          const double2 t = __ldg(reinterpret_cast<const double2*>(P.pts) + q);
          xi[0] = t.x; xi[1] = t.y;
          if constexpr (T::NEED_NORMAL) {
            const double2 n = __ldg(reinterpret_cast<const double2*>(P.normals) + q);
            nrm[0] = n.x; nrm[1] = n.y;
          }
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            xi[d] = __ldg(P.pts + 3 * q + d);
            if constexpr (T::NEED_NORMAL) nrm[d] = __ldg(P.normals + 3 * q + d);
          }
        }
        const double w = __ldg(P.w + q);
        accumulate_point<T>(accT, geo, xi, w, nrm, coef, gamma, fp);
      }
    }

    if constexpr (T::RANK == 0) {
      if (P.elem_out) {
        double v = acc[0];
#pragma unroll
        for (int s = G / 2; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (active && r == 0) P.out[k] = v;
      } else {
        msum += acc[0];  // m[0] += m_local[0] (assemble_scalar.py:82), reduced once per block below
      }
    } else {
      transpose_reduce_step<MPAD, G / 2>(acc, r);
      if (active) sink_entries<T, G, MPAD>(P, acc, r, k, cell);
    }
  }
  if constexpr (T::RANK == 0) {
    if (!P.elem_out) block_sum_atomic(msum, P.out);
  }
}

}  // namespace cq
