// cq_staged.cuh -- the bulk-copy (TMA) staged, warp-specialised, persistent assembly kernel.
//
// One CTA per SM.  NPW producer warps each own one shared-memory stage; NCW consumer warps do the
// arithmetic.  A stage holds CB = 32 consecutive entries of the qr list (one producer lane per
// entry):
//
//   producer warp p, forever:
//     wait until stage p is empty; grab the next stage index t from a global counter (dynamic
//     load balance: ragged rules make stages unequal); the stage's points / weights / normals are
//     one contiguous slice of the flat arrays -> two or three cp.async.bulk copies (UBLKCP) that
//     complete on the stage's mbarrier; meanwhile each lane gathers everything else its entry
//     needs by pointer chasing -- cell id, geometry dofmap -> vertex coordinates, dofmap -> CSR
//     row starts, scatter positions, packed coefficient -- into the stage header; finally the
//     lanes sort the entries by trip count (longest first) and the stage is published.
//   consumer warps:
//     wait for a full stage, claim 32/G entries at a time from its sorted list (a shared-memory
//     counter), G lanes per entry accumulate the element tensor from shared memory in registers,
//     transpose-reduce over the group, and scatter with red.global.add.f64 using the prefetched
//     row starts and positions -- no global loads at all on the consumer side.
//
// Replaces the same reference code as the direct kernel (assemble_matrix.py:84-111 etc.); the two
// give identical results up to FP64 summation order.
#pragma once
#include "cq_kernels.cuh"

#ifndef CQ_STAGED_UNROLL_VS  // point loop of the vector / scalar kernels with at most 16 accumulators: points in flight per lane
#define CQ_STAGED_UNROLL_VS 2
#endif
namespace cq {

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------------------
CQ_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

CQ_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
CQ_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CQ_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CQ_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CQ_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra CQ_DONE_%=;\n"
      "bra CQ_WAIT_%=;\n"
      "CQ_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned
CQ_DEVINL void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// The same copy for data that is read exactly once (the rule's points / weights / normals, the scatter map): L2
// evict-first, so the stream does not push the matrix lines (zeroed ahead of the scatter, cq_zero.cuh) out of the L2.
CQ_DEVINL uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
CQ_DEVINL void bulk_g2s_once(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, const uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}

// L2 prefetch of a global range (bytes % 16 == 0, address 16-B aligned): the producer warps issue it for the
// rule slice of the stage they will publish two ring turns later, so that the stage's bulk copy into shared memory
// is an L2 hit (~0.4 us) instead of a DRAM access under load (1-2 us).  The L2 acts as the deep end of the stage
// ring: the number of bytes in flight from HBM is no longer tied to the shared-memory slots.
CQ_DEVINL void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// ---- stage layout ------------------------------------------------------------------------------------
template <class T, int CB>
struct StageMeta {
  static constexpr int N = T::N, NV = T::NV, DIM = T::DIM;
  static constexpr int NPOS = round_up(N * N, 16);
  static constexpr bool MAT = T::RANK == 2;
  static constexpr bool VEC = T::RANK == 1;
  int nent;              // entries in this stage, -1: no more work
  int next;              // consumer claim counter
  long long k0;          // list position of the stage's first entry (element-tensor output of the owner-computes path, cq_owner.cuh)
  int order[CB];         // entry indices sorted by trip count, longest first
  int nq[CB];
  int cell[CB];
  const double* pts[CB];  // generic pointers: into this stage's shared memory, or global on overflow
  const double* w[CB];
  const double* nrm[T::NEED_NORMAL ? CB : 1];
  // Entry-major with an odd row stride: the producer writes with lane = entry (stride NRP, odd: conflict-free), the
  // transposed scatter reads with lane = value slot (a few distinct rows of ONE entry: neighbouring banks).  The
  // round-1 layout [N][CB] put every row of an entry into the same bank (CB * 8 B = 2 x 128 B): ncu counted 4
  // wavefronts per rowstart load of the scatter instead of 2 (12 of the 74 LSU wavefronts per Q2 cell; the kernels
  // are bound by the LSU data pipe, profiles/r2/lsu_study.md).
  static constexpr int NRP = (MAT ? N : 1) | 1;
  long long rowstart[CB][NRP];  // indptr[dof_i * bs]
  int rowlen[CB][NRP];          // length of that (blocked) row
  int dofs[VEC ? N : 1][CB];
  alignas(16) uint8_t pos[MAT ? CB : 1][MAT ? NPOS : 16];
  double coef[T::NEED_COEF ? T::NCOEF : 1][CB];
  // monomial moments handed from the moment warps to the expansion warps (cq_moment.cuh, split kernel)
  // (quadrilaterals of degree 1 and 2, hexahedra of degree 1)
  // The symmetric Nitsche form carries DIM + 1 moment sets: weights w and w n_d (quadrature normals).
  // (hexahedra only: on quadrilaterals the point-by-point kernel is faster, and the header space is needed for points)
  static constexpr bool NIT = T::FORM_ID == CQ_FORM_NITSCHE && T::DIM == 3 && T::DEG == 1;
  static constexpr bool MOMF = T::FORM_ID == CQ_FORM_MASS || T::FORM_ID == CQ_FORM_STIFFNESS || NIT;
  static constexpr bool MOM = MOMF && (T::DIM == 2 || (T::DIM == 3 && T::DEG == 1));
  static constexpr int PN = 2 * T::DEG + 1;
  static constexpr int NM1 = T::DIM == 2 ? PN * PN : PN * PN * PN;  // moments per set
  static constexpr int NSET = NIT ? T::DIM + 1 : 1;
  static constexpr int NMOM = MOM ? NSET * NM1 : 1;
  static constexpr int NMP = NMOM | 1;  // entry-major, odd stride: the expansion warps read with lane = entry
  // The vertex coordinates X are read by the point-by-point consumers only, the moments mu by the expansion warps of
  // the moment kernels only (those work from the cell constants kc): the two share their storage.
  union {
    double X[NV * DIM][CB];
    double mu[CB][NMP];
  };
  // axis-aligned cell constants (producer): [0] |det J|; stiffness: [1 + d] |det J| / J_dd^2;
  // Nitsche: [1 + d] |det J| / J_dd (signed), [1 + DIM] |det J| betaN / h_K
  double kc[MOM ? 5 : 1][CB];
};

template <class T, int CB>
struct StagedLayout {
  static constexpr int DIM = T::DIM;
  static constexpr size_t meta_bytes = (sizeof(StageMeta<T, CB>) + 127) / 128 * 128;
  // +2 points of slack: the slice starts at an even point index, and is rounded up to an even count
  CQ_HD static size_t pts_bytes(int pb) { return ((size_t)(pb + 2) * DIM * 8 + 127) / 128 * 128; }
  CQ_HD static size_t w_bytes(int pb) { return ((size_t)(pb + 2) * 8 + 127) / 128 * 128; }
  CQ_HD static size_t stage_bytes(int pb) { return meta_bytes + pts_bytes(pb) * (T::NEED_NORMAL ? 2 : 1) + w_bytes(pb); }
};

// ---- consumer: one entry per group of G lanes --------------------------------------------------------
// (i, j) of the accumulator slices a lane owns after the transpose-reduce: fixed for the whole kernel
template <class T, int G>
struct LaneEntries {
  static constexpr int B = round_up(T::NACC, G) / G;
  unsigned char i[B], j[B];
  CQ_DEVINL void init(const int r) {
#pragma unroll
    for (int kk = 0; kk < B; ++kk) {
      int ii = 0, jj = 0;
      if constexpr (T::RANK == 2) {
        if (r * B + kk < T::NACC) tri_unpack<T::N>(r * B + kk, ii, jj);
      }
      i[kk] = (unsigned char)ii;
      j[kk] = (unsigned char)jj;
    }
  }
};

template <class T, int G, int CB>
CQ_DEVINL void staged_process_entry(const KParams& P, const StageMeta<T, CB>& M, const int e, const bool active, const int r,
                                    double& msum, const LaneEntries<T, G>& LE) {
  constexpr int DIM = T::DIM, N = T::N, NV = T::NV;
  constexpr int MPAD = round_up(T::NACC, G);
  double acc[MPAD];
#pragma unroll
  for (int i = 0; i < MPAD; ++i) acc[i] = 0.0;
  const int nq = active ? M.nq[e] : 0;
  if (nq > 0) {
    double X[NV][DIM];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < DIM; ++k) X[v][k] = M.X[v * DIM + k][e];
    CellGeom<DIM> geo;
    geo.init(X, T::NEED_H);
    init_affine<DIM>(geo);
    const double gamma = T::NEED_H ? P.betaN / geo.hK : 0.0;
    double coef_reg[T::NEED_COEF ? T::NCOEF : 1];
    if constexpr (T::NEED_COEF) {
#pragma unroll
      for (int i = 0; i < T::NCOEF; ++i) coef_reg[i] = M.coef[i][e];
    }
    const double* fp = T::NEED_NORMAL ? P.p + 1 : P.p;
    const double* pp = M.pts[e];
    const double* wp = M.w[e];
    const double* np_ = T::NEED_NORMAL ? M.nrm[e] : nullptr;
    double (&accT)[T::NACC] = *reinterpret_cast<double (*)[T::NACC]>(&acc[0]);
    constexpr int UNR = (T::RANK < 2 && T::NACC <= 16) ? CQ_STAGED_UNROLL_VS : 2;  // points in flight per lane
#pragma unroll UNR
    for (int q = r; q < nq; q += G) {
      double xi[DIM], nrm[DIM];
      if constexpr (DIM == 2) {
        const double2 t = *reinterpret_cast<const double2*>(pp + 2 * q);
        xi[0] = t.x; xi[1] = t.y;
        if constexpr (T::NEED_NORMAL) {
          const double2 n = *reinterpret_cast<const double2*>(np_ + 2 * q);
          nrm[0] = n.x; nrm[1] = n.y;
        }
      } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          xi[d] = pp[3 * q + d];
          if constexpr (T::NEED_NORMAL) nrm[d] = np_[3 * q + d];
        }
      }
      const double w = wp[q];
      accumulate_point<T>(accT, geo, xi, w, nrm, coef_reg, gamma, fp);
    }
  }
  if constexpr (T::RANK == 0) {
    msum += acc[0];
  } else {
    transpose_reduce_step<MPAD, G / 2>(acc, r);
    if (active) {
      constexpr int B = MPAD / G;
      if constexpr (T::RANK == 2) {
        const int bs = P.bs;
#pragma unroll
        for (int kk = 0; kk < B; ++kk) {
          const int en = r * B + kk;
          if (en < T::NACC) {
            const int i = LE.i[kk], j = LE.j[kk];
            const double v = acc[kk];
            const unsigned pij = M.pos[e][i * N + j], pji = M.pos[e][j * N + i];
            double* pi = P.values + M.rowstart[e][i];
            double* pj = P.values + M.rowstart[e][j];
            if (bs == 1) {
              red_add_f64(pi + pij, v);
              if (i != j) red_add_f64(pj + pji, v);
            } else {
              const int li = M.rowlen[e][i], lj = M.rowlen[e][j];
#pragma unroll 1
              for (int a = 0; a < bs; ++a) {
                red_add_f64(pi + (long long)a * li + pij * bs + a, v);
                if (i != j) red_add_f64(pj + (long long)a * lj + pji * bs + a, v);
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < B; ++kk) {
          const int en = r * B + kk;
          if (en < T::NACC) {
            const int i = en / T::BS, a = en % T::BS;
            red_add_f64(P.out + (long long)M.dofs[i][e] * T::BS + a, acc[kk]);
          }
        }
      }
    }
  }
}

}  // namespace cq
