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

// ---- stage layout ------------------------------------------------------------------------------------
template <class T, int CB>
struct StageMeta {
  static constexpr int N = T::N, NV = T::NV, DIM = T::DIM;
  static constexpr int NPOS = round_up(N * N, 16);
  static constexpr bool MAT = T::RANK == 2;
  static constexpr bool VEC = T::RANK == 1;
  int nent;              // entries in this stage, -1: no more work
  int next;              // consumer claim counter
  int order[CB];         // entry indices sorted by trip count, longest first
  int nq[CB];
  int cell[CB];
  const double* pts[CB];  // generic pointers: into this stage's shared memory, or global on overflow
  const double* w[CB];
  const double* nrm[T::NEED_NORMAL ? CB : 1];
  double X[NV * DIM][CB];
  long long rowstart[MAT ? N : 1][CB];  // indptr[dof_i * bs]
  int rowlen[MAT ? N : 1][CB];          // length of that (blocked) row
  int dofs[VEC ? N : 1][CB];
  alignas(16) uint8_t pos[MAT ? CB : 1][MAT ? NPOS : 16];
  double coef[T::NEED_COEF ? T::NCOEF : 1][CB];
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
template <class T, int G, int CB>
CQ_DEVINL void staged_process_entry(const KParams& P, const StageMeta<T, CB>& M, const int e, const bool active, const int r,
                                    double& msum) {
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
#pragma unroll 2
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
            int i, j;
            tri_unpack<N>(en, i, j);
            const double v = acc[kk];
            const int pij = M.pos[e][i * N + j], pji = M.pos[e][j * N + i];
            const long long si = M.rowstart[i][e], sj = M.rowstart[j][e];
            if (bs == 1) {
              red_add_f64(P.values + si + pij, v);
              if (i != j) red_add_f64(P.values + sj + pji, v);
            } else {
              const int li = M.rowlen[i][e], lj = M.rowlen[j][e];
              for (int a = 0; a < bs; ++a) {
                red_add_f64(P.values + si + (long long)a * li + pij * bs + a, v);
                if (i != j) red_add_f64(P.values + sj + (long long)a * lj + pji * bs + a, v);
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

// ---- the kernel ------------------------------------------------------------------------------------------
template <class T, int G, int NPW, int NCW>
__global__ void __launch_bounds__((NPW + NCW) * 32, 1) assemble_staged_kernel(const KParams P) {
  constexpr int CB = 32;
  constexpr int DIM = T::DIM, N = T::N, NV = T::NV;
  using Meta = StageMeta<T, CB>;
  using Lay = StagedLayout<T, CB>;
  constexpr int CPW = 32 / G;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[NPW], empty_bar[NPW];
  const int PB = P.stage_points;
  const size_t stage_bytes = Lay::stage_bytes(PB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NPW; ++s) {
      mbar_init(&full_bar[s], 2);     // producer lane 0 twice: expect_tx at copy issue, arrive after the header
      mbar_init(&empty_bar[s], NCW);  // one elected lane per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t num_stages = (P.num_entries + CB - 1) / CB;
  double msum = 0.0;

  if (warp < NPW) {
    // =========================== producer warp: owns stage `warp` ===========================
    unsigned char* base = smem_raw + (size_t)warp * stage_bytes;
    Meta& M = *reinterpret_cast<Meta*>(base);
    double* s_pts = reinterpret_cast<double*>(base + Lay::meta_bytes);
    double* s_w = reinterpret_cast<double*>(base + Lay::meta_bytes + Lay::pts_bytes(PB));
    double* s_nrm = reinterpret_cast<double*>(base + Lay::meta_bytes + Lay::pts_bytes(PB) + Lay::w_bytes(PB));
    uint32_t parity = 0;
    for (int it = 0;; ++it) {
      if (it > 0) {
        mbar_wait(&empty_bar[warp], parity);
        parity ^= 1;
      }
      long long t = 0;
      if (lane == 0) t = atomicAdd(P.work_counter, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= num_stages) {
        if (lane == 0) {
          M.nent = -1;
          mbar_arrive(&full_bar[warp]);
          mbar_arrive(&full_bar[warp]);
        }
        break;
      }
      const int64_t k0 = t * CB;
      const int nent = (int)((P.num_entries - k0) < CB ? (P.num_entries - k0) : CB);
      const bool valid = lane < nent;
      const int64_t k = k0 + (valid ? lane : nent - 1);
      const int64_t o0 = __ldg(P.offsets + k), o1 = __ldg(P.offsets + k + 1);
      const int32_t cell = __ldg(P.cells + k);
      const int64_t P0 = __shfl_sync(0xffffffffu, o0, 0);
      const int64_t P1 = __shfl_sync(0xffffffffu, o1, nent - 1);
      const int64_t Pa = P0 & ~(int64_t)1;  // even start: 16-B aligned for 8 / 16 / 24-byte strides
      // ---- bulk copies of the staged prefix [Pa, Pa + cnt) --------------------------------------
      int64_t cnt = P1 - Pa;
      if (cnt > PB) cnt = PB & ~1;          // overflow: the tail entries read global memory directly
      int64_t cnt_copy = (cnt + 1) & ~(int64_t)1;
      int tail = 0;                          // one odd point at the very end of the arrays
      if (Pa + cnt_copy > P.num_points) { cnt_copy = cnt - 1; tail = 1; }
      if (lane == 0) {
        const uint32_t b_pts = (uint32_t)(cnt_copy * DIM * 8), b_w = (uint32_t)(cnt_copy * 8);
        mbar_arrive_expect_tx(&full_bar[warp], b_pts * (T::NEED_NORMAL ? 2u : 1u) + b_w);
        if (cnt_copy > 0) {
          bulk_g2s(s_pts, P.pts + Pa * DIM, b_pts, &full_bar[warp]);
          bulk_g2s(s_w, P.w + Pa, b_w, &full_bar[warp]);
          if constexpr (T::NEED_NORMAL) bulk_g2s(s_nrm, P.normals + Pa * DIM, b_pts, &full_bar[warp]);
        }
        if (tail) {
          const int64_t q = Pa + cnt - 1, l = cnt - 1;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            s_pts[l * DIM + d] = __ldg(P.pts + q * DIM + d);
            if constexpr (T::NEED_NORMAL) s_nrm[l * DIM + d] = __ldg(P.normals + q * DIM + d);
          }
          s_w[l] = __ldg(P.w + q);
        }
      }
      // ---- per-entry header, one lane per entry ------------------------------------------------
      const int nq = valid ? (int)(o1 - o0) : 0;
      const bool staged = (o1 - Pa) <= cnt;
      M.nq[lane] = nq;
      M.cell[lane] = cell;
      M.pts[lane] = staged ? s_pts + (o0 - Pa) * DIM : P.pts + o0 * DIM;
      M.w[lane] = staged ? s_w + (o0 - Pa) : P.w + o0;
      if constexpr (T::NEED_NORMAL) M.nrm[lane] = staged ? s_nrm + (o0 - Pa) * DIM : P.normals + o0 * DIM;
      if (valid) {
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
        int32_t dof[T::RANK >= 1 ? N : 1];
        if constexpr (T::RANK >= 1) {
#pragma unroll
          for (int i = 0; i < N; ++i) dof[i] = __ldg(P.dofmap + (int64_t)cell * N + i);
        }
        if constexpr (Meta::MAT) {
          const int4* src = reinterpret_cast<const int4*>(P.elem_pos + (int64_t)cell * Meta::NPOS);
          int4* dst = reinterpret_cast<int4*>(&M.pos[lane][0]);
#pragma unroll
          for (int i = 0; i < Meta::NPOS / 16; ++i) dst[i] = __ldg(src + i);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int d = 0; d < DIM; ++d) M.X[v * DIM + d][lane] = __ldg(P.coords + 3 * (int64_t)vid[v] + d);
        if constexpr (Meta::MAT) {
          const int bs = P.bs;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const long long rs = __ldg(P.indptr + (int64_t)dof[i] * bs);
            M.rowstart[i][lane] = rs;
            M.rowlen[i][lane] = bs == 1 ? 0 : (int)(__ldg(P.indptr + (int64_t)dof[i] * bs + 1) - rs);
          }
        }
        if constexpr (Meta::VEC) {
#pragma unroll
          for (int i = 0; i < N; ++i) M.dofs[i][lane] = dof[i];
        }
        if constexpr (T::NEED_COEF) {
          const double* c = P.coeffs + (int64_t)cell * T::NCOEF;
#pragma unroll
          for (int i = 0; i < T::NCOEF; ++i) M.coef[i][lane] = __ldg(c + i);
        }
      }
      // ---- sort entries by trip count, longest first (bitonic over the 32 lanes) ----------------
      int key = (((nq + G - 1) / G) << 5) | (31 - lane);  // ties: lower lane first
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride >= 1; stride >>= 1) {
          const int other = __shfl_xor_sync(0xffffffffu, key, stride);
          const bool desc = (lane & size) == 0;  // final order: descending across lanes
          const bool lower = (lane & stride) == 0;
          const int mx = key > other ? key : other, mn = key > other ? other : key;
          key = (desc == lower) ? mx : mn;
        }
      M.order[lane] = 31 - (key & 31);
      if (lane == 0) {
        M.nent = nent;
        M.next = 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[warp]);  // release: header + tail stores are visible
    }
  } else {
    // =========================== consumer warps ===========================================
    const int r = lane & (G - 1), grp = lane / G;
    uint32_t ended = 0, parity = 0;
    int s = 0;
    while (ended != (1u << NPW) - 1u) {
      if (!((ended >> s) & 1u)) {
        mbar_wait(&full_bar[s], parity);
        Meta& M = *reinterpret_cast<Meta*>(smem_raw + (size_t)s * stage_bytes);
        const int nent = M.nent;
        if (nent < 0) {
          ended |= 1u << s;
        } else {
          for (;;) {
            int e0 = 0;
            if (lane == 0) e0 = atomicAdd(&M.next, CPW);
            e0 = __shfl_sync(0xffffffffu, e0, 0);
            if (e0 >= nent) break;
            const bool active = e0 + grp < nent;
            const int e = active ? M.order[e0 + grp] : 0;
            staged_process_entry<T, G, CB>(P, M, e, active, r, msum);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
      }
      if (++s == NPW) {
        s = 0;
        parity ^= 1;
      }
    }
  }
  if constexpr (T::RANK == 0) block_sum_atomic(msum, P.out);
}

}  // namespace cq
