// cq_staged_kernel.cuh -- the staged persistent kernel (producer / consumer roles); see cq_staged.cuh
#pragma once
#include "cq_dmma.cuh"
#include "cq_moment.cuh"
#include "cq_staged.cuh"
#include "cq_zero.cuh"

#ifndef CQ_L2_PREFETCH  // 1: producers prefetch the rule slice of a stage into L2 one pipeline level before its gather (measured neutral, profiles/r2/l2_prefetch.txt: the ring is not bound by the copy latency)
#define CQ_L2_PREFETCH 0
#endif
#ifndef CQ_POS_BULK_MIN  // scatter maps of at least this many bytes per cell travel by cp.async.bulk
#define CQ_POS_BULK_MIN 64
#endif

namespace cq {

// ---- producer role (one warp per stage slot) ----------------------------------------------------------
// KC_ONLY (moment kernels): the consumers need the axis-aligned cell constants only, so the producer
// gathers DIM + 1 vertex coordinates per cell instead of all of them and does not publish M.X.
template <class T, int G, int CB, bool KC_ONLY = false>
CQ_DEVINL void staged_producer(const KParams& P, unsigned char* base, uint64_t* full_bar_w, uint64_t* empty_bar_w, const int lane,
                               const int PB, const int64_t num_stages, const int slot, const int nslots, const int sub = 0,
                               const int npp = 1, uint64_t* turn_bar_w = nullptr) {
  constexpr int DIM = T::DIM, N = T::N, NV = T::NV;
  using Meta = StageMeta<T, CB>;
  using Lay = StagedLayout<T, CB>;
  {
    // =========================== producer warp: owns stage `warp` ===========================
    // Filling a stage is a chain of dependent global loads (work counter -> offsets, cell ids ->
    // dofmap rows -> vertex coordinates, CSR row starts): ~4 round trips.  The chain is software-
    // pipelined over three stages held in registers -- while stage A is being published, level 3
    // of stage B, level 2 of stage C and level 1 of stage D are in flight -- so a free slot only
    // waits for its bulk copy.
    Meta& M = *reinterpret_cast<Meta*>(base);
    double* s_pts = reinterpret_cast<double*>(base + Lay::meta_bytes);
    double* s_w = reinterpret_cast<double*>(base + Lay::meta_bytes + Lay::pts_bytes(PB));
    double* s_nrm = reinterpret_cast<double*>(base + Lay::meta_bytes + Lay::pts_bytes(PB) + Lay::w_bytes(PB));
    constexpr bool POS_BULK = Meta::MAT && Meta::NPOS >= CQ_POS_BULK_MIN;  // larger maps: bulk copy instead of registers
    constexpr int NP16 = (Meta::MAT && !POS_BULK) ? Meta::NPOS / 16 : 1;
    struct Regs {
      long long t;  // stage index; -2: pipeline bubble, -1: no more work
      long long k0; // list position of the stage's first entry
      int nent;
      bool valid;
      int64_t o0, o1;
      int32_t cell;
      int32_t vid[NV];
      int32_t dof[T::RANK >= 1 ? N : 1];
      int4 pos[NP16];
      double X[KC_ONLY ? 2 * DIM : NV * DIM];  // KC_ONLY: vertex 0, then coordinate d of its +d neighbour
      long long rs[Meta::MAT ? N : 1];
      int rl[Meta::MAT ? N : 1];
    };
    Regs ra, rb, rc;
#ifndef CQ_NO_EVICT_FIRST
    const uint64_t pol = l2_policy_evict_first();
#define CQ_BULK_ONCE(dst, src, bytes, bar) bulk_g2s_once(dst, src, bytes, bar, pol)
#else
#define CQ_BULK_ONCE(dst, src, bytes, bar) bulk_g2s(dst, src, bytes, bar)
#endif
    ra.t = rb.t = rc.t = -2;
    ra.valid = rb.valid = rc.valid = false;
    const bool scat = P.elem_scalar == nullptr;  // owner-computes path: no scatter, so no row starts / scatter map
    // Stages are dealt out statically, interleaved over the CTAs and their slots: with hundreds of
    // stages per CTA the ragged rules average out (a dynamic ticket costs one more dependent
    // round trip per stage in this latency-bound warp).
    // With npp > 1 producer warps per slot, warp `sub` fills the slot's rounds sub, sub + npp, ...: its
    // gather chain overlaps the other warps' rounds.  Round r of a slot may be written once round
    // r - 1 has been published (turn barrier) and consumed (empty barrier).
    const int64_t round_t = (int64_t)nslots * gridDim.x;
    int64_t next_t = ((int64_t)sub * nslots + slot) * gridDim.x + blockIdx.x;
    const int64_t step_t = round_t * npp;
    int round = sub;
    // P.dynamic (long lists; npp == 1): the stages come from a global ticket counter instead, drawn one iteration
    // ahead so that the atomic's round trip hides in the pipeline.  The static interleaving lets the CTAs drift apart
    // -- nothing ties a fast SM to a slow one, and over the ~2400 rounds of a 6.7e7-cell list (config 5) the drift
    // reaches many rounds -- so that rows shared by neighbouring cells are long evicted from the L2 when the second
    // contribution arrives: ncu counted 3.3 DRAM reads and 2.4 writes per matrix line.  With tickets the stages in
    // flight are always ~grid * slots consecutive ones.
    const bool dyn = P.dynamic != 0;
    unsigned int ticket = 0;
    if (dyn && lane == 0) ticket = atomicAdd(P.work_counter, 1u);
    for (;;) {
      // ---------------- publish stage A ----------------
      if (ra.t == -3) break;  // beyond the end marker (another warp of this slot publishes it)
      if (ra.t == -1) {
        if (round >= 1) {
          if (npp > 1) mbar_wait(turn_bar_w, (uint32_t)(round - 1) & 1u);
          mbar_wait(empty_bar_w, (uint32_t)(round - 1) & 1u);
        }
        if (lane == 0) {
          M.nent = -1;
          mbar_arrive(full_bar_w);
          mbar_arrive(full_bar_w);
        }
        break;
      }
      if (ra.t >= 0) {
        if (round >= 1) {
          if (npp > 1) mbar_wait(turn_bar_w, (uint32_t)(round - 1) & 1u);
          mbar_wait(empty_bar_w, (uint32_t)(round - 1) & 1u);
        }
        round += npp;
        const int nent = ra.nent;
        const int64_t o0 = ra.o0, o1 = ra.o1;
        const int64_t P0 = __shfl_sync(0xffffffffu, o0, 0);
        const int64_t P1 = __shfl_sync(0xffffffffu, o1, nent - 1);
        const int64_t Pa = P0 & ~(int64_t)1;  // even start: 16-B aligned for 8 / 16 / 24-byte strides
        int64_t cnt = P1 - Pa;
        if (cnt > PB) cnt = PB & ~1;          // overflow: the tail entries read global memory directly
        int64_t cnt_copy = (cnt + 1) & ~(int64_t)1;
        int tail = 0;                          // one odd point at the very end of the arrays
        if (Pa + cnt_copy > P.num_points) { cnt_copy = cnt - 1; tail = 1; }
        if (lane == 0) {
          const uint32_t b_pts = (uint32_t)(cnt_copy * DIM * 8), b_w = (uint32_t)(cnt_copy * 8);
          mbar_arrive_expect_tx(full_bar_w, b_pts * (T::NEED_NORMAL ? 2u : 1u) + b_w + (POS_BULK && scat ? (uint32_t)(nent * Meta::NPOS) : 0u));
          if (cnt_copy > 0) {
            CQ_BULK_ONCE(s_pts, P.pts + Pa * DIM, b_pts, full_bar_w);
            CQ_BULK_ONCE(s_w, P.w + Pa, b_w, full_bar_w);
            if constexpr (T::NEED_NORMAL) CQ_BULK_ONCE(s_nrm, P.normals + Pa * DIM, b_pts, full_bar_w);
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
        const int nq = ra.valid ? (int)(o1 - o0) : 0;
        const bool staged = (o1 - Pa) <= cnt;
        M.nq[lane] = nq;
        M.cell[lane] = ra.cell;
        M.pts[lane] = staged ? s_pts + (o0 - Pa) * DIM : P.pts + o0 * DIM;
        M.w[lane] = staged ? s_w + (o0 - Pa) : P.w + o0;
        if constexpr (T::NEED_NORMAL) M.nrm[lane] = staged ? s_nrm + (o0 - Pa) * DIM : P.normals + o0 * DIM;
        if constexpr (POS_BULK) {
          // consecutive cells (the usual case): the 32 map rows are one contiguous slice -> one bulk copy.
          // Keep this at warp-uniform control flow: nesting the per-lane copy loop (UBLKCP is issued
          // through uniform registers, one lane at a time) under a second divergent branch hung the warp.
#ifndef CQ_NO_POS_FASTPATH
          const int cell0 = __shfl_sync(0xffffffffu, ra.cell, 0);
          const bool consec = __all_sync(0xffffffffu, ra.valid && ra.cell == cell0 + lane);
#else
          const bool consec = false;
#endif
          if (!scat) {
          } else if (consec) {
            if (lane == 0) CQ_BULK_ONCE(&M.pos[0][0], P.elem_pos + (int64_t)cell0 * Meta::NPOS, CB * Meta::NPOS, full_bar_w);
          } else if (ra.valid) {
            CQ_BULK_ONCE(&M.pos[lane][0], P.elem_pos + (int64_t)ra.cell * Meta::NPOS, Meta::NPOS, full_bar_w);
          }
        }
        if (ra.valid) {
          if constexpr (Meta::MAT) {
            if constexpr (!POS_BULK) {
              int4* dst = reinterpret_cast<int4*>(&M.pos[lane][0]);
#pragma unroll
              for (int i = 0; i < NP16; ++i) dst[i] = ra.pos[i];
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
              M.rowstart[lane][i] = ra.rs[i];
              M.rowlen[lane][i] = ra.rl[i];
            }
          }
          if constexpr (!KC_ONLY) {
#pragma unroll
            for (int i = 0; i < NV * DIM; ++i) M.X[i][lane] = ra.X[i];
          }
          if constexpr (Meta::MOM) {
            // axis-aligned cells: J = diag(J_dd), J_dd = x_d(+d neighbour) - x_d(vertex 0); the only divisions of
            // the moment kernels.  +d neighbour of local vertex 0: quadrilateral (Lattice<2,1>) vertex 2 is +x,
            // vertex 1 is +y; hexahedron (Lattice<3,1>) vertex 1 << d.
            double Jd[DIM], adet = 1.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
              const int nb = DIM == 2 ? (d == 0 ? 2 : 1) : (1 << d);
              Jd[d] = (KC_ONLY ? ra.X[DIM + d] : ra.X[nb * DIM + d]) - ra.X[d];
              adet *= Jd[d];
            }
            adet = fabs(adet);
            M.kc[0][lane] = adet;
            if constexpr (T::FORM_ID == CQ_FORM_STIFFNESS) {
#pragma unroll
              for (int d = 0; d < DIM; ++d) M.kc[1 + d][lane] = adet / (Jd[d] * Jd[d]);
            }
            if constexpr (Meta::NIT) {
              double h2 = 0.0;  // ufl.CellDiameter of a box: its diagonal
#pragma unroll
              for (int d = 0; d < DIM; ++d) {
                M.kc[1 + d][lane] = adet / Jd[d];
                h2 = fma(Jd[d], Jd[d], h2);
              }
              M.kc[1 + DIM][lane] = adet * P.betaN / sqrt(h2);
            }
          }
          if constexpr (Meta::VEC) {
#pragma unroll
            for (int i = 0; i < N; ++i) M.dofs[i][lane] = ra.dof[i];
          }
          if constexpr (T::NEED_COEF) {
            const double* c = P.coeffs + (int64_t)ra.cell * T::NCOEF;
#pragma unroll
            for (int i = 0; i < T::NCOEF; ++i) M.coef[i][lane] = __ldg(c + i);
          }
        }
        // sort entries by trip count, longest first (bitonic over the 32 lanes); nothing to sort when every entry has
        // the same number of points (a standard rule tiled over the cells: the bulk of configs 4 and 5)
        int key = (((nq + G - 1) / G) << 5) | (31 - lane);  // ties: lower lane first
        if (!__all_sync(0xffffffffu, nq == __shfl_sync(0xffffffffu, nq, 0))) {
#pragma unroll
          for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
            for (int stride = size >> 1; stride >= 1; stride >>= 1) {
              const int other = __shfl_xor_sync(0xffffffffu, key, stride);
              const bool desc = (lane & size) == 0;
              const bool lower = (lane & stride) == 0;
              const int mx = key > other ? key : other, mn = key > other ? other : key;
              key = (desc == lower) ? mx : mn;
            }
        }
        M.order[lane] = 31 - (key & 31);
        if (lane == 0) {
          M.nent = nent;
          M.next = 0;
          M.k0 = ra.k0;
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(full_bar_w);  // release: header + tail stores are visible
          if (npp > 1) mbar_arrive(turn_bar_w);
        }
      }
      // ---------------- advance the pipeline: A <- B (+ level 3), B <- C (+ level 2), C <- D (+ level 1) ----
      ra = rb;
      if (ra.t >= 0 && ra.valid) {
        if constexpr (!KC_ONLY) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int d = 0; d < DIM; ++d) ra.X[v * DIM + d] = __ldg(P.coords + 3 * (int64_t)ra.vid[v] + d);
        } else {
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            const int nb = DIM == 2 ? (d == 0 ? 2 : 1) : (1 << d);
            ra.X[d] = __ldg(P.coords + 3 * (int64_t)ra.vid[0] + d);
            ra.X[DIM + d] = __ldg(P.coords + 3 * (int64_t)ra.vid[nb] + d);
          }
        }
        if constexpr (Meta::MAT) {
          const int bs = P.bs;
          if (scat) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              // rows owned by the neighbour rank: slot table of ITS matrix (peer memory over NVLink)
              const long long row = (long long)ra.dof[i] * bs;
              const long long* ip = row >= P.peer_row0 ? P.peer_rowstart + (row - P.peer_row0) : reinterpret_cast<const long long*>(P.indptr) + row;
              const long long rs = __ldg(ip);
              ra.rs[i] = rs;
              ra.rl[i] = bs == 1 ? 0 : (int)(__ldg(ip + 1) - rs);
            }
          }
        }
      }
      rb = rc;
#if CQ_L2_PREFETCH
      if (rb.t >= 0) {
        // the slice this stage's bulk copies will read (same arithmetic as at publish time, rounded inwards)
        const int64_t P0 = __shfl_sync(0xffffffffu, rb.o0, 0);
        const int64_t P1 = __shfl_sync(0xffffffffu, rb.o1, rb.nent - 1);
        const int64_t Pa = P0 & ~(int64_t)1;
        int64_t cnt = (P1 - Pa) & ~(int64_t)1;
        if (cnt > PB) cnt = PB & ~1;
        bool pos_consec = false;
        if constexpr (POS_BULK) pos_consec = __all_sync(0xffffffffu, rb.valid && rb.cell == __shfl_sync(0xffffffffu, rb.cell, 0) + lane);
        if (lane == 0) {  // one lane, one branch level (bulk instructions issue through uniform registers)
          if (cnt > 0) {
            bulk_prefetch_l2(P.pts + Pa * DIM, (uint32_t)(cnt * DIM * 8));
            bulk_prefetch_l2(P.w + Pa, (uint32_t)(cnt * 8));
            if constexpr (T::NEED_NORMAL) bulk_prefetch_l2(P.normals + Pa * DIM, (uint32_t)(cnt * DIM * 8));
          }
          if constexpr (POS_BULK)
            if (pos_consec) bulk_prefetch_l2(P.elem_pos + (int64_t)rb.cell * Meta::NPOS, CB * Meta::NPOS);
        }
      }
#endif
      if (rb.t >= 0 && rb.valid) {
        const int32_t cell = rb.cell;
        if constexpr (DIM == 2) {
          const int4 v = __ldg(reinterpret_cast<const int4*>(P.geom_dofmap) + cell);
          rb.vid[0] = v.x; rb.vid[1] = v.y; rb.vid[2] = v.z; rb.vid[3] = v.w;
        } else {
          const int4 v0 = __ldg(reinterpret_cast<const int4*>(P.geom_dofmap) + 2 * (int64_t)cell);
          const int4 v1 = __ldg(reinterpret_cast<const int4*>(P.geom_dofmap) + 2 * (int64_t)cell + 1);
          rb.vid[0] = v0.x; rb.vid[1] = v0.y; rb.vid[2] = v0.z; rb.vid[3] = v0.w;
          rb.vid[4] = v1.x; rb.vid[5] = v1.y; rb.vid[6] = v1.z; rb.vid[7] = v1.w;
        }
        if constexpr (T::RANK >= 1) {
#pragma unroll
          for (int i = 0; i < N; ++i) rb.dof[i] = __ldg(P.dofmap + (int64_t)cell * N + i);
        }
        if constexpr (Meta::MAT && !POS_BULK) {
          const int4* src = reinterpret_cast<const int4*>(P.elem_pos + (int64_t)cell * Meta::NPOS);
#pragma unroll
          for (int i = 0; i < NP16; ++i) rb.pos[i] = __ldg(src + i);
        }
      }
      long long td = next_t < num_stages ? next_t : (next_t - round_t < num_stages ? -1 : -3);
      next_t += step_t;
      if (dyn) {
        const unsigned int tk = __shfl_sync(0xffffffffu, ticket, 0);
        const bool live = rb.t != -1 && rc.t != -1;  // no ticket after this warp's end marker
        td = !live ? -3 : ((long long)tk < num_stages ? (long long)tk : -1);
        if (td >= 0 && lane == 0) ticket = atomicAdd(P.work_counter, 1u);  // the next one, consumed one iteration from now
      }
      rc.t = td;
      rc.valid = false;
      if (td >= 0) {
        const int64_t k0 = stage_of(P, td, num_stages) * CB;  // td: position in the traversal, see stage_of
        const int nent = (int)((P.num_entries - k0) < CB ? (P.num_entries - k0) : CB);
        rc.nent = nent;
        rc.k0 = k0;
        rc.valid = lane < nent;
        const int64_t k = k0 + (rc.valid ? lane : nent - 1);
        rc.o0 = __ldg(P.offsets + k);
        rc.o1 = __ldg(P.offsets + k + 1);
        rc.cell = __ldg(P.cells + k);
      }
    }
  }
}

// ---- the kernel ------------------------------------------------------------------------------------------
// MODE 0: G lanes per entry accumulate the element tensor point by point (any form, any cell).
// MODE 1: G lanes per entry, monomial moments (cq_moment.cuh).
// MODE 2: one warp per entry, FP64 tensor-core contraction for the Q2 hexahedron (cq_dmma.cuh; G = 32).
template <class T, int G, int NPW, int NCW, int MODE>
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
    staged_producer<T, G, CB, false>(P, smem_raw + (size_t)warp * stage_bytes, &full_bar[warp], &empty_bar[warp], lane, PB, num_stages, warp, NPW);
  } else {
    // =========================== consumer warps ===========================================
    const int r = lane & (G - 1), grp = lane / G;
    LaneEntries<T, G> LE;
    LE.init(r);
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
          for (; CQ_DBG(P) != 1;) {
            int e0 = 0;
            if (lane == 0) e0 = atomicAdd(&M.next, CPW);
            e0 = __shfl_sync(0xffffffffu, e0, 0);
            if (e0 >= nent) break;
            const bool active = e0 + grp < nent;
            const int e = active ? M.order[e0 + grp] : 0;
            if constexpr (MODE == 2) dmma_process_entry<T, CB>(P, M, e, reinterpret_cast<double*>(smem_raw + (size_t)NPW * stage_bytes + (size_t)(warp - NPW) * DMMA_TILE_BYTES));
            else if constexpr (MODE == 1) moment_process_entry<T, G, CB>(P, M, e, active, r);
            else staged_process_entry<T, G, CB>(P, M, e, active, r, msum, LE);
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



// ---------------------------------------------------------------------------------------------
// Three-role variant for the moment path (2-D mass / stiffness on axis-aligned cells):
//   producer warps (as above)  ->  moment warps: point loops, moments into the stage header
//                              ->  expansion warps: moments -> n x n values -> REDs
// The expansion is almost pure LSU work (one RED per value, ~1.15 clk per lane per SM measured,
// profiles/microbench/red_bench.cu) while the point loops are FP64 / latency work; giving them to
// different warps lets the two overlap instead of alternating inside one warp.  Each expansion warp
// owns a fixed subset of the x-pairs (compile-time code per role) and handles all 32 entries of a
// stage with one thread per entry.
// ---------------------------------------------------------------------------------------------
// XT: row-owned expansion with the transposed scatter (RowTile, cq_moment.cuh); NXW = sets * (DEG + 1)
// NPP: producer warps per stage slot
// ZW: 1 = one more warp that zeroes the matrix ahead of the scatter (fused zeroEntries, cq_zero.cuh)
// 1: the wait for the zeroed ranges sits in the first moment warp (which has slack) instead of the expansion warps (which
// are the critical path of a one-set kernel): zero + Q2 stiffness 1.89 -> 1.77 ms, zero + Q2 mass 1.59 -> 1.54 ms
// (profiles/r2s3/call10_zero_wait_in_moment_warp.log)
#ifndef CQ_ZWAIT_IN_MOMENT
#define CQ_ZWAIT_IN_MOMENT 1
#endif
template <class T, int G, int NPW, int NMW, int NXW, bool XT = false, int NPP = 1, int ZW = 0>
__global__ void __launch_bounds__((NPW * NPP + NMW + NXW + ZW) * 32, 1) assemble_moment_split_kernel(const KParams P) {
  constexpr int CB = 32;
  using Meta = StageMeta<T, CB>;
  using Lay = StagedLayout<T, CB>;
  constexpr int CPW = 32 / G;
  constexpr int Pn = 2 * T::DEG + 1;
  static_assert(Meta::MOM, "moment kernel: mass / stiffness on quadrilaterals (degree 1, 2) and hexahedra (degree 1)");
  static_assert(XT || T::DIM == 2, "hexahedra: row-owned expansion only");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[NPW], mom_bar[NPW], empty_bar[NPW], turn_bar[NPW];
  __shared__ volatile long long zero_progress;  // the wave the moment phase of this CTA has reached (paces the zeroing warp)
  const int PB = P.stage_points;
  const size_t stage_bytes = Lay::stage_bytes(PB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NPW; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&mom_bar[s], NMW);
      mbar_init(&turn_bar[s], 1);
      mbar_init(&empty_bar[s], XT ? T::DEG + 1 : NXW);  // the expansion warps of one set
    }
    zero_progress = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t num_stages = (P.num_entries + CB - 1) / CB;

  if (ZW && warp == NPW * NPP + NMW + NXW) {
    // ------------------------------ zeroing warp (fused zeroEntries, cq_zero.cuh) ------------------------------
    // Off every critical path: zeroes this CTA's share of the ranges of ZB waves at a time, zero_ahead waves ahead
    // of the CTA's moment phase, then one fence and the ZB signals.
    if (P.zero_hi == nullptr) return;
    constexpr int ZB = 4;
    const long long waves = (num_stages + gridDim.x - 1) / gridDim.x;
    const int zahead = P.zero_ahead;
    zero_share_keep(P.values, __ldg(P.zero_hi + waves - 1), P.nnz, P.zero_keep_lo, P.zero_keep_hi, blockIdx.x, gridDim.x, lane);  // rows no entry touches
    long long hi_prev = 0;
    for (long long kz0 = 0; kz0 < waves; kz0 += ZB) {
      const long long kz1 = kz0 + ZB < waves ? kz0 + ZB : waves;
      const long long hi = __ldg(P.zero_hi + kz1 - 1);  // before the pacing wait: off the critical path
      while (zero_progress < kz1 - 1 - zahead) __nanosleep(128);
      zero_share_keep(P.values, hi_prev, hi, P.zero_keep_lo, P.zero_keep_hi, blockIdx.x, gridDim.x, lane);
      hi_prev = hi;
      __threadfence();  // the zero stores are performed (gpu scope) before the signals
      __syncwarp();
      if (lane < kz1 - kz0) red_relaxed_inc_u32(P.zero_done + kz0 + lane);
    }
  } else if (warp < NPW * NPP) {
    const int slot = warp % NPW, sub = warp / NPW;
    // short gather (DIM + 1 coordinates per cell) where it pays: measured, Q1 quadrilaterals are faster with the
    // full gather (0.78 vs 0.86-0.94 ms), Q2 quadrilaterals (1.42 vs 1.44) and hexahedra (0.55 vs 0.78) with the short one
    staged_producer<T, (Meta::NIT ? 1 : G), CB, (T::DIM == 3 || T::DEG == 2 || Meta::NIT)>(P, smem_raw + (size_t)slot * stage_bytes, &full_bar[slot], &empty_bar[slot], lane, PB, num_stages, slot, NPW, sub,
                              NPP, &turn_bar[slot]);
  } else if (warp < NPW * NPP + NMW) {
    // ------------------------------ moment warps ------------------------------
    const int r = lane & (G - 1), grp = lane / G;
    uint32_t ended = 0, parity = 0;
    int s = 0;
    const bool zlead = ZW && P.zero_hi != nullptr && warp == NPW * NPP && lane == 0;  // reports the progress to the zeroing warp
    long long rnd = 0;
    while (ended != (1u << NPW) - 1u) {
      if (!((ended >> s) & 1u)) {
        if (zlead) zero_progress = rnd * NPW + s;
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
            moment_accumulate_store<T, G, CB, (NMW >= 9)>(M, e, active, CQ_DBG(P) != 1, r);  // long-rule shapes: more points in flight
          }
        }
        __syncwarp();
#if CQ_ZWAIT_IN_MOMENT
        // fused zeroEntries: the first moment warp holds its arrival until every CTA has zeroed this wave's ranges
        // (the CTAs' zeroing warps go through the waves in order, so the counter of this wave covers the earlier ones);
        // the expansion warps then scatter without looking at the counters
        if (ZW && P.zero_hi != nullptr && warp == NPW * NPP && nent >= 0) {
          if (lane == 0) {
            const unsigned* zd = P.zero_done + (rnd * NPW + s);
            while (ld_relaxed_u32(zd) < gridDim.x) __nanosleep(64);
          }
          __syncwarp();
        }
#endif
        if (lane == 0) mbar_arrive(&mom_bar[s]);  // also passes the end marker on
      }
      if (++s == NPW) {
        s = 0;
        parity ^= 1;
        ++rnd;
      }
    }
    if (zlead) zero_progress = 0x7fffffffffffffffLL;
  } else {
    // ------------------------------ expansion warps ------------------------------
    // XT: warp = (set, role); role = x lattice index of the rows it owns, set k serves the stage
    // slots s with s % NXS == k (NXS = NXW / (DEG + 1) sets work on different stages concurrently)
    constexpr int NROLE = XT ? T::DEG + 1 : NXW;
    constexpr int NXS = NXW / NROLE;
    static_assert(NXW % NROLE == 0 && NPW % NXS == 0, "expansion warps: whole sets, slots divisible by sets");
    const int xw = warp - NPW * NPP - NMW;
    const int role = xw % NROLE, set = xw / NROLE;
    uint32_t mine = 0;
#pragma unroll
    for (int q = 0; q < NPW; ++q)
      if (q % NXS == set) mine |= 1u << q;
    uint32_t ended = 0, parity = 0;
    int s = 0;
    // fused zeroEntries (cq_zero.cuh): the zeroing warp zeroes the ranges zero_ahead waves ahead; here, before the
    // scatter of wave = rnd * NPW + s, every range up to this wave's must be complete on all CTAs.  This set saw the
    // waves <= wave - NXS complete in its previous iteration, so it looks at the counters of the last NXS waves; they
    // are loaded one iteration ahead (normally complete long ago: the load is off the critical path).
    const bool fz = ZW && P.zero_hi != nullptr && !CQ_ZWAIT_IN_MOMENT;  // ZW == 0: all of this folds away
    const long long waves = (num_stages + gridDim.x - 1) / gridDim.x;
    const unsigned zexpect = gridDim.x;
    unsigned zs[NXS];
    auto zload = [&](const long long w) {
#pragma unroll
      for (int q = 0; q < NXS; ++q) {
        const long long j = w - q;
        zs[q] = (fz && lane == 0 && j >= 0 && j < waves) ? ld_relaxed_u32(P.zero_done + j) : zexpect;
      }
    };
    auto zwait = [&](const long long w) {
#pragma unroll
      for (int q = 0; q < NXS; ++q)
        while (zs[q] < zexpect) zs[q] = ld_relaxed_u32(P.zero_done + (w - q));
      __syncwarp();
    };
    long long rnd = 0;
    zload(set);
    while (ended != mine) {
      if (((mine >> s) & 1u) && !((ended >> s) & 1u)) {
        const long long wave = rnd * NPW + s;
        mbar_wait(&mom_bar[s], parity);
        Meta& M = *reinterpret_cast<Meta*>(smem_raw + (size_t)s * stage_bytes);
        const int nent = M.nent;
        if (nent < 0) {
          ended |= 1u << s;
        } else {
          if constexpr (XT) {
            // this warp's tile lives in the stage's point buffer (dead after the moment phase)
            using RT = RowTile<T::DIM, T::DEG>;
            double* tile = reinterpret_cast<double*>(smem_raw + (size_t)s * stage_bytes + Lay::meta_bytes + (size_t)role * RT::bytes(CB));
            if (CQ_DBG(P) != 2) {
              static_for<0, T::DEG + 1>([&](auto RR) {
                if (role == decltype(RR)::value) {
                  if (lane < nent) moment_expand_rows<T, CB, decltype(RR)::value>(M, lane, tile + lane * RT::STRIDE);
                  if (fz) {
                    zwait(wave);
                    zload(wave + NXS);
                  }
                  __syncwarp();
                  scatter_rows_transposed<T, CB, decltype(RR)::value>(P, M, tile, nent, lane);
                }
              });
            }
          } else if constexpr (T::DIM == 2) {
           if (fz) {
             zwait(wave);
             zload(wave + NXS);
           }
           if (lane < nent && CQ_DBG(P) != 2) {
            double mu[Pn][Pn];
#pragma unroll
            for (int p = 0; p < Pn; ++p)
#pragma unroll
              for (int q = 0; q < Pn; ++q) mu[p][q] = M.mu[lane][p * Pn + q];
            static_for<0, NXW>([&](auto XW) {
              if (xw == decltype(XW)::value) moment_expand<T, 1, CB, decltype(XW)::value, NXW>(P, M, lane, 0, mu);
            });
           }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
      }
      if (++s == NPW) {
        s = 0;
        parity ^= 1;
        ++rnd;
      }
    }
  }
}

}  // namespace cq
