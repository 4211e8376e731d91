// cq_launch.cuh -- launch helpers shared by the cq_launch_*.cu translation units: the direct
// (grid-strided) kernel and the bulk-copy staged persistent kernel, and the choice between them.
#pragma once
#include <cstdlib>

#include "cq_internal.h"
#include "cq_staged_kernel.cuh"

namespace cq {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Tuning / experiment switches read from the environment exist in profiling builds only (-DCQ_PROFILING,
// profiles/build_variant*.sh): the shipped library has compile-time constants, so no environment variable can change
// which code runs or skip work.
#ifdef CQ_PROFILING
static inline int tune_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
static inline double tune_f64(const char* name, double dflt) { const char* v = getenv(name); return v ? atof(v) : dflt; }
#define CQ_TUNE_INT(name, dflt) tune_int(name, dflt)
#define CQ_TUNE_F64(name, dflt) tune_f64(name, dflt)
#else
#define CQ_TUNE_INT(name, dflt) (dflt)
#define CQ_TUNE_F64(name, dflt) (dflt)
#endif

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device and per kernel: remembered in the ctx (one per device)
static int ensure_smem_optin(cq_ctx* ctx, const void* kernel, size_t smem) {
  size_t& have = ctx->smem_optin[kernel];
  if (smem <= have) return CQ_OK;
  if (int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(smem)")) return rc;
  have = smem;
  return CQ_OK;
}

#ifndef CQ_DYNAMIC_ROUNDS
#define CQ_DYNAMIC_ROUNDS 512
#endif
// ---- fused zeroEntries (cq_zero.cuh) ----------------------------------------------------------------
// Sets P.zero_hi / zero_done / zero_ahead for a launch of `grid` CTAs (stream-ordered).  The caller has checked that
// the kernel supports it; on return P.zero_req is cleared.
static int prepare_fused_zero(cq_ctx* ctx, KParams& P, const int n, const int grid, cudaStream_t s) {
  P.tile_sp = P.tile_sg = 0;  // the zeroing front follows the list order
  const int64_t num_stages = (P.num_entries + 31) / 32;
  const int64_t waves = (num_stages + grid - 1) / grid;
  if ((size_t)waves > ctx->zero_waves) {
    if (ctx->zero_buf) cudaFree(ctx->zero_buf);
    ctx->zero_buf = nullptr;
    ctx->zero_waves = 0;
    const size_t cap = (size_t)waves + (size_t)waves / 4 + 64;
    if (int rc = check_cuda(cudaMalloc(&ctx->zero_buf, cap * 12), "cudaMalloc(zero ranges)")) return rc;
    ctx->zero_waves = cap;
  }
  unsigned long long* raw = static_cast<unsigned long long*>(ctx->zero_buf);
  unsigned int* done = reinterpret_cast<unsigned int*>(raw + ctx->zero_waves);
  if (int rc = check_cuda(cudaMemsetAsync(done, 0, (size_t)waves * 4, s), "memset(zero_done)")) return rc;
  const int64_t blocks = waves < (int64_t)ctx->num_sms * 16 ? waves : (int64_t)ctx->num_sms * 16;
  wave_hi_kernel<><<<(int)blocks, 256, 0, s>>>(P.cells, P.num_entries, P.dofmap, n, P.bs, P.indptr, P.nnz, grid, raw);
  wave_scan_kernel<><<<1, 1024, 0, s>>>(raw, waves);
  ctx->launches += 2;
  if (int rc = check_cuda(cudaGetLastError(), "fused-zero setup kernels")) return rc;
  const int ahead = CQ_TUNE_INT("CQ_ZERO_AHEAD", 6);
  P.zero_hi = reinterpret_cast<const long long*>(raw);
  P.zero_done = done;
  P.zero_ahead = ahead < 4 ? 4 : ahead;  // at least the batch of the zeroing warp (no deadlock: cq_staged_kernel.cuh)
  P.zero_req = 0;
  return CQ_OK;
}

// Stage distribution of a persistent launch: dynamic tickets for long lists (more than CQ_DYNAMIC_ROUNDS rounds of
// grid * slots stages; see staged_producer), the static interleaving otherwise (one dependent round trip less per stage).
// heavy: stages that take so long (Q2 hexahedra: a 27 x 27 tensor-core contraction and 729 bs scattered values per
// entry) that the ticket's round trip is noise -- always dynamic (measured on config 4, 128^3: 68.7 -> 61.7 ms).
static int choose_dynamic(cq_ctx* ctx, KParams& P, const int64_t num_stages, const int64_t per_round, cudaStream_t s, const bool heavy = false) {
  const int mode = CQ_TUNE_INT("CQ_DYNAMIC", -1);  // profiling builds: 0 / 1 force
  P.dynamic = mode >= 0 ? mode : ((heavy && num_stages > per_round) || num_stages > (int64_t)CQ_DYNAMIC_ROUNDS * per_round ? 1 : 0);
  if (!P.dynamic) return CQ_OK;
  return check_cuda(cudaMemsetAsync(ctx->work_counter, 0, sizeof(unsigned int), s), "memset(work_counter)");
}

// kernels without the fused zeroing: a memset does it
static int prezero(KParams& P, cudaStream_t s) {
  if (!P.zero_req) return CQ_OK;
  P.zero_req = 0;
  if (P.zero_keep_hi > P.zero_keep_lo) {  // all but the kept range (cq_csr.zero_keep_*)
    if (int rc = check_cuda(cudaMemsetAsync(P.values, 0, sizeof(double) * (size_t)P.zero_keep_lo, s), "zeroEntries")) return rc;
    return check_cuda(cudaMemsetAsync(P.values + P.zero_keep_hi, 0, sizeof(double) * (size_t)(P.nnz - P.zero_keep_hi), s), "zeroEntries");
  }
  return check_cuda(cudaMemsetAsync(P.values, 0, sizeof(double) * (size_t)P.nnz, s), "zeroEntries");
}


template <class T, int G>
static int launch_direct(cq_ctx* ctx, KParams P, cudaStream_t s) {
  if (int rc = prezero(P, s)) return rc;
  constexpr int THREADS = 256;
  constexpr int64_t ENTRIES_PER_BLOCK = (THREADS / 32) * (32 / G);
  const int grid = grid_for<assemble_direct_kernel<T, G, THREADS>>(ctx, THREADS, ENTRIES_PER_BLOCK, P.num_entries);
  assemble_direct_kernel<T, G, THREADS><<<grid, THREADS, 0, s>>>(P);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "assemble_direct_kernel launch");
}

// Can this call run the staged kernel?  (Otherwise the direct kernel handles it.)
template <class T>
static bool staged_ok(const cq_ctx* ctx, const KParams& P) {
  if (ctx->variant == 1 || P.elem_out) return false;
  if (T::RANK == 2 && !P.elem_pos) return false;  // needs the scatter map (cq_build_scatter_map)
  if (!aligned16(P.pts) || !aligned16(P.w) || (T::NEED_NORMAL && !aligned16(P.normals))) return false;
  return true;
}

template <class T, int G, int NPW, int NCW, int MODE = 0>
static int launch_staged(cq_ctx* ctx, KParams P, cudaStream_t s) {
  if (int rc = prezero(P, s)) return rc;
  constexpr int CB = 32;
  using Lay = StagedLayout<T, CB>;
  auto kernel = assemble_staged_kernel<T, G, NPW, NCW, MODE>;
  // stage capacity: ~1.25x the mean points of 32 entries, bounded by the shared memory per SM
  // (MODE 2: every consumer warp also owns a 27 x 27 tile for the row-by-row scatter, cq_dmma.cuh)
  const size_t tiles = MODE == 2 ? (size_t)NCW * DMMA_TILE_BYTES : 0;
  const size_t budget = (size_t)ctx->max_smem_optin - 1024 - tiles;
  const size_t per_point = (size_t)T::DIM * 8 * (T::NEED_NORMAL ? 2 : 1) + 8;
  const size_t fixed = Lay::meta_bytes + 3 * 128 + 2 * per_point;
  if (budget / NPW <= fixed + 64 * per_point) { set_error("staged kernel: stage header does not fit shared memory"); return CQ_ERR_UNSUPPORTED; }
  const int64_t pb_max = (int64_t)((budget / NPW - fixed) / per_point) & ~(int64_t)1;
  const double factor = CQ_TUNE_F64("CQ_STAGE_FACTOR", 1.25);  // stage capacity / mean stage size
  int64_t pb = (int64_t)((P.num_points / (P.num_entries > 0 ? P.num_entries : 1) + 1) * CB * factor) + 64;
  pb = (pb + 1) & ~(int64_t)1;
  if (pb > pb_max) pb = pb_max;
  if (pb < 64) pb = 64;
  P.stage_points = (int)pb;
  P.dbg = CQ_TUNE_INT("CQ_DEBUG", 0);
  const size_t smem = NPW * Lay::stage_bytes((int)pb) + tiles;
  if (int rc = ensure_smem_optin(ctx, (const void*)kernel, smem)) return rc;
  const int64_t num_stages = (P.num_entries + CB - 1) / CB;
  const int grid = (int)(num_stages < ctx->num_sms ? num_stages : ctx->num_sms);
  if (int rc = choose_dynamic(ctx, P, num_stages, grid * NPW, s, MODE == 2)) return rc;
  kernel<<<grid, (NPW + NCW) * 32, smem, s>>>(P);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "assemble_staged_kernel launch");
}

// Can / should the persistent moment kernels zero the matrix themselves (cq_zero.cuh)?
// ctx->fused_zero: 0 never, 1 wherever a kernel carries the zeroing warp, 2 (default) where it was measured to pay:
// the Q2 quadrilateral kernels in the shape 4 + 8 + 3 + 1 (one expansion set, the wait for the zeroed ranges in the
// first moment warp; bench workload, zero + assemble: mass 1.78 -> 1.54 ms, stiffness 1.83 -> 1.77 ms, DRAM traffic
// 9.2 -> 7.0 GB per matrix; as 4 + 5 + 6 + 1 with the wait in the expansion warps, round 1: 1.96 / 1.98 ms).  Not: Q1
// quadrilaterals (0.84 -> 0.93 ms), Q1 hexahedra (a 15th warp, no register cost, correct -- tests/test_gpu_fused_zero.py
// -- but 128^3 cells, 8 points: 0.707 vs 0.587 ms; config 3 at 256^3: 2.42 vs 2.38 ms per step).
// profiles/r2s3/call2_*.log, call5_*.log, call10_*.log
// (a matrix with peer rows qualifies: the zeroing touches this rank's own value array only, and the slots a neighbour
// adds into are excluded by the caller's zero_keep range)
static bool fused_zero_possible(const cq_ctx* ctx, const KParams& P) {
  return P.zero_req && aligned16(P.values) && ctx->coop_launch;
}
static bool fused_zero_ok(const cq_ctx* ctx, const KParams& P, const bool pays = false) {
  return (ctx->fused_zero == 1 || (ctx->fused_zero == 2 && pays)) && fused_zero_possible(ctx, P);
}
// ... and only where the stages are dealt out statically (the zeroing front follows the waves of the static
// interleaving; lists long enough for tickets / traversal tiles keep the memset)
static bool fused_zero_ok_static(const cq_ctx* ctx, const KParams& P, const int slots) {
  const int64_t num_stages = (P.num_entries + 31) / 32;
  return fused_zero_ok(ctx, P, false) && num_stages <= (int64_t)CQ_DYNAMIC_ROUNDS * ctx->num_sms * slots && P.tile_sp == 0;
}

template <class T, int G, int NPW, int NMW, int NXW, bool XT = false, int NPP = 1, int ZW = 0>
static int launch_moment_split(cq_ctx* ctx, KParams P, cudaStream_t s) {
  constexpr int CB = 32;
  using Lay = StagedLayout<T, CB>;
  auto kernel = assemble_moment_split_kernel<T, G, NPW, NMW, NXW, XT, NPP, ZW>;
  const size_t budget = (size_t)ctx->max_smem_optin - 1024;
  const size_t per_point = (size_t)T::DIM * 8 * (T::NEED_NORMAL ? 2 : 1) + 8;
  const size_t fixed = Lay::meta_bytes + 3 * 128 + 2 * per_point;
  if (budget / NPW <= fixed + 64 * per_point) { set_error("moment kernel: stage header does not fit shared memory"); return CQ_ERR_UNSUPPORTED; }
  const int64_t pb_max = (int64_t)((budget / NPW - fixed) / per_point) & ~(int64_t)1;
  const double factor = CQ_TUNE_F64("CQ_STAGE_FACTOR", 1.25);
  int64_t pb = (int64_t)((P.num_points / (P.num_entries > 0 ? P.num_entries : 1) + 1) * CB * factor) + 64;
  pb = (pb + 1) & ~(int64_t)1;
  if (pb > pb_max) pb = pb_max;
  if (pb < 64) pb = 64;
  if constexpr (XT) {
    // the expansion tiles of a stage reuse its point buffer: it has to hold (DEG + 1) of them
    // (the tiles may run over the weights / normals that follow the points: all dead by then)
    const int64_t need = (int64_t)(((T::DEG + 1) * RowTile<T::DIM, T::DEG>::bytes(CB) + per_point - 1) / per_point);
    if (pb < need) pb = (need + 1) & ~(int64_t)1;
    if (pb > pb_max) { set_error("moment kernel: expansion tiles do not fit the stage"); return CQ_ERR_UNSUPPORTED; }
  }
  P.stage_points = (int)pb;
  P.dbg = CQ_TUNE_INT("CQ_DEBUG", 0);
  const size_t smem = NPW * Lay::stage_bytes((int)pb);
  if (int rc = ensure_smem_optin(ctx, (const void*)kernel, smem)) return rc;
  const int64_t num_stages = (P.num_entries + CB - 1) / CB;
  const int grid = (int)(num_stages < ctx->num_sms ? num_stages : ctx->num_sms);
  if (P.zero_req) {
    // fused zeroEntries (cq_zero.cuh)
    if (ZW && fused_zero_possible(ctx, P)) {  // the caller chose this instantiation by the policy (fused_zero_ok)
      if (int rc = prepare_fused_zero(ctx, P, T::N, grid, s)) return rc;
      void* args[] = {(void*)&P};
      ctx->launches++;
      return check_cuda(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3((NPW * NPP + NMW + NXW + ZW) * 32), args, smem, s),
                        "assemble_moment_split_kernel cooperative launch");
    }
    if (int rc = prezero(P, s)) return rc;
  }
  if (NPP == 1)
    if (int rc = choose_dynamic(ctx, P, num_stages, grid * NPW, s)) return rc;
  kernel<<<grid, (NPW * NPP + NMW + NXW + ZW) * 32, smem, s>>>(P);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "assemble_moment_split_kernel launch");
}

// lanes per entry / consumer warps of the point-by-point kernel by accumulator count (measured on Q1 hexahedra,
// Nitsche, 128^3, U{8..27} points: G = 8, 6 warps 3.92 ms; G = 4, 8 warps 2.67 ms; 12 warps 3.8-4.8 ms (spills))
#ifndef CQ_STAGED_G36   // 17..36 accumulators
#define CQ_STAGED_G36 4
#endif
#ifndef CQ_STAGED_NCW36
#define CQ_STAGED_NCW36 8
#endif
#ifndef CQ_STAGED_G45   // more than 36 accumulators
#define CQ_STAGED_G45 8
#endif
#ifndef CQ_STAGED_NCW45
#define CQ_STAGED_NCW45 4
#endif
// consumer warps by accumulator size: the register file holds (NPW + NCW) * 32 threads
// (measured, profiles/r1/configs_session3.txt: 6 slots change nothing, 8 slots cost 10-60 %)
#ifndef CQ_STAGED_NPW_VS  // stage slots of the vector / scalar kernels with at most 16 accumulators (16 warps in all)
#define CQ_STAGED_NPW_VS 4
#endif
#ifndef CQ_STAGED_NCW_VS3  // consumer warps of the hexahedron kernels with at most 16 accumulators (Q1 vectors, scalars)
#define CQ_STAGED_NCW_VS3 12
#endif
template <class T>
struct StagedShape {
  static constexpr bool SMALL_VS = T::RANK < 2 && T::NACC <= 16;
  static constexpr int NPW = SMALL_VS ? CQ_STAGED_NPW_VS : 4;
  static constexpr int NCW = T::NACC <= 16 ? (T::DIM == 3 ? CQ_STAGED_NCW_VS3 : 16 - NPW) : (T::NACC <= 36 ? CQ_STAGED_NCW36 : CQ_STAGED_NCW45);
  static constexpr int G = T::NACC <= 16 ? 4 : (T::NACC <= 36 ? CQ_STAGED_G36 : CQ_STAGED_G45);
};

#ifndef CQ_MOMENT_G
#define CQ_MOMENT_G 2
#endif
#ifndef CQ_MOMENT_NPW
#define CQ_MOMENT_NPW 4
#endif
#ifndef CQ_MOMENT_NMW   // moment warps, degree 1 (4 + 8 + 2 = 14 warps)
#define CQ_MOMENT_NMW 8
#endif
#ifndef CQ_MOMENT_NMW2  // moment warps, degree 2 (4 + 6 + 6 = 16 warps: 128 registers per thread, no spills)
#define CQ_MOMENT_NMW2 6
#endif
#ifndef CQ_MOMENT_NXW1
#define CQ_MOMENT_NXW1 3
#endif
#ifndef CQ_MOMENT_SG
#define CQ_MOMENT_SG 4
#endif
#ifndef CQ_MOMENT_LONG2  // quadrilaterals: mean points per cell above which the long-rule shape is used (measured: the
#define CQ_MOMENT_LONG2 48  // bench workload, mean 40, loses 4 % in it; U{40..80}, mean 60, gains 23 % (Q2) / 4 % (Q1))
#endif
#ifndef CQ_MOMENT_NMW2Z  // Q2 quadrilaterals with the zeroing warp: moment warps / expansion sets (4 + NMW2Z + 3 NXS2Z + 1 <= 16;
#define CQ_MOMENT_NMW2Z 8  // 4 + 5 + 6 + 1: zero + mass 1.96, zero + stiffness 1.98 ms; 4 + 8 + 3 + 1: 1.54 / 1.77 ms)
#endif
#ifndef CQ_MOMENT_NXS2Z
#define CQ_MOMENT_NXS2Z 1
#endif
#ifndef CQ_MOMENT_G3    // Q1 hexahedra: lanes per entry in the point loop, moment warps, expansion warp sets
#define CQ_MOMENT_G3 2
#endif
#ifndef CQ_MOMENT_NPW3   // stage slots (= producer warps)
#define CQ_MOMENT_NPW3 4
#endif
#ifndef CQ_MOMENT_NPW3_SHORT  // ... for rules of at most 12 points per cell on average
#define CQ_MOMENT_NPW3_SHORT 6
#endif
#ifndef CQ_MOMENT_NMW3_SHORT  // ... its moment warps and expansion sets: 6 + 6 + 4 = 16 warps (128^3 cells, 8 points: stiffness
#define CQ_MOMENT_NMW3_SHORT 6  // 0.503 -> 0.470 ms, mass 0.495 -> 0.486 ms; three expansion sets instead, 6 + 4 + 6: 0.500 / 0.506)
#endif
#ifndef CQ_MOMENT_NXS3_SHORT
#define CQ_MOMENT_NXS3_SHORT CQ_MOMENT_NXS3
#endif
#ifndef CQ_MOMENT_NMW3
#define CQ_MOMENT_NMW3 4
#endif
#ifndef CQ_MOMENT_LONG3  // Q1 hexahedra: mean points per cell above which the long-rule shape is used
#define CQ_MOMENT_LONG3 32
#endif
#ifndef CQ_MOMENT_NMW3L  // ... its moment warps (NPW3L producers + NMW3L + 2 expansion warps)
#define CQ_MOMENT_NMW3L 10
#endif
#ifndef CQ_MOMENT_NPW3L  // ... its stage slots
#define CQ_MOMENT_NPW3L 4
#endif
#ifndef CQ_MOMENT_LONG3_3SLOT  // ... up to this mean three slots (a whole stage fits in 73 KB)
#define CQ_MOMENT_LONG3_3SLOT 56
#endif
#ifndef CQ_MOMENT_NPP3   // producer warps per stage slot, Q1 hexahedra
#define CQ_MOMENT_NPP3 1
#endif
#ifndef CQ_MOMENT_NXS3
#define CQ_MOMENT_NXS3 2
#endif
#ifndef CQ_NITSCHE_NMW3  // Nitsche moment kernel (Q1 hexahedra): moment warps / expansion sets (3 stage slots)
#define CQ_NITSCHE_NMW3 6
#endif
#ifndef CQ_NITSCHE_NXS3  // three sets = one per stage slot (15 warps, 128 registers): U{4..16} surface points per cell 0.61 -> 0.49 ms
#define CQ_NITSCHE_NXS3 3
#endif
#ifndef CQ_MOMENT_NPP  // producer warps per stage slot
#define CQ_MOMENT_NPP 1
#endif
#ifndef CQ_MOMENT_NXS1  // expansion warp sets, degree 1
#define CQ_MOMENT_NXS1 1
#endif
#ifndef CQ_MOMENT_NXS2  // expansion warp sets, degree 2
#define CQ_MOMENT_NXS2 2
#endif
// mesh flags (cq_mesh.flags)
constexpr int MESH_AXIS_ALIGNED = 1;

template <class T, int GDIRECT = 8>
static int launch_auto(cq_ctx* ctx, const KParams& P, cudaStream_t s) {
  using S = StagedShape<T>;
  if (staged_ok<T>(ctx, P)) {
    // mass / stiffness on quadrilaterals whose Jacobian is constant and diagonal: moment kernel
    if constexpr (T::DIM == 2 && (T::FORM_ID == CQ_FORM_MASS || T::FORM_ID == CQ_FORM_STIFFNESS)) {
      if ((P.mesh_flags & MESH_AXIS_ALIGNED) && ctx->variant != 3) {
        // three-role kernel for Q1 (measured 1.00 vs 1.06 ms on the bench workload); for Q2 both
        // organisations are bound by the L2 atomic rate of the 81 REDs per cell (2.56 vs 2.59 ms)
        // split = 2: row-owned expansion + transposed scatter (RowTile, cq_moment.cuh)
        const int split = CQ_TUNE_INT("CQ_MOMENT_SPLIT", 2);
        if (split == 2) {
          // fused zeroEntries (cq_ctx_set_fused_zero; policy: fused_zero_ok): one more warp.  Q1: 4 + 8 + 2 + 1 warps
          // still cost 16 warps' worth of registers; Q2: 4 + 6 + 6 = 16 warps at 128 registers and a 17th warp costs the
          // registers of 20 (96 per thread, spills in the point loop), so the zeroing warp and two more moment warps
          // take the place of the second expansion set: 4 + 8 + 3 + 1
          constexpr int NMW = T::DEG == 1 ? CQ_MOMENT_NMW : CQ_MOMENT_NMW2;
          constexpr int NXW = (T::DEG == 1 ? CQ_MOMENT_NXS1 : CQ_MOMENT_NXS2) * (T::DEG + 1);
          const bool long_rules = P.num_points > (int64_t)CQ_MOMENT_LONG2 * P.num_entries;  // (their shape below gains more: memset + it)
          if (fused_zero_ok(ctx, P, T::DEG == 2 && !long_rules))
            return launch_moment_split<T, CQ_MOMENT_SG, 4, (T::DEG == 2 ? CQ_MOMENT_NMW2Z : NMW), (T::DEG == 2 ? CQ_MOMENT_NXS2Z * 3 : NXW), true, CQ_MOMENT_NPP, 1>(ctx, P, s);
          // long rules: as on hexahedra, the serial point loops bind -> more moment warps (Q1 4 + 10 + 2, Q2 4 + 9 + 3:
          // one expansion set) and four points in flight per lane.  1024^2 cells: Q2 stiffness U{200..300} points per
          // cell 3.70 -> 2.11 ms (0.29 -> 0.51 of the HBM roofline), Q2 mass U{80..120} 1.57 -> 0.91, Q1 stiffness
          // U{80..120} 1.03 -> 0.66 (0.40 -> 0.63); profiles/r2s3/call3_quad_long_shape_unroll_nitsche_sets.log
          if (long_rules)
            return launch_moment_split<T, CQ_MOMENT_SG, 4, (T::DEG == 1 ? 10 : 9), T::DEG + 1, true, CQ_MOMENT_NPP, 0>(ctx, P, s);
          return launch_moment_split<T, CQ_MOMENT_SG, 4, NMW, NXW, true, CQ_MOMENT_NPP, 0>(ctx, P, s);
        }
        if (split) return launch_moment_split<T, CQ_MOMENT_SG, 4, CQ_MOMENT_NMW, T::DEG == 1 ? 2 : CQ_MOMENT_NXW1, false, 1, 0>(ctx, P, s);
        return launch_staged<T, CQ_MOMENT_G, CQ_MOMENT_NPW, CQ_MOMENT_NPW * CQ_MOMENT_G, 1>(ctx, P, s);
      }
    }
    // symmetric Nitsche form with quadrature normals on Q1 hexahedra: 4 moment sets (w, w n_d), one lane of the entry's
    // group per set.  Three stage slots (108 moments per entry in the header).  Not used for quadrilaterals: measured on
    // the bench-size Q1 mesh the point-by-point kernel is faster there (1.77 vs 4.2 ms: its per-point work is small and
    // it splits the points over the lanes, the moment sets run every point in every lane).
    if constexpr (T::FORM_ID == CQ_FORM_NITSCHE && T::DEG == 1 && T::DIM == 3) {
      if ((P.mesh_flags & MESH_AXIS_ALIGNED) && ctx->variant != 3)
        return launch_moment_split<T, 4, 3, CQ_NITSCHE_NMW3, CQ_NITSCHE_NXS3 * 2, true>(ctx, P, s);
    }
    // ... and on Q1 hexahedra (create_box meshes): 27 moments, two expansion roles
    if constexpr (T::DIM == 3 && T::DEG == 1 && (T::FORM_ID == CQ_FORM_MASS || T::FORM_ID == CQ_FORM_STIFFNESS)) {
      if ((P.mesh_flags & MESH_AXIS_ALIGNED) && ctx->variant != 3) {
        // short rules (the 8-point bulk rule of config 5): the stages are small and the ring is bound by the stage
        // rate, not by bytes -> six slots (measured on 128^3 cells, 8 points: stiffness 0.513 -> 0.475 ms, mass
        // 0.491 -> 0.466 ms; with U{8..27} points six slots no longer hold a stage and cost 0.60 -> 0.80 ms)
        // (fused zeroEntries: one more warp -- 6 + 4 + 4 + 1 = 15 or 4 + 4 + 4 + 1 = 13 in all, still 128 registers)
        if (P.num_points <= 12 * P.num_entries) {
          if (fused_zero_ok_static(ctx, P, CQ_MOMENT_NPW3_SHORT))
            return launch_moment_split<T, CQ_MOMENT_G3, CQ_MOMENT_NPW3_SHORT, CQ_MOMENT_NMW3, CQ_MOMENT_NXS3 * 2, true, CQ_MOMENT_NPP3, 1>(ctx, P, s);
          return launch_moment_split<T, CQ_MOMENT_G3, CQ_MOMENT_NPW3_SHORT, CQ_MOMENT_NMW3_SHORT, CQ_MOMENT_NXS3_SHORT * 2, true, CQ_MOMENT_NPP3>(ctx, P, s);
        }
        // long rules (real cut cells carry ~110 points in 3-D, config 3): the kernel is bound by the serial point
        // loops of its moment lanes, not by bytes (4 moment warps: 0.27 of the HBM roofline at U{40..60} points per cell,
        // 0.24 at U{100..120}, 0.23 at U{200..300}) -> ten moment warps and one expansion set (16 warps, 128 registers:
        // 0.36 / 0.33 / 0.29; cut-cell stiffness of config 3 0.66 -> 0.54 ms).  Where three stage slots hold a whole
        // stage and four do not (mean of 33..56 points per cell) three slots keep every entry in shared memory (U{40..60}:
        // 0.53; beyond that the entries that overflow are read from global memory by the point loops either way and the
        // deeper ring wins).  profiles/r2s3/call1_long_rule_shapes_sg8.log
        if (P.num_points > (int64_t)CQ_MOMENT_LONG3 * P.num_entries) {
          if (P.num_points <= (int64_t)CQ_MOMENT_LONG3_3SLOT * P.num_entries)
            return launch_moment_split<T, CQ_MOMENT_G3, 3, CQ_MOMENT_NMW3L, 2, true, CQ_MOMENT_NPP3>(ctx, P, s);
          return launch_moment_split<T, CQ_MOMENT_G3, CQ_MOMENT_NPW3L, CQ_MOMENT_NMW3L, 2, true, CQ_MOMENT_NPP3>(ctx, P, s);
        }
        if (fused_zero_ok_static(ctx, P, CQ_MOMENT_NPW3))
          return launch_moment_split<T, CQ_MOMENT_G3, CQ_MOMENT_NPW3, CQ_MOMENT_NMW3, CQ_MOMENT_NXS3 * 2, true, CQ_MOMENT_NPP3, 1>(ctx, P, s);
        return launch_moment_split<T, CQ_MOMENT_G3, CQ_MOMENT_NPW3, CQ_MOMENT_NMW3, CQ_MOMENT_NXS3 * 2, true, CQ_MOMENT_NPP3>(ctx, P, s);
      }
    }
    return launch_staged<T, S::G, S::NPW, S::NCW>(ctx, P, s);
  }
  if (ctx->variant == 2) { set_error("staged kernel requested but not applicable (scatter map / alignment / element output)"); return CQ_ERR_UNSUPPORTED; }
  return launch_direct<T, GDIRECT>(ctx, P, s);
}

}  // namespace cq
