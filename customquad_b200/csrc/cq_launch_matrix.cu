// cq_launch_matrix.cu -- instantiation + dispatch of the rank-2 (bilinear form) kernels
#include "cq_dmma.cuh"
#include "cq_launch.cuh"
#include "cq_owner.cuh"

// consumer warps (one entry per warp) of the staged Q2-hexahedron tensor-core kernel.  Measured on 48^3 cells with 27
// points each (profiles/r1/q2hex_ncw.txt): scalar spaces are bound by the contraction and gain from 13 warps
// (stiffness 2.19 -> 1.45 ms), blocked spaces (bs = 3, 9x the scatter per entry) are bound by the RED scatter and are
// fastest with 5 (3.52 ms; 3.89 ms with 13).
#ifndef CQ_DMMA_NCW
#define CQ_DMMA_NCW 13
#endif
#ifndef CQ_DMMA_NCW_BLOCKED
#define CQ_DMMA_NCW_BLOCKED 5
#endif

namespace cq {

template <int FORM>
static int launch_dmma(cq_ctx* ctx, KParams P, cudaStream_t s) {
  if (int rc = prezero(P, s)) return rc;
  constexpr int THREADS = 128;
  const int grid = grid_for<assemble_q2hex_dmma_kernel<FORM, THREADS>>(ctx, THREADS, THREADS / 32, P.num_entries);
  assemble_q2hex_dmma_kernel<FORM, THREADS><<<grid, THREADS, 0, s>>>(P);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "assemble_q2hex_dmma_kernel launch");
}

// Owner-computes path for blocked Q2-hexahedron matrices (cq_owner.cuh).  Returns CQ_OK with *done = false when the
// call does not qualify (the RED path takes it).
template <class TQ2, int NPW>
static int launch_owner_q2hex(cq_ctx* ctx, const KParams& P, cudaStream_t s, bool* done) {
  *done = false;
  constexpr int N = TQ2::N;
  if (!P.zero_req || P.bs < 2 || P.bs > 3 || P.peer_rowstart != nullptr || P.zero_keep_hi > P.zero_keep_lo || ctx->variant != 0 || P.num_dofs <= 0) return CQ_OK;
  if (P.num_entries >= (1LL << 26) || P.num_dofs >= 0x7fffffffLL) return CQ_OK;  // (k * 32 + li) and the pair offsets are 32-bit
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t nd1 = (size_t)P.num_dofs + 1;
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int*)nullptr, (unsigned int*)nullptr, (int)nd1, s);
  const size_t b_cnt = up(nd1 * 4), b_ptr = up(nd1 * 4), b_scan = up(scan_bytes), b_pairs = up((size_t)P.num_entries * N * sizeof(int2)),
               b_elem = up((size_t)P.num_entries * N * N * 8);
  const size_t need = b_cnt + b_ptr + b_scan + b_pairs + b_elem;
  if (need > ctx->owner_ws_bytes) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return CQ_OK;
    if (need > free_b + ctx->owner_ws_bytes || need + (1ull << 30) > free_b + ctx->owner_ws_bytes) return CQ_OK;  // not enough room: RED path
    if (ctx->owner_ws) cudaFree(ctx->owner_ws);
    ctx->owner_ws = nullptr;
    ctx->owner_ws_bytes = 0;
    if (cudaMalloc(&ctx->owner_ws, need) != cudaSuccess) { cudaGetLastError(); return CQ_OK; }
    ctx->owner_ws_bytes = need;
  }
  unsigned char* w = static_cast<unsigned char*>(ctx->owner_ws);
  if (int rc = check_cuda(cudaMemsetAsync(w, 0, b_cnt, s), "memset(owner counts)")) return rc;
  OwnerParams O;
  O.num_entries = P.num_entries; O.num_dofs = P.num_dofs; O.n = N; O.npos = StageMeta<TQ2, 32>::NPOS; O.bs = P.bs;
  O.cells = P.cells; O.dofmap = P.dofmap; O.elem_pos = P.elem_pos; O.indptr = P.indptr; O.values = P.values;
  O.cnt = reinterpret_cast<int*>(w);
  O.ptr = reinterpret_cast<unsigned int*>(w + b_cnt);
  void* scan_tmp = w + b_cnt + b_ptr;
  O.pairs = reinterpret_cast<int2*>(w + b_cnt + b_ptr + b_scan);
  double* elem = reinterpret_cast<double*>(w + b_cnt + b_ptr + b_scan + b_pairs);
  O.elem = elem;
  // phase 0: dof -> (entry, local index) lists
  const int64_t total = P.num_entries * N;
  const int blocks = (int)((total + 255) / 256 < (int64_t)ctx->num_sms * 16 ? (total + 255) / 256 : (int64_t)ctx->num_sms * 16);
  owner_count_kernel<<<blocks, 256, 0, s>>>(O);
  if (int rc = check_cuda(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, O.cnt, O.ptr, (int)nd1, s), "owner scan")) return rc;
  owner_fill_kernel<<<blocks, 256, 0, s>>>(O);
  ctx->launches += 3;
  if (int rc = check_cuda(cudaGetLastError(), "owner setup kernels")) return rc;
  // phase 1: element blocks (the scatter-free flavour of the staged tensor-core kernel: bound by the contraction, 13 consumer warps)
  KParams P1 = P;
  P1.elem_scalar = elem;
  P1.zero_req = 0;
  if (int rc = launch_staged<TQ2, 32, NPW, CQ_DMMA_NCW, 2>(ctx, P1, s)) return rc;
  // phase 2: every row stored once by its owner
  const int64_t nb = (P.num_dofs + OWNER_WARPS - 1) / OWNER_WARPS;
  const int grid = (int)(nb < (int64_t)ctx->num_sms * 8 ? nb : (int64_t)ctx->num_sms * 8);
  // bs = 2 runs the run-time-bs instantiation: the <2> one was observed to fail with "misaligned address" on the B200
  // (tests/test_gpu_owner.py, 60-cell mesh; its SASS holds no access wider than 8 bytes; compute-sanitizer is not
  // available on the pool) while <0> and <3> pass the same cases -- not understood, so not used
  if (P.bs == 2) owner_rows_kernel<0><<<grid, OWNER_WARPS * 32, 0, s>>>(O);
  else owner_rows_kernel<3><<<grid, OWNER_WARPS * 32, 0, s>>>(O);
  ctx->launches++;
  *done = true;
  return check_cuda(cudaGetLastError(), "owner_rows_kernel launch");
}

template <int FORM>
static int launch_matrix_form(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  const bool quad = f.cell_type == CQ_CELL_QUADRILATERAL;
  if (quad && f.degree == 1) return launch_auto<FormTraits<FORM, 2, 1, 1, CQ_FN_CONST>>(ctx, P, s);
  if (quad && f.degree == 2) return launch_auto<FormTraits<FORM, 2, 2, 1, CQ_FN_CONST>>(ctx, P, s);
  if (!quad && f.degree == 1) return launch_auto<FormTraits<FORM, 3, 1, 1, CQ_FN_CONST>>(ctx, P, s);
  // Q2 hexahedron: FP64 tensor-core contraction, staged (warp per entry) when the scatter map is there
  using TQ2 = FormTraits<FORM, 3, 2, 1, CQ_FN_CONST>;
  if (staged_ok<TQ2>(ctx, P)) {
    constexpr int NPW = 2;  // two stage slots: the consumer warps' scatter tiles take up to 78 KB of the shared memory
    bool done = false;
    if (int rc = launch_owner_q2hex<TQ2, NPW>(ctx, P, s, &done)) return rc;
    if (done) return CQ_OK;
    return P.bs == 1 ? launch_staged<TQ2, 32, NPW, CQ_DMMA_NCW, 2>(ctx, P, s)
                     : launch_staged<TQ2, 32, NPW, CQ_DMMA_NCW_BLOCKED, 2>(ctx, P, s);
  }
  if (ctx->variant == 2) { set_error("staged kernel requested but not applicable"); return CQ_ERR_UNSUPPORTED; }
  return launch_dmma<FORM>(ctx, P, s);
}

int launch_matrix(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  switch (f.form) {
    case CQ_FORM_MASS: return launch_matrix_form<CQ_FORM_MASS>(ctx, f, P, s);
    case CQ_FORM_STIFFNESS: return launch_matrix_form<CQ_FORM_STIFFNESS>(ctx, f, P, s);
    case CQ_FORM_NITSCHE: return launch_matrix_form<CQ_FORM_NITSCHE>(ctx, f, P, s);
    default: set_error("form %d is not a bilinear form", f.form); return CQ_ERR_INVALID;
  }
}

}  // namespace cq
