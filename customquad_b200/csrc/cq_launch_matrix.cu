// cq_launch_matrix.cu -- instantiation + dispatch of the rank-2 (bilinear form) kernels
#include "cq_dmma.cuh"
#include "cq_launch.cuh"

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
