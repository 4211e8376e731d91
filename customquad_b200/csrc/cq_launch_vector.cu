// cq_launch_vector.cu -- instantiation + dispatch of the rank-1 (linear form) kernels
#include "cq_internal.h"

namespace cq {

template <class T, int G>
static int launch_direct(cq_ctx* ctx, const KParams& P, cudaStream_t s) {
  constexpr int THREADS = 256;
  constexpr int64_t ENTRIES_PER_BLOCK = (THREADS / 32) * (32 / G);
  const int grid = grid_for<assemble_direct_kernel<T, G, THREADS>>(ctx, THREADS, ENTRIES_PER_BLOCK, P.num_entries);
  assemble_direct_kernel<T, G, THREADS><<<grid, THREADS, 0, s>>>(P);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "assemble_direct_kernel launch");
}

template <int FORM, int DIM, int DEG, int BS>
static int launch_fn(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  switch (f.fn_kind) {
    case CQ_FN_COEF: return launch_direct<FormTraits<FORM, DIM, DEG, BS, CQ_FN_COEF>, 8>(ctx, P, s);
    case CQ_FN_CONST: return launch_direct<FormTraits<FORM, DIM, DEG, BS, CQ_FN_CONST>, 8>(ctx, P, s);
    case CQ_FN_AFFINE: return launch_direct<FormTraits<FORM, DIM, DEG, BS, CQ_FN_AFFINE>, 8>(ctx, P, s);
    case CQ_FN_SINPROD: return launch_direct<FormTraits<FORM, DIM, DEG, BS, CQ_FN_SINPROD>, 8>(ctx, P, s);
    default: set_error("unknown fn_kind %d", f.fn_kind); return CQ_ERR_INVALID;
  }
}

template <int FORM, int DIM>
static int launch_elem(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  if (f.bs != 1 && f.bs != DIM) { set_error("bs must be 1 or gdim (got %d)", f.bs); return CQ_ERR_UNSUPPORTED; }
  if (f.degree == 1) return f.bs == 1 ? launch_fn<FORM, DIM, 1, 1>(ctx, f, P, s) : launch_fn<FORM, DIM, 1, DIM>(ctx, f, P, s);
  return f.bs == 1 ? launch_fn<FORM, DIM, 2, 1>(ctx, f, P, s) : launch_fn<FORM, DIM, 2, DIM>(ctx, f, P, s);
}

int launch_vector(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  const bool quad = f.cell_type == CQ_CELL_QUADRILATERAL;
  switch (f.form) {
    case CQ_FORM_LOAD: return quad ? launch_elem<CQ_FORM_LOAD, 2>(ctx, f, P, s) : launch_elem<CQ_FORM_LOAD, 3>(ctx, f, P, s);
    case CQ_FORM_NITSCHE_RHS:
      return quad ? launch_elem<CQ_FORM_NITSCHE_RHS, 2>(ctx, f, P, s) : launch_elem<CQ_FORM_NITSCHE_RHS, 3>(ctx, f, P, s);
    default: set_error("form %d is not a linear form", f.form); return CQ_ERR_INVALID;
  }
}

}  // namespace cq
