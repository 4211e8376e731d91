// cq_launch_vector.cu -- instantiation + dispatch of the rank-1 (linear form) kernels
#include "cq_launch.cuh"

namespace cq {

template <int FORM, int DIM, int DEG, int BS>
static int launch_fn(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  switch (f.fn_kind) {
    case CQ_FN_COEF: return launch_auto<FormTraits<FORM, DIM, DEG, BS, CQ_FN_COEF>>(ctx, P, s);
    case CQ_FN_CONST: return launch_auto<FormTraits<FORM, DIM, DEG, BS, CQ_FN_CONST>>(ctx, P, s);
    case CQ_FN_AFFINE: return launch_auto<FormTraits<FORM, DIM, DEG, BS, CQ_FN_AFFINE>>(ctx, P, s);
    case CQ_FN_SINPROD: return launch_auto<FormTraits<FORM, DIM, DEG, BS, CQ_FN_SINPROD>>(ctx, P, s);
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
