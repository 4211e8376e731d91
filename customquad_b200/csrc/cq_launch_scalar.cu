// cq_launch_scalar.cu -- instantiation + dispatch of the rank-0 (functional) kernels
#include "cq_launch.cuh"

namespace cq {

// closed-form data function (u_ex / f): COEF is not a legal choice here
template <int FORM, int DIM, int DEG, int BS>
static int launch_analytic(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  switch (f.fn_kind) {
    case CQ_FN_CONST: return launch_auto<FormTraits<FORM, DIM, DEG, BS, CQ_FN_CONST>>(ctx, P, s);
    case CQ_FN_AFFINE: return launch_auto<FormTraits<FORM, DIM, DEG, BS, CQ_FN_AFFINE>>(ctx, P, s);
    case CQ_FN_SINPROD: return launch_auto<FormTraits<FORM, DIM, DEG, BS, CQ_FN_SINPROD>>(ctx, P, s);
    default: set_error("form %d needs a closed-form fn_kind (got %d)", FORM, f.fn_kind); return CQ_ERR_INVALID;
  }
}

template <int FORM, int DIM>
static int launch_err(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  if (f.bs != 1 && f.bs != DIM) { set_error("bs must be 1 or gdim (got %d)", f.bs); return CQ_ERR_UNSUPPORTED; }
  if (f.degree == 1) return f.bs == 1 ? launch_analytic<FORM, DIM, 1, 1>(ctx, f, P, s) : launch_analytic<FORM, DIM, 1, DIM>(ctx, f, P, s);
  return f.bs == 1 ? launch_analytic<FORM, DIM, 2, 1>(ctx, f, P, s) : launch_analytic<FORM, DIM, 2, DIM>(ctx, f, P, s);
}

template <int DIM>
static int launch_sq(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  constexpr int FORM = CQ_FORM_INTEGRAL_SQ;
  if (f.bs != 1 && f.bs != DIM) { set_error("bs must be 1 or gdim (got %d)", f.bs); return CQ_ERR_UNSUPPORTED; }
  if (f.fn_kind != CQ_FN_COEF)  // closed-form f: the element is irrelevant
    return f.bs == 1 ? launch_analytic<FORM, DIM, 1, 1>(ctx, f, P, s) : launch_analytic<FORM, DIM, 1, DIM>(ctx, f, P, s);
  if (f.degree == 1)
    return f.bs == 1 ? launch_auto<FormTraits<FORM, DIM, 1, 1, CQ_FN_COEF>>(ctx, P, s)
                     : launch_auto<FormTraits<FORM, DIM, 1, DIM, CQ_FN_COEF>>(ctx, P, s);
  return f.bs == 1 ? launch_auto<FormTraits<FORM, DIM, 2, 1, CQ_FN_COEF>>(ctx, P, s)
                   : launch_auto<FormTraits<FORM, DIM, 2, DIM, CQ_FN_COEF>>(ctx, P, s);
}

int launch_scalar(cq_ctx* ctx, const cq_form_desc& f, const KParams& P, cudaStream_t s) {
  const bool quad = f.cell_type == CQ_CELL_QUADRILATERAL;
  switch (f.form) {
    case CQ_FORM_INTEGRAL:
      return quad ? launch_analytic<CQ_FORM_INTEGRAL, 2, 1, 1>(ctx, f, P, s) : launch_analytic<CQ_FORM_INTEGRAL, 3, 1, 1>(ctx, f, P, s);
    case CQ_FORM_INTEGRAL_SQ: return quad ? launch_sq<2>(ctx, f, P, s) : launch_sq<3>(ctx, f, P, s);
    case CQ_FORM_L2_ERR: return quad ? launch_err<CQ_FORM_L2_ERR, 2>(ctx, f, P, s) : launch_err<CQ_FORM_L2_ERR, 3>(ctx, f, P, s);
    case CQ_FORM_H10_ERR: return quad ? launch_err<CQ_FORM_H10_ERR, 2>(ctx, f, P, s) : launch_err<CQ_FORM_H10_ERR, 3>(ctx, f, P, s);
    default: set_error("form %d is not a functional", f.form); return CQ_ERR_INVALID;
  }
}

}  // namespace cq
