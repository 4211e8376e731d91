// cq_forms.cuh -- what the generated tabulate_tensor_runtime_float64 computes at ONE quadrature
// point (reference README.md:133-158; integrands: SURVEY.md A.1), written so that a thread can
// accumulate any number of points of one cell into private registers.
//
//   A_K += sum_q w_q |det J_q| integrand_q         (weights are reference-measure weights)
//
// Matrix forms are symmetric (mass, stiffness, and the symmetric Nitsche form of
// demo/poisson.py:156-158), so only the upper triangle of the scalar n x n block is accumulated;
// blocked spaces replicate that block on the bs diagonal components, which the
// scatter stage does.
#pragma once
#include "cq_common.cuh"

namespace cq {

template <int FORM, int DIM_, int DEG_, int BS_, int FN_>
struct FormTraits {
  static constexpr int FORM_ID = FORM;
  static constexpr int DIM = DIM_;
  static constexpr int DEG = DEG_;
  static constexpr int BS = BS_;
  static constexpr int FN = FN_;
  static constexpr int N = Lattice<DIM, DEG>::N;
  static constexpr int NV = 1 << DIM;
  static constexpr int RANK = FORM < 10 ? 2 : (FORM < 20 ? 1 : 0);
  static constexpr int NACC = RANK == 2 ? tri_count(N) : (RANK == 1 ? N * BS : 1);
  static constexpr bool IS_ERR = FORM == CQ_FORM_L2_ERR || FORM == CQ_FORM_H10_ERR;
  static constexpr bool NEED_NORMAL = FORM == CQ_FORM_NITSCHE || FORM == CQ_FORM_NITSCHE_RHS;
  static constexpr bool NEED_GRAD = FORM == CQ_FORM_STIFFNESS || NEED_NORMAL || FORM == CQ_FORM_H10_ERR;
  static constexpr bool NEED_COEF = RANK < 2 && (FN == CQ_FN_COEF || IS_ERR);
  static constexpr bool NEED_X = RANK < 2 && (FN == CQ_FN_AFFINE || FN == CQ_FN_SINPROD);
  static constexpr bool NEED_PSI = FORM == CQ_FORM_MASS || NEED_NORMAL || FORM == CQ_FORM_LOAD ||
                                   (NEED_COEF && FORM != CQ_FORM_H10_ERR);
  static constexpr bool NEED_H = NEED_NORMAL;
  static constexpr int NCOEF = N * BS;
};

// psi_i and reference gradients g_i[e] at xi, compile-time lattice => fully unrolled products
template <int DIM, int DEG, bool VAL, bool GRAD>
CQ_DEVINL void tabulate_point(const double (&xi)[DIM], double (&psi)[Lattice<DIM, DEG>::N],
                              double (&g)[Lattice<DIM, DEG>::N][DIM]) {
  using L = Lattice<DIM, DEG>;
  const Basis1D<DEG> b0(xi[0]);
  const Basis1D<DEG> b1(xi[1]);
  if constexpr (DIM == 2) {
#pragma unroll
    for (int j = 0; j < L::N; ++j) {
      const int t0 = L::t(j, 0), t1 = L::t(j, 1);
      if constexpr (VAL) psi[j] = b0.v[t0] * b1.v[t1];
      if constexpr (GRAD) {
        g[j][0] = b0.d[t0] * b1.v[t1];
        g[j][1] = b0.v[t0] * b1.d[t1];
      }
    }
  } else {
    const Basis1D<DEG> b2(xi[2]);
    // pairwise products shared between value and gradients
    double v01[DEG + 1][DEG + 1];
#pragma unroll
    for (int a = 0; a <= DEG; ++a)
#pragma unroll
      for (int b = 0; b <= DEG; ++b) v01[a][b] = b0.v[a] * b1.v[b];
#pragma unroll
    for (int j = 0; j < L::N; ++j) {
      const int t0 = L::t(j, 0), t1 = L::t(j, 1), t2 = L::t(j, 2);
      if constexpr (VAL) psi[j] = v01[t0][t1] * b2.v[t2];
      if constexpr (GRAD) {
        g[j][0] = (b0.d[t0] * b1.v[t1]) * b2.v[t2];
        g[j][1] = (b0.v[t0] * b1.d[t1]) * b2.v[t2];
        g[j][2] = v01[t0][t1] * b2.d[t2];
      }
    }
  }
}

// One quadrature point of one cell.  acc: per-thread partial element tensor.
//   rank 2: acc[tri_index(N, i, j)], i <= j   (scalar block)
//   rank 1: acc[i*BS + a]
//   rank 0: acc[0]
// coef: the cell's packed coefficient [N*BS] (blocked interleaved), gamma = betaN / h_K,
// p: parameters of the closed-form data function.
template <class T>
CQ_DEVINL void accumulate_point(double (&acc)[T::NACC], const CellGeom<T::DIM>& geo, const double (&xi)[T::DIM],
                                const double w, const double (&nrm)[T::DIM], const double* __restrict__ coef,
                                const double gamma, const double* __restrict__ p) {
  constexpr int DIM = T::DIM, N = T::N, BS = T::BS;
  double psi[N], g[N][DIM];
  tabulate_point<DIM, T::DEG, T::NEED_PSI, T::NEED_GRAD>(xi, psi, g);

  double adj[DIM][DIM], det, inv_det;
  if (geo.affine) {  // constant Jacobian: per-cell values (warp-uniform for tensor meshes)
    det = geo.det0;
    inv_det = geo.inv_det0;
    if constexpr (T::NEED_GRAD) {
#pragma unroll
      for (int e = 0; e < DIM; ++e)
#pragma unroll
        for (int k = 0; k < DIM; ++k) adj[e][k] = geo.adj0[e][k];
    }
  } else {
    double J[DIM][DIM];
    geo.jacobian(xi, J);
    if constexpr (T::NEED_GRAD) {
      det = det_adj<DIM>(J, adj);
      inv_det = 1.0 / det;
    } else {
      det = det_only<DIM>(J);
      inv_det = 0.0;
    }
  }
  const double adet = fabs(det);
  const double s = w * adet;  // w_q |det J_q|

  if constexpr (T::FORM_ID == CQ_FORM_MASS) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double ti = s * psi[i];
#pragma unroll
      for (int j = i; j < N; ++j) acc[tri_index(N, i, j)] = fma(ti, psi[j], acc[tri_index(N, i, j)]);
    }
  } else if constexpr (T::FORM_ID == CQ_FORM_STIFFNESS) {
    // s G_i.G_j = (w/|det|) (adj^T g_i).(adj^T g_j),  G = adj^T g / det
    const double sc = w * fabs(inv_det);
    double a[N][DIM];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        double v = adj[0][k] * g[i][0];
#pragma unroll
        for (int e = 1; e < DIM; ++e) v = fma(adj[e][k], g[i][e], v);
        a[i][k] = v;
      }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double t[DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k) t[k] = sc * a[i][k];
#pragma unroll
      for (int j = i; j < N; ++j) {
        double v = acc[tri_index(N, i, j)];
#pragma unroll
        for (int k = 0; k < DIM; ++k) v = fma(t[k], a[j][k], v);
        acc[tri_index(N, i, j)] = v;
      }
    }
  } else {
    // ---- everything that may need n.grad(psi_i), coefficient values, closed-form data --------
    double dn[N];  // n . grad psi_i  (physical)
    if constexpr (T::NEED_NORMAL) {
      double m[DIM];  // m_e = sum_k adj[e][k] n_k / det
#pragma unroll
      for (int e = 0; e < DIM; ++e) {
        double v = adj[e][0] * nrm[0];
#pragma unroll
        for (int k = 1; k < DIM; ++k) v = fma(adj[e][k], nrm[k], v);
        m[e] = v * inv_det;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double v = m[0] * g[i][0];
#pragma unroll
        for (int e = 1; e < DIM; ++e) v = fma(m[e], g[i][e], v);
        dn[i] = v;
      }
    }
    if constexpr (T::FORM_ID == CQ_FORM_NITSCHE) {
      // s( -dn_j psi_i - psi_j dn_i + gamma psi_i psi_j ) = psi_i u_j - e_i psi_j
      const double sg = s * gamma;
      double e_[N], u_[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        e_[i] = s * dn[i];
        u_[i] = fma(sg, psi[i], -e_[i]);
      }
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) {
          double v = fma(psi[i], u_[j], acc[tri_index(N, i, j)]);
          acc[tri_index(N, i, j)] = fma(-e_[i], psi[j], v);
        }
    } else {
      // data function(s): f_a (and grad f_a), coefficient uh_a (and grad uh_a)
      double x[DIM];
      if constexpr (T::NEED_X) geo.position(xi, x);
      double uh[BS], guh[BS][DIM], f[BS], gf[BS][DIM];
      constexpr bool GRADF = T::FORM_ID == CQ_FORM_H10_ERR;
      if constexpr (T::NEED_COEF) {
#pragma unroll
        for (int a = 0; a < BS; ++a) {
          if constexpr (!GRADF) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) v = fma(coef[k * BS + a], psi[k], v);
            uh[a] = v;
          } else {
            double r[DIM];  // reference gradient of uh_a
#pragma unroll
            for (int e = 0; e < DIM; ++e) r[e] = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k)
#pragma unroll
              for (int e = 0; e < DIM; ++e) r[e] = fma(coef[k * BS + a], g[k][e], r[e]);
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
              double v = adj[0][k] * r[0];
#pragma unroll
              for (int e = 1; e < DIM; ++e) v = fma(adj[e][k], r[e], v);
              guh[a][k] = v * inv_det;
            }
          }
        }
      }
#pragma unroll
      for (int a = 0; a < BS; ++a) {
        if constexpr (T::FN == CQ_FN_COEF) {
          f[a] = uh[a];
        } else {
          eval_fn<T::FN, DIM, GRADF>(p, a, x, f[a], gf[a]);
        }
      }
      if constexpr (T::FORM_ID == CQ_FORM_LOAD) {
#pragma unroll
        for (int a = 0; a < BS; ++a) {
          const double sf = s * f[a];
#pragma unroll
          for (int i = 0; i < N; ++i) acc[i * BS + a] = fma(sf, psi[i], acc[i * BS + a]);
        }
      } else if constexpr (T::FORM_ID == CQ_FORM_NITSCHE_RHS) {
        // s g_a ( -dn_i + gamma psi_i )
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double r = s * fma(gamma, psi[i], -dn[i]);
#pragma unroll
          for (int a = 0; a < BS; ++a) acc[i * BS + a] = fma(f[a], r, acc[i * BS + a]);
        }
      } else if constexpr (T::FORM_ID == CQ_FORM_INTEGRAL) {
        acc[0] = fma(s, f[0], acc[0]);
      } else if constexpr (T::FORM_ID == CQ_FORM_INTEGRAL_SQ) {
#pragma unroll
        for (int a = 0; a < BS; ++a) acc[0] = fma(s * f[a], f[a], acc[0]);
      } else if constexpr (T::FORM_ID == CQ_FORM_L2_ERR) {
#pragma unroll
        for (int a = 0; a < BS; ++a) {
          const double e = uh[a] - f[a];
          acc[0] = fma(s * e, e, acc[0]);
        }
      } else if constexpr (T::FORM_ID == CQ_FORM_H10_ERR) {
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
          for (int k = 0; k < DIM; ++k) {
            const double e = guh[a][k] - gf[a][k];
            acc[0] = fma(s * e, e, acc[0]);
          }
      }
    }
  }
}

}  // namespace cq
