// cq_common.cuh -- element tables, 1-D Lagrange bases, geometry pull-back, data functions.
// sm_100a only.  All arithmetic is FP64 (the reference's PETSc.ScalarType, setup_types.py:53-67).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cq_b200.h"

namespace cq {

#define CQ_DEVINL __device__ __forceinline__
#define CQ_HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Local dof lattice of Q_k on quad/hex in the frame the runtime points use.
//
// call_basix returns FE[q][j] = phi_perm[j](xi_q) (reference customquad/call_basix.h:60-81): for
// quads perm = {0,2,1,3} swaps basix's two axes (local vertex 1 of a dolfinx create_rectangle cell
// is the +y neighbour, so the points' axis 0 is physical x: SURVEY.md A.2; pinned by
// test/test_linear_basis.py:181), hexes use the identity.  T[j][d] is the lattice coordinate
// (0..k; node xi = T/k) of psi_j = phi_perm[j] along the points' axis d.  Degree-2 entity order:
// vertices, edges, (faces,) interior (test/create_high_order_mesh.py:46-64,153-217); the degree-2
// quad permutation is the corrected {0,2,1,3,5,4,7,6,8} (see oracle/cq_oracle.c, call_basix_perm).
// ---------------------------------------------------------------------------------------------
template <int DIM, int DEG>
struct Lattice;

template <>
struct Lattice<2, 1> {
  static constexpr int N = 4;
  CQ_HD static constexpr int t(int j, int d) {
    constexpr int T[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
    return T[j][d];
  }
};
template <>
struct Lattice<2, 2> {
  static constexpr int N = 9;
  CQ_HD static constexpr int t(int j, int d) {
    // basix quad entity order (0,0),(2,0),(0,2),(2,2),(1,0),(0,1),(2,1),(1,2),(1,1) with axes swapped
    constexpr int T[9][2] = {{0, 0}, {0, 2}, {2, 0}, {2, 2}, {0, 1}, {1, 0}, {1, 2}, {2, 1}, {1, 1}};
    return T[j][d];
  }
};
template <>
struct Lattice<3, 1> {
  static constexpr int N = 8;
  CQ_HD static constexpr int t(int j, int d) { return (j >> d) & 1; }
};
template <>
struct Lattice<3, 2> {
  static constexpr int N = 27;
  CQ_HD static constexpr int t(int j, int d) {
    constexpr int T[27][3] = {
        {0, 0, 0}, {2, 0, 0}, {0, 2, 0}, {2, 2, 0}, {0, 0, 2}, {2, 0, 2}, {0, 2, 2}, {2, 2, 2},  // vertices
        {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 1, 0}, {2, 0, 1}, {1, 2, 0},                        // edges 0-5
        {0, 2, 1}, {2, 2, 1}, {1, 0, 2}, {0, 1, 2}, {2, 1, 2}, {1, 2, 2},                        // edges 6-11
        {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {2, 1, 1}, {1, 2, 1}, {1, 1, 2},                        // faces
        {1, 1, 1}};
    return T[j][d];
  }
};

// 1-D equispaced Lagrange basis by lattice coordinate (lagrange_variant::unset, call_basix.h:30).
template <int DEG>
struct Basis1D;
template <>
struct Basis1D<1> {
  double v[2], d[2];
  CQ_DEVINL explicit Basis1D(double x) {
    v[0] = 1.0 - x; v[1] = x;
    d[0] = -1.0; d[1] = 1.0;
  }
};
template <>
struct Basis1D<2> {
  double v[3], d[3];
  CQ_DEVINL explicit Basis1D(double x) {
    const double y = 1.0 - x;
    v[0] = y * (1.0 - 2.0 * x);
    v[1] = 4.0 * x * y;
    v[2] = x * (2.0 * x - 1.0);
    d[0] = 4.0 * x - 3.0;
    d[1] = 4.0 - 8.0 * x;
    d[2] = 4.0 * x - 1.0;
  }
};

// ---------------------------------------------------------------------------------------------
// Per-cell geometry: the Q1 coordinate map x(xi) = sum_v x_v psi_v(xi) in
// multilinear monomial form x = c_0 + sum_m c_m * prod_{d in m} xi_d, m = 1..2^DIM-1 (bit d of m
// set <=> xi_d in the monomial).  The coefficients are differences of vertex coordinates, so the
// Jacobian of an (almost) axis-aligned cell carries no cancellation error from the absolute
// coordinates.  J is evaluated at every point (the reference's generated code does the same for
// its non-affine cell types).
// ---------------------------------------------------------------------------------------------
template <int DIM>
struct CellGeom {
  static constexpr int NV = 1 << DIM;
  double c[NV][DIM];  // c[m][component]
  double hK;          // ufl.CellDiameter: largest vertex-vertex distance (demo/poisson.py:143)
  // Affine cell (all multilinear coefficients of order >= 2 are exactly zero, as for every cell of
  // a create_rectangle / create_box mesh): J is the same at every point, so J, det J and adj J are
  // evaluated once per cell (init_affine) instead of once per point.  Same numbers: the per-point
  // formula reduces to J[k][e] = c[1<<e][k] when the higher coefficients vanish.
  bool affine;
  double det0, inv_det0, adj0[DIM][DIM];

  // X[v][comp] with v in the points' frame lattice order of Lattice<DIM,1>
  CQ_DEVINL void init(const double (&X)[NV][DIM], bool need_h) {
    // Moebius transform over the vertex lattice: c_m = sum_{s subset m} (-1)^{|m|-|s|} x_{vertex(s)}.
    // vertex(s): the Q1 local dof whose lattice coordinate along axis d is bit d of s.
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      double tmp[NV];
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        int v = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          bool same = true;
#pragma unroll
          for (int d = 0; d < DIM; ++d) same = same && (Lattice<DIM, 1>::t(j, d) == ((m >> d) & 1));
          if (same) v = j;
        }
        tmp[m] = X[v][k];
      }
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int m = 0; m < NV; ++m)
          if (m & (1 << d)) tmp[m] -= tmp[m ^ (1 << d)];
#pragma unroll
      for (int m = 0; m < NV; ++m) c[m][k] = tmp[m];
    }
    affine = true;
#pragma unroll
    for (int m = 0; m < NV; ++m)
      if ((m & (m - 1)) != 0)  // more than one bit set: bilinear / trilinear terms
#pragma unroll
        for (int k = 0; k < DIM; ++k) affine = affine && (c[m][k] == 0.0);
    hK = 0.0;
    if (need_h) {
      double h2 = 0.0;
#pragma unroll
      for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int b = a + 1; b < NV; ++b) {
          double d2 = 0.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k) {
            const double t = X[a][k] - X[b][k];
            d2 = fma(t, t, d2);
          }
          h2 = fmax(h2, d2);
        }
      hK = sqrt(h2);
    }
  }

  // J[comp][e] = d x_comp / d xi_e
  CQ_DEVINL void jacobian(const double (&xi)[DIM], double (&J)[DIM][DIM]) const {
    if constexpr (DIM == 2) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        J[k][0] = fma(c[3][k], xi[1], c[1][k]);
        J[k][1] = fma(c[3][k], xi[0], c[2][k]);
      }
    } else {
      const double yz = xi[1] * xi[2], xz = xi[0] * xi[2], xy = xi[0] * xi[1];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        J[k][0] = fma(c[7][k], yz, fma(c[5][k], xi[2], fma(c[3][k], xi[1], c[1][k])));
        J[k][1] = fma(c[7][k], xz, fma(c[6][k], xi[2], fma(c[3][k], xi[0], c[2][k])));
        J[k][2] = fma(c[7][k], xy, fma(c[6][k], xi[1], fma(c[5][k], xi[0], c[4][k])));
      }
    }
  }

  CQ_DEVINL void position(const double (&xi)[DIM], double (&x)[DIM]) const {
    if constexpr (DIM == 2) {
#pragma unroll
      for (int k = 0; k < 2; ++k) x[k] = fma(fma(c[3][k], xi[1], c[1][k]), xi[0], fma(c[2][k], xi[1], c[0][k]));
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double a0 = fma(c[4][k], xi[2], c[0][k]);   // 1
        const double a1 = fma(c[5][k], xi[2], c[1][k]);   // x
        const double a2 = fma(c[6][k], xi[2], c[2][k]);   // y
        const double a3 = fma(c[7][k], xi[2], c[3][k]);   // xy
        x[k] = fma(fma(a3, xi[1], a1), xi[0], fma(a2, xi[1], a0));
      }
    }
  }
};

// det J and adj J (K = adj / det):  adj[e][k] = det * d xi_e / d x_k
template <int DIM>
CQ_DEVINL double det_adj(const double (&J)[DIM][DIM], double (&adj)[DIM][DIM]) {
  if constexpr (DIM == 2) {
    adj[0][0] = J[1][1]; adj[0][1] = -J[0][1];
    adj[1][0] = -J[1][0]; adj[1][1] = J[0][0];
    return fma(J[0][0], J[1][1], -J[0][1] * J[1][0]);
  } else {
    adj[0][0] = fma(J[1][1], J[2][2], -J[1][2] * J[2][1]);
    adj[1][0] = fma(J[1][2], J[2][0], -J[1][0] * J[2][2]);
    adj[2][0] = fma(J[1][0], J[2][1], -J[1][1] * J[2][0]);
    adj[0][1] = fma(J[0][2], J[2][1], -J[0][1] * J[2][2]);
    adj[1][1] = fma(J[0][0], J[2][2], -J[0][2] * J[2][0]);
    adj[2][1] = fma(J[0][1], J[2][0], -J[0][0] * J[2][1]);
    adj[0][2] = fma(J[0][1], J[1][2], -J[0][2] * J[1][1]);
    adj[1][2] = fma(J[0][2], J[1][0], -J[0][0] * J[1][2]);
    adj[2][2] = fma(J[0][0], J[1][1], -J[0][1] * J[1][0]);
    return fma(J[0][0], adj[0][0], fma(J[0][1], adj[1][0], J[0][2] * adj[2][0]));
  }
}

// per-cell constants of an affine cell (CellGeom::affine): det J, 1/det J, adj J
template <int DIM>
CQ_DEVINL void init_affine(CellGeom<DIM>& geo) {
  if (geo.affine) {
    double J[DIM][DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k)
#pragma unroll
      for (int e = 0; e < DIM; ++e) J[k][e] = geo.c[1 << e][k];
    geo.det0 = det_adj<DIM>(J, geo.adj0);
    geo.inv_det0 = 1.0 / geo.det0;
  }
}

template <int DIM>
CQ_DEVINL double det_only(const double (&J)[DIM][DIM]) {
  if constexpr (DIM == 2) {
    return fma(J[0][0], J[1][1], -J[0][1] * J[1][0]);
  } else {
    const double c0 = fma(J[1][1], J[2][2], -J[1][2] * J[2][1]);
    const double c1 = fma(J[1][2], J[2][0], -J[1][0] * J[2][2]);
    const double c2 = fma(J[1][0], J[2][1], -J[1][1] * J[2][0]);
    return fma(J[0][0], c0, fma(J[0][1], c1, J[0][2] * c2));
  }
}

// ---------------------------------------------------------------------------------------------
// Closed-form data functions (cq_fn) of the physical coordinate and their gradients.
// ---------------------------------------------------------------------------------------------
struct FnParams {
  double p[CQ_MAX_CONSTS];
};

template <int FN, int DIM, bool GRAD>
CQ_DEVINL void eval_fn(const double* __restrict__ p, int a, const double (&x)[DIM], double& f, double (&g)[DIM]) {
  if constexpr (FN == CQ_FN_CONST) {
    f = p[a];
    if constexpr (GRAD)
#pragma unroll
      for (int d = 0; d < DIM; ++d) g[d] = 0.0;
  } else if constexpr (FN == CQ_FN_AFFINE) {
    f = p[4 * a];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      f = fma(p[4 * a + 1 + d], x[d], f);
      if constexpr (GRAD) g[d] = p[4 * a + 1 + d];
    }
  } else {  // CQ_FN_SINPROD
    double s[DIM], c[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) sincospi(x[d], &s[d], &c[d]);
    double v = p[a];
#pragma unroll
    for (int d = 0; d < DIM; ++d) v *= s[d];
    f = v;
    if constexpr (GRAD) {
#pragma unroll
      for (int e = 0; e < DIM; ++e) {
        double t = p[a] * 3.14159265358979323846 * c[e];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
          if (d != e) t *= s[d];
        g[e] = t;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
CQ_DEVINL void red_add_f64(double* addr, double v) {
  // REDG.E.ADD.F64.RN.STRONG.GPU: fire-and-forget FP64 reduction at L2 (no return value)
#if defined(CQ_PROFILING) && defined(CQ_EXPERIMENT_NORED)  // profiling builds only (profiles/build_variant.sh): value is formed, RED never issues
  if (v == 1.2345e300) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
#elif defined(CQ_PROFILING) && defined(CQ_EXPERIMENT_ST)   // profiling builds only: plain store instead of the reduction (wrong results)
  asm volatile("st.global.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
#else
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
#endif
}

CQ_DEVINL double shfl_xor_f64(double v, int lane_mask, unsigned mask = 0xffffffffu) {
  return __shfl_xor_sync(mask, v, lane_mask);
}

CQ_DEVINL double ld_nc_f64(const double* p) { return __ldg(p); }

// upper-triangular (row-major, i <= j) enumeration of an n x n symmetric matrix
CQ_HD constexpr int tri_count(int n) { return n * (n + 1) / 2; }
CQ_HD constexpr int tri_index(int n, int i, int j) { return i * n - i * (i - 1) / 2 + (j - i); }

}  // namespace cq
