/*
 * cq_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of customquad's runtime-quadrature assembly path, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker and the timed CPU baseline.  Nothing under customquad_b200/ may import, link or
 * call this file.
 *
 * PARITY STATUS (see DESIGN.md "Oracle"):  the reference itself cannot be imported or
 * compiled here -- its arithmetic lives in un-vendored dependencies that are absent from
 * /root/reference and from this image:
 *     ffcx-custom @ august/2023cq (generates tabulate_tensor_runtime_float64; docker/Dockerfile:5)
 *     basix 0.6.x from dolfinx/dolfinx:v0.6.0-r1 (Lagrange tabulation; docker/Dockerfile:1)
 *     dolfinx 0.6.0 / PETSc (sparsity, MatSetValuesLocal; assemble_matrix.py:19,103-111)
 * so this file restates their published algorithms (equispaced Lagrange Q1/Q2 on basix's
 * reference quadrilateral/hexahedron; UFL's pull-back) and anchors parity on the reference's
 * own call sites and on every known answer its test-suite holds for this path (tests/golden/,
 * tests/test_oracle_reference_cases.py).  What those tests pin: Q1 quads and hexes, mass-type
 * integrands, SpatialCoordinate, subset/repeated cells, weight scaling.  PARITY UNPINNED (no
 * passing reference test exists; the oracle is the mathematical definition):
 * degree 2, gradients (stiffness), normals / Nitsche, CellDiameter, blocked spaces, ragged nq.
 *
 * Each function cites the reference lines (paths into /root/reference) it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cq_b200.h" /* enums and constants only */

#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXN 27  /* Q2 hex */
#define MAXV 8   /* hex vertices */
#define MAXBS 3

/* ---------------------------------------------------------------------------------------
 * basix Lagrange tabulation (restated: basix is not vendored).  Dof order = basix entity
 * order (vertices, edges, faces, interior), lattice coordinate r_j in units of 1/degree
 * along the reference axes; cf. reference test/create_high_order_mesh.py:46-64,153-217.
 * --------------------------------------------------------------------------------------- */
static const int HEX_EDGES[12][2] = {{0,1},{0,2},{0,4},{1,3},{1,5},{2,3},{2,6},{3,7},{4,5},{4,6},{5,7},{6,7}};
static const int HEX_FACES[6][4] = {{0,1,2,3},{0,1,4,5},{0,2,4,6},{1,3,5,7},{2,3,6,7},{4,5,6,7}};
static const int QUAD_EDGES[4][2] = {{0,1},{0,2},{1,3},{2,3}};

static int basix_lattice(int cell_type, int degree, int r[MAXN][3]) {
  const int tdim = cell_type == CQ_CELL_QUADRILATERAL ? 2 : 3;
  const int nv = 1 << tdim;
  int v[MAXV][3];
  for (int i = 0; i < nv; ++i)
    for (int a = 0; a < 3; ++a) v[i][a] = a < tdim ? ((i >> a) & 1) * degree : 0;
  int n = 0;
  for (int i = 0; i < nv; ++i, ++n)
    for (int a = 0; a < 3; ++a) r[n][a] = v[i][a];
  if (degree == 2) {
    const int ne = tdim == 2 ? 4 : 12;
    for (int e = 0; e < ne; ++e, ++n) {
      const int* ed = tdim == 2 ? QUAD_EDGES[e] : HEX_EDGES[e];
      for (int a = 0; a < 3; ++a) r[n][a] = (v[ed[0]][a] + v[ed[1]][a]) / 2;
    }
    if (tdim == 3)
      for (int f = 0; f < 6; ++f, ++n)
        for (int a = 0; a < 3; ++a)
          r[n][a] = (v[HEX_FACES[f][0]][a] + v[HEX_FACES[f][1]][a] + v[HEX_FACES[f][2]][a] + v[HEX_FACES[f][3]][a]) / 4;
    for (int a = 0; a < 3; ++a) r[n][a] = a < tdim ? 1 : 0;
    ++n;
  }
  return n;
}

/* 1-D equispaced Lagrange basis attached to lattice coordinate t (lagrange_variant::unset,
 * call_basix.h:30; equispaced == GLL for degree <= 2). */
static void lagrange_1d(int degree, int t, double x, double* v, double* dv) {
  if (degree == 1) {
    if (t == 0) { *v = 1.0 - x; *dv = -1.0; } else { *v = x; *dv = 1.0; }
  } else if (t == 0) { *v = (1.0 - x) * (1.0 - 2.0 * x); *dv = 4.0 * x - 3.0; }
  else if (t == 1) { *v = 4.0 * x * (1.0 - x); *dv = 4.0 - 8.0 * x; }
  else { *v = x * (2.0 * x - 1.0); *dv = 4.0 * x - 1.0; }
}

/* FiniteElement::tabulate(nd, pts, {nq, gdim}) for a Lagrange element: tab[deriv][q][dof],
 * deriv 0 = value, 1 + e = d/dxi_e (call_basix.h:34-39). */
static void basix_tabulate(int cell_type, int degree, int nd, int nq, const double* pts, int gdim,
                           double* tab /* [(nd ? 1+gdim : 1)][nq][nbf] */) {
  int r[MAXN][3];
  const int nbf = basix_lattice(cell_type, degree, r);
  for (int q = 0; q < nq; ++q) {
    double v[3][3], dv[3][3];
    for (int d = 0; d < gdim; ++d)
      for (int t = 0; t <= degree; ++t) lagrange_1d(degree, t, pts[gdim * q + d], &v[d][t], &dv[d][t]);
    for (int j = 0; j < nbf; ++j) {
      double val = 1.0;
      for (int d = 0; d < gdim; ++d) val *= v[d][r[j][d]];
      tab[(size_t)q * nbf + j] = val;
      if (nd)
        for (int e = 0; e < gdim; ++e) {
          double g = 1.0;
          for (int d = 0; d < gdim; ++d) g *= (d == e) ? dv[d][r[j][d]] : v[d][r[j][d]];
          tab[((size_t)(1 + e) * nq + q) * nbf + j] = g;
        }
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * call_basix (customquad/call_basix.h:14-81): builds FE[0][0][q][j] = tab(deriv, q, perm[j], 0).
 *
 * Permutation table (call_basix.h:60-74): quad degree 1 {0,2,1,3}; hex identity.  For quad
 * degree 2 the reference's entry {0,2,6,8,1,3,5,7,4} treats basix's entity order as a
 * lexicographic grid, which it is not; every degree-2 test of the reference is xfail
 * (test/test_high_order_basis.py:30,54,67).  The intended invariant -- a standard rule must
 * reproduce standard assembly (test_high_order_basis.py:39) -- needs the same x<->y swap as
 * degree 1, which in entity order is {0,2,1,3,5,4,7,6,8}.  That is what is used here and on
 * the GPU ("do not copy it").
 * --------------------------------------------------------------------------------------- */
static int call_basix_perm(int cell_type, int degree, int perm[MAXN]) {
  static const int q1[4] = {0, 2, 1, 3};
  static const int q2[9] = {0, 2, 1, 3, 5, 4, 7, 6, 8};
  if (cell_type == CQ_CELL_QUADRILATERAL) {
    if (degree == 1) { memcpy(perm, q1, sizeof q1); return 4; }
    if (degree == 2) { memcpy(perm, q2, sizeof q2); return 9; }
  } else if (cell_type == CQ_CELL_HEXAHEDRON) {
    const int n = degree == 1 ? 8 : 27;
    if (degree == 1 || degree == 2) { for (int i = 0; i < n; ++i) perm[i] = i; return n; }
  }
  return 0; /* reference: assert(perm.size() == num_basis_functions) fails (call_basix.h:76) */
}

void call_basix(double***** FE, int num_quadrature_points, const double* quadrature_points,
                int basix_derivative, int family, int cell_type, int degree, int lattice_type, int gdim) {
  (void)family; (void)lattice_type;
  const int nd = basix_derivative == 0 ? 0 : 1; /* call_basix.h:34 */
  int perm[MAXN];
  const int nbf = call_basix_perm(cell_type, degree, perm);
  const int nq = num_quadrature_points;
  double* tab = (double*)malloc(sizeof(double) * (size_t)(nd ? 1 + gdim : 1) * (nq > 0 ? nq : 1) * nbf);
  basix_tabulate(cell_type, degree, nd, nq, quadrature_points, gdim, tab);
  /* Allocate (call_basix.h:49-57): nq + 3 heap blocks per table */
  *FE = (double****)malloc(sizeof(double***));
  (*FE)[0] = (double***)malloc(sizeof(double**));
  (*FE)[0][0] = (double**)malloc(sizeof(double*) * (nq > 0 ? nq : 1));
  for (int k = 0; k < nq; ++k) (*FE)[0][0][k] = (double*)malloc(sizeof(double) * nbf);
  /* Copy with permutation (call_basix.h:79-81) */
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < nbf; ++j) (*FE)[0][0][i][j] = tab[((size_t)basix_derivative * nq + i) * nbf + perm[j]];
  free(tab);
}

/* The reference never frees its tables; the oracle does. */
static void free_basix_table(double**** FE, int nq) {
  for (int k = 0; k < nq; ++k) free(FE[0][0][k]);
  free(FE[0][0]); free(FE[0]); free(FE);
}

/* Allocation-free flavour of the same table ("optimised CPU" baseline). */
static void basix_table_flat(double* out /* [nq][nbf] */, int nq, const double* pts, int deriv, int cell_type,
                             int degree, int gdim, double* scratch) {
  int perm[MAXN];
  const int nbf = call_basix_perm(cell_type, degree, perm);
  basix_tabulate(cell_type, degree, deriv != 0, nq, pts, gdim, scratch);
  for (int i = 0; i < nq; ++i)
    for (int j = 0; j < nbf; ++j) out[(size_t)i * nbf + j] = scratch[((size_t)deriv * nq + i) * nbf + perm[j]];
}

/* ---------------------------------------------------------------------------------------
 * The generated kernel (README.md:133-158).  ffcx-custom bakes form, element and geometry
 * into each generated function; here one generic body reads them from a thread-local
 * "compiled form" record that cq_oracle_select_kernel() binds, and the exported function
 * pointer keeps exactly the generated ABI.
 * --------------------------------------------------------------------------------------- */
typedef struct oracle_cfg {
  int form, cell_type, degree, bs, fn_kind, gdim;
  int faithful; /* 1: one call_basix (heap tables) per table per cell, like the reference */
} oracle_cfg;

static _Thread_local oracle_cfg g_cfg;

typedef void (*tabulate_tensor_runtime_fn)(double* A, const double* w, const double* c,
                                           const double* coordinate_dofs, const int* entity_local_index,
                                           const uint8_t* quadrature_permutation, int num_quadrature_points,
                                           const double* quadrature_points, const double* quadrature_weights,
                                           const double* quadrature_normals);

static void eval_fn(int kind, const double* p, int a, int gdim, const double x[3], double* f, double g[3]) {
  g[0] = g[1] = g[2] = 0.0;
  if (kind == CQ_FN_CONST) { *f = p[a]; }
  else if (kind == CQ_FN_AFFINE) {
    *f = p[4 * a];
    for (int d = 0; d < gdim; ++d) { *f += p[4 * a + 1 + d] * x[d]; g[d] = p[4 * a + 1 + d]; }
  } else { /* CQ_FN_SINPROD: demo/poisson.py:59-67 */
    double s[3], c[3];
    for (int d = 0; d < gdim; ++d) { s[d] = sin(M_PI * x[d]); c[d] = cos(M_PI * x[d]); }
    double v = p[a];
    for (int d = 0; d < gdim; ++d) v *= s[d];
    *f = v;
    for (int e = 0; e < gdim; ++e) {
      double t = p[a] * M_PI * c[e];
      for (int d = 0; d < gdim; ++d) if (d != e) t *= s[d];
      g[e] = t;
    }
  }
}

void tabulate_tensor_runtime_float64(double* A, const double* w, const double* c, const double* coordinate_dofs,
                                     const int* entity_local_index, const uint8_t* quadrature_permutation,
                                     int num_quadrature_points, const double* quadrature_points,
                                     const double* quadrature_weights, const double* quadrature_normals) {
  (void)entity_local_index; (void)quadrature_permutation;
  const oracle_cfg cfg = g_cfg;
  const int gdim = cfg.gdim, nq = num_quadrature_points, bs = cfg.bs;
  const int nv = 1 << gdim;
  const int n = cfg.degree == 1 ? nv : (gdim == 2 ? 9 : 27);
  const int nl = n * bs;
  const double* weights = quadrature_weights; /* README.md:144-145 */
  if (nq <= 0) return;

  /* Tables: FE value + one per derivative for the form element, and the same for the Q1
   * coordinate element -- "one call per table per cell" (README.md:152-156). */
  double *FE = (double*)malloc(sizeof(double) * nq * n), *dFE[3] = {0, 0, 0};
  double *XE = (double*)malloc(sizeof(double) * nq * nv), *dXE[3] = {0, 0, 0};
  double* scratch = (double*)malloc(sizeof(double) * (1 + gdim) * nq * MAXN);
  for (int e = 0; e < gdim; ++e) { dFE[e] = (double*)malloc(sizeof(double) * nq * n); dXE[e] = (double*)malloc(sizeof(double) * nq * nv); }
  if (cfg.faithful) {
    double**** T;
    for (int t = 0; t <= gdim; ++t) {
      call_basix(&T, nq, quadrature_points, t, 1, cfg.cell_type, cfg.degree, 0, gdim);
      for (int q = 0; q < nq; ++q) memcpy((t ? dFE[t - 1] : FE) + (size_t)q * n, T[0][0][q], sizeof(double) * n);
      free_basix_table(T, nq);
      call_basix(&T, nq, quadrature_points, t, 1, cfg.cell_type, 1, 0, gdim);
      for (int q = 0; q < nq; ++q) memcpy((t ? dXE[t - 1] : XE) + (size_t)q * nv, T[0][0][q], sizeof(double) * nv);
      free_basix_table(T, nq);
    }
  } else {
    for (int t = 0; t <= gdim; ++t) {
      basix_table_flat(t ? dFE[t - 1] : FE, nq, quadrature_points, t, cfg.cell_type, cfg.degree, gdim, scratch);
      basix_table_flat(t ? dXE[t - 1] : XE, nq, quadrature_points, t, cfg.cell_type, 1, gdim, scratch);
    }
  }

  /* ufl.CellDiameter: largest distance between two vertices of the cell */
  double h = 0.0;
  for (int a = 0; a < nv; ++a)
    for (int b = a + 1; b < nv; ++b) {
      double d2 = 0.0;
      for (int d = 0; d < gdim; ++d) { const double t = coordinate_dofs[3 * a + d] - coordinate_dofs[3 * b + d]; d2 += t * t; }
      if (d2 > h) h = d2;
    }
  h = sqrt(h);
  const int nitsche = cfg.form == CQ_FORM_NITSCHE || cfg.form == CQ_FORM_NITSCHE_RHS;
  const double gamma = nitsche ? c[0] / h : 0.0;   /* betaN / h, demo/poisson.py:157,159 */
  const double* p = nitsche ? c + 1 : c;           /* closed-form function parameters */

  for (int iq = 0; iq < nq; ++iq) {
    /* geometry: x = sum_v x_v psi_v, J = sum_v x_v (x) grad psi_v (non-affine: every point) */
    double J[3][3] = {{0}}, x[3] = {0, 0, 0};
    for (int v = 0; v < nv; ++v)
      for (int d = 0; d < gdim; ++d) {
        x[d] += coordinate_dofs[3 * v + d] * XE[(size_t)iq * nv + v];
        for (int e = 0; e < gdim; ++e) J[d][e] += coordinate_dofs[3 * v + d] * dXE[e][(size_t)iq * nv + v];
      }
    double detJ, K[3][3];
    if (gdim == 2) {
      detJ = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      K[0][0] = J[1][1] / detJ; K[0][1] = -J[0][1] / detJ; K[1][0] = -J[1][0] / detJ; K[1][1] = J[0][0] / detJ;
    } else {
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                   c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      detJ = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      K[0][0] = c00 / detJ; K[1][0] = c01 / detJ; K[2][0] = c02 / detJ;
      K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / detJ;
      K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / detJ;
      K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / detJ;
      K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / detJ;
      K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / detJ;
      K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / detJ;
    }
    const double s = weights[iq] * fabs(detJ);
    const double* psi = FE + (size_t)iq * n;
    double G[MAXN][3], dn[MAXN]; /* physical gradients, and n . grad psi_i */
    for (int i = 0; i < n; ++i) {
      dn[i] = 0.0;
      for (int d = 0; d < gdim; ++d) {
        double g = 0.0;
        for (int e = 0; e < gdim; ++e) g += K[e][d] * dFE[e][(size_t)iq * n + i];
        G[i][d] = g;
        if (nitsche) dn[i] += quadrature_normals[gdim * iq + d] * g;
      }
    }
    /* data function(s) at the point */
    double f[MAXBS] = {0, 0, 0}, gf[MAXBS][3] = {{0}}, uh[MAXBS] = {0, 0, 0}, guh[MAXBS][3] = {{0}};
    const int need_coef = cfg.fn_kind == CQ_FN_COEF || cfg.form == CQ_FORM_L2_ERR || cfg.form == CQ_FORM_H10_ERR;
    const int rank = cfg.form < 10 ? 2 : (cfg.form < 20 ? 1 : 0);
    if (rank < 2) {
      if (need_coef)
        for (int a = 0; a < bs; ++a)
          for (int k = 0; k < n; ++k) {
            uh[a] += w[k * bs + a] * psi[k];
            for (int d = 0; d < gdim; ++d) guh[a][d] += w[k * bs + a] * G[k][d];
          }
      for (int a = 0; a < bs; ++a) {
        if (cfg.fn_kind == CQ_FN_COEF) { f[a] = uh[a]; for (int d = 0; d < 3; ++d) gf[a][d] = guh[a][d]; }
        else eval_fn(cfg.fn_kind, p, a, gdim, x, &f[a], gf[a]);
      }
    }
    switch (cfg.form) {
      case CQ_FORM_MASS:
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            const double v = s * psi[i] * psi[j];
            for (int a = 0; a < bs; ++a) A[(size_t)(i * bs + a) * nl + j * bs + a] += v;
          }
        break;
      case CQ_FORM_STIFFNESS:
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            double v = 0.0;
            for (int d = 0; d < gdim; ++d) v += G[i][d] * G[j][d];
            v *= s;
            for (int a = 0; a < bs; ++a) A[(size_t)(i * bs + a) * nl + j * bs + a] += v;
          }
        break;
      case CQ_FORM_NITSCHE:
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            const double v = s * (-dn[j] * psi[i] - psi[j] * dn[i] + gamma * psi[i] * psi[j]);
            for (int a = 0; a < bs; ++a) A[(size_t)(i * bs + a) * nl + j * bs + a] += v;
          }
        break;
      case CQ_FORM_LOAD:
        for (int i = 0; i < n; ++i)
          for (int a = 0; a < bs; ++a) A[i * bs + a] += s * f[a] * psi[i];
        break;
      case CQ_FORM_NITSCHE_RHS:
        for (int i = 0; i < n; ++i)
          for (int a = 0; a < bs; ++a) A[i * bs + a] += s * f[a] * (-dn[i] + gamma * psi[i]);
        break;
      case CQ_FORM_INTEGRAL:
        A[0] += s * f[0];
        break;
      case CQ_FORM_INTEGRAL_SQ:
        for (int a = 0; a < bs; ++a) A[0] += s * f[a] * f[a];
        break;
      case CQ_FORM_L2_ERR:
        for (int a = 0; a < bs; ++a) A[0] += s * (uh[a] - f[a]) * (uh[a] - f[a]);
        break;
      case CQ_FORM_H10_ERR:
        for (int a = 0; a < bs; ++a)
          for (int d = 0; d < gdim; ++d) A[0] += s * (guh[a][d] - gf[a][d]) * (guh[a][d] - gf[a][d]);
        break;
      default: break;
    }
  }
  free(FE); free(XE); free(scratch);
  for (int e = 0; e < gdim; ++e) { free(dFE[e]); free(dXE[e]); }
}

/* "JIT": bind the form and hand back the kernel, like
 * form.ufcx_form.integrals(cell)[i].tabulate_tensor_runtime_float64 (assemble_matrix.py:27-30). */
tabulate_tensor_runtime_fn cq_oracle_select_kernel(int form, int cell_type, int degree, int bs, int fn_kind, int faithful) {
  oracle_cfg cfg = {form, cell_type, degree, bs, fn_kind, cell_type == CQ_CELL_QUADRILATERAL ? 2 : 3, faithful};
  g_cfg = cfg;
  return tabulate_tensor_runtime_float64;
}

/* ---------------------------------------------------------------------------------------
 * MatSetValuesLocal(A, nrow, irow, ncol, icol, y, ADD_VALUES) on a sorted-column CSR
 * (customquad/setup_types.py:106-113; call site assemble_matrix.py:103-111).
 * Returns the number of (row, col) pairs missing from the pattern.
 * --------------------------------------------------------------------------------------- */
static int mat_set_values_local_add(const int64_t* indptr, const int32_t* indices, double* values, int nrow,
                                    const int32_t* irow, int ncol, const int32_t* icol, const double* y, int atomic) {
  int missing = 0;
  for (int i = 0; i < nrow; ++i) {
    const int64_t lo0 = indptr[irow[i]], hi0 = indptr[irow[i] + 1];
    for (int j = 0; j < ncol; ++j) {
      int64_t lo = lo0, hi = hi0;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (indices[mid] < icol[j]) lo = mid + 1; else hi = mid; }
      if (lo < hi0 && indices[lo] == icol[j]) {
        if (atomic) {
#pragma omp atomic
          values[lo] += y[(size_t)i * ncol + j];
        } else values[lo] += y[(size_t)i * ncol + j];
      } else ++missing;
    }
  }
  return missing;
}

typedef struct oracle_args {
  int form, cell_type, degree, bs, fn_kind, faithful, nthreads;
  /* utils.get_vertices (utils.py:66-71) */
  const int32_t* vertices; const double* coords;
  /* utils.get_dofs before blocking (utils.py:33-63) */
  const int32_t* dofmap;
  /* packed coefficients [num_cells, ncoef] and constants (assemble_matrix.py:16-17) */
  const double* coeffs; int ncoef; const double* consts;
  /* qr tuple in ragged form; normals NULL for a 3-tuple */
  int64_t num_entries; const int32_t* cells; const int64_t* offsets; const double* pts; const double* w; const double* normals;
} oracle_args;

static int num_local_dofs(const oracle_args* a) {
  const int gdim = a->cell_type == CQ_CELL_QUADRILATERAL ? 2 : 3;
  return a->degree == 1 ? (1 << gdim) : (gdim == 2 ? 9 : 27);
}

/* Common body of the three assemble_cells loops (assemble_matrix.py:50-113,
 * assemble_vector.py:52-94, assemble_scalar.py:37-82). rank selects the sink. */
static int64_t assemble_cells(const oracle_args* a, int rank, const int64_t* indptr, const int32_t* indices,
                              double* out, double* element_out) {
  const int gdim = a->cell_type == CQ_CELL_QUADRILATERAL ? 2 : 3;
  const int nv = 1 << gdim, n = num_local_dofs(a), nl = n * a->bs;
  const int len = rank == 2 ? nl * nl : (rank == 1 ? nl : 1);
  const double* qr_n = a->normals ? a->normals : a->pts; /* "dummy" (assemble_matrix.py:67) */
  int64_t missing = 0;
  int nthreads = a->nthreads > 0 ? a->nthreads : 1;
#ifndef _OPENMP
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads) reduction(+ : missing)
  {
    tabulate_tensor_runtime_fn kernel = cq_oracle_select_kernel(a->form, a->cell_type, a->degree, a->bs, a->fn_kind, a->faithful);
    double cell_coords[MAXV * 3];
    memset(cell_coords, 0, sizeof cell_coords);            /* np.zeros((num_loc_vertices, 3)) */
    double* A_local = (double*)malloc(sizeof(double) * len);
    int32_t pos[MAXN * MAXBS];
    const int entity_local_index[1] = {0};
    const uint8_t perm[1] = {0};                           /* "Don't permute" (assemble_matrix.py:81-82) */
    double m_private = 0.0;
#pragma omp for schedule(static)
    for (int64_t k = 0; k < a->num_entries; ++k) {
      const int32_t cell = a->cells[k];
      for (int i = 0; i < n; ++i)                          /* pos = dofmap[cell, :], blocked (utils.py:59-62) */
        for (int b = 0; b < a->bs; ++b) pos[i * a->bs + b] = a->dofmap ? a->dofmap[(size_t)cell * n + i] * a->bs + b : 0;
      for (int v = 0; v < nv; ++v)                         /* cell_coords[:, :] = coords[vertices[cell, :]] */
        memcpy(cell_coords + 3 * v, a->coords + 3 * (size_t)a->vertices[(size_t)cell * nv + v], 3 * sizeof(double));
      const int nq = (int)(a->offsets[k + 1] - a->offsets[k]);
      memset(A_local, 0, sizeof(double) * len);            /* A_local.fill(0.0) */
      kernel(A_local, a->coeffs ? a->coeffs + (size_t)cell * a->ncoef : NULL, a->consts, cell_coords,
             entity_local_index, perm, nq, a->pts + (size_t)gdim * a->offsets[k], a->w + a->offsets[k],
             qr_n + (size_t)gdim * a->offsets[k]);
      if (element_out) memcpy(element_out + (size_t)k * len, A_local, sizeof(double) * len);
      else if (rank == 2) missing += mat_set_values_local_add(indptr, indices, out, nl, pos, nl, pos, A_local, nthreads > 1);
      else if (rank == 1) {
        for (int j = 0; j < nl; ++j) {                     /* b[dofs[cell, j]] += b_local[j] */
          if (nthreads > 1) {
#pragma omp atomic
            out[pos[j]] += A_local[j];
          } else out[pos[j]] += A_local[j];
        }
      } else m_private += A_local[0];                      /* m[0] += m_local[0] */
    }
    if (rank == 0 && !element_out) {
#pragma omp atomic
      out[0] += m_private;
    }
    free(A_local);
  }
  return missing;
}

int64_t cq_oracle_assemble_matrix(const oracle_args* a, const int64_t* indptr, const int32_t* indices, double* values) {
  return assemble_cells(a, 2, indptr, indices, values, NULL);
}
int64_t cq_oracle_assemble_vector(const oracle_args* a, double* b) { return assemble_cells(a, 1, NULL, NULL, b, NULL); }
int64_t cq_oracle_assemble_scalar(const oracle_args* a, double* m) { return assemble_cells(a, 0, NULL, NULL, m, NULL); }
/* rank: 2 matrix, 1 vector, 0 scalar; out [num_entries, len] */
int64_t cq_oracle_element_tensors(const oracle_args* a, int rank, double* out) { return assemble_cells(a, rank, NULL, NULL, NULL, out); }

/* call_basix on a flat point list: out [nq, nbf] (parity hook for cq_tabulate) */
int cq_oracle_tabulate(int cell_type, int degree, int basix_derivative, int64_t nq, const double* pts, double* out) {
  const int gdim = cell_type == CQ_CELL_QUADRILATERAL ? 2 : 3;
  int perm[MAXN];
  const int nbf = call_basix_perm(cell_type, degree, perm);
  if (!nbf) return -1;
  double**** T;
  call_basix(&T, (int)nq, pts, basix_derivative, 1, cell_type, degree, 0, gdim);
  for (int64_t q = 0; q < nq; ++q) memcpy(out + q * nbf, T[0][0][q], sizeof(double) * nbf);
  free_basix_table(T, (int)nq);
  return nbf;
}

/* ---------------------------------------------------------------------------------------
 * dolfinx.cpp.fem.petsc.create_matrix(form) (assemble_matrix.py:19): the AIJ sparsity pattern of the form's cell
 * integral.  Restated the way dolfinx's SparsityPattern / PETSc preallocation build it: for every cell of the
 * integration domain insert dofs[c] x dofs[c]; a row is a sorted array of distinct columns grown by insertion.
 * Blocked spaces: row r*bs + a holds the columns c*bs + b, b < bs, of every scalar coupling (r, c)
 * (utils.py:59-62 interleaving).  Own implementation: independent of customquad_b200/sparsity.py,
 * so that "bit-exact sparsity" compares two builders.
 * cells == NULL: all cells (dx_cut without subdomain id, demo/poisson.py:180,189); else the listed cells.
 * --------------------------------------------------------------------------------------- */
typedef struct oracle_pattern {
  int64_t num_dofs;
  int64_t* start;   /* [num_dofs + 1] capacity offsets */
  int32_t* len;     /* [num_dofs] columns currently in the row */
  int32_t* cols;    /* [start[num_dofs]] */
} oracle_pattern;

void cq_oracle_pattern_free(oracle_pattern* p) {
  if (!p) return;
  free(p->start); free(p->len); free(p->cols); free(p);
}

oracle_pattern* cq_oracle_pattern_build(const int32_t* dofmap, int64_t num_cells, int n, int64_t num_dofs, const int32_t* cells, int64_t ncells_listed) {
  oracle_pattern* p = (oracle_pattern*)calloc(1, sizeof *p);
  p->num_dofs = num_dofs;
  p->start = (int64_t*)calloc((size_t)num_dofs + 1, sizeof(int64_t));
  p->len = (int32_t*)calloc((size_t)num_dofs > 0 ? (size_t)num_dofs : 1, sizeof(int32_t));
  const int64_t nc = cells ? ncells_listed : num_cells;
  /* preallocation: a row gets room for n columns per cell that touches it (an upper bound) */
  for (int64_t k = 0; k < nc; ++k) {
    const int32_t* d = dofmap + (size_t)(cells ? cells[k] : k) * n;
    for (int i = 0; i < n; ++i) p->start[d[i] + 1] += n;
  }
  for (int64_t r = 0; r < num_dofs; ++r) p->start[r + 1] += p->start[r];
  p->cols = (int32_t*)malloc(sizeof(int32_t) * (size_t)(p->start[num_dofs] > 0 ? p->start[num_dofs] : 1));
  for (int64_t k = 0; k < nc; ++k) {
    const int32_t* d = dofmap + (size_t)(cells ? cells[k] : k) * n;
    for (int i = 0; i < n; ++i) {
      int32_t* row = p->cols + p->start[d[i]];
      int32_t m = p->len[d[i]];
      for (int j = 0; j < n; ++j) {
        const int32_t c = d[j];
        int32_t lo = 0, hi = m;
        while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (row[mid] < c) lo = mid + 1; else hi = mid; }
        if (lo < m && row[lo] == c) continue;
        memmove(row + lo + 1, row + lo, sizeof(int32_t) * (size_t)(m - lo));
        row[lo] = c;
        ++m;
      }
      p->len[d[i]] = m;
    }
  }
  return p;
}

int64_t cq_oracle_pattern_nnz(const oracle_pattern* p, int bs) {
  int64_t s = 0;
  for (int64_t r = 0; r < p->num_dofs; ++r) s += p->len[r];
  return s * bs * bs;
}

/* indptr [num_dofs*bs + 1] int64, indices [nnz] int32 */
void cq_oracle_pattern_export(const oracle_pattern* p, int bs, int64_t* indptr, int32_t* indices) {
  int64_t at = 0;
  indptr[0] = 0;
  for (int64_t r = 0; r < p->num_dofs; ++r)
    for (int a = 0; a < bs; ++a) {
      const int32_t* row = p->cols + p->start[r];
      for (int32_t k = 0; k < p->len[r]; ++k)
        for (int b = 0; b < bs; ++b) indices[at++] = row[k] * bs + b;
      indptr[r * bs + a + 1] = at;
    }
}

int cq_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Debug aid: print a native backtrace on SIGSEGV (set CQ_ORACLE_BACKTRACE=1 in oracle.py). */
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void segv_handler(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
void cq_oracle_install_segv_handler(void) { signal(SIGSEGV, segv_handler); }
