// cq_moment.cuh -- mass / stiffness matrices on axis-aligned quadrilateral cells through the
// monomial moments of the runtime quadrature rule.
//
// On a cell whose Jacobian is constant and diagonal (every cell of a dolfinx create_rectangle
// mesh: J = diag(hx, hy)), the integrands of the reference's bilinear forms are polynomials in
// the reference coordinate with per-cell constant coefficients:
//
//   mass       A_ij = |det J| sum_q w_q psi_i psi_j
//   stiffness  A_ij = sum_k (|det J| / J_kk^2) sum_q w_q d_k psi_i d_k psi_j
//
// and psi_i = l_a0(xi0) l_a1(xi1).  Each product l_a l_c (or l_a' l_c') is a polynomial of degree
// <= 2 deg in one variable, so every entry is a fixed linear combination of the (2 deg + 1)^2
// moments  mu[p][q] = sum_q w_q t0^p t1^q,  t = xi - 1/2  (centred monomials keep the combination
// well conditioned: measured max deviation from the direct evaluation 2.5e-15 of max|A| for Q2).
// Per point this costs 15 (Q1) / 37 (Q2) FP64 instructions instead of ~65 / ~190, which moves the
// Q1 AND the Q2 kernels from the FP64 roofline to the HBM roofline.  The reference computes the
// same sums entry by entry (README.md:133-158 generated kernel); results agree to FP64 rounding.
//
// One thread per qr entry (no cross-lane reduction at all); the points come from the staged
// shared-memory slice, so the per-thread streams cost no coalescing.  After the point loop the
// thread expands its moments into the n x n element tensor and scatters each value as soon as it
// is formed (no n x n register tile): for every unordered x-pair (a0 <= c0) contract mu with the
// x coefficient rows once, then for every ordered y-pair contract with the y rows.
#pragma once
#include <utility>

#include "cq_staged.cuh"

namespace cq {

template <int B, int E, class F>
CQ_DEVINL void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(static_cast<F&&>(f));
  }
}

// 1-D Lagrange basis in the centred variable t = xi - 1/2 and the coefficient rows of pair products
template <int DEG>
struct Poly1D {
  static constexpr int P = 2 * DEG + 1;
  // coefficient of t^u in l_a(t + 1/2)  (a = lattice coordinate 0..DEG)
  CQ_HD static constexpr double L(int a, int u) {
    if (DEG == 1) {
      constexpr double T[2][2] = {{0.5, -1.0}, {0.5, 1.0}};
      return u < 2 ? T[a][u] : 0.0;
    } else {
      // l_0 = 2t^2 - t, l_1 = 1 - 4t^2 (midpoint), l_2 = 2t^2 + t
      constexpr double T[3][3] = {{0.0, -1.0, 2.0}, {1.0, 0.0, -4.0}, {0.0, 1.0, 2.0}};
      return u < 3 ? T[a][u] : 0.0;
    }
  }
  CQ_HD static constexpr double dL(int a, int u) { return u < DEG ? (u + 1) * L(a, u + 1) : 0.0; }
  CQ_HD static constexpr double R00(int a, int c, int p) {
    double s = 0.0;
    for (int u = 0; u <= DEG; ++u)
      if (p - u >= 0 && p - u <= DEG) s += L(a, u) * L(c, p - u);
    return s;
  }
  // d/dt (l_a l_c): derivative of the product polynomial
  CQ_HD static constexpr double RD(int a, int c, int p) { return p + 1 < P ? (p + 1) * R00(a, c, p + 1) : 0.0; }
  // KIND 0: l_a l_c, 1: (l_a l_c)'
  template <int KIND>
  CQ_HD static constexpr double RK(int a, int c, int p) {
    return KIND == 0 ? R00(a, c, p) : RD(a, c, p);
  }
  CQ_HD static constexpr double R11(int a, int c, int p) {
    double s = 0.0;
    for (int u = 0; u < DEG; ++u)
      if (p - u >= 0 && p - u < DEG) s += dL(a, u) * dL(c, p - u);
    return s;
  }
};

// local dof of lattice point (a0, a1)
template <int DEG>
CQ_HD constexpr int quad_dof_of(int a0, int a1) {
  for (int j = 0; j < Lattice<2, DEG>::N; ++j)
    if (Lattice<2, DEG>::t(j, 0) == a0 && Lattice<2, DEG>::t(j, 1) == a1) return j;
  return -1;
}

// Per-entry scatter data held in registers: the scatter positions (one byte each, packed) and one
// pointer per local row (values + CSR row start of component `comp`), computed once per entry, so
// that a value costs: byte extract, one 64-bit multiply-add, one RED.
//   UNIT (bs == 1): slot = rowstart[i] + pos
//   blocked:        slot = rowstart[i] + comp*rowlen[i] + pos*bs + comp   (stride = bs)
template <class T, int CB, bool UNIT>
struct ScatterRegs {
  static constexpr int N = T::N;
  static constexpr int NW = StageMeta<T, CB>::NPOS / 4;
  uint32_t posw[NW];
  double* rowp[N];
  int stride;
  CQ_DEVINL void load(const KParams& P, const StageMeta<T, CB>& M, const int e, const int comp) {
    const int4* src = reinterpret_cast<const int4*>(&M.pos[e][0]);
#pragma unroll
    for (int i = 0; i < NW / 4; ++i) {
      const int4 v = src[i];
      posw[4 * i] = v.x; posw[4 * i + 1] = v.y; posw[4 * i + 2] = v.z; posw[4 * i + 3] = v.w;
    }
    stride = UNIT ? 1 : P.bs;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if constexpr (UNIT) rowp[i] = P.values + M.rowstart[e][i];
      else rowp[i] = P.values + M.rowstart[e][i] + (long long)comp * M.rowlen[e][i] + comp;
    }
  }
  template <int I, int J>
  CQ_DEVINL void add(const double v) const {
    constexpr int k = I * N + J;
    const unsigned pij = (posw[k >> 2] >> (8 * (k & 3))) & 0xffu;
    if constexpr (UNIT) red_add_f64(rowp[I] + pij, v);
    else red_add_f64(rowp[I] + pij * (unsigned)stride, v);
  }
};

// owner (expansion warp) of the unordered x-pair number idx when the expansion is split over NXW warps
template <int DEG, int NXW>
CQ_HD constexpr int pair_owner(int idx) {
  if (NXW == 1) return 0;
  if (DEG == 1) {  // pairs (0,0) (0,1) (1,1): 4 + 8 + 4 values
    constexpr int o2[3] = {1, 0, 1};
    return NXW == 2 ? o2[idx] : idx % NXW;
  }
  // pairs (0,0) (0,1) (0,2) (1,1) (1,2) (2,2): 9 18 18 9 18 9 values
  constexpr int o3[6] = {2, 0, 1, 1, 2, 0};
  constexpr int o2[6] = {0, 0, 1, 1, 0, 1};  // 9+18+18 = 45 | 18+9+9 = 36
  return NXW == 3 ? o3[idx] : (NXW == 2 ? o2[idx] : idx % NXW);
}

// points in flight per lane in the quadrilateral point loop: two on the bench workload (U{16..64} points per cell:
// three cost 5.23 -> 5.32 ms per step, four 5.42), four in the long-rule shape (cq_launch.cuh; Q2 stiffness at
// U{200..300} points per cell 2.71 -> 2.11 ms, Q1 at U{80..120} 0.87 -> 0.66 ms)
#ifndef CQ_MOMENT_UNROLL2
#define CQ_MOMENT_UNROLL2 2
#endif
#ifndef CQ_MOMENT_UNROLL2L
#define CQ_MOMENT_UNROLL2L 4
#endif
// T: FormTraits with DIM == 2 and FORM in {MASS, STIFFNESS}.
// Point loop: G lanes per entry, lane r takes the points r, r+G, ...; the moments are all-reduced over
// the group (every lane ends with the full set).
template <class T, int G, int CB, bool ALLREDUCE = true, int UNR = CQ_MOMENT_UNROLL2>
CQ_DEVINL void moment_accumulate(const StageMeta<T, CB>& M, const int e, const bool active, const int r,
                                 double (&mu)[2 * T::DEG + 1][2 * T::DEG + 1]) {
  constexpr int Pn = 2 * T::DEG + 1;
#pragma unroll
  for (int p = 0; p < Pn; ++p)
#pragma unroll
    for (int q = 0; q < Pn; ++q) mu[p][q] = 0.0;
  const int nq = active ? M.nq[e] : 0;
  const double* pp = M.pts[e];
  const double* wp = M.w[e];
#pragma unroll UNR
  for (int q = r; q < nq; q += G) {
    const double2 xy = *reinterpret_cast<const double2*>(pp + 2 * q);
    const double w = wp[q];
    const double tx = xy.x - 0.5, ty = xy.y - 0.5;
    double X[Pn], WY[Pn];
    X[0] = 1.0; X[1] = tx;
    WY[0] = w; WY[1] = w * ty;
#pragma unroll
    for (int p = 2; p < Pn; ++p) {
      X[p] = X[p - 1] * tx;
      WY[p] = WY[p - 1] * ty;
    }
#pragma unroll
    for (int q2 = 0; q2 < Pn; ++q2) mu[0][q2] += WY[q2];
#pragma unroll
    for (int p = 1; p < Pn; ++p)
#pragma unroll
      for (int q2 = 0; q2 < Pn; ++q2) mu[p][q2] = fma(X[p], WY[q2], mu[p][q2]);
  }
  if constexpr (G > 1 && ALLREDUCE) {
#pragma unroll
    for (int st = G / 2; st >= 1; st >>= 1)
#pragma unroll
      for (int p = 0; p < Pn; ++p)
#pragma unroll
        for (int q = 0; q < Pn; ++q) mu[p][q] += __shfl_xor_sync(0xffffffffu, mu[p][q], st);
  }
}

// Split kernels: the moments of an entry go to the stage header, so instead of an all-reduce the
// G lanes transpose-reduce (M (1 - 1/G) shuffles) and lane r stores its slice of the sums.
// Q1 hexahedron: mu[(p * 3 + q) * 3 + r] = sum w t0^p t1^q t2^r, p, q, r <= 2 (39 FP64 instructions per point)
// Position of moment k in the accumulator handed to transpose_reduce_step<MPAD, G / 2>: after the reduction lane r
// holds in acc[kk] the sum of position r * B + kk, so moment k = kk * G + r sits at (k % G) * B + k / G.  The G lanes
// of an entry then store G CONSECUTIVE moments per instruction (one 8 G-byte run per entry) instead of G runs a
// multiple of 256 B apart (same banks: ncu counted 9.4 wavefronts per store of the round-1 layout, ideal 2).
template <int G, int MPAD>
CQ_HD constexpr int mom_slot(int k) { return (k % G) * (MPAD / G) + k / G; }

// UNR: points in flight per lane.  Two for the short / medium shapes; the long-rule shape (ten moment warps,
// cq_launch.cuh) takes four: most of its entries overflow the stage and are read from global memory, where the loop is
// bound by the loads in flight (measured with ten moment warps, 96^3 cells: U{100..120} points per cell 1.59 -> 1.26 ms,
// U{200..300} 3.97 -> 3.36 ms; U{8..27} on the medium shape 0.300 -> 0.309 ms, so not there)
#ifndef CQ_MOMENT_UNROLL3
#define CQ_MOMENT_UNROLL3 2
#endif
#ifndef CQ_MOMENT_UNROLL3L
#define CQ_MOMENT_UNROLL3L 4
#endif
template <class T, int G, int CB, int MPAD, int UNR = CQ_MOMENT_UNROLL3>
CQ_DEVINL void moment_accumulate3(const StageMeta<T, CB>& M, const int e, const bool active, const int r, double (&acc)[MPAD]) {
  constexpr int Pn = 3;
  auto S = [](int k) constexpr { return mom_slot<G, MPAD>(k); };
#pragma unroll
  for (int k = 0; k < MPAD; ++k) acc[k] = 0.0;
  const int nq = active ? M.nq[e] : 0;
  const double* pp = M.pts[e];
  const double* wp = M.w[e];
  // Every entry starts its point loop at a different point (the sum is over all of them anyway).  With the fixed
  // 8-point rule of the uncut cells the entries lie 192 bytes apart in the stage: un-rotated, the sixteen entries of
  // a load instruction hit two bank groups (ncu: 8 wavefronts per cell for the point loads, ideal 2,
  // profiles/r2/ncu_src_hexq1_stiffness_8pt_before.txt); rotated by e / 2 they tile all 32 banks twice.
  int rot = 0;
  if (nq > 1) rot = (e >> 1) & ((1 << (31 - __clz(nq))) - 1);
#pragma unroll UNR
  for (int q0 = r; q0 < nq; q0 += G) {
    int q = q0 + rot;
    if (q >= nq) q -= nq;
    const double t0 = pp[3 * q] - 0.5, t1 = pp[3 * q + 1] - 0.5, t2 = pp[3 * q + 2] - 0.5;
    const double w = wp[q];
    double Z[Pn], YZ[Pn][Pn];
    Z[0] = w; Z[1] = w * t2; Z[2] = Z[1] * t2;
    const double t11 = t1 * t1, t00 = t0 * t0;
#pragma unroll
    for (int c = 0; c < Pn; ++c) {
      YZ[0][c] = Z[c];
      YZ[1][c] = t1 * Z[c];
      YZ[2][c] = t11 * Z[c];
    }
#pragma unroll
    for (int b = 0; b < Pn; ++b)
#pragma unroll
      for (int c = 0; c < Pn; ++c) {
        acc[S((0 * Pn + b) * Pn + c)] += YZ[b][c];
        acc[S((1 * Pn + b) * Pn + c)] = fma(t0, YZ[b][c], acc[S((1 * Pn + b) * Pn + c)]);
        acc[S((2 * Pn + b) * Pn + c)] = fma(t00, YZ[b][c], acc[S((2 * Pn + b) * Pn + c)]);
      }
  }
}

#ifndef CQ_NITSCHE_UNROLL  // points in flight per lane in the Nitsche moment-set loop (hexahedra 96^3, U{4..16} / U{8..27} surface
#define CQ_NITSCHE_UNROLL 2  // points per cell: 1: 0.472 / 0.569 ms, 2: 0.483 / 0.586, 4: 0.516 / 0.624 -- but on the real surface rules
#endif                      // of config 3 one point in flight costs 0.32 -> 0.36 ms, so two)
constexpr int kNitscheUnroll = CQ_NITSCHE_UNROLL;
// Nitsche: lane r of the entry's group of four owns moment set r (weights w, w n_0, w n_1 (, w n_2)) and runs
// over ALL points of the entry: no cross-lane reduction, and the point itself is a shared-memory broadcast.
template <class T, int CB>
CQ_DEVINL void moment_accumulate_sets(StageMeta<T, CB>& M, const int e, const bool active, const bool compute, const int r) {
  using Meta = StageMeta<T, CB>;
  constexpr int DIM = T::DIM, Pn = Meta::PN, NM1 = Meta::NM1, NSET = Meta::NSET;
  double acc[NM1];
#pragma unroll
  for (int k = 0; k < NM1; ++k) acc[k] = 0.0;
  const bool mine = active && r < NSET;
  const int nq = (mine && compute) ? M.nq[e] : 0;
  const double* pp = M.pts[e];
  const double* wp = M.w[e];
  const double* np_ = M.nrm[e];
#pragma unroll kNitscheUnroll
  for (int q = 0; q < nq; ++q) {
    double w = wp[q];
    if (r > 0) w *= np_[DIM * q + (r - 1)];
    if constexpr (DIM == 2) {
      const double2 xy = *reinterpret_cast<const double2*>(pp + 2 * q);
      const double tx = xy.x - 0.5, ty = xy.y - 0.5;
      double X[Pn], WY[Pn];
      X[0] = 1.0; X[1] = tx;
      WY[0] = w; WY[1] = w * ty;
#pragma unroll
      for (int p = 2; p < Pn; ++p) {
        X[p] = X[p - 1] * tx;
        WY[p] = WY[p - 1] * ty;
      }
#pragma unroll
      for (int b = 0; b < Pn; ++b) acc[b] += WY[b];
#pragma unroll
      for (int a = 1; a < Pn; ++a)
#pragma unroll
        for (int b = 0; b < Pn; ++b) acc[a * Pn + b] = fma(X[a], WY[b], acc[a * Pn + b]);
    } else {
      static_assert(DIM == 2 || Pn == 3, "hexahedra: degree 1");
      const double t0 = pp[3 * q] - 0.5, t1 = pp[3 * q + 1] - 0.5, t2 = pp[3 * q + 2] - 0.5;
      double Z[3], YZ[3][3];
      Z[0] = w; Z[1] = w * t2; Z[2] = Z[1] * t2;
      const double t11 = t1 * t1, t00 = t0 * t0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        YZ[0][c] = Z[c];
        YZ[1][c] = t1 * Z[c];
        YZ[2][c] = t11 * Z[c];
      }
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          acc[(0 * 3 + b) * 3 + c] += YZ[b][c];
          acc[(1 * 3 + b) * 3 + c] = fma(t0, YZ[b][c], acc[(1 * 3 + b) * 3 + c]);
          acc[(2 * 3 + b) * 3 + c] = fma(t00, YZ[b][c], acc[(2 * 3 + b) * 3 + c]);
        }
    }
  }
  if (mine) {
#pragma unroll
    for (int k = 0; k < NM1; ++k) M.mu[e][r * NM1 + k] = acc[k];
  }
}

template <class T, int G, int CB, bool LONG = false>
CQ_DEVINL void moment_accumulate_store(StageMeta<T, CB>& M, const int e, const bool active, const bool compute, const int r) {
  if constexpr (StageMeta<T, CB>::NIT) {
    static_assert(!StageMeta<T, CB>::NIT || G == 4, "Nitsche moment sets: four lanes per entry");
    moment_accumulate_sets<T, CB>(M, e, active, compute, r);
    return;
  } else {
  constexpr int Pn = 2 * T::DEG + 1, NM = StageMeta<T, CB>::NMOM;
  constexpr int MPAD = round_up(NM, G), B = MPAD / G;
  double acc[MPAD];
  if constexpr (T::DIM == 3) {
    moment_accumulate3<T, G, CB, MPAD, (LONG ? CQ_MOMENT_UNROLL3L : CQ_MOMENT_UNROLL3)>(M, e, active && compute, r, acc);
  } else {
    double mu[Pn][Pn];
    moment_accumulate<T, G, CB, false, (LONG ? CQ_MOMENT_UNROLL2L : CQ_MOMENT_UNROLL2)>(M, e, active && compute, r, mu);
#pragma unroll
    for (int k = 0; k < MPAD; ++k) acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NM; ++k) acc[mom_slot<G, MPAD>(k)] = mu[k / Pn][k % Pn];
  }
  transpose_reduce_step<MPAD, G / 2>(acc, r);
  if (active) {
#pragma unroll
    for (int kk = 0; kk < B; ++kk)
      if (kk * G + r < NM) M.mu[e][kk * G + r] = acc[kk];
  }
  }
}

// Expansion moments -> element tensor values, each scattered as soon as it is formed.
//   XW / NXW: this caller expands only the x-pairs owned by expansion warp XW (NXW == 1: all of them)
//   G, r:     with NXW == 1 and G lanes holding the same moments, lane r scatters the values k % G == r
template <class T, int G, int CB, int XW, int NXW, bool UNIT>
CQ_DEVINL void moment_expand_comp(const KParams& P, const StageMeta<T, CB>& M, const int e, const int r, const int comp,
                                  const double (&mu)[2 * T::DEG + 1][2 * T::DEG + 1]) {
  constexpr int DEG = T::DEG;
  using Q = Poly1D<DEG>;
  constexpr int Pn = Q::P;
  constexpr bool STIFF = T::FORM_ID == CQ_FORM_STIFFNESS;
  ScatterRegs<T, CB, UNIT> S;
  S.load(P, M, e, comp);
  // per-cell constants: J = diag(J00, J11) (axis-aligned cell; checked for the whole mesh at setup)
  const double k0 = STIFF ? M.kc[1][e] : M.kc[0][e];  // stiffness: |det|/J00^2 ; mass: |det|  (formed by the producer)
  const double k1 = STIFF ? M.kc[2][e] : 0.0;         // stiffness: |det|/J11^2

  static_for<0, DEG + 1>([&](auto A0) {
    static_for<decltype(A0)::value, DEG + 1>([&](auto C0) {
      constexpr int a0 = decltype(A0)::value, c0 = decltype(C0)::value;
      constexpr int pair_idx = a0 * (DEG + 1) - a0 * (a0 - 1) / 2 + (c0 - a0);
      if constexpr (pair_owner<DEG, NXW>(pair_idx) == XW) {
        // contract the x index:  U0[q] = sum_p R00x[p] mu[p][q],  U1[q] = sum_p R11x[p] mu[p][q]
        double U0[Pn], U1[Pn];
#pragma unroll
        for (int q = 0; q < Pn; ++q) U0[q] = U1[q] = 0.0;
        static_for<0, Pn>([&](auto PP) {
          constexpr int p = decltype(PP)::value;
          constexpr double r00 = Q::R00(a0, c0, p), r11 = Q::R11(a0, c0, p);
          if constexpr (r00 != 0.0) {
#pragma unroll
            for (int q = 0; q < Pn; ++q) U0[q] = fma(r00, mu[p][q], U0[q]);
          }
          if constexpr (STIFF && r11 != 0.0) {
#pragma unroll
            for (int q = 0; q < Pn; ++q) U1[q] = fma(r11, mu[p][q], U1[q]);
          }
        });
        static_for<0, DEG + 1>([&](auto A1) {
          static_for<0, DEG + 1>([&](auto C1) {
            constexpr int a1 = decltype(A1)::value, c1 = decltype(C1)::value;
            double m0 = 0.0, m1 = 0.0;  // y contractions
            static_for<0, Pn>([&](auto QQ) {
              constexpr int q = decltype(QQ)::value;
              constexpr double r00 = Q::R00(a1, c1, q), r11 = Q::R11(a1, c1, q);
              if constexpr (STIFF) {
                if constexpr (r00 != 0.0) m0 = fma(r00, U1[q], m0);  // d/dx d/dx: R11 in x, R00 in y
                if constexpr (r11 != 0.0) m1 = fma(r11, U0[q], m1);  // d/dy d/dy: R00 in x, R11 in y
              } else {
                if constexpr (r00 != 0.0) m0 = fma(r00, U0[q], m0);
              }
            });
            const double v = STIFF ? fma(k0, m0, k1 * m1) : k0 * m0;
            constexpr int i = quad_dof_of<DEG>(a0, a1), j = quad_dof_of<DEG>(c0, c1);
            if (G == 1 || NXW > 1 || (i * T::N + j) % G == r) S.template add<i, j>(v);
            if constexpr (a0 != c0) {  // x-swapped entry has the same value (the x rows are symmetric)
              constexpr int i2 = quad_dof_of<DEG>(c0, a1), j2 = quad_dof_of<DEG>(a0, c1);
              if (G == 1 || NXW > 1 || (i2 * T::N + j2) % G == r) S.template add<i2, j2>(v);
            }
          });
        });
      }
    });
  });
}

// ---- row-owned expansion + transposed scatter --------------------------------------------------
// Measured on the bench workload (profiles/r1/scatter_study.md): a RED (or a plain store) whose 32
// lanes go to 32 different rows costs one L1 tag cycle per lane, i.e. 81 * 32 cycles per stage for
// Q2 -- 1.2 ms of the 2.5 ms kernel -- whatever the L2 does.  With lane = entry (above) that is the
// access pattern.  Here an expansion warp owns the rows whose x lattice index is A0: it forms their
// (DEG+1) * N values per entry with lane = entry, parks them in a shared-memory tile (the stage's
// point buffer, dead once the moments are done), and scatters the tile with lane = position inside
// the tile: consecutive lanes hold consecutive columns (c1 fastest) of one row, which are adjacent
// CSR slots whenever the dofs along the points' axis 1 are numbered consecutively (dolfinx
// create_rectangle is y-fastest) -> one sector per run of DEG+1 values (ncu: 2.5x fewer RED sectors,
// 5x fewer L1 wavefronts).  The scatter loop is unrolled over the entries so that every address is
// base + constant.
#ifndef CQ_SCATTER_UNROLL
#define CQ_SCATTER_UNROLL 4
#endif
template <int DIM, int DEG>
struct RowTile {
  static constexpr int D1 = DEG + 1, N = DIM == 2 ? D1 * D1 : D1 * D1 * D1;
  static constexpr int NROLE = D1;          // expansion roles: the x lattice index A0 of the owned rows
  static constexpr int NVW = (N / D1) * N;  // values per entry owned by one expansion warp (Q1 quad 8, Q2 quad 27, Q1 hex 32)
  static constexpr int STRIDE = NVW | 1;    // odd stride: conflict-free one-thread-per-entry writes
  static constexpr int EPI = 32 / NVW;      // entries per scatter instruction (Q1 quad: 4, Q2 quad / Q1 hex: 1)
  static constexpr size_t bytes(int cb) { return (size_t)cb * STRIDE * 8; }
  // quadrilaterals: local dof of lattice point (x, y), 4 bits each, index x * D1 + y
  CQ_HD static constexpr uint64_t dof_table() {
    uint64_t t = 0;
    if (DIM == 2)
      for (int x = 0; x < D1; ++x)
        for (int y = 0; y < D1; ++y) t |= (uint64_t)quad_dof_of<DEG>(x, y) << (4 * (x * D1 + y));
    return t;
  }
};

// lane = entry e: values (a1, c0, c1) of the rows with x index A0 -> tile[(a1 * D1 + c0) * D1 + c1]
// Q1 hexahedron (local dof j = a0 + 2 a1 + 4 a2, Lattice<3,1>): rows i = A0 + 2 (a1 + 2 a2), all 8 columns
// j = c0 + 2 c1 + 4 c2 -> tile[(a1 + 2 a2) * 8 + j].  The three indices are contracted one after the other
// (x, y, z); stiffness carries the three derivative placements (R11 in x | y | z) through the chain.
template <class T, int CB, int A0>
CQ_DEVINL void moment_expand_rows3(const StageMeta<T, CB>& M, const int e, double* __restrict__ tile) {
  static_assert(T::DIM == 3 && T::DEG == 1, "Q1 hexahedron");
  using Q = Poly1D<1>;
  constexpr int Pn = 3;
  constexpr bool STIFF = T::FORM_ID == CQ_FORM_STIFFNESS;
  double mu[Pn][Pn][Pn];
#pragma unroll
  for (int p = 0; p < Pn; ++p)
#pragma unroll
    for (int q = 0; q < Pn; ++q)
#pragma unroll
      for (int r = 0; r < Pn; ++r) mu[p][q][r] = M.mu[e][(p * Pn + q) * Pn + r];
  const double kM = M.kc[0][e];
  const double kx = STIFF ? M.kc[1][e] : 0.0, ky = STIFF ? M.kc[2][e] : 0.0, kz = STIFF ? M.kc[3][e] : 0.0;
  static_for<0, 2>([&](auto C0) {
    constexpr int c0 = decltype(C0)::value;
    // x: U0 = R00x . mu, U1 = R11x . mu
    double U0[Pn][Pn], U1[Pn][Pn];
    bool i0 = false, i1 = false;
    static_for<0, Pn>([&](auto PP) {
      constexpr int p = decltype(PP)::value;
      constexpr double r00 = Q::R00(A0, c0, p), r11 = Q::R11(A0, c0, p);
      if constexpr (r00 != 0.0) {
#pragma unroll
        for (int q = 0; q < Pn; ++q)
#pragma unroll
          for (int r = 0; r < Pn; ++r) U0[q][r] = i0 ? fma(r00, mu[p][q][r], U0[q][r]) : r00 * mu[p][q][r];
        i0 = true;
      }
      if constexpr (STIFF && r11 != 0.0) {
#pragma unroll
        for (int q = 0; q < Pn; ++q)
#pragma unroll
          for (int r = 0; r < Pn; ++r) U1[q][r] = i1 ? fma(r11, mu[p][q][r], U1[q][r]) : r11 * mu[p][q][r];
        i1 = true;
      }
    });
    static_for<0, 2>([&](auto A1) {
      static_for<0, 2>([&](auto C1) {
        constexpr int a1 = decltype(A1)::value, c1 = decltype(C1)::value;
        // y: V00 = R00y . U0 (mass, d/dz d/dz), V01 = R11y . U0 (d/dy d/dy), V10 = R00y . U1 (d/dx d/dx)
        double V00[Pn], V01[Pn], V10[Pn];
        bool j00 = false, j01 = false, j10 = false;
        static_for<0, Pn>([&](auto QQ) {
          constexpr int q = decltype(QQ)::value;
          constexpr double r00 = Q::R00(a1, c1, q), r11 = Q::R11(a1, c1, q);
          if constexpr (r00 != 0.0) {
#pragma unroll
            for (int r = 0; r < Pn; ++r) V00[r] = j00 ? fma(r00, U0[q][r], V00[r]) : r00 * U0[q][r];
            j00 = true;
            if constexpr (STIFF) {
#pragma unroll
              for (int r = 0; r < Pn; ++r) V10[r] = j10 ? fma(r00, U1[q][r], V10[r]) : r00 * U1[q][r];
              j10 = true;
            }
          }
          if constexpr (STIFF && r11 != 0.0) {
#pragma unroll
            for (int r = 0; r < Pn; ++r) V01[r] = j01 ? fma(r11, U0[q][r], V01[r]) : r11 * U0[q][r];
            j01 = true;
          }
        });
        static_for<0, 2>([&](auto A2) {
          static_for<0, 2>([&](auto C2) {
            constexpr int a2 = decltype(A2)::value, c2 = decltype(C2)::value;
            double m = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
            bool km = false, ks = false, kzz = false;
            static_for<0, Pn>([&](auto RR) {
              constexpr int r = decltype(RR)::value;
              constexpr double r00 = Q::R00(a2, c2, r), r11 = Q::R11(a2, c2, r);
              if constexpr (!STIFF) {
                if constexpr (r00 != 0.0) { m = km ? fma(r00, V00[r], m) : r00 * V00[r]; km = true; }
              } else {
                if constexpr (r00 != 0.0) {
                  sx = ks ? fma(r00, V10[r], sx) : r00 * V10[r];
                  sy = ks ? fma(r00, V01[r], sy) : r00 * V01[r];
                  ks = true;
                }
                if constexpr (r11 != 0.0) { sz = kzz ? fma(r11, V00[r], sz) : r11 * V00[r]; kzz = true; }
              }
            });
            tile[(a1 + 2 * a2) * 8 + (c0 + 2 * c1 + 4 * c2)] = STIFF ? fma(kx, sx, fma(ky, sy, kz * sz)) : kM * m;
          });
        });
      });
    });
  });
}

// Symmetric Nitsche form on an axis-aligned cell (J = diag(J_dd)):
//   A_ij = sum_q w |det J| [ -(n.G_j) psi_i - psi_j (n.G_i) + betaN/h psi_i psi_j ],  n.G_i = sum_d n_d d_d psi_i / J_dd
//        = gamma W[psi_i psi_j] - sum_d kappa_d N_d[ d_d (psi_i psi_j) ],
// W[f] = sum w f, N_d[f] = sum w n_d f, gamma = |det J| betaN / h_K, kappa_d = |det J| / J_dd (kc, producer).
// Each term is a mass-like contraction of one moment set, with the derivative rows RD in dimension d.
template <class T, int CB, int A0>
CQ_DEVINL void moment_expand_rows_nitsche(const StageMeta<T, CB>& M, const int e, double* __restrict__ tile) {
  using Meta = StageMeta<T, CB>;
  constexpr int DIM = T::DIM, DEG = T::DEG, D1 = DEG + 1, Pn = Meta::PN, NM1 = Meta::NM1;
  using Q = Poly1D<DEG>;
  const double gam = M.kc[1 + DIM][e];
  double kap[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) kap[d] = M.kc[1 + d][e];
  if constexpr (DIM == 2) {
    static_for<0, D1>([&](auto C0) {
      constexpr int c0 = decltype(C0)::value;
      // x contraction of the three sets: (W, R00), (N0, RD), (N1, R00)
      double U[3][Pn];
      static_for<0, 3>([&](auto SS) {
        constexpr int set = decltype(SS)::value;
        bool init = false;
        static_for<0, Pn>([&](auto PP) {
          constexpr int p = decltype(PP)::value;
          constexpr double rc = Q::template RK<(set == 1) ? 1 : 0>(A0, c0, p);
          if constexpr (rc != 0.0) {
#pragma unroll
            for (int q = 0; q < Pn; ++q) {
              const double m = M.mu[e][set * NM1 + p * Pn + q];
              U[set][q] = init ? fma(rc, m, U[set][q]) : rc * m;
            }
            init = true;
          }
        });
        if (!init) {
#pragma unroll
          for (int q = 0; q < Pn; ++q) U[set][q] = 0.0;
        }
      });
      static_for<0, D1>([&](auto A1) {
        static_for<0, D1>([&](auto C1) {
          constexpr int a1 = decltype(A1)::value, c1 = decltype(C1)::value;
          double mW = 0.0, m0 = 0.0, m1 = 0.0;
          static_for<0, Pn>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            constexpr double r00 = Q::R00(a1, c1, q), rd = Q::RD(a1, c1, q);
            if constexpr (r00 != 0.0) {
              mW = fma(r00, U[0][q], mW);
              m0 = fma(r00, U[1][q], m0);
            }
            if constexpr (rd != 0.0) m1 = fma(rd, U[2][q], m1);
          });
          tile[(a1 * D1 + c0) * D1 + c1] = fma(gam, mW, -fma(kap[0], m0, kap[1] * m1));
        });
      });
    });
  } else {
    static_assert(DIM == 2 || DEG == 1, "hexahedra: degree 1");
    static_for<0, 2>([&](auto C0) {
      constexpr int c0 = decltype(C0)::value;
      // x: (W, R00), (N0, RD), (N1, R00), (N2, R00)
      double U[4][3][3];
      static_for<0, 4>([&](auto SS) {
        constexpr int set = decltype(SS)::value;
        bool init = false;
        static_for<0, 3>([&](auto PP) {
          constexpr int p = decltype(PP)::value;
          constexpr double rc = Q::template RK<(set == 1) ? 1 : 0>(A0, c0, p);
          if constexpr (rc != 0.0) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const double m = M.mu[e][set * NM1 + (p * 3 + q) * 3 + r];
                U[set][q][r] = init ? fma(rc, m, U[set][q][r]) : rc * m;
              }
            init = true;
          }
        });
        if (!init) {
#pragma unroll
          for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int r = 0; r < 3; ++r) U[set][q][r] = 0.0;
        }
      });
      static_for<0, 2>([&](auto A1) {
        static_for<0, 2>([&](auto C1) {
          constexpr int a1 = decltype(A1)::value, c1 = decltype(C1)::value;
          // y: RD for the N1 set, R00 for the others
          double V[4][3];
#pragma unroll
          for (int st = 0; st < 4; ++st)
#pragma unroll
            for (int r = 0; r < 3; ++r) V[st][r] = 0.0;
          static_for<0, 3>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            constexpr double r00 = Q::R00(a1, c1, q), rd = Q::RD(a1, c1, q);
            if constexpr (r00 != 0.0) {
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                V[0][r] = fma(r00, U[0][q][r], V[0][r]);
                V[1][r] = fma(r00, U[1][q][r], V[1][r]);
                V[3][r] = fma(r00, U[3][q][r], V[3][r]);
              }
            }
            if constexpr (rd != 0.0) {
#pragma unroll
              for (int r = 0; r < 3; ++r) V[2][r] = fma(rd, U[2][q][r], V[2][r]);
            }
          });
          static_for<0, 2>([&](auto A2) {
            static_for<0, 2>([&](auto C2) {
              constexpr int a2 = decltype(A2)::value, c2 = decltype(C2)::value;
              double mW = 0.0, m0 = 0.0, m1 = 0.0, m2 = 0.0;
              static_for<0, 3>([&](auto RR) {
                constexpr int r = decltype(RR)::value;
                constexpr double r00 = Q::R00(a2, c2, r), rd = Q::RD(a2, c2, r);
                if constexpr (r00 != 0.0) {
                  mW = fma(r00, V[0][r], mW);
                  m0 = fma(r00, V[1][r], m0);
                  m1 = fma(r00, V[2][r], m1);
                }
                if constexpr (rd != 0.0) m2 = fma(rd, V[3][r], m2);
              });
              tile[(a1 + 2 * a2) * 8 + (c0 + 2 * c1 + 4 * c2)] = fma(gam, mW, -fma(kap[0], m0, fma(kap[1], m1, kap[2] * m2)));
            });
          });
        });
      });
    });
  }
}

template <class T, int CB, int A0>
CQ_DEVINL void moment_expand_rows(const StageMeta<T, CB>& M, const int e, double* __restrict__ tile) {
  if constexpr (StageMeta<T, CB>::NIT) {
    moment_expand_rows_nitsche<T, CB, A0>(M, e, tile);
    return;
  } else if constexpr (T::DIM == 3) {
    moment_expand_rows3<T, CB, A0>(M, e, tile);
    return;
  } else {
  constexpr int DEG = T::DEG, D1 = DEG + 1;
  using Q = Poly1D<DEG>;
  constexpr int Pn = Q::P;
  constexpr bool STIFF = T::FORM_ID == CQ_FORM_STIFFNESS;
  double mu[Pn][Pn];
#pragma unroll
  for (int p = 0; p < Pn; ++p)
#pragma unroll
    for (int q = 0; q < Pn; ++q) mu[p][q] = M.mu[e][p * Pn + q];
  const double k0 = STIFF ? M.kc[1][e] : M.kc[0][e];  // stiffness: |det|/J00^2 ; mass: |det|
  const double k1 = STIFF ? M.kc[2][e] : 0.0;         // stiffness: |det|/J11^2
  static_for<0, D1>([&](auto C0) {
    constexpr int c0 = decltype(C0)::value;
    // contract the x index; the first non-zero term initialises (no 0 + x instructions)
    double U0[Pn], U1[Pn];
    bool i0 = false, i1 = false;
    static_for<0, Pn>([&](auto PP) {
      constexpr int p = decltype(PP)::value;
      constexpr double r00 = Q::R00(A0, c0, p), r11 = Q::R11(A0, c0, p);
      if constexpr (r00 != 0.0) {
#pragma unroll
        for (int q = 0; q < Pn; ++q) U0[q] = i0 ? fma(r00, mu[p][q], U0[q]) : r00 * mu[p][q];
        i0 = true;
      }
      if constexpr (STIFF && r11 != 0.0) {
#pragma unroll
        for (int q = 0; q < Pn; ++q) U1[q] = i1 ? fma(r11, mu[p][q], U1[q]) : r11 * mu[p][q];
        i1 = true;
      }
    });
    static_for<0, D1>([&](auto A1) {
      static_for<0, D1>([&](auto C1) {
        constexpr int a1 = decltype(A1)::value, c1 = decltype(C1)::value;
        double m0 = 0.0, m1 = 0.0;
        bool j0 = false, j1 = false;
        static_for<0, Pn>([&](auto QQ) {
          constexpr int q = decltype(QQ)::value;
          constexpr double r00 = Q::R00(a1, c1, q), r11 = Q::R11(a1, c1, q);
          if constexpr (STIFF) {
            if constexpr (r00 != 0.0) { m0 = j0 ? fma(r00, U1[q], m0) : r00 * U1[q]; j0 = true; }  // d/dx d/dx
            if constexpr (r11 != 0.0) { m1 = j1 ? fma(r11, U0[q], m1) : r11 * U0[q]; j1 = true; }  // d/dy d/dy
          } else {
            if constexpr (r00 != 0.0) { m0 = j0 ? fma(r00, U0[q], m0) : r00 * U0[q]; j0 = true; }
          }
        });
        tile[(a1 * D1 + c0) * D1 + c1] = STIFF ? fma(k0, m0, k1 * m1) : k0 * m0;
      });
    });
  });
  }
}

// All 32 lanes.  tile: nent entries of RowTile::NVW values at stride RowTile::STRIDE.  Lane ->
// (entry within the instruction, value slot) is loop invariant; lanes beyond EPI * NVW idle.
template <class T, int CB, int A0>
CQ_DEVINL void scatter_rows_transposed(const KParams& P, const StageMeta<T, CB>& M, const double* __restrict__ tile, const int nent,
                                       const int lane) {
  using RT = RowTile<T::DIM, T::DEG>;
  constexpr int D1 = RT::D1, N = RT::N, NVW = RT::NVW, EPI = RT::EPI;
  constexpr uint64_t TBL = RT::dof_table();
  const int sub = lane / NVW, s = lane - sub * NVW;
  if (sub >= EPI) return;
  int i, j;
  if constexpr (T::DIM == 2) {
    const int a1 = s / N, cc = s - a1 * N;  // cc = c0 * D1 + c1
    i = (int)((TBL >> (4 * (A0 * D1 + a1))) & 15u);
    j = (int)((TBL >> (4 * cc)) & 15u);
  } else {  // Q1 hexahedron: s = (a1 + 2 a2) * 8 + j
    i = A0 + 2 * (s >> 3);
    j = s & 7;
  }
  const double* tl = tile + sub * RT::STRIDE + s;
  const uint8_t* pp = &M.pos[sub][i * N + j];
  constexpr int NRP = StageMeta<T, CB>::NRP;
  const long long* rs = &M.rowstart[sub][i];
  double* const vals = P.values;
  if (P.bs == 1) {
    if (nent == CB) {
      constexpr int UNROLL = CQ_SCATTER_UNROLL;
#pragma unroll UNROLL
      for (int e0 = 0; e0 < CB; e0 += EPI)
        red_add_f64(vals + rs[e0 * NRP] + pp[e0 * StageMeta<T, CB>::NPOS], tl[e0 * RT::STRIDE]);
    } else {
#pragma unroll 1
      for (int e0 = 0; e0 + sub < nent; e0 += EPI)
        red_add_f64(vals + rs[e0 * NRP] + pp[e0 * StageMeta<T, CB>::NPOS], tl[e0 * RT::STRIDE]);
    }
  } else {
    const int bs = P.bs;
    const int* rl = &M.rowlen[sub][i];
#pragma unroll 1
    for (int e0 = 0; e0 + sub < nent; e0 += EPI) {
      const double v = tl[e0 * RT::STRIDE];
      double* row = vals + rs[e0 * NRP] + (unsigned)pp[e0 * StageMeta<T, CB>::NPOS] * bs;
      const int len = rl[e0 * NRP];
      for (int a = 0; a < bs; ++a) red_add_f64(row + (long long)a * len + a, v);
    }
  }
}

// bs == 1: unit-stride scatter; blocked spaces: the scalar block is scattered once per component
template <class T, int G, int CB, int XW, int NXW>
CQ_DEVINL void moment_expand(const KParams& P, const StageMeta<T, CB>& M, const int e, const int r,
                             const double (&mu)[2 * T::DEG + 1][2 * T::DEG + 1]) {
  if (P.bs == 1) {
    moment_expand_comp<T, G, CB, XW, NXW, true>(P, M, e, r, 0, mu);
  } else {
#pragma unroll 1
    for (int comp = 0; comp < P.bs; ++comp) moment_expand_comp<T, G, CB, XW, NXW, false>(P, M, e, r, comp, mu);
  }
}

// MODE 1 of the staged kernel: the same warp does the point loop and the expansion
template <class T, int G, int CB>
CQ_DEVINL void moment_process_entry(const KParams& P, const StageMeta<T, CB>& M, const int e, const bool active, const int r) {
  double mu[2 * T::DEG + 1][2 * T::DEG + 1];
  moment_accumulate<T, G, CB>(M, e, active, r, mu);
  if (!active || CQ_DBG(P) == 2) return;
  moment_expand<T, G, CB, 0, 1>(P, M, e, r, mu);
}

}  // namespace cq
