// cq_dmma.cuh -- Q2 hexahedron (27 dofs) matrix forms on the FP64 tensor-core path.
//
// The scalar element block is dense 27x27 and the contraction over the cell's points is a real
// GEMM:  S = B^T diag(w |det J|) B  with B [(q,k), j] the pulled-back gradient table
// (SURVEY.md 2.1, K2).  tcgen05 has no f64 kind, so this uses warp-level
// mma.sync.aligned.m8n8k4.f64 (DMMA.8x8x4 in SASS).  One warp per qr entry; each k-step covers
// 4 quadrature points.  Lane (g = lane>>2, t = lane&3) evaluates dofs {g, 8+g, 16+g, 24+g} at
// point 4*step + t, which is exactly the A- and B-fragment layout of m8n8k4 -- no shared memory,
// no shuffles.  The 27x27 block is padded to 32x32 = 4x4 tiles; only the 10 upper-triangular
// tiles are accumulated (20 FP64 registers per lane).
#pragma once
#include "cq_kernels.cuh"
#include "cq_staged.cuh"

namespace cq {

// packed lattice coordinates of Lattice<3,2>: t0 | t1 << 2 | t2 << 4, 27 dofs padded to 32
static __device__ __constant__ uint8_t c_q2hex_lat[32] = {
    0 | 0 << 2 | 0 << 4, 2 | 0 << 2 | 0 << 4, 0 | 2 << 2 | 0 << 4, 2 | 2 << 2 | 0 << 4,
    0 | 0 << 2 | 2 << 4, 2 | 0 << 2 | 2 << 4, 0 | 2 << 2 | 2 << 4, 2 | 2 << 2 | 2 << 4,
    1 | 0 << 2 | 0 << 4, 0 | 1 << 2 | 0 << 4, 0 | 0 << 2 | 1 << 4, 2 | 1 << 2 | 0 << 4,
    2 | 0 << 2 | 1 << 4, 1 | 2 << 2 | 0 << 4, 0 | 2 << 2 | 1 << 4, 2 | 2 << 2 | 1 << 4,
    1 | 0 << 2 | 2 << 4, 0 | 1 << 2 | 2 << 4, 2 | 1 << 2 | 2 << 4, 1 | 2 << 2 | 2 << 4,
    1 | 1 << 2 | 0 << 4, 1 | 0 << 2 | 1 << 4, 0 | 1 << 2 | 1 << 4, 2 | 1 << 2 | 1 << 4,
    1 | 2 << 2 | 1 << 4, 1 | 1 << 2 | 2 << 4, 1 | 1 << 2 | 1 << 4, 0, 0, 0, 0, 0};

CQ_DEVINL void dmma_m8n8k4(double (&d)[2], const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

CQ_DEVINL double sel3(const double (&v)[3], const int t) { return t == 0 ? v[0] : (t == 1 ? v[1] : v[2]); }


// point loop of one cell: accumulates the 10 upper-triangular 8x8 tiles D of the scalar 27x27 block.
// pts / w / nrm: the cell's points (global or shared memory), lat/valid: this lane's four dofs.
template <int FORM>
CQ_DEVINL void dmma_cell_tiles(double (&D)[10][2], const CellGeom<3>& geo, const double gamma, const double* pts, const double* w_,
                               const double* nrm_, const int nq, const int t, const int (&lat)[4][3], const bool (&valid)[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) D[i][0] = D[i][1] = 0.0;
  for (int qb = 0; qb < nq; qb += 4) {
    const int q = qb + t;
    const bool live = q < nq;
    double xi[3] = {0.5, 0.5, 0.5}, nrm[3] = {0.0, 0.0, 0.0}, w = 0.0;
    if (live) {
#pragma unroll
      for (int d = 0; d < 3; ++d) xi[d] = pts[3 * q + d];
      w = w_[q];
      if constexpr (FORM == CQ_FORM_NITSCHE)
#pragma unroll
        for (int d = 0; d < 3; ++d) nrm[d] = nrm_[3 * q + d];
    }
    const Basis1D<2> b0(xi[0]), b1(xi[1]), b2(xi[2]);
    double adj[3][3], det;
    if (geo.affine) {  // constant Jacobian: per-cell values
      det = geo.det0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) adj[a][b] = geo.adj0[a][b];
    } else {
      double J[3][3];
      geo.jacobian(xi, J);
      det = det_adj<3>(J, adj);
    }
    const double adet = fabs(det);
    if constexpr (FORM == CQ_FORM_MASS) {
      const double s = w * adet;
      double f[4], sf[4];
#pragma unroll
      for (int tile = 0; tile < 4; ++tile) {
        const double v = sel3(b0.v, lat[tile][0]) * sel3(b1.v, lat[tile][1]) * sel3(b2.v, lat[tile][2]);
        f[tile] = valid[tile] ? v : 0.0;
        sf[tile] = s * f[tile];
      }
      int e = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++e) dmma_m8n8k4(D[e], sf[ti], f[tj]);
    } else if constexpr (FORM == CQ_FORM_STIFFNESS) {
      const double sc = w / adet;
      double f[4][3], sf[4][3];
#pragma unroll
      for (int tile = 0; tile < 4; ++tile) {
        const double v0 = sel3(b0.v, lat[tile][0]), v1 = sel3(b1.v, lat[tile][1]), v2 = sel3(b2.v, lat[tile][2]);
        const double d0 = sel3(b0.d, lat[tile][0]), d1 = sel3(b1.d, lat[tile][1]), d2 = sel3(b2.d, lat[tile][2]);
        const double gr[3] = {d0 * v1 * v2, v0 * d1 * v2, v0 * v1 * d2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double a = fma(adj[2][c], gr[2], fma(adj[1][c], gr[1], adj[0][c] * gr[0]));
          f[tile][c] = valid[tile] ? a : 0.0;
          sf[tile][c] = sc * f[tile][c];
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int e = 0;
#pragma unroll
        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
          for (int tj = ti; tj < 4; ++tj, ++e) dmma_m8n8k4(D[e], sf[ti][c], f[tj][c]);
      }
    } else {  // CQ_FORM_NITSCHE:  psi_i u_j - e_i psi_j,  e = s dn,  u = s gamma psi - e
      const double s = w * adet, sg = s * gamma, inv_det = 1.0 / det;
      double m[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) m[e] = fma(adj[e][2], nrm[2], fma(adj[e][1], nrm[1], adj[e][0] * nrm[0])) * inv_det;
      double psi[4], ev[4], uv[4];
#pragma unroll
      for (int tile = 0; tile < 4; ++tile) {
        const double v0 = sel3(b0.v, lat[tile][0]), v1 = sel3(b1.v, lat[tile][1]), v2 = sel3(b2.v, lat[tile][2]);
        const double d0 = sel3(b0.d, lat[tile][0]), d1 = sel3(b1.d, lat[tile][1]), d2 = sel3(b2.d, lat[tile][2]);
        const double dn = fma(m[2], v0 * v1 * d2, fma(m[1], v0 * d1 * v2, m[0] * (d0 * v1 * v2)));
        psi[tile] = valid[tile] ? v0 * v1 * v2 : 0.0;
        ev[tile] = valid[tile] ? s * dn : 0.0;
        uv[tile] = fma(sg, psi[tile], -ev[tile]);
      }
      int e = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++e) {
          dmma_m8n8k4(D[e], psi[ti], uv[tj]);
          dmma_m8n8k4(D[e], -ev[ti], psi[tj]);
        }
    }
  }
}

CQ_DEVINL void dmma_lane_dofs(const int g, int (&lat)[4][3], bool (&valid)[4]) {
#pragma unroll
  for (int tile = 0; tile < 4; ++tile) {
    const int j = 8 * tile + g;
    const int pk = c_q2hex_lat[j];
    lat[tile][0] = pk & 3; lat[tile][1] = (pk >> 2) & 3; lat[tile][2] = (pk >> 4) & 3;
    valid[tile] = j < 27;
  }
}

// staged kernel, MODE 2: one warp per stage entry; everything comes from the stage header.
//
// Scatter: the accumulator fragments hold one value per (row group, column pair) -- scattered straight from the
// registers (round 1), a RED instruction's 32 lanes hit 8 rows x 4 column pairs, i.e. ~30 different lines, and every
// value needed its own scatter position / row start / row length from shared memory: ~8000 LSU wavefronts per cell,
// the bound of config 4 (Q2, bs = 3: 2187 values per cell).  Now the warp parks the symmetric 27 x 27 block in a
// shared-memory tile of its own and scatters it ROW BY ROW with lane = local column: the 27 values of a row go to 9
// runs of 3 consecutive columns (x-neighbours are numbered consecutively), ~10 lines per instruction, and the row's
// start / length are one broadcast load per row instead of one per value.
constexpr int DMMA_TILE_STRIDE = 28;                               // doubles per row of the 27 x 27 tile
constexpr int DMMA_TILE_BYTES = (27 * DMMA_TILE_STRIDE * 8 + 127) / 128 * 128;

template <class T, int CB>
CQ_DEVINL void dmma_process_entry(const KParams& P, const StageMeta<T, CB>& M, const int e, double* __restrict__ tile) {
  constexpr int N = 27, FORM = T::FORM_ID, TS = DMMA_TILE_STRIDE;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  int lat[4][3];
  bool valid[4];
  dmma_lane_dofs(g, lat, valid);
  double X[8][3];
#pragma unroll
  for (int v = 0; v < 8; ++v)
#pragma unroll
    for (int k = 0; k < 3; ++k) X[v][k] = M.X[v * 3 + k][e];
  CellGeom<3> geo;
  geo.init(X, FORM == CQ_FORM_NITSCHE);
  init_affine<3>(geo);
  const double gamma = FORM == CQ_FORM_NITSCHE ? P.betaN / geo.hK : 0.0;
  double D[10][2];
  dmma_cell_tiles<FORM>(D, geo, gamma, M.pts[e], M.w[e], T::NEED_NORMAL ? M.nrm[e] : nullptr, M.nq[e], t, lat, valid);
  // fragments -> tile (both triangles)
  __syncwarp();  // the previous entry's scatter has read the tile
  int en = 0;
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
#pragma unroll
    for (int tj = ti; tj < 4; ++tj, ++en) {
      const int i = 8 * ti + g;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = 8 * tj + 2 * t + c;
        if (i < N && j < N) {
          tile[i * TS + j] = D[en][c];
          if (ti != tj) tile[j * TS + i] = D[en][c];  // diagonal tiles are full
        }
      }
    }
  __syncwarp();
  // row-by-row scatter, lane = local column
  const int bs = P.bs;
  if (lane < N) {
    const uint8_t* pos = &M.pos[e][lane];
#pragma unroll 3
    for (int i = 0; i < N; ++i) {
      const double v = tile[i * TS + lane];
      const unsigned p = pos[i * N];
      double* row = P.values + M.rowstart[e][i] + p * bs;
      if (bs == 1) {
        red_add_f64(row, v);
      } else {
        const int len = M.rowlen[e][i];
        for (int a = 0; a < bs; ++a) red_add_f64(row + (long long)a * len + a, v);
      }
    }
  }
}

template <int FORM, int THREADS>
__global__ void __launch_bounds__(THREADS) assemble_q2hex_dmma_kernel(const KParams P) {
  constexpr int N = 27;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t warp_id = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
  const int64_t num_warps = (int64_t)gridDim.x * (THREADS / 32);

  int lat[4][3];
  bool valid[4];
  dmma_lane_dofs(g, lat, valid);

  for (int64_t k = warp_id; k < P.num_entries; k += num_warps) {
    const int32_t cell = __ldg(P.cells + k);
    const int64_t q0 = __ldg(P.offsets + k), q1 = __ldg(P.offsets + k + 1);
    double X[8][3];
    load_cell_vertices<3>(P, cell, X);
    CellGeom<3> geo;
    geo.init(X, FORM == CQ_FORM_NITSCHE);
    init_affine<3>(geo);
    const double gamma = FORM == CQ_FORM_NITSCHE ? P.betaN / geo.hK : 0.0;

    double D[10][2];  // upper-triangular tiles (ti <= tj), row-major enumeration
    dmma_cell_tiles<FORM>(D, geo, gamma, P.pts + 3 * q0, P.w + q0, FORM == CQ_FORM_NITSCHE ? P.normals + 3 * q0 : nullptr, (int)(q1 - q0), t, lat,
                          valid);

    // ---- sink: lane holds D[ti][tj] at row 8*ti + g, cols 8*tj + 2*t + {0,1} -------------------
    const int bs = P.bs;
    const int nl = N * bs;
    if (P.elem_pos != nullptr && !P.elem_out) {
      // scatter map: CSR row starts of this lane's 4 rows (i) and 8 mirrored rows (j), positions by byte lookup
      constexpr int NPOS = (N * N + 15) / 16 * 16;
      const uint8_t* __restrict__ pos = P.elem_pos + (int64_t)cell * NPOS;
      long long rsI[4], rsJ[4][2];
      int rlI[4], rlJ[4][2];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) {
        const int i = 8 * ti + g;
        rsI[ti] = 0; rlI[ti] = 0;
        if (i < N) {
          const int64_t row = (int64_t)__ldg(P.dofmap + (int64_t)cell * N + i) * bs;
          rsI[ti] = __ldg(P.indptr + row);
          rlI[ti] = bs == 1 ? 0 : (int)(__ldg(P.indptr + row + 1) - rsI[ti]);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = 8 * ti + 2 * t + c;
          rsJ[ti][c] = 0; rlJ[ti][c] = 0;
          if (j < N) {
            const int64_t row = (int64_t)__ldg(P.dofmap + (int64_t)cell * N + j) * bs;
            rsJ[ti][c] = __ldg(P.indptr + row);
            rlJ[ti][c] = bs == 1 ? 0 : (int)(__ldg(P.indptr + row + 1) - rsJ[ti][c]);
          }
        }
      }
      int e = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++e) {
          const int i = 8 * ti + g;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int j = 8 * tj + 2 * t + c;
            if (i < N && j < N) {
              const double v = D[e][c];
              const int pij = pos[i * N + j];
              for (int a = 0; a < bs; ++a) red_add_f64(P.values + rsI[ti] + (long long)a * rlI[ti] + pij * bs + a, v);
              if (ti != tj) {  // diagonal tiles are full
                const int pji = pos[j * N + i];
                for (int a = 0; a < bs; ++a) red_add_f64(P.values + rsJ[tj][c] + (long long)a * rlJ[tj][c] + pji * bs + a, v);
              }
            }
          }
        }
      continue;
    }
    int e = 0;
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = ti; tj < 4; ++tj, ++e) {
        const int i = 8 * ti + g;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = 8 * tj + 2 * t + c;
          if (i < N && j < N) {
            const double v = D[e][c];
            if (P.elem_out) {
              double* o = P.out + k * (int64_t)(nl * nl);
              for (int a = 0; a < bs; ++a) {
                o[(i * bs + a) * nl + j * bs + a] = v;
                if (ti != tj) o[(j * bs + a) * nl + i * bs + a] = v;
              }
            } else {
              const int32_t di = __ldg(P.dofmap + (int64_t)cell * N + i);
              const int32_t dj = __ldg(P.dofmap + (int64_t)cell * N + j);
              for (int a = 0; a < bs; ++a) {
                csr_add(P, (int64_t)di * bs + a, dj * bs + a, v);
                if (ti != tj) csr_add(P, (int64_t)dj * bs + a, di * bs + a, v);  // diagonal tiles are full
              }
            }
          }
        }
      }
  }
}

}  // namespace cq
