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

// ---- point loop ------------------------------------------------------------------------------------------
// Round 1 let every lane evaluate the 1-D bases of "its" point and pick its dofs' factors out of them with
// run-time selects: the compiler turned those into divergent branches and local-memory traffic, and the eight lanes
// that share a point repeated its basis and Jacobian (measured 0.37 of the DMMA peak; the tensor core waited on
// scalar code).  Now the warp first fills a TABLE in shared memory, one row per point, one lane per point:
//   [2 i, 2 i + 1]  the pair (l_a(xi_d), l_a'(xi_d)), i = d * 3 + a: one 16-byte load per factor
//   [18, 19]        (0, 0): the padding dofs 27..31 read these
//   [20..]          the point's factors: stiffness G = w / |det J| adj adj^T (6 values), mass w |det J|,
//                   Nitsche w |det J|, w |det J| gamma, m = adj n / det J
// and the k-steps read each lane's factors through per-lane row offsets fixed for the whole kernel.  The table lives
// in the warp's 27 x 28 scatter tile (28 rows of 26 doubles: a 27-point rule is one chunk).  Row stride 26 doubles =
// 13 sixteen-byte units: the four rows a load instruction touches (points 4 s + t) put their three pairs on bank
// groups {0,1,2} {5,6,7} {2,3,4} {7,0,1} -- at most two per group, the minimum for twelve 16-byte words (the first
// layout, separate value / derivative arrays at stride 27, measured 3-4 wavefronts per 8-byte load: the shared-memory
// pipe, not the tensor core, was the busiest unit, profiles/r2s2/ncu_q2hex_phase1_stride27.txt).
// Stiffness: S_ij = sum_q g_i^T G g_j with g the REFERENCE gradient: the A fragment is g itself, the B fragment G g.
constexpr int DT_STRIDE = 26;  // doubles per table row
constexpr int DT_POINTS = 28;  // rows: 28 * 26 <= 27 * 28 doubles, the scatter tile

// this lane's row offsets (in doubles, even): dof 8 * tile + g, factor of dimension d
CQ_DEVINL void dmma_lane_offsets(const int g, int (&off)[4][3]) {
#pragma unroll
  for (int tile = 0; tile < 4; ++tile) {
    const int j = 8 * tile + g;
    const int pk = c_q2hex_lat[j];
#pragma unroll
    for (int d = 0; d < 3; ++d) off[tile][d] = 2 * (j < 27 ? d * 3 + ((pk >> (2 * d)) & 3) : 9);
  }
}

template <int FORM>
CQ_DEVINL void dmma_fill_table(double* __restrict__ tab, const CellGeom<3>& geo, const double gamma, const double* pts, const double* w_,
                               const double* nrm_, const int q0, const int nqc, const int lane) {
  const int rows = (nqc + 3) & ~3;
  for (int r = lane; r < rows; r += 32) {
    double* __restrict__ row = tab + r * DT_STRIDE;
    if (r >= nqc) {  // padding point of the last k-step: contributes nothing
#pragma unroll
      for (int i = 0; i < DT_STRIDE; ++i) row[i] = 0.0;
      continue;
    }
    const int q = q0 + r;
    double xi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) xi[d] = pts[3 * q + d];
    const double w = w_[q];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const Basis1D<2> b(xi[d]);
#pragma unroll
      for (int a = 0; a < 3; ++a) *reinterpret_cast<double2*>(row + 2 * (d * 3 + a)) = make_double2(b.v[a], b.d[a]);
    }
    *reinterpret_cast<double2*>(row + 18) = make_double2(0.0, 0.0);
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
      row[20] = w * adet;
    } else if constexpr (FORM == CQ_FORM_STIFFNESS) {
      const double sc = w / adet;
      int k = 20;
#pragma unroll
      for (int e = 0; e < 3; ++e)
#pragma unroll
        for (int f = e; f < 3; ++f, ++k) row[k] = sc * fma(adj[e][2], adj[f][2], fma(adj[e][1], adj[f][1], adj[e][0] * adj[f][0]));
    } else {
      const double s = w * adet, inv_det = 1.0 / det;
      row[20] = s;
      row[21] = s * gamma;
#pragma unroll
      for (int e = 0; e < 3; ++e)
        row[22 + e] = fma(adj[e][2], nrm_[3 * q + 2], fma(adj[e][1], nrm_[3 * q + 1], adj[e][0] * nrm_[3 * q])) * inv_det;
    }
  }
}

// k-steps over the table rows [0, rows): accumulates into the 10 upper-triangular 8x8 tiles D
template <int FORM>
CQ_DEVINL void dmma_table_steps(double (&D)[10][2], const double* __restrict__ tab, const int rows, const int t, const int (&off)[4][3],
                                const bool diagG) {
  for (int qb = 0; qb < rows; qb += 4) {
    const double* __restrict__ row = tab + (qb + t) * DT_STRIDE;
    if constexpr (FORM == CQ_FORM_MASS) {
      const double s = row[20];
      double f[4], sf[4];
#pragma unroll
      for (int tile = 0; tile < 4; ++tile) {
        f[tile] = row[off[tile][0]] * row[off[tile][1]] * row[off[tile][2]];  // the values of the pairs
        sf[tile] = s * f[tile];
      }
      int e = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++e) dmma_m8n8k4(D[e], sf[ti], f[tj]);
    } else if constexpr (FORM == CQ_FORM_STIFFNESS) {
      double gr[4][3];
#pragma unroll
      for (int tile = 0; tile < 4; ++tile) {
        const double2 p0 = *reinterpret_cast<const double2*>(row + off[tile][0]);  // (value, derivative)
        const double2 p1 = *reinterpret_cast<const double2*>(row + off[tile][1]);
        const double2 p2 = *reinterpret_cast<const double2*>(row + off[tile][2]);
        gr[tile][0] = p0.y * (p1.x * p2.x);
        gr[tile][1] = (p0.x * p2.x) * p1.y;
        gr[tile][2] = (p0.x * p1.x) * p2.y;
      }
      const double G00 = row[20], G11 = row[23], G22 = row[25];
      if (diagG) {  // axis-aligned cell (every cell of a create_box mesh): G is diagonal
        const double Gd[3] = {G00, G11, G22};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double b[4];
#pragma unroll
          for (int tile = 0; tile < 4; ++tile) b[tile] = Gd[c] * gr[tile][c];
          int e = 0;
#pragma unroll
          for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int tj = ti; tj < 4; ++tj, ++e) dmma_m8n8k4(D[e], gr[ti][c], b[tj]);
        }
      } else {
        const double G01 = row[21], G02 = row[22], G12 = row[24];
        const double G[3][3] = {{G00, G01, G02}, {G01, G11, G12}, {G02, G12, G22}};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double b[4];
#pragma unroll
          for (int tile = 0; tile < 4; ++tile) b[tile] = fma(G[c][2], gr[tile][2], fma(G[c][1], gr[tile][1], G[c][0] * gr[tile][0]));
          int e = 0;
#pragma unroll
          for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int tj = ti; tj < 4; ++tj, ++e) dmma_m8n8k4(D[e], gr[ti][c], b[tj]);
        }
      }
    } else {  // CQ_FORM_NITSCHE:  psi_i u_j - e_i psi_j,  e = s dn,  u = s gamma psi - e
      const double s = row[20], sg = row[21], m0 = row[22], m1 = row[23], m2 = row[24];
      double psi[4], ev[4], uv[4];
#pragma unroll
      for (int tile = 0; tile < 4; ++tile) {
        const double2 p0 = *reinterpret_cast<const double2*>(row + off[tile][0]);  // (value, derivative)
        const double2 p1 = *reinterpret_cast<const double2*>(row + off[tile][1]);
        const double2 p2 = *reinterpret_cast<const double2*>(row + off[tile][2]);
        const double v12 = p1.x * p2.x;
        const double dn = fma(m2, (p0.x * p1.x) * p2.y, fma(m1, (p0.x * p2.x) * p1.y, m0 * (p0.y * v12)));
        psi[tile] = p0.x * v12;
        ev[tile] = s * dn;
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

// point loop of one cell: the 10 upper-triangular 8x8 tiles D of the scalar 27x27 block.
// pts / w / nrm: the cell's points (global or shared memory); tab: the warp's table (all 32 lanes call this).
template <int FORM>
CQ_DEVINL void dmma_cell_tiles(double (&D)[10][2], const CellGeom<3>& geo, const double gamma, const double* pts, const double* w_,
                               const double* nrm_, const int nq, const int lane, const int (&off)[4][3], double* __restrict__ tab) {
#pragma unroll
  for (int i = 0; i < 10; ++i) D[i][0] = D[i][1] = 0.0;
  // diagonal Jacobian of an affine cell: adj is diagonal, and so is G
  const bool diagG = geo.affine && geo.adj0[0][1] == 0.0 && geo.adj0[0][2] == 0.0 && geo.adj0[1][0] == 0.0 && geo.adj0[1][2] == 0.0 &&
                     geo.adj0[2][0] == 0.0 && geo.adj0[2][1] == 0.0;
  for (int q0 = 0; q0 < nq; q0 += DT_POINTS) {
    const int nqc = nq - q0 < DT_POINTS ? nq - q0 : DT_POINTS;
    __syncwarp();  // the table's previous readers (k-steps of the last chunk, scatter of the last entry) are done
    dmma_fill_table<FORM>(tab, geo, gamma, pts, w_, nrm_, q0, nqc, lane);
    __syncwarp();
    dmma_table_steps<FORM>(D, tab, (nqc + 3) & ~3, lane & 3, off, diagG);
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
  int off[4][3];
  dmma_lane_offsets(g, off);
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
  dmma_cell_tiles<FORM>(D, geo, gamma, M.pts[e], M.w[e], T::NEED_NORMAL ? M.nrm[e] : nullptr, M.nq[e], lane, off, tile);
  // fragments -> tile (both triangles)
  __syncwarp();  // the k-steps have read the table, which lives in the tile
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
  if (P.elem_scalar != nullptr) {
    // owner-computes path (cq_owner.cuh): the block leaves as 729 consecutive doubles, row-major, at list position k
    double* __restrict__ o = P.elem_scalar + (M.k0 + e) * (long long)(N * N);
#pragma unroll 4
    for (int idx = lane; idx < N * N; idx += 32) {
      const int i = idx / N;
      o[idx] = tile[i * TS + (idx - i * N)];
    }
    return;
  }
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

  int off[4][3];
  dmma_lane_offsets(g, off);
  __shared__ __align__(16) double tab_all[THREADS / 32][DT_POINTS * DT_STRIDE];
  double* tab = tab_all[threadIdx.x >> 5];

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
    dmma_cell_tiles<FORM>(D, geo, gamma, P.pts + 3 * q0, P.w + q0, FORM == CQ_FORM_NITSCHE ? P.normals + 3 * q0 : nullptr, (int)(q1 - q0), lane, off,
                          tab);

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
