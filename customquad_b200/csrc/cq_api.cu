// cq_api.cu -- the C-ABI of libcq_b200.so (include/cq_b200.h): argument checking, context,
// parameter packing, and the host-buffer (end-to-end) entry points.
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "cq_internal.h"
#include "cq_levelset.cuh"

namespace cq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return CQ_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return CQ_ERR_CUDA;
}

// base address and size of the allocation containing p (driver API cuMemGetAddressRange, resolved through the
// runtime so that the library keeps linking against libcudart only)
static int cuMemGetAddressRange_fallback(const void* p, void** base, size_t* size) {
  typedef int (*fn_t)(unsigned long long*, size_t*, unsigned long long);
  static fn_t fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
    fn = reinterpret_cast<fn_t>(f);
  }
  unsigned long long b = 0;
  size_t s = 0;
  if (fn(&b, &s, (unsigned long long)(uintptr_t)p) != 0) return -1;
  *base = reinterpret_cast<void*>((uintptr_t)b);
  *size = s;
  return 0;
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

static int gdim_of(int cell_type) { return cell_type == CQ_CELL_QUADRILATERAL ? 2 : 3; }

static int rank_of(int form) { return form < 10 ? 2 : (form < 20 ? 1 : 0); }

// Validates the descriptor and packs what the kernels need.
static int pack(cq_ctx* ctx, const cq_form_desc* f, const cq_mesh* mesh, const cq_space* space, const cq_quadrature* qr,
                KParams& P) {
  if (!ctx || !f || !mesh || !qr) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (f->cell_type != CQ_CELL_QUADRILATERAL && f->cell_type != CQ_CELL_HEXAHEDRON) {
    // the reference's permutation table has no entry for other cells (call_basix.h:60-76)
    set_error("cell_type %d unsupported: only quadrilateral (4) and hexahedron (5)", f->cell_type);
    return CQ_ERR_UNSUPPORTED;
  }
  if (f->degree != 1 && f->degree != 2) { set_error("degree %d unsupported: only 1 and 2 (call_basix.h:63-73)", f->degree); return CQ_ERR_UNSUPPORTED; }
  if (f->bs < 1 || f->bs > 3) { set_error("bs %d unsupported", f->bs); return CQ_ERR_UNSUPPORTED; }
  if (f->nconsts < 0 || f->nconsts > CQ_MAX_CONSTS || (f->nconsts > 0 && !f->consts)) { set_error("bad consts"); return CQ_ERR_INVALID; }
  const int rank = rank_of(f->form);
  const int gdim = gdim_of(f->cell_type);
  const bool nitsche = f->form == CQ_FORM_NITSCHE || f->form == CQ_FORM_NITSCHE_RHS;
  const bool need_coef = rank < 2 && (f->fn_kind == CQ_FN_COEF || f->form == CQ_FORM_L2_ERR || f->form == CQ_FORM_H10_ERR);
  if (qr->num_entries < 0 || qr->num_points < 0) { set_error("negative sizes"); return CQ_ERR_INVALID; }
  if (qr->num_entries > 0 && (!qr->cells || !qr->offsets)) { set_error("qr.cells / qr.offsets null"); return CQ_ERR_INVALID; }
  if (qr->num_points > 0 && (!qr->pts || !qr->weights)) { set_error("qr.pts / qr.weights null"); return CQ_ERR_INVALID; }
  if (nitsche && qr->num_points > 0 && !qr->normals) { set_error("this form needs quadrature normals (a 4-tuple qr)"); return CQ_ERR_INVALID; }
  if (!mesh->geom_dofmap || !mesh->coords) { set_error("mesh arrays null"); return CQ_ERR_INVALID; }
  if ((rank >= 1 || need_coef) && (!space || !space->dofmap)) { set_error("space / dofmap null"); return CQ_ERR_INVALID; }
  if (need_coef && !space->coeffs) { set_error("form needs a packed coefficient (space.coeffs)"); return CQ_ERR_INVALID; }
  // vectorised loads: 16 B for 2-D point/normal pairs and the geometry dofmap, 8 B elsewhere
  if (!aligned(mesh->geom_dofmap, 16) || (gdim == 2 && (!aligned(qr->pts, 16) || !aligned(qr->normals, 16))) ||
      !aligned(qr->pts, 8) || !aligned(qr->weights, 8) || !aligned(qr->offsets, 8) || !aligned(mesh->coords, 8)) {
    set_error("alignment: geom_dofmap and 2-D pts/normals need 16 B, f64/i64 arrays 8 B");
    return CQ_ERR_ALIGN;
  }
  memset(&P, 0, sizeof P);
  P.peer_row0 = 0x7fffffffffffffffLL;
  P.num_entries = qr->num_entries;
  P.cells = qr->cells; P.offsets = qr->offsets; P.pts = qr->pts; P.w = qr->weights; P.normals = qr->normals;
  P.geom_dofmap = mesh->geom_dofmap; P.coords = mesh->coords;
  P.dofmap = space ? space->dofmap : nullptr;
  P.coeffs = space ? space->coeffs : nullptr;
  P.err_flag = ctx->err_flag;
  P.elem_pos = space ? space->elem_pos : nullptr;
  P.work_counter = ctx->work_counter;
  P.mesh_flags = mesh->flags;
  P.num_points = qr->num_points;
  if (qr->plane_entries >= 64 && qr->plane_tiles > 1 && qr->num_entries > qr->plane_entries) {
    P.tile_sp = qr->plane_entries / 32;
    P.tile_sg = (P.tile_sp + qr->plane_tiles - 1) / qr->plane_tiles;
    if (P.tile_sg < 1) P.tile_sg = 1;
  }
  P.bs = f->bs;
  for (int i = 0; i < f->nconsts; ++i) P.p[i] = f->consts[i];
  P.betaN = nitsche ? (f->nconsts > 0 ? f->consts[0] : 0.0) : 0.0;
  if (nitsche && f->nconsts < 1) { set_error("Nitsche forms need consts[0] = betaN"); return CQ_ERR_INVALID; }
  return CQ_OK;
}

static int set_device(cq_ctx* ctx) { return check_cuda(cudaSetDevice(ctx->device), "cudaSetDevice"); }

__global__ void accumulate_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += x[i];
}

// values[indptr[row0 + r] + i] += recv[recv_indptr[r] + i]: one warp per row
__global__ void accumulate_rows_kernel(double* __restrict__ values, const int64_t* __restrict__ indptr, int64_t row0, int64_t nrows,
                                       const int64_t* __restrict__ recv_indptr, const double* __restrict__ recv) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t dst = indptr[row0 + r], src = recv_indptr[r], n = recv_indptr[r + 1] - src;
    for (int64_t i = lane; i < n; i += 32) values[dst + i] += recv[src + i];
  }
}

// *flag &= (every cell has a constant, diagonal, non-singular Jacobian)
template <int DIM>
__global__ void classify_kernel(const KParams P, int64_t num_cells, int* flag) {
  bool ok = true;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < num_cells; c += (int64_t)gridDim.x * blockDim.x) {
    double X[1 << DIM][DIM];
    load_cell_vertices<DIM>(P, (int32_t)c, X);
    CellGeom<DIM> geo;
    geo.init(X, false);
    ok = ok && geo.affine;
#pragma unroll
    for (int e = 0; e < DIM; ++e)
#pragma unroll
      for (int k = 0; k < DIM; ++k) ok = ok && (k == e ? geo.c[1 << e][k] != 0.0 : geo.c[1 << e][k] == 0.0);
  }
  if (!ok) atomicAnd(flag, 0);
}

// cut-cell quadrature of a disk (cq_levelset.cuh): one thread per cell / per cut cell
__global__ void circle_classify_kernel(const CircleGrid G, const int64_t ncells, int32_t* __restrict__ cls, int32_t* __restrict__ nbulk,
                                       int32_t* __restrict__ nsurf) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  int32_t nb, ns;
  cls[c] = ls_classify_count(G, c, &nb, &ns);
  nbulk[c] = nb;
  nsurf[c] = ns;
}

__global__ void circle_rules_kernel(const CircleGrid G, const int64_t ncut, const int32_t* __restrict__ cut_cells,
                                    const int64_t* __restrict__ off_bulk, const int64_t* __restrict__ off_surf, double* __restrict__ pts,
                                    double* __restrict__ w, double* __restrict__ pts_b, double* __restrict__ w_b, double* __restrict__ nrm) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncut) return;
  double lo[2], hi[2], wsum;
  ls_cell_box(G, cut_cells[k], lo, hi);
  const int64_t ob = off_bulk[k], os = off_surf[k];
  ls_bulk_rule(G, lo, hi, pts + 2 * ob, w + ob, &wsum);
  ls_surface_rule(G, lo, hi, pts_b + 2 * os, w_b + os, nrm + 2 * os);
}

// Classification in two passes.  Counting the points of a cut cell means running the whole rule construction
// (ls_sphere_rules without outputs), ~1e5 FP64 operations, while the box test of the other 99 % of the cells is a few
// dozen: with one thread per CELL doing both (round 1), nearly every warp that meets the surface dragged 31 idle lanes
// through one cell's rule construction -- 511 ms for the 256^3 mesh of config 3.  Pass 1 is the box test alone and
// compacts the cut cells; pass 2 counts with one thread per CUT cell (dense warps).
__global__ void sphere_classify_box_kernel(const SphereGrid G, const int64_t ncells, int32_t* __restrict__ cls, int32_t* __restrict__ nbulk,
                                           int32_t* __restrict__ nsurf, int32_t* __restrict__ cut_list, unsigned int* __restrict__ cut_count) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  double lo[3], hi[3];
  ls_cell_box3(G, c, lo, hi);
  const int32_t k = ls_classify_box3(G, lo, hi);
  cls[c] = k;
  nbulk[c] = 0;
  nsurf[c] = 0;
  if (k == LS_CUT) cut_list[atomicAdd(cut_count, 1u)] = (int32_t)c;
}

__global__ void sphere_count_kernel(const SphereGrid G, const int32_t* __restrict__ cut_list, const unsigned int* __restrict__ cut_count,
                                    int32_t* __restrict__ cls, int32_t* __restrict__ nbulk, int32_t* __restrict__ nsurf) {
  const unsigned int n = *cut_count;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int32_t c = cut_list[k];
    int32_t nb, ns;
    const int32_t kind = ls_classify_count3(G, c, &nb, &ns);  // may demote a grazed cell to "outside"
    cls[c] = kind;
    nbulk[c] = nb;
    nsurf[c] = ns;
  }
}

__global__ void sphere_rules_kernel(const SphereGrid G, const int64_t ncut, const int32_t* __restrict__ cut_cells,
                                    const int64_t* __restrict__ off_bulk, const int64_t* __restrict__ off_surf, double* __restrict__ pts,
                                    double* __restrict__ w, double* __restrict__ pts_b, double* __restrict__ w_b, double* __restrict__ nrm) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncut) return;
  double lo[3], hi[3], wsum;
  int nb, ns;
  ls_cell_box3(G, cut_cells[k], lo, hi);
  const int64_t ob = off_bulk[k], os = off_surf[k];
  ls_sphere_rules(G, lo, hi, pts + 3 * ob, w + ob, pts_b + 3 * os, w_b + os, nrm + 3 * os, &nb, &ns, &wsum);
}

// scatter positions: one thread per (cell, i, j)
__global__ void scatter_map_kernel(const int32_t* __restrict__ dofmap, int64_t num_cells, int n, int stride, int bs,
                                   const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   uint8_t* __restrict__ out, int* err) {
  const int64_t total = num_cells * n * n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / (n * n);
    const int ij = (int)(t - c * n * n), i = ij / n, j = ij - i * n;
    const int64_t row = (int64_t)dofmap[c * n + i] * bs;
    const int32_t col = dofmap[c * n + j] * bs;
    const int64_t s = csr_find(indptr, indices, row, col);
    const int64_t rel = s < 0 ? -1 : (s - indptr[row]) / bs;
    if (s < 0 || rel > 255 || (s - indptr[row]) % bs != 0) atomicExch(err, s < 0 ? 1 : 2);
    out[c * stride + ij] = (uint8_t)(rel < 0 ? 0 : rel);
  }
}

// blocked row r*bs + a: columns c*bs + b of the scalar row r; one warp per scalar row
__global__ void expand_blocks_kernel(const int64_t num_rows, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const int bs,
                                     int64_t* __restrict__ indptr_out, int32_t* __restrict__ indices_out) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < num_rows; r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t s = indptr[r], len = indptr[r + 1] - s;
    const int64_t base = s * bs * bs, blen = len * bs;  // blocked rows of r start at base, base + blen, ...
    if (lane < bs) indptr_out[r * bs + lane] = base + lane * blen;
    if (r == num_rows - 1 && lane == 0) indptr_out[num_rows * bs] = base + bs * blen;
    for (int64_t i = lane; i < blen; i += 32) {
      const int32_t col = indices[s + i / bs] * bs + (int32_t)(i % bs);
      for (int a = 0; a < bs; ++a) indices_out[base + a * blen + i] = col;
    }
  }
}

// Ghost penalty on axis-aligned cells: one thread per (facet, i, j), i, j in the 2N dofs of the two cells.
//   M_ij = betas * avg(h) * |F| * sum_q gw_q J_i(q) J_j(q),  J = (d_ax psi^lo / J_ax^lo, -d_ax psi^hi / J_ax^hi) at the facet points
// (tensor Gauss rule with DEG + 1 points per facet direction: exact, the integrand has degree 2 DEG per direction).
template <int DIM, int DEG>
__global__ void ghost_penalty_kernel(const KParams P, const int64_t num_facets, const int32_t* __restrict__ facets, const double betas) {
  constexpr int N = Lattice<DIM, DEG>::N, N2 = 2 * N, NG = DEG + 1, NQ = DIM == 2 ? NG : NG * NG;
  const double g2[2] = {0.21132486540518713, 0.7886751345948129}, w2[2] = {0.5, 0.5};
  const double g3[3] = {0.1127016653792583, 0.5, 0.8872983346207417}, w3[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  const int64_t total = num_facets * N2 * N2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = t / (N2 * N2);
    const int ij = (int)(t - f * N2 * N2), i = ij / N2, j = ij - i * N2;
    const int32_t cell[2] = {facets[3 * f], facets[3 * f + 1]};
    const int ax = facets[3 * f + 2];
    double Jax[2], hK[2], area = 1.0;
    for (int s = 0; s < 2; ++s) {
      double X[1 << DIM][DIM];
      load_cell_vertices<DIM>(P, cell[s], X);
      CellGeom<DIM> geo;
      geo.init(X, true);
      hK[s] = geo.hK;
      Jax[s] = geo.c[1 << ax][ax];
      if (s == 0)
        for (int d = 0; d < DIM; ++d)
          if (d != ax) area *= fabs(geo.c[1 << d][d]);
    }
    double m = 0.0;
    for (int q = 0; q < NQ; ++q) {
      // facet point: Gauss coordinates in the directions != ax; xi_ax = 1 in the low cell, 0 in the high cell
      double gw = 1.0, xi[2][DIM];
      int qq = q;
      for (int d = 0; d < DIM; ++d) {
        if (d == ax) { xi[0][d] = 1.0; xi[1][d] = 0.0; continue; }
        const int k = qq % NG;
        qq /= NG;
        const double gp = DEG == 1 ? g2[k] : g3[k];
        gw *= DEG == 1 ? w2[k] : w3[k];
        xi[0][d] = xi[1][d] = gp;
      }
      double Ji = 0.0, Jj = 0.0;
      for (int s = 0; s < 2; ++s) {
        double psi[N], gr[N][DIM];
        tabulate_point<DIM, DEG, false, true>(xi[s], psi, gr);
        const double sgn = (s == 0 ? 1.0 : -1.0) / Jax[s];
        for (int k = 0; k < N; ++k) {
          double v = 0.0;
          for (int d = 0; d < DIM; ++d) v = d == ax ? gr[k][d] : v;
          if (i == s * N + k) Ji = sgn * v;
          if (j == s * N + k) Jj = sgn * v;
        }
      }
      m = fma(gw * Ji, Jj, m);
    }
    m *= betas * 0.5 * (hK[0] + hK[1]) * area;
    const int32_t di = __ldg(P.dofmap + (int64_t)cell[i / N] * N + (i % N));
    const int32_t dj = __ldg(P.dofmap + (int64_t)cell[j / N] * N + (j % N));
    for (int a = 0; a < P.bs; ++a) csr_add(P, (int64_t)di * P.bs + a, dj * P.bs + a, m);
  }
}

// rowmin[c] = min over the entries k in chunk c of the smallest (blocked) row cell cells[k] touches
__global__ void chunk_rowmin_kernel(const int32_t* __restrict__ cells, int64_t num_entries, int nchunks, const int32_t* __restrict__ dofmap,
                                    int n, int bs, int32_t* __restrict__ rowmin) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < num_entries; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c0 = cells[k];
    int32_t m = 0x7fffffff;
    for (int i = 0; i < n; ++i) m = min(m, dofmap[c0 * n + i]);
    // chunk of entry k: chunks are [C*c/nchunks, C*(c+1)/nchunks)
    int c = (int)((k * nchunks) / num_entries);
    while (c + 1 < nchunks && k >= num_entries * (c + 1) / nchunks) ++c;
    while (c > 0 && k < num_entries * c / nchunks) --c;
    m = __reduce_min_sync(__activemask(), m);  // same chunk for (almost) the whole warp: one atomic per warp in the common case
    const int cw = __shfl_sync(__activemask(), c, __ffs(__activemask()) - 1);
    if (__all_sync(__activemask(), c == cw)) {
      if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicMin(rowmin + c, m * bs);
    } else {
      int32_t mine = 0x7fffffff;
      for (int i = 0; i < n; ++i) mine = min(mine, dofmap[c0 * n + i]);
      atomicMin(rowmin + c, mine * bs);
    }
  }
}

// call_basix on the device: one thread per (point, dof)
template <int DIM, int DEG>
__global__ void tabulate_kernel(const double* __restrict__ pts, int64_t nq, int deriv, double* __restrict__ out) {
  constexpr int N = Lattice<DIM, DEG>::N;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    double xi[DIM], psi[N], g[N][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) xi[d] = pts[DIM * q + d];
    tabulate_point<DIM, DEG, true, true>(xi, psi, g);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double v = psi[j];
#pragma unroll
      for (int e = 0; e < DIM; ++e)
        if (deriv == e + 1) v = g[j][e];
      out[q * N + j] = v;
    }
  }
}

}  // namespace cq

using namespace cq;

extern "C" {

int cq_abi_version(void) { return CQ_B200_ABI_VERSION; }

const char* cq_last_error(void) { return g_err; }

int cq_ctx_create(int device, cq_ctx** out) {
  if (!out) { set_error("null out"); return CQ_ERR_INVALID; }
  *out = nullptr;
  int n = 0;
  if (int rc = check_cuda(cudaGetDeviceCount(&n), "cudaGetDeviceCount")) return rc;
  if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return CQ_ERR_INVALID; }
  if (int rc = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return rc;
  cudaDeviceProp prop;
  if (int rc = check_cuda(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return rc;
  if (prop.major != 10) { set_error("libcq_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor); return CQ_ERR_UNSUPPORTED; }
  cq_ctx* c = new (std::nothrow) cq_ctx();
  if (!c) { set_error("out of host memory"); return CQ_ERR_INVALID; }
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->coop_launch = prop.cooperativeLaunch;
  c->fused_zero = 2;
  int rc = check_cuda(cudaMalloc(&c->err_flag, 2 * sizeof(int)), "cudaMalloc(err_flag)");
  if (!rc) rc = check_cuda(cudaMemset(c->err_flag, 0, 2 * sizeof(int)), "cudaMemset(err_flag)");
  if (!rc) rc = check_cuda(cudaMalloc(&c->work_counter, sizeof(unsigned int)), "cudaMalloc(work_counter)");
  if (!rc) rc = check_cuda(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device), "cudaDeviceGetAttribute");
  if (!rc) rc = check_cuda(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  if (!rc) rc = check_cuda(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  for (int i = 0; i < 12 && !rc; ++i) rc = check_cuda(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming), "cudaEventCreate");
  for (int i = 0; i < 16 && !rc; ++i) rc = check_cuda(cudaEventCreateWithFlags(&c->kev[i], cudaEventDisableTiming), "cudaEventCreate");
  if (!rc) rc = check_cuda(cudaMalloc(&c->chunk_rowmin, 16 * sizeof(int32_t)), "cudaMalloc(chunk_rowmin)");
  if (!rc) rc = check_cuda(cudaHostAlloc(&c->host_rowmin, 16 * sizeof(int32_t), cudaHostAllocDefault), "cudaHostAlloc");
  if (!rc) rc = check_cuda(cudaHostAlloc(&c->host_vallo, 16 * sizeof(int64_t), cudaHostAllocDefault), "cudaHostAlloc");
  if (!rc) rc = check_cuda(cudaHostAlloc(&c->host_flag, sizeof(int), cudaHostAllocDefault), "cudaHostAlloc");
  if (rc) { cq_ctx_destroy(c); return rc; }
  *out = c;
  return CQ_OK;
}

int cq_ctx_destroy(cq_ctx* ctx) {
  if (!ctx) return CQ_OK;
  cudaSetDevice(ctx->device);
  comm_release(ctx);
  if (ctx->err_flag) cudaFree(ctx->err_flag);
  if (ctx->work_counter) cudaFree(ctx->work_counter);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->scalar_dev) cudaFree(ctx->scalar_dev);
  if (ctx->zero_buf) cudaFree(ctx->zero_buf);
  if (ctx->owner_ws) cudaFree(ctx->owner_ws);
  if (ctx->cut_list) cudaFree(ctx->cut_list);
  if (ctx->cut_count) cudaFree(ctx->cut_count);
  if (ctx->host_flag) cudaFreeHost(ctx->host_flag);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->kev) if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  if (ctx->chunk_rowmin) cudaFree(ctx->chunk_rowmin);
  if (ctx->host_rowmin) cudaFreeHost(ctx->host_rowmin);
  if (ctx->host_vallo) cudaFreeHost(ctx->host_vallo);
  delete ctx;
  return CQ_OK;
}

int cq_ctx_check(cq_ctx* ctx, void* stream) {
  if (!ctx) { set_error("null ctx"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = check_cuda(cudaMemcpyAsync(ctx->host_flag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s), "err_flag D2H")) return rc;
  if (int rc = check_cuda(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), s), "err_flag reset")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize")) return rc;
  if (*ctx->host_flag) {
    set_error("a cell produced a (row, col) pair that is absent from the CSR sparsity pattern");
    return CQ_ERR_SPARSITY;
  }
  return CQ_OK;
}

int64_t cq_ctx_launch_count(const cq_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int cq_ctx_set_variant(cq_ctx* ctx, int variant) {
  if (!ctx || variant < 0 || variant > 3) { set_error("bad variant"); return CQ_ERR_INVALID; }
  ctx->variant = variant;
  return CQ_OK;
}

static int assemble_matrix_impl(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                                const cq_quadrature* qr, const cq_csr* A, void* stream, const bool zero_first) {
  KParams P;
  if (int rc = pack(ctx, form, mesh, space, qr, P)) return rc;
  if (rank_of(form->form) != 2) { set_error("form %d is not a bilinear form", form->form); return CQ_ERR_INVALID; }
  if (!A || !A->indptr || !A->indices || !A->values) { set_error("CSR arrays null"); return CQ_ERR_INVALID; }
  if (A->num_rows != space->num_dofs * form->bs) { set_error("CSR rows %lld != num_dofs*bs %lld", (long long)A->num_rows, (long long)(space->num_dofs * form->bs)); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  const bool keep = A->zero_keep_hi > A->zero_keep_lo;
  if (keep && (A->zero_keep_lo < 0 || A->zero_keep_hi > A->nnz)) { set_error("zero_keep range outside [0, nnz]"); return CQ_ERR_INVALID; }
  if (qr->num_entries == 0) {
    if (!zero_first) return CQ_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!keep) return check_cuda(cudaMemsetAsync(A->values, 0, sizeof(double) * (size_t)A->nnz, s), "zeroEntries");
    if (int rc = check_cuda(cudaMemsetAsync(A->values, 0, sizeof(double) * (size_t)A->zero_keep_lo, s), "zeroEntries")) return rc;
    return check_cuda(cudaMemsetAsync(A->values + A->zero_keep_hi, 0, sizeof(double) * (size_t)(A->nnz - A->zero_keep_hi), s), "zeroEntries");
  }
  if (zero_first && (A->flags & CQ_CSR_PEER_TARGET) && !keep) {
    set_error("this matrix receives a neighbour rank's peer scatter: zero the interface rows, fence, then assemble with cq_csr.zero_keep_* naming them (cq_peer_fence / SlabExchange.fence)");
    return CQ_ERR_UNSUPPORTED;
  }
  P.zero_req = zero_first ? 1 : 0;
  P.zero_keep_lo = keep ? A->zero_keep_lo : 0;
  P.zero_keep_hi = keep ? A->zero_keep_hi : 0;
  P.num_dofs = space->num_dofs;
  P.nnz = A->nnz;
  P.indptr = A->indptr; P.indices = A->indices; P.values = A->values;
  if (A->peer_rowstart) {
    if (A->peer_row0 < 0 || A->peer_row0 > A->num_rows) { set_error("peer_row0 out of range"); return CQ_ERR_INVALID; }
    if (!space->elem_pos || ctx->variant == 1) { set_error("peer scatter needs the scatter map (staged kernels)"); return CQ_ERR_UNSUPPORTED; }
    P.peer_row0 = A->peer_row0;
    P.peer_rowstart = reinterpret_cast<const long long*>(A->peer_rowstart);
  }
  return launch_matrix(ctx, *form, P, static_cast<cudaStream_t>(stream));
}

int cq_ctx_set_fused_zero(cq_ctx* ctx, int on) {
  if (!ctx) { set_error("null ctx"); return CQ_ERR_INVALID; }
  ctx->fused_zero = on < 0 ? 2 : (on != 0);
  return CQ_OK;
}

int cq_assemble_matrix(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                       const cq_quadrature* qr, const cq_csr* A, void* stream) {
  return assemble_matrix_impl(ctx, form, mesh, space, qr, A, stream, false);
}

int cq_assemble_matrix_zero(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                            const cq_quadrature* qr, const cq_csr* A, void* stream) {
  return assemble_matrix_impl(ctx, form, mesh, space, qr, A, stream, true);
}

int cq_assemble_vector(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                       const cq_quadrature* qr, double* b, void* stream) {
  KParams P;
  if (int rc = pack(ctx, form, mesh, space, qr, P)) return rc;
  if (rank_of(form->form) != 1) { set_error("form %d is not a linear form", form->form); return CQ_ERR_INVALID; }
  if (!b) { set_error("b null"); return CQ_ERR_INVALID; }
  if (qr->num_entries == 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  P.out = b;
  return launch_vector(ctx, *form, P, static_cast<cudaStream_t>(stream));
}

int cq_assemble_scalar(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                       const cq_quadrature* qr, double* m, void* stream) {
  KParams P;
  if (int rc = pack(ctx, form, mesh, space, qr, P)) return rc;
  if (rank_of(form->form) != 0) { set_error("form %d is not a functional", form->form); return CQ_ERR_INVALID; }
  if (!m) { set_error("m null"); return CQ_ERR_INVALID; }
  if (qr->num_entries == 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  P.out = m;
  return launch_scalar(ctx, *form, P, static_cast<cudaStream_t>(stream));
}

int cq_element_tensors(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                       const cq_quadrature* qr, double* out, void* stream) {
  KParams P;
  if (int rc = pack(ctx, form, mesh, space, qr, P)) return rc;
  if (!out) { set_error("out null"); return CQ_ERR_INVALID; }
  if (qr->num_entries == 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  P.out = out;
  P.elem_out = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (rank_of(form->form)) {
    case 2: return launch_matrix(ctx, *form, P, s);
    case 1: return launch_vector(ctx, *form, P, s);
    default: return launch_scalar(ctx, *form, P, s);
  }
}

int cq_mesh_classify(cq_ctx* ctx, int cell_type, const cq_mesh* mesh, int32_t* flags, void* stream) {
  if (!ctx || !mesh || !flags || !mesh->geom_dofmap || !mesh->coords) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (cell_type != CQ_CELL_QUADRILATERAL && cell_type != CQ_CELL_HEXAHEDRON) { set_error("bad cell_type"); return CQ_ERR_UNSUPPORTED; }
  *flags = 0;
  if (mesh->num_cells <= 0) return CQ_OK;
  if (!aligned(mesh->geom_dofmap, 16)) { set_error("geom_dofmap needs 16-B alignment"); return CQ_ERR_ALIGN; }
  if (int rc = set_device(ctx)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KParams P;
  memset(&P, 0, sizeof P);
  P.geom_dofmap = mesh->geom_dofmap;
  P.coords = mesh->coords;
  int one = 1;
  int* scratch = ctx->err_flag + 1;  // own word: a pending sparsity violation in err_flag[0] survives this call
  if (int rc = check_cuda(cudaMemcpyAsync(scratch, &one, sizeof(int), cudaMemcpyHostToDevice, s), "flag init")) return rc;
  const int grid = (int)((mesh->num_cells + 255) / 256 < (int64_t)ctx->num_sms * 8 ? (mesh->num_cells + 255) / 256 : (int64_t)ctx->num_sms * 8);
  if (cell_type == CQ_CELL_QUADRILATERAL) classify_kernel<2><<<grid, 256, 0, s>>>(P, mesh->num_cells, scratch);
  else classify_kernel<3><<<grid, 256, 0, s>>>(P, mesh->num_cells, scratch);
  ctx->launches++;
  if (int rc = check_cuda(cudaGetLastError(), "classify_kernel launch")) return rc;
  if (int rc = check_cuda(cudaMemcpyAsync(ctx->host_flag, scratch, sizeof(int), cudaMemcpyDeviceToHost, s), "flag D2H")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize")) return rc;
  if (*ctx->host_flag) *flags |= CQ_MESH_AXIS_ALIGNED;
  return CQ_OK;
}

int cq_scatter_map_stride(int cell_type, int degree) {
  if ((cell_type != CQ_CELL_QUADRILATERAL && cell_type != CQ_CELL_HEXAHEDRON) || (degree != 1 && degree != 2)) return -1;
  int n = degree + 1;
  n = cell_type == CQ_CELL_QUADRILATERAL ? n * n : n * n * n;
  return (n * n + 15) / 16 * 16;
}

int cq_build_scatter_map(cq_ctx* ctx, int cell_type, int degree, int bs, int64_t num_cells, const int32_t* dofmap,
                         const cq_csr* A, uint8_t* out, void* stream) {
  const int stride = cq_scatter_map_stride(cell_type, degree);
  if (!ctx || !dofmap || !A || !A->indptr || !A->indices || !out || stride < 0 || bs < 1 || bs > 3) { set_error("bad argument"); return CQ_ERR_INVALID; }
  if (num_cells <= 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  int n = degree + 1;
  n = cell_type == CQ_CELL_QUADRILATERAL ? n * n : n * n * n;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = check_cuda(cudaMemsetAsync(out, 0, (size_t)num_cells * stride, s), "memset map")) return rc;
  const int64_t total = num_cells * n * n;
  const int grid = (int)((total + 255) / 256 < (int64_t)ctx->num_sms * 16 ? (total + 255) / 256 : (int64_t)ctx->num_sms * 16);
  int* scratch = ctx->err_flag + 1;  // own word (see cq_mesh_classify)
  if (int rc = check_cuda(cudaMemsetAsync(scratch, 0, sizeof(int), s), "scratch reset")) return rc;
  scatter_map_kernel<<<grid, 256, 0, s>>>(dofmap, num_cells, n, stride, bs, A->indptr, A->indices, out, scratch);
  ctx->launches++;
  if (int rc = check_cuda(cudaGetLastError(), "scatter_map_kernel launch")) return rc;
  // setup call: synchronous check so the caller knows whether the map is usable
  if (int rc = check_cuda(cudaMemcpyAsync(ctx->host_flag, scratch, sizeof(int), cudaMemcpyDeviceToHost, s), "scratch D2H")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize")) return rc;
  if (*ctx->host_flag == 1) { set_error("scatter map: a (row, col) of the dofmap is absent from the CSR pattern"); return CQ_ERR_SPARSITY; }
  if (*ctx->host_flag == 2) { set_error("scatter map: a row holds more than 255 scalar entries"); return CQ_ERR_UNSUPPORTED; }
  return CQ_OK;
}

int cq_csr_expand_blocks(cq_ctx* ctx, int64_t num_rows, const int64_t* indptr, const int32_t* indices, int bs, int64_t* indptr_out,
                         int32_t* indices_out, void* stream) {
  if (!ctx || num_rows < 0 || !indptr || !indptr_out || bs < 1 || bs > 3) { set_error("bad argument"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (num_rows == 0) return check_cuda(cudaMemsetAsync(indptr_out, 0, sizeof(int64_t), s), "memset");
  if (!indices || !indices_out) { set_error("indices null"); return CQ_ERR_INVALID; }
  const int64_t blocks = (num_rows + 7) / 8 < (int64_t)ctx->num_sms * 16 ? (num_rows + 7) / 8 : (int64_t)ctx->num_sms * 16;
  expand_blocks_kernel<<<(int)blocks, 256, 0, s>>>(num_rows, indptr, indices, bs, indptr_out, indices_out);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "expand_blocks_kernel launch");
}

int cq_assemble_ghost_penalty(cq_ctx* ctx, int cell_type, int degree, int bs, double betas, const cq_mesh* mesh, const cq_space* space,
                              int64_t num_facets, const int32_t* facets, const cq_csr* A, void* stream) {
  if (!ctx || !mesh || !space || !space->dofmap || !mesh->geom_dofmap || !mesh->coords || !A || !A->indptr || !A->indices || !A->values || num_facets < 0) {
    set_error("null argument");
    return CQ_ERR_INVALID;
  }
  if ((cell_type != CQ_CELL_QUADRILATERAL && cell_type != CQ_CELL_HEXAHEDRON) || (degree != 1 && degree != 2) || bs < 1 || bs > 3) {
    set_error("ghost penalty: quadrilaterals / hexahedra, degree 1 or 2, bs 1..3");
    return CQ_ERR_UNSUPPORTED;
  }
  if (!(mesh->flags & CQ_MESH_AXIS_ALIGNED)) { set_error("ghost penalty: axis-aligned cells only (cq_mesh.flags)"); return CQ_ERR_UNSUPPORTED; }
  if (num_facets == 0) return CQ_OK;
  if (!facets) { set_error("facets null"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  KParams P;
  memset(&P, 0, sizeof P);
  P.geom_dofmap = mesh->geom_dofmap; P.coords = mesh->coords; P.dofmap = space->dofmap;
  P.indptr = A->indptr; P.indices = A->indices; P.values = A->values; P.err_flag = ctx->err_flag; P.bs = bs;
  int n = degree + 1;
  n = cell_type == CQ_CELL_QUADRILATERAL ? n * n : n * n * n;
  const int64_t total = num_facets * 4 * n * n;
  const int grid = (int)((total + 127) / 128 < (int64_t)ctx->num_sms * 16 ? (total + 127) / 128 : (int64_t)ctx->num_sms * 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool quad = cell_type == CQ_CELL_QUADRILATERAL;
  if (quad && degree == 1) ghost_penalty_kernel<2, 1><<<grid, 128, 0, s>>>(P, num_facets, facets, betas);
  else if (quad) ghost_penalty_kernel<2, 2><<<grid, 128, 0, s>>>(P, num_facets, facets, betas);
  else if (degree == 1) ghost_penalty_kernel<3, 1><<<grid, 128, 0, s>>>(P, num_facets, facets, betas);
  else ghost_penalty_kernel<3, 2><<<grid, 128, 0, s>>>(P, num_facets, facets, betas);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "ghost_penalty_kernel launch");
}

int cq_tabulate(cq_ctx* ctx, int cell_type, int degree, int basix_derivative, int64_t num_points, const double* pts,
                double* out, void* stream) {
  if (!ctx || (num_points > 0 && (!pts || !out))) { set_error("null argument"); return CQ_ERR_INVALID; }
  if ((cell_type != CQ_CELL_QUADRILATERAL && cell_type != CQ_CELL_HEXAHEDRON) || (degree != 1 && degree != 2)) {
    set_error("call_basix: no permutation for cell_type %d degree %d (call_basix.h:60-76)", cell_type, degree);
    return CQ_ERR_UNSUPPORTED;
  }
  const int gdim = gdim_of(cell_type);
  if (basix_derivative < 0 || basix_derivative > gdim) { set_error("basix_derivative %d out of range", basix_derivative); return CQ_ERR_INVALID; }
  if (num_points == 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = 128;
  const int grid = (int)((num_points + threads - 1) / threads < 65535 ? (num_points + threads - 1) / threads : 65535);
  if (gdim == 2 && degree == 1) tabulate_kernel<2, 1><<<grid, threads, 0, s>>>(pts, num_points, basix_derivative, out);
  else if (gdim == 2) tabulate_kernel<2, 2><<<grid, threads, 0, s>>>(pts, num_points, basix_derivative, out);
  else if (degree == 1) tabulate_kernel<3, 1><<<grid, threads, 0, s>>>(pts, num_points, basix_derivative, out);
  else tabulate_kernel<3, 2><<<grid, threads, 0, s>>>(pts, num_points, basix_derivative, out);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "tabulate_kernel launch");
}

int cq_csr_accumulate_rows(cq_ctx* ctx, const cq_csr* A, int64_t row0, int64_t nrows, const int64_t* recv_indptr,
                           const double* recv, void* stream) {
  if (!ctx || !A || !A->indptr || !A->values || (nrows > 0 && (!recv || !recv_indptr))) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (row0 < 0 || nrows < 0 || row0 + nrows > A->num_rows) { set_error("row range out of bounds"); return CQ_ERR_INVALID; }
  if (nrows == 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  const int64_t warps = nrows;  // one warp per row
  const int grid = (int)((warps + 7) / 8 < (int64_t)ctx->num_sms * 8 ? (warps + 7) / 8 : (int64_t)ctx->num_sms * 8);
  accumulate_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A->values, A->indptr, row0, nrows, recv_indptr, recv);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "accumulate_rows_kernel launch");
}

int cq_ipc_export(cq_ctx* ctx, const void* dev_ptr, unsigned char handle[64], int64_t* offset) {
  if (!ctx || !dev_ptr || !handle || !offset) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  if (int rc = check_cuda(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)), "cudaIpcGetMemHandle")) return rc;
  void* base = nullptr;
  size_t size = 0;
  if (cuMemGetAddressRange_fallback(dev_ptr, &base, &size) != 0) { set_error("cannot find the allocation base of the pointer"); return CQ_ERR_CUDA; }
  memcpy(handle, &h, 64);
  *offset = (int64_t)((const char*)dev_ptr - (const char*)base);
  return CQ_OK;
}

int cq_ipc_open(cq_ctx* ctx, const unsigned char handle[64], int64_t offset, void** mapped_ptr, void** mapped_base) {
  if (!ctx || !handle || !mapped_ptr || !mapped_base || offset < 0) { set_error("bad argument"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* base = nullptr;
  if (int rc = check_cuda(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return rc;
  *mapped_base = base;
  *mapped_ptr = (char*)base + offset;
  return CQ_OK;
}

int cq_ipc_close(cq_ctx* ctx, void* mapped_base) {
  if (!ctx || !mapped_base) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  return check_cuda(cudaIpcCloseMemHandle(mapped_base), "cudaIpcCloseMemHandle");
}

static_assert(sizeof(cq_circle) == sizeof(CircleGrid), "cq_circle (C-ABI) and cq::CircleGrid must have the same layout");

int cq_circle_classify(cq_ctx* ctx, const cq_circle* grid, int32_t* cls, int32_t* nbulk, int32_t* nsurf, void* stream) {
  if (!ctx || !grid || !cls || !nbulk || !nsurf) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (grid->n1 < 2 || grid->n1 > LS_MAX_N1 || grid->n[0] < 0 || grid->n[1] < 0 || !(grid->h[0] > 0.0) || !(grid->h[1] > 0.0) || !(grid->radius > 0.0)) {
    set_error("bad cq_circle (n1 in 2..8, positive cell size and radius)");
    return CQ_ERR_INVALID;
  }
  const int64_t ncells = grid->n[0] * grid->n[1];
  if (ncells == 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  const int64_t blocks = (ncells + 255) / 256;
  circle_classify_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*reinterpret_cast<const CircleGrid*>(grid), ncells, cls, nbulk, nsurf);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "circle_classify_kernel launch");
}

int cq_circle_rules(cq_ctx* ctx, const cq_circle* grid, int64_t ncut, const int32_t* cut_cells, const int64_t* off_bulk,
                    const int64_t* off_surf, double* pts, double* w, double* pts_b, double* w_b, double* nrm, void* stream) {
  if (!ctx || !grid || ncut < 0) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (ncut == 0) return CQ_OK;
  if (!cut_cells || !off_bulk || !off_surf || !pts || !w || !pts_b || !w_b || !nrm) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (grid->n1 < 2 || grid->n1 > LS_MAX_N1) { set_error("bad cq_circle.n1"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  const int64_t blocks = (ncut + 127) / 128;
  circle_rules_kernel<<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(*reinterpret_cast<const CircleGrid*>(grid), ncut, cut_cells, off_bulk, off_surf, pts, w,
                                                                                     pts_b, w_b, nrm);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "circle_rules_kernel launch");
}

static_assert(sizeof(cq_sphere) == sizeof(SphereGrid), "cq_sphere (C-ABI) and cq::SphereGrid must have the same layout");

int cq_sphere_classify(cq_ctx* ctx, const cq_sphere* grid, int32_t* cls, int32_t* nbulk, int32_t* nsurf, void* stream) {
  if (!ctx || !grid || !cls || !nbulk || !nsurf) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (grid->n1 < 2 || grid->n1 > LS_MAX_N1 || grid->axis < 0 || grid->axis > 2 || !(grid->radius > 0.0) || !(grid->h[0] > 0.0) || !(grid->h[1] > 0.0) ||
      !(grid->h[2] > 0.0) || grid->n[0] < 0 || grid->n[1] < 0 || grid->n[2] < 0) {
    set_error("bad cq_sphere (n1 in 2..8, axis in 0..2, positive cell size and radius)");
    return CQ_ERR_INVALID;
  }
  const int64_t ncells = grid->n[0] * grid->n[1] * grid->n[2];
  if (ncells == 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  if (ncells >= (int64_t)1 << 31) { set_error("more than 2^31 cells"); return CQ_ERR_UNSUPPORTED; }
  if (ctx->cut_list_cap < (size_t)ncells) {
    if (ctx->cut_list) cudaFree(ctx->cut_list);
    ctx->cut_list = nullptr;
    ctx->cut_list_cap = 0;
    if (int rc = check_cuda(cudaMalloc(&ctx->cut_list, sizeof(int32_t) * (size_t)ncells), "cudaMalloc(cut list)")) return rc;
    ctx->cut_list_cap = (size_t)ncells;
  }
  if (!ctx->cut_count)
    if (int rc = check_cuda(cudaMalloc(&ctx->cut_count, sizeof(unsigned int)), "cudaMalloc(cut count)")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = check_cuda(cudaMemsetAsync(ctx->cut_count, 0, sizeof(unsigned int), s), "memset(cut count)")) return rc;
  const int64_t blocks = (ncells + 255) / 256;
  const SphereGrid& G = *reinterpret_cast<const SphereGrid*>(grid);
  sphere_classify_box_kernel<<<(unsigned)blocks, 256, 0, s>>>(G, ncells, cls, nbulk, nsurf, ctx->cut_list, ctx->cut_count);
  sphere_count_kernel<<<ctx->num_sms * 8, 64, 0, s>>>(G, ctx->cut_list, ctx->cut_count, cls, nbulk, nsurf);
  ctx->launches += 2;
  return check_cuda(cudaGetLastError(), "sphere classification kernels launch");
}

int cq_sphere_rules(cq_ctx* ctx, const cq_sphere* grid, int64_t ncut, const int32_t* cut_cells, const int64_t* off_bulk,
                    const int64_t* off_surf, double* pts, double* w, double* pts_b, double* w_b, double* nrm, void* stream) {
  if (!ctx || !grid || ncut < 0) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (ncut == 0) return CQ_OK;
  if (!cut_cells || !off_bulk || !off_surf || !pts || !w || !pts_b || !w_b || !nrm) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (grid->n1 < 2 || grid->n1 > LS_MAX_N1 || grid->axis < 0 || grid->axis > 2) { set_error("bad cq_sphere"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  const int64_t blocks = (ncut + 63) / 64;
  sphere_rules_kernel<<<(unsigned)blocks, 64, 0, static_cast<cudaStream_t>(stream)>>>(*reinterpret_cast<const SphereGrid*>(grid), ncut, cut_cells, off_bulk, off_surf, pts, w, pts_b,
                                                                                    w_b, nrm);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "sphere_rules_kernel launch");
}

int cq_axpy(cq_ctx* ctx, int64_t n, const double* x, double* y, void* stream) {
  if (!ctx || (n > 0 && (!x || !y))) { set_error("null argument"); return CQ_ERR_INVALID; }
  if (n <= 0) return CQ_OK;
  if (int rc = set_device(ctx)) return rc;
  const int grid = (int)((n + 255) / 256 < (int64_t)ctx->num_sms * 8 ? (n + 255) / 256 : (int64_t)ctx->num_sms * 8);
  accumulate_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, n);
  ctx->launches++;
  return check_cuda(cudaGetLastError(), "accumulate_kernel launch");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// End-to-end entry points: qr data in host memory.  The ragged arrays are streamed to the device
// in chunks of entries on ctx->copy_stream; the kernels for chunk c run on `stream` as soon as its
// copy has landed, overlapping with the copy of chunk c+1.  Because every kernel accumulates into
// the same sink (ADD_VALUES), chunking does not change the result.
// ---------------------------------------------------------------------------------------------
static int ensure_stage(cq_ctx* ctx, size_t bytes) {
  if (ctx->stage_bytes >= bytes) return CQ_OK;
  if (ctx->stage) cudaFree(ctx->stage);
  ctx->stage = nullptr;
  ctx->stage_bytes = 0;
  if (int rc = check_cuda(cudaMalloc(&ctx->stage, bytes), "cudaMalloc(stage)")) return rc;
  ctx->stage_bytes = bytes;
  return CQ_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int no_after_chunk(int, int) { return CQ_OK; }

constexpr int MAX_CHUNKS = 8;
constexpr int MAX_HOST_JOBS = 8;

// What the chunk loop needs to know about a matrix whose finished rows go back early (see cq_assemble_matrices_host)
struct RowsBack {
  int n = 0;            // scalar dofs per cell of the job's space
  int bs = 1;
  const int32_t* dofmap = nullptr;
  const cq_csr* A = nullptr;
  int64_t val_lo[MAX_CHUNKS + 1] = {};  // out: CSR offset of the smallest row any chunk >= c touches
};

// launch(qc): enqueue the kernels of one chunk on `s`; after(c, nchunks): whatever follows chunk c.
// rb[0..nrb): matrices for which val_lo is worked out before the first chunk is launched.
template <class LaunchFn, class AfterFn>
static int run_host_chunked(cq_ctx* ctx, const int cell_type, const cq_quadrature* qh, cudaStream_t s, RowsBack* rb, const int nrb,
                            LaunchFn&& launch, AfterFn&& after) {
  if (!ctx || !qh) { set_error("null argument"); return CQ_ERR_INVALID; }
  const int gdim = gdim_of(cell_type);
  const int64_t C = qh->num_entries, Pn = qh->num_points;
  if (C == 0) return CQ_OK;
  if (!qh->cells || !qh->offsets || (Pn > 0 && (!qh->pts || !qh->weights))) { set_error("qr host arrays null"); return CQ_ERR_INVALID; }
  if (int rc = set_device(ctx)) return rc;
  // device layout inside the staging buffer (256-B aligned sub-arrays)
  const size_t o_cells = 0;
  const size_t o_off = align_up(o_cells + sizeof(int32_t) * C, 256);
  const size_t o_pts = align_up(o_off + sizeof(int64_t) * (C + 1), 256);
  const size_t o_w = align_up(o_pts + sizeof(double) * gdim * Pn, 256);
  const size_t o_n = align_up(o_w + sizeof(double) * Pn, 256);
  const size_t total = align_up(o_n + (qh->normals ? sizeof(double) * gdim * Pn : 0), 256);
  if (int rc = ensure_stage(ctx, total)) return rc;
  char* base = static_cast<char*>(ctx->stage);
  cq_quadrature qd = *qh;
  qd.cells = reinterpret_cast<int32_t*>(base + o_cells);
  qd.offsets = reinterpret_cast<int64_t*>(base + o_off);
  qd.pts = reinterpret_cast<double*>(base + o_pts);
  qd.weights = reinterpret_cast<double*>(base + o_w);
  qd.normals = qh->normals ? reinterpret_cast<double*>(base + o_n) : nullptr;

  // copy stream must not start before prior work on `s` that may still read the staging buffer
  if (int rc = check_cuda(cudaEventRecord(ctx->ev[11], s), "cudaEventRecord")) return rc;
  if (int rc = check_cuda(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[11], 0), "cudaStreamWaitEvent")) return rc;

  const int nchunks = C >= (1 << 16) ? MAX_CHUNKS : 1;
  cudaStream_t cs = ctx->copy_stream;
  bool cells_up = false;
  if (nchunks > 1 && nrb > 0) {
    // all cell ids first (small), then per matrix the smallest row of every chunk and its CSR offset
    int rc = check_cuda(cudaMemcpyAsync(const_cast<int32_t*>(qd.cells), qh->cells, sizeof(int32_t) * C, cudaMemcpyHostToDevice, cs), "H2D cells");
    if (rc) return rc;
    cells_up = true;
    for (int j = 0; j < nrb; ++j) {
      RowsBack& R = rb[j];
      rc = check_cuda(cudaMemsetAsync(ctx->chunk_rowmin, 0x7f, 16 * sizeof(int32_t), cs), "rowmin init");
      if (rc) return rc;
      chunk_rowmin_kernel<<<ctx->num_sms * 4, 256, 0, cs>>>(qd.cells, C, nchunks, R.dofmap, R.n, R.bs, ctx->chunk_rowmin);
      ctx->launches++;
      rc = check_cuda(cudaGetLastError(), "chunk_rowmin_kernel launch");
      if (!rc) rc = check_cuda(cudaMemcpyAsync(ctx->host_rowmin, ctx->chunk_rowmin, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, cs), "rowmin D2H");
      if (!rc) rc = check_cuda(cudaStreamSynchronize(cs), "cudaStreamSynchronize");
      if (rc) return rc;
      int32_t suffix = 0x7fffffff;
      for (int c = nchunks - 1; c >= 0; --c) {
        if (ctx->host_rowmin[c] < suffix) suffix = ctx->host_rowmin[c];
        ctx->host_rowmin[c] = suffix;
      }
      for (int c = 0; c < nchunks && !rc; ++c) {
        const int64_t row = ctx->host_rowmin[c] < R.A->num_rows ? ctx->host_rowmin[c] : R.A->num_rows;
        rc = check_cuda(cudaMemcpyAsync(ctx->host_vallo + c, R.A->indptr + row, sizeof(int64_t), cudaMemcpyDeviceToHost, cs), "indptr D2H");
      }
      if (!rc) rc = check_cuda(cudaStreamSynchronize(cs), "cudaStreamSynchronize");
      if (rc) return rc;
      for (int c = 0; c < nchunks; ++c) R.val_lo[c] = ctx->host_vallo[c];
      R.val_lo[nchunks] = R.A->nnz;
    }
  }
  for (int c = 0; c < nchunks; ++c) {
    const int64_t k0 = C * c / nchunks, k1 = C * (c + 1) / nchunks;
    if (k1 == k0) continue;
    const int64_t p0 = qh->offsets[k0], p1 = qh->offsets[k1];
    int rc = CQ_OK;
    if (!cells_up) rc = check_cuda(cudaMemcpyAsync(const_cast<int32_t*>(qd.cells) + k0, qh->cells + k0, sizeof(int32_t) * (k1 - k0), cudaMemcpyHostToDevice, cs), "H2D cells");
    if (!rc) rc = check_cuda(cudaMemcpyAsync(const_cast<int64_t*>(qd.offsets) + k0, qh->offsets + k0, sizeof(int64_t) * (k1 - k0 + 1), cudaMemcpyHostToDevice, cs), "H2D offsets");
    if (!rc && p1 > p0) {
      rc = check_cuda(cudaMemcpyAsync(const_cast<double*>(qd.pts) + gdim * p0, qh->pts + gdim * p0, sizeof(double) * gdim * (p1 - p0), cudaMemcpyHostToDevice, cs), "H2D pts");
      if (!rc) rc = check_cuda(cudaMemcpyAsync(const_cast<double*>(qd.weights) + p0, qh->weights + p0, sizeof(double) * (p1 - p0), cudaMemcpyHostToDevice, cs), "H2D weights");
      if (!rc && qh->normals)
        rc = check_cuda(cudaMemcpyAsync(const_cast<double*>(qd.normals) + gdim * p0, qh->normals + gdim * p0, sizeof(double) * gdim * (p1 - p0), cudaMemcpyHostToDevice, cs), "H2D normals");
    }
    if (!rc) rc = check_cuda(cudaEventRecord(ctx->ev[c], cs), "cudaEventRecord");
    if (!rc) rc = check_cuda(cudaStreamWaitEvent(s, ctx->ev[c], 0), "cudaStreamWaitEvent");
    if (rc) return rc;
    cq_quadrature qc = qd;
    qc.plane_entries = 0;  // a chunk is a sub-list: no traversal hint
    qc.plane_tiles = 0;
    qc.cells = qd.cells + k0;
    qc.offsets = qd.offsets + k0;
    qc.num_entries = k1 - k0;
    if (int lrc = launch(qc)) return lrc;
    if (int arc = after(c, nchunks)) return arc;
  }
  return CQ_OK;
}

extern "C" {

int cq_assemble_matrices_host(cq_ctx* ctx, int njobs, const cq_matrix_job* jobs, const cq_mesh* mesh, const cq_quadrature* qr_host,
                              void* stream) {
  if (!ctx || !jobs || !mesh || !qr_host || njobs < 1 || njobs > MAX_HOST_JOBS) { set_error("bad argument (1..%d jobs)", MAX_HOST_JOBS); return CQ_ERR_INVALID; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = set_device(ctx)) return rc;
  RowsBack rb[MAX_HOST_JOBS];
  int rb_of[MAX_HOST_JOBS];
  int64_t copied[MAX_HOST_JOBS] = {};
  int nrb = 0;
  for (int j = 0; j < njobs; ++j) {
    const cq_matrix_job& J = jobs[j];
    if (!J.form || !J.space || !J.A || !J.A->values || !J.A->indptr) { set_error("job %d: null form / space / CSR", j); return CQ_ERR_INVALID; }
    if (J.form->cell_type != jobs[0].form->cell_type) { set_error("jobs on different cell types"); return CQ_ERR_INVALID; }
    if (J.A->flags & CQ_CSR_PEER_TARGET) { set_error("job %d: the matrix is the target of a neighbour's peer scatter (cq_csr.flags)", j); return CQ_ERR_UNSUPPORTED; }
    if (int rc = check_cuda(cudaMemsetAsync(J.A->values, 0, sizeof(double) * (size_t)J.A->nnz, s), "zeroEntries")) return rc;  // A.zeroEntries()
    rb_of[j] = -1;
    if (J.values_host && J.space->dofmap) {
      int n = J.form->degree + 1;
      n = gdim_of(J.form->cell_type) == 2 ? n * n : n * n * n;
      rb[nrb].n = n;
      rb[nrb].bs = J.form->bs;
      rb[nrb].dofmap = J.space->dofmap;
      rb[nrb].A = J.A;
      rb_of[j] = nrb++;
    }
  }
  // Rows are final once no later chunk touches them: values below val_lo[c + 1] (the CSR offset of the smallest row
  // any chunk after c touches) go back to the host on a second stream while the points of the later chunks are still
  // coming in (PCIe is full duplex).  For a sorted cell list with a locality-preserving dof numbering that hides all
  // but the last chunk's rows; for an arbitrary list val_lo is 0 and everything is copied at the end.
  auto after_chunk = [&](int c, int nchunks) -> int {
    if (c + 1 >= nchunks) return CQ_OK;
    bool recorded = false;
    for (int j = 0; j < njobs; ++j) {
      if (rb_of[j] < 0) continue;
      const int64_t upto = rb[rb_of[j]].val_lo[c + 1];
      if (upto <= copied[j]) continue;
      int rc = CQ_OK;
      if (!recorded) {
        rc = check_cuda(cudaEventRecord(ctx->kev[c], s), "cudaEventRecord");
        if (!rc) rc = check_cuda(cudaStreamWaitEvent(ctx->d2h_stream, ctx->kev[c], 0), "cudaStreamWaitEvent");
        recorded = true;
      }
      if (!rc) rc = check_cuda(cudaMemcpyAsync(jobs[j].values_host + copied[j], jobs[j].A->values + copied[j], sizeof(double) * (upto - copied[j]), cudaMemcpyDeviceToHost, ctx->d2h_stream), "D2H rows");
      if (rc) return rc;
      copied[j] = upto;
    }
    return CQ_OK;
  };
  auto launch = [&](const cq_quadrature& qc) -> int {
    for (int j = 0; j < njobs; ++j)
      if (int rc = cq_assemble_matrix(ctx, jobs[j].form, mesh, jobs[j].space, &qc, jobs[j].A, stream)) return rc;
    return CQ_OK;
  };
  if (int rc = run_host_chunked(ctx, jobs[0].form->cell_type, qr_host, s, rb, nrb, launch, after_chunk)) return rc;
  // the rest: after the last kernel, on the result stream (keeps the order of the host buffer writes simple)
  if (int rc = check_cuda(cudaEventRecord(ctx->kev[15], s), "cudaEventRecord")) return rc;
  if (int rc = check_cuda(cudaStreamWaitEvent(ctx->d2h_stream, ctx->kev[15], 0), "cudaStreamWaitEvent")) return rc;
  for (int j = 0; j < njobs; ++j)
    if (jobs[j].values_host && jobs[j].A->nnz > copied[j])
      if (int rc = check_cuda(cudaMemcpyAsync(jobs[j].values_host + copied[j], jobs[j].A->values + copied[j], sizeof(double) * (jobs[j].A->nnz - copied[j]), cudaMemcpyDeviceToHost, ctx->d2h_stream), "D2H values")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize")) return rc;
  return check_cuda(cudaStreamSynchronize(ctx->d2h_stream), "cudaStreamSynchronize(d2h)");
}

int cq_assemble_matrix_host(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                            const cq_quadrature* qr_host, const cq_csr* A, double* values_host, void* stream) {
  const cq_matrix_job job = {form, space, A, values_host};
  return cq_assemble_matrices_host(ctx, 1, &job, mesh, qr_host, stream);
}

int cq_assemble_vector_host(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                            const cq_quadrature* qr_host, double* b_dev, double* b_host, void* stream) {
  if (!ctx || !form || !b_dev || !space) { set_error("null argument"); return CQ_ERR_INVALID; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = set_device(ctx)) return rc;
  const size_t bytes = sizeof(double) * space->num_dofs * form->bs;
  if (int rc = check_cuda(cudaMemsetAsync(b_dev, 0, bytes, s), "zero b")) return rc;
  int rc = run_host_chunked(ctx, form->cell_type, qr_host, s, nullptr, 0,
                            [&](const cq_quadrature& qc) { return cq_assemble_vector(ctx, form, mesh, space, &qc, b_dev, stream); }, no_after_chunk);
  if (rc) return rc;
  if (b_host)
    if (int rc2 = check_cuda(cudaMemcpyAsync(b_host, b_dev, bytes, cudaMemcpyDeviceToHost, s), "D2H b")) return rc2;
  return check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
}

int cq_assemble_scalar_host(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh, const cq_space* space,
                            const cq_quadrature* qr_host, double* m_host, void* stream) {
  if (!ctx || !form || !m_host) { set_error("null argument"); return CQ_ERR_INVALID; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = set_device(ctx)) return rc;
  if (!ctx->scalar_dev)
    if (int rc = check_cuda(cudaMalloc(&ctx->scalar_dev, sizeof(double)), "cudaMalloc(m)")) return rc;
  double* m_dev = ctx->scalar_dev;
  if (int rc = check_cuda(cudaMemsetAsync(m_dev, 0, sizeof(double), s), "zero m")) return rc;
  int rc = run_host_chunked(ctx, form->cell_type, qr_host, s, nullptr, 0,
                            [&](const cq_quadrature& qc) { return cq_assemble_scalar(ctx, form, mesh, space, &qc, m_dev, stream); }, no_after_chunk);
  if (rc) return rc;
  if (int rc2 = check_cuda(cudaMemcpyAsync(m_host, m_dev, sizeof(double), cudaMemcpyDeviceToHost, s), "D2H m")) return rc2;
  return check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
}

}  // extern "C"
