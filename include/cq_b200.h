/*
 * cq_b200.h -- C-ABI of libcq_b200.so: runtime-quadrature finite-element assembly
 * (customquad's assemble_matrix / assemble_vector / assemble_scalar) on one B200.
 *
 * Boundary rules:
 *   - extern "C" only, plain pointers and sizes, no torch / C++ types.
 *   - every function returns int: 0 = CQ_OK, negative = cq_status; cq_last_error()
 *     gives a thread-local message.  Nothing throws or exits across the boundary.
 *   - all array arguments of the cq_assemble_* entry points are caller-owned DEVICE
 *     pointers on the device the ctx was created for (the *_host entry points take
 *     HOST pointers and do their own staged copies); the library allocates nothing
 *     persistent except the opaque cq_ctx (scratch, error flag, copy streams).
 *   - calls are asynchronous and ordered on the cudaStream_t passed in (as void*).
 *
 * What each entry point replaces in the reference (paths into /root/reference):
 *   cq_assemble_matrix  <- customquad/assemble_matrix.py:50-113 (assemble_cells: gather,
 *                          generated tabulate_tensor_runtime_float64, MatSetValuesLocal ADD)
 *                          + customquad/call_basix.h:14-81 (basis tables at runtime points)
 *                          + customquad/setup_types.py:106-113 (MatSetValuesLocal binding)
 *   cq_assemble_vector  <- customquad/assemble_vector.py:52-94
 *   cq_assemble_scalar  <- customquad/assemble_scalar.py:37-82
 *   cq_tabulate         <- customquad/call_basix.h:14-81 alone (debug / parity hook; the
 *                          analogue of its debug_output dump, call_basix.h:83-126)
 *   cq_element_tensors  <- one generated-kernel call per cell, README.md:133-158, without
 *                          the scatter (the analogue of utils.print_for_header, utils.py:185-225)
 *   cq_csr_accumulate_rows <- the receive side of PETSc MatAssemblyBegin/End after
 *                          A.assemble() (README.md:95, demo/poisson.py:194)
 *   cq_exchange_offrank <- that whole stash exchange (send the rows a neighbour owns, receive and
 *                          add the neighbour's contributions to the rows this rank owns), NCCL send/recv
 *   cq_assemble_matrices_host <- several assemble_matrix calls on ONE qr tuple (demo/poisson.py:193,209
 *                          reuse qr_bulk for a matrix and a vector): the rule crosses PCIe once
 */
#ifndef CQ_B200_H
#define CQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CQ_B200_ABI_VERSION 5

/* ---- status codes ------------------------------------------------------------------ */
typedef enum cq_status {
  CQ_OK = 0,
  CQ_ERR_INVALID = -1,     /* bad argument (null pointer, unsupported cell/degree/form ...) */
  CQ_ERR_CUDA = -2,        /* a CUDA runtime call failed; see cq_last_error()               */
  CQ_ERR_UNSUPPORTED = -3, /* combination not implemented                                   */
  CQ_ERR_SPARSITY = -4,    /* a (row, col) produced by a cell is absent from the CSR pattern */
  CQ_ERR_ALIGN = -5        /* pointer alignment requirement violated (8 B for all f64/i64)   */
} cq_status;

/* ---- what is integrated --------------------------------------------------------------
 * The reference gets its integrand from UFL through ffcx-custom; there is no UFL here, so
 * the integrand is one of the forms the reference's demos and tests use. */
typedef enum cq_form {
  /* rank 2 -> cq_assemble_matrix */
  CQ_FORM_MASS = 0,        /* inner(u, v)                         test/common.py:91            */
  CQ_FORM_STIFFNESS = 1,   /* inner(grad u, grad v)               demo/poisson.py:154          */
  CQ_FORM_NITSCHE = 2,     /* -(n.grad u) v - u (n.grad v) + betaN/h u v  demo/poisson.py:156-158 */
  /* rank 1 -> cq_assemble_vector */
  CQ_FORM_LOAD = 10,       /* inner(f, v)      demo/poisson.py:155, test/common.py:68, test_linear_basis.py:129 */
  CQ_FORM_NITSCHE_RHS = 11,/* -g (n.grad v) + betaN/h g v         demo/poisson.py:159          */
  /* rank 0 -> cq_assemble_scalar */
  CQ_FORM_INTEGRAL = 20,   /* f                test_volume_area.py:58, test_linear_basis.py:84-86 */
  CQ_FORM_INTEGRAL_SQ = 21,/* inner(f, f)      test/common.py:44                               */
  CQ_FORM_L2_ERR = 22,     /* (uh - u_ex)**2   demo/poisson.py:275                             */
  CQ_FORM_H10_ERR = 23     /* (grad uh - grad u_ex)**2   demo/poisson.py:279                   */
} cq_form;

/* The data function f / g / u_ex of a form.  CQ_FN_COEF is a finite-element function of the
 * form's own space (packed per cell like dolfinx.cpp.fem.pack_coefficients,
 * assemble_matrix.py:16,92); the others are closed-form in the physical coordinate x
 * (ufl.SpatialCoordinate in the reference) with parameters in `consts`. Component a of a
 * blocked (bs > 1) function uses its own parameters. */
typedef enum cq_fn {
  CQ_FN_COEF = 0,    /* sum_k w[k*bs + a] psi_k(xi)                                             */
  CQ_FN_CONST = 1,   /* p[a]                                              (1 parameter / comp.) */
  CQ_FN_AFFINE = 2,  /* p[4a] + p[4a+1] x + p[4a+2] y + p[4a+3] z         (4 parameters / comp.) */
  CQ_FN_SINPROD = 3  /* p[a] * prod_d sin(pi x_d)   demo/poisson.py:59-67 (1 parameter / comp.) */
} cq_fn;

/* basix::cell::type values used by the reference (call_basix.h:61,68; README.md:153) */
#define CQ_CELL_QUADRILATERAL 4
#define CQ_CELL_HEXAHEDRON 5

#define CQ_MAX_CONSTS 16

/* Form descriptor.  consts is a HOST pointer (copied into the launch parameters):
 *   Nitsche forms: consts[0] = betaN, function parameters p = consts + 1;
 *   all other forms: p = consts.
 * For L2_ERR / H10_ERR the coefficient is uh (always CQ_FN_COEF) and fn_kind describes u_ex. */
typedef struct cq_form_desc {
  int32_t form;       /* cq_form */
  int32_t cell_type;  /* CQ_CELL_QUADRILATERAL | CQ_CELL_HEXAHEDRON */
  int32_t degree;     /* 1 | 2 : Lagrange degree of the form's space (call_basix.h:63-73) */
  int32_t bs;         /* block size of the space: 1, or gdim for VectorFunctionSpace (utils.py:47-62) */
  int32_t fn_kind;    /* cq_fn */
  int32_t nconsts;    /* <= CQ_MAX_CONSTS */
  const double* consts;
} cq_form_desc;

/* Mesh geometry exactly as utils.get_vertices returns it (utils.py:66-71):
 *   geom_dofmap [num_cells, nv] int32 row-major, nv = 4 (quad) | 8 (hex), dolfinx local order
 *   coords      [num_nodes, 3]  f64 row-major (z = 0 in 2-D)                                  */
typedef struct cq_mesh {
  int64_t num_cells;
  int64_t num_nodes;
  const int32_t* geom_dofmap;
  const double* coords;
  /* Optional hints from cq_mesh_classify (0 = unknown, always safe):
   *   CQ_MESH_AXIS_ALIGNED: every cell has a constant diagonal Jacobian (what create_rectangle /
   *   create_box produce); enables the moment kernels for mass / stiffness on quadrilaterals
   *   (degree 1, 2) and hexahedra (degree 1). */
  int32_t flags;
} cq_mesh;

#define CQ_MESH_AXIS_ALIGNED 1

/* Function-space data as utils.get_dofs returns it (utils.py:33-63), kept un-expanded:
 *   dofmap [num_cells, n] int32, n = (degree+1)^gdim scalar dofs per cell; the blocked index
 *   of component a of local dof i is dofmap[c, i]*bs + a (utils.py:59-62).
 *   coeffs [num_cells, n*bs] f64, packed coefficient of the form, indexed by MESH cell id
 *   (assemble_matrix.py:92), or NULL when the form has none.                                  */
typedef struct cq_space {
  int64_t num_dofs;        /* scalar dofs; vector length = num_dofs * bs */
  const int32_t* dofmap;
  const double* coeffs;
  /* Optional (may be NULL): scatter positions built once per (space, pattern) by
   * cq_build_scatter_map -- the stand-in for the local-to-global map a PETSc Mat carries.
   * With it cq_assemble_matrix runs the bulk-copy staged kernel and scatters without searching
   * CSR rows; without it every (row, col) is found by binary search like MatSetValuesLocal. */
  const uint8_t* elem_pos;
} cq_space;

/* One qr_data tuple (cells, qr_pts, qr_w[, qr_n]) (assemble_matrix.py:65-73) in ragged
 * CSR form: entry k of the list integrates mesh cell cells[k] with the points
 * offsets[k] .. offsets[k+1]-1.
 *   pts     [num_points, gdim] f64, point-major interleaved reference coordinates (call_basix.h:38-39)
 *   weights [num_points]       f64, reference-measure weights (kernel multiplies by |det J|)
 *   normals [num_points, gdim] f64 physical normals, or NULL for a 3-tuple                   */
typedef struct cq_quadrature {
  int64_t num_entries;
  int64_t num_points;
  const int32_t* cells;
  const int64_t* offsets;  /* [num_entries + 1], offsets[0] = 0 */
  const double* pts;
  const double* weights;
  const double* normals;
  /* Optional traversal hint (0 = none; never changes the result beyond FP64 summation order).  When the list walks a
   * structured 3-D mesh plane by plane (plane_entries consecutive entries per cell layer, entry 0 at the start of a
   * layer), a matrix whose dof planes are larger than the L2 is scattered into twice per plane from DRAM: once per
   * adjacent cell layer.  With plane_tiles > 1 the persistent kernels walk the list tile by tile instead -- for each
   * of plane_tiles groups of rows, all layers -- so that the rows a tile scatters into stay in L2 between two
   * consecutive layers (config 5: 512^2-cell layers, 57 MB of matrix rows per dof plane). */
  int64_t plane_entries;
  int32_t plane_tiles;
  int32_t pad_;
} cq_quadrature;

/* CSR matrix with the layout dolfinx.cpp.fem.petsc.create_matrix gives an AIJ Mat
 * (assemble_matrix.py:19): rows = num_dofs*bs, columns ascending inside each row.            */
typedef struct cq_csr {
  int64_t num_rows;
  int64_t nnz;
  const int64_t* indptr;   /* [num_rows + 1] */
  const int32_t* indices;  /* [nnz] */
  double* values;          /* [nnz] -- accumulated into (ADD_VALUES); caller zeroes */
  /* Optional (multi-GPU, cell-partitioned assembly): the rows >= peer_row0 of this matrix are owned by a
   * neighbour rank.  With peer_rowstart != NULL the kernels add the contributions to those rows straight
   * into the owner's value array over NVLink (red.global.add.f64 on peer memory) instead of into `values`:
   * the slot of entry p of row r >= peer_row0 is values + peer_rowstart[r - peer_row0] + p, where
   * peer_rowstart (device, [num_rows - peer_row0 + 1] int64) already contains the distance between the
   * two allocations in doubles (customquad_b200/dist.py builds it from a cq_ipc_open'ed pointer).  This
   * replaces the separate stash exchange of PETSc's MatAssemblyBegin/End (README.md:95).  Set
   * peer_row0 = num_rows and peer_rowstart = NULL when unused. */
  int64_t peer_row0;
  const int64_t* peer_rowstart;
  /* CQ_CSR_PEER_TARGET: a neighbour rank's kernels add into `values` over NVLink (this rank is the owner side of
   * a peer scatter).  Zeroing such a matrix has to be ordered against the neighbour by the caller: zero the slots
   * the neighbour adds into, fence, assemble, fence.  cq_assemble_matrix_zero accepts it only with a non-empty
   * zero_keep range (below); the *_host entry points refuse it with CQ_ERR_UNSUPPORTED. */
  int32_t flags;
  int32_t pad_;
  /* cq_assemble_matrix_zero leaves values[zero_keep_lo, zero_keep_hi) alone and zeroes the rest (ABI 5).  For the
   * owner side of a peer scatter: the slots of the interface rows, which the caller zeroed BEFORE the fence that
   * precedes this call -- the neighbour's kernels may already have added into them, while every other slot of the
   * matrix is touched by this rank's kernels only and can be zeroed by the call (memset, or inside the kernel).
   * Empty (lo >= hi, e.g. both 0) when unused. */
  int64_t zero_keep_lo;
  int64_t zero_keep_hi;
} cq_csr;

#define CQ_CSR_PEER_TARGET 1

typedef struct cq_ctx cq_ctx;

/* ---- context ------------------------------------------------------------------------ */
int cq_abi_version(void);
const char* cq_last_error(void);
int cq_ctx_create(int device, cq_ctx** out);
int cq_ctx_destroy(cq_ctx* ctx);
/* Synchronises `stream` and returns CQ_ERR_SPARSITY if any launch since the last check hit a
 * (row, col) missing from the pattern (the reference drops PETSc's return code,
 * assemble_matrix.py:103; here it is reported). Resets the flag. */
int cq_ctx_check(cq_ctx* ctx, void* stream);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
int64_t cq_ctx_launch_count(const cq_ctx* ctx);
/* Kernel variant: 0 = auto, 1 = direct-load kernel only, 2 = bulk-copy (TMA) staged persistent
 * kernel only (error when not applicable), 3 = like 0 but never the moment kernels. */
int cq_ctx_set_variant(cq_ctx* ctx, int variant);
/* cq_assemble_matrix_zero zeroes inside the persistent moment kernels (cooperative launch) instead of a memset
 * before them: 1 wherever a kernel can (mass / stiffness on quadrilaterals and Q1 hexahedra), 0 never, -1 = the
 * default: where it was measured to pay (the Q2 quadrilateral kernels: zero + assemble 1.78 -> 1.54 ms (mass) and
 * 1.83 -> 1.77 ms (stiffness) on the bench workload; the Q1 kernels are bound by their load/store pipe and lose).
 * Same result either way. */
int cq_ctx_set_fused_zero(cq_ctx* ctx, int on);

/* ---- the hot path (device pointers) ---------------------------------------------------- */
int cq_assemble_matrix(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                       const cq_space* space, const cq_quadrature* qr, const cq_csr* A,
                       void* stream);
/* A.zeroEntries() and cq_assemble_matrix in one pass (customquad/assemble_matrix.py:20 followed by
 * :26-45 for the first integral): A->values need not be initialised.  On the persistent moment
 * kernels the zeroing runs inside the kernel, a few stages ahead of the scatter, so the zeroed lines
 * are still in L2 when the element tensors arrive and each line of the matrix crosses the HBM
 * interface once instead of three times (memset write, read-modify-write) -- when the ctx has the
 * fused zeroing applies (cq_ctx_set_fused_zero; by default on the Q2 quadrilateral kernels only: it
 * moves 7.0 instead of 9.2 GB per Q2 matrix, but the Q1 kernels are no faster for it,
 * profiles/r1/fused_zero_study.md, profiles/r2s3/).  Otherwise, for every other kernel, for lists long
 * enough for ticketed stages (hexahedra), and for a matrix with peer rows, a memset runs first.  Same result as memset + cq_assemble_matrix. */
int cq_assemble_matrix_zero(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                            const cq_space* space, const cq_quadrature* qr, const cq_csr* A,
                            void* stream);
/* b [num_dofs*bs] f64, accumulated into: b[dofs[cell, j]] += b_local[j] (assemble_vector.py:93-94) */
int cq_assemble_vector(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                       const cq_space* space, const cq_quadrature* qr, double* b, void* stream);
/* m [1] f64 on the device, accumulated into: m[0] += m_local[0] (assemble_scalar.py:82).
 * space may be NULL when the form has no coefficient. */
int cq_assemble_scalar(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                       const cq_space* space, const cq_quadrature* qr, double* m, void* stream);

/* ---- mesh classification (setup, not the timed path) ------------------------------------------
 * Sets *flags to the cq_mesh.flags bits that hold for every cell of the mesh. Synchronous. */
int cq_mesh_classify(cq_ctx* ctx, int cell_type, const cq_mesh* mesh, int32_t* flags, void* stream);

/* ---- scatter map (setup, not the timed path) -----------------------------------------------
 * out [num_cells, cq_scatter_map_stride()] uint8: out[c*stride + i*n + j] = position of column
 * dofmap[c,j] inside the sorted SCALAR row dofmap[c,i] (for bs > 1 the blocked position of
 * component (a,a) is pos*bs + a in blocked row dofmap[c,i]*bs + a).  Rows longer than 255 scalar
 * entries cannot be mapped: CQ_ERR_UNSUPPORTED (callers then leave elem_pos NULL).
 * A is the pattern the matrices will use (values may be NULL). */
int cq_scatter_map_stride(int cell_type, int degree);
int cq_build_scatter_map(cq_ctx* ctx, int cell_type, int degree, int bs, int64_t num_cells,
                         const int32_t* dofmap, const cq_csr* A, uint8_t* out, void* stream);

/* ---- blocked pattern (setup, not the timed path) -------------------------------------------------
 * The AIJ pattern of a blocked space (VectorFunctionSpace, bs = gdim; demo/vector_poisson.py:95) from the pattern of
 * its scalar space: blocked row r*bs + a holds the columns c*bs + b, b < bs, of every scalar coupling (r, c)
 * (utils.py:59-62 interleaving; what dolfinx's create_matrix gives a PETSc AIJ Mat with block size bs).
 *   indptr_out [num_rows*bs + 1] int64, indices_out [indptr[num_rows]*bs*bs] int32 (device, caller-allocated).
 * One pass, no temporaries: config 4 (Q2 hexahedra 128^3, bs = 3) expands 1.08e9 scalar couplings into 9.69e9. */
int cq_csr_expand_blocks(cq_ctx* ctx, int64_t num_rows, const int64_t* indptr, const int32_t* indices, int bs,
                         int64_t* indptr_out, int32_t* indices_out, void* stream);

/* ---- ghost penalty (the step beside the path) ---------------------------------------
 * betas * avg(h) * inner(jump(n, grad u), jump(n, grad v)) * dS(ghost_penalty_tag) (demo/poisson.py:160,168-170;
 * demo/vector_poisson.py:127 per component), which the reference leaves to dolfinx's standard assembler, on the
 * interior facets utils.get_facetags tags (utils.py:155-182), added into the same device CSR as the runtime-
 * quadrature terms so that A = Ax + Ac1 + Ac2 never leaves the GPU.  Axis-aligned cells (cq_mesh.flags must carry
 * CQ_MESH_AXIS_ALIGNED), Lagrange degree 1 or 2, bs >= 1 (block-diagonal).
 *   facets [num_facets, 3] int32 on the device: (cell on the low side, cell on the high side, axis of the normal).
 * The pattern of A has to contain the couplings between the dofs of the two cells of every facet
 * (customquad_b200/sparsity.py:add_macro_couplings); a missing slot sets the sparsity error flag (cq_ctx_check). */
int cq_assemble_ghost_penalty(cq_ctx* ctx, int cell_type, int degree, int bs, double betas, const cq_mesh* mesh,
                              const cq_space* space, int64_t num_facets, const int32_t* facets, const cq_csr* A,
                              void* stream);

/* ---- parity / debug hooks --------------------------------------------------------------- */
/* call_basix for every point of `qr`: out [num_points, nbf] = D^alpha psi_j(xi_q) with
 * basix_derivative 0 value, 1 d/dxi0, 2 d/dxi1, 3 d/dxi2 (call_basix.h:34-39,79-81). */
int cq_tabulate(cq_ctx* ctx, int cell_type, int degree, int basix_derivative, int64_t num_points,
                const double* pts, double* out, void* stream);
/* Un-scattered element tensors: out [num_entries, len], len = (n*bs)^2 | n*bs | 1. */
int cq_element_tensors(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                       const cq_space* space, const cq_quadrature* qr, double* out, void* stream);

/* ---- multi-GPU receive side --------------------------------------------------------------- */
/* For r in [0, nrows): values[indptr[row0 + r] + i] += recv[recv_indptr[r] + i],
 * i < recv_indptr[r+1] - recv_indptr[r].  Adds a neighbour rank's contributions to rows this
 * rank owns; a sender row is a prefix of the owner's row (customquad_b200/dist.py).
 * recv_indptr [nrows + 1] int64 and recv are device pointers. */
int cq_csr_accumulate_rows(cq_ctx* ctx, const cq_csr* A, int64_t row0, int64_t nrows,
                           const int64_t* recv_indptr, const double* recv, void* stream);
/* y[i] += x[i], i < n (vector interface dofs) */
int cq_axpy(cq_ctx* ctx, int64_t n, const double* x, double* y, void* stream);

/* ---- peer memory (one process per GPU) --------------------------------------------------------
 * cq_ipc_export: the 64-byte CUDA IPC handle of the allocation containing dev_ptr and dev_ptr's byte
 * offset inside it; cq_ipc_open (in the neighbour's process, on its own ctx): map that allocation and
 * return the address of the same byte there (peer access is enabled lazily); cq_ipc_close unmaps. */
int cq_ipc_export(cq_ctx* ctx, const void* dev_ptr, unsigned char handle[64], int64_t* offset);
int cq_ipc_open(cq_ctx* ctx, const unsigned char handle[64], int64_t offset, void** mapped_ptr, void** mapped_base);
int cq_ipc_close(cq_ctx* ctx, void* mapped_base);

/* ---- cut-cell quadrature producer (the step in front of the path) --------
 * What the reference's demo gets from Algoim: demo/algoim_utils.hpp:84-153 with the level set of
 * demo/circle.hpp:14-34, phi = |x - center|^2 - radius^2, on a tensor quadrilateral mesh with cell
 * id = ix * n[1] + iy (dolfinx create_rectangle order).  Own scheme (Algoim is not available):
 * customquad_b200/csrc/cq_levelset.cuh.  g / gw: Gauss-Legendre nodes / weights on [0, 1].        */
typedef struct cq_circle {
  double xmin[2];
  double h[2];
  int64_t n[2];
  double center[2];
  double radius;
  int32_t n1;              /* Gauss points per direction, 2..8 */
  int32_t pad_;
  double g[8];
  double gw[8];
} cq_circle;
/* Pass 1, one thread per cell: cls[c] = 0 outside / 1 cut / 2 uncut (algoim_utils.hpp:98-100,
 * 145-152); nbulk[c] / nsurf[c] = points of the bulk / surface rule of a cut cell (0 otherwise).
 * All three [n[0]*n[1]] int32 on the device. */
int cq_circle_classify(cq_ctx* ctx, const cq_circle* grid, int32_t* cls, int32_t* nbulk,
                       int32_t* nsurf, void* stream);
/* Pass 2, one thread per cut cell k (cell cut_cells[k]): the bulk rule into pts[2*off_bulk[k]..],
 * w[off_bulk[k]..] and the surface rule into pts_b / w_b / nrm at off_surf[k]: points in unit-cell
 * coordinates (algoim_utils.hpp:134-143), weights relative to the cell volume (:127-132), unit
 * outward normals -- the qr_pts / qr_w / qr_n arrays of a qr_data tuple in ragged form, device
 * resident, ready for cq_assemble_*. */
int cq_circle_rules(cq_ctx* ctx, const cq_circle* grid, int64_t ncut, const int32_t* cut_cells,
                    const int64_t* off_bulk, const int64_t* off_surf, double* pts, double* w,
                    double* pts_b, double* w_b, double* nrm, void* stream);

/* The same for a ball on a tensor hexahedral mesh (demo/sphere.hpp:14-34; cell id = (iz*n[1] + iy)*n[0] + ix,
 * dolfinx create_box order): the ball is sliced along `axis`, the cross-sections are disk-rectangle problems.
 * gu / gwu: Gauss-Legendre on [0, 1] with n1 + 2 points (slicing variable). */
typedef struct cq_sphere {
  double xmin[3];
  double h[3];
  int64_t n[3];
  double center[3];
  double radius;
  int32_t n1;
  int32_t axis;
  double g[8];
  double gw[8];
  double gu[10];
  double gwu[10];
} cq_sphere;
int cq_sphere_classify(cq_ctx* ctx, const cq_sphere* grid, int32_t* cls, int32_t* nbulk,
                       int32_t* nsurf, void* stream);
/* points / normals: 3 doubles per point */
int cq_sphere_rules(cq_ctx* ctx, const cq_sphere* grid, int64_t ncut, const int32_t* cut_cells,
                    const int64_t* off_bulk, const int64_t* off_surf, double* pts, double* w,
                    double* pts_b, double* w_b, double* nrm, void* stream);

/* ---- end-to-end entry points (HOST pointers; copies are part of the call) ------------------
 * The reference-facing call with host buffers: qr arrays, and the result, live in (ideally
 * pinned) host memory; mesh/space/pattern are device-resident setup data like the PETSc Mat
 * and dolfinx mesh are in the reference.  Points are streamed host->device in chunks that
 * overlap with the kernel; the result is copied back before the call returns.              */
int cq_assemble_matrix_host(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                            const cq_space* space, const cq_quadrature* qr_host, const cq_csr* A,
                            double* values_host, void* stream);
int cq_assemble_vector_host(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                            const cq_space* space, const cq_quadrature* qr_host, double* b_dev,
                            double* b_host, void* stream);
int cq_assemble_scalar_host(cq_ctx* ctx, const cq_form_desc* form, const cq_mesh* mesh,
                            const cq_space* space, const cq_quadrature* qr_host, double* m_host,
                            void* stream);

/* Several bilinear forms over the SAME qr tuple (the reference's demo integrates its bulk rule into a matrix
 * and a vector, demo/poisson.py:193,209; bench.py assembles stiffness and mass of two spaces): the rule is
 * uploaded once, in chunks; every chunk runs the kernels of all jobs; the finished rows of every matrix go
 * back to the host while later chunks are still coming in.  Each A is zeroed first (assemble_matrix.py:20).
 * Equivalent to njobs calls of cq_assemble_matrix_host, minus (njobs - 1) uploads of the rule. */
typedef struct cq_matrix_job {
  const cq_form_desc* form;
  const cq_space* space;
  const cq_csr* A;
  double* values_host;  /* [A->nnz] host (ideally pinned) or NULL: keep the result on the device */
} cq_matrix_job;
int cq_assemble_matrices_host(cq_ctx* ctx, int njobs, const cq_matrix_job* jobs, const cq_mesh* mesh,
                              const cq_quadrature* qr_host, void* stream);

/* ---- multi-GPU exchange inside the library (one process per GPU) ------------------------------------
 * The NCCL communicator lives in the ctx, so that a non-Python host can run the cell-partitioned assembly
 * (cq_exchange_offrank).  libnccl.so.2 is resolved at run time (the copy already loaded into the
 * process, e.g. torch's, else the system one): the library does not link against it.
 *   rank 0: cq_comm_unique_id(id); broadcast id by any means; every rank: cq_comm_init(ctx, n, rank, id). */
#define CQ_COMM_ID_BYTES 128
int cq_comm_unique_id(unsigned char id[CQ_COMM_ID_BYTES]);
int cq_comm_init(cq_ctx* ctx, int nranks, int rank, const unsigned char id[CQ_COMM_ID_BYTES]);
int cq_comm_destroy(cq_ctx* ctx);
/* The stash exchange of A.assemble() for a slab partition (customquad_b200/dist.py): this rank's value slots
 * [send_lo, send_hi) -- its contributions to rows owned by rank `up` -- are sent to `up` (up < 0: nothing to
 * send); `recv_count` values arrive from rank `down` (down < 0: nothing) and are added to the rows
 * [recv_row0, recv_row0 + recv_nrows): values[indptr[recv_row0 + r] + i] += recv[recv_indptr[r] + i]
 * (a sender row is a prefix of the owner's row; recv_indptr [recv_nrows + 1] int64 on the device,
 * recv_indptr[recv_nrows] == recv_count).  One grouped ncclSend / ncclRecv plus cq_csr_accumulate_rows,
 * all on `stream`. */
int cq_exchange_offrank(cq_ctx* ctx, const cq_csr* A, int up, int64_t send_lo, int64_t send_hi, int down,
                        int64_t recv_row0, int64_t recv_nrows, const int64_t* recv_indptr, int64_t recv_count,
                        void* stream);
/* The same for a vector: b[send_lo .. send_hi) goes up, recv_count values from `down` are added at b[recv_lo ..). */
int cq_exchange_offrank_vector(cq_ctx* ctx, double* b, int up, int64_t send_lo, int64_t send_hi, int down,
                               int64_t recv_lo, int64_t recv_count, void* stream);
/* In-place sum of n doubles (device) over all ranks: assemble_scalar's missing MPI reduction (assemble_scalar.py:34). */
int cq_allreduce_sum(cq_ctx* ctx, double* x, int64_t n, void* stream);

/* ---- neighbour fence through peer memory (no NCCL in the step) ------------------------------------------
 * The fused peer scatter needs two stream-ordered barriers with the two slab neighbours per matrix (after the
 * zeroing, after the kernel).  Each ctx owns two flag words; the neighbours map them (cq_ipc_export of
 * cq_peer_flags(ctx), cq_ipc_open on their side) and hand the mapped addresses to cq_peer_fence_setup:
 * flag_in_down / flag_in_up = the word of the DOWN / UP neighbour that this rank writes (NULL: no such
 * neighbour).  cq_peer_fence enqueues one single-thread kernel on `stream`: release-store the next epoch into
 * the neighbours' words, then acquire-spin on the own words until both neighbours have reached it.  Every rank
 * has to call it the same number of times. */
void* cq_peer_flags(cq_ctx* ctx);  /* device address of this ctx's two flag words: [0] written by down, [1] by up */
int cq_peer_fence_setup(cq_ctx* ctx, void* flag_in_down, void* flag_in_up);
int cq_peer_fence(cq_ctx* ctx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CQ_B200_H */
