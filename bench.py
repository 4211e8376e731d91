#!/usr/bin/env python
"""bench.py -- runtime-quadrature assembly throughput on B200 (BASELINE.json metric).

Headline workload (BASELINE.json configs[1], SURVEY.md 8d "cfg 2"): 2-D Q1 and Q2 stiffness + mass matrices on a
2048x2048 quadrilateral mesh with synthetic ragged runtime quadrature, nq_c ~ U{16..64} points per cell, every cell
listed.  One step = one pass of the hot path over that batch: four ``assemble_matrix`` calls (Q1 stiffness, Q1 mass,
Q2 stiffness, Q2 mass), each zeroing its CSR values first as the reference does (assemble_matrix.py:19-20).

    value = element tensors per second = 4 * num_cells / step time   (inputs resident in HBM)
    e2e   = the same through cq_assemble_matrices_host: the qr arrays start in pinned host memory, are uploaded ONCE
            per step and serve all four forms (the reference's demo reuses one rule for a matrix and a vector,
            demo/poisson.py:193,209); the CSR values of all four matrices come back to pinned host memory.

N>1 (torchrun): the mesh grows with N (weak scaling): rank r owns the slab of 2048 cell-columns [r*2048, (r+1)*2048) of
a (2048*N) x 2048 mesh, assembles into its own rows; the one interface dof line it does not own is added into the
owner's matrix over NVLink by the assembly kernels themselves (customquad_b200/dist.py).  After the timed region every
rank > 0 re-assembles the two cell layers around its interface with the CPU oracle and compares its interface rows
(``parity``; a mismatch makes the run exit non-zero).  ``extra.config5`` (N>1) is BASELINE.json configs[4]: Q1
hexahedra 512^3 strong-scaled over the N GPUs (bulk + cut-cell stiffness, Nitsche with normals, mass), with its own
interface parity check.

``--impl reference`` times the CPU oracle (oracle/, the restatement of the reference's numba loop + generated kernel:
the reference itself needs dolfinx/basix/ffcx/PETSc, absent here) with all host cores on a bounded sample of the same
workload.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MESH_N = 2048
NQ_LO, NQ_HI = 16, 64
XMIN, XMAX = [-1.11, -1.51], [1.55, 1.22]  # demo/poisson.py:42-43
if os.environ.get("CQ_BENCH_XSHIFT"):  # development: the same mesh moved along x (the coordinate magnitudes of the 8-GPU slabs on fewer GPUs); shown in config.workload
    XMIN, XMAX = [XMIN[0] + float(os.environ["CQ_BENCH_XSHIFT"]), XMIN[1]], [XMAX[0] + float(os.environ["CQ_BENCH_XSHIFT"]), XMAX[1]]
BOX3 = [[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]]  # demo/poisson.py:48-49
FORMS = [("q1_stiffness", 1, "stiffness"), ("q1_mass", 1, "mass"), ("q2_stiffness", 2, "stiffness"), ("q2_mass", 2, "mass")]
RTOL = 1e-12  # north_star: values within 1e-12 relative


def log(msg):
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def host_cores() -> int:
    """Cores this process may use -- NOT omp_get_max_threads(): torch.distributed.run exports OMP_NUM_THREADS=1."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_fp64():
    """FP64 peaks from profiles/microbench/fp64_peak.cu (run on this pool's B200, profiles/r2/fp64_peak.json)."""
    p = os.path.join(ROOT, "profiles", "r2", "fp64_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["dfma_tflops"]), float(d["dmma_8x8x4_tflops"]), "measured (profiles/r2/fp64_peak.json)"
    return 37.2, 37.2, "computed (148 SM x 64 FMA/clk x 1.965 GHz)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}  # fmt: skip


@contextlib.contextmanager
def near_gpu(local_rank):
    """Run a block on the cores next to GPU `local_rank` (its PCIe root / NUMA node), so that pinned host buffers
    allocated and first touched inside it are NUMA-local to the copy engine.  No-op when sysfs has no answer."""
    old = None
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{getattr(pr, 'pci_device_id', 0):02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        old = os.sched_getaffinity(0)
        cpus &= old
        if cpus and cpus != old:
            os.sched_setaffinity(0, cpus)
        else:
            old = None
    except Exception:  # noqa: BLE001 -- best effort
        old = None
    try:
        yield
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)


def algorithmic_bytes(num_cells, num_points, gdim, nv, n, num_nodes, nnz, normals=False):
    """per listed cell 4 (cell id) + 8 (int64 offset) + nq*8*(d+1 (+d with normals)) + 4*nv (geometry
    dofmap) + 4*n (dofmap); + 24 per distinct geometry node; + 8*nnz output.  No credit for CSR structure reads,
    atomics read-modify-write, or the memset."""
    return num_cells * (4 + 8 + 4 * nv + 4 * n) + num_points * 8 * (gdim + 1 + (gdim if normals else 0)) + 24 * num_nodes + 8 * nnz


def point_flops(deg, d, kind):
    """SURVEY.md 8d "algorithmic flops per unit of work": per quadrature point, point-by-point evaluation."""
    n, nv = (deg + 1) ** d, 2**d
    tab = 4 * deg * d * (deg + 1) + n * (d - 1) + n * d * (d - 1)
    geom = 2 * nv * d * d + (8 if d == 2 else 50)
    if kind == "stiffness":
        return tab + geom + 2 * n * d * d + 2 * n * n * d + n * d
    return tab + geom + 2 * n * n + n


def build_problem(nx, ny, rank=0, world=1, distort=0.0):
    """Mesh (slab `rank` of a (nx*world) x ny mesh) + spaces + synthetic ragged rule for the cells this rank assembles
    (SURVEY.md 8d generator, seed 20231 + rank).  distort > 0: every interior vertex is moved by up to distort*h/2 in
    each direction (no cell of the mesh stays axis-aligned: the point-by-point kernels run)."""
    import customquad_b200 as cq
    from customquad_b200 import dist as cqdist, forms
    from customquad_b200.quadrature import synthetic_ragged

    hx = (XMAX[0] - XMIN[0]) / MESH_N
    slab = cqdist.make_slab([[XMIN[0], XMIN[1]], [XMIN[0] + nx * world * hx, XMAX[1]]], [nx * world, ny], rank, world)
    mesh = slab.mesh
    if distort > 0.0:
        x = mesh.geometry.x
        h = mesh.h()
        rng = np.random.default_rng(99)
        lx, ly = mesh.shape[0] + 1, mesh.shape[1] + 1
        interior = np.zeros((lx, ly), dtype=bool)
        interior[1:-1, 1:-1] = True
        interior = interior.reshape(-1)
        x[interior, 0] += distort * h[0] * (rng.random(interior.sum()) - 0.5)
        x[interior, 1] += distort * h[1] * (rng.random(interior.sum()) - 0.5)
    spaces = {1: cq.FunctionSpace(mesh, ("Lagrange", 1)), 2: cq.FunctionSpace(mesh, ("Lagrange", 2))}
    fl = [(name, getattr(forms, kind)(spaces[deg]), deg) for name, deg, kind in FORMS]
    qr = synthetic_ragged(slab.cells, 2, NQ_LO, NQ_HI, seed=20231 + rank)
    return mesh, spaces, fl, qr, slab


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle on all host cores, a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from oracle.oracle import build_pattern  # the oracle's own builder (oracle/cq_oracle.c)

    oracle.build()
    threads = host_cores()  # the same at every N (torchrun's OMP_NUM_THREADS=1 is overridden by the explicit count)
    sample_cols = int(os.environ.get("CQ_BENCH_SAMPLE_COLS", "192"))  # 192 x 2048 = 393,216 cells per step and form (~9% of the workload); the override is for tests/test_bench_contract.py
    mesh, spaces, fl, qr, _ = build_problem(sample_cols, MESH_N)
    pats = {d: build_pattern(V.dofmap.list, V.dofmap.num_dofs) for d, V in spaces.items()}

    def step():
        for _, form, deg in fl:
            oracle.assemble_matrix(form, [qr], nthreads=threads, pattern=pats[deg])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = len(fl) * mesh.num_cells / dt
    sample = f"{sample_cols}x{MESH_N} cell slab ({mesh.num_cells} cells, {qr.num_points} points) of the 2048x2048 workload, all 4 forms per step"
    line = {
        "impl": "reference", "metric": "element_tensors_per_s", "value": value, "unit": "element tensors/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "element tensors/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "element tensors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "csr_nnz_per_s": 2 * (pats[1].nnz + pats[2].nnz) / dt,
    }  # fmt: skip
    emit(line)


def workload_config(n_gpus, kernel_path="moment", fused_zero=None):
    return {
        "workload": f"2-D Q1+Q2 stiffness+mass, {MESH_N * n_gpus}x{MESH_N} quad mesh ({MESH_N}x{MESH_N} per GPU), synthetic ragged runtime quadrature U{{{NQ_LO}..{NQ_HI}}} pts/cell, all cells listed (BASELINE.json configs[1])" + (f" [development: mesh shifted along x, CQ_BENCH_XSHIFT={os.environ['CQ_BENCH_XSHIFT']}]" if os.environ.get("CQ_BENCH_XSHIFT") else ""),
        "cells_per_gpu": MESH_N * MESH_N, "forms": [f[0] for f in FORMS], "mean_points_per_cell": (NQ_LO + NQ_HI) / 2,
        "l2": "inputs (>= 4 GB of points per pass) are larger than the 126 MB L2; no explicit flush",
        "parallelism": f"cells{n_gpus}" if n_gpus > 1 else "single",
        "kernel_path": kernel_path + (" (monomial-moment kernels: every cell of the tensor mesh is axis-aligned; the general-cell point-by-point path is timed in extra.general_cells)" if kernel_path == "moment" else " (point-by-point kernels)"),
        "zeroing": {True: "in-kernel wherever a kernel can (cq_ctx_set_fused_zero(1))", False: "memset before every kernel (cq_ctx_set_fused_zero(0))",
                    None: "library default (cq_ctx_set_fused_zero(-1)): a memset before the Q1 kernels; the Q2 matrices are zeroed inside their kernels"}[fused_zero],
        "parity_pinned_by_reference_tests": "Q1 mass-type only (test/test_linear_basis.py); Q1/Q2 stiffness and Q2 mass are pinned by sympy closed forms and the demo's convergence (DESIGN.md 6)",
        "exchange": "interface rows of the neighbour rank: red.global.add.f64 into peer memory over NVLink from the assembly kernels, neighbour fences through peer-memory flags (CQ_PEER_SCATTER=0: ncclSend/ncclRecv + accumulate kernel, cq_exchange_offrank)" if n_gpus > 1 else "none",
    }  # fmt: skip


def cpu_baseline_sample():
    """Oracle (port) on the GPU box's host cores: a bounded sample of the same workload."""
    from oracle import oracle
    from oracle.oracle import build_pattern  # the oracle's own builder (oracle/cq_oracle.c)

    threads = host_cores()
    cols = 512  # 512 x 2048 = 1,048,576 cells (a quarter of the workload), all 4 forms once
    mesh, spaces, fl, qr, _ = build_problem(cols, MESH_N)
    pats = {d: build_pattern(V.dofmap.list, V.dofmap.num_dofs) for d, V in spaces.items()}
    oracle.assemble_matrix(fl[0][1], [qr.select(np.arange(4096))], nthreads=threads, pattern=pats[1])  # warm up threads
    t0 = time.perf_counter()
    for _, form, deg in fl:
        oracle.assemble_matrix(form, [qr], nthreads=threads, pattern=pats[deg])
    dt = time.perf_counter() - t0
    # one core (the reference's cell loop is serial, assemble_matrix.py:84): the first 64 columns of the same slab
    n1 = 64 * MESH_N
    q1 = qr.select(np.arange(n1))
    t1 = time.perf_counter()
    for _, form, deg in fl:
        oracle.assemble_matrix(form, [q1], nthreads=1, pattern=pats[deg])
    dt1 = time.perf_counter() - t1
    # structure-faithful flavour: call_basix rebuilds heap tables per cell and per table like call_basix.h:49-57
    nf = 32 * MESH_N
    qf = qr.select(np.arange(nf))
    t2 = time.perf_counter()
    for _, form, deg in fl:
        oracle.assemble_matrix(form, [qf], nthreads=threads, faithful=True, pattern=pats[deg])
    dt2 = time.perf_counter() - t2
    return {"value": len(fl) * mesh.num_cells / dt, "unit": "element tensors/s", "cores": threads, "kind": "port", "seconds": dt,
            "sample": f"{cols}x{MESH_N} cell slab ({mesh.num_cells} cells, {qr.num_points} points) of the workload, all 4 forms once; oracle/cq_oracle.c allocation-free flavour, OpenMP over cells",
            "single_core": {"value": len(fl) * n1 / dt1, "unit": "element tensors/s", "cores": 1, "seconds": dt1,
                            "sample": f"64x{MESH_N} cell slab ({n1} cells, {q1.num_points} points), all 4 forms once"},
            "structure_faithful": {"value": len(fl) * nf / dt2, "unit": "element tensors/s", "cores": threads, "seconds": dt2,
                                   "sample": f"32x{MESH_N} cell slab ({nf} cells), all 4 forms once; heap basis tables rebuilt per cell and table (call_basix.h:49-57 allocation pattern; basix::create_element itself cannot be reproduced)"}}  # fmt: skip


# ------------------------------------------------------------------------------------------------
def parity_tolerance(mesh):
    """1e-12 (north_star), widened by the conditioning of the geometry itself: the reference's generated kernel forms
    J = sum_v x_v grad(psi_v) in floating point, whose own relative rounding error is eps * kappa, kappa = max|x| / h
    (the weak-scaling mesh grows with N: at N = 8 its x-coordinates reach 20 with h = 1.3e-3, kappa = 1.6e4).  Two
    correct FP64 evaluations of the same element tensor -- vertex differences here, the table sum in the oracle / the
    reference -- may differ by a few eps * kappa; the sparsity stays bit-exact.  (Measured on one GPU with rank 7's
    slab, profiles/r2s2/far_mesh_parity.jsonl: oracle vs the oracle on the recentred geometry 5e-13, CUDA path vs the
    recentred oracle 2e-14.  The 2.1e-12 the 8-GPU run first showed was something else: neighbouring slab meshes that
    disagreed on their shared vertices by an ulp, fixed in dist.make_slab.)"""
    x = mesh.geometry.x
    kappa = float(np.max(np.abs(x[:, : mesh.geometry.dim]))) / float(np.min(mesh.h()))
    return max(RTOL, 4.0 * np.finfo(np.float64).eps * kappa), kappa


def compare_interface_rows(A, R, r0, r1, tol=RTOL):
    """Rows [r0, r1) of the device matrix A against the same rows of the oracle's scipy matrix R (built on the pattern
    of the cells around the interface: for these rows that is the full pattern).  Returns (rows, max_rel, ok)."""
    ip = A.indptr[r0 : r1 + 1].cpu().numpy()
    lo, hi = int(ip[0]), int(ip[-1])
    idx = A.indices[lo:hi].cpu().numpy()
    val = A.data[lo:hi].cpu().numpy()
    rip = R.indptr[r0 : r1 + 1].astype(np.int64)
    ridx = R.indices[rip[0] : rip[-1]]
    rval = R.data[rip[0] : rip[-1]]
    same_pattern = np.array_equal(ip - ip[0], rip - rip[0]) and np.array_equal(idx, ridx)
    if not same_pattern:
        return r1 - r0, float("inf"), False
    scale = float(np.max(np.abs(rval))) if rval.size else 1.0
    err = float(np.max(np.abs(val - rval))) / scale if rval.size else 0.0
    nerr = float(np.linalg.norm(val - rval) / max(np.linalg.norm(rval), 1e-300))
    return r1 - r0, max(err, nerr), err <= tol and nerr <= tol


def cat_rules(a, b):
    """Two host rules one after the other."""
    from customquad_b200.quadrature import RaggedQuadrature

    off = np.concatenate([a.offsets, a.offsets[-1] + b.offsets[1:]])
    n = None if a.n is None else np.concatenate([a.n, b.n])
    return RaggedQuadrature(np.concatenate([a.cells, b.cells]).astype(np.int32), off, np.concatenate([a.pts, b.pts]), np.concatenate([a.w, b.w]), n, a.gdim)


def interface_parity_cfg2(rank, world, slab, fl, mats, exch, qr):
    """Every rank > 0: oracle on its ghost cell layer (the neighbour's top layer, rule regenerated from the neighbour's
    seed) + its own first layer; compare the owned interface rows of all four matrices."""
    from customquad_b200 import dist as cqdist
    from customquad_b200.quadrature import synthetic_ragged
    from oracle import oracle

    if rank == 0:
        return {"rows": 0, "max_rel": 0.0, "max_rel_recentred": 0.0, "ok": True, "tol": RTOL, "kappa": 0.0}
    per_layer = MESH_N
    hx = (XMAX[0] - XMIN[0]) / MESH_N
    nb = cqdist.make_slab([[XMIN[0], XMIN[1]], [XMIN[0] + MESH_N * world * hx, XMAX[1]]], [MESH_N * world, MESH_N], rank - 1, world)
    nbq = synthetic_ragged(nb.cells, 2, NQ_LO, NQ_HI, seed=20231 + rank - 1)
    top = nbq.select(np.arange(nbq.num_entries - per_layer, nbq.num_entries))
    top.cells = np.arange(per_layer, dtype=np.int32)  # the neighbour's top layer is this rank's ghost layer (local cells 0..)
    own = qr.select(np.arange(per_layer))
    rule = cat_rules(top, own)
    rows, worst, worst0, ok = 0, 0.0, 0.0, True
    tol, kappa = parity_tolerance(slab.mesh)
    x = slab.mesh.geometry.x
    x_iface = float(x[:, 0].min()) + hx  # the ghost layer is the slab's first cell layer
    for name, form, _deg in fl:
        V = form.function_spaces[0]
        pat = oracle.build_pattern(V.dofmap.list, V.dofmap.num_dofs, V.bs, cells=rule.cells)
        ex = exch[name]
        # (1) identical inputs, tolerance widened by the conditioning of the reference's own Jacobian sum
        R = oracle.assemble_matrix(form, [rule], nthreads=host_cores(), pattern=pat)
        n, e, good = compare_interface_rows(mats[name], R, ex.recv_row0, ex.recv_row0 + ex.line, tol)
        # (2) the oracle on the same cells with the host copy of the geometry translated so that the interface
        # sits at x = 0 (these four integrands do not depend on x: in exact arithmetic the same matrix): 1e-12
        try:
            slab.mesh.geometry.x = x - np.array([x_iface, 0.0, 0.0])[: x.shape[1]]
            R0 = oracle.assemble_matrix(form, [rule], nthreads=host_cores(), pattern=pat)
        finally:
            slab.mesh.geometry.x = x
        _, e0, good0 = compare_interface_rows(mats[name], R0, ex.recv_row0, ex.recv_row0 + ex.line, RTOL)
        rows, worst, worst0, ok = rows + n, max(worst, e), max(worst0, e0), ok and good and good0
    return {"rows": int(rows), "max_rel": worst, "max_rel_recentred": worst0, "ok": bool(ok), "tol": tol, "kappa": kappa}


# ------------------------------------------------------------------------------------------------
def config5_leg(args, rank, world, dev, peak):
    """BASELINE.json configs[4]: full bulk + cut runtime-quadrature assembly, Q1 hexahedra 512^3, cell-partitioned
    (strong scaling) over the N GPUs with the off-rank row exchange.  Unit sphere level set (demo/sphere.hpp): cells
    cut by the surface get U{8..27} bulk points and a surface rule of U{4..16} points with normals, all other cells the
    fixed 8-point rule.  One step: A <- 0; A += stiffness(all cells); A += Nitsche(cut cells); M <- 0; M += mass(all
    cells); interface rows of both matrices exchanged.  The rules are generated on the device."""
    import torch
    import torch.distributed as dist

    import customquad_b200 as cq
    from customquad_b200 import assemble as asm
    from customquad_b200 import dist as cqdist, forms
    from customquad_b200.quadrature import RaggedQuadrature

    N5 = int(os.environ.get("CQ_BENCH_C5_N", "512"))
    lo, hi = np.asarray(BOX3[0]), np.asarray(BOX3[1])
    log(f"config 5: classification of {N5}^3 cells")
    # cut cells per z-layer (device, every rank the same): phi = |x|^2 - 1 at the vertices
    ax = [torch.linspace(lo[d], hi[d], N5 + 1, dtype=torch.float64, device=dev) for d in range(3)]

    def cut_mask(z0, z1):  # cells of layers [z0, z1): [nz, ny, nx] bool
        ins = (ax[2][z0 : z1 + 1, None, None] ** 2 + ax[1][None, :, None] ** 2 + ax[0][None, None, :] ** 2) < 1.0
        cnt = torch.zeros((z1 - z0, N5, N5), dtype=torch.int8, device=dev)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    cnt += ins[dz : dz + z1 - z0, dy : dy + N5, dx : dx + N5].to(torch.int8)
        return (cnt > 0) & (cnt < 8)

    per_layer_cut = torch.cat([cut_mask(z, min(z + 32, N5)).sum(dim=(1, 2)) for z in range(0, N5, 32)]).cpu().numpy()
    # work per layer ~ sum of nq * n^2: 8 points per cell, cut cells 17.5 + 10 (surface) on average
    weights = (N5 * N5 - per_layer_cut) * 8.0 + per_layer_cut * 27.5
    bounds = cqdist.partition_layers(N5, world, weights=weights)
    slab = cqdist.make_slab(BOX3, [N5, N5, N5], rank, world, bounds=bounds)
    mesh = slab.mesh
    per_layer = N5 * N5
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    log(f"config 5: slab of {slab.count} layers ({len(slab.cells)} cells) on rank 0, bounds {list(map(int, bounds))}")
    # ---- rules, born on the device -------------------------------------------------------------
    g = torch.Generator(device=dev)
    g.manual_seed(5150 + rank)
    C = len(slab.cells)
    cells_t = torch.arange(slab.ghost * per_layer, slab.ghost * per_layer + C, dtype=torch.int32, device=dev)
    cut = cut_mask(slab.first, slab.first + slab.count).reshape(-1)
    cut_idx = torch.nonzero(cut).reshape(-1)
    ncut = int(cut_idx.shape[0])
    nq = torch.full((C,), 8, dtype=torch.int64, device=dev)
    nq[cut_idx] = torch.randint(8, 28, (ncut,), generator=g, device=dev)
    off = torch.zeros(C + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(nq, 0)
    P = int(off[-1])
    pts = torch.rand(3 * P, dtype=torch.float64, device=dev, generator=g)
    w = torch.rand(P, dtype=torch.float64, device=dev, generator=g).add_(0.5)
    w.div_(torch.repeat_interleave(nq.to(torch.float64), nq))
    q_bulk = RaggedQuadrature.on_device(cells_t, off, pts, w, None, 3)
    nqs = torch.randint(4, 17, (ncut,), generator=g, device=dev)
    offs = torch.zeros(ncut + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(nqs, 0)
    Ps = int(offs[-1])
    nrm = torch.randn((Ps, 3), dtype=torch.float64, device=dev, generator=g)
    nrm = (nrm / nrm.norm(dim=1, keepdim=True)).reshape(-1)
    ws = torch.rand(Ps, dtype=torch.float64, device=dev, generator=g).add_(0.5)
    ws.div_(torch.repeat_interleave(nqs.to(torch.float64), nqs))
    q_surf = RaggedQuadrature.on_device(cells_t[cut_idx].contiguous(), offs, torch.rand(3 * Ps, dtype=torch.float64, device=dev, generator=g), ws, nrm, 3)
    del cut, nq, nqs
    # ---- matrices, exchange -----------------------------------------------------------------------
    f_stiff, f_nit, f_mass = forms.stiffness(V), forms.nitsche(V, 10.0), forms.mass(V)
    log("config 5: sparsity pattern + scatter map")
    A, M = asm.create_matrix(f_stiff), asm.create_matrix(f_mass)
    exA, exM = cqdist.SlabExchange(V, slab), cqdist.SlabExchange(V, slab)
    exA.set_pattern(A.indptr, A.indices)
    exM.set_pattern(M.indptr, M.indices)
    fused = os.environ.get("CQ_PEER_SCATTER", "1") != "0" and all([exA.enable_peer_scatter(A), exM.enable_peer_scatter(M)])
    state = {"fused": fused}

    def step(timers=None):
        for mat, ex, work in ((A, exA, ((f_stiff, q_bulk), (f_nit, q_surf))), (M, exM, ((f_mass, q_bulk),))):
            mat.data.zero_()  # A.zeroEntries()
            if state["fused"]:
                ex.fence(dev)  # every rank's matrix is zeroed before any rank's kernel adds into a neighbour
            for form, q in work:
                asm.assemble_matrix(form, [q], A=mat)
            if timers is not None:
                timers.append(torch.cuda.Event(enable_timing=True))
                timers[-1].record()
            ex.exchange_matrix(mat.data)  # fused: completion fence; else ncclSend/ncclRecv + accumulate
            if timers is not None:
                timers.append(torch.cuda.Event(enable_timing=True))
                timers[-1].record()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(reps):
        timers = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        for _ in range(reps):
            step(timers)
        t1.record()
        barrier()
        v = torch.tensor([t0.elapsed_time(t1) / reps, sum(timers[i].elapsed_time(timers[i + 1]) for i in range(0, len(timers), 2)) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v[0]), float(v[1])

    log("config 5: warm-up")
    for _ in range(2):
        step()
    A.assemble()
    ms_step, ms_exchange = timed(max(2, min(args.steps, 3)))
    # the same step with the NCCL exchange (cq_exchange_offrank), for the record
    ms_nccl = None
    if fused:
        saved = [(m.peer_row0, m.peer_rowstart, m.peer_target, m.peer_exchange) for m in (A, M)]
        for m, ex in ((A, exA), (M, exM)):
            m.peer_row0, m.peer_rowstart, m.peer_target, m.peer_exchange, ex.peer = None, None, False, None, False
        state["fused"] = False
        step()
        a, b = timed(2)
        ms_nccl = {"ms_per_step": a, "exchange_ms": b}
        for (m, ex), sv in zip(((A, exA), (M, exM)), saved):
            m.peer_row0, m.peer_rowstart, m.peer_target, m.peer_exchange = sv
            ex.peer = True
        state["fused"] = True
        step()  # the matrices checked below come from the fused path
        barrier()
    # ---- parity on the interface: oracle on the ghost layer (the neighbour's rule, sent up) + own first layer ----
    log("config 5: interface parity against the oracle")

    def layer_slice(q, first_cell, ncells):  # host rule of the entries whose cell lies in [first_cell, first_cell + ncells)
        _, (c, o, p_, w_, n_) = q.device_arrays()
        k = torch.nonzero((c >= first_cell) & (c < first_cell + ncells)).reshape(-1)
        if k.numel() == 0:
            e = np.zeros(0)
            return RaggedQuadrature(np.zeros(0, np.int32), np.zeros(1, np.int64), e, e, None if n_ is None else e, 3)
        k0, k1 = int(k[0]), int(k[-1]) + 1  # entries are cell-ascending: a contiguous range
        p0, p1 = int(o[k0]), int(o[k1])
        hn = None if n_ is None else n_[3 * p0 : 3 * p1].cpu().numpy()
        return RaggedQuadrature(c[k0:k1].cpu().numpy(), (o[k0 : k1 + 1] - p0).cpu().numpy(), p_[3 * p0 : 3 * p1].cpu().numpy(), w_[p0:p1].cpu().numpy(), hn, 3)

    top_cell = (slab.ghost + slab.count - 1) * per_layer
    mine_up = [layer_slice(q, top_cell, per_layer) for q in (q_bulk, q_surf)]
    for r in mine_up:
        r.cells = (r.cells - top_cell).astype(np.int32)  # the receiver's ghost layer is its local cells 0..per_layer-1
    got = [None] * world
    dist.all_gather_object(got, None if rank == world - 1 else [(r.cells, r.offsets, r.pts, r.w, r.n) for r in mine_up])
    par = {"rows": 0, "max_rel": 0.0, "ok": True}
    if rank > 0:
        from oracle import oracle

        below = [RaggedQuadrature(*t, 3) for t in got[rank - 1]]
        own = [layer_slice(q, slab.ghost * per_layer, per_layer) for q in (q_bulk, q_surf)]
        rule_bulk, rule_surf = cat_rules(below[0], own[0]), cat_rules(below[1], own[1])
        pat = oracle.build_pattern(V.dofmap.list, V.dofmap.num_dofs, 1, cells=np.arange(2 * per_layer, dtype=np.int32))
        T = host_cores()
        RA = oracle.assemble_matrix(f_stiff, [rule_bulk], nthreads=T, pattern=pat)
        if rule_surf.num_entries:
            RA.data += oracle.assemble_matrix(f_nit, [rule_surf], nthreads=T, pattern=pat).data  # same pattern object: same slots
        RM = oracle.assemble_matrix(f_mass, [rule_bulk], nthreads=T, pattern=pat)
        rows, worst, ok = 0, 0.0, True
        tol5, _ = parity_tolerance(mesh)
        for mat, R, ex in ((A, RA, exA), (M, RM, exM)):
            n_, e_, g_ = compare_interface_rows(mat, R, ex.recv_row0, ex.recv_row0 + ex.line, tol5)
            rows, worst, ok = rows + n_, max(worst, e_), ok and g_
        par = {"rows": int(rows), "max_rel": worst, "ok": bool(ok)}
    del got
    allp = [None] * world
    dist.all_gather_object(allp, par)
    # ---- bookkeeping ---------------------------------------------------------------------------------
    nodes = mesh.geometry.x.shape[0]
    ab = (algorithmic_bytes(C, P, 3, 8, 8, nodes, A.nnz) + algorithmic_bytes(ncut, Ps, 3, 8, 8, 0, 0, normals=True)
          + algorithmic_bytes(C, P, 3, 8, 8, nodes, M.nnz))  # fmt: skip
    tot = torch.tensor([float(C), float(ncut), float(P + Ps), float(ab)], device=dev, dtype=torch.float64)
    mx = tot.clone()
    dist.all_reduce(tot)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    iface = (N5 + 1) ** 2
    return {
        "workload": f"Q1 hex {N5}^3, strong scaling over {world} GPUs (work-balanced z-slabs): stiffness on all cells (8 pts, cut cells U{{8..27}}) + Nitsche with normals on the cut cells (U{{4..16}}) into A, mass on all cells into M; unit-sphere level set, rules generated on the device (BASELINE.json configs[4])",
        "cells": int(tot[0]), "cut_cells": int(tot[1]), "points": int(tot[2]), "ms_per_step": ms_step,
        "element_tensors_per_s": (2 * float(tot[0]) + float(tot[1])) / (ms_step * 1e-3),
        "csr_nnz_per_s": 2.0 * 27 * (N5 + 1) ** 3 / (ms_step * 1e-3),
        "exchange": "fused: red.global.add.f64 into the owner's memory over NVLink + two fences per matrix" if fused else "ncclSend/ncclRecv of the interface rows + accumulate kernel (cq_exchange_offrank)",
        "exchange_ms": ms_exchange, "interface_rows": iface, "interface_bytes_per_matrix": iface * 18 * 8,
        "nccl_exchange_path": ms_nccl,
        "per_gpu_roofline": {"bound": "hbm", "algorithmic_bytes_max_rank": float(mx[3]), "achieved_gbs": float(mx[3]) / ms_step / 1e6, "peak": peak,
                             "frac": float(mx[3]) / ms_step / 1e6 / peak},
        "partition": {"bounds": [int(b) for b in bounds], "max_cells_per_rank": int(mx[0])},
        "parity": {"checked": True, "rows": int(sum(p["rows"] for p in allp)), "max_rel": max(p["max_rel"] for p in allp), "ok": all(p["ok"] for p in allp),
                   "what": "owned interface rows of A and M on every rank > 0 vs the CPU oracle on the two cell layers around the interface"},
        "parity_pinned_by_reference_tests": "Q1 hexahedron mass (test/test_linear_basis.py:41-66); stiffness and Nitsche/normals are pinned by sympy closed forms (DESIGN.md 6)",
    }  # fmt: skip


# ------------------------------------------------------------------------------------------------
def sample_parity(form, sub, A, nthreads):
    """`sub` (a host rule on a sample of cells) assembled by the CUDA path into the zeroed full-pattern matrix A against
    the oracle on the same entries: every (row, column) the oracle produced is looked up in the device CSR (sparsity),
    values to 1e-12.  A is left holding the sample's contributions."""
    from customquad_b200 import assemble as asm
    from oracle import oracle

    V = form.function_spaces[0]
    asm.assemble_matrix(form, [sub], A=A, zero=True)
    A.assemble()
    pat = oracle.build_pattern(V.dofmap.list, V.dofmap.num_dofs, V.bs, cells=sub.cells)
    R = oracle.assemble_matrix(form, [sub], nthreads=nthreads, pattern=pat).tocoo()
    keep = R.data != 0.0
    rr, cc, vv = R.row[keep].astype(np.int64), R.col[keep].astype(np.int64), R.data[keep]
    rows = np.unique(rr)
    ip_lo, ip_hi = A.indptr[rows].cpu().numpy(), A.indptr[rows + 1].cpu().numpy()
    worst, found = 0.0, True
    scale = float(np.max(np.abs(vv)))
    order = np.argsort(rr, kind="stable")
    rr, cc, vv = rr[order], cc[order], vv[order]
    starts = np.searchsorted(rr, rows)
    ends = np.append(starts[1:], len(rr))
    for i in range(0, len(rows), 4096):  # device rows in batches
        lo, hi = ip_lo[i : i + 4096], ip_hi[i : i + 4096]
        for j in range(len(lo)):
            cols = A.indices[int(lo[j]) : int(hi[j])].cpu().numpy() if hi[j] - lo[j] > 4096 else None
            if cols is None:
                seg = slice(int(lo[j]), int(hi[j]))
                cols = A.indices[seg].cpu().numpy()
            vals = A.data[int(lo[j]) : int(hi[j])].cpu().numpy()
            a, b = starts[i + j], ends[i + j]
            pos = np.searchsorted(cols, cc[a:b])
            ok = (pos < len(cols)) & (cols[np.minimum(pos, len(cols) - 1)] == cc[a:b])
            found = found and bool(ok.all())
            if ok.all():
                worst = max(worst, float(np.max(np.abs(vals[pos] - vv[a:b]))) / scale)
    return {"entries": int(sub.num_entries), "rows": int(len(rows)), "values": int(len(vv)), "sparsity_ok": found, "max_rel": worst, "ok": bool(found and worst <= RTOL)}


def event_ms(fn, reps=5, warm=2):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def config3_leg(dev, peak):
    """BASELINE.json configs[2]: sphere on Q1 hexahedra 256^3 -- cut-cell volume integrals plus the surface Nitsche terms
    with quadrature normals, on the rules the device producer generates (cq_sphere_*; demo/poisson.py:100-113 with
    -domain sphere), and the standard 8-point rule on the uncut cells.  One step: A <- 0; A += stiffness(uncut, 8
    points); A += stiffness(cut, runtime rule); A += Nitsche(cut, surface rule with normals)."""
    import torch

    import customquad_b200 as cq
    from customquad_b200 import assemble as asm
    from customquad_b200 import forms, levelset
    from customquad_b200.quadrature import RaggedQuadrature, gauss_tensor_rule

    N3 = int(os.environ.get("CQ_BENCH_C3_N", "256"))
    mesh = cq.create_box(BOX3, [N3, N3, N3])
    levelset.generate_qr_sphere_device(cq.create_box([[-1, -1, -1], [1, 1, 1]], [4, 4, 4]), 0.9)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cut, uncut, _outside, qb, qs = levelset.generate_qr_sphere_device(mesh, 1.0, (0.0, 0.0, 0.0), 4)
    torch.cuda.synchronize()
    producer_ms = 1e3 * (time.perf_counter() - t0)
    pts, w = gauss_tensor_rule(3, 2)
    nq, U = len(w), len(uncut)
    qu = RaggedQuadrature.on_device(torch.from_numpy(np.ascontiguousarray(uncut, dtype=np.int32)).to(dev), torch.arange(U + 1, dtype=torch.int64, device=dev) * nq,
                                    torch.from_numpy(pts.reshape(-1)).to(dev).repeat(U), torch.from_numpy(w).to(dev).repeat(U), None, 3)  # fmt: skip
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    fK, fN = forms.stiffness(V), forms.nitsche(V, 10.0)
    A = asm.create_matrix(fK)
    one = forms.integral(mesh, 1.0)
    vol = float(cq.assemble_scalar(one, [qb])) + float(np.prod(mesh.h())) * U
    area = float(cq.assemble_scalar(one, [qs]))

    def step():
        asm.assemble_matrix(fK, [qu], A=A, zero=True)
        asm.assemble_matrix(fK, [qb], A=A)
        asm.assemble_matrix(fN, [qs], A=A)

    ms = event_ms(step)
    ms_k = {"stiffness_uncut_8pt": event_ms(lambda: asm.assemble_matrix(fK, [qu], A=A)), "stiffness_cut": event_ms(lambda: asm.assemble_matrix(fK, [qb], A=A)),
            "nitsche_cut": event_ms(lambda: asm.assemble_matrix(fN, [qs], A=A)), "zero": event_ms(lambda: A.data.zero_())}  # fmt: skip
    nodes = mesh.geometry.x.shape[0]
    ab = (algorithmic_bytes(U, U * nq, 3, 8, 8, nodes, A.nnz) + algorithmic_bytes(qb.num_entries, qb.num_points, 3, 8, 8, 0, 0)
          + algorithmic_bytes(qs.num_entries, qs.num_points, 3, 8, 8, 0, 0, normals=True))  # fmt: skip
    T = host_cores()
    rng = np.random.default_rng(3)
    pick = np.sort(rng.choice(qb.num_entries, min(32768, qb.num_entries), replace=False))
    par = {"stiffness_cut": sample_parity(fK, qb.select(pick), A, T), "nitsche_cut": sample_parity(fN, qs.select(pick), A, T)}
    out = {"workload": f"sphere on Q1 hexahedra {N3}^3: 8-point rule on the {U} uncut cells, device-generated cut-cell rules on the {len(cut)} cut cells (volume) + surface rules with normals (Nitsche) (BASELINE.json configs[2])",
           "cells": int(mesh.num_cells), "cut_cells": int(len(cut)), "bulk_points": int(qb.num_points), "surface_points": int(qs.num_points),
           "producer_ms": producer_ms, "ms_per_step": ms, "per_kernel_ms": ms_k,
           "zeroing": "A <- 0 as part of the first call of the step (cq_assemble_matrix_zero: a memset before the kernel unless --fused-zero); per_kernel_ms times each kernel alone, adding to A, and a memset of A", "element_tensors_per_s": (U + 2 * len(cut)) / (ms * 1e-3), "csr_nnz_per_s": A.nnz / (ms * 1e-3),
           "roofline": {"bound": "hbm", "algorithmic_bytes": ab, "achieved_gbs": ab / ms / 1e6, "peak": peak, "frac": ab / ms / 1e6 / peak},
           "volume_rel_err": abs(vol - 4 * np.pi / 3) / (4 * np.pi / 3), "area_rel_err": abs(area - 4 * np.pi) / (4 * np.pi),
           "parity": {"checked": True, **{k: v for k, v in par.items()}, "ok": all(v["ok"] for v in par.values()),
                      "what": "a sample of 32768 cut cells assembled alone by the CUDA path and by the CPU oracle: every value the oracle produced, sparsity by lookup in the device CSR, values to 1e-12"},
           "parity_pinned_by_reference_tests": "none of these forms (stiffness, Nitsche with normals): sympy closed forms, DESIGN.md 6"}  # fmt: skip
    del A
    for q in (qu, qb, qs):
        q._dev.clear()
    torch.cuda.empty_cache()
    return out


def config4_leg(dev, peak):
    """BASELINE.json configs[3]: demo/vector_poisson.py-style 3-D vector Poisson on Q2 hexahedra 128^3 (bs = 3), the
    3^3 Gauss rule in every cell as a runtime rule: blocked stiffness matrix (FP64 tensor cores, owner-computes row
    stores, csrc/cq_owner.cuh) -- 9.69e9 non-zeros, 116 GB of CSR on one GPU.  One step: A <- 0 and A += stiffness in
    one call (assemble_matrix(form, qr))."""
    import torch

    import customquad_b200 as cq
    from customquad_b200 import assemble as asm
    from customquad_b200 import forms
    from customquad_b200.quadrature import RaggedQuadrature, gauss_tensor_rule, normalise_qr

    N4 = int(os.environ.get("CQ_BENCH_C4_N", "128"))
    free_b, _ = torch.cuda.mem_get_info(dev)
    need = 5.7e4 * N4**3 + 8e9  # CSR + element blocks + rule, measured at 128^3: 133 GB
    if free_b < need:
        return {"skipped": f"needs {need / 1e9:.0f} GB of free HBM, {free_b / 1e9:.0f} GB free"}
    mesh = cq.create_box(BOX3, [N4, N4, N4])
    V = cq.VectorFunctionSpace(mesh, ("Lagrange", 2))
    pts, w = gauss_tensor_rule(3, 4)
    C, nq = mesh.num_cells, len(w)
    qr = RaggedQuadrature.on_device(torch.arange(C, dtype=torch.int32, device=dev), torch.arange(C + 1, dtype=torch.int64, device=dev) * nq,
                                    torch.from_numpy(pts.reshape(-1)).to(dev).repeat(C), torch.from_numpy(w).to(dev).repeat(C), None, 3)  # fmt: skip
    fK = forms.stiffness(V)
    K = asm.create_matrix(fK, zero=False)
    ms = event_ms(lambda: asm.assemble_matrix(fK, [qr], A=K, zero=True), reps=3, warm=1)
    K.assemble()
    chunks = [K.data[i : i + (1 << 28)] for i in range(0, K.nnz, 1 << 28)]  # no full-size temporaries
    total_rel = abs(sum(float(c.sum()) for c in chunks)) / sum(float(c.abs().sum()) for c in chunks)  # 1^T K 1 = 0 (rigid translations)
    ab = algorithmic_bytes(C, C * nq, 3, 8, 27, mesh.geometry.x.shape[0], K.nnz)
    flops = 2.0 * 3 * 27 * 27 * C * nq  # B^T D B over the three gradient components
    _dfma, dmma, _ = measured_fp64()
    t_hbm, t_f64 = ab / (peak * 1e6), flops / (dmma * 1e9)
    pick = np.sort(np.random.default_rng(2).choice(C, 2048, replace=False))
    sub = normalise_qr((pick, np.tile(pts.reshape(-1), [len(pick), 1]), np.tile(w, [len(pick), 1])), 3)
    par = sample_parity(fK, sub, K, host_cores())
    out = {"workload": f"vector Poisson stiffness, Q2 hexahedra {N4}^3, bs = 3, {nq} points per cell (BASELINE.json configs[3])", "cells": int(C), "dofs": int(V.num_dofs),
           "nnz": int(K.nnz), "csr_GB": K.nnz * 12 / 1e9, "ms_per_step": ms, "element_tensors_per_s": C / (ms * 1e-3), "csr_nnz_per_s": K.nnz / (ms * 1e-3),
           "path": "owner-computes (cq_owner.cuh): element blocks by the tensor-core kernel, every blocked row stored once, zeroEntries fused",
           "roofline": {"bound": "hbm" if t_hbm >= t_f64 else "fp64", "algorithmic_bytes": ab, "algorithmic_flops": flops, "achieved_gbs": ab / ms / 1e6, "peak": peak,
                        "frac": ab / ms / 1e6 / peak, "frac_fp64_dmma": t_f64 / ms},
           "sum_of_all_entries_rel": total_rel,
           "parity": {"checked": True, **par, "what": "2048 random cells assembled alone by the CUDA path (same kernels) and by the CPU oracle: every value the oracle produced, values to 1e-12"},
           "parity_pinned_by_reference_tests": "none (the reference's Q2 tests are xfail): sympy closed forms and the convergence of examples/vector_poisson.py, DESIGN.md 6"}  # fmt: skip
    del K, chunks
    qr._dev.clear()
    torch.cuda.empty_cache()
    return out


def general_cells_leg(dev, peak):
    """The headline workload on a DISTORTED mesh: no cell is axis-aligned, cq_mesh_classify says so, and the same four
    assemble_matrix calls run the point-by-point staged kernels (MODE 0) instead of the moment kernels."""
    import torch

    from customquad_b200 import assemble as asm

    mesh, spaces, fl, qr, _ = build_problem(MESH_N, MESH_N, distort=0.3)
    asm.qr_on_device(qr, dev)
    dfma, _dmma, src = measured_fp64()
    out = {"mesh": "2048x2048 quadrilaterals, every interior vertex moved by up to 0.15 h in x and y (seeded): bilinear, non-affine cells",
           "kernel_path": "point (assemble_staged_kernel MODE 0)", "fp64_peak_tflops": dfma, "fp64_peak_source": src, "per_kernel": []}
    assert not (asm.mesh_on_device(mesh, dev)[2] & 1), "the distorted mesh was classified axis-aligned"
    total = 0.0
    for name, form, deg in fl:
        A = asm.create_matrix(form)
        ts = []
        for _ in range(4):
            A.data.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            asm.assemble_matrix(form, [qr], A=A)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        A.assemble()
        ms = float(np.mean(ts[1:]))
        total += ms
        ab = algorithmic_bytes(qr.num_entries, qr.num_points, 2, 4, (deg + 1) ** 2, mesh.geometry.x.shape[0], A.nnz)
        fl_ = point_flops(deg, 2, name.split("_")[1]) * qr.num_points
        t_hbm, t_f64 = ab / (peak * 1e6), fl_ / (dfma * 1e9)  # ms
        out["per_kernel"].append({"kernel": name, "ms": ms, "frac_hbm": t_hbm / ms, "frac_fp64": t_f64 / ms, "bound": "fp64" if t_f64 > t_hbm else "hbm",
                                  "frac_binding": max(t_hbm, t_f64) / ms, "algorithmic_flops": fl_})  # fmt: skip
        del A
    out["ms_per_step_kernels"] = total
    out["element_tensors_per_s"] = len(fl) * qr.num_entries / (total * 1e-3)
    qr._dev.clear()
    torch.cuda.empty_cache()
    return out


def list_input_leg(qr):
    """The reference's own input form: lists of per-cell arrays (demo/poisson.py:109-113).  Time of normalising the
    full 4.2 M-cell rule into the flat ragged arrays (customquad_b200/quadrature.py; csrc/cq_pyflatten.c)."""
    from customquad_b200.quadrature import _flatten_helper, normalise_qr

    off = qr.offsets
    t0 = time.perf_counter()
    pts = [qr.pts[2 * off[k] : 2 * off[k + 1]] for k in range(qr.num_entries)]
    w = [qr.w[off[k] : off[k + 1]] for k in range(qr.num_entries)]
    t_make = time.perf_counter() - t0
    t0 = time.perf_counter()
    q = normalise_qr((qr.cells, pts, w), 2)
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(q.offsets, off) and np.array_equal(q.w, qr.w) and np.array_equal(q.pts, qr.pts))
    return {"seconds": dt, "cells": int(qr.num_entries), "ns_per_cell": dt / qr.num_entries * 1e9, "identical_to_flat_rule": ok,
            "helper": "csrc/cq_pyflatten.c" if _flatten_helper() is not None else "numpy (helper not built)",
            "note": f"once per rule, outside the timed e2e region (building the {qr.num_entries} Python arrays themselves took {t_make:.1f} s)"}  # fmt: skip


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import customquad_b200 as cq  # noqa: F401
    from customquad_b200 import _lib, assemble as asm
    from customquad_b200.quadrature import RaggedQuadrature

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    if args.fused_zero or args.no_fused_zero:
        asm.set_fused_zero(args.fused_zero)

    log("building problem")
    mesh, spaces, fl, qr, slab = build_problem(MESH_N, MESH_N, rank, world)
    C_cells, P_pts = qr.num_entries, qr.num_points
    mats = {name: asm.create_matrix(form) for name, form, _ in fl}
    kernel_path = "moment" if (asm.mesh_on_device(mesh, dev)[2] & 1) else "point"
    exch = None
    if world > 1:
        from customquad_b200 import dist as cqdist

        exch = {}
        for name, _, deg in fl:
            ex = cqdist.SlabExchange(spaces[deg], slab)
            ex.set_pattern(mats[name].indptr, mats[name].indices)
            exch[name] = ex
        # fused exchange: the assembly kernels add the interface rows straight into the owner's matrix over
        # NVLink (peer memory); falls back to the NCCL send/recv + accumulate path if peer mapping fails
        if os.environ.get("CQ_PEER_SCATTER", "1") != "0":
            fused = all([exch[name].enable_peer_scatter(mats[name]) for name, _, _ in fl])
            if not fused:
                for name in mats:
                    mats[name].peer_row0, mats[name].peer_rowstart, mats[name].peer_target, mats[name].peer_exchange = None, None, False, None
                    exch[name].peer = False
            log(f"interface exchange: {'peer RED over NVLink (fused in the assembly kernels)' if fused else 'NCCL send/recv'}")
    asm.qr_on_device(qr, dev)  # inputs resident in HBM before the timed region
    fused = exch is not None and all(getattr(e, "peer", False) for e in exch.values())

    def step(events=None):
        for i, (name, form, _) in enumerate(fl):
            A = mats[name]
            if events is not None:
                events[i][0].record()
            # A.zeroEntries() (assemble_matrix.py:20) + assembly: cq_assemble_matrix_zero; with the peer scatter the call
            # zeroes, fences with the neighbours (their kernels add into this matrix), then assembles
            asm.assemble_matrix(form, [qr], A=A, zero=True)
            if exch is not None:
                exch[name].exchange_matrix(A.data)  # fused: only the completion fence
            if events is not None:
                events[i][1].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    for name in mats:
        mats[name].assemble()  # checks the sparsity error flag once, outside the timed region
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in fl] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = _lib.launch_count(local_rank)
    wall0 = time.time()
    t_start.record()
    for s in range(args.steps):
        step(ev[s])
    t_end.record()
    barrier()
    wall1 = time.time()
    launches = _lib.launch_count(local_rank) - launches0
    ms_total = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_step = ms_total / args.steps
    per_form_ms = [float(np.mean([ev[s][i][0].elapsed_time(ev[s][i][1]) for s in range(args.steps)])) for i in range(len(fl))]
    log(f"timed region done: {ms_step:.3f} ms/step")

    # ---- multi-GPU parity: the matrices of the last timed step against the oracle around every interface ----
    parity = None
    if world > 1:
        log("interface parity against the oracle")
        mine = interface_parity_cfg2(rank, world, slab, fl, mats, exch, qr)
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        parity = {"checked": True, "rows": int(sum(p["rows"] for p in allp)), "max_rel": max(p["max_rel"] for p in allp), "ok": all(p["ok"] for p in allp),
                  "tol": max(p["tol"] for p in allp), "geometry_kappa": max(p["kappa"] for p in allp),
                  "per_rank_max_rel": [p["max_rel"] for p in allp], "max_rel_recentred": max(p["max_rel_recentred"] for p in allp), "tol_recentred": RTOL,
                  "what": "owned interface rows of the 4 matrices on every rank > 0 (result of the last timed step, exchange included) vs the CPU oracle on the two cell layers around the interface; sparsity bit-exact, values (a) on identical inputs to tol = max(1e-12, 4 eps kappa), kappa = max|x| / h of the rank's slab (bench.py:parity_tolerance: the weak-scaling mesh grows with N, and the reference's J = sum_v x_v grad psi_v carries eps kappa of its own), and (b) to 1e-12 against the oracle on the same cells with the geometry translated to put the interface at x = 0 (max_rel_recentred)"}  # fmt: skip

    # ---- per-kernel roofline (kernel-only times: separate pass, events around the launch alone) ----
    # For every form the kernel the timed step runs: where the library zeroes the matrix inside the assembly kernel
    # (config.zeroing; seen from the launch count of the zeroing call: two small setup kernels + the cooperative kernel)
    # that is the zeroing kernel, timed through the same call as in the step; otherwise the add-only kernel, timed after a
    # separate fill.  ms_add_only: the add-only kernel in both cases (the basis of the earlier rounds' numbers).
    peak, peak_src = measured_peak()
    kern = []
    for i, (name, form, deg) in enumerate(fl):
        A = mats[name]

        def timed_add_only():
            ts = []
            for _ in range(max(3, args.steps)):
                A.data.zero_()
                if fused:
                    exch[name].fence(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                asm.assemble_matrix(form, [qr], A=A)
                e1.record()
                if fused:
                    exch[name].fence(dev)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.mean(ts))

        in_kernel = False
        if world == 1:  # (at N > 1 the form's time in the step includes the exchange: the add-only kernel is reported)
            l0 = _lib.launch_count(local_rank)
            asm.assemble_matrix(form, [qr], A=A, zero=True)
            in_kernel = _lib.launch_count(local_rank) - l0 == 3
        ms_add = timed_add_only()
        ms = per_form_ms[i] if in_kernel else ms_add  # in-kernel zeroing: the call inside the timed region is this kernel (+ its two setup kernels)
        n = (deg + 1) ** 2
        ab = algorithmic_bytes(C_cells, P_pts, 2, 4, n, mesh.geometry.x.shape[0], A.nnz)
        # ms_in_step: what the form costs inside the timed step, zeroing included (memset + kernel, or the zeroing kernel) and,
        # at N > 1, the fence
        kern.append({"kernel": name, "ms": ms, "algorithmic_bytes": ab, "achieved_gbs": ab / ms / 1e6, "frac": ab / ms / 1e6 / peak,
                     "zeroing": "inside this kernel (ms includes it and the two setup kernels of the zeroing ranges)" if in_kernel else "memset before the kernel (a separate launch, not in ms)",
                     "in_kernel_zeroing": in_kernel, "ms_add_only": ms_add, "frac_add_only": ab / ms_add / 1e6 / peak,
                     "ms_in_step": per_form_ms[i], "frac_in_step": ab / per_form_ms[i] / 1e6 / peak,
                     "nnz": A.nnz, "cells_per_s": C_cells / ms * 1e3})  # fmt: skip
    dom = max(kern, key=lambda k: k["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))  # Q2 kernels: the zeroing kernel's traffic, or the add-only kernel's (_<name>_add_only)
        traffic = tj.get(dom["kernel"]) if dom["in_kernel_zeroing"] else tj.get("_" + dom["kernel"] + "_add_only", tj.get(dom["kernel"]))
    ab_step = sum(k["algorithmic_bytes"] for k in kern)
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src, "frac_add_only": dom["frac_add_only"],
                "basis": "the kernel the timed step runs for this form; where it zeroes the matrix itself (per_kernel[i].zeroing) its time includes the zeroing, for which the algorithmic bytes give no credit -- per_kernel[i].frac_add_only is the add-only kernel on a zeroed matrix, the basis of round 1's 0.66",
                "step": {"algorithmic_bytes": ab_step, "ms": ms_step, "frac": ab_step / ms_step / 1e6 / peak, "what": "all four forms, zeroing included"},
                "per_kernel": kern}  # fmt: skip

    # ---- end to end: host (pinned) qr buffers in ONCE per step, CSR values of all four matrices out ----
    e2e = None
    if not args.no_e2e:
        log("e2e: pinning host buffers")
        with near_gpu(local_rank):
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
            hq = RaggedQuadrature(pin(qr.cells), pin(qr.offsets), pin(qr.pts), pin(qr.w), None, 2)
            vals = {name: torch.empty(mats[name].nnz, dtype=torch.float64).pin_memory() for name in mats}
            for v in vals.values():
                v.zero_()  # first touch on the local node
        hq_np = RaggedQuadrature(hq.cells.numpy(), hq.offsets.numpy(), hq.pts.numpy(), hq.w.numpy(), None, 2)

        iface = {}
        if exch is not None:
            for name in mats:
                ex = exch[name]
                ip = mats[name].indptr[ex.recv_row0 : ex.recv_row0 + ex.line + 1 : ex.line].cpu().numpy() if ex.has_down else [0, 0]
                iface[name] = (int(ip[0]), int(ip[-1]))
        saved = {name: (mats[name].peer_row0, mats[name].peer_rowstart, mats[name].peer_target, mats[name].peer_exchange) for name in mats}
        for name in mats:  # the host-buffer call zeroes inside the call: it uses the NCCL exchange
            mats[name].peer_row0, mats[name].peer_rowstart, mats[name].peer_target, mats[name].peer_exchange = None, None, False, None
        names = [name for name, _, _ in fl]

        def e2e_step():
            asm.assemble_matrices_host([f for _, f, _ in fl], hq_np, [mats[n] for n in names], [vals[n] for n in names])
            if exch is not None:
                # the neighbour's contributions to this rank's interface rows arrive after the call's own D2H:
                # refresh those rows of the host copy (one lattice line, ~100 kB)
                for name in names:
                    ex = exch[name]
                    ex.exchange_matrix(mats[name].data, force_nccl=True)
                    if ex.has_down:
                        lo, hi = int(iface[name][0]), int(iface[name][1])
                        vals[name][lo:hi].copy_(mats[name].data[lo:hi], non_blocking=True)
                torch.cuda.current_stream().synchronize()

        log("e2e: warm-up")
        for _ in range(2):
            e2e_step()
        barrier()
        log("e2e: timing")
        e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the host copy of the last e2e step against the device result it was copied from
        chk = float((vals[names[-1]][:1000000].to(dev) - mats[names[-1]].data[:1000000]).abs().max())
        for name in mats:
            mats[name].peer_row0, mats[name].peer_rowstart, mats[name].peer_target, mats[name].peer_exchange = saved[name]
        h2d = qr.cells.nbytes + qr.offsets.nbytes + qr.pts.nbytes + qr.w.nbytes
        d2h = sum(8 * mats[name].nnz for name in mats) + sum(8 * (hi - lo) for lo, hi in iface.values())
        e2e = {"value": len(fl) * C_cells * world / dt, "unit": "element tensors/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "host_copy_max_abs_diff_vs_device": chk,
               "api": "cq_assemble_matrices_host: the pinned host rule is uploaded once per step (8 chunks overlapping the kernels) and serves all 4 forms; finished CSR rows of every matrix stream back while later chunks upload"}  # fmt: skip
        del hq, hq_np, vals

    extra = {}
    if world > 1 and os.environ.get("CQ_BENCH_C5", "1") != "0" and not args.no_extra:
        # free the headline problem first
        for name in list(mats):
            del mats[name]
        exch = None
        qr._dev.clear()
        torch.cuda.empty_cache()
        try:
            extra["config5"] = config5_leg(args, rank, world, dev, peak)
        except Exception as e:  # noqa: BLE001 -- the headline line must still be printed
            import traceback

            traceback.print_exc()
            extra["config5"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        log("cpu baseline sample")
        cpu = cpu_baseline_sample()
        log("cpu baseline done")
    if world == 1 and not args.no_extra:
        log("extra: general (distorted) cells, list-of-arrays input")
        for name in list(mats):
            del mats[name]
        torch.cuda.empty_cache()
        try:
            extra["list_input"] = list_input_leg(qr)
            qr._dev.clear()
            extra["general_cells"] = general_cells_leg(dev, peak)
            if os.environ.get("CQ_BENCH_C3", "1") != "0":
                log("extra: config 3 (sphere, Q1 hexahedra 256^3)")
                extra["config3"] = config3_leg(dev, peak)
            if os.environ.get("CQ_BENCH_C4", "1") != "0":
                log("extra: config 4 (vector Poisson, Q2 hexahedra 128^3)")
                extra["config4"] = config4_leg(dev, peak)
        except Exception as e:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            extra["error"] = f"{type(e).__name__}: {e}"[:300]
    dfma, dmma, fsrc = measured_fp64()
    line = {
        "metric": "element_tensors_per_s", "value": len(fl) * C_cells * world / (ms_step * 1e-3), "unit": "element tensors/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world, kernel_path, True if args.fused_zero else (False if args.no_fused_zero else None)),
        "csr_nnz_per_s": sum(k["nnz"] for k in kern) * world / (ms_step * 1e-3),
        "per_form_ms": dict(zip([f[0] for f in fl], per_form_ms)),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "fp64_peak": {"dfma_tflops": dfma, "dmma_8x8x4_tflops": dmma, "source": fsrc},
    }  # fmt: skip
    if parity is not None:
        line["parity"] = parity
    if extra:
        line["extra"] = extra
    # flush before any teardown: stdout is block-buffered when redirected, and a line lost to a late
    # failure (seen once: an empty output file from a 4-GPU run whose timed region had finished) is a lost run
    emit(line)
    bad = (parity is not None and not parity["ok"]) or any(isinstance(extra.get(k), dict) and extra[k].get("parity", {}).get("ok") is False for k in ("config3", "config4", "config5"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        log("PARITY FAILURE (partition interfaces or an extra config)")
        sys.exit(3)


_JSON_FD = None


def emit(line: dict):
    """The one JSON line of the contract, on the real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    # Libraries write banners to the C stdout (NCCL prints "NCCL version ..." on the first collective):
    # keep file descriptor 1 for the JSON line alone and send everything else to stderr.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (N=1: general cells, list input; N>1: config 5, also CQ_BENCH_C5=0)")
    ap.add_argument("--fused-zero", action="store_true", help="zero the matrices inside the moment kernels wherever a kernel can (cq_ctx_set_fused_zero(1); default: the library's policy, a memset first)")
    ap.add_argument("--no-fused-zero", action="store_true", help="development: never zero inside the kernels (memset first)")
    ap.add_argument("--forms", default="", help="development only: comma-separated subset of the workload's forms (changes the workload; shown in config.forms)")
    ap.add_argument("--only-config", type=int, default=0, choices=[0, 3, 4], help="run just the extra leg of BASELINE.json configs[N-1] on one GPU and print its JSON (tests/test_gpu_full_size.py; not the contract line)")
    args = ap.parse_args()
    if args.only_config:
        import torch

        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        peak, _ = measured_peak()
        if args.no_fused_zero or args.fused_zero:
            from customquad_b200 import assemble as asm

            asm.set_fused_zero(args.fused_zero)
        out = (config3_leg if args.only_config == 3 else config4_leg)(dev, peak)
        emit({"only_config": args.only_config, **out})
        sys.exit(0 if out.get("skipped") or out["parity"]["ok"] else 3)
    if args.forms:
        global FORMS
        FORMS = [f for f in FORMS if f[0] in args.forms.split(",")]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
