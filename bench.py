#!/usr/bin/env python
"""bench.py -- runtime-quadrature assembly throughput on B200 (BASELINE.json metric).

Workload at N=1 (BASELINE.json configs[1], SURVEY.md 8d "cfg 2"): 2-D Q1 and Q2 stiffness + mass
matrices on a 2048x2048 quadrilateral mesh with synthetic ragged runtime quadrature, nq_c ~ U{16..64}
points per cell, every cell listed.  One step = one pass of the hot path over that batch:
four ``assemble_matrix`` calls (Q1 stiffness, Q1 mass, Q2 stiffness, Q2 mass), each zeroing its CSR
values first as the reference does (assemble_matrix.py:19-20).

    value = element tensors per second = 4 * num_cells / step time   (inputs resident in HBM)
    e2e   = the same through cq_assemble_matrix_host: qr arrays in pinned host memory, H2D copies and
            the D2H copy of the CSR values inside the timed region.

N>1 (torchrun): the mesh grows with N (weak scaling): rank r owns the slab of 2048 cell-columns
[r*2048, (r+1)*2048) of a (2048*N) x 2048 mesh, assembles into its own rows and ships the one
interface dof line it does not own to rank r+1 over NCCL (customquad_b200/dist.py).

``--impl reference`` times the CPU oracle (oracle/, the restatement of the reference's numba loop +
generated kernel: the reference itself needs dolfinx/basix/ffcx/PETSc, absent here) with all host
threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
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
FORMS = [("q1_stiffness", 1, "stiffness"), ("q1_mass", 1, "mass"), ("q2_stiffness", 2, "stiffness"), ("q2_mass", 2, "mass")]


def log(msg):
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def algorithmic_bytes(num_cells, num_points, gdim, nv, n, num_nodes, nnz):
    """per listed cell 4 (cell id) + 8 (int64 offset) + nq*8*(d+1) + 4*nv (geometry
    dofmap) + 4*n (dofmap); + 24 per distinct geometry node; + 8*nnz output.  No credit for CSR
    structure reads, atomics read-modify-write, or the memset."""
    return num_cells * (4 + 8 + 4 * nv + 4 * n) + num_points * 8 * (gdim + 1) + 24 * num_nodes + 8 * nnz


def build_problem(nx, ny, rank=0, world=1):
    """Mesh (slab `rank` of a (nx*world) x ny mesh) + spaces + synthetic ragged rule for the cells
    this rank assembles (SURVEY.md 8d generator, seed 20231 + rank)."""
    import customquad_b200 as cq
    from customquad_b200 import dist as cqdist, forms
    from customquad_b200.quadrature import synthetic_ragged

    hx = (XMAX[0] - XMIN[0]) / MESH_N
    slab = cqdist.make_slab([[XMIN[0], XMIN[1]], [XMIN[0] + nx * world * hx, XMAX[1]]], [nx * world, ny], rank, world)
    mesh = slab.mesh
    spaces = {1: cq.FunctionSpace(mesh, ("Lagrange", 1)), 2: cq.FunctionSpace(mesh, ("Lagrange", 2))}
    fl = [(name, getattr(forms, kind)(spaces[deg]), deg) for name, deg, kind in FORMS]
    qr = synthetic_ragged(slab.cells, 2, NQ_LO, NQ_HI, seed=20231 + rank)
    return mesh, spaces, fl, qr, slab


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle on all host threads, a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    from oracle.oracle import build_pattern  # the oracle's own builder (oracle/cq_oracle.c)

    oracle.build()
    threads = oracle.max_threads()
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


def workload_config(n_gpus):
    return {
        "workload": f"2-D Q1+Q2 stiffness+mass, {MESH_N * n_gpus}x{MESH_N} quad mesh ({MESH_N}x{MESH_N} per GPU), synthetic ragged runtime quadrature U{{{NQ_LO}..{NQ_HI}}} pts/cell, all cells listed (BASELINE.json configs[1])",
        "cells_per_gpu": MESH_N * MESH_N, "forms": [f[0] for f in FORMS], "mean_points_per_cell": (NQ_LO + NQ_HI) / 2,
        "l2": "inputs (>= 4 GB of points per pass) are larger than the 126 MB L2; no explicit flush",
        "parallelism": f"cells{n_gpus}" if n_gpus > 1 else "single",
        "exchange": "interface rows of the neighbour rank: red.global.add.f64 into peer memory over NVLink from the assembly kernels (CQ_PEER_SCATTER=0: NCCL send/recv + accumulate kernel)" if n_gpus > 1 else "none",
    }  # fmt: skip


def cpu_baseline_sample():
    """Oracle (port) on the GPU box's host cores: a bounded sample of the same workload."""
    from oracle import oracle
    from oracle.oracle import build_pattern  # the oracle's own builder (oracle/cq_oracle.c)

    threads = oracle.max_threads()
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
    return {"value": len(fl) * mesh.num_cells / dt, "unit": "element tensors/s", "cores": threads, "kind": "port", "seconds": dt,
            "sample": f"{cols}x{MESH_N} cell slab ({mesh.num_cells} cells, {qr.num_points} points) of the workload, all 4 forms once; oracle/cq_oracle.c allocation-free flavour, OpenMP over cells",
            "single_core": {"value": len(fl) * n1 / dt1, "unit": "element tensors/s", "cores": 1, "seconds": dt1,
                            "sample": f"64x{MESH_N} cell slab ({n1} cells, {q1.num_points} points), all 4 forms once"}}  # fmt: skip


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

    log("building problem")
    mesh, spaces, fl, qr, slab = build_problem(MESH_N, MESH_N, rank, world)
    C_cells, P_pts = qr.num_entries, qr.num_points
    mats = {name: asm.create_matrix(form) for name, form, _ in fl}
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
                    mats[name].peer_row0, mats[name].peer_rowstart = None, None
                    exch[name].peer = False
            log(f"interface exchange: {'peer RED over NVLink (fused in the assembly kernels)' if fused else 'NCCL send/recv'}")
    asm.qr_on_device(qr, dev)  # inputs resident in HBM before the timed region
    fused = exch is not None and all(getattr(e, "peer", False) for e in exch.values())

    def step(events=None):
        for i, (name, form, _) in enumerate(fl):
            A = mats[name]
            if events is not None:
                events[i][0].record()
            if fused:
                A.data.zero_()  # A.zeroEntries() (assemble_matrix.py:20)
                exch[name].fence(dev)  # the neighbour's kernel adds into this matrix: it must be zeroed first
                asm.assemble_matrix(form, [qr], A=A)
            else:
                asm.assemble_matrix(form, [qr], A=A, zero=True)  # A.zeroEntries() + assembly in one call (cq_assemble_matrix_zero)
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
    # ---- per-kernel roofline (kernel-only times: separate pass, events around the launch alone) ----
    peak, peak_src = measured_peak()
    kern = []
    for i, (name, form, deg) in enumerate(fl):
        A = mats[name]
        ts = []
        for _ in range(max(3, args.steps)):
            A.data.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            asm.assemble_matrix(form, [qr], A=A)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts))
        n = (deg + 1) ** 2
        ab = algorithmic_bytes(C_cells, P_pts, 2, 4, n, mesh.geometry.x.shape[0], A.nnz)
        kern.append({"kernel": name, "ms": ms, "algorithmic_bytes": ab, "achieved_gbs": ab / ms / 1e6, "frac": ab / ms / 1e6 / peak,
                     "nnz": A.nnz, "cells_per_s": C_cells / ms * 1e3})  # fmt: skip
    dom = max(kern, key=lambda k: k["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(dom["kernel"])
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic, "peak_source": peak_src, "per_kernel": kern}  # fmt: skip

    # ---- end to end: host (pinned) qr buffers in, CSR values out to pinned host memory --------------
    e2e = None
    if not args.no_e2e:
        log("e2e: pinning host buffers")
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
        hq = RaggedQuadrature(pin(qr.cells), pin(qr.offsets), pin(qr.pts), pin(qr.w), None, 2)
        hq_np = RaggedQuadrature(hq.cells.numpy(), hq.offsets.numpy(), hq.pts.numpy(), hq.w.numpy(), None, 2)
        vals = {name: torch.empty(mats[name].nnz, dtype=torch.float64).pin_memory() for name in mats}

        iface = {}
        if exch is not None:
            for name in mats:
                ex = exch[name]
                ip = mats[name].indptr[ex.recv_row0 : ex.recv_row0 + ex.line + 1 : ex.line].cpu().numpy() if ex.has_down else [0, 0]
                iface[name] = (int(ip[0]), int(ip[-1]))
        saved = {name: (mats[name].peer_row0, mats[name].peer_rowstart) for name in mats}
        for name in mats:  # the host-buffer call zeroes inside the call: it uses the NCCL exchange
            mats[name].peer_row0, mats[name].peer_rowstart = None, None

        def e2e_step():
            for name, form, _ in fl:
                asm.assemble_matrix_host(form, hq_np, mats[name], vals[name])
                if exch is not None:
                    # the neighbour's contributions to this rank's interface rows arrive after the call's own D2H:
                    # refresh those rows of the host copy (one lattice line, ~100 kB)
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
        for name in mats:
            mats[name].peer_row0, mats[name].peer_rowstart = saved[name]
        h2d = len(fl) * (qr.cells.nbytes + qr.offsets.nbytes + qr.pts.nbytes + qr.w.nbytes)
        d2h = sum(8 * mats[name].nnz for name in mats) + sum(8 * (hi - lo) for lo, hi in iface.values())
        e2e = {"value": len(fl) * C_cells * world / dt, "unit": "element tensors/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "api": "cq_assemble_matrix_host (pinned host qr in, CSR values out)"}  # fmt: skip

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
    line = {
        "metric": "element_tensors_per_s", "value": len(fl) * C_cells * world / (ms_step * 1e-3), "unit": "element tensors/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "csr_nnz_per_s": sum(mats[n].nnz for n in mats) * world / (ms_step * 1e-3),
        "per_form_ms": dict(zip([f[0] for f in fl], per_form_ms)),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }  # fmt: skip
    # flush before any teardown: stdout is block-buffered when redirected, and a line lost to a late
    # failure (seen once: an empty output file from a 4-GPU run whose timed region had finished) is a lost run
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--forms", default="", help="development only: comma-separated subset of the workload's forms (changes the workload; shown in config.forms)")
    args = ap.parse_args()
    if args.forms:
        global FORMS
        FORMS = [f for f in FORMS if f[0] in args.forms.split(",")]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
