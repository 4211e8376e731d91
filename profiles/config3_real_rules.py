"""Config 3 on real rules: sphere on Q1 hexahedra N^3, cut-cell rules from the GPU producer (cq_sphere_*), then the
cut-cell stiffness matrix, the Nitsche matrix with quadrature normals and the volume / area functionals.
Prints one JSON line per stage (flushed: partial results survive a timeout)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import assemble as asm, forms, levelset  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = lambda **k: print(json.dumps(k), flush=True)  # noqa: E731
t0 = time.perf_counter()
mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [N, N, N])
out(stage="mesh", N=N, s=round(time.perf_counter() - t0, 2))
levelset.generate_qr_sphere_device(cq.create_box([[-1, -1, -1], [1, 1, 1]], [4, 4, 4]), 0.9)  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
cut, uncut, outside, qb, qs = levelset.generate_qr_sphere_device(mesh, 1.0, (0.0, 0.0, 0.0), 4)
torch.cuda.synchronize()
out(stage="producer", ms=round(1e3 * (time.perf_counter() - t0), 1), cut_cells=int(len(cut)), uncut_cells=int(len(uncut)),
    bulk_points=int(qb.num_points), surface_points=int(qs.num_points))
one = forms.integral(mesh, 1.0)
vol = float(cq.assemble_scalar(one, [qb])) + float(np.prod(mesh.h())) * len(uncut)
area = float(cq.assemble_scalar(one, [qs]))
out(stage="functionals", volume_rel_err=abs(vol - 4 * np.pi / 3) / (4 * np.pi / 3), area_rel_err=abs(area - 4 * np.pi) / (4 * np.pi))
t0 = time.perf_counter()
V = cq.FunctionSpace(mesh, ("Lagrange", 1))
K = asm.create_matrix(forms.stiffness(V))
torch.cuda.synchronize()
out(stage="space+pattern", s=round(time.perf_counter() - t0, 2), nnz=int(K.nnz))


def timed(fn, reps=5):
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


out(stage="cut-cell stiffness (kernel, accumulate)", ms=round(timed(lambda: cq.assemble_matrix(forms.stiffness(V), [qb], A=K)), 4), entries=int(qb.num_entries))
out(stage="Nitsche with normals (kernel, accumulate)", ms=round(timed(lambda: cq.assemble_matrix(forms.nitsche(V, 10.0), [qs], A=K)), 4), entries=int(qs.num_entries))
out(stage="volume functional", ms=round(timed(lambda: cq.assemble_scalar(one, [qb])), 4))
