"""Diagnostic for the N = 8 weak-scaling parity figure (max_rel 2.1e-12 at x ~ 20, h = 1.3e-3): one GPU assembles the
two cell layers around the interface of rank 7's slab; the result is held against the oracle on identical inputs and
against the oracle on the translated geometry (interface at x = 0).  Prints one JSON line per form."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import dist as cqdist, forms  # noqa: E402
from customquad_b200.quadrature import synthetic_ragged  # noqa: E402
from oracle import oracle  # noqa: E402

MESH_N, XMIN, XMAX = bench.MESH_N, bench.XMIN, bench.XMAX
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = world - 1
hx = (XMAX[0] - XMIN[0]) / MESH_N
slab = cqdist.make_slab([[XMIN[0], XMIN[1]], [XMIN[0] + MESH_N * world * hx, XMAX[1]]], [MESH_N * world, MESH_N], rank, world)
mesh = slab.mesh
tol, kappa = bench.parity_tolerance(mesh)
cells = np.arange(2 * MESH_N, dtype=np.int32)
qr = synthetic_ragged(cells, 2, 16, 64, seed=3)
x = mesh.geometry.x
for variant in (0, 1):
    cq.assemble.set_variant(variant)
    for deg in (1, 2):
        V = cq.FunctionSpace(mesh, ("Lagrange", deg))
        for name in ("stiffness", "mass"):
            f = getattr(forms, name)(V)
            A = cq.assemble_matrix(f, [qr])
            A.assemble()
            S = A.to_scipy()
            pat = oracle.build_pattern(V.dofmap.list, V.dofmap.num_dofs, V.bs)
            R = oracle.assemble_matrix(f, [qr], pattern=pat, nthreads=16)
            mesh.geometry.x = x - np.array([x[:, 0].min() + hx, 0, 0])
            R0 = oracle.assemble_matrix(f, [qr], pattern=pat, nthreads=16)
            mesh.geometry.x = x
            sc = np.abs(R0.data).max()
            print(json.dumps({"variant": variant, "deg": deg, "form": name, "kappa": kappa, "same_pattern": bool(np.array_equal(S.indices, R.indices)),
                              "gpu_vs_oracle": float(np.abs(S.data - R.data).max() / sc), "gpu_vs_recentred": float(np.abs(S.data - R0.data).max() / sc),
                              "oracle_vs_recentred": float(np.abs(R.data - R0.data).max() / sc)}), flush=True)
