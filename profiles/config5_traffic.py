"""DRAM traffic of the Q1-hexahedron stiffness kernel on one rank's share of config 5 for different traversal tilings
(run under ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:assemble_).
usage: python profiles/config5_traffic.py nz tile_bytes [tile_bytes ...]"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import assemble as asm  # noqa: E402
from customquad_b200 import forms  # noqa: E402
from customquad_b200.quadrature import RaggedQuadrature  # noqa: E402

nz = int(sys.argv[1])
N = int(os.environ.get("C5_N", "512"))
dev = torch.device("cuda", 0)
mesh = cq.create_box(bench.BOX3, [N, N, nz])
V = cq.FunctionSpace(mesh, ("Lagrange", 1))
C = mesh.num_cells
g = torch.Generator(device=dev).manual_seed(1)
q = RaggedQuadrature.on_device(torch.arange(C, dtype=torch.int32, device=dev), torch.arange(C + 1, dtype=torch.int64, device=dev) * 8,
                               torch.rand(3 * 8 * C, dtype=torch.float64, device=dev, generator=g), torch.rand(8 * C, dtype=torch.float64, device=dev, generator=g) / 8, None, 3)
form = forms.stiffness(V)
A = asm.create_matrix(form)
for tile in sys.argv[2:]:
    asm.L2_TILE_BYTES = int(tile)
    A.data.zero_()
    asm.assemble_matrix(form, [q], A=A)
    torch.cuda.synchronize()
    print("tile", tile, "tiles", asm._Call(form, q, dev, A=A).qr.plane_tiles, flush=True)
