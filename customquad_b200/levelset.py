"""Cut-cell quadrature for a circular domain on a tensor quadrilateral mesh: the input *producer*
of the hot path, i.e. what the reference's demo obtains from Algoim through
``demo/algoim_utils.py:92-180`` + ``demo/algoim_utils.hpp:84-153`` with the level set of
``demo/circle.hpp:14-34`` (phi = x^2 + y^2 - R^2).

Algoim is not available here, so this is an own scheme for that one level set, with the same
outputs and conventions as ``generate_qr`` (algoim_utils.py:169-180):

* classification cut / uncut / outside (algoim_utils.hpp:98-100,145-152);
* bulk rule of ``cell ∩ {phi < 0}`` by dimension reduction along a height direction chosen per
  cell (the coordinate in which the circle is a graph with slope <= ~1 inside the cell), Gauss
  points in the abscissa between the break points (cell bounds and the crossings of the circle
  with the cell's far edges), Gauss points in the height direction inside the interval the disk
  cuts out -- the same idea as Algoim's quadGen, specialised to a convex quadric;
* surface rule on ``cell ∩ {phi = 0}``: Gauss points in the polar angle on every arc inside the
  cell, unit outward normals (the demo passes the un-normalised gradient, a FIXME in
  algoim_utils.hpp:107-114; unit normals are what the Nitsche terms of demo/poisson.py:156-159
  need);
* weights divided by the cell volume (algoim_utils.hpp:127-132) and points mapped to the unit
  cell (x - xmin)/dx (algoim_utils.hpp:134-143).

The sphere (``demo/sphere.hpp:14-34``, ``generate_qr_sphere``) is sliced along one axis: every
cross-section of ball ∩ cell is a disk ∩ rectangle, for which the 2-D rules above apply; the
slicing variable is integrated with Gauss points between the heights where the cross-section
changes type (the circle passes a corner of the rectangle or touches an edge, or the ball ends),
after a smooth-step change of variable that removes the square-root end-point behaviour at the
tangencies.  The surface rule uses the hat-box identity dS = R dtheta du.

Host numpy code: it runs once per mesh before the timed path, like the reference's generator.
"""
from __future__ import annotations

import numpy as np


def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _bulk_rule_cell(lo, hi, c, R, n1):
    """Points (physical) and weights (physical measure) of [lo, hi] ∩ disk(c, R)."""
    cx = 0.5 * (lo + hi) - c
    # height direction: the axis along which the centre of the cell is farther from the disk centre
    hdir = 1 if abs(cx[1]) >= abs(cx[0]) else 0
    adir = 1 - hdir
    a0, a1 = lo[adir] - c[adir], hi[adir] - c[adir]
    t0, t1 = lo[hdir] - c[hdir], hi[hdir] - c[hdir]
    brk = {a0, a1}
    for tt in (t0, t1):  # crossings of the circle with the two edges t = const
        if abs(tt) < R:
            s = np.sqrt(R * R - tt * tt)
            for v in (-s, s):
                if a0 < v < a1:
                    brk.add(v)
    for v in (-R, R):  # the disk ends inside the abscissa range
        if a0 < v < a1:
            brk.add(v)
    brk = sorted(brk)
    g, gw = _gauss01(n1)
    P, W = [], []
    for b0, b1 in zip(brk[:-1], brk[1:]):
        if b1 - b0 < 1e-14 * max(1.0, abs(b0)):
            continue
        a = b0 + (b1 - b0) * g
        wa = (b1 - b0) * gw
        half = np.sqrt(np.maximum(R * R - a * a, 0.0))
        tl, th = np.maximum(t0, -half), np.minimum(t1, half)
        ok = th > tl
        for ai, wi, l, h in zip(a[ok], wa[ok], tl[ok], th[ok]):
            t = l + (h - l) * g
            p = np.empty((n1, 2))
            p[:, adir] = ai + c[adir]
            p[:, hdir] = t + c[hdir]
            P.append(p)
            W.append(wi * (h - l) * gw)
    if not P:
        return np.zeros((0, 2)), np.zeros(0)
    return np.concatenate(P), np.concatenate(W)


def _surface_rule_cell(lo, hi, c, R, n1):
    """Points, weights (arc length) and unit outward normals of [lo, hi] ∩ circle(c, R)."""
    ang = []
    for d in (0, 1):
        for edge in (lo[d] - c[d], hi[d] - c[d]):
            if abs(edge) < R:
                other = np.sqrt(R * R - edge * edge)
                for o in (-other, other):
                    if lo[1 - d] - c[1 - d] - 1e-14 <= o <= hi[1 - d] - c[1 - d] + 1e-14:
                        xy = np.empty(2)
                        xy[d], xy[1 - d] = edge, o
                        ang.append(np.arctan2(xy[1], xy[0]))
    inside = lambda p: bool(np.all(p >= lo - 1e-13) and np.all(p <= hi + 1e-13))  # noqa: E731
    arcs = []
    if len(ang) < 2:
        # no crossing: the whole circle is inside the cell (small cross-sections near the poles of a sphere) or outside
        if inside(c + np.array([R, 0.0])) and inside(c - np.array([R, 0.0])) and inside(c + np.array([0.0, R])) and inside(c - np.array([0.0, R])):
            arcs = [(0.0, 2 * np.pi)]
    else:
        ang = np.unique(np.round(np.sort(np.asarray(ang)), 14))
        m = len(ang)
        for k in range(m):
            th0 = ang[k]
            th1 = ang[(k + 1) % m] + (2 * np.pi if k == m - 1 else 0.0)
            if th1 - th0 < 1e-13:
                continue
            mid = 0.5 * (th0 + th1)
            if inside(c + R * np.array([np.cos(mid), np.sin(mid)])):
                arcs.append((th0, th1))
    g, gw = _gauss01(n1)
    P, W, Nn = [], [], []
    for a0, a1 in arcs:
        pieces = max(1, int(np.ceil((a1 - a0) / (0.5 * np.pi))))  # Gauss on arcs of at most a quarter turn
        for j in range(pieces):
            th0, th1 = a0 + (a1 - a0) * j / pieces, a0 + (a1 - a0) * (j + 1) / pieces
            th = th0 + (th1 - th0) * g
            n = np.stack([np.cos(th), np.sin(th)], axis=1)
            P.append(c + R * n)
            W.append(R * (th1 - th0) * gw)
            Nn.append(n)
    if not P:
        return np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2))
    return np.concatenate(P), np.concatenate(W), np.concatenate(Nn)


def _slice_nodes(lo_u, hi_u, brk, n1):
    """Gauss nodes / weights in u on [lo_u, hi_u] split at `brk`, each piece mapped by the smooth step
    u = a + (b - a) x^2 (3 - 2x): du vanishes at both ends, which turns sqrt(u - a) type integrands
    (a cross-section appearing through a tangency) into smooth ones."""
    pts = sorted({lo_u, hi_u, *[b for b in brk if lo_u < b < hi_u]})
    g, gw = _gauss01(n1)
    U, W = [], []
    for a, b in zip(pts[:-1], pts[1:]):
        if b - a < 1e-13 * max(1.0, abs(a)):
            continue
        U.append(a + (b - a) * g * g * (3.0 - 2.0 * g))
        W.append((b - a) * 6.0 * g * (1.0 - g) * gw)
    if not U:
        return np.zeros(0), np.zeros(0)
    return np.concatenate(U), np.concatenate(W)


def _sphere_rules_cell(lo, hi, c, R, n1, axis):
    """Bulk (points, weights) and surface (points, weights, unit normals) rules of [lo, hi] ∩ ball(c, R),
    sliced along `axis`."""
    st = [d for d in range(3) if d != axis]
    u0, u1 = max(lo[axis] - c[axis], -R), min(hi[axis] - c[axis], R)
    if u1 <= u0:
        return (np.zeros((0, 3)), np.zeros(0)), (np.zeros((0, 3)), np.zeros(0), np.zeros((0, 3)))
    lo2, hi2, c2 = lo[st], hi[st], c[st]
    # heights where the cross-section circle (radius sqrt(R^2 - u^2)) passes a corner or touches an edge line
    brk = []
    rel = [np.array([lo2[0] - c2[0], hi2[0] - c2[0]]), np.array([lo2[1] - c2[1], hi2[1] - c2[1]])]
    for a in rel[0]:
        for b in rel[1]:
            d2 = R * R - a * a - b * b
            if d2 > 0:
                brk += [-np.sqrt(d2), np.sqrt(d2)]
    for k in (0, 1):
        for a in rel[k]:
            d2 = R * R - a * a
            if d2 > 0 and rel[1 - k][0] <= 0.0 <= rel[1 - k][1]:
                brk += [-np.sqrt(d2), np.sqrt(d2)]
    n_u = n1 + 2  # the change of variable costs a little accuracy on smooth pieces
    U, WU = _slice_nodes(u0, u1, brk, n_u)
    Pb, Wb, Ps, Ws, Ns = [], [], [], [], []
    for u, wu in zip(U, WU):
        rho2 = R * R - u * u
        if rho2 <= 0:
            continue
        rho = np.sqrt(rho2)
        p2, w2 = _bulk_rule_cell(lo2, hi2, c2, rho, n1)
        if w2.size:
            p = np.empty((w2.size, 3))
            p[:, st] = p2
            p[:, axis] = u + c[axis]
            Pb.append(p)
            Wb.append(w2 * wu)
        q2, wq, nq = _surface_rule_cell(lo2, hi2, c2, rho, n1)
        if wq.size:
            p = np.empty((wq.size, 3))
            p[:, st] = q2
            p[:, axis] = u + c[axis]
            Ps.append(p)
            Ws.append(wq / rho * R * wu)  # arc length rho dtheta -> dtheta; dS = R dtheta du
            Ns.append((p - c) / R)
    cat = lambda L, shape: np.concatenate(L) if L else np.zeros(shape)  # noqa: E731
    return (cat(Pb, (0, 3)), cat(Wb, 0)), (cat(Ps, (0, 3)), cat(Ws, 0), cat(Ns, (0, 3)))


def generate_qr_sphere(mesh, radius=1.0, center=(0.0, 0.0, 0.0), order=4, axis=2):
    """``generate_qr(mesh, NN, degree, "sphere", opts)`` (demo/algoim_utils.py:92-180, demo/sphere.hpp):
    same outputs and conventions as generate_qr_circle, for hexahedral meshes."""
    assert mesh.topology.dim == 3, "sphere: 3-D meshes"
    nx, ny, nz = mesh.shape
    h = mesh.h()
    c = np.asarray(center, dtype=np.float64)
    R = float(radius)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ix, iy, iz = ix.reshape(-1), iy.reshape(-1), iz.reshape(-1)  # cell = (iz * ny + iy) * nx + ix (Mesh.cell_index)
    lo = np.stack([mesh.xmin[0] + ix * h[0], mesh.xmin[1] + iy * h[1], mesh.xmin[2] + iz * h[2]], axis=1)
    hi = lo + h[None, :]
    near = np.clip(c[None, :], lo, hi) - c[None, :]
    far = np.maximum(np.abs(lo - c[None, :]), np.abs(hi - c[None, :]))
    rmin2, rmax2 = np.sum(near * near, axis=1), np.sum(far * far, axis=1)
    outside = rmin2 >= R * R
    uncut = rmax2 <= R * R
    cut_cells = np.nonzero(~(outside | uncut))[0].astype(np.int32)
    vol = float(np.prod(h))
    n1 = max(2, (order + 2) // 2)
    keep, qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n = [], [], [], [], [], []
    for cell in cut_cells:
        (p, w), (pb, wb, nb) = _sphere_rules_cell(lo[cell], hi[cell], c, R, n1, axis)
        if w.size == 0 or w.sum() < 1e-14 * vol:
            outside[cell] = True
            continue
        keep.append(cell)
        qr_pts.append(((p - lo[cell]) / h).reshape(-1))
        qr_w.append(w / vol)
        qr_pts_b.append(((pb - lo[cell]) / h).reshape(-1))
        qr_w_b.append(wb / vol)
        qr_n.append(nb.reshape(-1))
    return (np.asarray(keep, dtype=np.int32), np.nonzero(uncut)[0].astype(np.int32), np.nonzero(outside)[0].astype(np.int32),
            qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n)  # fmt: skip


def generate_qr_circle(mesh, radius=1.0, center=(0.0, 0.0), order=4):
    """``generate_qr(mesh, NN, degree, "circle", opts)`` of demo/algoim_utils.py:92-180.

    Returns (cut_cells, uncut_cells, outside_cells, qr_pts, qr_w, qr_pts_bdry, qr_w_bdry, qr_n):
    the qr_* lists hold one flat float64 array per *cut* cell (demo/poisson.py:107-113 keeps only
    those), points in unit-cell coordinates interleaved, weights cell-relative."""
    assert mesh.topology.dim == 2, "circle: 2-D meshes"
    nx, ny = mesh.shape
    h = mesh.h()
    c = np.asarray(center, dtype=np.float64)
    R = float(radius)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ix, iy = ix.reshape(-1), iy.reshape(-1)  # cell = ix * ny + iy (Mesh.cell_index)
    lo = np.stack([mesh.xmin[0] + ix * h[0], mesh.xmin[1] + iy * h[1]], axis=1)
    hi = lo + h[None, :]
    # nearest / farthest point of each cell from the centre
    near = np.clip(c[None, :], lo, hi) - c[None, :]
    far = np.maximum(np.abs(lo - c[None, :]), np.abs(hi - c[None, :]))
    rmin2, rmax2 = np.sum(near * near, axis=1), np.sum(far * far, axis=1)
    outside = rmin2 >= R * R
    uncut = rmax2 <= R * R
    cut = ~(outside | uncut)
    cut_cells = np.nonzero(cut)[0].astype(np.int32)
    vol = float(np.prod(h))
    n1 = max(2, (order + 2) // 2)
    qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n = [], [], [], [], []
    keep = []
    for cell in cut_cells:
        p, w = _bulk_rule_cell(lo[cell], hi[cell], c, R, n1)
        pb, wb, nb = _surface_rule_cell(lo[cell], hi[cell], c, R, n1)
        if w.size == 0 or w.sum() < 1e-14 * vol:  # the circle only touches the cell
            outside[cell] = True
            continue
        keep.append(cell)
        qr_pts.append(((p - lo[cell]) / h).reshape(-1))
        qr_w.append(w / vol)
        qr_pts_b.append(((pb - lo[cell]) / h).reshape(-1))
        qr_w_b.append(wb / vol)
        qr_n.append(nb.reshape(-1))
    cut_cells = np.asarray(keep, dtype=np.int32)
    return (cut_cells, np.nonzero(uncut)[0].astype(np.int32), np.nonzero(outside)[0].astype(np.int32),
            qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n)  # fmt: skip


def circle_desc(mesh, radius=1.0, center=(0.0, 0.0), order=4):
    """The ``cq_circle`` argument of cq_circle_classify / cq_circle_rules (include/cq_b200.h) for a tensor mesh."""
    from . import _lib

    assert mesh.topology.dim == 2, "circle: 2-D meshes"
    n1 = max(2, (order + 2) // 2)
    assert n1 <= 8, "at most 8 Gauss points per direction"
    g, gw = _gauss01(n1)
    h = mesh.h()
    D = _lib.CircleDesc()
    D.xmin[:] = [float(mesh.xmin[0]), float(mesh.xmin[1])]
    D.h[:] = [float(h[0]), float(h[1])]
    D.n[:] = [int(mesh.shape[0]), int(mesh.shape[1])]
    D.center[:] = [float(center[0]), float(center[1])]
    D.radius = float(radius)
    D.n1 = n1
    D.g[:n1] = list(g)
    D.gw[:n1] = list(gw)
    return D


def sphere_desc(mesh, radius=1.0, center=(0.0, 0.0, 0.0), order=4, axis=2):
    """The ``cq_sphere`` argument of cq_sphere_classify / cq_sphere_rules for a tensor hexahedral mesh."""
    from . import _lib

    assert mesh.topology.dim == 3, "sphere: 3-D meshes"
    n1 = max(2, (order + 2) // 2)
    assert n1 <= 8, "at most 8 Gauss points per direction"
    g, gw = _gauss01(n1)
    gu, gwu = _gauss01(n1 + 2)
    h = mesh.h()
    D = _lib.SphereDesc()
    D.xmin[:] = [float(v) for v in mesh.xmin[:3]]
    D.h[:] = [float(v) for v in h[:3]]
    D.n[:] = [int(v) for v in mesh.shape[:3]]
    D.center[:] = [float(v) for v in center]
    D.radius = float(radius)
    D.n1 = n1
    D.axis = int(axis)
    D.g[:n1] = list(g)
    D.gw[:n1] = list(gw)
    D.gu[: n1 + 2] = list(gu)
    D.gwu[: n1 + 2] = list(gwu)
    return D


def _produce_on_device(mesh, D, classify, rules, gdim):
    """Two passes: per-cell classification + point counts, prefix sums, then the rules of the cut cells written
    into the ragged layout of a qr tuple.  Everything stays in HBM; the cell lists also come back to the host."""
    import ctypes as C

    import torch

    from . import _lib
    from .quadrature import RaggedQuadrature

    dev = torch.device("cuda", torch.cuda.current_device())
    ctx = _lib.ctx(dev.index)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ncells = mesh.num_cells
    cls = torch.empty(ncells, dtype=torch.int32, device=dev)
    nb = torch.empty(ncells, dtype=torch.int32, device=dev)
    ns = torch.empty(ncells, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(classify(ctx, C.byref(D), p(cls), p(nb), p(ns), stream))
    cut = torch.nonzero(cls == 1).reshape(-1).to(torch.int32)
    ncut = int(cut.shape[0])
    off_b = torch.zeros(ncut + 1, dtype=torch.int64, device=dev)
    off_s = torch.zeros(ncut + 1, dtype=torch.int64, device=dev)
    if ncut:
        off_b[1:] = torch.cumsum(nb[cut.long()].to(torch.int64), 0)
        off_s[1:] = torch.cumsum(ns[cut.long()].to(torch.int64), 0)
    Pb, Ps = int(off_b[-1]), int(off_s[-1])
    f64 = lambda n: torch.empty(max(n, 1), dtype=torch.float64, device=dev)[:n]  # noqa: E731
    pts, w, pts_b, w_b, nrm = f64(gdim * Pb), f64(Pb), f64(gdim * Ps), f64(Ps), f64(gdim * Ps)
    _lib.check(rules(ctx, C.byref(D), ncut, p(cut), p(off_b), p(off_s), p(pts), p(w), p(pts_b), p(w_b), p(nrm), stream))
    # the rules are born on the device and stay there (``q.host()`` copies them back on demand: at 256^3 the copies of
    # 25 M points into pageable memory were most of round 1's 511 ms "producer time"); only the cell lists go back
    qb = RaggedQuadrature.on_device(cut, off_b, pts, w, None, gdim)
    qs = RaggedQuadrature.on_device(cut, off_s, pts_b, w_b, nrm, gdim)
    host = lambda t: t.cpu().numpy()  # noqa: E731
    uncut = torch.nonzero(cls == 2).reshape(-1).to(torch.int32)
    outside = torch.nonzero(cls == 0).reshape(-1).to(torch.int32)
    return (host(cut), host(uncut), host(outside), qb, qs)


def generate_qr_sphere_device(mesh, radius=1.0, center=(0.0, 0.0, 0.0), order=4, axis=2):
    """``generate_qr_sphere`` on the GPU (cq_sphere_classify + cq_sphere_rules); same return value as
    ``generate_qr_circle_device``."""
    from . import _lib

    L = _lib.load()
    return _produce_on_device(mesh, sphere_desc(mesh, radius, center, order, axis), L.cq_sphere_classify, L.cq_sphere_rules, 3)


def generate_qr_circle_device(mesh, radius=1.0, center=(0.0, 0.0), order=4):
    """``generate_qr_circle`` on the GPU (cq_circle_classify + cq_circle_rules, csrc/cq_levelset.cuh): the same
    classification and rules, produced in HBM in the ragged layout the assembly kernels read.

    Returns (cut_cells, uncut_cells, outside_cells, qr_bulk, qr_bdry): int32 numpy cell lists and two
    ``RaggedQuadrature`` objects (bulk: points + weights; bdry: points + weights + unit normals) whose device copies
    are already in place -- ``assemble_*(form, [qr_bulk])`` uploads nothing."""
    from . import _lib

    L = _lib.load()
    return _produce_on_device(mesh, circle_desc(mesh, radius, center, order), L.cq_circle_classify, L.cq_circle_rules, 2)
