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
    if len(ang) < 2:
        return np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2))
    ang = np.unique(np.round(np.sort(np.asarray(ang)), 14))
    g, gw = _gauss01(n1)
    P, W, Nn = [], [], []
    m = len(ang)
    for k in range(m):
        th0 = ang[k]
        th1 = ang[(k + 1) % m] + (2 * np.pi if k == m - 1 else 0.0)
        mid = 0.5 * (th0 + th1)
        pm = c + R * np.array([np.cos(mid), np.sin(mid)])
        if np.all(pm >= lo - 1e-13) and np.all(pm <= hi + 1e-13) and th1 - th0 < np.pi:  # arc inside the cell
            th = th0 + (th1 - th0) * g
            n = np.stack([np.cos(th), np.sin(th)], axis=1)
            P.append(c + R * n)
            W.append(R * (th1 - th0) * gw)
            Nn.append(n)
    if not P:
        return np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2))
    return np.concatenate(P), np.concatenate(W), np.concatenate(Nn)


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
