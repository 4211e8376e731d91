"""Cut-cell quadrature producers beyond the circle and the sphere: the reference obtains its
rules from Algoim for ANY level set handed to ``generate_qr`` (demo/algoim_utils.py:92-180, demo/algoim_utils.hpp:84-153)
and ships three of them: demo/circle.hpp, demo/sphere.hpp and demo/square.hpp.  Algoim is absent here; ``levelset.py``
covers the first two in closed form, this module adds

* ``generate_qr_box``: the square of demo/square.hpp:19-23 (phi = max(|x| - 1, |y| - 1)) and, more generally, any
  axis-aligned box in 2-D or 3-D.  cell ∩ box is again a box and box ∩ cell-interior faces are rectangles, so the
  tensor Gauss rules are EXACT for polynomials (what Algoim also returns for a piecewise-linear phi); normals are
  the unit axis vectors (square.hpp:36-46 returns the same);
* ``generate_qr_levelset``: a smooth level set given as Python callables ``phi(x)`` / ``grad(x)`` on 2-D tensor meshes:
  dimension reduction along a height direction in which phi is monotone inside the cell (Saye's construction, the idea
  behind algoim::quadGen), with numerical roots instead of closed forms and recursive subdivision of the cell where no
  coordinate direction qualifies.  Same outputs and conventions as ``levelset.generate_qr_circle``.

3-D level sets other than the sphere and the box are not covered (Algoim recurses over the dimension with the
restricted functions phi(., face); out of scope here).  Host numpy code: it runs once per mesh before the
timed path, like the reference's generator."""
from __future__ import annotations

import numpy as np

from .levelset import _gauss01


def _cells_lo(mesh):
    """Lower corners [num_cells, dim] in the cell numbering of mesh.py (2-D: ix * ny + iy; 3-D: (iz * ny + iy) * nx + ix)."""
    h = mesh.h()
    if mesh.topology.dim == 2:
        nx, ny = mesh.shape
        ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        idx = [ix.reshape(-1), iy.reshape(-1)]
    else:
        nx, ny, nz = mesh.shape
        iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        idx = [ix.reshape(-1), iy.reshape(-1), iz.reshape(-1)]
    return np.stack([mesh.xmin[d] + idx[d] * h[d] for d in range(mesh.topology.dim)], axis=1)


# ------------------------------------------------------------------------------------------------
# axis-aligned box (demo/square.hpp)
# ------------------------------------------------------------------------------------------------
def generate_qr_box(mesh, box_lo=None, box_hi=None, order=4):
    """``generate_qr(mesh, NN, degree, "square", opts)`` (demo/algoim_utils.py:92-180 with demo/square.hpp) for the box
    [box_lo, box_hi] (default [-1, 1]^d, the reference's square).  Returns (cut_cells, uncut_cells, outside_cells,
    qr_pts, qr_w, qr_pts_bdry, qr_w_bdry, qr_n) like ``levelset.generate_qr_circle``.

    A box face that coincides with a mesh plane would leave the surface integral without a cell (the cells on its
    inside are uncut); such a placement is refused."""
    d = mesh.topology.dim
    bl = -np.ones(d) if box_lo is None else np.asarray(box_lo, dtype=np.float64)
    bh = np.ones(d) if box_hi is None else np.asarray(box_hi, dtype=np.float64)
    h = mesh.h()
    for k in range(d):
        for f in (bl[k], bh[k]):
            t = (f - mesh.xmin[k]) / h[k]
            if -1e-9 < t < mesh.shape[k] + 1e-9:
                assert abs(t - round(t)) > 1e-9, f"box face x_{k} = {f} lies on a mesh plane"
    lo = _cells_lo(mesh)
    hi = lo + h[None, :]
    ilo, ihi = np.maximum(lo, bl[None, :]), np.minimum(hi, bh[None, :])  # cell ∩ box
    overlap = np.all(ihi > ilo, axis=1)
    uncut = np.all((lo >= bl[None, :]) & (hi <= bh[None, :]), axis=1)
    cut_cells = np.nonzero(overlap & ~uncut)[0].astype(np.int32)
    n1 = max(2, (order + 2) // 2)
    g, gw = _gauss01(n1)
    vol = float(np.prod(h))
    grids = np.meshgrid(*([g] * d), indexing="ij")
    G = np.stack([x.reshape(-1) for x in grids], axis=1)  # [n1^d, d]
    GW = np.prod(np.stack([x.reshape(-1) for x in np.meshgrid(*([gw] * d), indexing="ij")], axis=1), axis=1)
    if d > 1:
        Gf = np.stack([x.reshape(-1) for x in np.meshgrid(*([g] * (d - 1)), indexing="ij")], axis=1)
        GWf = np.prod(np.stack([x.reshape(-1) for x in np.meshgrid(*([gw] * (d - 1)), indexing="ij")], axis=1), axis=1)
    qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n = [], [], [], [], []
    for c in cut_cells:
        a, b = ilo[c], ihi[c]
        p = a[None, :] + (b - a)[None, :] * G
        qr_pts.append(((p - lo[c]) / h).reshape(-1))
        qr_w.append(GW * np.prod(b - a) / vol)
        Pb, Wb, Nb = [], [], []
        for k in range(d):
            others = [j for j in range(d) if j != k]
            for f, sgn in ((bl[k], -1.0), (bh[k], 1.0)):
                if lo[c, k] < f < hi[c, k]:  # this face of the box crosses the cell: the rectangle (cell ∩ box) at x_k = f
                    q = np.empty((Gf.shape[0], d))
                    q[:, k] = f
                    q[:, others] = a[others][None, :] + (b - a)[others][None, :] * Gf
                    Pb.append(q)
                    Wb.append(GWf * np.prod((b - a)[others]))
                    n = np.zeros((Gf.shape[0], d))
                    n[:, k] = sgn
                    Nb.append(n)
        assert Pb, "a cut cell of a box holds a piece of one of its faces"
        pb = np.concatenate(Pb)
        qr_pts_b.append(((pb - lo[c]) / h).reshape(-1))
        qr_w_b.append(np.concatenate(Wb) / vol)
        qr_n.append(np.concatenate(Nb).reshape(-1))
    return (cut_cells, np.nonzero(uncut)[0].astype(np.int32), np.nonzero(~overlap)[0].astype(np.int32),
            qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n)  # fmt: skip


# ------------------------------------------------------------------------------------------------
# smooth level sets in 2-D
# ------------------------------------------------------------------------------------------------
def _roots_1d(f, a, b, samples=16):
    """All sign changes of the vectorised scalar function f on [a, b] found on a uniform sampling, refined by bisection."""
    x = np.linspace(a, b, samples + 1)
    v = f(x)
    out = []
    for i in range(samples):
        if v[i] == 0.0:
            out.append(x[i])
        elif v[i] * v[i + 1] < 0.0:
            l, r, fl = x[i], x[i + 1], v[i]
            for _ in range(60):
                m = 0.5 * (l + r)
                fm = float(f(np.array([m]))[0])
                if fl * fm <= 0.0:
                    r = m
                else:
                    l, fl = m, fm
            out.append(0.5 * (l + r))
    return out


def _bisect_vec(f, l, r, fl, iters=60):
    """Roots of the monotone functions f_i on [l_i, r_i] (f vectorised over i; fl = f(l))."""
    l, r, fl = l.copy(), r.copy(), fl.copy()
    for _ in range(iters):
        m = 0.5 * (l + r)
        fm = f(m)
        left = fl * fm <= 0.0
        r = np.where(left, m, r)
        l = np.where(left, l, m)
        fl = np.where(left, fl, fm)
    return 0.5 * (l + r)


def _height_rule(phi, grad, lo, hi, k, n1):
    """Rules of the cell [lo, hi] in which phi is monotone in x_k: returns (P, W), (Pb, Wb, Nb) in physical units."""
    a_dir = 1 - k
    a0, a1, t0, t1 = lo[a_dir], hi[a_dir], lo[k], hi[k]

    def at(a, t):
        x = np.empty((np.size(a), 2))
        x[:, a_dir], x[:, k] = a, t
        return x

    brk = {a0, a1}
    for tt in (t0, t1):  # where the interface leaves through the faces t = const
        for r in _roots_1d(lambda a, tt=tt: phi(at(a, np.full(np.size(a), tt))), a0, a1):
            if a0 < r < a1:
                brk.add(float(r))
    brk = sorted(brk)
    g, gw = _gauss01(n1)
    P, W, Pb, Wb, Nb = [], [], [], [], []
    for b0, b1 in zip(brk[:-1], brk[1:]):
        if b1 - b0 < 1e-13 * max(1.0, abs(b0), abs(b1)):
            continue
        a = b0 + (b1 - b0) * g
        wa = (b1 - b0) * gw
        f0, f1 = phi(at(a, np.full(n1, t0))), phi(at(a, np.full(n1, t1)))
        cross = f0 * f1 < 0.0
        ts = np.where(cross, 0.0, t0)
        if cross.any():
            idx = np.nonzero(cross)[0]
            ts[idx] = _bisect_vec(lambda t, idx=idx: phi(at(a[idx], t)), np.full(idx.size, t0), np.full(idx.size, t1), f0[idx])
        # inside interval in t for every abscissa
        tl = np.where(cross, np.where(f0 < 0.0, t0, ts), t0)
        th = np.where(cross, np.where(f0 < 0.0, ts, t1), np.where(f0 < 0.0, t1, t0))
        ok = th > tl
        for ai, wi, l, u in zip(a[ok], wa[ok], tl[ok], th[ok]):
            P.append(at(np.full(n1, ai), l + (u - l) * g))
            W.append(wi * (u - l) * gw)
        if cross.any():
            xb = at(a[cross], ts[cross])
            gb = grad(xb)
            nrm = np.linalg.norm(gb, axis=1)
            Pb.append(xb)
            Wb.append(wa[cross] * nrm / np.abs(gb[:, k]))  # ds = sqrt(1 + t'(a)^2) da = |grad phi| / |d_k phi| da
            Nb.append(gb / nrm[:, None])
    cat = lambda L, shape: np.concatenate(L) if L else np.zeros(shape)  # noqa: E731
    return (cat(P, (0, 2)), cat(W, 0)), (cat(Pb, (0, 2)), cat(Wb, 0), cat(Nb, (0, 2)))


def _lattice(lo, hi, m):
    s = np.linspace(0.0, 1.0, m)
    X, Y = np.meshgrid(lo[0] + (hi[0] - lo[0]) * s, lo[1] + (hi[1] - lo[1]) * s, indexing="ij")
    return np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)


def _rules_2d(phi, grad, lo, hi, n1, depth, max_depth):
    """Recursive: a height direction if one qualifies, otherwise four sub-cells."""
    L = _lattice(lo, hi, 5)
    v = phi(L)
    if np.all(v < 0.0):  # full (sub-)cell: tensor Gauss rule
        g, gw = _gauss01(n1)
        X, Y = np.meshgrid(g, g, indexing="ij")
        p = lo[None, :] + (hi - lo)[None, :] * np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
        w = np.outer(gw, gw).reshape(-1) * np.prod(hi - lo)
        return (p, w), (np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2)))
    if np.all(v > 0.0):
        return (np.zeros((0, 2)), np.zeros(0)), (np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2)))
    G = grad(L)
    nrm = np.linalg.norm(G, axis=1)
    best, k = -1.0, -1
    for kk in (0, 1):
        gk = G[:, kk]
        if np.all(gk > 0.0) or np.all(gk < 0.0):
            q = float(np.min(np.abs(gk) / np.maximum(nrm, 1e-300)))  # how far the interface stays from vertical
            if q > best:
                best, k = q, kk
    if k >= 0 and (best > 0.35 or depth >= max_depth):
        return _height_rule(phi, grad, lo, hi, k, n1)
    if depth >= max_depth:  # no monotone direction even at the finest level: the better of the two, as is
        return _height_rule(phi, grad, lo, hi, int(np.argmax(np.mean(np.abs(G), axis=0))), n1)
    mid = 0.5 * (lo + hi)
    parts = []
    for sx in (0, 1):
        for sy in (0, 1):
            l = np.array([lo[0] if sx == 0 else mid[0], lo[1] if sy == 0 else mid[1]])
            u = np.array([mid[0] if sx == 0 else hi[0], mid[1] if sy == 0 else hi[1]])
            parts.append(_rules_2d(phi, grad, l, u, n1, depth + 1, max_depth))
    cat = lambda xs: np.concatenate(xs)  # noqa: E731
    return ((cat([p[0][0] for p in parts]), cat([p[0][1] for p in parts])),
            (cat([p[1][0] for p in parts]), cat([p[1][1] for p in parts]), cat([p[1][2] for p in parts])))  # fmt: skip


def generate_qr_levelset(mesh, phi, grad, order=4, max_depth=3, classify_samples=5):
    """``generate_qr`` for a smooth level set on a 2-D tensor mesh: the domain is {phi < 0}.

    ``phi(x)`` -> [n] and ``grad(x)`` -> [n, 2] take points x [n, 2] (vectorised).  A cell is cut when phi changes sign on
    a ``classify_samples``^2 lattice of the cell (the reference decides by Algoim's interval bounds; a feature smaller
    than the lattice spacing can be missed).  Normals are grad phi / |grad phi| (unit; the demo passes the un-normalised
    gradient, a FIXME in algoim_utils.hpp:107-114)."""
    assert mesh.topology.dim == 2, "smooth level sets: 2-D meshes (3-D: sphere and box only, levelset.py / generate_qr_box)"
    h = mesh.h()
    lo = _cells_lo(mesh)
    s = np.linspace(0.0, 1.0, classify_samples)
    SX, SY = np.meshgrid(s, s, indexing="ij")
    S = np.stack([SX.reshape(-1), SY.reshape(-1)], axis=1)  # [m, 2]
    C = lo.shape[0]
    vals = phi((lo[:, None, :] + S[None, :, :] * h[None, None, :]).reshape(-1, 2)).reshape(C, -1)
    uncut = np.all(vals < 0.0, axis=1)
    outside = np.all(vals >= 0.0, axis=1)
    vol = float(np.prod(h))
    n1 = max(2, (order + 2) // 2)
    keep, qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n = [], [], [], [], [], []
    for c in np.nonzero(~(uncut | outside))[0]:
        (p, w), (pb, wb, nb) = _rules_2d(phi, grad, lo[c], lo[c] + h, n1, 0, max_depth)
        if w.size == 0 or w.sum() < 1e-14 * vol or wb.size == 0:  # the interface only touches the cell
            if w.size and w.sum() > 0.5 * vol:
                uncut[c] = True
            else:
                outside[c] = True
            continue
        keep.append(c)
        qr_pts.append(((p - lo[c]) / h).reshape(-1))
        qr_w.append(w / vol)
        qr_pts_b.append(((pb - lo[c]) / h).reshape(-1))
        qr_w_b.append(wb / vol)
        qr_n.append(nb.reshape(-1))
    return (np.asarray(keep, dtype=np.int32), np.nonzero(uncut)[0].astype(np.int32), np.nonzero(outside)[0].astype(np.int32),
            qr_pts, qr_w, qr_pts_b, qr_w_b, qr_n)  # fmt: skip
