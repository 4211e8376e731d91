// cq_levelset.cuh -- cut-cell quadrature of a disk on a tensor quadrilateral mesh, per cell.
//
// The producer in front of the assembly path: what the reference's demo gets from Algoim
// through demo/algoim_utils.hpp:84-153 with the level set of demo/circle.hpp:14-34 (phi = |x - c|^2 - R^2):
// classification cut / uncut / outside (:98-100, :145-152), a bulk rule of cell ∩ {phi < 0}, a surface rule of
// cell ∩ {phi = 0} with unit normals, weights divided by the cell volume (:127-132), points in unit-cell
// coordinates (:134-143).  Algoim is not available; the scheme is the one of customquad_b200/levelset.py
// (dimension reduction along a per-cell height direction; Gauss points in the polar angle on every arc),
// restated here as host/device functions so that the same code runs in the CUDA kernels (cq_api.cu) and in the
// g++-compiled harness the CPU tests compare with levelset.py (tests/test_levelset_device_logic.py).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define CQ_LS_HD __host__ __device__ __forceinline__
#else
#define CQ_LS_HD inline
#endif

namespace cq {

constexpr int LS_MAX_N1 = 8;  // Gauss points per direction

struct CircleGrid {   // == cq_circle in include/cq_b200.h
  double xmin[2];     // lower-left corner of the mesh
  double h[2];        // cell size
  int64_t n[2];       // cells per direction; cell id = ix * n[1] + iy (dolfinx create_rectangle order, mesh.py)
  double center[2];
  double radius;
  int32_t n1;         // Gauss points per direction, 2 <= n1 <= LS_MAX_N1
  int32_t pad_;
  double g[LS_MAX_N1];   // Gauss-Legendre nodes on [0, 1]
  double gw[LS_MAX_N1];  // ... and weights
};

enum : int32_t { LS_OUTSIDE = 0, LS_CUT = 1, LS_UNCUT = 2 };

CQ_LS_HD void ls_cell_box(const CircleGrid& G, const int64_t cell, double lo[2], double hi[2]) {
  const int64_t ix = cell / G.n[1], iy = cell - ix * G.n[1];
  lo[0] = G.xmin[0] + (double)ix * G.h[0];
  lo[1] = G.xmin[1] + (double)iy * G.h[1];
  hi[0] = lo[0] + G.h[0];
  hi[1] = lo[1] + G.h[1];
}

// nearest / farthest point of the cell from the centre (levelset.py:generate_qr_circle)
CQ_LS_HD int32_t ls_classify_box(const CircleGrid& G, const double lo[2], const double hi[2]) {
  double rmin2 = 0.0, rmax2 = 0.0;
  for (int d = 0; d < 2; ++d) {
    const double c = G.center[d];
    const double nearest = (c < lo[d] ? lo[d] : (c > hi[d] ? hi[d] : c)) - c;
    const double fl = fabs(lo[d] - c), fh = fabs(hi[d] - c);
    const double farthest = fl > fh ? fl : fh;
    rmin2 += nearest * nearest;
    rmax2 += farthest * farthest;
  }
  const double R2 = G.radius * G.radius;
  if (rmin2 >= R2) return LS_OUTSIDE;
  if (rmax2 <= R2) return LS_UNCUT;
  return LS_CUT;
}

CQ_LS_HD void ls_sort_unique(double* v, int& n) {
  for (int i = 1; i < n; ++i) {  // insertion sort: n <= 18
    const double x = v[i];
    int j = i - 1;
    while (j >= 0 && v[j] > x) {
      v[j + 1] = v[j];
      --j;
    }
    v[j + 1] = x;
  }
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (m == 0 || v[i] != v[m - 1]) v[m++] = v[i];
  n = m;
}

// A disk ∩ rectangle problem in physical coordinates (the circle's cells, and the cross-sections of the sphere's)
struct Disk2 {
  double lo[2], hi[2], c[2];
  double R;
  int n1;
  const double* g;   // Gauss-Legendre nodes / weights on [0, 1]
  const double* gw;
};

// Bulk rule of [lo, hi] ∩ disk: emit(index, x, y, weight) in physical coordinates / measure, in the order of
// levelset.py:_bulk_rule_cell.  Returns the number of points; *wsum = sum of the weights.
template <class Emit>
CQ_LS_HD int ls_bulk_core(const Disk2& D, Emit&& emit, double* wsum) {
  const double R = D.R, R2 = R * R;
  const double* c = D.c;
  const double* lo = D.lo;
  const double* hi = D.hi;
  const double cx0 = 0.5 * (lo[0] + hi[0]) - c[0], cx1 = 0.5 * (lo[1] + hi[1]) - c[1];
  const int hdir = fabs(cx1) >= fabs(cx0) ? 1 : 0, adir = 1 - hdir;
  const double a0 = lo[adir] - c[adir], a1 = hi[adir] - c[adir];
  const double t0 = lo[hdir] - c[hdir], t1 = hi[hdir] - c[hdir];
  double brk[8];
  int nb = 0;
  brk[nb++] = a0;
  brk[nb++] = a1;
  for (int e = 0; e < 2; ++e) {  // crossings of the circle with the two edges t = const
    const double tt = e == 0 ? t0 : t1;
    if (fabs(tt) < R) {
      const double s = sqrt(R2 - tt * tt);
      if (a0 < -s && -s < a1) brk[nb++] = -s;
      if (a0 < s && s < a1) brk[nb++] = s;
    }
  }
  if (a0 < -R && -R < a1) brk[nb++] = -R;  // the disk ends inside the abscissa range
  if (a0 < R && R < a1) brk[nb++] = R;
  ls_sort_unique(brk, nb);
  const int n1 = D.n1;
  int np = 0;
  double sum = 0.0;
  for (int k = 0; k + 1 < nb; ++k) {
    const double b0 = brk[k], b1 = brk[k + 1];
    const double sc = fabs(b0) > 1.0 ? fabs(b0) : 1.0;
    if (b1 - b0 < 1e-14 * sc) continue;
    for (int i = 0; i < n1; ++i) {
      const double a = b0 + (b1 - b0) * D.g[i];
      const double wa = (b1 - b0) * D.gw[i];
      const double r2 = R2 - a * a;
      const double half = sqrt(r2 > 0.0 ? r2 : 0.0);
      const double tl = t0 > -half ? t0 : -half, th = t1 < half ? t1 : half;
      if (!(th > tl)) continue;
      for (int j = 0; j < n1; ++j) {
        const double t = tl + (th - tl) * D.g[j];
        const double wt = wa * (th - tl) * D.gw[j];
        sum += wt;
        double p[2];
        p[adir] = a + c[adir];
        p[hdir] = t + c[hdir];
        emit(np, p[0], p[1], wt);
        ++np;
      }
    }
  }
  *wsum = sum;
  return np;
}

CQ_LS_HD Disk2 ls_disk_of(const CircleGrid& G, const double lo[2], const double hi[2]) {
  Disk2 D;
  for (int d = 0; d < 2; ++d) {
    D.lo[d] = lo[d];
    D.hi[d] = hi[d];
    D.c[d] = G.center[d];
  }
  D.R = G.radius;
  D.n1 = G.n1;
  D.g = G.g;
  D.gw = G.gw;
  return D;
}

// Bulk rule of a cell of the circle problem.  pts / w may be null (count and weight sum only).  Points are written in
// unit-cell coordinates, weights relative to the cell volume.  *wsum = sum of the weights (physical measure).
CQ_LS_HD int ls_bulk_rule(const CircleGrid& G, const double lo[2], const double hi[2], double* pts, double* w, double* wsum) {
  const Disk2 D = ls_disk_of(G, lo, hi);
  const double inv_vol = 1.0 / (G.h[0] * G.h[1]);
  return ls_bulk_core(
      D,
      [&](const int k, const double x, const double y, const double wt) {
        if (pts) {
          pts[2 * k] = (x - lo[0]) / G.h[0];
          pts[2 * k + 1] = (y - lo[1]) / G.h[1];
          w[k] = wt * inv_vol;
        }
      },
      wsum);
}

CQ_LS_HD bool ls_inside(const double lo[2], const double hi[2], const double x, const double y) {
  return x >= lo[0] - 1e-13 && y >= lo[1] - 1e-13 && x <= hi[0] + 1e-13 && y <= hi[1] + 1e-13;
}

// Surface rule of [lo, hi] ∩ circle: Gauss points in the polar angle on every arc inside the cell;
// emit(index, x, y, arc-length weight, nx, ny) in physical coordinates, order of levelset.py:_surface_rule_cell.
// With emit_points == false only the count is formed.  Returns the number of points.
template <class Emit>
CQ_LS_HD int ls_surface_core(const Disk2& D, const bool emit_points, Emit&& emit) {
  const double R = D.R, R2 = R * R;
  const double* c = D.c;
  const double* lo = D.lo;
  const double* hi = D.hi;
  const double PI = 3.141592653589793;
  double ang[8];
  int na = 0;
  for (int d = 0; d < 2; ++d)
    for (int e = 0; e < 2; ++e) {
      const double edge = (e == 0 ? lo[d] : hi[d]) - c[d];
      if (fabs(edge) < R) {
        const double other = sqrt(R2 - edge * edge);
        for (int sgn = 0; sgn < 2; ++sgn) {
          const double o = sgn == 0 ? -other : other;
          if (lo[1 - d] - c[1 - d] - 1e-14 <= o && o <= hi[1 - d] - c[1 - d] + 1e-14) {
            double xy[2];
            xy[d] = edge;
            xy[1 - d] = o;
            ang[na++] = atan2(xy[1], xy[0]);
          }
        }
      }
    }
  double arc0[8], arc1[8];
  int narc = 0;
  if (na < 2) {
    // no crossing: the whole circle is inside the cell, or outside
    if (ls_inside(lo, hi, c[0] + R, c[1]) && ls_inside(lo, hi, c[0] - R, c[1]) && ls_inside(lo, hi, c[0], c[1] + R) &&
        ls_inside(lo, hi, c[0], c[1] - R)) {
      arc0[0] = 0.0;
      arc1[0] = 2.0 * PI;
      narc = 1;
    }
  } else {
    for (int i = 0; i < na; ++i) ang[i] = rint(ang[i] * 1e14) / 1e14;  // numpy.round(., 14)
    ls_sort_unique(ang, na);
    for (int k = 0; k < na; ++k) {
      const double th0 = ang[k];
      const double th1 = ang[(k + 1) % na] + (k == na - 1 ? 2.0 * PI : 0.0);
      if (th1 - th0 < 1e-13) continue;
      const double mid = 0.5 * (th0 + th1);
      if (ls_inside(lo, hi, c[0] + R * cos(mid), c[1] + R * sin(mid))) {
        arc0[narc] = th0;
        arc1[narc] = th1;
        ++narc;
      }
    }
  }
  const int n1 = D.n1;
  int np = 0;
  for (int a = 0; a < narc; ++a) {
    const double a0 = arc0[a], a1 = arc1[a];
    int pieces = (int)ceil((a1 - a0) / (0.5 * PI));  // Gauss on arcs of at most a quarter turn
    if (pieces < 1) pieces = 1;
    for (int j = 0; j < pieces; ++j) {
      const double th0 = a0 + (a1 - a0) * j / pieces, th1 = a0 + (a1 - a0) * (j + 1) / pieces;
      for (int i = 0; i < n1; ++i) {
        if (emit_points) {
          const double th = th0 + (th1 - th0) * D.g[i];
          const double nx = cos(th), ny = sin(th);
          emit(np, c[0] + R * nx, c[1] + R * ny, R * (th1 - th0) * D.gw[i], nx, ny);
        }
        ++np;
      }
    }
  }
  return np;
}

// Surface rule of a cell of the circle problem: unit-cell points, weights = arc length / cell volume, unit outward
// normals.  pts / w / nrm may be null (count only).
CQ_LS_HD int ls_surface_rule(const CircleGrid& G, const double lo[2], const double hi[2], double* pts, double* w, double* nrm) {
  const Disk2 D = ls_disk_of(G, lo, hi);
  const double inv_vol = 1.0 / (G.h[0] * G.h[1]);
  return ls_surface_core(D, pts != nullptr, [&](const int k, const double x, const double y, const double wt, const double nx, const double ny) {
    pts[2 * k] = (x - lo[0]) / G.h[0];
    pts[2 * k + 1] = (y - lo[1]) / G.h[1];
    w[k] = wt * inv_vol;
    nrm[2 * k] = nx;
    nrm[2 * k + 1] = ny;
  });
}

// Classification with the producer's last rule (levelset.py: a cut cell whose bulk rule is empty -- the circle only
// touches it -- counts as outside) and the point counts of a cut cell.
CQ_LS_HD int32_t ls_classify_count(const CircleGrid& G, const int64_t cell, int32_t* nbulk, int32_t* nsurf) {
  double lo[2], hi[2];
  ls_cell_box(G, cell, lo, hi);
  int32_t cls = ls_classify_box(G, lo, hi);
  *nbulk = *nsurf = 0;
  if (cls == LS_CUT) {
    double wsum;
    const int nb = ls_bulk_rule(G, lo, hi, nullptr, nullptr, &wsum);
    if (nb == 0 || wsum < 1e-14 * (G.h[0] * G.h[1])) return LS_OUTSIDE;
    *nbulk = nb;
    *nsurf = ls_surface_rule(G, lo, hi, nullptr, nullptr, nullptr);
  }
  return cls;
}

// ---------------------------------------------------------------------------------------------------------------
// Ball on a tensor hexahedral mesh (demo/sphere.hpp:14-34), scheme of levelset.py:_sphere_rules_cell: the ball is
// sliced along `axis`; every cross-section of ball ∩ cell is a disk ∩ rectangle (the 2-D rules above, radius
// sqrt(R^2 - u^2)); the slicing variable u is integrated with Gauss points between the heights where the
// cross-section changes type, after the smooth-step change of variable u = a + (b - a) x^2 (3 - 2x); the surface
// rule uses dS = R dtheta du.
// ---------------------------------------------------------------------------------------------------------------
constexpr int LS_MAX_NU = LS_MAX_N1 + 2;

struct SphereGrid {   // == cq_sphere in include/cq_b200.h
  double xmin[3];
  double h[3];
  int64_t n[3];       // cell id = (iz * n[1] + iy) * n[0] + ix (dolfinx create_box order, mesh.py)
  double center[3];
  double radius;
  int32_t n1;         // Gauss points per direction in the cross-sections, 2 <= n1 <= LS_MAX_N1
  int32_t axis;       // slicing direction
  double g[LS_MAX_N1], gw[LS_MAX_N1];    // Gauss-Legendre on [0, 1], n1 points
  double gu[LS_MAX_NU], gwu[LS_MAX_NU];  // ... n1 + 2 points (slicing variable)
};

CQ_LS_HD void ls_cell_box3(const SphereGrid& G, const int64_t cell, double lo[3], double hi[3]) {
  const int64_t ix = cell % G.n[0], r = cell / G.n[0];
  const int64_t iy = r % G.n[1], iz = r / G.n[1];
  const int64_t id[3] = {ix, iy, iz};
  for (int d = 0; d < 3; ++d) {
    lo[d] = G.xmin[d] + (double)id[d] * G.h[d];
    hi[d] = lo[d] + G.h[d];
  }
}

CQ_LS_HD int32_t ls_classify_box3(const SphereGrid& G, const double lo[3], const double hi[3]) {
  double rmin2 = 0.0, rmax2 = 0.0;
  for (int d = 0; d < 3; ++d) {
    const double c = G.center[d];
    const double nearest = (c < lo[d] ? lo[d] : (c > hi[d] ? hi[d] : c)) - c;
    const double fl = fabs(lo[d] - c), fh = fabs(hi[d] - c);
    const double farthest = fl > fh ? fl : fh;
    rmin2 += nearest * nearest;
    rmax2 += farthest * farthest;
  }
  const double R2 = G.radius * G.radius;
  if (rmin2 >= R2) return LS_OUTSIDE;
  if (rmax2 <= R2) return LS_UNCUT;
  return LS_CUT;
}

// Bulk and surface rules of one cell.  Output pointers may be null (counts and bulk weight sum only).  Points in
// unit-cell coordinates, weights relative to the cell volume, unit outward normals.
CQ_LS_HD void ls_sphere_rules(const SphereGrid& G, const double lo[3], const double hi[3], double* pts, double* w, double* pts_b,
                              double* w_b, double* nrm, int* nbulk, int* nsurf, double* wsum) {
  const double R = G.radius, R2 = R * R;
  const int axis = G.axis;
  const int s0 = axis == 0 ? 1 : 0, s1 = axis == 2 ? 1 : 2;
  *nbulk = *nsurf = 0;
  *wsum = 0.0;
  const double ua = lo[axis] - G.center[axis], ub = hi[axis] - G.center[axis];
  const double u0 = ua > -R ? ua : -R, u1 = ub < R ? ub : R;
  if (u1 <= u0) return;
  Disk2 D;
  D.lo[0] = lo[s0]; D.lo[1] = lo[s1];
  D.hi[0] = hi[s0]; D.hi[1] = hi[s1];
  D.c[0] = G.center[s0]; D.c[1] = G.center[s1];
  D.n1 = G.n1;
  D.g = G.g;
  D.gw = G.gw;
  // heights where the cross-section circle passes a corner of the rectangle or touches an edge line
  const double rel[2][2] = {{D.lo[0] - D.c[0], D.hi[0] - D.c[0]}, {D.lo[1] - D.c[1], D.hi[1] - D.c[1]}};
  double brk[18];
  int nb = 0;
  brk[nb++] = u0;
  brk[nb++] = u1;
  auto add = [&](const double v) {
    if (u0 < v && v < u1) brk[nb++] = v;
  };
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      const double d2 = R2 - rel[0][i] * rel[0][i] - rel[1][j] * rel[1][j];
      if (d2 > 0.0) {
        add(-sqrt(d2));
        add(sqrt(d2));
      }
    }
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < 2; ++i) {
      const double d2 = R2 - rel[k][i] * rel[k][i];
      if (d2 > 0.0 && rel[1 - k][0] <= 0.0 && 0.0 <= rel[1 - k][1]) {
        add(-sqrt(d2));
        add(sqrt(d2));
      }
    }
  ls_sort_unique(brk, nb);
  const int nu = G.n1 + 2;
  const double inv_vol = 1.0 / (G.h[0] * G.h[1] * G.h[2]);
  int npb = 0, nps = 0;
  double sum = 0.0;
  for (int k = 0; k + 1 < nb; ++k) {
    const double a = brk[k], b = brk[k + 1];
    const double sc = fabs(a) > 1.0 ? fabs(a) : 1.0;
    if (b - a < 1e-13 * sc) continue;
    for (int i = 0; i < nu; ++i) {
      const double x = G.gu[i];
      const double u = a + (b - a) * x * x * (3.0 - 2.0 * x);
      const double wu = (b - a) * 6.0 * x * (1.0 - x) * G.gwu[i];
      const double rho2 = R2 - u * u;
      if (rho2 <= 0.0) continue;
      const double rho = sqrt(rho2);
      D.R = rho;
      const double pu = u + G.center[axis];
      double ws2;
      const int base_b = npb;
      npb += ls_bulk_core(
          D,
          [&](const int q, const double x0, const double x1, const double wt) {
            if (pts) {
              double p[3];
              p[s0] = x0; p[s1] = x1; p[axis] = pu;
              double* o = pts + 3 * (base_b + q);
              for (int d = 0; d < 3; ++d) o[d] = (p[d] - lo[d]) / G.h[d];
              w[base_b + q] = wt * wu * inv_vol;
            }
          },
          &ws2);
      sum += ws2 * wu;
      const int base_s = nps;
      nps += ls_surface_core(D, pts_b != nullptr, [&](const int q, const double x0, const double x1, const double wt, const double, const double) {
        double p[3];
        p[s0] = x0; p[s1] = x1; p[axis] = pu;
        double* o = pts_b + 3 * (base_s + q);
        double* nn = nrm + 3 * (base_s + q);
        for (int d = 0; d < 3; ++d) {
          o[d] = (p[d] - lo[d]) / G.h[d];
          nn[d] = (p[d] - G.center[d]) / R;
        }
        w_b[base_s + q] = wt / rho * R * wu * inv_vol;  // arc length rho dtheta -> dtheta; dS = R dtheta du
      });
    }
  }
  *nbulk = npb;
  *nsurf = nps;
  *wsum = sum;
}

CQ_LS_HD int32_t ls_classify_count3(const SphereGrid& G, const int64_t cell, int32_t* nbulk, int32_t* nsurf) {
  double lo[3], hi[3];
  ls_cell_box3(G, cell, lo, hi);
  const int32_t cls = ls_classify_box3(G, lo, hi);
  *nbulk = *nsurf = 0;
  if (cls == LS_CUT) {
    int nb, ns;
    double wsum;
    ls_sphere_rules(G, lo, hi, nullptr, nullptr, nullptr, nullptr, nullptr, &nb, &ns, &wsum);
    if (nb == 0 || wsum < 1e-14 * (G.h[0] * G.h[1] * G.h[2])) return LS_OUTSIDE;
    *nbulk = nb;
    *nsurf = ns;
  }
  return cls;
}

}  // namespace cq
