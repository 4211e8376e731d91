// Host build (g++) of customquad_b200/csrc/cq_levelset.cuh: the per-cell functions the CUDA kernels call, exposed
// to ctypes so that the CPU suite can compare them with customquad_b200/levelset.py (test_levelset_device_logic.py).
#include "../customquad_b200/csrc/cq_levelset.cuh"

extern "C" {
int ls_sizeof_grid(void) { return (int)sizeof(cq::CircleGrid); }
// cls / nbulk / nsurf for every cell
void ls_classify_all(const cq::CircleGrid* G, int32_t* cls, int32_t* nbulk, int32_t* nsurf) {
  const int64_t n = G->n[0] * G->n[1];
  for (int64_t c = 0; c < n; ++c) cls[c] = cq::ls_classify_count(*G, c, nbulk + c, nsurf + c);
}
// the rules of one cell (buffers sized from ls_classify_all)
void ls_rules_cell(const cq::CircleGrid* G, int64_t cell, double* pts, double* w, double* pts_b, double* w_b, double* nrm) {
  double lo[2], hi[2], wsum;
  cq::ls_cell_box(*G, cell, lo, hi);
  cq::ls_bulk_rule(*G, lo, hi, pts, w, &wsum);
  cq::ls_surface_rule(*G, lo, hi, pts_b, w_b, nrm);
}

int ls_sizeof_grid3(void) { return (int)sizeof(cq::SphereGrid); }
void ls_classify_all3(const cq::SphereGrid* G, int32_t* cls, int32_t* nbulk, int32_t* nsurf) {
  const int64_t n = G->n[0] * G->n[1] * G->n[2];
  for (int64_t c = 0; c < n; ++c) cls[c] = cq::ls_classify_count3(*G, c, nbulk + c, nsurf + c);
}
void ls_rules_cell3(const cq::SphereGrid* G, int64_t cell, double* pts, double* w, double* pts_b, double* w_b, double* nrm) {
  double lo[3], hi[3], wsum;
  int nb, ns;
  cq::ls_cell_box3(*G, cell, lo, hi);
  cq::ls_sphere_rules(*G, lo, hi, pts, w, pts_b, w_b, nrm, &nb, &ns, &wsum);
}
}
