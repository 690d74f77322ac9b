// nms3d_pair.cuh -- the two "Qhull" stages of the 3D NMS cascade as plain single-thread routines
// (host + device).  The CUDA kernels in nms3d.cu distribute the per-plane loop over a warp; these
// serial versions define the arithmetic and are what tests/hostcheck compares with oracle/_ref.
// See geom3d.cuh (Q) for the relation to the reference's qhull_overlap_kernel /
// qhull_overlap_convex_hulls (stardist3d_impl.cpp:830-939).
#pragma once
#include "geom3d.cuh"

namespace sd3 {

constexpr int SD3_MAX_RAYS = 256;
constexpr int SD3_MAX_FACES = 512;     // 2*n_rays - 4

struct PlaneArray {
  const Plane* p;
  SD3_HD Plane operator()(int i) const { return p[i]; }
};

SD3_HD inline double extent_bound(const float* pv1, const float* pv2, int n_rays, const double* p) {
  double m = 0;
  for (int i = 0; i < 3 * n_rays; ++i) {
    const double a = fabs((double)pv1[i] - p[i % 3]), b = fabs((double)pv2[i] - p[i % 3]);
    m = a > m ? a : m; m = b > m ? b : m;
  }
  return 4.0 * m + 1.0;
}

// stage 3: volume of kernel(P1) ∩ kernel(P2); 0 when the centre midpoint is not strictly inside
SD3_HD inline float overlap_kernel_volume(const float* pv1, const float* c1, const float* pv2, const float* c2,
                                          const int* faces, int n_rays, int n_faces, Plane* planes /* 2*n_faces */) {
  for (int i = 0; i < n_faces; ++i) {
    double hs[4];
    build_halfspace(&pv1[3 * faces[3 * i]], &pv1[3 * faces[3 * i + 1]], &pv1[3 * faces[3 * i + 2]], hs);
    planes[2 * i].n0 = hs[0]; planes[2 * i].n1 = hs[1]; planes[2 * i].n2 = hs[2]; planes[2 * i].d = hs[3];
    build_halfspace(&pv2[3 * faces[3 * i]], &pv2[3 * faces[3 * i + 1]], &pv2[3 * faces[3 * i + 2]], hs);
    planes[2 * i + 1].n0 = hs[0]; planes[2 * i + 1].n1 = hs[1]; planes[2 * i + 1].n2 = hs[2]; planes[2 * i + 1].d = hs[3];
  }
  // interior_point[k] = .5*(center1[k]+center2[k]) : float add, double multiply (:856-858)
  double p[3];
  for (int k = 0; k < 3; ++k) p[k] = .5 * (double)(c1[k] + c2[k]);
  const int n = 2 * n_faces;
  for (int i = 0; i < n; ++i) if (!plane_feasible(planes[i], p)) return 0.f;
  const double L = extent_bound(pv1, pv2, n_rays, p);
  PlaneArray PA{planes};
  double vol = 0; int ovf = 0;
  for (int k = 0; k < n; ++k) vol += face_cone_volume(PA, n, k, p, L, &ovf);
  return (float)vol;
}
SD3_HD inline float overlap_kernel_volume(const float* pv1, const float* c1, const float* pv2, const float* c2,
                                          const int* faces, int n_rays, int n_faces) {
  Plane planes[2 * SD3_MAX_FACES];
  if (n_faces > SD3_MAX_FACES) return 0.f;
  return overlap_kernel_volume(pv1, c1, pv2, c2, faces, n_rays, n_faces, planes);
}

// stage 4: volume of hull(P1) ∩ hull(P2); 1e10 when the midpoint is not strictly inside both hulls
// (or the hull construction fails), like the reference's err_value
SD3_HD inline float overlap_convex_volume(const float* pv1, const float* c1, const float* pv2, const float* c2, int n_rays) {
  if (n_rays > SD3_MAX_RAYS) return 1.e10f;
  Plane planes[2 * SD3_MAX_FACES];
  double pts[3 * SD3_MAX_RAYS];
  uint32_t edge_done[(SD3_MAX_RAYS * SD3_MAX_RAYS + 31) / 32];
  int16_t stack[3 * 4 * SD3_MAX_RAYS];
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv1[i];
  demote_duplicate_points(pts, n_rays);
  const int n1 = convex_hull_planes(pts, n_rays, planes, SD3_MAX_FACES, edge_done, stack, 4 * SD3_MAX_RAYS);
  if (n1 < 4) return 1.e10f;
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv2[i];
  demote_duplicate_points(pts, n_rays);
  const int n2 = convex_hull_planes(pts, n_rays, planes + n1, SD3_MAX_FACES, edge_done, stack, 4 * SD3_MAX_RAYS);
  if (n2 < 4) return 1.e10f;
  double p[3];
  for (int k = 0; k < 3; ++k) p[k] = .5 * ((double)c1[k] + (double)c2[k]);
  const int n = n1 + n2;
  for (int i = 0; i < n; ++i) if (!plane_feasible(planes[i], p)) return 1.e10f;
  const double L = extent_bound(pv1, pv2, n_rays, p);
  PlaneArray PA{planes};
  double vol = 0; int ovf = 0;
  for (int k = 0; k < n; ++k) vol += face_cone_volume(PA, n, k, p, L, &ovf);
  return (float)vol;
}

// ---- the same two stages on pre-normalised planes (face_cone_volume_n): bit-identical results, see geom3d.cuh ----------
SD3_HD inline float overlap_kernel_volume_n(const float* pv1, const float* c1, const float* pv2, const float* c2,
                                            const int* faces, int n_rays, int n_faces) {
  Plane planes[2 * SD3_MAX_FACES];
  if (n_faces > SD3_MAX_FACES) return 0.f;
  for (int i = 0; i < n_faces; ++i) {
    double hs[4];
    build_halfspace(&pv1[3 * faces[3 * i]], &pv1[3 * faces[3 * i + 1]], &pv1[3 * faces[3 * i + 2]], hs);
    planes[2 * i].n0 = hs[0]; planes[2 * i].n1 = hs[1]; planes[2 * i].n2 = hs[2]; planes[2 * i].d = hs[3];
    build_halfspace(&pv2[3 * faces[3 * i]], &pv2[3 * faces[3 * i + 1]], &pv2[3 * faces[3 * i + 2]], hs);
    planes[2 * i + 1].n0 = hs[0]; planes[2 * i + 1].n1 = hs[1]; planes[2 * i + 1].n2 = hs[2]; planes[2 * i + 1].d = hs[3];
  }
  double p[3];
  for (int k = 0; k < 3; ++k) p[k] = .5 * (double)(c1[k] + c2[k]);
  const int n = 2 * n_faces;
  for (int i = 0; i < n; ++i) if (!plane_feasible(planes[i], p)) return 0.f;
  const double L = extent_bound(pv1, pv2, n_rays, p);
  for (int i = 0; i < n; ++i) planes[i] = normalized_plane(planes[i]);
  PlaneArray PA{planes};
  double vol = 0; int ovf = 0;
  for (int k = 0; k < n; ++k) vol += face_cone_volume_n(PA, n, k, p, L, &ovf);
  return (float)vol;
}

SD3_HD inline float overlap_convex_volume_n(const float* pv1, const float* c1, const float* pv2, const float* c2, int n_rays) {
  if (n_rays > SD3_MAX_RAYS) return 1.e10f;
  Plane planes[2 * SD3_MAX_FACES];
  double pts[3 * SD3_MAX_RAYS];
  uint32_t edge_done[(SD3_MAX_RAYS * SD3_MAX_RAYS + 31) / 32];
  int16_t stack[3 * 4 * SD3_MAX_RAYS];
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv1[i];
  demote_duplicate_points(pts, n_rays);
  const int n1 = convex_hull_planes(pts, n_rays, planes, SD3_MAX_FACES, edge_done, stack, 4 * SD3_MAX_RAYS);
  if (n1 < 4) return 1.e10f;
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv2[i];
  demote_duplicate_points(pts, n_rays);
  const int n2 = convex_hull_planes(pts, n_rays, planes + n1, SD3_MAX_FACES, edge_done, stack, 4 * SD3_MAX_RAYS);
  if (n2 < 4) return 1.e10f;
  double p[3];
  for (int k = 0; k < 3; ++k) p[k] = .5 * ((double)c1[k] + (double)c2[k]);
  const int n = n1 + n2;
  for (int i = 0; i < n; ++i) if (!plane_feasible(planes[i], p)) return 1.e10f;
  const double L = extent_bound(pv1, pv2, n_rays, p);
  for (int i = 0; i < n; ++i) planes[i] = normalized_plane(planes[i]);
  PlaneArray PA{planes};
  double vol = 0; int ovf = 0;
  for (int k = 0; k < n; ++k) vol += face_cone_volume_n(PA, n, k, p, L, &ovf);
  return (float)vol;
}

}  // namespace sd3

// ---- ray-fan bounds of a convex polytope's volume (serial definition of the arithmetic of fan_bounds / k_s3_bound_warp in
// nms3d.cu; used by tests/hostcheck to check  lower <= volume <= upper  on the CPU) --------------------------------------
// planes[0..np): n.x + d <= 0 inside; p strictly inside; verts[R][3] ray directions, faces[F][3] their triangulation.
// subdiv: one extra ray per face along v_a + v_b + v_c, every face cone split into three sub-cones.
namespace sd3 {
SD3_HD inline double fan_det(const double* A, double ta, const double* B, double tb, const double* C, double tc) {
  const double Az = ta * A[0], Ay = ta * A[1], Ax = ta * A[2], Bz = tb * B[0], By = tb * B[1], Bx = tb * B[2], Cz = tc * C[0], Cy = tc * C[1], Cx = tc * C[2];
  const double M00 = Bz - Az, M01 = By - Ay, M02 = Bx - Ax, M10 = Cz - Az, M11 = Cy - Ay, M12 = Cx - Ax, M20 = -Az, M21 = -Ay, M22 = -Ax;
  return M00 * (M11 * M22 - M21 * M12) - M01 * (M10 * M22 - M12 * M20) + M02 * (M10 * M21 - M11 * M20);
}
SD3_HD inline void fan_first_plane(const Plane* planes, int np, const double* p, double Lext, const double* v, double* t, int* jhit) {
  double tn = Lext, td = 1.0; int jb = -1;
  for (int j = 0; j < np; ++j) {
    const Plane& P = planes[j];
    const double a = P.n0 * v[0] + P.n1 * v[1] + P.n2 * v[2];
    if (a > 0) {
      const double sd = -(P.d + P.n0 * p[0] + P.n1 * p[1] + P.n2 * p[2]);
      if (sd * td < tn * a) { tn = sd; td = a; jb = j; }
    }
  }
  *t = tn / td; *jhit = jb;
}
SD3_HD inline bool fan_cone(const Plane* planes, const double* p, const double* A, double tA, int jA, const double* B, double tB, int jB,
                            const double* C, double tC, int jC, double* lo, double* up) {
  if (!(fan_det(A, 1.0, B, 1.0, C, 1.0) > 0)) return false;
  const double l = fan_det(A, tA, B, tB, C, tC);
  *lo += l > 0 ? l : 0.0;
  double u = 1e300;
  const int js[3] = {jA, jB, jC};
  for (int e = 0; e < 3; ++e) {
    const int j = js[e];
    if (j < 0) continue;
    const Plane& P = planes[j];
    const double sd = -(P.d + P.n0 * p[0] + P.n1 * p[1] + P.n2 * p[2]);
    const double qa = P.n0 * A[0] + P.n1 * A[1] + P.n2 * A[2], qb = P.n0 * B[0] + P.n1 * B[1] + P.n2 * B[2], qc = P.n0 * C[0] + P.n1 * C[1] + P.n2 * C[2];
    if (qa > 0 && qb > 0 && qc > 0) { const double d = fan_det(A, sd / qa, B, sd / qb, C, sd / qc); u = d < u ? d : u; }
  }
  if (u >= 1e299) return false;
  *up += u;
  return true;
}
// t / jh: scratch of n_rays entries
SD3_HD inline void fan_bounds_serial(const Plane* planes, int np, const double* p, double Lext, const float* verts, const int* faces, int n_rays,
                                     int n_faces, int subdiv, double* t, int* jh, double* lower, double* upper) {
  for (int k = 0; k < n_rays; ++k) {
    const double v[3] = {verts[3 * k], verts[3 * k + 1], verts[3 * k + 2]};
    fan_first_plane(planes, np, p, Lext, v, &t[k], &jh[k]);
  }
  double lo = 0, up = 0; bool bad = false;
  for (int f = 0; f < n_faces; ++f) {
    const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
    const double va[3] = {verts[3 * ia], verts[3 * ia + 1], verts[3 * ia + 2]}, vb[3] = {verts[3 * ib], verts[3 * ib + 1], verts[3 * ib + 2]},
                 vc[3] = {verts[3 * ic], verts[3 * ic + 1], verts[3 * ic + 2]};
    if (!(fan_det(va, 1.0, vb, 1.0, vc, 1.0) > 0)) { bad = true; continue; }
    bool ok;
    if (subdiv) {
      const double vm[3] = {va[0] + vb[0] + vc[0], va[1] + vb[1] + vc[1], va[2] + vb[2] + vc[2]};
      double tm; int jm;
      fan_first_plane(planes, np, p, Lext, vm, &tm, &jm);
      ok = fan_cone(planes, p, va, t[ia], jh[ia], vb, t[ib], jh[ib], vm, tm, jm, &lo, &up);
      ok = fan_cone(planes, p, vb, t[ib], jh[ib], vc, t[ic], jh[ic], vm, tm, jm, &lo, &up) && ok;
      ok = fan_cone(planes, p, vc, t[ic], jh[ic], va, t[ia], jh[ia], vm, tm, jm, &lo, &up) && ok;
    } else {
      ok = fan_cone(planes, p, va, t[ia], jh[ia], vb, t[ib], jh[ib], vc, t[ic], jh[ic], &lo, &up);
    }
    if (!ok) bad = true;
  }
  *lower = lo / 6.0;
  *upper = bad ? 1e300 : up / 6.0;
}
}  // namespace sd3
