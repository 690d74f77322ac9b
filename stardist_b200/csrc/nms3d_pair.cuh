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
  const int n1 = convex_hull_planes(pts, n_rays, planes, SD3_MAX_FACES, edge_done, stack, 4 * SD3_MAX_RAYS);
  if (n1 < 4) return 1.e10f;
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv2[i];
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
  const int n1 = convex_hull_planes(pts, n_rays, planes, SD3_MAX_FACES, edge_done, stack, 4 * SD3_MAX_RAYS);
  if (n1 < 4) return 1.e10f;
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv2[i];
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
