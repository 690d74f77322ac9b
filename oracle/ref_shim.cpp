// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
//
// Thin C-ABI shim over the *reference's own* compiled sources so tests can check single
// pair decisions, not only whole NMS runs.  Nothing is restated here: the 3D helpers are
// pulled in by compiling the reference translation unit where it lies
// (stardist/lib/stardist3d_impl.cpp) and the 2D pair area calls the vendored Clipper
// exactly the way stardist/lib/stardist2d.cpp:152-165 does.
#include <vector>
#include <cstring>
#include "clipper.hpp"
#include "stardist3d_impl.cpp"   // reference TU, via -I$(REF)/stardist/lib

extern "C" {

// Clipper ctIntersection / pftNonZero of two closed int paths (stardist2d.cpp:152-165).
// a is added as ptClip, b as ptSubject (same order as the reference).
// out_xy receives the result paths back to back as (X,Y) pairs, out_counts the vertex count
// of each path. Returns the number of result paths, or -1 if the buffers are too small.
int sdref_clip_intersection(const long long* a_xy, int na, const long long* b_xy, int nb,
                            long long* out_xy, int* out_counts, int max_paths, int max_pts) {
  ClipperLib::Path pa, pb;
  for (int i = 0; i < na; i++) pa << ClipperLib::IntPoint(a_xy[2*i], a_xy[2*i+1]);
  for (int i = 0; i < nb; i++) pb << ClipperLib::IntPoint(b_xy[2*i], b_xy[2*i+1]);
  ClipperLib::Clipper c;
  ClipperLib::Paths res;
  c.Clear();
  c.AddPath(pa, ClipperLib::ptClip, true);
  c.AddPath(pb, ClipperLib::ptSubject, true);
  c.Execute(ClipperLib::ctIntersection, res, ClipperLib::pftNonZero, ClipperLib::pftNonZero);
  if ((int)res.size() > max_paths) return -1;
  int tot = 0;
  for (size_t r = 0; r < res.size(); r++) {
    out_counts[r] = (int)res[r].size();
    if (tot + (int)res[r].size() > max_pts) return -1;
    for (size_t k = 0; k < res[r].size(); k++) {
      out_xy[2*tot]   = res[r][k].X;
      out_xy[2*tot+1] = res[r][k].Y;
      tot++;
    }
  }
  return (int)res.size();
}

// float area exactly as stardist2d.cpp:128-138 (float accumulator over int64 cross products)
static float area_like_ref(const ClipperLib::Path& p) {
  float area = 0;
  const int n = p.size();
  for (int i = 0; i < n; i++)
    area += p[i].X * p[(i+1)%n].Y - p[i].Y * p[(i+1)%n].X;
  area = 0.5 * std::abs(area);
  return area;
}

// batch version: n_pairs pairs of n-gons -> intersection area (float) per pair.
void sdref_clip_area_batch(const long long* a_xy, const long long* b_xy, int n_pairs, int n,
                           float* out_area) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int p = 0; p < n_pairs; p++) {
    ClipperLib::Path pa, pb;
    for (int i = 0; i < n; i++) pa << ClipperLib::IntPoint(a_xy[2*((long)p*n+i)], a_xy[2*((long)p*n+i)+1]);
    for (int i = 0; i < n; i++) pb << ClipperLib::IntPoint(b_xy[2*((long)p*n+i)], b_xy[2*((long)p*n+i)+1]);
    ClipperLib::Clipper c;
    ClipperLib::Paths res;
    c.AddPath(pa, ClipperLib::ptClip, true);
    c.AddPath(pb, ClipperLib::ptSubject, true);
    c.Execute(ClipperLib::ctIntersection, res, ClipperLib::pftNonZero, ClipperLib::pftNonZero);
    float a = 0;
    for (size_t r = 0; r < res.size(); r++) a += area_like_ref(res[r]);
    out_area[p] = a;
  }
}

// ---- 3D inner functions of the reference (stardist3d_impl.cpp), exposed unchanged ----------
// polyverts* are (n_rays,3) float arrays as produced by polyhedron_polyverts (:570-585)
void sdref_polyverts(const float* dist, const float* center, const float* verts, int n_rays, float* out) {
  polyhedron_polyverts(dist, center, verts, n_rays, out);
}
float sdref_overlap_kernel(const float* pv1, const float* c1, const float* pv2, const float* c2,
                           const int* faces, int n_rays, int n_faces) {
  return qhull_overlap_kernel(pv1, c1, pv2, c2, faces, n_rays, n_faces);
}
float sdref_overlap_convex(const float* pv1, const float* c1, const float* pv2, const float* c2,
                           const int* faces, int n_rays, int n_faces) {
  return qhull_overlap_convex_hulls(pv1, c1, pv2, c2, faces, n_rays, n_faces);
}
void sdref_bbox(const float* dist, const float* center, const float* verts, int n_rays, int* bbox) {
  polyhedron_bbox(dist, center, verts, n_rays, bbox);
}
float sdref_volume(const float* dist, const float* verts, const int* faces, int n_rays, int n_faces) {
  return polyhedron_volume(dist, verts, faces, n_rays, n_faces);
}
// render polyhedron 1 into its bbox, then count overlap with polyhedron 2 (early exit like :608-636)
int sdref_overlap_render(const float* dist1, const float* c1, const float* pv1,
                         const float* dist2, const float* c2, const float* pv2,
                         const float* verts, const int* faces, int n_rays, int n_faces, float overlap_maximal) {
  int bbox[6];
  polyhedron_bbox(dist1, c1, verts, n_rays, bbox);
  int Nz = bbox[1]-bbox[0]+1, Ny = bbox[3]-bbox[2]+1, Nx = bbox[5]-bbox[4]+1;
  bool* rendered = new bool[(size_t)Nz*Ny*Nx];
  render_polyhedron(dist1, c1, bbox, pv1, faces, n_rays, n_faces, rendered, Nz, Ny, Nx);
  int r = overlap_render_polyhedron(dist2, c2, bbox, pv2, faces, n_rays, n_faces, rendered, Nz, Ny, Nx, overlap_maximal);
  delete[] rendered;
  return r;
}
float sdref_sphere_iso(float r1, const float* p1, float r2, const float* p2, const float* aniso) {
  return intersect_sphere_isotropic(r1, p1, r2, p2, aniso);
}
float sdref_bbox_inter(const int* b1, const int* b2) { return intersect_bbox(b1, b2); }
float sdref_radius_outer_iso(const float* dist, const float* verts, int n_rays, const float* aniso) {
  return bounding_radius_outer_isotropic(dist, verts, n_rays, aniso);
}
float sdref_radius_inner_iso(const float* dist, const float* verts, const int* faces, int n_rays, int n_faces, const float* aniso) {
  return bounding_radius_inner_isotropic(dist, verts, faces, n_rays, n_faces, aniso);
}
float sdref_radius_outer(const float* dist, int n_rays) { return bounding_radius_outer(dist, n_rays); }

} // extern "C"
