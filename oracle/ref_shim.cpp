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

} // extern "C"
