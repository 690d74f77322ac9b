// tests/hostcheck/hostcheck.cpp -- TEST INFRASTRUCTURE ONLY.
// Host (g++) build of the device-side geometry headers so that their bit-exactness against
// oracle/_ref can be checked on the GPU-less dev box over millions of cases.  The product
// never loads this library: the shipped path is the CUDA build of the same headers.
#define SDC_STATS 1
#include "../../stardist_b200/csrc/clip2d.cuh"
#include <vector>
using namespace sdclip;

extern "C" {

int hc_hw[8];
void hc_clip_area_batch(const int32_t* a_xy, const int32_t* b_xy, int n_pairs, int n,
                        float* out_area, int* out_status) {
  for (int k = 0; k < 8; k++) hc_hw[k] = 0;
#pragma omp parallel
  {
    ClipSweep<128>* S = new ClipSweep<128>();
    std::vector<int32_t> ax(n), ay(n), bx(n), by(n);
#pragma omp for schedule(dynamic, 64)
    for (int p = 0; p < n_pairs; p++) {
      for (int i = 0; i < n; i++) {
        ax[i] = a_xy[2 * ((long)p * n + i)]; ay[i] = a_xy[2 * ((long)p * n + i) + 1];
        bx[i] = b_xy[2 * ((long)p * n + i)]; by[i] = b_xy[2 * ((long)p * n + i) + 1];
      }
      int st;
      SplitXY va{ax.data(), ay.data()}, vb{bx.data(), by.data()};
      out_area[p] = clip_intersection_area(va, vb, n, *S, &st);
      out_status[p] = st;
#pragma omp critical
      for (int k = 0; k < 8; k++) if (S->hw[k] > hc_hw[k]) hc_hw[k] = S->hw[k];
    }
    delete S;
  }
}

// single pair with the result paths written out (BuildResult order) for debugging
int hc_clip_paths(const int32_t* a_xy, int na, const int32_t* b_xy, int nb,
                  int32_t* out_xy, int* out_counts, int max_paths, int max_pts, int* status) {
  ClipSweep<128>* S = new ClipSweep<128>();
  std::vector<int32_t> ax(na), ay(na), bx(nb), by(nb);
  for (int i = 0; i < na; i++) { ax[i] = a_xy[2*i]; ay[i] = a_xy[2*i+1]; }
  for (int i = 0; i < nb; i++) { bx[i] = b_xy[2*i]; by[i] = b_xy[2*i+1]; }
  S->init();
  SplitXY va{ax.data(), ay.data()}, vb{bx.data(), by.data()};
  S->add_path(va, na, ptClip);
  S->add_path(vb, nb, ptSubject);
  bool ok = S->execute();
  *status = S->err;
  int np = 0, tot = 0;
  if (ok) for (int i = 0; i < S->nR; i++) {
    if (S->R[i].pts == SDC_NIL) continue;
    ix p0 = S->P[S->R[i].pts].prev;
    int cnt = S->point_count(p0);
    if (cnt < 2) continue;
    if (np >= max_paths || tot + cnt > max_pts) { delete S; return -1; }
    out_counts[np++] = cnt;
    ix p = p0;
    for (int k = 0; k < cnt; k++) { out_xy[2*tot] = S->P[p].x; out_xy[2*tot+1] = S->P[p].y; tot++; p = S->P[p].prev; }
  }
  delete S;
  return np;
}

}

// ---------------------------------------------------------------- fast overlap integral (polyfast.cuh)
#include "../../stardist_b200/csrc/polyfast.cuh"
extern "C" {
// out: I (overlap integral), bound (clipper_bound), K (crossings)
void hc_fast_batch(const int32_t* a_xy, const int32_t* b_xy, int n_pairs, int n,
                   double* out_I, double* out_bound, int* out_K) {
#pragma omp parallel
  {
    std::vector<sdfast::Edge> ea(n), eb(n);
    std::vector<double> fa(n), fb(n);
#pragma omp for schedule(dynamic, 256)
    for (int p = 0; p < n_pairs; p++) {
      const int32_t* A = a_xy + 2 * (long)p * n; const int32_t* B = b_xy + 2 * (long)p * n;
      double maxlen_a = 0, maxlen_b = 0, maxc = 0;
      for (int i = 0; i < n; i++) {
        int k = (i + 1) % n;
        ea[i] = {A[2*i], A[2*i+1], A[2*k], A[2*k+1]};
        eb[i] = {B[2*i], B[2*i+1], B[2*k], B[2*k+1]};
        maxlen_a = fmax(maxlen_a, hypot((double)ea[i].x1 - ea[i].x0, (double)ea[i].y1 - ea[i].y0));
        maxlen_b = fmax(maxlen_b, hypot((double)eb[i].x1 - eb[i].x0, (double)eb[i].y1 - eb[i].y0));
        maxc = fmax(maxc, fmax(fmax(fabs((double)A[2*i]), fabs((double)A[2*i+1])), fmax(fabs((double)B[2*i]), fabs((double)B[2*i+1]))));
      }
      double sa = 0, sb = 0;
      for (int i = n - 1; i >= 0; i--) {
        fa[i] = sa; sa += sdfast::edge_F(ea[i].x0, ea[i].y0, ea[i].x1, ea[i].y1);
        fb[i] = sb; sb += sdfast::edge_F(eb[i].x0, eb[i].y0, eb[i].x1, eb[i].y1);
      }
      int wq = 0, wp = 0;
      for (int i = 0; i < n; i++) { wq += sdfast::wind_Q_edge(eb[i], A[0], A[1]); wp += sdfast::wind_P_edge(ea[i], B[0], B[1]); }
      sdfast::Accum acc; acc.clear();
      const bool small = maxc <= 8191.0;
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
          if (small) sdfast::edge_pair<int32_t>(ea[i], &fa[i], eb[j], &fb[j], acc);
          else sdfast::edge_pair<long long>(ea[i], &fa[i], eb[j], &fb[j], acc);
        }
      out_I[p] = acc.I + wq * sa + wp * sb;
      out_bound[p] = sdfast::clipper_bound(acc, maxlen_a + maxlen_b, maxc, n);
      out_K[p] = acc.K;
    }
  }
}
}

// ---------------------------------------------------------------- 3D geometry (geom3d.cuh)
#include "../../stardist_b200/csrc/geom3d.cuh"
#include "../../stardist_b200/csrc/nms3d_pair.cuh"

extern "C" {
float hc_overlap_kernel(const float* pv1, const float* c1, const float* pv2, const float* c2,
                        const int* faces, int n_rays, int n_faces) {
  return sd3::overlap_kernel_volume(pv1, c1, pv2, c2, faces, n_rays, n_faces);
}
float hc_overlap_convex(const float* pv1, const float* c1, const float* pv2, const float* c2,
                        int n_rays) {
  return sd3::overlap_convex_volume(pv1, c1, pv2, c2, n_rays);
}
float hc_overlap_kernel_n(const float* pv1, const float* c1, const float* pv2, const float* c2,
                          const int* faces, int n_rays, int n_faces) {
  return sd3::overlap_kernel_volume_n(pv1, c1, pv2, c2, faces, n_rays, n_faces);
}
float hc_overlap_convex_n(const float* pv1, const float* c1, const float* pv2, const float* c2, int n_rays) {
  return sd3::overlap_convex_volume_n(pv1, c1, pv2, c2, n_rays);
}
// double results (before the float rounding) of both formulations, for a stricter comparison
void hc_overlap_kernel_pair(const float* pv1, const float* c1, const float* pv2, const float* c2, const int* faces, int n_rays,
                            int n_faces, double* out2) {
  using namespace sd3;
  std::vector<Plane> planes(2 * (size_t)n_faces), pn(2 * (size_t)n_faces);
  for (int i = 0; i < n_faces; ++i) {
    double hs[4];
    build_halfspace(&pv1[3 * faces[3 * i]], &pv1[3 * faces[3 * i + 1]], &pv1[3 * faces[3 * i + 2]], hs);
    planes[2 * i] = Plane{hs[0], hs[1], hs[2], hs[3]};
    build_halfspace(&pv2[3 * faces[3 * i]], &pv2[3 * faces[3 * i + 1]], &pv2[3 * faces[3 * i + 2]], hs);
    planes[2 * i + 1] = Plane{hs[0], hs[1], hs[2], hs[3]};
  }
  double p[3];
  for (int k = 0; k < 3; ++k) p[k] = .5 * (double)(c1[k] + c2[k]);
  const int n = 2 * n_faces;
  const double L = extent_bound(pv1, pv2, n_rays, p);
  for (int i = 0; i < n; ++i) pn[i] = normalized_plane(planes[i]);
  PlaneArray PA{planes.data()}, PN{pn.data()};
  double a = 0, b = 0; int ovf = 0;
  for (int k = 0; k < n; ++k) { a += face_cone_volume(PA, n, k, p, L, &ovf); b += face_cone_volume_n(PN, n, k, p, L, &ovf); }
  out2[0] = a; out2[1] = b;
}

// Serial host restatement of nms3d.cu (k_pre1 / k_aniso / k_pre2 / k_frontier+k_pretest as the greedy loop they resolve /
// k_heavy) with the SAME header arithmetic (geom3d.cuh, nms3d_pair.cuh): pins S1, S2, S5, the radius query and the anisotropy
// on the CPU box against the reference's c_non_max_suppression_inds for ray counts the GPU goldens do not cover.
// norm_planes: use the *_n volume stages (sdb_nms3d_set_variant(1)).
void hc_nms3d_serial(const float* dist, const float* points, const float* verts, const int* faces, int n, int R, int F,
                     float threshold, int use_bbox, int use_kdtree, int norm_planes, unsigned char* keep, int* stage_counts /*[5]*/) {
  using namespace sd3;
  std::vector<float> volume(n), r_outer(n), r_outer_iso(n), r_inner_iso(n), terms(3 * (size_t)n);
  std::vector<int> bbox(6 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const float* d = dist + (size_t)i * R; const float* c = points + 3 * i;
    volume[i] = polyhedron_volume(d, verts, faces, F);
    polyhedron_bbox(d, c, verts, R, &bbox[6 * i]);
    terms[i] = (float)(bbox[6 * i + 1] - bbox[6 * i]) / n;
    terms[n + i] = (float)(bbox[6 * i + 3] - bbox[6 * i + 2]) / n;
    terms[2 * (size_t)n + i] = (float)(bbox[6 * i + 5] - bbox[6 * i + 4]) / n;
    r_outer[i] = bounding_radius_outer(d, R);
  }
  float an[3];
  for (int a = 0; a < 3; ++a) { float acc = 0.f; for (int i = 0; i < n; ++i) acc = acc + terms[(size_t)a * n + i]; an[a] = acc; }
  { const float tmp = fmaxf(fmaxf(an[0], an[1]), an[2]); const float a0 = an[0], a1 = an[1], a2 = an[2]; an[0] = tmp / a0; an[1] = tmp / a1; an[2] = tmp / a2; }
  float max_dist = 0.f;
  for (int i = 0; i < n; ++i) {
    const float* d = dist + (size_t)i * R;
    r_outer_iso[i] = bounding_radius_outer_isotropic(d, verts, R, an);
    r_inner_iso[i] = bounding_radius_inner_isotropic(d, verts, faces, F, an);
    max_dist = fmaxf(max_dist, r_outer[i]);
  }
  std::vector<unsigned char> sup(n, 0);
  std::vector<float> pv1(3 * (size_t)R), pv2(3 * (size_t)R);
  for (int k = 0; k < 5; ++k) stage_counts[k] = 0;
  for (int h = 0; h < n; ++h) {
    if (sup[h]) continue;
    const float* ph = points + 3 * h; const float* d1 = dist + (size_t)h * R;
    for (int j = 0; j < R; ++j) for (int k = 0; k < 3; ++k) pv1[3 * j + k] = ph[k] + d1[j] * verts[3 * j + k];
    for (int c = h + 1; c < n; ++c) {
      if (sup[c]) continue;
      const float* pc = points + 3 * c;
      if (use_kdtree) {
        const float d0 = ph[0] - pc[0], dd1 = ph[1] - pc[1], d2 = ph[2] - pc[2];
        const float dd = d0 * d0 + dd1 * dd1 + d2 * d2;
        const float rr = max_dist + r_outer[h];
        if (!(dd < rr * rr)) continue;
      }
      stage_counts[0]++;
      const float A_min = fminf(volume[h], volume[c]);
      float A_inter = fminf(intersect_sphere_isotropic(r_outer_iso[h], ph, r_outer_iso[c], pc, an), intersect_bbox(&bbox[6 * h], &bbox[6 * c]));
      float iou = (float)fmin(1.0, (double)A_inter / ((double)A_min + 1e-10));
      if (use_bbox && (((double)A_inter < 1.e-10) || (iou <= threshold))) continue;
      stage_counts[1]++;
      A_inter = intersect_sphere_isotropic(r_inner_iso[h], ph, r_inner_iso[c], pc, an);
      iou = (float)fmax(0.0, (double)A_inter / ((double)A_min + 1e-10));
      if (iou > threshold) { sup[c] = 1; continue; }
      stage_counts[2]++;
      const float* d2p = dist + (size_t)c * R;
      for (int j = 0; j < R; ++j) for (int k = 0; k < 3; ++k) pv2[3 * j + k] = pc[k] + d2p[j] * verts[3 * j + k];
      const double den = (double)A_min + 1e-10;
      const float vk = norm_planes ? overlap_kernel_volume_n(pv1.data(), ph, pv2.data(), pc, faces, R, F)
                                   : overlap_kernel_volume(pv1.data(), ph, pv2.data(), pc, faces, R, F);
      iou = (float)((double)vk / den);
      if (iou > threshold) { sup[c] = 1; continue; }
      stage_counts[3]++;
      const float vc = norm_planes ? overlap_convex_volume_n(pv1.data(), ph, pv2.data(), pc, R) : overlap_convex_volume(pv1.data(), ph, pv2.data(), pc, R);
      iou = (float)((double)vc / den);
      if (iou <= threshold) continue;
      stage_counts[4]++;
      // S5 exactly as the reference loop (:608-636): count voxels of bbox(h) inside both, return early once res > overlap_maximal
      const int* bb = &bbox[6 * h];
      const int Nz = bb[1] - bb[0] + 1, Ny = bb[3] - bb[2] + 1, Nx = bb[5] - bb[4] + 1;
      const float overlap_maximal = (float)(den * (double)threshold);
      int res = 0; bool stop = false;
      for (int z = 0; z < Nz && !stop; ++z) for (int y = 0; y < Ny && !stop; ++y) for (int x = 0; x < Nx; ++x) {
        const float fz = (float)(z + bb[0]), fy = (float)(y + bb[2]), fx = (float)(x + bb[4]);
        res += (inside_polyhedron(fz, fy, fx, ph, pv1.data(), faces, F) && inside_polyhedron(fz, fy, fx, pc, pv2.data(), faces, F)) ? 1 : 0;
        if ((float)res > overlap_maximal) { stop = true; break; }
      }
      const float iou5 = (float)((double)(float)res / den);
      if (iou5 > threshold) sup[c] = 1;
    }
  }
  for (int i = 0; i < n; ++i) keep[i] = !sup[i];
}

// host restatement of k_paint3d (label3d.cu), render mode "full", ONE polyhedron: the same header functions in the same
// order -- vertices and integer bbox in float, kernel half-spaces in double, then inside_polyhedron -- so that the
// device rendering rule can be compared with the reference's c_polyhedron_to_label on the CPU box.
void hc_paint3d(const float* dist, const float* center, const float* verts, const int* faces, int n_rays, int n_faces,
                int nz, int ny, int nx, unsigned char* out) {
  std::vector<float> pv(3 * (size_t)n_rays);
  std::vector<double> hs(4 * (size_t)n_faces);
  int z1 = INT32_MAX, z2 = -1, y1 = INT32_MAX, y2 = -1, x1 = INT32_MAX, x2 = -1;
  for (int j = 0; j < n_rays; ++j) {
    pv[3 * j] = center[0] + dist[j] * verts[3 * j];
    pv[3 * j + 1] = center[1] + dist[j] * verts[3 * j + 1];
    pv[3 * j + 2] = center[2] + dist[j] * verts[3 * j + 2];
    const int iz = sd3::round_to_int(center[0] + dist[j] * verts[3 * j]);
    const int iy = sd3::round_to_int(center[1] + dist[j] * verts[3 * j + 1]);
    const int ix = sd3::round_to_int(center[2] + dist[j] * verts[3 * j + 2]);
    z1 = std::min(z1, iz); z2 = std::max(z2, iz); y1 = std::min(y1, iy); y2 = std::max(y2, iy); x1 = std::min(x1, ix); x2 = std::max(x2, ix);
  }
  z1 = std::max(0, z1); z2 = std::min(nz - 1, z2); y1 = std::max(0, y1); y2 = std::min(ny - 1, y2); x1 = std::max(0, x1); x2 = std::min(nx - 1, x2);
  for (int f = 0; f < n_faces; ++f)
    sd3::build_halfspace(&pv[3 * faces[3 * f]], &pv[3 * faces[3 * f + 1]], &pv[3 * faces[3 * f + 2]], &hs[4 * f]);
  for (int z = z1; z <= z2; ++z)
    for (int y = y1; y <= y2; ++y)
      for (int x = x1; x <= x2; ++x) {
        const float fz = (float)z, fy = (float)y, fx = (float)x;
        bool ker = true;
        for (int f = 0; f < n_faces && ker; ++f)
          if (hs[4 * f] * fz + hs[4 * f + 1] * fy + hs[4 * f + 2] * fx + hs[4 * f + 3] > 0) ker = false;
        if (ker || sd3::inside_polyhedron(fz, fy, fx, center, pv.data(), faces, n_faces))
          out[((size_t)z * ny + y) * nx + x] = 1;
      }
}
}

// ray-fan bounds vs the volumes they bound (nms3d.cu fan_bounds): out = {feasible, V_kernel, lo, up, lo_refined, up_refined,
//                                                                      hull_ok, V_hull, lo, up, lo_refined, up_refined}
static int g_fan_ovf[2];
extern "C" void hc_fan_last_overflow(int* two) { two[0] = g_fan_ovf[0]; two[1] = g_fan_ovf[1]; }
extern "C" void hc_fan_bounds_pair(const float* pv1, const float* c1, const float* pv2, const float* c2, const float* verts, const int* faces,
                                   int n_rays, int n_faces, double* out12) {
  using namespace sd3;
  for (int i = 0; i < 12; ++i) out12[i] = 0;
  g_fan_ovf[0] = g_fan_ovf[1] = 0;
  std::vector<Plane> planes(2 * (size_t)SD3_MAX_FACES);
  std::vector<double> t(n_rays); std::vector<int> jh(n_rays);
  double p[3];
  // kernel halfspaces (stage S3)
  for (int f = 0; f < n_faces; ++f) {
    double hs[4];
    build_halfspace(&pv1[3 * faces[3 * f]], &pv1[3 * faces[3 * f + 1]], &pv1[3 * faces[3 * f + 2]], hs);
    planes[2 * f] = Plane{hs[0], hs[1], hs[2], hs[3]};
    build_halfspace(&pv2[3 * faces[3 * f]], &pv2[3 * faces[3 * f + 1]], &pv2[3 * faces[3 * f + 2]], hs);
    planes[2 * f + 1] = Plane{hs[0], hs[1], hs[2], hs[3]};
  }
  for (int k = 0; k < 3; ++k) p[k] = .5 * (double)(c1[k] + c2[k]);
  int np = 2 * n_faces; bool feas = true;
  for (int i = 0; i < np; ++i) if (!plane_feasible(planes[i], p)) feas = false;
  if (feas) {
    const double L = extent_bound(pv1, pv2, n_rays, p);
    PlaneArray PA{planes.data()};
    double vol = 0; int ovf = 0;
    for (int k = 0; k < np; ++k) vol += face_cone_volume(PA, np, k, p, L, &ovf);
    out12[0] = 1; out12[1] = vol; g_fan_ovf[0] = ovf;
    fan_bounds_serial(planes.data(), np, p, L, verts, faces, n_rays, n_faces, 0, t.data(), jh.data(), &out12[2], &out12[3]);
    fan_bounds_serial(planes.data(), np, p, L, verts, faces, n_rays, n_faces, 1, t.data(), jh.data(), &out12[4], &out12[5]);
  }
  // hull halfspaces (stage S4)
  std::vector<double> pts(3 * (size_t)n_rays);
  std::vector<uint32_t> edge_done(((size_t)n_rays * n_rays + 31) / 32);
  std::vector<int16_t> stack(3 * 4 * (size_t)n_rays);
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv1[i];
  const int n1 = convex_hull_planes(pts.data(), n_rays, planes.data(), SD3_MAX_FACES, edge_done.data(), stack.data(), 4 * n_rays);
  if (n1 < 4) return;
  for (int i = 0; i < 3 * n_rays; ++i) pts[i] = (double)pv2[i];
  const int n2 = convex_hull_planes(pts.data(), n_rays, planes.data() + n1, SD3_MAX_FACES, edge_done.data(), stack.data(), 4 * n_rays);
  if (n2 < 4) return;
  for (int k = 0; k < 3; ++k) p[k] = .5 * ((double)c1[k] + (double)c2[k]);
  np = n1 + n2; feas = true;
  for (int i = 0; i < np; ++i) if (!plane_feasible(planes[i], p)) feas = false;
  if (!feas) return;
  const double L = extent_bound(pv1, pv2, n_rays, p);
  PlaneArray PA{planes.data()};
  double vol = 0; int ovf = 0;
  for (int k = 0; k < np; ++k) vol += face_cone_volume(PA, np, k, p, L, &ovf);
  out12[6] = 1; out12[7] = vol; g_fan_ovf[1] = ovf;
  fan_bounds_serial(planes.data(), np, p, L, verts, faces, n_rays, n_faces, 0, t.data(), jh.data(), &out12[8], &out12[9]);
  fan_bounds_serial(planes.data(), np, p, L, verts, faces, n_rays, n_faces, 1, t.data(), jh.data(), &out12[10], &out12[11]);
}

// direction bins of the 3-D rendering / S5 (bins3d.cuh): bin of a direction, and whether a bin lists a face
#include "../../stardist_b200/csrc/bins3d.cuh"
extern "C" int hc_bin_of(float u0, float u1, float u2) { return sdbins::bin_of(u0, u1, u2); }
extern "C" int hc_bin_takes_face(int b, const float* verts, const int* faces, int f) { return sdbins::bin_takes_face(b, verts, faces, f) ? 1 : 0; }
extern "C" int hc_bin_count() { return sdbins::BIN_N; }

// hull facet planes of one polyhedron (geom3d.cuh convex_hull_planes): planes[4 * n] out, returns the facet count
extern "C" int hc_convex_hull_planes(const double* pts, int n, double* planes_out, int max_planes) {
  std::vector<sd3::Plane> pl(max_planes);
  std::vector<uint32_t> edge_done(((size_t)n * n + 31) / 32);
  std::vector<int16_t> stack(3 * 4 * (size_t)n);
  const int nf = sd3::convex_hull_planes(pts, n, pl.data(), max_planes, edge_done.data(), stack.data(), 4 * n);
  for (int i = 0; i < nf; ++i) { planes_out[4 * i] = pl[i].n0; planes_out[4 * i + 1] = pl[i].n1; planes_out[4 * i + 2] = pl[i].n2; planes_out[4 * i + 3] = pl[i].d; }
  return nf;
}

// volume of the intersection of n halfspaces around the interior point p (geom3d.cuh face_cone_volume), per-face parts out
extern "C" double hc_planes_volume(const double* planes4, int n, const double* p, double L, double* parts, int* overflow) {
  std::vector<sd3::Plane> pl(n);
  for (int i = 0; i < n; ++i) pl[i] = sd3::Plane{planes4[4 * i], planes4[4 * i + 1], planes4[4 * i + 2], planes4[4 * i + 3]};
  sd3::PlaneArray PA{pl.data()};
  double vol = 0; *overflow = 0;
  for (int k = 0; k < n; ++k) { const double v = sd3::face_cone_volume(PA, n, k, p, L, overflow); if (parts) parts[k] = v; vol += v; }
  return vol;
}

// ---------------------------------------------------------------- 3-D label rendering rule (label3d.cu k_paint3d, serial)
// The DEFINING per-voxel rule of the device rendering, without its result-neutral short cuts (sphere culling, direction bins):
//   mode 0 "full":   kernel planes  ||  inside_polyhedron          (the reference: kernel || (hull && inside_polyhedron))
//   mode 1 "kernel": kernel planes;   mode 2 "hull": hull facets (gift wrapping);   mode 3 "bbox": every voxel of the box
// polyhedra in the given order, first cover wins (all labels non-zero, no overlap label).  out: int32 [nz][ny][nx], zeroed here.
extern "C" void hc_polyhedron_to_label(const float* dist, const float* points, const float* verts, const int* faces, int n_polys, int n_rays,
                                       int n_faces, const int* labels, int nz, int ny, int nx, int mode, int* out) {
  using namespace sd3;
  for (long long v = 0; v < (long long)nz * ny * nx; ++v) out[v] = 0;
  // mode 0 for a ray set with degenerate faces: the reference's hull conjunct is applied (sdb_polyhedron_to_label does the same)
  bool hull_conj = false;
  if (mode == 0) for (int f = 0; f < n_faces; ++f) hull_conj = hull_conj || ray_face_is_degenerate(verts, faces, f);
  std::vector<double> hh(4 * (size_t)std::max(n_faces, 4));
  std::vector<float> pv(3 * (size_t)n_rays);
  std::vector<double> hs(4 * (size_t)std::max(n_faces, 1)), pts(3 * (size_t)n_rays);
  std::vector<Plane> hull(n_faces > 4 ? n_faces : 4);
  std::vector<uint32_t> edge_done(((size_t)n_rays * n_rays + 31) / 32);
  std::vector<int16_t> stack(3 * 4 * (size_t)n_rays);
  for (int i = 0; i < n_polys; ++i) {
    const float* d = dist + (size_t)i * n_rays;
    const float center[3] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
    int z1 = INT32_MAX, z2 = -1, y1 = INT32_MAX, y2 = -1, x1 = INT32_MAX, x2 = -1;
    for (int j = 0; j < n_rays; ++j) {
      pv[3 * j] = center[0] + d[j] * verts[3 * j]; pv[3 * j + 1] = center[1] + d[j] * verts[3 * j + 1]; pv[3 * j + 2] = center[2] + d[j] * verts[3 * j + 2];
      const int iz = round_to_int(center[0] + d[j] * verts[3 * j]), iy = round_to_int(center[1] + d[j] * verts[3 * j + 1]), ix = round_to_int(center[2] + d[j] * verts[3 * j + 2]);
      z1 = std::min(z1, iz); z2 = std::max(z2, iz); y1 = std::min(y1, iy); y2 = std::max(y2, iy); x1 = std::min(x1, ix); x2 = std::max(x2, ix);
    }
    z1 = std::max(0, z1); z2 = std::min(nz - 1, z2); y1 = std::max(0, y1); y2 = std::min(ny - 1, y2); x1 = std::max(0, x1); x2 = std::min(nx - 1, x2);
    int n_planes = n_faces, n_hull = 0;
    if (mode == 0 && hull_conj) {
      for (int k = 0; k < 3 * n_rays; ++k) pts[k] = (double)pv[k];
      demote_duplicate_points(pts.data(), n_rays);
      n_hull = convex_hull_planes(pts.data(), n_rays, hull.data(), n_faces, edge_done.data(), stack.data(), 4 * n_rays);
      for (int f = 0; f < n_hull; ++f) { hh[4 * f] = hull[f].n0; hh[4 * f + 1] = hull[f].n1; hh[4 * f + 2] = hull[f].n2; hh[4 * f + 3] = hull[f].d; }
    }
    if (mode == 2) {
      for (int k = 0; k < 3 * n_rays; ++k) pts[k] = (double)pv[k];
      demote_duplicate_points(pts.data(), n_rays);
      n_planes = convex_hull_planes(pts.data(), n_rays, hull.data(), n_faces, edge_done.data(), stack.data(), 4 * n_rays);
      for (int f = 0; f < n_planes; ++f) { hs[4 * f] = hull[f].n0; hs[4 * f + 1] = hull[f].n1; hs[4 * f + 2] = hull[f].n2; hs[4 * f + 3] = hull[f].d; }
    } else {
      for (int f = 0; f < n_faces; ++f) build_halfspace(&pv[3 * faces[3 * f]], &pv[3 * faces[3 * f + 1]], &pv[3 * faces[3 * f + 2]], &hs[4 * f]);
    }
    for (int z = z1; z <= z2; ++z) for (int y = y1; y <= y2; ++y) for (int x = x1; x <= x2; ++x) {
      const float fz = (float)z, fy = (float)y, fx = (float)x;
      auto in_planes = [&](int cnt) {
        for (int f = 0; f < cnt; ++f) if (hs[4 * f] * fz + hs[4 * f + 1] * fy + hs[4 * f + 2] * fx + hs[4 * f + 3] > 0) return false;
        return true;
      };
      bool inside;
      auto in_hull = [&]() {
        if (n_hull < 4) return false;
        for (int f = 0; f < n_hull; ++f) if (hh[4 * f] * fz + hh[4 * f + 1] * fy + hh[4 * f + 2] * fx + hh[4 * f + 3] > 0) return false;
        return true;
      };
      if (mode == 0 && hull_conj) inside = in_planes(n_faces) || (in_hull() && inside_polyhedron(fz, fy, fx, center, pv.data(), faces, n_faces));
      else if (mode == 0) inside = in_planes(n_faces) || inside_polyhedron(fz, fy, fx, center, pv.data(), faces, n_faces);
      else if (mode == 1) inside = in_planes(n_faces);
      else if (mode == 2) inside = n_planes >= 4 && in_planes(n_planes);
      else inside = true;
      int& o = out[((size_t)z * ny + y) * nx + x];
      if (inside && o == 0) o = labels[i];
    }
  }
}
