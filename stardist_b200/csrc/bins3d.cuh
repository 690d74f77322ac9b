// bins3d.cuh -- direction bins of the ray triangulation (shared by label3d.cu and nms3d.cu; -fmad=false TUs).
#pragma once
#include "geom3d.cuh"

namespace sdbins {

// inside_polyhedron is an OR over the F tetrahedra (centre, A, B, C); a voxel can only lie in a tetrahedron whose cone
// (over the spherical triangle of the ray directions, the same for every polyhedron of a call) contains its direction from
// the centre.  The directions are binned on a cube map (6 faces x 4 x 4 cells); every bin lists the faces whose cone's
// bounding cap meets the bin's bounding cap widened by 0.05 rad -- a superset of the faces that can answer true, so testing
// only those gives the same boolean as testing all F (the float determinants of a tetrahedron whose cone is > 0.05 rad away
// are negative by a margin ~1e4 x their rounding).  The same lists put the kernel plane most likely to be violated first.
constexpr int BIN_B = 4, BIN_N = 6 * BIN_B * BIN_B, BIN_CAP = 64;
struct FaceBins { int* count; int* faces; };     // count[BIN_N] (-1: list overflow -> all faces), faces[BIN_N][BIN_CAP]

SD3_HD inline int bin_of(float u0, float u1, float u2) {
  const float a0 = fabsf(u0), a1 = fabsf(u1), a2 = fabsf(u2);
  int m = 0; float am = a0, um = u0, p = u1, q = u2;
  if (a1 > am) { m = 1; am = a1; um = u1; p = u0; q = u2; }
  if (a2 > am) { m = 2; am = a2; um = u2; p = u0; q = u1; }
  if (!(am > 0.f)) return -1;
  const float inv = 1.f / am;
  int ia = (int)((p * inv + 1.f) * (0.5f * BIN_B)), ib = (int)((q * inv + 1.f) * (0.5f * BIN_B));
  ia = ia < 0 ? 0 : (ia > BIN_B - 1 ? BIN_B - 1 : ia); ib = ib < 0 ? 0 : (ib > BIN_B - 1 ? BIN_B - 1 : ib);
  return ((m * 2 + (um < 0.f ? 1 : 0)) * BIN_B + ia) * BIN_B + ib;
}

// does bin b list face f?  bounding cap of the bin's cube-map cell (centre direction, largest angle to a corner) against the
// bounding cap of the face's spherical triangle (axis = sum of the unit vertex directions, largest angle to a vertex), widened
// by 0.05 rad.  Host + device: tests/hostcheck checks the superset property on the CPU with this very function.
SD3_HD inline bool bin_takes_face(int b, const float* verts, const int* faces, int f) {
  const int ib = b % BIN_B, ia = (b / BIN_B) % BIN_B, ms = b / (BIN_B * BIN_B), m = ms >> 1;
  const double sgn = (ms & 1) ? -1.0 : 1.0;
  const double a0 = -1.0 + 2.0 * ia / BIN_B, a1 = -1.0 + 2.0 * (ia + 1) / BIN_B, b0 = -1.0 + 2.0 * ib / BIN_B, b1 = -1.0 + 2.0 * (ib + 1) / BIN_B;
  const double ca[5] = {0.5 * (a0 + a1), a0, a0, a1, a1}, cb[5] = {0.5 * (b0 + b1), b0, b1, b0, b1};
  double c[3] = {0, 0, 0}, hb = 0;
  for (int q = 0; q < 5; ++q) {
    double v[3], k[3];
    v[m] = sgn; v[m == 0 ? 1 : 0] = ca[q]; v[m == 2 ? 1 : 2] = cb[q];
    const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    k[0] = v[0] / l; k[1] = v[1] / l; k[2] = v[2] / l;
    if (q == 0) { c[0] = k[0]; c[1] = k[1]; c[2] = k[2]; }
    else { const double dt = c[0] * k[0] + c[1] * k[1] + c[2] * k[2]; const double an = acos(dt < 1.0 ? dt : 1.0); hb = an > hb ? an : hb; }
  }
  double n[3][3], ax[3] = {0, 0, 0};
  bool ok = true;
  for (int e = 0; e < 3; ++e) {
    const float* v = verts + 3 * faces[3 * f + e];
    const double l = sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
    if (!(l > 0)) ok = false;
    for (int d = 0; d < 3; ++d) { n[e][d] = v[d] / (l > 0 ? l : 1.0); ax[d] += n[e][d]; }
  }
  const double la = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  if (!(ok && la > 1e-6)) return true;
  // a face whose three directions are (nearly) coplanar with the centre -- Rays_Cartesian's pole faces, two of whose rays
  // coincide -- spans a degenerate tetrahedron: all four determinants of inside_tetrahedron vanish on a whole PLANE through
  // the centre, and the reference's `det >= 0` then answers true for voxels far from the face's direction.  Such a face is
  // listed in every bin, so the binned OR stays the OR over all faces.
  const double det = n[0][0] * (n[1][1] * n[2][2] - n[1][2] * n[2][1]) - n[0][1] * (n[1][0] * n[2][2] - n[1][2] * n[2][0]) +
                     n[0][2] * (n[1][0] * n[2][1] - n[1][1] * n[2][0]);
  if (fabs(det) < 1e-6) return true;
  auto clampc = [](double x) { return x > 1.0 ? 1.0 : (x < -1.0 ? -1.0 : x); };
  double hf = 0;
  for (int e = 0; e < 3; ++e) { const double an = acos(clampc((ax[0] * n[e][0] + ax[1] * n[e][1] + ax[2] * n[e][2]) / la)); hf = an > hf ? an : hf; }
  const double ang = acos(clampc((ax[0] * c[0] + ax[1] * c[1] + ax[2] * c[2]) / la));
  return ang <= hf + hb + 0.05;
}

#if defined(__CUDACC__)
// one block per bin, the faces across its threads (the order of a bin's list is irrelevant: the lists feed an OR / an AND)
static __global__ void __launch_bounds__(128) k_build_bins(const float* __restrict__ verts, const int* __restrict__ faces, int n_faces, FaceBins B) {
  const int b = blockIdx.x;
  if (b >= BIN_N) return;
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int f = threadIdx.x; f < n_faces; f += blockDim.x)
    if (bin_takes_face(b, verts, faces, f)) { const int pos = atomicAdd(&s_cnt, 1); if (pos < BIN_CAP) B.faces[b * BIN_CAP + pos] = f; }
  __syncthreads();
  if (threadIdx.x == 0) B.count[b] = s_cnt <= BIN_CAP ? s_cnt : -1;
}

// inside_polyhedron (geom3d.cuh) restricted to the tetrahedra of the voxel's direction bin: same boolean as the full loop
__device__ __forceinline__ bool inside_polyhedron_binned(float z, float y, float x, const float* center, const float* pv, const int* faces, int n_faces,
                                                         const FaceBins& B) {
  const int b = B.count ? bin_of(z - center[0], y - center[1], x - center[2]) : -1;
  const int cnt = b >= 0 ? __ldg(B.count + b) : -1;
  if (cnt < 0) return sd3::inside_polyhedron(z, y, x, center, pv, faces, n_faces);
  const int* lst = B.faces + b * BIN_CAP;
  for (int i = 0; i < cnt; ++i) {
    const int f = __ldg(lst + i);
    const int iA = faces[3 * f], iB = faces[3 * f + 1], iC = faces[3 * f + 2];
    if (sd3::inside_tetrahedron(z, y, x, center[0], center[1], center[2], pv[3 * iA], pv[3 * iA + 1], pv[3 * iA + 2],
                                pv[3 * iB], pv[3 * iB + 1], pv[3 * iB + 2], pv[3 * iC], pv[3 * iC + 1], pv[3 * iC + 2])) return true;
  }
  return false;
}

#endif  // __CUDACC__

}  // namespace sdbins
