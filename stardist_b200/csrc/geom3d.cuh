// geom3d.cuh -- star-convex polyhedron geometry for the 3D NMS / label painting (host + device).
//
// Reference: stardist/lib/stardist3d_impl.cpp.  Two classes of routines:
//
//  (E) EXACT float32 restatements -- same operations in the same order, no FMA (the TU that includes
//      this header for the device must be compiled with -fmad=false):
//        inside_halfspace :89-106, inside_tetrahedron :109-148, inside_polyhedron :153-190,
//        tetrahedron_volume :234-253, polyhedron_volume :257-291, bounding_radius_* :343-467,
//        intersect_sphere_isotropic :494-520, intersect_bbox :523-532, polyhedron_bbox :536-567,
//        polyhedron_polyverts :570-585, build_halfspace :744-764, point_in_halfspaces :818-827.
//      Implicit promotions of the reference are reproduced (e.g. `1.f/(sqrt(..)+1.e-10)` is a double
//      division, `M_PI*4.f/3*r*r*r` is double arithmetic, `fmin(1.f, A/(B+1e-10))` is double).
//
//  (Q) Replacements for the two Qhull calls (qhull_overlap_kernel :830-869 -> "halfspace
//      intersection volume about the centre midpoint", qhull_overlap_convex_hulls :872-939 -> the same
//      on the facet planes of the two convex hulls).  Qhull (vendored 2018.0.1.r, ~30 kLoC) is not
//      restated; these are independent double-precision computations of the same quantities:
//        * feasibility test exactly as qh_sethalfspace (geom2_r.c:1937-1943): dist = offset + n.p
//          accumulated in that order; dist > 0 -> error -> err_value (0 for the kernel stage,
//          1e10 for the hull stage);
//        * volume = sum over planes of (1/3) * h_k * area(face_k), face_k = plane k clipped by all
//          other halfspaces (2D Sutherland-Hodgman in the plane's basis);
//        * convex hull facets by gift wrapping.
//      They agree with Qhull to ~1e-12 relative; the result is rounded to float like the reference's
//      return type, so a decision can only differ from the reference on a pair whose IoU sits within
//      one float ulp of the threshold, or where Qhull itself reports a precision error.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define SD3_HD __host__ __device__
#else
#define SD3_HD
#endif

namespace sd3 {

// ------------------------------------------------------------------------------------------ (E)
SD3_HD inline bool inside_halfspace(float z, float y, float x, float Az, float Ay, float Ax,
                                    float Bz, float By, float Bx, float Cz, float Cy, float Cx) {
  const float M00 = Bz - Az, M01 = By - Ay, M02 = Bx - Ax;
  const float M10 = Cz - Az, M11 = Cy - Ay, M12 = Cx - Ax;
  const float M20 = z - Az, M21 = y - Ay, M22 = x - Ax;
  const float det = M00 * (M11 * M22 - M21 * M12) - M01 * (M10 * M22 - M12 * M20) + M02 * (M10 * M21 - M11 * M20);
  return det >= 0;
}

SD3_HD inline bool inside_tetrahedron(float z, float y, float x, float Rz, float Ry, float Rx,
                                      float Az, float Ay, float Ax, float Bz, float By, float Bx,
                                      float Cz, float Cy, float Cx) {
  return inside_halfspace(z, y, x, Az, Ay, Ax, Bz, By, Bx, Cz, Cy, Cx) &&
         inside_halfspace(z, y, x, Rz, Ry, Rx, Bz, By, Bx, Az, Ay, Ax) &&
         inside_halfspace(z, y, x, Rz, Ry, Rx, Cz, Cy, Cx, Bz, By, Bx) &&
         inside_halfspace(z, y, x, Rz, Ry, Rx, Az, Ay, Ax, Cz, Cy, Cx);
}

// pv: [n_rays][3] polyhedron vertices (z,y,x); faces: [n_faces][3]
SD3_HD inline bool inside_polyhedron(float z, float y, float x, const float* center, const float* pv,
                                     const int* faces, int n_faces) {
  const float Rz = center[0], Ry = center[1], Rx = center[2];
  for (int i = 0; i < n_faces; ++i) {
    const int iA = faces[3 * i], iB = faces[3 * i + 1], iC = faces[3 * i + 2];
    if (inside_tetrahedron(z, y, x, Rz, Ry, Rx, pv[3 * iA], pv[3 * iA + 1], pv[3 * iA + 2],
                           pv[3 * iB], pv[3 * iB + 1], pv[3 * iB + 2], pv[3 * iC], pv[3 * iC + 1], pv[3 * iC + 2]))
      return true;
  }
  return false;
}

SD3_HD inline float tetrahedron_volume0(float Az, float Ay, float Ax, float Bz, float By, float Bx,
                                        float Cz, float Cy, float Cx) {
  // tetrahedron_volume(R = 0,0,0, A, B, C)
  const float M00 = Bz - Az, M01 = By - Ay, M02 = Bx - Ax;
  const float M10 = Cz - Az, M11 = Cy - Ay, M12 = Cx - Ax;
  const float M20 = 0.f - Az, M21 = 0.f - Ay, M22 = 0.f - Ax;
  const float det = M00 * (M11 * M22 - M21 * M12) - M01 * (M10 * M22 - M12 * M20) + M02 * (M10 * M21 - M11 * M20);
  return det / 6.f;
}

SD3_HD inline float polyhedron_volume(const float* dist, const float* verts, const int* faces, int n_faces) {
  float vol = 0.f;
  for (int i = 0; i < n_faces; ++i) {
    const int iA = faces[3 * i], iB = faces[3 * i + 1], iC = faces[3 * i + 2];
    const float Az = dist[iA] * verts[3 * iA], Ay = dist[iA] * verts[3 * iA + 1], Ax = dist[iA] * verts[3 * iA + 2];
    const float Bz = dist[iB] * verts[3 * iB], By = dist[iB] * verts[3 * iB + 1], Bx = dist[iB] * verts[3 * iB + 2];
    const float Cz = dist[iC] * verts[3 * iC], Cy = dist[iC] * verts[3 * iC + 1], Cx = dist[iC] * verts[3 * iC + 2];
    vol += tetrahedron_volume0(Az, Ay, Ax, Bz, By, Bx, Cz, Cy, Cx);
  }
  return vol;
}

SD3_HD inline int round_to_int(float r) {
#if defined(__CUDA_ARCH__)
  return __float2int_rn(r);        // lrint: round half to even
#else
  return (int)lrintf(r);
#endif
}

SD3_HD inline void polyhedron_bbox(const float* dist, const float* center, const float* verts, int n_rays, int* bbox) {
  int z1 = INT32_MAX, z2 = -1, y1 = INT32_MAX, y2 = -1, x1 = INT32_MAX, x2 = -1;
  for (int j = 0; j < n_rays; ++j) {
    const float z = center[0] + dist[j] * verts[3 * j];
    const float y = center[1] + dist[j] * verts[3 * j + 1];
    const float x = center[2] + dist[j] * verts[3 * j + 2];
    const int iz = round_to_int(z), iy = round_to_int(y), ix = round_to_int(x);
    z1 = iz < z1 ? iz : z1; z2 = iz > z2 ? iz : z2;
    y1 = iy < y1 ? iy : y1; y2 = iy > y2 ? iy : y2;
    x1 = ix < x1 ? ix : x1; x2 = ix > x2 ? ix : x2;
  }
  bbox[0] = z1; bbox[1] = z2; bbox[2] = y1; bbox[3] = y2; bbox[4] = x1; bbox[5] = x2;
}

SD3_HD inline void polyhedron_polyverts(const float* dist, const float* center, const float* verts, int n_rays, float* pv) {
  for (int j = 0; j < n_rays; ++j) {
    pv[3 * j] = center[0] + dist[j] * verts[3 * j];
    pv[3 * j + 1] = center[1] + dist[j] * verts[3 * j + 1];
    pv[3 * j + 2] = center[2] + dist[j] * verts[3 * j + 2];
  }
}

SD3_HD inline float bounding_radius_outer(const float* dist, int n_rays) {
  float r = 0;
  for (int i = 0; i < n_rays; ++i) r = fmaxf(r, dist[i]);
  return r;
}

SD3_HD inline float bounding_radius_outer_isotropic(const float* dist, const float* verts, int n_rays, const float* aniso) {
  float r2max = 0;
  for (int i = 0; i < n_rays; ++i) {
    const float z = aniso[0] * dist[i] * verts[3 * i];
    const float y = aniso[1] * dist[i] * verts[3 * i + 1];
    const float x = aniso[2] * dist[i] * verts[3 * i + 2];
    const float r2 = z * z + y * y + x * x;
    r2max = fmaxf(r2, r2max);
  }
  return sqrtf(r2max);
}

SD3_HD inline float bounding_radius_inner_isotropic(const float* dist, const float* verts, const int* faces,
                                                    int n_faces, const float* aniso) {
  float r_min = INFINITY;
  for (int i = 0; i < n_faces; ++i) {
    const int iA = faces[3 * i], iB = faces[3 * i + 1], iC = faces[3 * i + 2];
    const float Az = aniso[0] * dist[iA] * verts[3 * iA], Ay = aniso[1] * dist[iA] * verts[3 * iA + 1], Ax = aniso[2] * dist[iA] * verts[3 * iA + 2];
    const float Bz = aniso[0] * dist[iB] * verts[3 * iB], By = aniso[1] * dist[iB] * verts[3 * iB + 1], Bx = aniso[2] * dist[iB] * verts[3 * iB + 2];
    const float Cz = aniso[0] * dist[iC] * verts[3 * iC], Cy = aniso[1] * dist[iC] * verts[3 * iC + 1], Cx = aniso[2] * dist[iC] * verts[3 * iC + 2];
    const float pz = Bz - Az, py = By - Ay, px = Bx - Ax;
    const float qz = Cz - Az, qy = Cy - Ay, qx = Cx - Ax;
    float Nz = (px * qy - py * qx);
    float Ny = (pz * qx - px * qz);
    float Nx = (py * qz - pz * qy);
    // float normz = 1.f/(sqrt(Nz*Nz+Ny*Ny+Nx*Nx)+1.e-10);  sqrt(float) -> float; + double; 1.f/double
    const float normz = (float)(1.0 / ((double)sqrtf(Nz * Nz + Ny * Ny + Nx * Nx) + 1.e-10));
    Nz *= normz; Ny *= normz; Nx *= normz;
    const float r = Az * Nz + Ay * Ny + Ax * Nx;
    r_min = fminf(r_min, r);
  }
  return r_min;
}

SD3_HD inline float intersect_sphere_isotropic(float r1, const float* p1, float r2, const float* p2, const float* an) {
  const float dz = an[0] * (p1[0] - p2[0]);
  const float dy = an[1] * (p1[1] - p2[1]);
  const float dx = an[2] * (p1[2] - p2[2]);
  const float d = sqrtf(dz * dz + dy * dy + dx * dx);
  const float rmin = fminf(r1, r2), rmax = fmaxf(r1, r2);
  if (d > (r1 + r2)) return 0;
  // if (rmax >= d+rmin-1.e-10) return M_PI*4.f/3*rmin*rmin*rmin;   (double arithmetic)
  if ((double)rmax >= (double)(d + rmin) - 1.e-10)
    return (float)(M_PI * 4.f / 3 * rmin * rmin * rmin);
  const float t = (r1 + r2 - d) / 2 / d;
  const float h1 = (r2 - r1 + d) * t;
  const float h2 = (r1 - r2 + d) * t;
  const float v1 = (float)(M_PI / 3 * h1 * h1 * (3 * r1 - h1));
  const float v2 = (float)(M_PI / 3 * h2 * h2 * (3 * r2 - h2));
  return (v1 + v2) / (an[0] * an[1] * an[2]);
}

SD3_HD inline float intersect_bbox(const int* b1, const int* b2) {
  // fmax(0, fmin(int,int) - fmax(int,int)) : double arithmetic on ints, stored to float
  const double wz0 = fmin((double)b1[1], (double)b2[1]) - fmax((double)b1[0], (double)b2[0]);
  const double wy0 = fmin((double)b1[3], (double)b2[3]) - fmax((double)b1[2], (double)b2[2]);
  const double wx0 = fmin((double)b1[5], (double)b2[5]) - fmax((double)b1[4], (double)b2[4]);
  const float wz = (float)fmax(0.0, wz0), wy = (float)fmax(0.0, wy0), wx = (float)fmax(0.0, wx0);
  return wx * wy * wz;
}

// build_halfspace :744-764 : float normal/offset, stored as double
SD3_HD inline void build_halfspace(const float* A, const float* B, const float* C, double* hs) {
  const float Az = A[0], Ay = A[1], Ax = A[2];
  const float Pz = B[0] - Az, Py = B[1] - Ay, Px = B[2] - Ax;
  const float Qz = C[0] - Az, Qy = C[1] - Ay, Qx = C[2] - Ax;
  const float Nz = -(Py * Qx - Px * Qy);
  const float Ny = -(Px * Qz - Pz * Qx);
  const float Nx = -(Pz * Qy - Py * Qz);
  hs[0] = Nz; hs[1] = Ny; hs[2] = Nx;
  hs[3] = -(Az * Nz + Ay * Ny + Ax * Nx);
}

// point_in_halfspaces on the kernel halfspaces of a polyhedron (:799-827): outside iff any plane > 0
SD3_HD inline bool inside_kernel(float z, float y, float x, const float* pv, const int* faces, int n_faces) {
  for (int i = 0; i < n_faces; ++i) {
    double hs[4];
    build_halfspace(&pv[3 * faces[3 * i]], &pv[3 * faces[3 * i + 1]], &pv[3 * faces[3 * i + 2]], hs);
    if (hs[0] * z + hs[1] * y + hs[2] * x + hs[3] > 0) return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------ (Q)
struct Plane { double n0, n1, n2, d; };     // n.x + d <= 0 is inside

// qh_sethalfspace feasibility (geom2_r.c:1937-1943): true when p is "clearly inside"
SD3_HD inline bool plane_feasible(const Plane& P, const double* p) {
  double dist = P.d;
  dist += P.n0 * p[0]; dist += P.n1 * p[1]; dist += P.n2 * p[2];
  return dist < 0;
}

constexpr int SD3_MAXPOLY = 48;

// Order in which face_cone_volume clips face k against the other planes.  Attempt 0 is the index order.  The intermediate
// polygon is face k of the polytope of the planes seen so far; in index order its vertex count can pass SD3_MAXPOLY although
// the final face has a handful (hull facets come out of the gift wrapping as a growing patch: a far face is bounded by the
// whole rim of the patch -- observed from 128 rays on, never for <= 96).  When that happens the face is clipped again in a
// scattered order, t -> (t * stride) mod n with a stride coprime to n near n / phi (then n / phi^2): the planes seen so far
// are spread over the sphere and the intermediate faces stay small.  Faces that do not overflow are untouched, bit for bit.
SD3_HD inline int clip_order_stride(int n, int attempt) {
  if (attempt == 0 || n < 3) return 1;
  int s = (int)((double)n * (attempt == 1 ? 0.6180339887498949 : 0.3819660112501051));
  if (s < 2) s = 2;
  for (int guard = 0; guard < n; ++guard, ++s) {
    if (s >= n) s = 2;
    int a = s, b = n;
    while (b) { const int r = a % b; a = b; b = r; }
    if (a == 1) return s;
  }
  return 1;
}
constexpr int SD3_CLIP_ATTEMPTS = 3;

// (1/3) * h_k * area(face k) of the polytope {x : planes[j].x + d_j <= 0 for all j}, p strictly inside.
// L = half edge of the initial square (any bound on the polytope's extent around p).
template <typename PlaneAt>
SD3_HD inline double face_cone_volume(const PlaneAt& planes, int n, int k, const double* p, double L, int* overflow) {
  Plane Pk = planes(k);
  const double len = sqrt(Pk.n0 * Pk.n0 + Pk.n1 * Pk.n1 + Pk.n2 * Pk.n2);
  if (!(len > 0)) return 0.0;
  const double nz = Pk.n0 / len, ny = Pk.n1 / len, nx = Pk.n2 / len, dk = Pk.d / len;
  const double sdist = nz * p[0] + ny * p[1] + nx * p[2] + dk;      // < 0
  const double h = -sdist;
  const double q0 = p[0] - sdist * nz, q1 = p[1] - sdist * ny, q2 = p[2] - sdist * nx;   // foot point
  // orthonormal basis (u, v) of the plane
  double u0, u1, u2;
  if (fabs(nz) <= fabs(ny) && fabs(nz) <= fabs(nx)) { u0 = 0; u1 = -nx; u2 = ny; }
  else if (fabs(ny) <= fabs(nx)) { u0 = -nx; u1 = 0; u2 = nz; }
  else { u0 = -ny; u1 = nz; u2 = 0; }
  const double ul = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
  u0 /= ul; u1 /= ul; u2 /= ul;
  const double v0 = ny * u2 - nx * u1, v1 = nx * u0 - nz * u2, v2 = nz * u1 - ny * u0;
  double pa[SD3_MAXPOLY], pb[SD3_MAXPOLY], qa[SD3_MAXPOLY], qb[SD3_MAXPOLY];
  int m = 4, ovf = 0;
  const double eps_dup = 1e-9;
  for (int attempt = 0; attempt < SD3_CLIP_ATTEMPTS; ++attempt) {
  m = 4; ovf = 0;
  pa[0] = -L; pb[0] = -L; pa[1] = L; pb[1] = -L; pa[2] = L; pb[2] = L; pa[3] = -L; pb[3] = L;
  const int stride = clip_order_stride(n, attempt);
  for (int jt = 0; jt < n && m > 0 && !ovf; ++jt) {
    const int j = attempt == 0 ? jt : (int)(((long long)jt * stride) % n);
    if (j == k) continue;
    Plane Pj = planes(j);
    const double lj = sqrt(Pj.n0 * Pj.n0 + Pj.n1 * Pj.n1 + Pj.n2 * Pj.n2);
    if (!(lj > 0)) continue;
    const double a0 = Pj.n0 / lj, a1 = Pj.n1 / lj, a2 = Pj.n2 / lj, dj = Pj.d / lj;
    const double A = a0 * u0 + a1 * u1 + a2 * u2;
    const double B = a0 * v0 + a1 * v1 + a2 * v2;
    const double C = a0 * q0 + a1 * q1 + a2 * q2 + dj;
    const double g = sqrt(A * A + B * B);
    if (g < 1e-12) {
      // parallel planes: j cuts the whole face, nothing, or is a duplicate of k (owned by the lower index)
      if (C > eps_dup * (1.0 + fabs(dj))) { m = 0; break; }
      if (fabs(C) <= eps_dup * (1.0 + fabs(dj)) && (a0 * nz + a1 * ny + a2 * nx) > 0 && j < k) { m = 0; break; }
      continue;
    }
    // Sutherland-Hodgman against A*a + B*b + C <= 0
    int mo = 0;
    double sa = pa[m - 1], sb = pb[m - 1];
    double fs = A * sa + B * sb + C;
    for (int t = 0; t < m; ++t) {
      const double ea = pa[t], eb = pb[t];
      const double fe = A * ea + B * eb + C;
      if (fe <= 0) {
        if (fs > 0) {
          const double w = fs / (fs - fe);
          if (mo < SD3_MAXPOLY) { qa[mo] = sa + w * (ea - sa); qb[mo] = sb + w * (eb - sb); mo++; } else ovf = 1;
        }
        if (mo < SD3_MAXPOLY) { qa[mo] = ea; qb[mo] = eb; mo++; } else ovf = 1;
      } else if (fs <= 0) {
        const double w = fs / (fs - fe);
        if (mo < SD3_MAXPOLY) { qa[mo] = sa + w * (ea - sa); qb[mo] = sb + w * (eb - sb); mo++; } else ovf = 1;
      }
      sa = ea; sb = eb; fs = fe;
    }
    m = mo;
    for (int t = 0; t < m; ++t) { pa[t] = qa[t]; pb[t] = qb[t]; }
  }
  if (!ovf) break;
  }
  if (ovf) *overflow = 1;      // every order overflowed: the value below is an under-estimate
  if (m < 3) return 0.0;
  double area2 = 0;
  for (int t = 0; t < m; ++t) {
    const int t2 = (t + 1 == m) ? 0 : t + 1;
    area2 += pa[t] * pb[t2] - pa[t2] * pb[t];
  }
  return fabs(area2) * 0.5 * h / 3.0;
}

// ---- the same volume with the per-plane work hoisted out of the (k, j) loop -----------------------------------------
// face_cone_volume re-derives  n/|n|, d/|n|  of plane j (one sqrt, four divisions) for every (k, j) pair and takes
// another square root for the parallel test: ~2F * 2F * (2 sqrt + 4 div) in double per pair of polyhedra.
// normalized_plane() performs exactly those operations ONCE per plane; face_cone_volume_n() consumes the results.
// Every double it produces is bit-identical to face_cone_volume's (tests/test_cpu_oracle.py compares the two on fuzzed
// pairs, host build):
//   * the scaled plane is the same four IEEE operations on the same operands;
//   * `sqrt(s) < 1e-12` is replaced by `s < 1e-24`: sqrt is correctly rounded (host libm, CUDA sqrt.rn.f64), hence monotone,
//     and 1e-24 (= 0x1.357c299a88ea7p-80) is the smallest double whose square root is not below 1e-12;
//   * a plane with no polygon vertex on its outer side leaves Sutherland-Hodgman's output equal to its input (same
//     vertices, same order), so that pass -- the common case once the face has shrunk -- is skipped.
// A plane with !(|n| > 0) is stored as all zeros and skipped (face_cone_volume: `continue` / `return 0`).
SD3_HD inline Plane normalized_plane(const Plane& P) {
  const double len = sqrt(P.n0 * P.n0 + P.n1 * P.n1 + P.n2 * P.n2);
  Plane Q;
  if (!(len > 0)) { Q.n0 = 0; Q.n1 = 0; Q.n2 = 0; Q.d = 0; return Q; }
  Q.n0 = P.n0 / len; Q.n1 = P.n1 / len; Q.n2 = P.n2 / len; Q.d = P.d / len;
  return Q;
}

template <typename PlaneAt>
SD3_HD inline double face_cone_volume_n(const PlaneAt& planes /* normalized_plane() of each */, int n, int k, const double* p,
                                        double L, int* overflow) {
  const Plane Pk = planes(k);
  if (Pk.n0 == 0 && Pk.n1 == 0 && Pk.n2 == 0) return 0.0;
  const double nz = Pk.n0, ny = Pk.n1, nx = Pk.n2, dk = Pk.d;
  const double sdist = nz * p[0] + ny * p[1] + nx * p[2] + dk;      // < 0
  const double h = -sdist;
  const double q0 = p[0] - sdist * nz, q1 = p[1] - sdist * ny, q2 = p[2] - sdist * nx;   // foot point
  double u0, u1, u2;
  if (fabs(nz) <= fabs(ny) && fabs(nz) <= fabs(nx)) { u0 = 0; u1 = -nx; u2 = ny; }
  else if (fabs(ny) <= fabs(nx)) { u0 = -nx; u1 = 0; u2 = nz; }
  else { u0 = -ny; u1 = nz; u2 = 0; }
  const double ul = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
  u0 /= ul; u1 /= ul; u2 /= ul;
  const double v0 = ny * u2 - nx * u1, v1 = nx * u0 - nz * u2, v2 = nz * u1 - ny * u0;
  double pa[SD3_MAXPOLY], pb[SD3_MAXPOLY], qa[SD3_MAXPOLY], qb[SD3_MAXPOLY];
  int m = 4, ovf = 0;
  const double eps_dup = 1e-9;
  for (int attempt = 0; attempt < SD3_CLIP_ATTEMPTS; ++attempt) {
  m = 4; ovf = 0;
  pa[0] = -L; pb[0] = -L; pa[1] = L; pb[1] = -L; pa[2] = L; pb[2] = L; pa[3] = -L; pb[3] = L;
  const int stride = clip_order_stride(n, attempt);
  for (int jt = 0; jt < n && m > 0 && !ovf; ++jt) {
    const int j = attempt == 0 ? jt : (int)(((long long)jt * stride) % n);
    if (j == k) continue;
    const Plane Pj = planes(j);
    if (Pj.n0 == 0 && Pj.n1 == 0 && Pj.n2 == 0) continue;
    const double a0 = Pj.n0, a1 = Pj.n1, a2 = Pj.n2, dj = Pj.d;
    const double A = a0 * u0 + a1 * u1 + a2 * u2;
    const double B = a0 * v0 + a1 * v1 + a2 * v2;
    const double C = a0 * q0 + a1 * q1 + a2 * q2 + dj;
    if (A * A + B * B < 1e-24) {
      if (C > eps_dup * (1.0 + fabs(dj))) { m = 0; break; }
      if (fabs(C) <= eps_dup * (1.0 + fabs(dj)) && (a0 * nz + a1 * ny + a2 * nx) > 0 && j < k) { m = 0; break; }
      continue;
    }
    bool any_out = false;
    for (int t = 0; t < m; ++t) any_out = any_out || (A * pa[t] + B * pb[t] + C > 0);
    if (!any_out) continue;
    int mo = 0;
    double sa = pa[m - 1], sb = pb[m - 1];
    double fs = A * sa + B * sb + C;
    for (int t = 0; t < m; ++t) {
      const double ea = pa[t], eb = pb[t];
      const double fe = A * ea + B * eb + C;
      if (fe <= 0) {
        if (fs > 0) {
          const double w = fs / (fs - fe);
          if (mo < SD3_MAXPOLY) { qa[mo] = sa + w * (ea - sa); qb[mo] = sb + w * (eb - sb); mo++; } else ovf = 1;
        }
        if (mo < SD3_MAXPOLY) { qa[mo] = ea; qb[mo] = eb; mo++; } else ovf = 1;
      } else if (fs <= 0) {
        const double w = fs / (fs - fe);
        if (mo < SD3_MAXPOLY) { qa[mo] = sa + w * (ea - sa); qb[mo] = sb + w * (eb - sb); mo++; } else ovf = 1;
      }
      sa = ea; sb = eb; fs = fe;
    }
    m = mo;
    for (int t = 0; t < m; ++t) { pa[t] = qa[t]; pb[t] = qb[t]; }
  }
  if (!ovf) break;
  }
  if (ovf) *overflow = 1;      // every order overflowed: the value below is an under-estimate
  if (m < 3) return 0.0;
  double area2 = 0;
  for (int t = 0; t < m; ++t) {
    const int t2 = (t + 1 == m) ? 0 : t + 1;
    area2 += pa[t] * pb[t2] - pa[t2] * pb[t];
  }
  return fabs(area2) * 0.5 * h / 3.0;
}

// Duplicate points break the gift wrapping below (a zero-length pivot edge), Qhull -- the reference -- ignores them.  They only
// occur for ray sets with coincident directions (Rays_Cartesian's pole rings) and equal distances on those rays.  Every point
// that repeats a lower-indexed one is moved to the centroid of the set: an interior point never becomes a hull vertex, so the
// facets are those of the distinct points.  Points without a twin are not touched.  Serial, O(n^2) compares; returns the count.
SD3_HD inline int demote_duplicate_points(double* pts, int n) {
  double c0 = 0, c1 = 0, c2 = 0;
  for (int i = 0; i < n; ++i) { c0 += pts[3 * i]; c1 += pts[3 * i + 1]; c2 += pts[3 * i + 2]; }
  c0 /= n; c1 /= n; c2 /= n;
  int moved = 0;
  for (int i = n - 1; i > 0; --i) {              // downwards: the lower-indexed twins are still in place
    bool twin = false;
    for (int j = 0; j < i && !twin; ++j) twin = pts[3 * j] == pts[3 * i] && pts[3 * j + 1] == pts[3 * i + 1] && pts[3 * j + 2] == pts[3 * i + 2];
    if (twin) { pts[3 * i] = c0; pts[3 * i + 1] = c1; pts[3 * i + 2] = c2; ++moved; }
  }
  return moved;
}

// true when two of the n ray directions coincide (relative 1e-5): only then can a polyhedron have duplicate vertices
inline bool rays_have_coincident_directions(const float* verts, int n) {
  for (int i = 1; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      const double d0 = (double)verts[3 * i] - verts[3 * j], d1 = (double)verts[3 * i + 1] - verts[3 * j + 1], d2 = (double)verts[3 * i + 2] - verts[3 * j + 2];
      const double l2 = (double)verts[3 * i] * verts[3 * i] + (double)verts[3 * i + 1] * verts[3 * i + 1] + (double)verts[3 * i + 2] * verts[3 * i + 2];
      if (d0 * d0 + d1 * d1 + d2 * d2 <= 1e-10 * l2) return true;
    }
  return false;
}

// A face of the ray triangulation whose three directions are coplanar with the centre (|det of the unit directions| < 1e-6;
// Rays_Cartesian's pole faces, two of whose rays coincide).  Its tetrahedron (centre, A, B, C) is degenerate: all four
// determinants of inside_tetrahedron vanish on a whole plane through the centre, where `det >= 0` answers true.
SD3_HD inline bool ray_face_is_degenerate(const float* verts, const int* faces, int f) {
  double n[3][3];
  for (int e = 0; e < 3; ++e) {
    const float* v = verts + 3 * faces[3 * f + e];
    const double l = sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
    if (!(l > 0)) return true;
    n[e][0] = v[0] / l; n[e][1] = v[1] / l; n[e][2] = v[2] / l;
  }
  const double det = n[0][0] * (n[1][1] * n[2][2] - n[1][2] * n[2][1]) - n[0][1] * (n[1][0] * n[2][2] - n[1][2] * n[2][0]) +
                     n[0][2] * (n[1][0] * n[2][1] - n[1][1] * n[2][0]);
  return fabs(det) < 1e-6;
}

// Convex hull facet planes (outward, unit normal) of n points (double, [n][3]) by gift wrapping.
// out: up to max_planes planes; returns the number of facets, or -1 on failure (degenerate input).
// scratch: edge_done bit matrix n*n bits (uint32 words), stack of directed edges (int16 triples).
SD3_HD inline int convex_hull_planes(const double* pts, int n, Plane* out, int max_planes,
                                     uint32_t* edge_done /* (n*n+31)/32 words */, int16_t* stack /* 3*max_edges */,
                                     int max_stack) {
  if (n < 4) return -1;
  for (int i = 0; i < (n * n + 31) / 32; ++i) edge_done[i] = 0;
#define SD3_P(i, c) pts[3 * (i) + (c)]
  // orientation of d relative to triangle (a,b,c): > 0 if d is on the side the normal (b-a)x(c-a) points to
  auto orient = [&](int a, int b, int c, int d) -> double {
    const double b0 = SD3_P(b, 0) - SD3_P(a, 0), b1 = SD3_P(b, 1) - SD3_P(a, 1), b2 = SD3_P(b, 2) - SD3_P(a, 2);
    const double c0 = SD3_P(c, 0) - SD3_P(a, 0), c1 = SD3_P(c, 1) - SD3_P(a, 1), c2 = SD3_P(c, 2) - SD3_P(a, 2);
    const double d0 = SD3_P(d, 0) - SD3_P(a, 0), d1 = SD3_P(d, 1) - SD3_P(a, 1), d2 = SD3_P(d, 2) - SD3_P(a, 2);
    return d0 * (b1 * c2 - b2 * c1) + d1 * (b2 * c0 - b0 * c2) + d2 * (b0 * c1 - b1 * c0);
  };
  // first point: lexicographic minimum
  int p0 = 0;
  for (int i = 1; i < n; ++i) {
    if (SD3_P(i, 0) < SD3_P(p0, 0) || (SD3_P(i, 0) == SD3_P(p0, 0) && (SD3_P(i, 1) < SD3_P(p0, 1) ||
        (SD3_P(i, 1) == SD3_P(p0, 1) && SD3_P(i, 2) < SD3_P(p0, 2))))) p0 = i;
  }
  // second point: hull edge from p0 -- minimise the angle to the plane coord0 = const, i.e. maximise
  // direction "flatness": choose p1 such that all others are on one side of the plane through p0,p1
  // parallel to axis 2 ... done by 2D gift wrapping in the (0,1) projection with tie-break on distance.
  int p1 = -1;
  for (int i = 0; i < n; ++i) {
    if (i == p0) continue;
    if (p1 < 0) { p1 = i; continue; }
    const double ax = SD3_P(p1, 0) - SD3_P(p0, 0), ay = SD3_P(p1, 1) - SD3_P(p0, 1);
    const double bx = SD3_P(i, 0) - SD3_P(p0, 0), by = SD3_P(i, 1) - SD3_P(p0, 1);
    const double cr = ax * by - ay * bx;
    if (cr < 0 || (cr == 0 && (bx * bx + by * by) > (ax * ax + ay * ay))) p1 = i;
  }
  if (p1 < 0) return -1;
  // third point: pivot around (p0,p1)
  auto pivot = [&](int a, int b, int skip) -> int {
    int q = -1;
    for (int r = 0; r < n; ++r) {
      if (r == a || r == b || r == skip) continue;
      if (q < 0) { q = r; continue; }
      if (orient(a, b, q, r) > 0) q = r;      // r is outside the half-space left of (a,b,q): turn further
    }
    return q;
  };
  int p2 = pivot(p0, p1, -1);
  if (p2 < 0) return -1;
  // make (p0,p1,p2) outward: all other points must have orient <= 0
  {
    int pos = 0, neg = 0;
    for (int r = 0; r < n; ++r) { if (r == p0 || r == p1 || r == p2) continue; const double o = orient(p0, p1, p2, r); if (o > 0) pos++; else if (o < 0) neg++; }
    if (pos > 0 && neg > 0) {
      // the (0,1)-projection start edge was not a 3D hull edge pivot result; retry pivot with flipped edge
      int t = p0; p0 = p1; p1 = t;
      p2 = pivot(p0, p1, -1);
      pos = neg = 0;
      for (int r = 0; r < n; ++r) { if (r == p0 || r == p1 || r == p2) continue; const double o = orient(p0, p1, p2, r); if (o > 0) pos++; else if (o < 0) neg++; }
      if (pos > 0 && neg > 0) return -1;
    }
    if (pos > 0) { int t = p1; p1 = p2; p2 = t; }
  }
  int nf = 0, sp = 0;
  auto mark = [&](int a, int b) { const int e = a * n + b; edge_done[e >> 5] |= (1u << (e & 31)); };
  auto done = [&](int a, int b) -> bool { const int e = a * n + b; return (edge_done[e >> 5] >> (e & 31)) & 1u; };
  auto emit = [&](int a, int b, int c) -> bool {
    if (nf >= max_planes) return false;
    const double b0 = SD3_P(b, 0) - SD3_P(a, 0), b1 = SD3_P(b, 1) - SD3_P(a, 1), b2 = SD3_P(b, 2) - SD3_P(a, 2);
    const double c0 = SD3_P(c, 0) - SD3_P(a, 0), c1 = SD3_P(c, 1) - SD3_P(a, 1), c2 = SD3_P(c, 2) - SD3_P(a, 2);
    double n0 = b1 * c2 - b2 * c1, n1 = b2 * c0 - b0 * c2, n2 = b0 * c1 - b1 * c0;
    const double l = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
    if (l > 0) { n0 /= l; n1 /= l; n2 /= l; }
    out[nf].n0 = n0; out[nf].n1 = n1; out[nf].n2 = n2;
    out[nf].d = -(n0 * SD3_P(a, 0) + n1 * SD3_P(a, 1) + n2 * SD3_P(a, 2));
    nf++;
    mark(a, b); mark(b, c); mark(c, a);
    return true;
  };
  auto push = [&](int a, int b, int c) { if (sp < max_stack) { stack[3 * sp] = (int16_t)a; stack[3 * sp + 1] = (int16_t)b; stack[3 * sp + 2] = (int16_t)c; sp++; } };
  if (!emit(p0, p1, p2)) return -1;
  push(p1, p0, p2); push(p2, p1, p0); push(p0, p2, p1);      // reversed edges to cross, with the opposite vertex
  int guard = 0;
  while (sp > 0) {
    if (++guard > 16 * n + 64) return -1;
    --sp;
    const int a = stack[3 * sp], b = stack[3 * sp + 1], opp = stack[3 * sp + 2];
    if (done(a, b)) continue;
    // new facet (a,b,q) on the other side of edge (b,a): q such that all points are on the inner side
    int q = -1;
    for (int r = 0; r < n; ++r) {
      if (r == a || r == b) continue;
      if (q < 0) { if (r != opp) q = r; continue; }
      if (orient(a, b, q, r) > 0) q = r;
    }
    if (q < 0) return -1;
    // guard against picking the facet we came from when everything else is coplanar-behind
    if (!emit(a, b, q)) return -1;
    if (!done(q, b)) push(q, b, a);
    if (!done(a, q)) push(a, q, b);
  }
#undef SD3_P
  return nf;
}

}  // namespace sd3
