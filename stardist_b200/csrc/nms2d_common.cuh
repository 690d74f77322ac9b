// nms2d_common.cuh -- shared declarations of the 2D NMS translation units (see nms2d.cu).
#pragma once
#include <math.h>
#include "common.cuh"

namespace sdnms {

using sdb::cdiv;

enum { ST_UNDECIDED = 0, ST_SUPPRESSED = 1, ST_KEPT_BASE = 2 };  // kept in round r: 2 + r

struct GridDesc {
  float minx, miny, cell;
  int gx, gy;
  int all_pairs;   // use_kdtree == 0: single cell, no distance predicate
};

__device__ __forceinline__ int cell_of(float v, float mn, float cell, int g) {
  int c = (int)((v - mn) / cell);
  return c < 0 ? 0 : (c >= g ? g - 1 : c);
}

struct NmsArrays {
  const float* points; const float* radius; const float* area;
  const int4* bbox; const int2* verts;
  const unsigned int* cell_start;  // [n_cells+1]
  const int* items;
  int* state;
  int n, R;
  float max_dist, threshold;
  int use_bbox;
  GridDesc G;
  // pre-filter data (polyfast.cuh): per-edge suffix sums of F = ∫ x dy, per-polygon ∮ x dy and longest
  // edge, largest |coordinate| over all polygons; filter: 0 = exact sweep only, 1 = filter + sweep for the
  // undecided pairs, 2 = verify (both on every pair, disagreements counted)
  const double* suf; const double* sarea; const float* maxlen;
  double max_abs_coord;
  int filter;
};

// would suppressor h (higher score) test candidate c ?  (stardist2d.cpp:548-549,572-576)
__device__ __forceinline__ bool reaches(const NmsArrays& A, int h, int c, float cy, float cx, const int4& bc) {
  if (!A.G.all_pairs) {
    // nanoflann L2_Simple: sum over dims of (q - p)^2 accumulated in float, strict < radius
    const float d0 = A.points[2 * h] - cy, d1 = A.points[2 * h + 1] - cx;
    const float dd = d0 * d0 + d1 * d1;
    const float rr = A.max_dist + A.radius[h];
    if (!(dd < rr * rr)) return false;
  }
  if (A.use_bbox) {
    const int4 bh = A.bbox[h];
    if (!(bc.x <= bh.y && bh.x <= bc.y && bc.z <= bh.w && bh.z <= bc.w)) return false;
  }
  return true;
}


// pre-filter mode and cumulative statistics (nms2d.cu)
extern int g_filter_mode;
extern int g_tail_mode;      // 1 (default): rounds >= 1 in one cooperative launch (k_tail); 0: host loop only
extern unsigned long long g_filter_stats[4];   // pairs, pairs sent to the exact sweep, verify mismatches, calls

// frontier-peeling rounds, instantiated per polygon capacity in nms2d_nv32.cu / nms2d_nv128.cu
int run_rounds_nv32(NmsArrays A, int* d_slow, unsigned int* d_counters, cudaStream_t st, int verbose, unsigned int* h_pin);
int run_rounds_nv128(NmsArrays A, int* d_slow, unsigned int* d_counters, cudaStream_t st, int verbose, unsigned int* h_pin);

}  // namespace sdnms
