// nms2d_rounds.cuh -- frontier / suppress kernels of the 2D NMS, templated on the polygon
// capacity NV.  Included by exactly one TU per NV (the clipping sweep is large; separate TUs
// compile in parallel).  Must be compiled with -fmad=false.
#pragma once
#include "nms2d_common.cuh"
#include "clip2d.cuh"

namespace sdnms {
namespace {

__global__ void k_frontier(NmsArrays A, int round, unsigned int* __restrict__ counters /* [0]=undecided */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.n) return;
  if (A.state[c] != ST_UNDECIDED) return;
  atomicAdd(&counters[0], 1u);
  const float cy = A.points[2 * c], cx = A.points[2 * c + 1];
  const int4 bc = A.bbox[c];
  const int kept_now = ST_KEPT_BASE + round;
  int ccx = 0, ccy = 0;
  if (!A.G.all_pairs) { ccx = cell_of(cx, A.G.minx, A.G.cell, A.G.gx); ccy = cell_of(cy, A.G.miny, A.G.cell, A.G.gy); }
  bool blocked = false;
  for (int yy = max(ccy - 1, 0); yy <= min(ccy + 1, A.G.gy - 1) && !blocked; ++yy)
    for (int xx = max(ccx - 1, 0); xx <= min(ccx + 1, A.G.gx - 1) && !blocked; ++xx) {
      const int cell = yy * A.G.gx + xx;
      const unsigned int e = A.cell_start[cell + 1];
      for (unsigned int t = A.cell_start[cell]; t < e; ++t) {
        const int h = A.items[t];
        if (h >= c) continue;
        const int sh = A.state[h];
        if (sh != ST_UNDECIDED && sh != kept_now) continue;
        if (reaches(A, h, c, cy, cx, bc)) { blocked = true; break; }
      }
    }
  if (!blocked) A.state[c] = kept_now;
}

struct DevVerts {
  const int2* v;
  __device__ int32_t x(int i) const { return v[i].x; }
  __device__ int32_t y(int i) const { return v[i].y; }
};

// overlap test of kept h against candidate c; returns 1 suppressed, 0 not, -1 pool overflow
template <int NV, int SC>
__device__ int pair_suppresses(const NmsArrays& A, int h, int c, sdclip::ClipSweep<NV, SC>& S) {
  DevVerts va{A.verts + (size_t)h * A.R}, vb{A.verts + (size_t)c * A.R};
  int status;
  const float inter = sdclip::clip_intersection_area(va, vb, A.R, S, &status);
  if (status == sdclip::CLIP_OVERFLOW) return -1;
  // overlap = area_inter / fmin(areas[i]+1e-10, areas[j]+1e-10)  (double), stored to float (:580)
  const double den = fmin((double)A.area[h] + 1.e-10, (double)A.area[c] + 1.e-10);
  const float overlap = (float)((double)inter / den);
  return overlap > A.threshold ? 1 : 0;
}

template <int NV, int SC>
__device__ void suppress_candidate(const NmsArrays& A, int c, int round, sdclip::ClipSweep<NV, SC>& S,
                                   int* __restrict__ slow_list, unsigned int* __restrict__ counters) {
  const float cy = A.points[2 * c], cx = A.points[2 * c + 1];
  const int4 bc = A.bbox[c];
  const int kept_now = ST_KEPT_BASE + round;
  int ccx = 0, ccy = 0;
  if (!A.G.all_pairs) { ccx = cell_of(cx, A.G.minx, A.G.cell, A.G.gx); ccy = cell_of(cy, A.G.miny, A.G.cell, A.G.gy); }
  for (int yy = max(ccy - 1, 0); yy <= min(ccy + 1, A.G.gy - 1); ++yy)
    for (int xx = max(ccx - 1, 0); xx <= min(ccx + 1, A.G.gx - 1); ++xx) {
      const int cell = yy * A.G.gx + xx;
      const unsigned int e = A.cell_start[cell + 1];
      for (unsigned int t = A.cell_start[cell]; t < e; ++t) {
        const int h = A.items[t];
        if (h >= c) continue;
        if (A.state[h] != kept_now) continue;
        if (!reaches(A, h, c, cy, cx, bc)) continue;
        atomicAdd(&counters[2], 1u);                 // pair evaluations (stats)
        const int r = pair_suppresses<NV, SC>(A, h, c, S);
        if (r == 1) { A.state[c] = ST_SUPPRESSED; return; }
        if (r < 0) {
          if (slow_list) { unsigned int k = atomicAdd(&counters[1], 1u); slow_list[k] = c; }
          else atomicAdd(&counters[3], 1u);          // overflow in the slow path: hard error
          return;
        }
      }
    }
}

template <int NV>
__global__ void __launch_bounds__(128) k_suppress(NmsArrays A, int round, int* __restrict__ slow_list,
                                                  unsigned int* __restrict__ counters) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.n) return;
  if (A.state[c] != ST_UNDECIDED) return;
  sdclip::ClipSweep<NV, 1> S;
  suppress_candidate<NV, 1>(A, c, round, S, slow_list, counters);
}

template <int NV>
__global__ void __launch_bounds__(64) k_suppress_slow(NmsArrays A, int round, const int* __restrict__ slow_list,
                                                     unsigned int* __restrict__ counters) {
  const unsigned int n_slow = counters[1];
  if (blockIdx.x * blockDim.x >= n_slow) return;
  sdclip::ClipSweep<NV, 4> S;
  for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_slow; t += gridDim.x * blockDim.x)
    suppress_candidate<NV, 4>(A, slow_list[t], round, S, nullptr, counters);
}

__global__ void k_reset_counters(unsigned int* counters) {
  if (threadIdx.x < 2) counters[threadIdx.x] = 0;     // undecided, slow-list length
}

template <int NV>
int run_rounds(NmsArrays A, int* d_slow, unsigned int* d_counters, cudaStream_t st, int verbose,
               unsigned int* h_pin /* pinned [4*BATCH] */) {
  const int n = A.n;
  constexpr int BATCH = 4;     // rounds launched per host synchronisation
  int round = 0;
  for (;;) {
    for (int b = 0; b < BATCH; ++b, ++round) {
      SDB_LAUNCH(k_reset_counters, 1, 32, 0, st, d_counters);
      SDB_LAUNCH(k_frontier, cdiv(n, 256), 256, 0, st, A, round, d_counters);
      SDB_LAUNCH((k_suppress<NV>), cdiv(n, 128), 128, 0, st, A, round, d_slow, d_counters);
      // slow path (pool overflow in the fast path): usually zero entries; grid-stride over the list
      SDB_LAUNCH((k_suppress_slow<NV>), 8, 64, 0, st, A, round, d_slow, d_counters);
      SDB_CUDA(cudaMemcpyAsync(h_pin + 4 * b, d_counters, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    }
    SDB_CUDA(cudaStreamSynchronize(st));
    bool done = false;
    for (int b = 0; b < BATCH; ++b) {
      const unsigned int* c = h_pin + 4 * b;
      if (c[3] != 0) { sdb::set_error("nms2d: polygon clipping pools overflowed in the slow path"); return 1; }
      if (c[0] == 0) { done = true; break; }
    }
    if (verbose) printf("NMS2D(b200): rounds=%d undecided(last batch)=%u pair tests so far=%u\n", round, h_pin[4 * (BATCH - 1)], h_pin[4 * (BATCH - 1) + 2]);
    if (done) break;
    if (round > 4 * n + 8) { sdb::set_error("nms2d: no progress"); return 1; }
  }
  return 0;
}


}  // namespace
}  // namespace sdnms
