// nms2d_nv128.cu -- instantiates the 2D NMS rounds for polygons with up to 128 rays.
#include "nms2d_rounds.cuh"
namespace sdnms {
int run_rounds_nv128(NmsArrays A, int* d_slow, unsigned int* d_counters, cudaStream_t st, int verbose, unsigned int* h_pin) {
  return run_rounds<128>(A, d_slow, d_counters, st, verbose, h_pin);
}
}  // namespace sdnms
