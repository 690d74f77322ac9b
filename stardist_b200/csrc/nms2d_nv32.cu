// nms2d_nv32.cu -- instantiates the 2D NMS rounds for polygons with up to 32 rays.
#include "nms2d_rounds.cuh"
namespace sdnms {
int run_rounds_nv32(NmsArrays A, int* d_slow, unsigned int* d_counters, cudaStream_t st, int verbose, unsigned int* h_pin) {
  return run_rounds<32>(A, d_slow, d_counters, st, verbose, h_pin);
}
}  // namespace sdnms
