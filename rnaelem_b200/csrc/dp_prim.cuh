// dp_prim.cuh -- warp-level primitives shared by the log-space (dp_warp.cuh) and linear-space (dp_lin.cuh) passes.
// Under RELEM_HOST_EMU a "warp" is one host thread (WARP_N = 1) and every collective is the identity.
#ifndef RELEM_DP_PRIM_CUH
#define RELEM_DP_PRIM_CUH
#include "dp_common.cuh"

namespace relem {
namespace dp {

#ifdef RELEM_HOST_EMU
#define WARP_N 1
RDEV int lane_id() { return 0; }
RDEV int warp_id() { return 0; }
RDEV int n_warps() { return 1; }
RDEV unsigned w_ballot(bool p) { return p ? 1u : 0u; }
RDEV double w_shfl_down(double v, int) { return v; }
RDEV int w_shfl_down(int v, int) { return v; }
RDEV int w_shfl_up(int v, int) { return v; }
RDEV double w_shfl(double v, int) { return v; }
RDEV int w_shfl(int v, int) { return v; }
RDEV void w_sync() {}
RDEV void w_fence() {}
RDEV int w_ffs(unsigned b) { return __builtin_ffs((int)b); }
RDEV void sm_add(double* p, double v) { *p += v; }
RDEV int ctr_next(int* c) { return (*c)++; }
#else
#define WARP_N 32
RDEV int lane_id() { return (int)(threadIdx.x & 31); }
RDEV int warp_id() { return (int)(threadIdx.x >> 5); }
RDEV int n_warps() { return (int)(blockDim.x >> 5); }
RDEV unsigned w_ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
RDEV double w_shfl_down(double v, int o) { return __shfl_down_sync(0xFFFFFFFFu, v, o); }
RDEV int w_shfl_down(int v, int o) { return __shfl_down_sync(0xFFFFFFFFu, v, o); }
RDEV int w_shfl_up(int v, int o) { return __shfl_up_sync(0xFFFFFFFFu, v, o); }
RDEV double w_shfl(double v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
RDEV int w_shfl(int v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
RDEV void w_sync() { __syncwarp(); }
RDEV void w_fence() { __threadfence(); }
RDEV int w_ffs(unsigned b) { return __ffs((int)b); }
RDEV void sm_add(double* p, double v) { atomicAdd(p, v); }
RDEV int ctr_next(int* c) { return atomicAdd(c, 1); }
#endif

#ifdef RELEM_HOST_EMU
RDEV bool w_any(bool p) { return p; }
RDEV unsigned lanemask_lt() { return 0u; }
RDEV int w_popc(unsigned b) { return __builtin_popcount(b); }
RDEV double w_sum(double v) { return v; }
#else
RDEV bool w_any(bool p) { return __any_sync(0xFFFFFFFFu, p); }
RDEV unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
RDEV int w_popc(unsigned b) { return __popc(b); }
RDEV double w_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
#endif

}  // namespace dp
}  // namespace relem
#endif
