// Arithmetic of the depth-guided sampler shared by the standalone sampler kernels (cnf_sampler.cu) and the fused training step's
// in-kernel sampling (cnf_fused_b.inc): every fp32 operation separately rounded, in the reference's order
// (src/scene_cateogries.py:62-77, 484-539), so that both produce the reference's depths bit for bit.
#pragma once
#include <stdint.h>

namespace cnf {

// 0 invalid depth, 1 valid & this object, 2 valid & other / unknown (scene_cateogries.py:484, 507, 533)
__device__ __forceinline__ int ray_case(float d, uint8_t label, float min_bound, int this_obj) {
  if (d <= min_bound) return 0;
  return (label == this_obj) ? 1 : 2;
}

// stratified z for bin k: (range*lin[k] + min) + u*(range/n)   -- scene_cateogries.py:62-77
__device__ __forceinline__ float strat(float mn, float mxv, float lin_k, float u, float nf) {
  const float range = __fsub_rn(mxv, mn);
  const float lower = __fadd_rn(__fmul_rn(range, lin_k), mn);
  const float len = __fdiv_rn(range, nf);
  return __fadd_rn(lower, __fmul_rn(u, len));
}

}  // namespace cnf
