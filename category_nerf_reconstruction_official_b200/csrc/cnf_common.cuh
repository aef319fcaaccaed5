// Shared helpers for libcnf_sm100.so (sm_100a only; no CPU fallback anywhere in this library).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/cnf.h"
#ifdef CNF_DEV
#include "../../include/cnf_dev.h"
#endif

namespace cnf {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return CNF_OK;
}

#define CNF_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      cnf::set_error(__VA_ARGS__);        \
      return CNF_ERR_ARG;                 \
    }                                     \
  } while (0)

#define CNF_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      cnf::set_error("%s failed: %s", #call, cudaGetErrorString(e__));     \
      return (int)e__;                                                     \
    }                                                                      \
  } while (0)

// "Has this call site configured the CURRENT device yet?"  Function attributes (the > 48 KB dynamic shared-memory opt-in) are per
// device, so a process that drives several GPUs must set them once on each; a plain `static bool` would skip every device but
// the first.  One bit per device ordinal, relaxed atomics (setting an attribute twice is harmless).
struct DeviceOnce {
  unsigned long long bits[4] = {0, 0, 0, 0};
  bool pending() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 256) return true;
    return ((__atomic_load_n(&bits[dev >> 6], __ATOMIC_RELAXED) >> (dev & 63)) & 1ull) == 0;
  }
  void done() {
    int d = 0;
    if (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 256) __atomic_fetch_or(&bits[d >> 6], 1ull << (d & 63), __ATOMIC_RELAXED);
  }
};

// Shared-memory carve-out (per cent of the SM's 228 KB) requested by the small kernels that run CONCURRENTLY inside one training step
// (small-batch sampler on one stream, prepare on another): two kernels whose carve-outs differ are serialised by the hardware
// (measured: the step lost its 9 us of overlap when the sampler's carve-out changed).
#define CNF_SMALL_STEP_CARVEOUT 28

inline cudaStream_t as_stream(cnf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// cnf_composite.cu: per-ray compositing of already-activated points pt[n_rays*S] = {occ, r, g, b} -> out_rays [n_rays,6]
// = {D, var, C0, C1, C2, O}, optional termination weights out_w [n_rays,S] (used by cnf_render_rays in cnf_fused.cu)
int render_composite_launch(int n_rays, int S, const float* pt, const float* z, float* out_rays, float* out_w, cudaStream_t st);

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace cnf
