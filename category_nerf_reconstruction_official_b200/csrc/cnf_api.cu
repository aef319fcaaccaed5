// Library-wide entry points: version, error string, device check, fused AdamW.
#include "cnf_common.cuh"
#include <string.h>

namespace cnf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// torch.optim.AdamW semantics (decoupled weight decay; reference train.py:40,183): one thread per element.
__global__ void adamw_kernel(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, float lr, float b1, float b2, float eps, float wd, float bc1,
                             float bc2_sqrt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  float pi = p[i] * (1.0f - lr * wd);
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= (lr / bc1) * (mi / denom);
  p[i] = pi;
}
}  // namespace cnf

using namespace cnf;

extern "C" int cnf_version(void) { return 100; }
extern "C" const char* cnf_last_error_string(void) { return g_err; }

extern "C" int cnf_device_check(int device) {
  cudaDeviceProp prop;
  CNF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libcnf_sm100 only runs on sm_100 (B200) and has no fallback", device, prop.major,
              prop.minor);
    return CNF_ERR_ARCH;
  }
  return CNF_OK;
}

extern "C" int cnf_adamw_step(size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int step, cnf_stream_t stream) {
  CNF_REQUIRE(param && grad && exp_avg && exp_avg_sq && step >= 1, "cnf_adamw_step: bad arguments");
  if (n == 0) return CNF_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  adamw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1,
                                                                           beta2, eps, weight_decay, bc1, sqrtf(bc2));
  return check_launch("cnf_adamw_step");
}
