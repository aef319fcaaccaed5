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

// Several parameter groups in ONE launch with the step count on the DEVICE (state[0] = optimizer steps done, state[1] = block
// ticket): the bias corrections 1 - beta^t are computed in the kernel from state[0] + 1, so a launch captured into a CUDA graph
// advances t on every replay (a host-side t would be frozen into the captured arguments).  The last block to finish increments
// state[0]; every block has read it by then.
struct AdamGroups {
  cnf_adamw_group g[CNF_ADAMW_MAX_GROUPS];
  unsigned first_block[CNF_ADAMW_MAX_GROUPS + 1];
  int n;
};
__global__ void __launch_bounds__(256)
adamw_groups_kernel(const AdamGroups a, float b1, float b2, float eps, int32_t* __restrict__ state) {
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double t = (double)(*reinterpret_cast<volatile int32_t*>(state) + 1);
    s_bc[0] = (float)(1.0 - pow((double)b1, t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  int gi = 0;
  while (gi + 1 < a.n && blockIdx.x >= a.first_block[gi + 1]) ++gi;
  const cnf_adamw_group& G = a.g[gi];
  const size_t i = (size_t)(blockIdx.x - a.first_block[gi]) * 256 + threadIdx.x;
  if (i < G.n) {
    const float gr = G.grad[i];
    float p = G.param[i] * (1.0f - G.lr * G.weight_decay);
    const float m = b1 * G.exp_avg[i] + (1.0f - b1) * gr;
    const float v = b2 * G.exp_avg_sq[i] + (1.0f - b2) * gr * gr;
    G.exp_avg[i] = m; G.exp_avg_sq[i] = v;
    p -= (G.lr / s_bc[0]) * (m / (sqrtf(v) / s_bc[1] + eps));
    G.param[i] = p;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 1, 1) == (int)gridDim.x - 1) { state[1] = 0; atomicAdd(state, 1); }
  }
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

extern "C" int cnf_adamw_groups(int n_groups, const cnf_adamw_group* groups, float beta1, float beta2, float eps, int32_t* state,
                                cnf_stream_t stream) {
  CNF_REQUIRE(n_groups >= 1 && n_groups <= CNF_ADAMW_MAX_GROUPS && groups && state, "cnf_adamw_groups: bad arguments (1..%d groups)",
              CNF_ADAMW_MAX_GROUPS);
  AdamGroups a;
  a.n = n_groups;
  unsigned blocks = 0;
  for (int i = 0; i < n_groups; ++i) {
    CNF_REQUIRE(groups[i].n == 0 || (groups[i].param && groups[i].grad && groups[i].exp_avg && groups[i].exp_avg_sq),
                "cnf_adamw_groups: null pointer in group %d", i);
    a.g[i] = groups[i];
    a.first_block[i] = blocks;
    blocks += (unsigned)((groups[i].n + 255) / 256);
  }
  a.first_block[n_groups] = blocks;
  if (blocks == 0) return CNF_OK;
  adamw_groups_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a, beta1, beta2, eps, state);
  return check_launch("cnf_adamw_groups");
}
