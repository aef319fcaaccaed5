// Standalone positional/Fourier embedding: UniDirsEmbed.forward (reference src/embedding.py:82-92) + backward.
// emb[p] = cat(t, sin(pi * 2^f * (t . B_d)))  with t = x/scale, f-major / direction-minor (3 + f*21 + d).
// HBM-bound (12 B in, 516 B out per point): the tile is built in shared memory and written back as
// one contiguous, fully coalesced run of 129*PTS floats; one libm sincos per direction + double-angle recurrence.
#include "cnf_common.cuh"

namespace cnf {

constexpr int EP = 64;          // points per block
constexpr int ET = 256;         // threads per block

__global__ void __launch_bounds__(ET)
embed_fwd_kernel(int P, const float* __restrict__ x, const float* __restrict__ B, const float* __restrict__ scale,
                 float* __restrict__ emb) {
  const int k = blockIdx.y;
  const int p0 = blockIdx.x * EP;
  const int np = min(EP, P - p0);
  __shared__ float s_t[EP][3];
  __shared__ float s_B[CNF_N_DIRS * 3];
  __shared__ float s_e[EP * CNF_EMB_DIM];
  const float sc = scale[k];
  for (int i = threadIdx.x; i < np * 3; i += ET) {
    const float v = __fdiv_rn(x[((size_t)k * P + p0) * 3 + i], sc);   // embedding.py:83
    s_t[i / 3][i % 3] = v;
    s_e[(i / 3) * CNF_EMB_DIM + (i % 3)] = v;
  }
  if (threadIdx.x < CNF_N_DIRS * 3) s_B[threadIdx.x] = B[(size_t)k * CNF_N_DIRS * 3 + threadIdx.x];
  __syncthreads();
  const float pi_f = 3.14159274101257324f;   // float32(np.pi): the fp32 tensor is multiplied by a python scalar
  for (int i = threadIdx.x; i < np * CNF_N_DIRS; i += ET) {
    const int p = i / CNF_N_DIRS, d = i - p * CNF_N_DIRS;
    // t @ B^T (embedding.py:84): fp32 dot of length 3, accumulated left to right
    float proj = __fmul_rn(s_t[p][0], s_B[d * 3 + 0]);
    proj = __fmaf_rn(s_t[p][1], s_B[d * 3 + 1], proj);
    proj = __fmaf_rn(s_t[p][2], s_B[d * 3 + 2], proj);
    // embedding.py:85-88: sin(pi * 2^f * proj), f = 0..5.  ONE accurate sincos (f = 0) + the double-angle recurrence: 126
    // libm sinf per point made this kernel compute-bound (1.3 G sinf per 10 M points); the recurrence doubles the absolute
    // error per octave (<= 32 x 1.2e-7 ~ 4e-6, inside the 3e-5 parity bound) and leaves the kernel HBM-bound.
    float sn, cs;
    sincosf(__fmul_rn(proj, pi_f), &sn, &cs);
#pragma unroll
    for (int f = 0; f < CNF_N_FREQ; ++f) {
      s_e[p * CNF_EMB_DIM + 3 + f * CNF_N_DIRS + d] = sn;
      const float s2 = 2.0f * sn * cs;
      cs = fmaf(-2.0f * sn, sn, 1.0f);
      sn = s2;
    }
  }
  __syncthreads();
  float* out = emb + ((size_t)k * P + p0) * CNF_EMB_DIM;
  for (int i = threadIdx.x; i < np * CNF_EMB_DIM; i += ET) out[i] = s_e[i];
}

// backward: g_B[d,:] += sum_p (sum_f g[3+f*21+d] * pi*2^f*cos(pi*2^f*proj_d)) * t_p ; g_x = (g_t + sum_d gp_d B_d)/scale
__global__ void __launch_bounds__(ET)
embed_bwd_kernel(int P, const float* __restrict__ x, const float* __restrict__ B, const float* __restrict__ scale,
                 const float* __restrict__ g_emb, float* __restrict__ g_B, float* __restrict__ g_x) {
  const int k = blockIdx.y;
  const int p0 = blockIdx.x * EP;
  const int np = min(EP, P - p0);
  __shared__ float s_t[EP][3];
  __shared__ float s_B[CNF_N_DIRS * 3];
  __shared__ float s_g[EP * CNF_EMB_DIM];
  __shared__ float s_gp[EP][CNF_N_DIRS + 1];
  const float sc = scale[k];
  for (int i = threadIdx.x; i < np * 3; i += ET) s_t[i / 3][i % 3] = __fdiv_rn(x[((size_t)k * P + p0) * 3 + i], sc);
  if (threadIdx.x < CNF_N_DIRS * 3) s_B[threadIdx.x] = B[(size_t)k * CNF_N_DIRS * 3 + threadIdx.x];
  const float* gin = g_emb + ((size_t)k * P + p0) * CNF_EMB_DIM;
  for (int i = threadIdx.x; i < np * CNF_EMB_DIM; i += ET) s_g[i] = gin[i];
  __syncthreads();
  const float pi_f = 3.14159274101257324f;
  for (int i = threadIdx.x; i < EP * CNF_N_DIRS; i += ET) {
    const int p = i / CNF_N_DIRS, d = i - p * CNF_N_DIRS;
    float gp = 0.f;
    if (p < np) {
      float proj = __fmul_rn(s_t[p][0], s_B[d * 3 + 0]);
      proj = __fmaf_rn(s_t[p][1], s_B[d * 3 + 1], proj);
      proj = __fmaf_rn(s_t[p][2], s_B[d * 3 + 2], proj);
      float band = 1.0f, sn, cs;
      sincosf(__fmul_rn(proj, pi_f), &sn, &cs);                         // one accurate sincos + double-angle recurrence (see forward)
#pragma unroll
      for (int f = 0; f < CNF_N_FREQ; ++f) {
        gp += s_g[p * CNF_EMB_DIM + 3 + f * CNF_N_DIRS + d] * (pi_f * band) * cs;
        const float s2 = 2.0f * sn * cs;
        cs = fmaf(-2.0f * sn, sn, 1.0f);
        sn = s2;
        band *= 2.0f;
      }
    }
    s_gp[p][d] = gp;
  }
  __syncthreads();
  // g_B: 63 outputs, each a sum over the block's points; one thread per output, then one atomic per block
  if (threadIdx.x < CNF_N_DIRS * 3) {
    const int d = threadIdx.x / 3, a = threadIdx.x % 3;
    float acc = 0.f;
    for (int p = 0; p < np; ++p) acc += s_gp[p][d] * s_t[p][a];
    atomicAdd(&g_B[(size_t)k * CNF_N_DIRS * 3 + threadIdx.x], acc);
  }
  if (g_x) {
    for (int i = threadIdx.x; i < np * 3; i += ET) {
      const int p = i / 3, a = i % 3;
      float acc = s_g[p * CNF_EMB_DIM + a];
      for (int d = 0; d < CNF_N_DIRS; ++d) acc += s_gp[p][d] * s_B[d * 3 + a];
      g_x[((size_t)k * P + p0) * 3 + i] = acc / sc;
    }
  }
}

}  // namespace cnf

using namespace cnf;

extern "C" int cnf_embed_fwd(int K, int P, const float* x, const float* B, const float* scale, float* emb,
                             cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && P >= 0 && x && B && scale && emb, "cnf_embed_fwd: bad arguments");
  if (P == 0) return CNF_OK;
  embed_fwd_kernel<<<dim3((P + EP - 1) / EP, K), ET, 0, as_stream(stream)>>>(P, x, B, scale, emb);
  return check_launch("cnf_embed_fwd");
}

extern "C" int cnf_embed_bwd(int K, int P, const float* x, const float* B, const float* scale, const float* g_emb,
                             float* g_B, float* g_x, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && P >= 0 && x && B && scale && g_emb && g_B, "cnf_embed_bwd: bad arguments");
  if (P == 0) return CNF_OK;
  embed_bwd_kernel<<<dim3((P + EP - 1) / EP, K), ET, 0, as_stream(stream)>>>(P, x, B, scale, g_emb, g_B, g_x);
  return check_launch("cnf_embed_bwd");
}
