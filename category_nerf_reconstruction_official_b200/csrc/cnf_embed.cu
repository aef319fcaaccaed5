// Standalone positional/Fourier embedding: UniDirsEmbed.forward (reference src/embedding.py:82-92) + backward.
// emb[p] = cat(t, sin(pi * 2^f * (t . B_d)))  with t = x/scale, f-major / direction-minor (3 + f*21 + d).
// HBM-bound (12 B in, 516 B out per point): the tile is built in shared memory and written back as
// one contiguous, fully coalesced run of 129*PTS floats; one libm sincos per direction + double-angle recurrence.
#include "cnf_common.cuh"
#include "cnf_umma.cuh"      // mbarrier + 1-D bulk-copy (TMA) helpers

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
// Persistent CTAs, two stages: a tile's 64 x 129 upstream gradients (33 KB, contiguous) and its 64 points arrive by 1-D bulk copies
// (TMA) on an mbarrier while the previous tile is reduced -- the first version's load loop (one LDG + STS per element, then a
// barrier) held 60 % of the stall samples and kept the kernel at 63 % of the HBM roofline.  The 63 partial sums of g_B stay in
// registers across the CTA's tiles (one atomic per output and CTA).  cos(pi 2^f proj) comes from ONE hardware sin/cos after an exact
// argument reduction + the double-angle recurrence (as in the fused kernels; the libm sincosf was 21 % of the instructions).
constexpr int EB_STAGE_FLOATS = EP * CNF_EMB_DIM + EP * 3 + 4;        // g tile | x tile | pad (stage stays 16-byte aligned)
__global__ void __launch_bounds__(ET)
embed_bwd_kernel(int P, const float* __restrict__ x, const float* __restrict__ B, const float* __restrict__ scale,
                 const float* __restrict__ g_emb, float* __restrict__ g_B, float* __restrict__ g_x) {
  using namespace umma;
  extern __shared__ __align__(128) float sm_e[];
  __shared__ __align__(8) uint64_t full[2];
  __shared__ float s_B[CNF_N_DIRS * 3];
  __shared__ float s_gp[EP][CNF_N_DIRS + 1];
  const int k = blockIdx.y, t = threadIdx.x;
  const int ntiles = (P + EP - 1) / EP;
  const size_t kbase = (size_t)k * P;
  const float sc = scale[k];
  const bool aligned = ((reinterpret_cast<uintptr_t>(g_emb + kbase * CNF_EMB_DIM) | reinterpret_cast<uintptr_t>(x + kbase * 3)) & 15) == 0;
  auto bulk_ok = [&](int tile) { return aligned && tile < ntiles && (tile + 1) * EP <= P; };
  if (t == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  if (t < CNF_N_DIRS * 3) s_B[t] = B[(size_t)k * CNF_N_DIRS * 3 + t];
  __syncthreads();
  auto issue = [&](int tile, int st) {
    float* base = sm_e + st * EB_STAGE_FLOATS;
    const size_t p0 = kbase + (size_t)tile * EP;
    mbar_expect_tx(&full[st], (uint32_t)(EP * CNF_EMB_DIM * 4 + EP * 3 * 4));
    bulk_g2s(base, g_emb + p0 * CNF_EMB_DIM, EP * CNF_EMB_DIM * 4, &full[st]);
    bulk_g2s(base + EP * CNF_EMB_DIM, x + p0 * 3, EP * 3 * 4, &full[st]);
  };
  const float pi_f = 3.14159274101257324f;
  float accB = 0.f;                                  // threads 0..251: this CTA's partial sum of g_B[d][a] over a quarter of every tile
  uint32_t ph0 = 0, ph1 = 0;
  int tile = blockIdx.x;
  if (t == 0 && bulk_ok(tile)) issue(tile, 0);
  for (int j = 0; tile < ntiles; ++j, tile += gridDim.x) {
    const int st = j & 1;
    float* s_g = sm_e + st * EB_STAGE_FLOATS;
    float* s_x = s_g + EP * CNF_EMB_DIM;
    const int p0 = tile * EP, np = min(EP, P - p0);
    const int nxt = tile + gridDim.x;
    if (t == 0 && bulk_ok(nxt)) issue(nxt, st ^ 1);  // stage st^1 was released by the barrier that ended iteration j-1
    if (bulk_ok(tile)) {
      mbar_wait(&full[st], st ? ph1 : ph0);
      if (st) ph1 ^= 1; else ph0 ^= 1;
      if (t < EP * 3) s_x[t] = __fdiv_rn(s_x[t], sc);                                  // embedding.py:83, once per coordinate
    } else {
      const float* gin = g_emb + (kbase + p0) * CNF_EMB_DIM;
      for (int i = t; i < np * CNF_EMB_DIM; i += ET) s_g[i] = gin[i];
      for (int i = t; i < EP * 3; i += ET) s_x[i] = i < np * 3 ? __fdiv_rn(x[(kbase + p0) * 3 + i], sc) : 0.f;   // rows >= np: finite
    }
    __syncthreads();
    for (int i = t; i < EP * CNF_N_DIRS; i += ET) {
      const int p = i / CNF_N_DIRS, d = i - p * CNF_N_DIRS;
      float gp = 0.f;
      if (p < np) {
        const float t0 = s_x[p * 3 + 0], t1 = s_x[p * 3 + 1], t2 = s_x[p * 3 + 2];
        float proj = __fmul_rn(t0, s_B[d * 3 + 0]);
        proj = __fmaf_rn(t1, s_B[d * 3 + 1], proj);
        proj = __fmaf_rn(t2, s_B[d * 3 + 2], proj);
        // sin / cos(pi 2^f proj): exact reduction of 2^f proj to [-1, 1] (period 2), then the MUFU approximations (|error| <= 4e-7)
        // at octaves 0 and 3, the double-angle recurrence for the two octaves after each (error x4: <= 2e-6 on any cosine)
        const float* gcol = s_g + p * CNF_EMB_DIM + 3 + d;
        float band = pi_f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float pr = h ? 8.0f * proj : proj;
          float sn, cs;
          __sincosf(pi_f * (pr - 2.0f * rintf(0.5f * pr)), &sn, &cs);
#pragma unroll
          for (int f = 0; f < 3; ++f) {
            gp = fmaf(gcol[(3 * h + f) * CNF_N_DIRS] * band, cs, gp);
            const float tw = sn + sn;                 // sin 2a = 2 sin a cos a ; cos 2a = 1 - 2 sin^2 a
            const float c2 = fmaf(-tw, sn, 1.0f);
            sn = tw * cs;
            cs = c2;
            band *= 2.0f;
          }
        }
      }
      s_gp[p][d] = gp;
    }
    __syncthreads();
    if (t < 4 * CNF_N_DIRS * 3) {                    // thread (quarter, output): 16 of the tile's points each
      const int q = t / (CNF_N_DIRS * 3), o = t - q * (CNF_N_DIRS * 3), d = o / 3, a = o - 3 * d;
      float acc = 0.f;
#pragma unroll 4
      for (int p = q * (EP / 4); p < (q + 1) * (EP / 4); ++p) acc = fmaf(s_gp[p][d], s_x[p * 3 + a], acc);     // rows >= np hold gp = 0
      accB += acc;
    }
    if (g_x) {
      for (int i = t; i < np * 3; i += ET) {
        const int p = i / 3, a = i - 3 * p;
        float acc = s_g[p * CNF_EMB_DIM + a];
        for (int d = 0; d < CNF_N_DIRS; ++d) acc += s_gp[p][d] * s_B[d * 3 + a];
        g_x[(kbase + p0) * 3 + i] = acc / sc;
      }
    }
    __syncthreads();                                 // stage st and s_gp are free again
  }
  if (t < 4 * CNF_N_DIRS * 3) atomicAdd(&g_B[(size_t)k * CNF_N_DIRS * 3 + t % (CNF_N_DIRS * 3)], accB);
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
  const int smem = 2 * EB_STAGE_FLOATS * (int)sizeof(float);
  static DeviceOnce once;
  if (once.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CNF_CUDA(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    once.done();
  }
  int nb = (P + EP - 1) / EP;
  if (nb > 148 * 3) nb = 148 * 3;                  // persistent: three 72 KB CTAs per SM, grid-stride over the tiles
  embed_bwd_kernel<<<dim3(nb, K), ET, smem, as_stream(stream)>>>(P, x, B, scale, g_emb, g_B, g_x);
  return check_launch("cnf_embed_bwd");
}
