// tcgen05 self-test: one 128-row UMMA tile with operands staged in the blocked no-swizzle layout used by the
// fused kernel.  Pins the shared-memory / instruction descriptor encodings on hardware for the three operand
// views the fused step needs: forward (A,B K-major), data-gradient (B MN-major), weight-gradient (A,B MN-major).
#include "cnf_common.cuh"
#include "cnf_umma.cuh"

namespace cnf {
using namespace umma;

__global__ void __launch_bounds__(128)
umma_selftest_kernel(int mode, int N, int Kd, const float* __restrict__ A, const float* __restrict__ B,
                     float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * Kd * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  // stage A
  if (mode == 2) {      // A given [Kd, 128]
    for (int i = tid; i < Kd * 128; i += 128) {
      const int r = i / 128, c = i % 128;
      *reinterpret_cast<__nv_bfloat16*>(sA + blk_off(r, c, 128)) = __float2bfloat16(A[i]);
    }
  } else {              // A given [128, Kd]
    for (int i = tid; i < 128 * Kd; i += 128) {
      const int r = i / Kd, c = i % Kd;
      *reinterpret_cast<__nv_bfloat16*>(sA + blk_off(r, c, Kd)) = __float2bfloat16(A[i]);
    }
  }
  if (mode == 0) {      // B given [N, Kd]
    for (int i = tid; i < N * Kd; i += 128) {
      const int r = i / Kd, c = i % Kd;
      *reinterpret_cast<__nv_bfloat16*>(sB + blk_off(r, c, Kd)) = __float2bfloat16(B[i]);
    }
  } else {              // B given [Kd, N]
    for (int i = tid; i < Kd * N; i += 128) {
      const int r = i / N, c = i % N;
      *reinterpret_cast<__nv_bfloat16*>(sB + blk_off(r, c, N)) = __float2bfloat16(B[i]);
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, N, mode == 2, mode >= 1);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int ks = 0; ks < Kd / 16; ++ks) {
      uint64_t ad, bd;
      if (mode == 2) ad = smem_desc(a0 + ks * 2 * (16 * 128), /*LBO*/ 16 * 128, /*SBO*/ 128);
      else           ad = smem_desc(a0 + ks * 2 * 128, /*LBO*/ 128, /*SBO*/ 16 * Kd);
      if (mode == 0) bd = smem_desc(b0 + ks * 2 * 128, 128, 16 * Kd);
      else           bd = smem_desc(b0 + ks * 2 * (16 * N), 16 * N, 128);
      mma_ss(tmem, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t tl = tmem_lane_base(tmem);
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tl + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Probe of the M=64 accumulator layout: sentinel-fill 128 lanes x N columns, run ONE K=16.. chain of M=64 MMAs with the
// D address carrying `lane_off` in its lane field, dump all 128 lanes.  A [64,Kd], B [N,Kd] both K-major.
__global__ void __launch_bounds__(128)
umma_m64_probe_kernel(int N, int Kd, int lane_off, const float* __restrict__ A, const float* __restrict__ B,
                      float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * Kd * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * Kd; i += 128) {
    const int r = i / Kd, c = i % Kd;
    *reinterpret_cast<__nv_bfloat16*>(sA + blk_off(r, c, Kd)) = __float2bfloat16(r < 64 ? A[i] : 0.f);
  }
  for (int i = tid; i < N * Kd; i += 128) {
    const int r = i / Kd, c = i % Kd;
    *reinterpret_cast<__nv_bfloat16*>(sB + blk_off(r, c, Kd)) = __float2bfloat16(B[i]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tl = tmem_lane_base(tmem);
  uint32_t sent[16];
  for (int j = 0; j < 16; ++j) sent[j] = __float_as_uint(-777.0f);
  for (int c0 = 0; c0 < N; c0 += 16) tmem_st16(tl + c0, sent);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(64, N, 0, 0);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int ks = 0; ks < Kd / 16; ++ks)
      mma_ss(tmem + ((uint32_t)lane_off << 16), smem_desc(a0 + ks * 256, 128, 16 * Kd), smem_desc(b0 + ks * 256, 128, 16 * Kd), idesc, ks > 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tl + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Modes 3..5 (operand views of the two-role fused kernel):
//   3: forward with the A operand in TENSOR MEMORY (each thread packs its row's bf16 pairs, one 32-bit column per two K
//      elements, tcgen05.st) and B [N,Kd] K-major in shared memory; D [128,N].
//   4/5: weight-gradient view with M = 64: A given [Kd, 64] and B given [Kd, N], both MN-major in shared memory; the D
//      address carries lane offset 0 (mode 4) / 16 (mode 5).  ALL 128 lanes are dumped (sentinel -777 where untouched):
//      accumulator row r must land in lane 32*(r/16) + r%16 (+16 in mode 5).
__global__ void __launch_bounds__(128)
umma_selftest2_kernel(int mode, int N, int Kd, const float* __restrict__ A, const float* __restrict__ B,
                      float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128 * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 128 * 2 + 256 * 128 * 2) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (mode >= 4) {
    for (int i = tid; i < Kd * 64; i += 128) {
      const int r = i / 64, c = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(sA + blk_off(r, c, 64)) = __float2bfloat16(A[i]);
    }
    for (int i = tid; i < Kd * N; i += 128) {
      const int r = i / N, c = i % N;
      *reinterpret_cast<__nv_bfloat16*>(sB + blk_off(r, c, N)) = __float2bfloat16(B[i]);
    }
  } else {
    for (int i = tid; i < N * Kd; i += 128) {
      const int r = i / Kd, c = i % Kd;
      *reinterpret_cast<__nv_bfloat16*>(sB + blk_off(r, c, Kd)) = __float2bfloat16(B[i]);
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tl = tmem_lane_base(tmem);
  if (mode == 3) {
    for (int c0 = 0; c0 < Kd; c0 += 16) {     // 16 K elements -> 8 columns
      uint32_t w[8];
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(A[tid * Kd + c0 + 2 * j], A[tid * Kd + c0 + 2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st8(tl + 256 + c0 / 2, w);
    }
  } else {
    uint32_t sent[16];
    for (int j = 0; j < 16; ++j) sent[j] = __float_as_uint(-777.0f);
    for (int c0 = 0; c0 < N; c0 += 16) tmem_st16(tl + c0, sent);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0 && elect_one()) {
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (mode == 3) {
      const uint32_t idesc = idesc_bf16(128, N, 0, 0);
      for (int ks = 0; ks < Kd / 16; ++ks)
        mma_ts(tmem, tmem + 256 + ks * 8, smem_desc(b0 + ks * 256, 128, 16 * Kd), idesc, ks > 0);
    } else {
      const uint32_t idesc = idesc_bf16(64, N, 1, 1);
      const uint32_t d = tmem + (mode == 5 ? (16u << 16) : 0u);
      for (int ks = 0; ks < Kd / 16; ++ks)
        mma_ss(d, smem_desc(a0 + ks * 2 * (16 * 64), 16 * 64, 128), smem_desc(b0 + ks * 2 * (16 * N), 16 * N, 128), idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tl + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Issue-rate probe: ONE thread issues `reps` back-to-back tcgen05.mma of shape M x N x 16 and the CTA measures the SM
// cycles until the commit arrives.  mode 0: A,B from shared memory, K-major (forward view); 1: A from TMEM, B K-major;
// 2: A,B from shared memory MN-major (weight-gradient view); 3: B MN-major (data-gradient view).
// Operand contents are whatever the (zeroed) shared memory holds -- only the timing is of interest.
__global__ void __launch_bounds__(128)
umma_rate_probe_kernel(int mode, int M, int N, int reps, int ndst, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (96 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0, t2 = 0;
  const bool issuer = warp == 0 && elect_one();
  if (issuer) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
    const uint32_t idesc = idesc_bf16(M, N, mode == 2, mode >= 2);
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      if (mode == 2) ad[ks] = smem_desc(a0 + ks * 2 * (16 * 128), 16 * 128, 128);
      else           ad[ks] = smem_desc(a0 + ks * 256, 128, 16 * 128);
      if (mode >= 2) bd[ks] = smem_desc(b0 + ks * 2 * (16 * N), 16 * N, 128);
      else           bd[ks] = smem_desc(b0 + ks * 256, 128, 16 * 128);
    }
    const uint32_t d0 = tmem, d1 = tmem + (ndst > 1 ? N : 0);
    t0 = clock64();
    for (int i = 0; i < reps; i += 8) {
      const uint32_t acc = i > 0;
      if (mode == 1) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ts((ks & 1) ? d1 : d0, tmem + 256 + ks * 8, bd[ks], idesc, acc | (ks > 1));
      } else {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ss((ks & 1) ? d1 : d0, ad[ks], bd[ks], idesc, acc | (ks > 1));
      }
    }
    t1 = clock64();
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  t2 = clock64();
  tc_fence_after();
  if (issuer) { out[blockIdx.x * 2 + 0] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace cnf

using namespace cnf;

extern "C" int cnf_umma_m64_probe(int N, int Kdim, int lane_off, const float* A, const float* B, float* D, cnf_stream_t stream) {
  CNF_REQUIRE(N >= 8 && N <= 256 && N % 8 == 0 && Kdim >= 16 && Kdim <= 256 && Kdim % 16 == 0 && A && B && D, "cnf_umma_m64_probe: bad arguments");
  const size_t smem = (size_t)128 * Kdim * 2 + (size_t)N * Kdim * 2;
  CNF_CUDA(cudaFuncSetAttribute(umma_m64_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_m64_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(N, Kdim, lane_off, A, B, D);
  return check_launch("cnf_umma_m64_probe");
}

extern "C" int cnf_umma_selftest(int mode, int N, int Kdim, const float* A, const float* B, float* D, cnf_stream_t stream) {
  CNF_REQUIRE(mode >= 0 && mode <= 5 && N >= 16 && N <= 256 && N % 16 == 0 && Kdim >= 16 && Kdim <= 256 && Kdim % 16 == 0,
              "cnf_umma_selftest: need mode in 0..5, N%%16==0 (16..256), K%%16==0 (16..256)");
  CNF_REQUIRE(A && B && D, "cnf_umma_selftest: null pointer");
  if (mode >= 3) {
    CNF_REQUIRE(Kdim <= 128, "cnf_umma_selftest: modes 3..5 need K <= 128");
    const int smem2 = 128 * 128 * 2 + 256 * 128 * 2;
    CNF_CUDA(cudaFuncSetAttribute(umma_selftest2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    umma_selftest2_kernel<<<1, 128, smem2, as_stream(stream)>>>(mode, N, Kdim, A, B, D);
    return check_launch("cnf_umma_selftest");
  }
  const size_t smem = (size_t)128 * Kdim * 2 + (size_t)N * Kdim * 2;
  CNF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(mode, N, Kdim, A, B, D);
  return check_launch("cnf_umma_selftest");
}

extern "C" int cnf_umma_rate_probe(int mode, int M, int N, int reps, int ndst, int n_cta, long long* out_cycles, cnf_stream_t stream) {
  CNF_REQUIRE(mode >= 0 && mode <= 3 && (M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && reps > 0 && ndst >= 1 &&
                  ndst * N <= 256 && n_cta >= 1 && out_cycles, "cnf_umma_rate_probe: bad arguments");
  const int smem = 96 * 1024;
  CNF_CUDA(cudaFuncSetAttribute(umma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_probe_kernel<<<n_cta, 128, smem, as_stream(stream)>>>(mode, M, N, reps, ndst, out_cycles);
  return check_launch("cnf_umma_rate_probe");
}
