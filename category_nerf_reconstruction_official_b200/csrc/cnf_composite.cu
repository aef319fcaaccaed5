// Occupancy -> termination compositing and the depth / colour / opacity losses, forward and backward.
// Reference: src/render_rays.py:3-95 (occupancy_activation, occupancy_to_termination, render, render_loss,
// reduce_batch_loss) and src/loss.py:5-74 (step_batch_loss, step_batch_loss_reg); math in SURVEY appendix A.
//
// S <= 32 (training): tile kernels -- coalesced staging of 128 rays in shared memory, one thread per ray (see below).
// S > 32 (e.g. the 96-sample uncertainty renders): one ray occupies a power-of-two segment of a warp; the exclusive
// cumprod of the free-space factors and the reverse (suffix) sum of the backward are warp-shuffle scans over 32-sample
// chunks with a carry.  HBM-bound: 20 B/sample in, 24 B/ray out (fwd).
#include "cnf_common.cuh"
#include "cnf_umma.cuh"      // mbarrier + 1-D bulk-copy (TMA) helpers

namespace cnf {

constexpr int CT = 256;  // threads per block

__device__ __forceinline__ float seg_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive product scan inside a G-lane segment
__device__ __forceinline__ float seg_scan_mul(float v, int G, int sl) {
  for (int o = 1; o < G; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (sl >= o) v *= n;
  }
  return v;
}
// inclusive suffix-sum scan inside a G-lane segment
__device__ __forceinline__ float seg_scan_add_rev(float v, int G, int sl) {
  for (int o = 1; o < G; o <<= 1) {
    const float n = __shfl_down_sync(0xffffffffu, v, o);
    if (sl + o < G) v += n;
  }
  return v;
}
__device__ __forceinline__ float sigmoid_acc(float a) { return 1.0f / (1.0f + expf(-a)); }
// <= 3 ulp (ex2.approx of a pre-scaled argument + a correctly rounded reciprocal); saturates to exactly 0 / 1 like the accurate form
__device__ __forceinline__ float sigmoid_fast(float a) { return __frcp_rn(1.0f + __expf(-a)); }

struct RayOut { float D, var, C0, C1, C2, O; };

// Forward quantities of one ray held by a G-lane segment.  Returns totals in every lane of the segment.
__device__ __forceinline__ RayOut ray_forward(const float* __restrict__ alpha, const float* __restrict__ color,
                                              const float* __restrict__ z, int S, int G, int sl) {
  RayOut o{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float carry = 1.0f;
  for (int c0 = 0; c0 < S; c0 += G) {
    const int i = c0 + sl;
    const bool act = i < S;
    const float occ = act ? sigmoid_acc(alpha[i]) : 0.f;
    const float f = act ? __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f) : 1.0f;   // render_rays.py:28
    const float incl = seg_scan_mul(f, G, sl);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (sl == 0) excl = 1.0f;
    const float T = carry * excl;
    const float w = occ * T;                                                // render_rays.py:33
    carry *= __shfl_sync(0xffffffffu, incl, G - 1, G);
    if (act) {
      o.D += w * z[i]; o.O += w;
      o.C0 += w * color[i * 3 + 0]; o.C1 += w * color[i * 3 + 1]; o.C2 += w * color[i * 3 + 2];
    }
  }
  o.D = seg_sum(o.D, G); o.O = seg_sum(o.O, G);
  o.C0 = seg_sum(o.C0, G); o.C1 = seg_sum(o.C1, G); o.C2 = seg_sum(o.C2, G);
  // variance needs D: second pass (loss.py:45)
  carry = 1.0f;
  for (int c0 = 0; c0 < S; c0 += G) {
    const int i = c0 + sl;
    const bool act = i < S;
    const float occ = act ? sigmoid_acc(alpha[i]) : 0.f;
    const float f = act ? __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f) : 1.0f;
    const float incl = seg_scan_mul(f, G, sl);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (sl == 0) excl = 1.0f;
    const float w = occ * carry * excl;
    carry *= __shfl_sync(0xffffffffu, incl, G - 1, G);
    if (act) { const float dz = z[i] - o.D; o.var += w * dz * dz; }
  }
  o.var = seg_sum(o.var, G);
  return o;
}

__device__ __forceinline__ int seg_width(int S) {
  int G = 1;
  while (G < S && G < 32) G <<= 1;
  return G;
}

__global__ void mask_counts_kernel(int R, const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask_depth,
                                   int32_t* __restrict__ counts) {
  const int k = blockIdx.y;
  int cd = 0, co = 0, cs = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    const uint8_t lab = labels[(size_t)k * R + r];
    const bool mo = lab != 0, ms = lab != 2, md = mask_depth[(size_t)k * R + r] != 0;
    cd += (md && mo); co += mo; cs += ms;
  }
  cd = warp_sum_i(cd); co = warp_sum_i(co); cs = warp_sum_i(cs);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counts[k * 4 + 0], cd); atomicAdd(&counts[k * 4 + 1], co); atomicAdd(&counts[k * 4 + 2], cs);
  }
}

// flags: term t (0 depth,1 obj,2 sem) is empty in ANY category -> that loss term is 0 everywhere (render_rays.py:67-72)
__device__ __forceinline__ void empty_flags(int K, const int32_t* __restrict__ counts, bool& e0, bool& e1, bool& e2) {
  e0 = e1 = e2 = false;
  for (int k = 0; k < K; ++k) {
    e0 |= counts[k * 4 + 0] == 0; e1 |= counts[k * 4 + 1] == 0; e2 |= counts[k * 4 + 2] == 0;
  }
}

__global__ void __launch_bounds__(CT)
composite_fwd_kernel(int K, int R, int S, const float* __restrict__ alpha, const float* __restrict__ color,
                     const float* __restrict__ z, const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                     const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask_depth,
                     float* __restrict__ out_rays, float* __restrict__ partial) {
  const int k = blockIdx.y;
  const int G = seg_width(S), rpw = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = lane / G, sl = lane - seg * G;
  const int rays_per_block = (CT / 32) * rpw;
  float ld = 0.f, lc = 0.f, lo = 0.f;
  for (int rb = blockIdx.x * rays_per_block; rb < R; rb += gridDim.x * rays_per_block) {
    const int r = rb + warp * rpw + seg;
    const int rr = min(r, R - 1);                 // keep all lanes in the shuffles
    const size_t g = (size_t)k * R + rr;
    const RayOut o = ray_forward(alpha + g * S, color + g * S * 3, z + g * S, S, G, sl);
    if (r < R && sl == 0) {
      float* orow = out_rays + g * 6;
      orow[0] = o.D; orow[1] = o.var; orow[2] = o.C0; orow[3] = o.C1; orow[4] = o.C2; orow[5] = o.O;
      const uint8_t lab = labels[g];
      const float mo = (lab != 0) ? 1.f : 0.f, ms = (lab != 2) ? 1.f : 0.f;
      const float md = (mask_depth[g] != 0 && lab != 0) ? 1.f : 0.f;
      const float info = 1.0f / (sqrtf(o.var) + 1e-4f);                       // render_rays.py:73-80
      ld += fabsf(o.D - gt_depth[g]) * md * info;                             // loss.py:52-54
      lc += (fabsf(o.C0 - gt_color[g * 3]) + fabsf(o.C1 - gt_color[g * 3 + 1]) + fabsf(o.C2 - gt_color[g * 3 + 2])) * mo;
      lo += fabsf(o.O - mo) * ms;                                             // loss.py:64-66
    }
  }
  __shared__ float s_p[CT / 32][3];
  ld = warp_sum(ld); lc = warp_sum(lc); lo = warp_sum(lo);
  if (lane == 0) { s_p[warp][0] = ld; s_p[warp][1] = lc; s_p[warp][2] = lo; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int i = 0; i < CT / 32; ++i) t += s_p[i][threadIdx.x];
    partial[((size_t)k * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;    // deterministic 2-stage reduce
  }
}

// out_loss = {loss, e_depth, e_obj, e_sem, depth[K], color[K], opacity[K]}
// Loss outputs from the per-CTA partial sums, by ONE CTA of any size (a multiple of 32 threads, <= 1024): fixed summation order
// (thread-strided partial sums, a shuffle tree per warp, the warp sums in warp order) -> deterministic.  Called by the finalize
// kernel, and by the LAST CTA of the dense tile kernel (ticket counter), which saves the second launch.
__device__ __forceinline__ void finalize_block(int K, int nblk, const float* __restrict__ partial, const int32_t* __restrict__ counts,
                                               float cs, float os, float* __restrict__ out_loss) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  __shared__ float s_p[32][3];
  bool e0, e1, e2;
  empty_flags(K, counts, e0, e1, e2);
  float total = 0.f;
  for (int k = 0; k < K; ++k) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
      const float* p = partial + ((size_t)k * nblk + b) * 3;
      s0 += __ldcg(p); s1 += __ldcg(p + 1); s2 += __ldcg(p + 2);             // written by other CTAs: bypass L1
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (nwarp > 1) {
      __syncthreads();                        // previous category's s_p has been read
      if (lane == 0) { s_p[warp][0] = s0; s_p[warp][1] = s1; s_p[warp][2] = s2; }
      __syncthreads();
      s0 = s1 = s2 = 0.f;
      for (int w = 0; w < nwarp; ++w) { s0 += s_p[w][0]; s1 += s_p[w][1]; s2 += s_p[w][2]; }
    }
    const float d = e0 ? 0.f : s0 / ((float)counts[k * 4 + 0] + 1e-10f);      // render_rays.py:86
    const float c = e1 ? 0.f : s1 / ((float)counts[k * 4 + 1] + 1e-10f);
    const float o = e2 ? 0.f : s2 / ((float)counts[k * 4 + 2] + 1e-10f);
    if (threadIdx.x == 0) { out_loss[4 + k] = d; out_loss[4 + K + k] = c; out_loss[4 + 2 * K + k] = o; }
    total += d + c * cs + o * os;                                             // loss.py:69-70
  }
  if (threadIdx.x == 0) { out_loss[0] = total; out_loss[1] = e0; out_loss[2] = e1; out_loss[3] = e2; }
}

__global__ void composite_finalize_kernel(int K, int nblk, const float* __restrict__ partial,
                                          const int32_t* __restrict__ counts, float cs, float os,
                                          float* __restrict__ out_loss) {
  finalize_block(K, nblk, partial, counts, cs, os, out_loss);
}

__global__ void __launch_bounds__(CT)
composite_bwd_kernel(int K, int R, int S, const float* __restrict__ alpha, const float* __restrict__ color,
                     const float* __restrict__ z, const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                     const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask_depth,
                     const int32_t* __restrict__ counts, float cs, float os, const float* __restrict__ g_loss,
                     float* __restrict__ g_alpha, float* __restrict__ g_color) {
  const int k = blockIdx.y;
  const int G = seg_width(S), rpw = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = lane / G, sl = lane - seg * G;
  const int rays_per_block = (CT / 32) * rpw;
  bool e0, e1, e2;
  empty_flags(K, counts, e0, e1, e2);
  const float gl = g_loss ? g_loss[0] : 1.0f;
  const float iNd = e0 ? 0.f : gl / ((float)counts[k * 4 + 0] + 1e-10f);
  const float iNo = e1 ? 0.f : gl * cs / ((float)counts[k * 4 + 1] + 1e-10f);
  const float iNs = e2 ? 0.f : gl * os / ((float)counts[k * 4 + 2] + 1e-10f);
  const int nchunk = (S + G - 1) / G;
  for (int rb = blockIdx.x * rays_per_block; rb < R; rb += gridDim.x * rays_per_block) {
    const int r = rb + warp * rpw + seg;
    const int rr = min(r, R - 1);
    const size_t g = (size_t)k * R + rr;
    const float* al = alpha + g * S; const float* co = color + g * S * 3; const float* zz = z + g * S;
    const RayOut o = ray_forward(al, co, zz, S, G, sl);
    const uint8_t lab = labels[g];
    const float mo = (lab != 0) ? 1.f : 0.f, ms = (lab != 2) ? 1.f : 0.f;
    const float md = (mask_depth[g] != 0 && lab != 0) ? 1.f : 0.f;
    auto sgn = [](float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); };
    const float gD = sgn(o.D - gt_depth[g]) * md * (1.0f / (sqrtf(o.var) + 1e-4f)) * iNd;
    const float gC0 = sgn(o.C0 - gt_color[g * 3 + 0]) * mo * iNo;
    const float gC1 = sgn(o.C1 - gt_color[g * 3 + 1]) * mo * iNo;
    const float gC2 = sgn(o.C2 - gt_color[g * 3 + 2]) * mo * iNo;
    const float gO = sgn(o.O - mo) * ms * iNs;
    // total free-space product of whole chunks before each chunk, walking chunks in reverse with a suffix carry
    float suffix_carry = 0.f;
    for (int ch = nchunk - 1; ch >= 0; --ch) {
      // transmittance carry into this chunk = product of f over all earlier chunks
      float carry = 1.0f;
      for (int c2 = 0; c2 < ch; ++c2) {
        const int i2 = c2 * G + sl;
        const float f2 = (i2 < S) ? __fadd_rn(__fsub_rn(1.0f, sigmoid_acc(al[i2])), 1e-10f) : 1.0f;
        const float inc2 = seg_scan_mul(f2, G, sl);
        carry *= __shfl_sync(0xffffffffu, inc2, G - 1, G);
      }
      const int i = ch * G + sl;
      const bool act = i < S;
      const float occ = act ? sigmoid_acc(al[i]) : 0.f;
      const float f = act ? __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f) : 1.0f;
      const float incl = seg_scan_mul(f, G, sl);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (sl == 0) excl = 1.0f;
      const float T = carry * excl;
      const float w = occ * T;
      float c0 = 0.f, c1 = 0.f, c2v = 0.f, zi = 0.f;
      if (act) { c0 = co[i * 3]; c1 = co[i * 3 + 1]; c2v = co[i * 3 + 2]; zi = zz[i]; }
      const float Gi = act ? (gD * zi + gC0 * c0 + gC1 * c1 + gC2 * c2v + gO) : 0.f;
      const float Gw = Gi * w;
      const float incl_s = seg_scan_add_rev(Gw, G, sl);
      const float suffix = (incl_s - Gw) + suffix_carry;                     // sum_{i>j} G_i w_i
      suffix_carry += __shfl_sync(0xffffffffu, incl_s, 0, G);
      if (act && r < R) {
        const float docc = Gi * T - suffix / f;
        g_alpha[g * S + i] = docc * occ * (1.0f - occ);
        g_color[(g * S + i) * 3 + 0] = w * gC0;
        g_color[(g * S + i) * 3 + 1] = w * gC1;
        g_color[(g * S + i) * 3 + 2] = w * gC2;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Tile kernels for S <= 32 (the training configurations): a block of TR threads stages TR consecutive rays (alpha, z, colour:
// 20 B per sample) in shared memory with fully coalesced loads, then ONE THREAD PER RAY runs the compositing in registers
// (sequential exclusive cumprod / suffix sum, no shuffles); the backward writes its per-sample gradients back into the
// staged tile and the block copies them out coalesced.  Row strides are odd (bank-conflict free).  HBM-bound.
// ---------------------------------------------------------------------------------------------------------
constexpr int TR = 128;
__host__ __device__ inline int tile_smem_floats(int S, bool bwd) { return TR * ((S | 1) * (bwd ? 3 : 2) + (3 * S | 1)); }

// asynchronous global -> shared copies (cp.async, 4 B each): no register round trip, so ALL of a thread's loads are in flight
// at once (a plain load/store loop keeps only the unrolled few in flight and reaches a fraction of the HBM bandwidth)
__device__ __forceinline__ void tile_load(float* __restrict__ dst, int stride, const float* __restrict__ src, int n_rows, int cols) {
  const unsigned inv = 0xFFFFFFFFu / (unsigned)cols + 1u;            // exact e / cols for e < 2^16
  const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
  for (int e = threadIdx.x; e < n_rows * cols; e += TR) {
    const int r = (int)__umulhi((unsigned)e, inv), c = e - r * cols;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * (uint32_t)(r * stride + c)), "l"(src + e) : "memory");
  }
}
__device__ __forceinline__ void tile_load_wait() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tile_store(float* __restrict__ dst, const float* __restrict__ src, int stride, int n_rows, int cols) {
  const unsigned inv = 0xFFFFFFFFu / (unsigned)cols + 1u;
  for (int e = threadIdx.x; e < n_rows * cols; e += TR) {
    const int r = (int)__umulhi((unsigned)e, inv), c = e - r * cols;
    dst[e] = src[r * stride + c];
  }
}

// MODE 0: forward (rays + loss partials); 1: backward (gradients); 2: both in ONE pass (the backward only needs per-ray quantities
// and the input-only mask counts, so the loss value and its gradients can leave the same kernel: small-batch steps save a launch
// and a second read of alpha / colour / z).
template <int MODE>
__global__ void __launch_bounds__(TR)
composite_tile_kernel(int K, int R, int S, const float* __restrict__ alpha, const float* __restrict__ color,
                      const float* __restrict__ z, const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                      const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask_depth,
                      const int32_t* __restrict__ counts, float cs, float os, const float* __restrict__ g_loss,
                      float* __restrict__ out_rays, float* __restrict__ partial, float* __restrict__ g_alpha, float* __restrict__ g_color) {
  extern __shared__ float sm_t[];
  constexpr bool BWD = MODE != 0, FWD = MODE != 1;
  const int k = blockIdx.y, t = threadIdx.x;
  const int SP = S | 1, CP = (3 * S) | 1;
  float* sA = sm_t;                    // alpha -> (fwd) weights w_i / (bwd) occupancy -> d alpha
  float* sZ = sA + TR * SP;
  float* sC = sZ + TR * SP;            // colour -> (bwd) d colour
  float* sT = sC + TR * CP;            // (bwd) transmittance T_i
  float iNd = 0.f, iNo = 0.f, iNs = 0.f;
  if (BWD) {
    bool e0, e1, e2;
    empty_flags(K, counts, e0, e1, e2);
    const float gl = g_loss ? g_loss[0] : 1.0f;
    iNd = e0 ? 0.f : gl / ((float)counts[k * 4 + 0] + 1e-10f);
    iNo = e1 ? 0.f : gl * cs / ((float)counts[k * 4 + 1] + 1e-10f);
    iNs = e2 ? 0.f : gl * os / ((float)counts[k * 4 + 2] + 1e-10f);
  }
  float ld = 0.f, lc = 0.f, lo = 0.f;
  for (int r0 = blockIdx.x * TR; r0 < R; r0 += gridDim.x * TR) {
    const int n = min(TR, R - r0);
    const size_t g0 = (size_t)k * R + r0;
    tile_load(sA, SP, alpha + g0 * S, n, S);
    tile_load(sZ, SP, z + g0 * S, n, S);
    tile_load(sC, CP, color + g0 * S * 3, n, 3 * S);
    tile_load_wait();
    __syncthreads();
    if (t < n) {
      const size_t g = g0 + t;
      float* a = sA + t * SP; const float* zz = sZ + t * SP; float* c = sC + t * CP; float* tt = sT + t * SP;
      float T = 1.0f, D = 0.f, O = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
      for (int i = 0; i < S; ++i) {
        const float occ = sigmoid_acc(a[i]);
        const float w = occ * T;                                               // render_rays.py:33
        D += w * zz[i]; O += w; C0 += w * c[3 * i]; C1 += w * c[3 * i + 1]; C2 += w * c[3 * i + 2];
        if (BWD) { tt[i] = T; a[i] = occ; } else a[i] = w;
        T *= __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f);                          // render_rays.py:28
      }
      float var = 0.f;
      for (int i = 0; i < S; ++i) {
        const float w = BWD ? a[i] * tt[i] : a[i];
        const float dz = zz[i] - D;
        var += w * dz * dz;                                                    // loss.py:45-46 (detached)
      }
      const uint8_t lab = labels[g];
      const float mo = (lab != 0) ? 1.f : 0.f, ms = (lab != 2) ? 1.f : 0.f;
      const float md = (mask_depth[g] != 0 && lab != 0) ? 1.f : 0.f;
      const float info = 1.0f / (sqrtf(var) + 1e-4f);                          // render_rays.py:73-80
      const float gd = gt_depth[g], q0 = gt_color[g * 3], q1 = gt_color[g * 3 + 1], q2 = gt_color[g * 3 + 2];
      if (FWD) {
        float* orow = out_rays + g * 6;
        orow[0] = D; orow[1] = var; orow[2] = C0; orow[3] = C1; orow[4] = C2; orow[5] = O;
        ld += fabsf(D - gd) * md * info;                                       // loss.py:52-54
        lc += (fabsf(C0 - q0) + fabsf(C1 - q1) + fabsf(C2 - q2)) * mo;
        lo += fabsf(O - mo) * ms;                                              // loss.py:64-66
      }
      if (BWD) {
        auto sgn = [](float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); };
        const float gD = sgn(D - gd) * md * info * iNd;
        const float gC0 = sgn(C0 - q0) * mo * iNo, gC1 = sgn(C1 - q1) * mo * iNo, gC2 = sgn(C2 - q2) * mo * iNo;
        const float gO = sgn(O - mo) * ms * iNs;
        float suffix = 0.f;                                                    // sum_{i>j} G_i w_i
        for (int j = S - 1; j >= 0; --j) {
          const float occ = a[j], Tj = tt[j], w = occ * Tj;
          const float Gi = gD * zz[j] + gC0 * c[3 * j] + gC1 * c[3 * j + 1] + gC2 * c[3 * j + 2] + gO;
          const float f = __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f);
          const float docc = Gi * Tj - suffix / f;
          a[j] = docc * occ * (1.0f - occ);
          c[3 * j] = w * gC0; c[3 * j + 1] = w * gC1; c[3 * j + 2] = w * gC2;
          suffix += Gi * w;
        }
      }
    }
    __syncthreads();
    if (BWD) {
      tile_store(g_alpha + g0 * S, sA, SP, n, S);
      tile_store(g_color + g0 * S * 3, sC, CP, n, 3 * S);
      __syncthreads();
    }
  }
  if (FWD) {
    __shared__ float s_p[TR / 32][3];
    const int lane = t & 31, warp = t >> 5;
    ld = warp_sum(ld); lc = warp_sum(lc); lo = warp_sum(lo);
    if (lane == 0) { s_p[warp][0] = ld; s_p[warp][1] = lc; s_p[warp][2] = lo; }
    __syncthreads();
    if (t < 3) {
      float v = 0.f;
      for (int i = 0; i < TR / 32; ++i) v += s_p[i][t];
      partial[((size_t)k * gridDim.x + blockIdx.x) * 3 + t] = v;               // deterministic 2-stage reduce
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Dense, double-buffered variant of the tile kernel (the default whenever gcd(S, 32) <= 4, e.g. S = 10 and 14).
// ncu of composite_tile_kernel<0> at 1 Mi rays x 10 samples: 45 % of its 1120 instructions per ray staged the tile -- one 4-byte
// cp.async per element with a division to find the padded (odd-stride) row -- and a block alternated load / compute phases.
// Here a tile keeps the GLOBAL layout (rows of S floats back to back), so each of alpha / z / colour of a 128-ray tile is ONE
// contiguous chunk that thread 0 fetches with a 1-D bulk copy (TMA) completing on an mbarrier; two stages, so tile i + 1 arrives while
// tile i is composited, and the backward's gradients leave as two bulk stores from the same buffers.  Row stride S words means a
// gcd(S, 32)-way bank conflict on the per-thread row accesses (2-way at S = 10 / 14: ~5 M shared wavefronts per 1 Mi rays, far from
// the limit); S with gcd > 4 (8, 16, 24, 32) keep the padded kernel above.  Tiles that a bulk copy cannot move (the category's base
// offset not 16-byte aligned, or the ragged last tile) are staged with dense 4-byte cp.async (no division) and stored by a plain loop.
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ inline int dense_stage_floats(int S, bool bwd) { return TR * S * (bwd ? 3 : 2) + TR * 3 * S; }

template <int MODE>
__global__ void __launch_bounds__(TR)
composite_dense_kernel(int K, int R, int S, const float* __restrict__ alpha, const float* __restrict__ color,
                       const float* __restrict__ z, const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                       const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask_depth,
                       const int32_t* __restrict__ counts, float cs, float os, const float* __restrict__ g_loss,
                       float* __restrict__ out_rays, float* __restrict__ partial, float* __restrict__ g_alpha, float* __restrict__ g_color,
                       unsigned* __restrict__ ticket /* zeroed by the host; the last CTA to finish writes out_loss */, float* __restrict__ out_loss) {
  using namespace umma;
  extern __shared__ __align__(128) float sm_d[];
  __shared__ __align__(8) uint64_t full[2];
  constexpr bool BWD = MODE != 0, FWD = MODE != 1;
  const int k = blockIdx.y, t = threadIdx.x;
  const int stage_floats = dense_stage_floats(S, BWD);
  const int ntiles = (R + TR - 1) / TR;
  const size_t kbase = (size_t)k * R;
  // bulk copies need 16-byte aligned global addresses: every full tile of this category qualifies iff its first one does
  const bool aligned = ((reinterpret_cast<uintptr_t>(alpha + kbase * S) | reinterpret_cast<uintptr_t>(z + kbase * S) |
                         reinterpret_cast<uintptr_t>(color + kbase * S * 3) |
                         (BWD ? (reinterpret_cast<uintptr_t>(g_alpha + kbase * S) | reinterpret_cast<uintptr_t>(g_color + kbase * S * 3)) : 0)) & 15) == 0;
  auto bulk_ok = [&](int tile) { return aligned && tile < ntiles && (tile + 1) * TR <= R; };
  if (t == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_mbar_init(); }
  __syncthreads();
  auto issue = [&](int tile, int st) {          // thread 0: the three chunks of a full tile -> stage st
    float* base = sm_d + st * stage_floats;
    const size_t g0 = kbase + (size_t)tile * TR;
    const uint32_t bA = TR * S * 4u;
    mbar_expect_tx(&full[st], 5u * bA);
    bulk_g2s(base, alpha + g0 * S, bA, &full[st]);
    bulk_g2s(base + TR * S, z + g0 * S, bA, &full[st]);
    bulk_g2s(base + 2 * TR * S, color + g0 * S * 3, 3u * bA, &full[st]);
  };
  float iNd = 0.f, iNo = 0.f, iNs = 0.f;
  if (BWD) {
    bool e0, e1, e2;
    empty_flags(K, counts, e0, e1, e2);
    const float gl = g_loss ? g_loss[0] : 1.0f;
    iNd = e0 ? 0.f : gl / ((float)counts[k * 4 + 0] + 1e-10f);
    iNo = e1 ? 0.f : gl * cs / ((float)counts[k * 4 + 1] + 1e-10f);
    iNs = e2 ? 0.f : gl * os / ((float)counts[k * 4 + 2] + 1e-10f);
  }
  float ld = 0.f, lc = 0.f, lo = 0.f;
  uint32_t ph0 = 0, ph1 = 0;
  int tile = blockIdx.x;
  if (t == 0 && bulk_ok(tile)) issue(tile, 0);
  for (int j = 0; tile < ntiles; ++j, tile += gridDim.x) {
    const int st = j & 1;
    float* sA = sm_d + st * stage_floats;          // alpha -> (fwd) weights w_i / (bwd) occupancy -> d alpha
    float* sZ = sA + TR * S;
    float* sC = sZ + TR * S;                       // colour -> (bwd) d colour
    float* sT = sC + TR * 3 * S;                   // (bwd) transmittance T_i
    const int r0 = tile * TR, n = min(TR, R - r0);
    const size_t g0 = kbase + r0;
    const int nxt = tile + gridDim.x;
    if (t == 0 && bulk_ok(nxt)) {                  // prefetch: stage st^1 was released by the barrier that ended iteration j-1
      if (BWD) bulk_wait_read<0>();                // ... and its gradient stores have finished reading it
      issue(nxt, st ^ 1);
    }
    if (bulk_ok(tile)) {
      mbar_wait(&full[st], st ? ph1 : ph0);
      if (st) ph1 ^= 1; else ph0 ^= 1;
    } else {
      if (BWD && t == 0) bulk_wait_read<0>();      // an earlier bulk store may still read this stage
      __syncthreads();
      const uint32_t dA = smem_u32(sA), dZ = smem_u32(sZ), dC = smem_u32(sC);
      const float* pa = alpha + g0 * S; const float* pz = z + g0 * S; const float* pc = color + g0 * S * 3;
      for (int e = t; e < n * S; e += TR) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dA + 4u * e), "l"(pa + e) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dZ + 4u * e), "l"(pz + e) : "memory");
      }
      for (int e = t; e < n * S * 3; e += TR) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dC + 4u * e), "l"(pc + e) : "memory");
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    if (t < n) {
      const size_t g = g0 + t;
      float* a = sA + t * S; const float* zz = sZ + t * S; float* c = sC + t * 3 * S; float* tt = sT + t * S;
      // per-ray inputs first: their latency overlaps the compositing loop
      const uint8_t lab = labels[g];
      const uint8_t mdep = mask_depth[g];
      const float gd = gt_depth[g], q0 = gt_color[g * 3], q1 = gt_color[g * 3 + 1], q2 = gt_color[g * 3 + 2];
      float T = 1.0f, D = 0.f, O = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
      for (int i = 0; i < S; ++i) {
        const float occ = sigmoid_fast(a[i]);
        const float w = occ * T;                                               // render_rays.py:33
        D += w * zz[i]; O += w; C0 += w * c[3 * i]; C1 += w * c[3 * i + 1]; C2 += w * c[3 * i + 2];
        if (BWD) { tt[i] = T; a[i] = occ; } else a[i] = w;
        T *= __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f);                          // render_rays.py:28
      }
      float var = 0.f;
      for (int i = 0; i < S; ++i) {
        const float w = BWD ? a[i] * tt[i] : a[i];
        const float dz = zz[i] - D;
        var += w * dz * dz;                                                    // loss.py:45-46 (detached)
      }
      const float mo = (lab != 0) ? 1.f : 0.f, ms = (lab != 2) ? 1.f : 0.f;
      const float md = (mdep != 0 && lab != 0) ? 1.f : 0.f;
      const float info = 1.0f / (sqrtf(var) + 1e-4f);                          // render_rays.py:73-80
      if (FWD) {
        float* orow = out_rays + g * 6;
        orow[0] = D; orow[1] = var; orow[2] = C0; orow[3] = C1; orow[4] = C2; orow[5] = O;
        ld += fabsf(D - gd) * md * info;                                       // loss.py:52-54
        lc += (fabsf(C0 - q0) + fabsf(C1 - q1) + fabsf(C2 - q2)) * mo;
        lo += fabsf(O - mo) * ms;                                              // loss.py:64-66
      }
      if (BWD) {
        auto sgn = [](float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); };
        const float gD = sgn(D - gd) * md * info * iNd;
        const float gC0 = sgn(C0 - q0) * mo * iNo, gC1 = sgn(C1 - q1) * mo * iNo, gC2 = sgn(C2 - q2) * mo * iNo;
        const float gO = sgn(O - mo) * ms * iNs;
        float suffix = 0.f;                                                    // sum_{i>j} G_i w_i
        for (int jj = S - 1; jj >= 0; --jj) {
          const float occ = a[jj], Tj = tt[jj], w = occ * Tj;
          const float Gi = gD * zz[jj] + gC0 * c[3 * jj] + gC1 * c[3 * jj + 1] + gC2 * c[3 * jj + 2] + gO;
          const float f = __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f);
          const float docc = Gi * Tj - __fdividef(suffix, f);
          a[jj] = docc * occ * (1.0f - occ);
          c[3 * jj] = w * gC0; c[3 * jj + 1] = w * gC1; c[3 * jj + 2] = w * gC2;
          suffix += Gi * w;
        }
      }
    }
    if (BWD) {
      if (bulk_ok(tile)) {
        fence_async_smem();                        // generic-proxy writes of the gradients -> visible to the bulk-copy engine
        __syncthreads();
        if (t == 0) {
          bulk_s2g(g_alpha + g0 * S, sA, TR * S * 4u);
          bulk_s2g(g_color + g0 * S * 3, sC, TR * S * 12u);
          bulk_commit();
        }
      } else {
        __syncthreads();
        for (int e = t; e < n * S; e += TR) g_alpha[g0 * S + e] = sA[e];
        for (int e = t; e < n * S * 3; e += TR) g_color[g0 * S * 3 + e] = sC[e];
        __syncthreads();
      }
    } else {
      __syncthreads();                             // every thread is done with stage st before it is refilled
    }
  }
  if (BWD && t == 0) bulk_wait<0>();               // the last gradient stores have left shared memory (and are complete) before exit
  if (FWD) {
    __shared__ float s_p[TR / 32][3];
    const int lane = t & 31, warp = t >> 5;
    ld = warp_sum(ld); lc = warp_sum(lc); lo = warp_sum(lo);
    if (lane == 0) { s_p[warp][0] = ld; s_p[warp][1] = lc; s_p[warp][2] = lo; }
    __syncthreads();
    if (t < 3) {
      float v = 0.f;
      for (int i = 0; i < TR / 32; ++i) v += s_p[i][t];
      partial[((size_t)k * gridDim.x + blockIdx.x) * 3 + t] = v;               // deterministic 2-stage reduce
      __threadfence();
    }
    __shared__ int s_last;
    __syncthreads();
    if (t == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (s_last) {                                  // every other CTA's partial sums are visible (fence + ticket)
      __threadfence();
      finalize_block(K, gridDim.x, partial, counts, cs, os, out_loss);
    }
  }
}

__global__ void termination_kernel(int n_rays, int S, const float* __restrict__ occ_in, float* __restrict__ w_out) {
  const int G = seg_width(S), rpw = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = lane / G, sl = lane - seg * G;
  const int rays_per_block = (blockDim.x / 32) * rpw;
  for (int rb = blockIdx.x * rays_per_block; rb < n_rays; rb += gridDim.x * rays_per_block) {
    const int r = rb + warp * rpw + seg;
    const int rr = min(r, n_rays - 1);
    float carry = 1.0f;
    for (int c0 = 0; c0 < S; c0 += G) {
      const int i = c0 + sl;
      const bool act = i < S;
      const float occ = act ? occ_in[(size_t)rr * S + i] : 0.f;
      const float f = act ? __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f) : 1.0f;
      const float incl = seg_scan_mul(f, G, sl);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (sl == 0) excl = 1.0f;
      if (act && r < n_rays) w_out[(size_t)rr * S + i] = occ * (carry * excl);
      carry *= __shfl_sync(0xffffffffu, incl, G - 1, G);
    }
  }
}

// Uncertainty fields of the registration sweep (src/category_registration.py:152-162): from the termination weights of a ray,
// entropy = sum_i -w_i log(w_i + 1e-10), mass = sum_i w_i and heuristic = mass * exp(-0.5 entropy).  One warp per ray (S = 96 there);
// when `occ_in` is given the weights are composited on the fly from the occupancies (scan with a carry, as termination_kernel).
__global__ void ray_uncertainty_kernel(int n_rays, int S, const float* __restrict__ w_in, float* __restrict__ out /* [n_rays,3] */) {
  const int lane = threadIdx.x & 31, warps = (blockDim.x >> 5);
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < n_rays; r += gridDim.x * warps) {
    const float* w = w_in + (size_t)r * S;
    float ent = 0.f, mass = 0.f;
    for (int i = lane; i < S; i += 32) {
      const float v = w[i];
      ent -= v * logf(v + 1e-10f);
      mass += v;
    }
    ent = warp_sum(ent); mass = warp_sum(mass);
    if (lane == 0) { out[(size_t)r * 3] = ent; out[(size_t)r * 3 + 1] = mass; out[(size_t)r * 3 + 2] = mass * expf(-0.5f * ent); }
  }
}

// Backward of occupancy_to_termination (src/render_rays.py:25-33 under autograd): w_i = occ_i T_i, T_i = prod_{j<i} f_j,
// f_j = 1 - occ_j + 1e-10.  d occ_j = T_j (g_j - Q_j) with Q_j = sum_{i>j} g_i occ_i prod_{j<m<i} f_m, i.e. the reverse recurrence
// Q_{j} = g_{j+1} occ_{j+1} + f_{j+1} Q_{j+1} -- division-free, so saturated samples (f = 1e-10) need no special case.  One thread
// per ray (module-level API, off the training hot path): pass 1 parks T_j in the output, pass 2 walks backwards.
__global__ void termination_bwd_kernel(int n_rays, int S, const float* __restrict__ occ_in, const float* __restrict__ g_w,
                                       float* __restrict__ g_occ) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const float* occ = occ_in + (size_t)r * S;
  const float* g = g_w + (size_t)r * S;
  float* out = g_occ + (size_t)r * S;
  float T = 1.0f;
  for (int i = 0; i < S; ++i) { out[i] = T; T *= __fadd_rn(__fsub_rn(1.0f, occ[i]), 1e-10f); }
  float Q = 0.f;
  for (int j = S - 1; j >= 0; --j) {
    out[j] = out[j] * (g[j] - Q);
    Q = fmaf(__fadd_rn(__fsub_rn(1.0f, occ[j]), 1e-10f), Q, g[j] * occ[j]);
  }
}

// Inference compositing of points whose occupancy is already activated: pt[r*S+i] = {occ, r, g, b}; one ray per G-lane warp
// segment, exclusive cumprod by shuffle scans over chunks of G samples with a carry (render_rays.py:25-50, loss.py:44-48).
__global__ void __launch_bounds__(CT)
render_composite_kernel(int n_rays, int S, const float4* __restrict__ pt, const float* __restrict__ z,
                        float* __restrict__ out_rays, float* __restrict__ out_w) {
  const int G = seg_width(S), rpw = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = lane / G, sl = lane - seg * G;
  const int rays_per_block = (CT / 32) * rpw;
  for (int rb = blockIdx.x * rays_per_block; rb < n_rays; rb += gridDim.x * rays_per_block) {
    const int r = rb + warp * rpw + seg;
    const int rr = min(r, n_rays - 1);
    const float4* p = pt + (size_t)rr * S;
    const float* zr = z + (size_t)rr * S;
    float carry = 1.0f, D = 0.f, O = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    for (int c0 = 0; c0 < S; c0 += G) {
      const int i = c0 + sl;
      const bool act = i < S;
      const float4 v = act ? p[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float f = act ? __fadd_rn(__fsub_rn(1.0f, v.x), 1e-10f) : 1.0f;     // render_rays.py:28
      const float incl = seg_scan_mul(f, G, sl);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (sl == 0) excl = 1.0f;
      const float w = v.x * (carry * excl);                                     // render_rays.py:33
      carry *= __shfl_sync(0xffffffffu, incl, G - 1, G);
      if (act) {
        D += w * zr[i]; O += w; C0 += w * v.y; C1 += w * v.z; C2 += w * v.w;
        if (out_w != nullptr && r < n_rays) out_w[(size_t)rr * S + i] = w;
      }
    }
    D = seg_sum(D, G); O = seg_sum(O, G); C0 = seg_sum(C0, G); C1 = seg_sum(C1, G); C2 = seg_sum(C2, G);
    float var = 0.f;                                                            // second pass: needs D (loss.py:45)
    carry = 1.0f;
    for (int c0 = 0; c0 < S; c0 += G) {
      const int i = c0 + sl;
      const bool act = i < S;
      const float occ = act ? p[i].x : 0.f;
      const float f = act ? __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f) : 1.0f;
      const float incl = seg_scan_mul(f, G, sl);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (sl == 0) excl = 1.0f;
      const float w = occ * (carry * excl);
      carry *= __shfl_sync(0xffffffffu, incl, G - 1, G);
      if (act) { const float dz = zr[i] - D; var += w * dz * dz; }
    }
    var = seg_sum(var, G);
    if (sl == 0 && r < n_rays) {
      float* o = out_rays + (size_t)r * 6;
      o[0] = D; o[1] = var; o[2] = C0; o[3] = C1; o[4] = C2; o[5] = O;
    }
  }
}

int render_composite_launch(int n_rays, int S, const float* pt, const float* z, float* out_rays, float* out_w, cudaStream_t st) {
  if (n_rays <= 0) return CNF_OK;
  const int G = S >= 32 ? 32 : (S > 16 ? 32 : (S > 8 ? 16 : (S > 4 ? 8 : (S > 2 ? 4 : (S > 1 ? 2 : 1)))));
  const int rays_per_block = (CT / 32) * (32 / G);
  long long blocks = ((long long)n_rays + rays_per_block - 1) / rays_per_block;
  if (blocks > 148 * 32) blocks = 148 * 32;
  render_composite_kernel<<<(int)blocks, CT, 0, st>>>(n_rays, S, reinterpret_cast<const float4*>(pt), z, out_rays, out_w);
  return check_launch("render_composite");
}

// one block; warp per code row
__global__ void code_reg_kernel(int n_obj, int L, const float* __restrict__ codes, float weight,
                                float* __restrict__ out, float* __restrict__ g_codes) {
  __shared__ float s_n[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float acc = 0.f;
  for (int i = warp; i < n_obj; i += nw) {
    float s = 0.f;
    for (int j = lane; j < L; j += 32) { const float v = codes[(size_t)i * L + j]; s += v * v; }
    s = warp_sum(s);
    const float nrm = sqrtf(s);
    acc += nrm;
    if (g_codes && n_obj > 1) {
      const float inv = nrm > 0.f ? weight / nrm : 0.f;
      for (int j = lane; j < L; j += 32) g_codes[(size_t)i * L + j] += inv * codes[(size_t)i * L + j];
    }
  }
  if (lane == 0) s_n[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += s_n[i];
    out[0] = (n_obj > 1) ? t : 0.f;                                          // loss.py:12
  }
}

}  // namespace cnf

using namespace cnf;

static int comp_blocks(int R, int S) {
  int G = 1; while (G < S && G < 32) G <<= 1;
  const int rpb = (CT / 32) * (32 / G);
  int nb = (R + rpb - 1) / rpb;
  if (nb > 1184) nb = 1184;      // 8 x 148 SMs, grid-stride beyond
  return nb < 1 ? 1 : nb;
}

// dense double-buffered kernel: chosen when the row stride S gives at most a 4-way bank conflict
static bool use_dense(int S) { int g = 1; while (g < 32 && S % (g * 2) == 0) g *= 2; return g <= 4; }
static int dense_smem_bytes(int S, bool bwd) { return 2 * dense_stage_floats(S, bwd) * (int)sizeof(float); }
static int dense_blocks(int R, int S, bool bwd) {
  int per_sm = (220 * 1024) / (dense_smem_bytes(S, bwd) + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int nb = (R + TR - 1) / TR;
  if (nb > 148 * per_sm) nb = 148 * per_sm;      // persistent CTAs, grid-stride over the tiles (<= 1184: the partial-sum workspace)
  return nb < 1 ? 1 : nb;
}
template <int MODE>
static int dense_attr(int S) {
  static DeviceOnce once;
  if (once.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(composite_dense_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dense_smem_bytes(31, MODE != 0)));
    CNF_CUDA(cudaFuncSetAttribute(composite_dense_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    once.done();
  }
  (void)S;
  return CNF_OK;
}

static int tile_blocks(int R) {
  int nb = (R + TR - 1) / TR;
  if (nb > 1184) nb = 1184;      // partial-sum workspace is sized for 8 x 148 blocks; grid-stride beyond
  return nb < 1 ? 1 : nb;
}

extern "C" int cnf_termination_fwd(int n_rays, int S, const float* occ, float* w, cnf_stream_t stream) {
  CNF_REQUIRE(n_rays >= 0 && S > 0 && occ && w, "cnf_termination_fwd: bad arguments");
  if (n_rays == 0) return CNF_OK;
  termination_kernel<<<comp_blocks(n_rays, S), CT, 0, as_stream(stream)>>>(n_rays, S, occ, w);
  return check_launch("cnf_termination_fwd");
}

extern "C" int cnf_ray_uncertainty(int n_rays, int S, const float* w, float* out, cnf_stream_t stream) {
  CNF_REQUIRE(n_rays >= 0 && S > 0 && w && out, "cnf_ray_uncertainty: bad arguments");
  if (n_rays == 0) return CNF_OK;
  int nb = (n_rays + 7) / 8; if (nb > 148 * 8) nb = 148 * 8;
  ray_uncertainty_kernel<<<nb, 256, 0, as_stream(stream)>>>(n_rays, S, w, out);
  return check_launch("cnf_ray_uncertainty");
}

extern "C" int cnf_termination_bwd(int n_rays, int S, const float* occ, const float* g_w, float* g_occ, cnf_stream_t stream) {
  CNF_REQUIRE(n_rays >= 0 && S > 0 && occ && g_w && g_occ, "cnf_termination_bwd: bad arguments");
  if (n_rays == 0) return CNF_OK;
  termination_bwd_kernel<<<(n_rays + 127) / 128, 128, 0, as_stream(stream)>>>(n_rays, S, occ, g_w, g_occ);
  return check_launch("cnf_termination_bwd");
}

extern "C" int cnf_mask_counts(int K, int R, const uint8_t* labels, const uint8_t* mask_depth, int32_t* counts,
                               cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && R > 0 && labels && mask_depth && counts, "cnf_mask_counts: bad arguments");
  CNF_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4 * K, as_stream(stream)));
  int nb = (R + 255) / 256; if (nb > 296) nb = 296;
  mask_counts_kernel<<<dim3(nb, K), 256, 0, as_stream(stream)>>>(R, labels, mask_depth, counts);
  return check_launch("cnf_mask_counts");
}

extern "C" size_t cnf_composite_workspace_bytes(int K, int R) { (void)R; return (size_t)K * 1184 * 3 * sizeof(float) + 16; }   // partial sums | ticket

extern "C" int cnf_composite_loss_fwd(int K, int R, int S, const float* alpha, const float* color, const float* z,
                                      const float* gt_depth, const float* gt_color, const uint8_t* labels,
                                      const uint8_t* mask_depth, const int32_t* counts, float cs, float os,
                                      float* out_rays, float* out_loss, void* ws, size_t ws_bytes, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && R > 0 && S > 0, "cnf_composite_loss_fwd: bad sizes");
  CNF_REQUIRE(alpha && color && z && gt_depth && gt_color && labels && mask_depth && counts && out_rays && out_loss && ws,
              "cnf_composite_loss_fwd: null pointer");
  if (ws_bytes < cnf_composite_workspace_bytes(K, R)) { set_error("cnf_composite_loss_fwd: workspace too small"); return CNF_ERR_WORKSPACE; }
  int nb;
  if (S <= 32 && use_dense(S)) {
    nb = dense_blocks(R, S, false);
    if (int rc = dense_attr<0>(S)) return rc;
    unsigned* ticket = reinterpret_cast<unsigned*>((float*)ws + (size_t)K * 1184 * 3);
    CNF_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), as_stream(stream)));
    composite_dense_kernel<0><<<dim3(nb, K), TR, dense_smem_bytes(S, false), as_stream(stream)>>>(K, R, S, alpha, color, z, gt_depth, gt_color, labels, mask_depth,
                                                                                                   counts, cs, os, nullptr, out_rays, (float*)ws, nullptr, nullptr,
                                                                                                   ticket, out_loss);
    return check_launch("cnf_composite_loss_fwd");
  } else if (S <= 32) {
    nb = tile_blocks(R);
    const int smem = tile_smem_floats(S, false) * 4;
    static DeviceOnce attr;
    if (attr.pending()) { CNF_CUDA(cudaFuncSetAttribute(composite_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem_floats(32, false) * 4)); attr.done(); }
    composite_tile_kernel<0><<<dim3(nb, K), TR, smem, as_stream(stream)>>>(K, R, S, alpha, color, z, gt_depth, gt_color, labels, mask_depth,
                                                                              counts, cs, os, nullptr, out_rays, (float*)ws, nullptr, nullptr);
  } else {
    nb = comp_blocks(R, S);
    composite_fwd_kernel<<<dim3(nb, K), CT, 0, as_stream(stream)>>>(K, R, S, alpha, color, z, gt_depth, gt_color, labels,
                                                                    mask_depth, out_rays, (float*)ws);
  }
  composite_finalize_kernel<<<1, nb > 64 ? 1024 : 32, 0, as_stream(stream)>>>(K, nb, (const float*)ws, counts, cs, os, out_loss);
  return check_launch("cnf_composite_loss_fwd");
}

extern "C" int cnf_composite_loss_bwd(int K, int R, int S, const float* alpha, const float* color, const float* z,
                                      const float* gt_depth, const float* gt_color, const uint8_t* labels,
                                      const uint8_t* mask_depth, const int32_t* counts, float cs, float os,
                                      const float* g_loss, float* g_alpha, float* g_color, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && R > 0 && S > 0, "cnf_composite_loss_bwd: bad sizes");
  CNF_REQUIRE(alpha && color && z && gt_depth && gt_color && labels && mask_depth && counts && g_alpha && g_color,
              "cnf_composite_loss_bwd: null pointer");
  if (S <= 32 && use_dense(S)) {
    if (int rc = dense_attr<1>(S)) return rc;
    composite_dense_kernel<1><<<dim3(dense_blocks(R, S, true), K), TR, dense_smem_bytes(S, true), as_stream(stream)>>>(
        K, R, S, alpha, color, z, gt_depth, gt_color, labels, mask_depth, counts, cs, os, g_loss, nullptr, nullptr, g_alpha, g_color, nullptr, nullptr);
    return check_launch("cnf_composite_loss_bwd");
  }
  if (S <= 32) {
    const int smem = tile_smem_floats(S, true) * 4;
    static DeviceOnce attr;
    if (attr.pending()) { CNF_CUDA(cudaFuncSetAttribute(composite_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem_floats(32, true) * 4)); attr.done(); }
    composite_tile_kernel<1><<<dim3(tile_blocks(R), K), TR, smem, as_stream(stream)>>>(K, R, S, alpha, color, z, gt_depth, gt_color, labels,
                                                                                        mask_depth, counts, cs, os, g_loss, nullptr, nullptr, g_alpha, g_color);
    return check_launch("cnf_composite_loss_bwd");
  }
  composite_bwd_kernel<<<dim3(comp_blocks(R, S), K), CT, 0, as_stream(stream)>>>(
      K, R, S, alpha, color, z, gt_depth, gt_color, labels, mask_depth, counts, cs, os, g_loss, g_alpha, g_color);
  return check_launch("cnf_composite_loss_bwd");
}

// forward and backward in one pass (upstream gradient of the loss = 1): S <= 32
extern "C" int cnf_composite_loss_fwd_bwd(int K, int R, int S, const float* alpha, const float* color, const float* z,
                                          const float* gt_depth, const float* gt_color, const uint8_t* labels,
                                          const uint8_t* mask_depth, const int32_t* counts, float cs, float os,
                                          float* out_rays, float* out_loss, float* g_alpha, float* g_color,
                                          void* ws, size_t ws_bytes, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && R > 0 && S > 0, "cnf_composite_loss_fwd_bwd: bad sizes");
  CNF_REQUIRE(alpha && color && z && gt_depth && gt_color && labels && mask_depth && counts && out_rays && out_loss && g_alpha && g_color && ws,
              "cnf_composite_loss_fwd_bwd: null pointer");
  if (S > 32) { set_error("cnf_composite_loss_fwd_bwd: S=%d > 32 (use the separate forward / backward entry points)", S); return CNF_ERR_UNSUPPORTED; }
  if (ws_bytes < cnf_composite_workspace_bytes(K, R)) { set_error("cnf_composite_loss_fwd_bwd: workspace too small"); return CNF_ERR_WORKSPACE; }
  if (use_dense(S)) {
    const int nbd = dense_blocks(R, S, true);
    if (int rc = dense_attr<2>(S)) return rc;
    unsigned* ticket = reinterpret_cast<unsigned*>((float*)ws + (size_t)K * 1184 * 3);
    CNF_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), as_stream(stream)));
    composite_dense_kernel<2><<<dim3(nbd, K), TR, dense_smem_bytes(S, true), as_stream(stream)>>>(K, R, S, alpha, color, z, gt_depth, gt_color, labels, mask_depth,
                                                                                                   counts, cs, os, nullptr, out_rays, (float*)ws, g_alpha, g_color,
                                                                                                   ticket, out_loss);
    return check_launch("cnf_composite_loss_fwd_bwd");
  }
  const int nb = tile_blocks(R);
  const int smem = tile_smem_floats(S, true) * 4;
  static DeviceOnce attr;
  if (attr.pending()) { CNF_CUDA(cudaFuncSetAttribute(composite_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem_floats(32, true) * 4)); attr.done(); }
  composite_tile_kernel<2><<<dim3(nb, K), TR, smem, as_stream(stream)>>>(K, R, S, alpha, color, z, gt_depth, gt_color, labels, mask_depth,
                                                                        counts, cs, os, nullptr, out_rays, (float*)ws, g_alpha, g_color);
  composite_finalize_kernel<<<1, nb > 64 ? 1024 : 32, 0, as_stream(stream)>>>(K, nb, (const float*)ws, counts, cs, os, out_loss);
  return check_launch("cnf_composite_loss_fwd_bwd");
}

extern "C" int cnf_code_reg(int n_obj, int L, const float* codes, float weight, float* out, float* g_codes,
                            cnf_stream_t stream) {
  CNF_REQUIRE(n_obj > 0 && L > 0 && codes && out, "cnf_code_reg: bad arguments");
  code_reg_kernel<<<1, 256, 0, as_stream(stream)>>>(n_obj, L, codes, weight, out, g_codes);
  return check_launch("cnf_code_reg");
}
