// Depth-guided ray sampler: sceneCategory.sample_3d_points / stratified_bins / normal_bins_sampling
// (reference src/scene_cateogries.py:51-96, 453-546).  Bit-exact against the reference: every fp32 op is
// separately rounded (__fmul_rn/__fadd_rn/__fdiv_rn, no FMA contraction), linspace tables and random
// draws are inputs, and rows of the random tensors are consumed by rank-within-case in ray order.
//
// HBM-bound.  Two launches (one for up to 4096 rays per category): (A) per-block case counts + max depth, whose last CTA per
// category also (B) scans the block counts, (C) emit: one thread per ray (M <= 16 surface bins) or per sample (generic).
#include "cnf_common.cuh"
#include "cnf_sample_math.cuh"

namespace cnf {

constexpr int SB = 256;  // rays per block

// (A) counts per 256-ray block: {nI, nO, nT} plus loss-mask counts {Nd, No, Ns} and block max depth.  One CTA covers COUNT_RB
// consecutive ray blocks with all of a thread's loads in flight (one ray block per CTA made the kernel a chain of tiny
// latency-bound CTAs: 770 GB/s); warp w counts the rays [32w, 32w + 32) of every ray block by ballots.  The CTA writes, per ray
// block, the exclusive prefix of the case counts INSIDE the CTA, and its own totals.
// (B) The LAST CTA of a category to finish (ticket counter) scans the CTA totals (nb / COUNT_RB entries: one round trip) into
// cta_prefix and writes the category's stats row; an emit block's running counts are cta_prefix[b / COUNT_RB] + blk_counts[b].
// A separate one-CTA scan launch over all ray blocks cost 15-20 us of a 150 us sampler at 1 Mi rays.
constexpr int COUNT_RB = 8, SCAN_BATCH = 4;
__global__ void __launch_bounds__(SB, 4)
sample_count_kernel(int R, int nb, const uint8_t* __restrict__ rgbs, const float* __restrict__ depth,
                    float min_bound, int this_obj, int* __restrict__ blk_counts /* [K][nb][4] */,
                    int4* __restrict__ cta_lo /* [K][gridDim.x] {nI, nO, nT, bits(max depth)} */,
                    int4* __restrict__ cta_hi /* [K][gridDim.x] {Nd, No, Ns, -} */,
                    int4* __restrict__ cta_prefix /* [K][gridDim.x] exclusive prefix of {nI, nO, nT} over the CTAs */,
                    unsigned* __restrict__ ticket /* [K], zeroed by the host */, const float* __restrict__ max_override,
                    int32_t* __restrict__ stats) {
  const int k = blockIdx.y, b0 = blockIdx.x * COUNT_RB;
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  const uint32_t* px = reinterpret_cast<const uint32_t*>(rgbs);
  __shared__ int s_c[COUNT_RB][SB / 32][6];
  __shared__ float s_m[COUNT_RB][SB / 32];
  __shared__ int s_t[COUNT_RB][8];
  {
    float d[COUNT_RB];
    uint32_t p[COUNT_RB];
#pragma unroll
    for (int j = 0; j < COUNT_RB; ++j) {
      const int r = (b0 + j) * SB + t;
      const bool in = r < R;
      d[j] = in ? depth[(size_t)k * R + r] : 0.f;
      p[j] = in ? px[(size_t)k * R + r] : 0u;
    }
#pragma unroll
    for (int j = 0; j < COUNT_RB; ++j) {
      const int r = (b0 + j) * SB + t;
      const bool in = r < R;
      const uint8_t lab = (uint8_t)(p[j] >> 24);
      const int c = in ? ray_case(d[j], lab, min_bound, this_obj) : 3;
      const bool mo = in && lab != 0, ms = in && lab != 2;
      const unsigned q0 = __ballot_sync(0xffffffffu, c == 0), q1 = __ballot_sync(0xffffffffu, c == 1), q2 = __ballot_sync(0xffffffffu, c == 2);
      const unsigned q3 = __ballot_sync(0xffffffffu, c != 0 && mo), q4 = __ballot_sync(0xffffffffu, mo), q5 = __ballot_sync(0xffffffffu, ms);
      const float mx = warp_max(in ? d[j] : -INFINITY);
      if (l == 0) {
        int* o = s_c[j][w];
        o[0] = __popc(q0); o[1] = __popc(q1); o[2] = __popc(q2); o[3] = __popc(q3); o[4] = __popc(q4); o[5] = __popc(q5);
        s_m[j][w] = mx;
      }
    }
  }
  __syncthreads();
  if (t < COUNT_RB * 8) {                             // thread (j, cc): ray block j, counter cc (6 = max depth): totals of the ray block
    const int j = t >> 3, cc = t & 7;
    int tt = 0;
    if (cc < 6) {
#pragma unroll
      for (int i = 0; i < SB / 32; ++i) tt += s_c[j][i][cc];
    } else if (cc == 6) {
      float m = s_m[j][0];
#pragma unroll
      for (int i = 1; i < SB / 32; ++i) m = fmaxf(m, s_m[j][i]);
      tt = __float_as_int(m);
    }
    s_t[j][cc] = tt;
  }
  __syncthreads();
  if (t < COUNT_RB) {                                 // ray block t: exclusive prefix of the case counts inside this CTA
    int p0 = 0, p1 = 0, p2 = 0;
    for (int j = 0; j < t; ++j) { p0 += s_t[j][0]; p1 += s_t[j][1]; p2 += s_t[j][2]; }
    if (b0 + t < nb) reinterpret_cast<int4*>(blk_counts)[(size_t)k * nb + b0 + t] = make_int4(p0, p1, p2, 0);
  } else if (t == 32) {                               // totals of the CTA (rays beyond R counted nothing)
    int tot[6] = {0, 0, 0, 0, 0, 0};
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < COUNT_RB; ++j) {
#pragma unroll
      for (int q = 0; q < 6; ++q) tot[q] += s_t[j][q];
      m = fmaxf(m, __int_as_float(s_t[j][6]));
    }
    cta_lo[(size_t)k * gridDim.x + blockIdx.x] = make_int4(tot[0], tot[1], tot[2], __float_as_int(m));
    cta_hi[(size_t)k * gridDim.x + blockIdx.x] = make_int4(tot[3], tot[4], tot[5], 0);
    __threadfence();
  }
  // ---- (B) last CTA of the category: exclusive scan of the CTA totals, category totals -> stats
  __shared__ int s_last;
  __syncthreads();
  if (t == 0) s_last = atomicAdd(ticket + k, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nc = gridDim.x;
  const int4* lo4 = cta_lo + (size_t)k * nc;
  const int4* hi4 = cta_hi + (size_t)k * nc;
  int4* pre4 = cta_prefix + (size_t)k * nc;
  const int per = (nc + SB - 1) / SB;
  const int c0 = min(nc, t * per), c1 = min(nc, c0 + per);
  int sum[6] = {0, 0, 0, 0, 0, 0};
  float mx = -INFINITY;
  for (int bb = c0; bb < c1; bb += SCAN_BATCH) {
    int4 lo[SCAN_BATCH], hi[SCAN_BATCH];
#pragma unroll
    for (int j = 0; j < SCAN_BATCH; ++j) {
      const bool in = bb + j < c1;
      lo[j] = in ? __ldcg(lo4 + bb + j) : make_int4(0, 0, 0, __float_as_int(-INFINITY));
      hi[j] = in ? __ldcg(hi4 + bb + j) : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < SCAN_BATCH; ++j) {
      sum[0] += lo[j].x; sum[1] += lo[j].y; sum[2] += lo[j].z; sum[3] += hi[j].x; sum[4] += hi[j].y; sum[5] += hi[j].z;
      mx = fmaxf(mx, __int_as_float(lo[j].w));
    }
  }
  __shared__ int s_ws[SB / 32][6];
  __shared__ float s_wm[SB / 32];
  int excl[3];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    int v = sum[c];
    if (c < 3) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if (l >= o) v += n; }
      excl[c] = v - sum[c];
      if (l == 31) s_ws[w][c] = v;
    } else {
      v = warp_sum_i(v);
      if (l == 0) s_ws[w][c] = v;
    }
  }
  mx = warp_max(mx);
  if (l == 0) s_wm[w] = mx;
  __syncthreads();
  int tot[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    int before = 0, all = 0;
#pragma unroll
    for (int i = 0; i < SB / 32; ++i) { const int v = s_ws[i][c]; if (i < w) before += v; all += v; }
    if (c < 3) excl[c] += before;
    tot[c] = all;
  }
  int run[3] = {excl[0], excl[1], excl[2]};
  for (int bb = c0; bb < c1; bb += SCAN_BATCH) {
    int4 lo[SCAN_BATCH];
#pragma unroll
    for (int j = 0; j < SCAN_BATCH; ++j) lo[j] = (bb + j < c1) ? __ldcg(lo4 + bb + j) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < SCAN_BATCH; ++j) {
      if (bb + j < c1) {
        pre4[bb + j] = make_int4(run[0], run[1], run[2], 0);
        run[0] += lo[j].x; run[1] += lo[j].y; run[2] += lo[j].z;
      }
    }
  }
  if (t == 0) {
    float m = s_wm[0];
#pragma unroll
    for (int i = 1; i < SB / 32; ++i) m = fmaxf(m, s_wm[i]);
    int32_t* st = stats + k * 8;
    st[0] = tot[0];              // nI
    st[1] = tot[1] + tot[2];     // nV
    st[2] = tot[1];              // nO
    st[3] = tot[2];              // nT
    st[4] = tot[3];              // N_depth&obj
    st[5] = tot[4];              // N_obj
    st[6] = tot[5];              // N_sem
    st[7] = __float_as_int(max_override ? max_override[k] : m);   // max depth bits
  }
}


// (C) emit: one block of SB rays; `base` = number of earlier rays of the category in case 0 / 1 / 2 (ray order)
constexpr int EMIT_T = 1024;      // threads of the single-launch small-batch kernel: the first SB rank the rays, all of them write samples
constexpr int EMIT_BIG_T = 512;   // threads per 256-ray block of the large-batch emit kernel (4 blocks per SM hide each other's phases)
struct EmitSmem {
  int row[SB];        // rank of the ray within its case (row of the random tensor)
  int rowv[SB];       // rank among valid rays (row of u_val)
  uint8_t cs[SB];
  float depth[SB];
  float od[SB][6];
  int w[SB / 32][3];
  float nrm[SB * 16];   // the block's rows of the normal draws (case-1 rays, rows are consecutive in ray order), n_bins <= 16
};
__device__ __forceinline__ void
emit_block(EmitSmem& sh, int k, int b, int R, int N, int M, const uint8_t* __restrict__ rgbs, const float* __restrict__ depth,
           const float* __restrict__ origins, const float* __restrict__ dirs,
           const float* __restrict__ u_inv, const float* __restrict__ u_val,
           const float* __restrict__ n_obj, const float* __restrict__ u_oth,
           const float* __restrict__ lin_S, const float* __restrict__ lin_N, const float* __restrict__ lin_M,
           float min_bound, float eps, float stop_eps, int this_obj, int base0_in, int base1_in, int base2_in, float max_bound,
           float* __restrict__ z, float* __restrict__ pcs, uint8_t* __restrict__ valid, uint8_t* __restrict__ labels,
           float* __restrict__ zbuf /* shared, [SB * S]: the block's depths, staged so that z AND pcs leave as linear coalesced streams */) {
  const int S = N + M;
  const int r0 = b * SB;
  int (&s_row)[SB] = sh.row; int (&s_rowv)[SB] = sh.rowv; uint8_t (&s_case)[SB] = sh.cs; float (&s_depth)[SB] = sh.depth;
  float (&s_od)[SB][6] = sh.od; int (&s_w)[SB / 32][3] = sh.w;

  // ---- phase 1: thread per ray (threads 0..SB-1): case + exclusive rank in ray order
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  if (t < SB) {           // warp-uniform (SB is a multiple of 32)
    const int r = r0 + t;
    int c = 3;
    float d = 0.f;
    if (r < R) {
      const size_t gi = (size_t)k * R + r;
      d = depth[gi];
      const uint8_t lab = rgbs[gi * 4 + 3];
      c = ray_case(d, lab, min_bound, this_obj);
      valid[gi] = (c != 0);
      labels[gi] = lab;
#pragma unroll
      for (int j = 0; j < 3; ++j) { s_od[t][j] = origins[gi * 3 + j]; s_od[t][3 + j] = dirs[gi * 3 + j]; }
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, c == 0);
    const unsigned m1 = __ballot_sync(0xffffffffu, c == 1);
    const unsigned m2 = __ballot_sync(0xffffffffu, c == 2);
    const unsigned lt = (1u << l) - 1u;
    if (l == 0) { s_w[w][0] = __popc(m0); s_w[w][1] = __popc(m1); s_w[w][2] = __popc(m2); }
    s_case[t] = (uint8_t)c;
    s_depth[t] = d;
    s_row[t] = (c == 0) ? __popc(m0 & lt) : (c == 1 ? __popc(m1 & lt) : __popc(m2 & lt));     // rank inside the warp
    s_rowv[t] = __popc(m1 & lt) + __popc(m2 & lt);
  }
  __syncthreads();
  if (t < SB) {           // add the block base and the counts of the preceding warps
    int base0 = base0_in, base1 = base1_in, base2 = base2_in;
    for (int i = 0; i < w; ++i) { base0 += s_w[i][0]; base1 += s_w[i][1]; base2 += s_w[i][2]; }
    const int c = s_case[t];
    s_row[t] += (c == 0) ? base0 : (c == 1 ? base1 : base2);
    s_rowv[t] += base1 + base2;     // valid rays = case1 U case2, both counted in ray order
  }
  // stage the normal draws of the block's this-object rays (every sample ranks itself against the ray's M draws)
  const bool stage = M <= 16;
  if (stage) {
    int cnt1 = 0;
#pragma unroll
    for (int i = 0; i < SB / 32; ++i) cnt1 += s_w[i][1];
    const float* src = n_obj + ((size_t)k * R + base1_in) * M;
    for (int e = t; e < cnt1 * M; e += blockDim.x) sh.nrm[e] = src[e];
  }
  __syncthreads();

  // ---- phase 2: thread per sample (coalesced z / pcs writes)
  const int nrays = min(SB, R - r0);
  const float fS = (float)S, fN = (float)N, fM = (float)M;
  const unsigned invS = 0xFFFFFFFFu / (unsigned)S + 1u;      // e / S by multiplication (exact for e < SB * 32; S = 1 wraps to 0: see div_S)
  auto div_S = [&](int e) { return S > 1 ? (int)__umulhi((unsigned)e, invS) : e; }; 
  for (int e = t; e < nrays * S; e += blockDim.x) {
    const int rr = div_S(e), i = e - rr * S;
    const int cc = s_case[rr];
    const float dd = s_depth[rr];
    float zz;
    if (cc == 0) {                               // :484-493 invalid depth: S bins in [min_bound, max depth]
      const float u = u_inv[((size_t)k * R + s_row[rr]) * S + i];
      zz = strat(min_bound, max_bound, lin_S[i], u, fS);
    } else if (i < N) {                          // :502-504 camera -> surface - eps
      const float u = u_val[((size_t)k * R + s_rowv[rr]) * N + i];
      zz = strat(min_bound, __fsub_rn(dd, eps), lin_N[i], u, fN);
    } else if (cc == 1) {                        // :519-524 sorted normals around the surface
      const float* row = stage ? sh.nrm + (s_row[rr] - base1_in) * M : n_obj + ((size_t)k * R + s_row[rr]) * M;
      const int j = i - N;
      const float v = row[j];
      int rank = 0;
      for (int q = 0; q < M; ++q) {
        const float o = row[q];
        rank += (o < v) || (o == v && q < j);
      }
      const float cl = fminf(fmaxf(v, -eps), eps);
      zbuf[rr * S + N + rank] = __fadd_rn(dd, cl);           // the sorted position inside the same ray
      continue;
    } else {                                     // :536-539 other objects: [d-eps, d+stop_eps]
      const float u = u_oth[((size_t)k * R + s_row[rr]) * M + (i - N)];
      zz = strat(__fsub_rn(dd, eps), __fadd_rn(dd, stop_eps), lin_M[i - N], u, fM);
    }
    zbuf[e] = zz;
  }
  __syncthreads();
  // ---- phase 3: the block's z [nrays * S] and pcs [nrays * S * 3] are contiguous in global memory: linear, coalesced stores
  const size_t g0 = ((size_t)k * R + r0) * S;
  const int nz = nrays * S;
  for (int f = t; f < nz; f += blockDim.x) z[g0 + f] = zbuf[f];
  if (pcs) {
    for (int f = t; f < nz * 3; f += blockDim.x) {
      const int smp = (int)(__umulhi((unsigned)f, 0xAAAAAAABu) >> 1), a = f - 3 * smp;
      const int rr = div_S(smp);
      pcs[g0 * 3 + f] = __fadd_rn(s_od[rr][a], __fmul_rn(s_od[rr][3 + a], zbuf[smp]));      // scene_cateogries.py:543-544
    }
  }
}

__global__ void __launch_bounds__(EMIT_T)
sample_emit_kernel(int R, int N, int M, const uint8_t* __restrict__ rgbs, const float* __restrict__ depth,
                   const float* __restrict__ origins, const float* __restrict__ dirs,
                   const float* __restrict__ u_inv, const float* __restrict__ u_val,
                   const float* __restrict__ n_obj, const float* __restrict__ u_oth,
                   const float* __restrict__ lin_S, const float* __restrict__ lin_N, const float* __restrict__ lin_M,
                   float min_bound, float eps, float stop_eps, int this_obj,
                   const int* __restrict__ blk_counts, const int4* __restrict__ cta_prefix, const int32_t* __restrict__ stats,
                   float* __restrict__ z, float* __restrict__ pcs, uint8_t* __restrict__ valid,
                   uint8_t* __restrict__ labels) {
  __shared__ EmitSmem sh;
  extern __shared__ float zbuf_dyn[];
  const int k = blockIdx.y, nb = gridDim.x, b = blockIdx.x;
  const int4 pc = cta_prefix[(size_t)k * ((nb + COUNT_RB - 1) / COUNT_RB) + b / COUNT_RB];
  const int4 pb = reinterpret_cast<const int4*>(blk_counts)[(size_t)k * nb + b];
  emit_block(sh, k, b, R, N, M, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth, lin_S, lin_N, lin_M, min_bound, eps, stop_eps,
             this_obj, pc.x + pb.x, pc.y + pb.y, pc.z + pb.z, __int_as_float(stats[k * 8 + 7]), z, pcs, valid, labels, zbuf_dyn);
}

// (C') emit, fast path for M <= 16 surface bins (every training / rendering shape of the reference except the 96-bin uncertainty
// sweep): ONE THREAD PER RAY.  The generic kernel above spends ~340 instructions per SAMPLE (index arithmetic per element, an
// O(M) rank loop per sample, one IEEE division per sample) and is issue-bound at 21 % of the HBM roofline (ncu: issue slots 79 %
// busy).  Here a block of 256 rays
//   1. loads its ray data (one coalesced round) and ranks the rays inside their case by warp ballots,
//   2. copies its rows of the four random tensors -- rows of one case are consecutive in ray order, so each is ONE contiguous chunk --
//      into shared memory with coalesced loads (nrays * S floats in total, whatever the case mix),
//   3. each thread turns its row into the ray's S depths: the two IEEE divisions of `stratified_bins` once per ray, the M normal
//      draws sorted by a compile-time odd-even transposition network in registers (any correct sort gives the reference's values),
//      and stages z [S] and pcs [S,3] of its ray in shared memory,
//   4. the block's z and pcs leave as linear 16-byte stores.
// Same separately rounded fp32 operations as the reference (bit-exact); ~400 instructions per RAY.
template <int M>
__device__ __forceinline__ void sort_network(float (&v)[M]) {
#pragma unroll
  for (int p = 0; p < M; ++p) {
#pragma unroll
    for (int i = (p & 1); i + 1 < M; i += 2) {
      const float a = v[i], b = v[i + 1];
      v[i] = fminf(a, b); v[i + 1] = fmaxf(a, b);
    }
  }
}

template <int M>
__device__ __forceinline__ void
emit_fast_body(const int k, const int b, int R, int N, const uint8_t* __restrict__ rgbs, const float* __restrict__ depth,
               const float* __restrict__ origins, const float* __restrict__ dirs,
               const float* __restrict__ u_inv, const float* __restrict__ u_val,
               const float* __restrict__ n_obj, const float* __restrict__ u_oth,
               const float* __restrict__ lin_S, const float* __restrict__ lin_N, const float* __restrict__ lin_M,
               float min_bound, float eps, float stop_eps, int this_obj,
               const int base0, const int base1, const int base2, const float max_bound,
               float* __restrict__ z, float* __restrict__ pcs, uint8_t* __restrict__ valid, uint8_t* __restrict__ labels, float* sm_f) {
  __shared__ int s_w[SB / 32][3];
  __shared__ float s_lin[99];                 // linspace tables: [0, 33) S bins | [33, 66) N bins | [66, 99) M bins (S <= 32)
  const int S = N + M;
  const int t = threadIdx.x, w = t >> 5, l = t & 31;
  if (t <= S) s_lin[t] = lin_S[t];
  else if (t >= 64 && t - 64 <= N) s_lin[33 + t - 64] = lin_N[t - 64];
  else if (t >= 128 && t - 128 <= M) s_lin[66 + t - 128] = lin_M[t - 128];
  const int r0 = b * SB, r = r0 + t, nrays = min(SB, R - r0);
  float* rbuf = sm_f;                         // [nrays * S]: random rows, then the block's z
  float* pbuf = sm_f + SB * S;                // [nrays * 3S]: the block's pcs
  // ---- 1. ray data (issued first: the longest-latency loads of the block)
  float d = 0.f, o[3] = {0.f, 0.f, 0.f}, dr[3] = {0.f, 0.f, 0.f};
  uint32_t px = 0;
  const size_t gi = (size_t)k * R + r;
  if (r < R) {
    d = depth[gi];
    px = reinterpret_cast<const uint32_t*>(rgbs)[gi];
#pragma unroll
    for (int j = 0; j < 3; ++j) { o[j] = origins[gi * 3 + j]; dr[j] = dirs[gi * 3 + j]; }
  }
  // case, rank inside the case (ray order)
  int c = 3;
  if (r < R) {
    const uint8_t lab = (uint8_t)(px >> 24);
    c = ray_case(d, lab, min_bound, this_obj);
    valid[gi] = (c != 0);
    labels[gi] = lab;
  }
  const unsigned m0 = __ballot_sync(0xffffffffu, c == 0), m1 = __ballot_sync(0xffffffffu, c == 1), m2 = __ballot_sync(0xffffffffu, c == 2);
  const unsigned lt = (1u << l) - 1u;
  if (l == 0) { s_w[w][0] = __popc(m0); s_w[w][1] = __popc(m1); s_w[w][2] = __popc(m2); }
  __syncthreads();
  int before[3] = {0, 0, 0}, cnt[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < SB / 32; ++i) {
#pragma unroll
    for (int q = 0; q < 3; ++q) { const int v = s_w[i][q]; cnt[q] += v; if (i < w) before[q] += v; }
  }
  // rank of this ray inside the block's rays of its case / of the valid rays
  const int rk0 = before[0] + __popc(m0 & lt), rk1 = before[1] + __popc(m1 & lt), rk2 = before[2] + __popc(m2 & lt);

  // ---- 2. the block's rows of the random tensors: four contiguous chunks -> rbuf (asynchronous 4-byte copies, all in flight)
  //      [0, a0) u_inv rows | [a0, a1) u_val rows of the valid rays | [a1, a2) normal draws | [a2, a3) u_oth rows
  const int nv = cnt[1] + cnt[2];
  const int a0 = cnt[0] * S, a1 = a0 + nv * N, a2 = a1 + cnt[1] * M, a3 = a2 + cnt[2] * M;
  {
    const size_t kr = (size_t)k * R;
    const float* p0 = u_inv + (kr + base0) * S;
    const float* p1 = u_val + (kr + base1 + base2) * N;
    const float* p2 = n_obj + (kr + base1) * M;
    const float* p3 = u_oth + (kr + base2) * M;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(rbuf);
    auto copy = [&](const float* src, int lo, int hi) {
      for (int e = t; e < hi - lo; e += SB)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + 4u * (lo + e)), "l"(src + e) : "memory");
    };
    copy(p0, 0, a0); copy(p1, a0, a1); copy(p2, a1, a2); copy(p3, a2, a3);
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  // ---- 3. this thread's ray, branch-free over the three cases (rays of different cases share a warp): both segments of a ray are
  //      z_i = (range * lin_i + mn) + u_i * (range / n) with per-case {range, mn, n, lin table, row of draws}; the this-object case
  //      replaces its surface segment by depth + clip(sorted normals).  z goes to the first S slots of the thread's pbuf row
  //      (rbuf is still being read by other threads).
  float* mine = pbuf + t * 3 * S;
  {
    const bool inval = (c == 0);
    // segment A = samples [0, N): invalid rays use bins 0..N-1 of their S-bin stratification (:484-493), valid rays camera -> surface - eps (:502-504)
    const float rangeI = __fsub_rn(max_bound, min_bound), lenI = __fdiv_rn(rangeI, (float)S);
    const float rangeV = __fsub_rn(__fsub_rn(d, eps), min_bound), lenV = __fdiv_rn(rangeV, (float)N);
    const float mn2 = __fsub_rn(d, eps), range2 = __fsub_rn(__fadd_rn(d, stop_eps), mn2), len2 = __fdiv_rn(range2, (float)M);
    const float* rowI = rbuf + rk0 * S;                          // only dereferenced by the case that owns it (others read in-bounds junk)
    const float* uA = inval ? rowI : rbuf + a0 + (rk1 + rk2) * N;
    const float* linA = inval ? s_lin : s_lin + 33;
    const float rangeA = inval ? rangeI : rangeV, lenA = inval ? lenI : lenV;
    for (int i = 0; i < N; ++i) mine[i] = __fadd_rn(__fadd_rn(__fmul_rn(rangeA, linA[i]), min_bound), __fmul_rn(uA[i], lenA));
    // segment B = samples [N, S): invalid: bins N..S-1; this object: sorted normals (:519-524); other objects: [d - eps, d + stop_eps] (:536-539)
    const float* uB = inval ? rowI + N : (c == 1 ? rbuf + a1 + rk1 * M : rbuf + a2 + rk2 * M);
    const float* linB = inval ? s_lin + N : s_lin + 66;
    const float rangeB = inval ? rangeI : range2, lenB = inval ? lenI : len2, mnB = inval ? min_bound : mn2;
    float v[M];
    if (c == 3) uB = rbuf;                                       // rays beyond R: any valid address
#pragma unroll
    for (int j = 0; j < M; ++j) v[j] = uB[j];
    float sv[M];
#pragma unroll
    for (int j = 0; j < M; ++j) sv[j] = v[j];
    sort_network<M>(sv);
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const float strat_z = __fadd_rn(__fadd_rn(__fmul_rn(rangeB, linB[j]), mnB), __fmul_rn(v[j], lenB));
      const float norm_z = __fadd_rn(d, fminf(fmaxf(sv[j], -eps), eps));
      mine[N + j] = (c == 1) ? norm_z : strat_z;
    }
  }
  __syncthreads();
  if (c != 3) {
    // in place, last sample first: pcs of sample i go to slots 3i..3i+2, the unread depths sit in slots < i
    for (int i = S - 1; i >= 0; --i) {
      const float zz = mine[i];
      rbuf[t * S + i] = zz;
      if (pcs) {
#pragma unroll
        for (int q = 0; q < 3; ++q) mine[3 * i + q] = __fadd_rn(o[q], __fmul_rn(dr[q], zz));      // scene_cateogries.py:543-544
      }
    }
  }
  __syncthreads();

  // ---- 4. linear copy-out (16-byte stores when the block's output offset allows)
  const size_t g0 = ((size_t)k * R + r0) * S;
  const int nz = nrays * S;
  float* zo = z + g0;
  if ((reinterpret_cast<uintptr_t>(zo) & 15) == 0) {
    for (int f = t; f < (nz >> 2); f += SB) reinterpret_cast<float4*>(zo)[f] = reinterpret_cast<const float4*>(rbuf)[f];
    for (int f = (nz & ~3) + t; f < nz; f += SB) zo[f] = rbuf[f];
  } else {
    for (int f = t; f < nz; f += SB) zo[f] = rbuf[f];
  }
  if (pcs) {
    float* po = pcs + g0 * 3;
    const int np = nz * 3;
    if ((reinterpret_cast<uintptr_t>(po) & 15) == 0) {
      for (int f = t; f < (np >> 2); f += SB) reinterpret_cast<float4*>(po)[f] = reinterpret_cast<const float4*>(pbuf)[f];
      for (int f = (np & ~3) + t; f < np; f += SB) po[f] = pbuf[f];
    } else {
      for (int f = t; f < np; f += SB) po[f] = pbuf[f];
    }
  }
}

template <int M>
__global__ void __launch_bounds__(SB, 5)
sample_emit_fast_kernel(int R, int N, const uint8_t* __restrict__ rgbs, const float* __restrict__ depth,
                        const float* __restrict__ origins, const float* __restrict__ dirs,
                        const float* __restrict__ u_inv, const float* __restrict__ u_val,
                        const float* __restrict__ n_obj, const float* __restrict__ u_oth,
                        const float* __restrict__ lin_S, const float* __restrict__ lin_N, const float* __restrict__ lin_M,
                        float min_bound, float eps, float stop_eps, int this_obj,
                        const int* __restrict__ blk_counts, const int4* __restrict__ cta_prefix, const int32_t* __restrict__ stats,
                        float* __restrict__ z, float* __restrict__ pcs, uint8_t* __restrict__ valid, uint8_t* __restrict__ labels) {
  extern __shared__ __align__(16) float sm_f[];
  const int k = blockIdx.y, nb = gridDim.x, b = blockIdx.x;
  // running case counts before this block = prefix over the count kernel's CTAs + prefix inside its CTA; the category's maximum depth
  const int4 pc = cta_prefix[(size_t)k * ((nb + COUNT_RB - 1) / COUNT_RB) + b / COUNT_RB];
  const int4 pb = reinterpret_cast<const int4*>(blk_counts)[(size_t)k * nb + b];
  emit_fast_body<M>(k, b, R, N, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth, lin_S, lin_N, lin_M, min_bound, eps, stop_eps, this_obj,
                    pc.x + pb.x, pc.y + pb.y, pc.z + pb.z, __int_as_float(stats[k * 8 + 7]), z, pcs, valid, labels, sm_f);
}

// Single launch for small ray batches (R <= SAMPLE_SMALL_RAYS), thread-per-ray: grid (ray block, category).  Every CTA counts the
// cases / loss masks / maximum depth of its whole category (<= 16 rays per thread, all loads in flight), derives the running case
// counts in front of ITS block and emits that block with the fast body; block 0 writes the stats row.
constexpr int SAMPLE_SMALL_RAYS = 4096;
template <int M>
__global__ void __launch_bounds__(SB, 5)
sample_small_fast_kernel(int R, int N, const uint8_t* __restrict__ rgbs, const float* __restrict__ depth,
                         const float* __restrict__ origins, const float* __restrict__ dirs,
                         const float* __restrict__ u_inv, const float* __restrict__ u_val,
                         const float* __restrict__ n_obj, const float* __restrict__ u_oth,
                         const float* __restrict__ lin_S, const float* __restrict__ lin_N, const float* __restrict__ lin_M,
                         float min_bound, float eps, float stop_eps, int this_obj, const float* __restrict__ max_override,
                         int32_t* __restrict__ stats, float* __restrict__ z, float* __restrict__ pcs, uint8_t* __restrict__ valid,
                         uint8_t* __restrict__ labels) {
  extern __shared__ __align__(16) float sm_f[];
  constexpr int NBMAX = SAMPLE_SMALL_RAYS / SB;
  __shared__ int s_blk[NBMAX][3];                     // per ray block: number of rays in case 0 / 1 / 2
  __shared__ int s_cnt[3];                            // loss-mask counts Nd, No, Ns
  __shared__ float s_mx[SB / 32];
  const int k = blockIdx.y, b = blockIdx.x, t = threadIdx.x, w = t >> 5, l = t & 31;
  const int nb = (R + SB - 1) / SB;
  if (t < 3 * NBMAX) (&s_blk[0][0])[t] = 0;
  if (t < 3) s_cnt[t] = 0;
  __syncthreads();
  const uint32_t* px = reinterpret_cast<const uint32_t*>(rgbs);
  float dd[NBMAX];
  uint32_t pp[NBMAX];
#pragma unroll
  for (int j = 0; j < NBMAX; ++j) {
    const int r = j * SB + t;
    const bool in = j < nb && r < R;
    dd[j] = in ? depth[(size_t)k * R + r] : 0.f;
    pp[j] = in ? px[(size_t)k * R + r] : 0u;
  }
  int cD = 0, cO = 0, cS = 0;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NBMAX; ++j) {
    if (j < nb) {                                     // warp-uniform
      const int r = j * SB + t;
      const bool in = r < R;
      const uint8_t lab = (uint8_t)(pp[j] >> 24);
      const int c = in ? ray_case(dd[j], lab, min_bound, this_obj) : 3;
      if (in) { cD += (c != 0) && lab != 0; cO += lab != 0; cS += lab != 2; mx = fmaxf(mx, dd[j]); }
      const unsigned m0 = __ballot_sync(0xffffffffu, c == 0), m1 = __ballot_sync(0xffffffffu, c == 1), m2 = __ballot_sync(0xffffffffu, c == 2);
      if (l == 0) { atomicAdd(&s_blk[j][0], __popc(m0)); atomicAdd(&s_blk[j][1], __popc(m1)); atomicAdd(&s_blk[j][2], __popc(m2)); }
    }
  }
  cD = warp_sum_i(cD); cO = warp_sum_i(cO); cS = warp_sum_i(cS); mx = warp_max(mx);
  if (l == 0) { atomicAdd(&s_cnt[0], cD); atomicAdd(&s_cnt[1], cO); atomicAdd(&s_cnt[2], cS); s_mx[w] = mx; }
  __syncthreads();
  int base[3] = {0, 0, 0}, tot[3] = {0, 0, 0};
  for (int j = 0; j < nb; ++j) {
#pragma unroll
    for (int q = 0; q < 3; ++q) { const int v = s_blk[j][q]; if (j < b) base[q] += v; tot[q] += v; }
  }
  float m = s_mx[0];
#pragma unroll
  for (int i = 1; i < SB / 32; ++i) m = fmaxf(m, s_mx[i]);
  const float max_bound = max_override ? max_override[k] : m;
  if (b == 0 && t == 0) {
    int32_t* st = stats + k * 8;
    st[0] = tot[0]; st[1] = tot[1] + tot[2]; st[2] = tot[1]; st[3] = tot[2];
    st[4] = s_cnt[0]; st[5] = s_cnt[1]; st[6] = s_cnt[2]; st[7] = __float_as_int(max_bound);
  }
  emit_fast_body<M>(k, b, R, N, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth, lin_S, lin_N, lin_M, min_bound, eps, stop_eps, this_obj,
                    base[0], base[1], base[2], max_bound, z, pcs, valid, labels, sm_f);
}

typedef void (*small_fast_fn)(int, int, const uint8_t*, const float*, const float*, const float*, const float*, const float*, const float*,
                              const float*, const float*, const float*, const float*, float, float, float, int, const float*, int32_t*,
                              float*, float*, uint8_t*, uint8_t*);
typedef void (*emit_fast_fn)(int, int, const uint8_t*, const float*, const float*, const float*, const float*, const float*, const float*,
                             const float*, const float*, const float*, const float*, float, float, float, int, const int*, const int4*,
                             const int32_t*, float*, float*, uint8_t*, uint8_t*);
template <int M> struct EmitFastTable {
  static void fill(emit_fast_fn* tab, small_fast_fn* tab2) {
    tab[M] = sample_emit_fast_kernel<M>; tab2[M] = sample_small_fast_kernel<M>;
    EmitFastTable<M - 1>::fill(tab, tab2);
  }
};
template <> struct EmitFastTable<0> { static void fill(emit_fast_fn*, small_fast_fn*) {} };
constexpr int EMIT_FAST_MAX_M = 16;

// Single-launch sampler for small ray batches (R <= SAMPLE_SMALL_RAYS): grid (category, ray block).  Every CTA counts the cases /
// loss masks and the maximum depth of its whole category (a few thousand rays: cheaper than a second launch), derives the running
// case offsets of ITS block and emits that block; block 0 writes the stats row.  (Generic path: more than 16 surface bins.)
__global__ void __launch_bounds__(EMIT_T)
sample_small_kernel(int R, int N, int M, const uint8_t* __restrict__ rgbs, const float* __restrict__ depth,
                    const float* __restrict__ origins, const float* __restrict__ dirs,
                    const float* __restrict__ u_inv, const float* __restrict__ u_val,
                    const float* __restrict__ n_obj, const float* __restrict__ u_oth,
                    const float* __restrict__ lin_S, const float* __restrict__ lin_N, const float* __restrict__ lin_M,
                    float min_bound, float eps, float stop_eps, int this_obj, const float* __restrict__ max_override,
                    int32_t* __restrict__ stats, float* __restrict__ z, float* __restrict__ pcs, uint8_t* __restrict__ valid,
                    uint8_t* __restrict__ labels) {
  __shared__ EmitSmem sh;
  extern __shared__ float zbuf_dyn[];
  __shared__ int s_c[EMIT_T / 32][6];
  __shared__ float s_m[EMIT_T / 32];
  __shared__ int s_tot[6];
  __shared__ float s_max;
  __shared__ int s_blk[SAMPLE_SMALL_RAYS / SB][3];     // per ray block: number of rays in case 0 / 1 / 2
  const int k = blockIdx.x, t = threadIdx.x, w = t >> 5, l = t & 31;
  const int nb = (R + SB - 1) / SB;
  if (t < 3 * (SAMPLE_SMALL_RAYS / SB)) (&s_blk[0][0])[t] = 0;
  __syncthreads();
  // ---- counts: thread per ray (R <= 4 * EMIT_T)
  int cD = 0, cObj = 0, cSem = 0;
  float mx = -INFINITY;
  for (int r = t; r < nb * SB; r += EMIT_T) {
    int cI = 0, cO = 0, cT = 0;
    if (r < R) {
      const float d = depth[(size_t)k * R + r];
      const uint8_t lab = rgbs[((size_t)k * R + r) * 4 + 3];
      const int c = ray_case(d, lab, min_bound, this_obj);
      cI = (c == 0); cO = (c == 1); cT = (c == 2);
      const bool mo = lab != 0;
      cD += (c != 0) && mo; cObj += mo; cSem += lab != 2;
      mx = fmaxf(mx, d);
    }
    // a warp covers 32 consecutive rays of ONE block (SB is a multiple of 32): per-block case counts by ballot
    const unsigned m0 = __ballot_sync(0xffffffffu, cI), m1 = __ballot_sync(0xffffffffu, cO), m2 = __ballot_sync(0xffffffffu, cT);
    if (l == 0 && r - l < nb * SB) { const int b = (r - l) / SB; atomicAdd(&s_blk[b][0], __popc(m0)); atomicAdd(&s_blk[b][1], __popc(m1)); atomicAdd(&s_blk[b][2], __popc(m2)); }
  }
  cD = warp_sum_i(cD); cObj = warp_sum_i(cObj); cSem = warp_sum_i(cSem); mx = warp_max(mx);
  if (l == 0) { s_c[w][3] = cD; s_c[w][4] = cObj; s_c[w][5] = cSem; s_m[w] = mx; }
  __syncthreads();
  if (t < 3) {
    int v = 0;
    for (int i = 0; i < EMIT_T / 32; ++i) v += s_c[i][3 + t];
    s_tot[3 + t] = v;
  }
  if (t == 32) {
    float m = s_m[0];
    for (int i = 1; i < EMIT_T / 32; ++i) m = fmaxf(m, s_m[i]);
    s_max = max_override ? max_override[k] : m;
    int a0 = 0, a1 = 0, a2 = 0;                      // exclusive scan of the per-block case counts (in place)
    for (int b = 0; b < nb; ++b) {
      const int v0 = s_blk[b][0], v1 = s_blk[b][1], v2 = s_blk[b][2];
      s_blk[b][0] = a0; s_blk[b][1] = a1; s_blk[b][2] = a2;
      a0 += v0; a1 += v1; a2 += v2;
    }
    s_tot[0] = a0; s_tot[1] = a1; s_tot[2] = a2;
  }
  __syncthreads();
  if (t == 0 && blockIdx.y == 0) {
    int32_t* st = stats + k * 8;
    st[0] = s_tot[0]; st[1] = s_tot[1] + s_tot[2]; st[2] = s_tot[1]; st[3] = s_tot[2];
    st[4] = s_tot[3]; st[5] = s_tot[4]; st[6] = s_tot[5]; st[7] = __float_as_int(s_max);
  }
  // ---- emit this CTA's block
  const int b = blockIdx.y;
  emit_block(sh, k, b, R, N, M, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth, lin_S, lin_N, lin_M, min_bound, eps, stop_eps,
             this_obj, s_blk[b][0], s_blk[b][1], s_blk[b][2], s_max, z, pcs, valid, labels, zbuf_dyn);
}

__global__ void stratified_bins_kernel(int n_rays, int n_bins, const float* __restrict__ mn, int mn_s,
                                       const float* __restrict__ mx, int mx_s, const float* __restrict__ u,
                                       const float* __restrict__ lin, float* __restrict__ z) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n_rays * n_bins) return;
  const int r = (int)(e / n_bins), i = (int)(e - (size_t)r * n_bins);
  z[e] = strat(mn[mn_s ? 0 : r], mx[mx_s ? 0 : r], lin[i], u[e], (float)n_bins);
}

// origin_dirs_O / origin_dirs_W (scene_cateogries.py:24-47): per-ray camera pose -> ray origin and direction in the object
// (invert = 1: T_OC = inv(T_CO), general 4x4 inverse by cofactors) or world (invert = 0) frame.  One thread per ray; each
// thread reads its own 64-byte matrix as four float4 (full sectors).
__device__ __forceinline__ void pose_rows(int invert, const float4 r0, const float4 r1, const float4 r2, const float4 r3, float b[3][4]) {
  if (invert) {
    const float m00 = r0.x, m01 = r0.y, m02 = r0.z, m03 = r0.w, m10 = r1.x, m11 = r1.y, m12 = r1.z, m13 = r1.w;
    const float m20 = r2.x, m21 = r2.y, m22 = r2.z, m23 = r2.w, m30 = r3.x, m31 = r3.y, m32 = r3.z, m33 = r3.w;
    const float s0 = m00 * m11 - m10 * m01, s1 = m00 * m12 - m10 * m02, s2 = m00 * m13 - m10 * m03;
    const float s3 = m01 * m12 - m11 * m02, s4 = m01 * m13 - m11 * m03, s5 = m02 * m13 - m12 * m03;
    const float c5 = m22 * m33 - m32 * m23, c4 = m21 * m33 - m31 * m23, c3 = m21 * m32 - m31 * m22;
    const float c2 = m20 * m33 - m30 * m23, c1 = m20 * m32 - m30 * m22, c0 = m20 * m31 - m30 * m21;
    const float inv = 1.0f / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
    b[0][0] = (m11 * c5 - m12 * c4 + m13 * c3) * inv;  b[0][1] = (-m01 * c5 + m02 * c4 - m03 * c3) * inv;
    b[0][2] = (m31 * s5 - m32 * s4 + m33 * s3) * inv;  b[0][3] = (-m21 * s5 + m22 * s4 - m23 * s3) * inv;
    b[1][0] = (-m10 * c5 + m12 * c2 - m13 * c1) * inv; b[1][1] = (m00 * c5 - m02 * c2 + m03 * c1) * inv;
    b[1][2] = (-m30 * s5 + m32 * s2 - m33 * s1) * inv; b[1][3] = (m20 * s5 - m22 * s2 + m23 * s1) * inv;
    b[2][0] = (m10 * c4 - m11 * c2 + m13 * c0) * inv;  b[2][1] = (-m00 * c4 + m01 * c2 - m03 * c0) * inv;
    b[2][2] = (m30 * s4 - m31 * s2 + m33 * s0) * inv;  b[2][3] = (-m20 * s4 + m21 * s2 - m23 * s0) * inv;
  } else {
    b[0][0] = r0.x; b[0][1] = r0.y; b[0][2] = r0.z; b[0][3] = r0.w;
    b[1][0] = r1.x; b[1][1] = r1.y; b[1][2] = r1.z; b[1][3] = r1.w;
    b[2][0] = r2.x; b[2][1] = r2.y; b[2][2] = r2.z; b[2][3] = r2.w;
  }
}

__global__ void origin_dirs_kernel(int n, int invert, const float4* __restrict__ T, const float* __restrict__ dirs_c,
                                   float* __restrict__ origins, float* __restrict__ dirs_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float b[3][4];
  pose_rows(invert, T[(size_t)r * 4 + 0], T[(size_t)r * 4 + 1], T[(size_t)r * 4 + 2], T[(size_t)r * 4 + 3], b);
  const float dx = dirs_c[(size_t)r * 3 + 0], dy = dirs_c[(size_t)r * 3 + 1], dz = dirs_c[(size_t)r * 3 + 2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dirs_out[(size_t)r * 3 + i] = b[i][0] * dx + b[i][1] * dy + b[i][2] * dz;
    origins[(size_t)r * 3 + i] = b[i][3];
  }
}

// Device-resident ray pools (sceneCategory.get_training_samples, scene_cateogries.py:352-451): one launch cuts this step's
// rays of EVERY category out of the pools and turns camera poses into ray origins / directions.  The pools of all categories
// are concatenated structure-of-arrays (rgbs u8x4 | depth | camera-frame direction | pose index | instance index); the
// reference's epoch reshuffle `pool = pool[randperm]` (:439-449) is kept as a permutation `perm` over pool rows instead of
// moving 100 B per ray, so the k-th category's r-th ray of the step is pool row perm[seg[k] + start[k] + r].  `pose_index`
// NULL = one pose per pool row (t_co_batch_all / t_wc_batch_all, :424-431), else a row of the pose table (the per-frame
// t_wc_batch_dict[frame_batch] of the background category, :373-379).  Same arithmetic as origin_dirs_kernel.
__global__ void pool_gather_kernel(int R, int invert, const int64_t* __restrict__ seg, const int64_t* __restrict__ start,
                                   const int32_t* __restrict__ perm, const uint32_t* __restrict__ rgbs, const float* __restrict__ depth,
                                   const float* __restrict__ dirs_c, const float4* __restrict__ poses, const int32_t* __restrict__ pose_index,
                                   const int64_t* __restrict__ inst, uint32_t* __restrict__ o_rgbs, float* __restrict__ o_depth,
                                   float* __restrict__ o_origins, float* __restrict__ o_dirs, int64_t* __restrict__ o_inst) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (r >= R) return;
  const size_t j = (size_t)perm[seg[k] + start[k] + r];
  const size_t o = (size_t)k * R + r;
  const size_t pj = pose_index ? (size_t)pose_index[j] : j;
  float b[3][4];
  pose_rows(invert, poses[pj * 4 + 0], poses[pj * 4 + 1], poses[pj * 4 + 2], poses[pj * 4 + 3], b);
  const float dx = dirs_c[j * 3 + 0], dy = dirs_c[j * 3 + 1], dz = dirs_c[j * 3 + 2];
  o_rgbs[o] = rgbs[j];
  o_depth[o] = depth[j];
  if (o_inst) o_inst[o] = inst ? inst[j] : 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o_dirs[o * 3 + i] = b[i][0] * dx + b[i][1] * dy + b[i][2] * dz;
    o_origins[o * 3 + i] = b[i][3];
  }
}

// perm'[i] = perm[seg + idx[i]] for one category's segment (the composition that `pool = pool[rand_idx]` performs).
__global__ void pool_compose_kernel(int64_t n, const int32_t* __restrict__ perm_in, const int64_t* __restrict__ idx, int32_t* __restrict__ perm_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) perm_out[i] = perm_in[idx[i]];
}

// Ray split across GPUs: sum the loss-mask counts and take the maximum depth over the ranks' stats rows (all-gathered).
__global__ void reduce_stats_kernel(int world, int K, const int32_t* __restrict__ gathered /* [world][K][8] */,
                                    int32_t* __restrict__ counts /* [K][4] */, float* __restrict__ max_depth /* [K] */) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int c0 = 0, c1 = 0, c2 = 0;
  float m = -INFINITY;
  for (int r = 0; r < world; ++r) {
    const int32_t* s = gathered + ((size_t)r * K + k) * 8;
    c0 += s[4]; c1 += s[5]; c2 += s[6];
    m = fmaxf(m, __int_as_float(s[7]));
  }
  counts[k * 4 + 0] = c0; counts[k * 4 + 1] = c1; counts[k * 4 + 2] = c2; counts[k * 4 + 3] = 0;
  max_depth[k] = m;
}

}  // namespace cnf

using namespace cnf;

extern "C" int cnf_origin_dirs(int n_rays, int invert, const float* T, const float* dirs_c, float* origins, float* dirs_out,
                               cnf_stream_t stream) {
  CNF_REQUIRE(n_rays >= 0 && (n_rays == 0 || (T && dirs_c && origins && dirs_out)), "cnf_origin_dirs: bad arguments");
  CNF_REQUIRE((reinterpret_cast<uintptr_t>(T) & 15) == 0, "cnf_origin_dirs: T must be 16-byte aligned");
  if (n_rays == 0) return CNF_OK;
  origin_dirs_kernel<<<(n_rays + 255) / 256, 256, 0, as_stream(stream)>>>(n_rays, invert, reinterpret_cast<const float4*>(T), dirs_c,
                                                                           origins, dirs_out);
  return check_launch("cnf_origin_dirs");
}

extern "C" int cnf_pool_gather(int n_cat, int n_rays, int invert, const int64_t* seg, const int64_t* start, const int32_t* perm,
                               const uint8_t* rgbs, const float* depth, const float* dirs_c, const float* poses, const int32_t* pose_index,
                               const int64_t* inst, uint8_t* out_rgbs, float* out_depth, float* out_origins, float* out_dirs,
                               int64_t* out_inst, cnf_stream_t stream) {
  CNF_REQUIRE(n_cat >= 0 && n_rays >= 0, "cnf_pool_gather: bad sizes");
  if (n_cat == 0 || n_rays == 0) return CNF_OK;
  CNF_REQUIRE(seg && start && perm && rgbs && depth && dirs_c && poses && out_rgbs && out_depth && out_origins && out_dirs,
              "cnf_pool_gather: null pointer");
  CNF_REQUIRE((reinterpret_cast<uintptr_t>(poses) & 15) == 0 && (reinterpret_cast<uintptr_t>(rgbs) & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(out_rgbs) & 3) == 0,
              "cnf_pool_gather: poses must be 16-byte aligned, rgbs 4-byte aligned");
  CNF_REQUIRE(n_cat <= 65535, "cnf_pool_gather: more than 65535 categories");
  pool_gather_kernel<<<dim3((n_rays + 255) / 256, n_cat), 256, 0, as_stream(stream)>>>(
      n_rays, invert, seg, start, perm, reinterpret_cast<const uint32_t*>(rgbs), depth, dirs_c, reinterpret_cast<const float4*>(poses),
      pose_index, inst, reinterpret_cast<uint32_t*>(out_rgbs), out_depth, out_origins, out_dirs, out_inst);
  return check_launch("cnf_pool_gather");
}

extern "C" int cnf_pool_compose(int64_t n, const int32_t* perm_in, const int64_t* idx, int32_t* perm_out, cnf_stream_t stream) {
  CNF_REQUIRE(n >= 0 && (n == 0 || (perm_in && idx && perm_out && perm_in != perm_out)), "cnf_pool_compose: bad arguments (out of place only)");
  if (n == 0) return CNF_OK;
  pool_compose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(n, perm_in, idx, perm_out);
  return check_launch("cnf_pool_compose");
}

extern "C" size_t cnf_sample_workspace_bytes(int n_cat, int n_rays) {
  const size_t nb = (size_t)(n_rays + SB - 1) / SB;
  const size_t nc = (nb + COUNT_RB - 1) / COUNT_RB;
  return align_up((size_t)n_cat * nb * 4 * sizeof(int), 256) + 3 * align_up((size_t)n_cat * nc * 4 * sizeof(int), 256) +
         align_up((size_t)n_cat * sizeof(unsigned), 256);            // block prefixes | CTA totals lo, hi | CTA prefixes | tickets
}

extern "C" int cnf_sample_stats(int K, int R, const uint8_t* rgbs, const float* depth, float min_bound, int this_obj, int32_t* stats,
                                void* ws, size_t ws_bytes, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && R > 0 && rgbs && depth && stats && ws, "cnf_sample_stats: bad arguments");
  if (ws_bytes < cnf_sample_workspace_bytes(K, R)) { set_error("cnf_sample_stats: workspace too small"); return CNF_ERR_WORKSPACE; }
  const int nb = (R + SB - 1) / SB, nc = (nb + COUNT_RB - 1) / COUNT_RB;
  char* w8 = reinterpret_cast<char*>(ws);
  const size_t cta_bytes = align_up((size_t)K * nc * 4 * sizeof(int), 256);
  int4* cta_lo = reinterpret_cast<int4*>(w8 + align_up((size_t)K * nb * 4 * sizeof(int), 256));
  int4* cta_hi = reinterpret_cast<int4*>(reinterpret_cast<char*>(cta_lo) + cta_bytes);
  int4* cta_prefix = reinterpret_cast<int4*>(reinterpret_cast<char*>(cta_hi) + cta_bytes);
  unsigned* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(cta_prefix) + cta_bytes);
  cudaStream_t st = as_stream(stream);
  CNF_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned) * K, st));
  sample_count_kernel<<<dim3(nc, K), SB, 0, st>>>(R, nb, rgbs, depth, min_bound, this_obj, reinterpret_cast<int*>(w8), cta_lo, cta_hi, cta_prefix,
                                                  ticket, nullptr, stats);
  return check_launch("cnf_sample_stats");
}

extern "C" int cnf_reduce_stats(int world, int K, const int32_t* gathered, int32_t* counts, float* max_depth, cnf_stream_t stream) {
  CNF_REQUIRE(world > 0 && K > 0 && gathered && counts && max_depth, "cnf_reduce_stats: bad arguments");
  reduce_stats_kernel<<<(K + 127) / 128, 128, 0, as_stream(stream)>>>(world, K, gathered, counts, max_depth);
  return check_launch("cnf_reduce_stats");
}

extern "C" int cnf_sample_fwd(int K, int R, int N, int M, const uint8_t* rgbs, const float* depth,
                              const float* origins, const float* dirs, const float* u_inv, const float* u_val,
                              const float* n_obj, const float* u_oth, const float* lin_S, const float* lin_N,
                              const float* lin_M, float min_bound, float eps, float stop_eps, int this_obj,
                              const float* max_depth_override, float* z, float* pcs, uint8_t* valid, uint8_t* labels, int32_t* stats, void* ws,
                              size_t ws_bytes, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && R > 0 && N >= 0 && M > 0, "cnf_sample_fwd: bad sizes K=%d R=%d N=%d M=%d", K, R, N, M);
  CNF_REQUIRE(rgbs && depth && origins && dirs && z && valid && labels && stats && ws, "cnf_sample_fwd: null pointer");
  CNF_REQUIRE(u_inv && u_val && n_obj && u_oth && lin_S && lin_M && (N == 0 || lin_N), "cnf_sample_fwd: null random/linspace input");
  if (ws_bytes < cnf_sample_workspace_bytes(K, R)) { set_error("cnf_sample_fwd: workspace too small"); return CNF_ERR_WORKSPACE; }
  const int nb = (R + SB - 1) / SB;
  const int nc = (nb + COUNT_RB - 1) / COUNT_RB;
  char* w8 = reinterpret_cast<char*>(ws);
  int* blk_counts = reinterpret_cast<int*>(w8);
  const size_t cta_bytes = align_up((size_t)K * nc * 4 * sizeof(int), 256);
  int4* cta_lo = reinterpret_cast<int4*>(w8 + align_up((size_t)K * nb * 4 * sizeof(int), 256));
  int4* cta_hi = reinterpret_cast<int4*>(reinterpret_cast<char*>(cta_lo) + cta_bytes);
  int4* cta_prefix = reinterpret_cast<int4*>(reinterpret_cast<char*>(cta_hi) + cta_bytes);
  unsigned* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(cta_prefix) + cta_bytes);
  cudaStream_t st = as_stream(stream);
  const size_t zbytes = (size_t)SB * (N + M) * sizeof(float);
  CNF_REQUIRE(N + M <= 96, "cnf_sample_fwd: more than 96 samples per ray");
  static DeviceOnce attr;
  if (attr.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(sample_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB * 96 * (int)sizeof(float)));
    CNF_CUDA(cudaFuncSetAttribute(sample_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SB * 96 * (int)sizeof(float)));
    attr.done();
  }
  const bool fast = M <= EMIT_FAST_MAX_M && N + M <= 32;      // thread-per-ray kernels
  static emit_fast_fn tab[EMIT_FAST_MAX_M + 1] = {nullptr};
  static small_fast_fn tab2[EMIT_FAST_MAX_M + 1] = {nullptr};
  const int fbytes = SB * 4 * (N + M) * (int)sizeof(float);
  if (fast) {
    if (!tab[1]) EmitFastTable<EMIT_FAST_MAX_M>::fill(tab, tab2);
    static DeviceOnce fattr[EMIT_FAST_MAX_M + 1];
    if (fattr[M].pending()) {
      for (const void* f : {(const void*)tab[M], (const void*)tab2[M]})
        CNF_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SB * 4 * 32 * (int)sizeof(float)));
      // the large-batch kernel wants five 41 KB CTAs per SM; the small-batch kernel keeps the default carve-out: it runs next to
      // other kernels of the step (prepare, random draws), and kernels with different carve-outs cannot share an SM
      CNF_CUDA(cudaFuncSetAttribute((const void*)tab[M], cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CNF_CUDA(cudaFuncSetAttribute((const void*)tab2[M], cudaFuncAttributePreferredSharedMemoryCarveout, CNF_SMALL_STEP_CARVEOUT));
      fattr[M].done();
    }
  }
  if (R <= SAMPLE_SMALL_RAYS) {
    // Thread-per-ray kernel when the points are wanted (module-level sample_3d_points, background step: 12.3 vs 14.3 us per call,
    // background step 4 % faster).  The z-only call is the fused object trainer's, where the sampler runs NEXT TO the prepare kernel
    // on another stream: there the generic kernel (whole SMs per CTA) leaves the prepare kernel's latency-bound CTAs alone and the
    // step is 6 us faster (same-box A/B: 88.4 vs 94.0 us), although the kernel itself is 2 us slower.
    if (fast && pcs != nullptr)
      tab2[M]<<<dim3(nb, K), SB, fbytes, st>>>(R, N, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth, lin_S, lin_N, lin_M, min_bound, eps,
                                                stop_eps, this_obj, max_depth_override, stats, z, pcs, valid, labels);
    else
      sample_small_kernel<<<dim3(K, nb), EMIT_T, zbytes, st>>>(R, N, M, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth, lin_S, lin_N, lin_M,
                                                               min_bound, eps, stop_eps, this_obj, max_depth_override, stats, z, pcs, valid, labels);
    return check_launch("cnf_sample_fwd");
  }
  CNF_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned) * K, st));
  sample_count_kernel<<<dim3(nc, K), SB, 0, st>>>(R, nb, rgbs, depth, min_bound, this_obj, blk_counts, cta_lo, cta_hi, cta_prefix, ticket,
                                                  max_depth_override, stats);
  if (fast) {
    tab[M]<<<dim3(nb, K), SB, fbytes, st>>>(R, N, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth, lin_S, lin_N, lin_M, min_bound,
                                             eps, stop_eps, this_obj, blk_counts, cta_prefix, stats, z, pcs, valid, labels);
    return check_launch("cnf_sample_fwd");
  }
  sample_emit_kernel<<<dim3(nb, K), EMIT_BIG_T, zbytes, st>>>(R, N, M, rgbs, depth, origins, dirs, u_inv, u_val, n_obj, u_oth,
                                                 lin_S, lin_N, lin_M, min_bound, eps, stop_eps, this_obj,
                                                 blk_counts, cta_prefix, stats, z, pcs, valid, labels);
  return check_launch("cnf_sample_fwd");
}

extern "C" int cnf_stratified_bins(int n_rays, int n_bins, const float* min_d, int min_is_scalar, const float* max_d,
                                   int max_is_scalar, const float* u, const float* lin, float* z, cnf_stream_t stream) {
  CNF_REQUIRE(n_rays >= 0 && n_bins > 0 && min_d && max_d && u && lin && z, "cnf_stratified_bins: bad arguments");
  if (n_rays == 0) return CNF_OK;
  const size_t n = (size_t)n_rays * n_bins;
  stratified_bins_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(n_rays, n_bins, min_d, min_is_scalar,
                                                                                     max_d, max_is_scalar, u, lin, z);
  return check_launch("cnf_stratified_bins");
}
