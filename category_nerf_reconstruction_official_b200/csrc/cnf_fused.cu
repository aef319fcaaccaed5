// Fused training step on Blackwell tensor cores (tcgen05 / TMEM), bf16 operands, fp32 accumulation.
//
//   embed -> CodeNeRF (W=32) -> occupancy compositing -> depth/colour/opacity losses, forward AND backward,
//   for K categories in ONE launch (reference train.py:136-182 minus the optimizer; math: src/embedding.py:82-92,
//   src/model.py:56-84, src/render_rays.py:25-95, src/loss.py:18-74).
//
// Rays are first grouped by instance (counting sort), so a tile of 128 sample points (whole rays, warp-aligned) belongs to
// ONE instance: the latent-code terms of the conditioned layers fold into a per-tile uniform bias c_l[inst] = W_l z_l[inst] + b_l
// and their gradients into one accumulator row (a constant-one column next to the activation tile).
// One CTA = 576 threads = two epilogue roles + two MMA issuers, pipelined over tiles:
//   * forward role (8 warps, two threads per tile row): embedding in registers (one MUFU sin/cos per direction +
//     double-angle recurrence) written as bf16 straight into TENSOR MEMORY as the A operand, then per layer
//     tcgen05.ld -> bias + relu -> bf16 -> next A operand (TMEM) + activation tile (shared memory, kept for the backward role);
//     rgb head, warp-shuffle compositing, losses and loss gradients close the forward of a tile;
//   * backward role (8 warps) works one tile behind: per layer tcgen05.ld of the data gradient -> relu mask (re-read from
//     the activation tile) -> bf16 -> A operand (TMEM) + gradient tile (shared memory); embedding backward at the end;
//   * forward / data-gradient GEMMs take A from tensor memory (16 cycles per 128x32x16 step instead of 40 from shared
//     memory); weight-gradient GEMMs contract over the tile's 128 points with both operands MN-major in shared memory,
//     as M=64 tiles wherever a layer has <= 64 inputs, two such accumulators sharing TMEM columns (lane offset 16);
//   * weight gradients accumulate in TMEM across ALL tiles of the CTA and are flushed once at the end; the per-instance
//     bias-gradient rows are flushed incrementally when the CTA's tile range crosses an instance boundary;
//   * encoding_shape (linear, no activation) is folded algebraically into its two linear consumers
//     (W6' = Wvd[:, :32] W5, wsig' = wsig W5); finalize un-folds the gradients by the chain rule;
//   * biases of layers 1 / 6 / sigma / rgb ride in constant-one columns (free in the GEMM; the matching weight-gradient
//     row is the bias gradient); the PE projection gradient is one more M=64 weight-gradient tile ([t,1]^T d(proj)).
// A prep kernel builds the bf16 operand images and per-instance biases; a finalize kernel maps the raw TMEM
// accumulators back to the reference's nn.Linear-shaped gradients (incl. latent layers, code tables, regulariser).
#include "cnf_common.cuh"
#include "cnf_umma.cuh"
#include "cnf_sample_math.cuh"

namespace cnf {
using namespace umma;

#include "cnf_fused_a.inc"
#include "cnf_fused_b.inc"
#include "cnf_fused_c.inc"

// ---------------------------------------------------------------------------------------------------------
// finalize: raw accumulators -> reference-shaped gradients (one CTA per category) + code regulariser
// ---------------------------------------------------------------------------------------------------------
// out_loss = {loss, e_depth, e_obj, e_sem, depth[K], color[K], opacity[K], reg_shape[K], reg_tex[K]}; executed by one warp
__device__ __forceinline__ void loss_outputs(int K, const float* __restrict__ raw, const int32_t* __restrict__ counts, int cstride,
                                             const float* reg, float cs, float os, float reg_scaling, int add_reg, int secondary,
                                             float* __restrict__ out_loss) {
  const int lane = threadIdx.x & 31;
  int f0 = 0, f1 = 0, f2 = 0;
  for (int k = lane; k < K; k += 32) { f0 |= counts[k * cstride + 0] == 0; f1 |= counts[k * cstride + 1] == 0; f2 |= counts[k * cstride + 2] == 0; }
  const bool e0 = __any_sync(0xffffffffu, f0), e1 = __any_sync(0xffffffffu, f1), e2 = __any_sync(0xffffffffu, f2);
  float total = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float* l = raw + (size_t)k * RAW_FLOATS + RLOSS;
    const float r0 = __ldcg(reg + k * 2), r1 = __ldcg(reg + k * 2 + 1);     // written by other CTAs of this launch: bypass L1
    const float d = e0 ? 0.f : l[0] / ((float)counts[k * cstride + 0] + 1e-10f);
    const float c = e1 ? 0.f : l[1] / ((float)counts[k * cstride + 1] + 1e-10f);
    const float o = e2 ? 0.f : l[2] / ((float)counts[k * cstride + 2] + 1e-10f);
    out_loss[4 + k] = d; out_loss[4 + K + k] = c; out_loss[4 + 2 * K + k] = o;
    out_loss[4 + 3 * K + k] = secondary ? 0.f : r0; out_loss[4 + 4 * K + k] = secondary ? 0.f : r1;
    total += d + c * cs + o * os;
    if (add_reg) total += reg_scaling * (r0 + r1);
  }
  total = warp_sum(total);
  // secondary part of a ray-split step (outputs are summed over the parts): flags and regulariser values are reported by the first part only
  if (lane == 0) { out_loss[0] = total; out_loss[1] = secondary ? 0.f : (float)e0; out_loss[2] = secondary ? 0.f : (float)e1; out_loss[3] = secondary ? 0.f : (float)e2; }
}

// grid = (K, FIN_JOBS): one CTA per (category, job); a job is a SLICE of ONE family of outputs, so a CTA's chain of dependent
// global-load rounds is short (the kernel is latency-bound: ~6 M instructions in total).  kind / slot / part of a job:
//   0 encoding_xyz, rgb layers, PE matrix   1 encoding_viewdir (sliced)   2 encoding_shape, its bias, sigma weights
//   3 conditioned layer s (cat_layer sliced)   4 latent layer s (sliced; needs d z of slot s)
//   5 code tables (4 parts; need d z of all slots + code norms)   6 regulariser values + loss outputs
struct FinJob { int kind, s, part, nparts; };
constexpr int FIN_P_VD = 3, FIN_P_CAT = 3, FIN_P_LAT = 3, FIN_P_CODE = 4;     // parts of the sliced jobs
constexpr int FIN_JOBS = 1 + FIN_P_VD + 1 + (3 + FIN_P_CAT) + 4 * FIN_P_LAT + FIN_P_CODE + 1;     // 28: K = 20 categories fit one wave
// pm = part multiplier: with few categories the sliced jobs are cut into pm x more parts so that the grid still fills the machine
// (K = 1: 182 CTAs instead of 28); grid.y = fin_jobs(pm) = 6 + 22 pm
__host__ __device__ constexpr int fin_jobs(int pm) { return 1 + FIN_P_VD * pm + 1 + (3 + FIN_P_CAT * pm) + 4 * FIN_P_LAT * pm + FIN_P_CODE * pm + 1; }
static_assert(fin_jobs(1) == FIN_JOBS, "job table");
__device__ __forceinline__ FinJob fin_job(int j, int pm) {
  if (j == 0) return {0, 0, 0, 1};
  j -= 1;
  if (j < FIN_P_VD * pm) return {1, 0, j, FIN_P_VD * pm};
  j -= FIN_P_VD * pm;
  if (j == 0) return {2, 0, 0, 1};
  j -= 1;
  if (j == 0) return {3, 0, 0, 1};
  j -= 1;
  if (j < FIN_P_CAT * pm) return {3, 1, j, FIN_P_CAT * pm};
  j -= FIN_P_CAT * pm;
  if (j < 2) return {3, 2 + j, 0, 1};
  j -= 2;
  if (j < 4 * FIN_P_LAT * pm) return {4, j / (FIN_P_LAT * pm), j % (FIN_P_LAT * pm), FIN_P_LAT * pm};
  j -= 4 * FIN_P_LAT * pm;
  if (j < FIN_P_CODE * pm) return {5, 0, j, FIN_P_CODE * pm};
  return {6, 0, 0, 1};
}
// (256, 4): 4 CTAs per SM so that the K x FIN_JOBS grid (560 CTAs at K = 20) is ONE wave; at 108 registers only 2 CTAs per SM fit
// and the second half of the jobs started 5-11 us late
__global__ void __launch_bounds__(256, 4)
fused_finalize_kernel(int L, int n_obj_max, const float* __restrict__ params, const float* __restrict__ pe_scale,
                      const float* __restrict__ shape_codes, const float* __restrict__ tex_codes,
                      const int32_t* __restrict__ n_obj_arr, const float* __restrict__ raw, const float* __restrict__ rawi,
                      const float* __restrict__ zlat,
                      float reg_scaling, int add_reg, int secondary, float* __restrict__ g_params, float* __restrict__ g_B,
                      float* __restrict__ g_shape, float* __restrict__ g_tex, float* __restrict__ reg_out,
                      const int32_t* __restrict__ counts, int cstride, float cs, float os, int* __restrict__ done_counter,
                      float* __restrict__ out_loss, long long* __restrict__ dbg /* development aid: per-CTA %globaltimer {start, end} */) {
  const int k = blockIdx.x, tid = threadIdx.x;
  struct Stamp {
    long long* p;
    __device__ Stamp(long long* q) : p(q) { if (p && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p[0] = t; } }
    __device__ ~Stamp() { __syncthreads(); if (p && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p[1] = t; } }
  } stamp_(dbg ? dbg + 16384 + 2 * (blockIdx.x * gridDim.y + blockIdx.y) : nullptr);
  const FinJob jb = fin_job(blockIdx.y, ((int)gridDim.y - 6) / 22);
  auto part_lo = [&](int n) { return (int)((long long)n * jb.part / jb.nparts); };
  auto part_hi = [&](int n) { return (int)((long long)n * (jb.part + 1) / jb.nparts); };
  // parameter offsets: built once per CTA in shared memory (dynamically indexed below; a per-thread table lived on the stack and
  // its construction was 12 % of the kernel's instructions)
  __shared__ __align__(16) unsigned char s_lay_raw[sizeof(PLay)];
  if (tid == 0) { const PLay tmp(L); *reinterpret_cast<PLay*>(s_lay_raw) = tmp; }
  __syncthreads();
  const PLay& lay = *reinterpret_cast<const PLay*>(s_lay_raw);
  const unsigned invL = 0xFFFFFFFFu / (unsigned)L + 1u;      // i / L by multiplication (exact for i < 2^32 / (4 L): every index range below)
  const float* P = params + (size_t)k * lay.total;
  float* G = g_params + (size_t)k * lay.total;
  const float* rk = raw + (size_t)k * RAW_FLOATS;
  const float* z = zlat + (size_t)k * 4 * n_obj_max * NL;            // [4][n_obj_max][32]
  const float* ri = rawi + (size_t)k * n_obj_max * 4 * 32;            // [n_obj_max][4][32] per-instance bias gradients
  const int n_obj = min(n_obj_arr[k], n_obj_max);
  __shared__ float s_dpre[4][MAXOBJ][NL];      // relu-masked gradient wrt the latent pre-activations, one chunk of MAXOBJ instances
  __shared__ float s_norm[2][MAXOBJ];
  (void)pe_scale;                              // t = x/scale is already applied in the accumulated operand
  const int warp = tid >> 5, lane = tid & 31;
  // conditioned layers: parameter index of the consumer weights, their row length, raw region; latent layer parameter index
  auto cons_w = [](int s) { return s == 0 ? 4 : (s == 1 ? 10 : (s == 2 ? 8 : 20)); };
  auto cons_ld = [](int s) { return s == 1 ? 119 : 32; };
  auto rw = [](int s) { return s == 0 ? RW2 : (s == 1 ? RW3 : (s == 2 ? RW4 : RW7)); };
  auto lat_w = [](int s) { return s == 0 ? 2 : (s == 1 ? 12 : (s == 2 ? 6 : 22)); };

  // d z[inst][c] = sum_n dC[inst][n] W_cons[n][c], masked by relu(z), for slots [s0, s1) and instances [i0, i0 + nc)
  auto compute_dpre = [&](int s0, int s1, int i0, int nc) {
    for (int i = tid; i < (s1 - s0) * nc * NL; i += 256) {
      const int s = s0 + i / (nc * NL), li = (i / NL) % nc, inst = i0 + li, c = i % NL;
      float a = 0.f;
      if (z[((size_t)s * n_obj_max + inst) * NL + c] > 0.f) {
        const int ld = cons_ld(s);
        const float* w = P + lay.o[cons_w(s)] + c;
        const float* r = ri + (inst * 4 + s) * 32;
#pragma unroll
        for (int n = 0; n < 32; ++n) a = fmaf(r[n], w[(size_t)n * ld], a);
      }
      s_dpre[s][li][c] = a;
    }
  };
  auto code_norm = [&](int which, int inst) {           // whole warp; every lane gets |code|
    const float* code = (which ? tex_codes : shape_codes) + ((size_t)k * n_obj_max + inst) * L;
    float a = 0.f;
    for (int j = lane; j < L; j += 32) a = fmaf(code[j], code[j], a);
    return sqrtf(warp_sum(a));
  };
  auto compute_norms = [&](int i0, int nc) {
    for (int o = warp; o < 2 * nc; o += 8) {
      const float a = code_norm(o / nc, i0 + o % nc);
      if (lane == 0) s_norm[o / nc][o % nc] = a;
    }
  };

  if (jb.kind == 0) {
    // PE projection matrix: dB[d][a] = sum_p d(proj)[p][d] * t[p][a]  (accumulated as [t0 t1 t2 ONE]^T d(proj))
    for (int i = tid; i < 63; i += 256) { const int d = i / 3, ax = i % 3; g_B[(size_t)k * 63 + i] = rk[RPE + ax * 32 + pe_col(d)]; }
    // plain layers (transposed copies out of the raw accumulators; xg / xc lanes follow the tile column order).  Everywhere in this
    // kernel consecutive threads READ consecutive addresses of the raw accumulators / parameters (or one broadcast address) and the
    // strided side is the fire-and-forget store: the strided reads of the first version made the kernel L2-sector bound.
    for (int i = tid; i < 32 * 87; i += 256) { const int e = i >> 5, n = i & 31; G[lay.o[0] + n * 87 + e] = rk[RW1 + xg_col(e) * 32 + n]; }
    for (int i = tid; i < 16 * 32; i += 256) { const int c = i >> 4, n = i & 15; G[lay.o[24] + n * 32 + c] = rk[RW89 + c * 16 + n]; }
    for (int i = tid; i < 3 * 16; i += 256) { const int j = i / 16, c = i % 16; G[lay.o[26] + i] = rk[RW89 + (32 + c) * 16 + j]; }
    for (int j = tid; j < 32; j += 256) G[lay.o[1] + j] = rk[RW1 + XG_BIAS_COL * 32 + j];              // b_xyz (row of the ones column)
    for (int n = tid; n < 16; n += 256) G[lay.o[25] + n] = rk[RDB8 + n];
    for (int j = tid; j < 3; j += 256) G[lay.o[27] + j] = rk[RDB9 + j];
  } else if (jb.kind == 1 || jb.kind == 2) {
    // encoding_viewdir / encoding_shape / sigma: the kernel trained the folded weights
    //   W6a' = Wvd[:, :32] W5, b6' = Wvd[:, :32] b5 + b_vd, wsig' = wsig W5, bsig' = wsig b5 + bsig.
    // With Ga'[n][k] = dL/dW6a'[n][k], gb6'[n], gws'[k], gbs':
    //   dWvd[n][j<32] = sum_k Ga'[n][k] W5[j][k] + gb6'[n] b5[j]         dWvd[n][32+e] direct
    //   dW5[j][k]     = sum_n Wvd[n][j] Ga'[n][k] + wsig[j] gws'[k]      db5[j] = sum_n Wvd[n][j] gb6'[n] + wsig[j] gbs'
    //   dwsig[j]      = sum_k gws'[k] W5[j][k] + gbs' b5[j]              db_vd = gb6', dbsig = gbs'
    const int BL = 32 + XC_BIAS_COL;                                   // lane of the ones column in the [r4 | xc] tile
    const float gbs = rk[RW6 + BL * 48 + 32];
    if (jb.kind == 1) {
      for (int i = part_lo(32 * 74) + tid; i < part_hi(32 * 74); i += 256) {
        const int j = i >> 5, n = i & 31;
        float a;
        if (j < 32) {
          a = rk[RW6 + BL * 48 + n] * P[lay.o[15] + j];
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) a = fmaf(rk[RW6 + kk * 48 + n], P[lay.o[14] + j * 32 + kk], a);
        } else {
          a = rk[RW6 + (32 + xc_col(87 + (j - 32))) * 48 + n];
        }
        G[lay.o[18] + n * 74 + j] = a;
      }
      if (jb.part == 0) {
        for (int j = tid; j < 32; j += 256) G[lay.o[19] + j] = rk[RW6 + BL * 48 + j];                    // b_vd
        if (tid == 0) G[lay.o[17]] = gbs;                                                                 // b_sigma
      }
    } else {
      for (int i = tid; i < 32 * 32; i += 256) {                      // dW5[j][k]
        const int kk = i >> 5, j = i & 31;
        float a = P[lay.o[16] + j] * rk[RW6 + kk * 48 + 32];
#pragma unroll
        for (int n = 0; n < 32; ++n) a = fmaf(P[lay.o[18] + (size_t)n * 74 + j], rk[RW6 + kk * 48 + n], a);
        G[lay.o[14] + j * 32 + kk] = a;
      }
      for (int j = tid; j < 32; j += 256) {
        float b5 = P[lay.o[16] + j] * gbs;                             // db5[j]
#pragma unroll
        for (int n = 0; n < 32; ++n) b5 = fmaf(P[lay.o[18] + (size_t)n * 74 + j], rk[RW6 + BL * 48 + n], b5);
        G[lay.o[15] + j] = b5;
        float ws = gbs * P[lay.o[15] + j];                             // dwsig[j]
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) ws = fmaf(rk[RW6 + kk * 48 + 32], P[lay.o[14] + j * 32 + kk], ws);
        G[lay.o[16] + j] = ws;
      }
    }
  } else if (jb.kind == 3) {
    // layer with per-instance bias c = W[:, :32] z + b : slots (zsl1->s1), (zcl->cat), (zsl2->s2), (ztl->t1)
    const int s = jb.s, ld = cons_ld(s);
    for (int i = part_lo(32 * ld) + tid; i < part_hi(32 * ld); i += 256) {
      const int c = i >> 5, n = i & 31;
      float a = rk[rw(s) + (c < 32 ? c : 32 + xg_col(c - 32)) * 32 + n];     // only the cat layer (s = 1) has c >= 32
      if (c < 32)
        for (int inst = 0; inst < n_obj; ++inst) a = fmaf(ri[(inst * 4 + s) * 32 + n], z[((size_t)s * n_obj_max + inst) * NL + c], a);
      G[lay.o[cons_w(s)] + n * ld + c] = a;
    }
    if (jb.part == 0)
      for (int n = tid; n < 32; n += 256) {
        float a = 0.f;
        for (int inst = 0; inst < n_obj; ++inst) a += ri[(inst * 4 + s) * 32 + n];
        G[lay.o[cons_w(s) + 1] + n] = a;
      }
  } else if (jb.kind == 4) {
    // latent layer: dW_lat[c][j] = sum_inst dpre[inst][c] code[inst][j]; db_lat[c] = sum_inst dpre[inst][c]
    const int s = jb.s;
    const float* codes = (s == 3 ? tex_codes : shape_codes) + (size_t)k * n_obj_max * L;
    for (int i0 = 0; i0 == 0 || i0 < n_obj; i0 += MAXOBJ) {         // chunks of MAXOBJ instances; later chunks add to the first
      const int nc = max(0, min(MAXOBJ, n_obj - i0));
      if (i0) __syncthreads();
      compute_dpre(s, s + 1, i0, nc);
      __syncthreads();
      for (int i = part_lo(32 * L) + tid; i < part_hi(32 * L); i += 256) {
        const int c = (int)__umulhi((unsigned)i, invL), j = i - c * L;
        float a = i0 ? G[lay.o[lat_w(s)] + i] : 0.f;
        const float* dp = &s_dpre[s][0][c];
        const float* cd = codes + (size_t)i0 * L + j;
#pragma unroll 4
        for (int li = 0; li < nc; ++li) a = fmaf(dp[li * NL], cd[(size_t)li * L], a);
        G[lay.o[lat_w(s)] + i] = a;
      }
      if (jb.part == 0)
      for (int c = tid; c < 32; c += 256) {
        float a = i0 ? G[lay.o[lat_w(s) + 1] + c] : 0.f;
        for (int li = 0; li < nc; ++li) a += s_dpre[s][li][c];
        G[lay.o[lat_w(s) + 1] + c] = a;
      }
    }
  } else if (jb.kind == 5) {
    // code tables: d code[inst][j] = sum_s sum_c dpre[s][inst][c] W_lat_s[c][j]  (+ regulariser)
    // this part's slice [lo, hi) of the [n_obj_max][L] tables, walked in chunks of MAXOBJ instances (d z only for those)
    const int total = n_obj_max * L;
    const int lo = part_lo(total), hi = part_hi(total);
    for (int i0 = lo / L; i0 * L < hi; i0 += MAXOBJ) {
      const int nc = max(0, min(min(MAXOBJ, n_obj - i0), (hi - 1) / L + 1 - i0));     // live instances of the chunk
      if (i0 != lo / L) __syncthreads();
      compute_dpre(0, 4, i0, nc);
      if (add_reg && n_obj > 1) compute_norms(i0, nc);
      __syncthreads();
      const int c_lo = max(lo, i0 * L), c_hi = min(hi, (i0 + MAXOBJ) * L);
      for (int i = c_lo + tid; i < c_hi; i += 256) {
        const int inst = (int)__umulhi((unsigned)i, invL), li = inst - i0, j = i - inst * L;
        float gs = 0.f, gt = 0.f;
        if (inst < n_obj) {
#pragma unroll 16
          for (int c = 0; c < 32; ++c) {
            gs = fmaf(s_dpre[0][li][c], P[lay.o[2] + (size_t)c * L + j], gs);
            gs = fmaf(s_dpre[1][li][c], P[lay.o[12] + (size_t)c * L + j], gs);
            gs = fmaf(s_dpre[2][li][c], P[lay.o[6] + (size_t)c * L + j], gs);
            gt = fmaf(s_dpre[3][li][c], P[lay.o[22] + (size_t)c * L + j], gt);
          }
          if (add_reg && n_obj > 1) {                                            // loss.py:12, train.py:165-167
            const float ns = s_norm[0][li], nt = s_norm[1][li];
            if (ns > 0.f) gs += reg_scaling * shape_codes[((size_t)k * n_obj_max + inst) * L + j] / ns;
            if (nt > 0.f) gt += reg_scaling * tex_codes[((size_t)k * n_obj_max + inst) * L + j] / nt;
          }
        }
        g_shape[(size_t)k * n_obj_max * L + i] = gs;
        g_tex[(size_t)k * n_obj_max * L + i] = gt;
      }
    }
  } else {
    float part[2] = {0.f, 0.f};                                    // this warp's share of sum |shape code|, sum |texture code|
    if (n_obj > 1)
      for (int o = warp; o < 2 * n_obj; o += 8) { const float a = code_norm(o / n_obj, o % n_obj); part[o / n_obj] += a; }
    if (lane == 0) { s_norm[0][warp] = part[0]; s_norm[1][warp] = part[1]; }
    __syncthreads();
    __shared__ int s_last;
    if (tid == 0) {
      float rs = 0.f, rt = 0.f;
      for (int w = 0; w < 8; ++w) { rs += s_norm[0][w]; rt += s_norm[1][w]; }
      reg_out[k * 2 + 0] = rs; reg_out[k * 2 + 1] = rt;
      __threadfence();
      s_last = atomicAdd(done_counter, 1) == (int)gridDim.x - 1;     // the last category to finish writes the loss outputs
    }
    __syncthreads();
    if (s_last && tid < 32) {
      __threadfence();
      loss_outputs(gridDim.x, raw, counts, cstride, reg_out, cs, os, reg_scaling, add_reg, secondary, out_loss);
    }
  }
}

struct FusedWs {
  size_t prep, ctab, zlat, raw, rawi, done, reg, hist, base, seg_ray0, seg_tile0, perm, rec, rank_tmp, total;
  int nblk;
  FusedWs(size_t K, size_t R, size_t nom) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    nblk = (int)((R + 255) / 256);
    prep = take(K * PREP_BYTES);
    ctab = take(K * 4 * nom * NL * 4);
    zlat = take(K * 4 * nom * NL * 4);
    raw = take(K * RAW_FLOATS * 4);          // raw and rawi are adjacent: one memset
    rawi = take(K * nom * 4 * 32 * 4);
    done = take(256 + K * 4);                  // finalize completion counter | instance-sort tickets [K] at +256 (zeroed with raw / rawi)
    reg = take(K * 2 * 4);
    hist = take(K * nom * nblk * 4);
    base = take(K * nom * nblk * 4);
    seg_ray0 = take(K * (nom + 1) * 4);
    seg_tile0 = take(K * (nom + 1) * 4);
    perm = take(K * R * 4);
    rec = take(K * R * 16);                    // in-kernel sampling: one record per sorted ray position (small batches)
    rank_tmp = take(K * R * 8);
    total = o;
  }
};

}  // namespace cnf

using namespace cnf;

static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
#ifdef CNF_DEV      // development build only (include/cnf_dev.h): clock-stamp buffer and the instrumented kernel variant
static thread_local long long* g_dbg = nullptr;
extern "C" int cnf_set_debug_buffer(void* p) { g_dbg = reinterpret_cast<long long*>(p); return CNF_OK; }
#else
static constexpr long long* g_dbg = nullptr;
#endif
extern "C" int cnf_set_kernel_events(void* ev_start, void* ev_stop) {
  g_ev_start = reinterpret_cast<cudaEvent_t>(ev_start);
  g_ev_stop = reinterpret_cast<cudaEvent_t>(ev_stop);
  return CNF_OK;
}

extern "C" size_t cnf_train_step_workspace_bytes(int K, int R, int S, int L, int n_obj_max) {
  (void)L; (void)S;
  return FusedWs(K, R, n_obj_max).total;
}

// prep + instance sort + accumulator clear: everything of a step that does not depend on the sampled depths.
// what: bit0 = operand half (fp16 weight images, constants, per-instance latent biases: functions of the PARAMETERS only, so a
// training loop refreshes them once per optimizer step), bit1 = input half (instance sort of this step's rays, accumulator clear).
static int train_step_prepare(int K, int R, int S, int L, int nom, const float* params, const float* pe_B, const float* pe_scale,
                              const float* shape_codes, const float* tex_codes, const int32_t* n_obj, const int32_t* inst,
                              uint8_t* w8, const FusedWs& wl, cudaStream_t st, int what = 3, const SampleIn* si = nullptr,
                              int32_t* stats = nullptr) {
  uint8_t* prep = w8 + wl.prep;
  float* ctab = reinterpret_cast<float*>(w8 + wl.ctab);
  float* zlat = reinterpret_cast<float*>(w8 + wl.zlat);
  float* raw = reinterpret_cast<float*>(w8 + wl.raw);
  int32_t* hist = reinterpret_cast<int32_t*>(w8 + wl.hist);
  int32_t* base = reinterpret_cast<int32_t*>(w8 + wl.base);
  int32_t* seg_ray0 = reinterpret_cast<int32_t*>(w8 + wl.seg_ray0);
  int32_t* seg_tile0 = reinterpret_cast<int32_t*>(w8 + wl.seg_tile0);
  int32_t* perm = reinterpret_cast<int32_t*>(w8 + wl.perm);
  const int rpw = 32 / S, RPT = 4 * rpw, nblk = wl.nblk;
  const int n_prep = PREP_IMG_PARTS + prep_lat_parts(nom);
  if (what == 1) {                                         // operand half only
    fused_prep_kernel<<<dim3(K, n_prep), 256, 0, st>>>(L, nom, params, pe_B, pe_scale, shape_codes, tex_codes, n_obj, prep, zlat, ctab);
    return check_launch("cnf_train_step_refresh");
  }
  if (nblk <= SORT_SMALL_BLOCKS) {
    // one launch: [images + biases] | instance sort | clear of raw, rawi and the finalize completion counter (16-byte chunks)
    const int first = (what & 1) ? 0 : n_prep;             // input half only: skip the operand parts
    static DeviceOnce pattr;
    if (pattr.pending()) {
      CNF_CUDA(cudaFuncSetAttribute(prepare_small_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, CNF_SMALL_STEP_CARVEOUT));
      pattr.done();
    }
    prepare_small_kernel<<<dim3(K, n_prep + 1 + PREP_ZERO_PARTS - first), 256, nom * sizeof(int), st>>>(
        L, nom, params, pe_B, pe_scale, shape_codes, tex_codes, n_obj, prep, zlat, ctab, R, RPT, inst, seg_ray0, seg_tile0, perm,
        reinterpret_cast<uint4*>(raw), (wl.reg - wl.raw) / 16, first, si ? *si : SampleIn{}, reinterpret_cast<int4*>(w8 + wl.rec),
        reinterpret_cast<int2*>(w8 + wl.rank_tmp), stats, g_dbg);
    return check_launch("cnf_train_step_prepare");
  }
  CNF_CUDA(cudaMemsetAsync(raw, 0, wl.reg - wl.raw, st));            // raw, rawi and the finalize completion counter
  if (what & 1)
    fused_prep_kernel<<<dim3(K, n_prep), 256, 0, st>>>(L, nom, params, pe_B, pe_scale, shape_codes, tex_codes, n_obj, prep, zlat, ctab);
  {
    sort_count_kernel<<<dim3(nblk, K), 256, nom * sizeof(int), st>>>(R, nom, nblk, inst, n_obj, hist);
    sort_scan_kernel<<<dim3(nom, K), SORT_SCAN_T, nom * sizeof(int), st>>>(nom, nblk, RPT, hist, base, seg_ray0, seg_tile0,
                                                                           reinterpret_cast<unsigned*>(w8 + wl.done + 256));
    sort_scatter_kernel<<<dim3(nblk, K), 256, 0, st>>>(R, nom, nblk, inst, n_obj, base, seg_ray0, perm);
  }
  return check_launch("cnf_train_step_prepare");
}

static int check_step_sizes(const char* who, int K, int R, int S, int L, int n_obj_max) {
  if (!(K > 0 && R > 0 && L > 0)) { set_error("%s: bad sizes K=%d R=%d L=%d", who, K, R, L); return CNF_ERR_ARG; }
  if (S < 1 || S > 32) { set_error("%s: samples per ray S=%d outside 1..32 (one ray per warp segment)", who, S); return CNF_ERR_UNSUPPORTED; }
  if (n_obj_max < 1 || n_obj_max > MAX_INSTANCES) { set_error("%s: n_obj_max=%d outside 1..%d", who, n_obj_max, MAX_INSTANCES); return CNF_ERR_UNSUPPORTED; }
  return CNF_OK;
}

extern "C" int cnf_train_step_prepare(int K, int R, int S, int L, int n_obj_max, const float* params, const float* pe_B,
                                      const float* pe_scale, const float* shape_codes, const float* tex_codes,
                                      const int32_t* n_obj, const int32_t* inst, int what, void* ws, size_t ws_bytes,
                                      cnf_stream_t stream) {
  if (int e = check_step_sizes("cnf_train_step_prepare", K, R, S, L, n_obj_max)) return e;
  CNF_REQUIRE(what >= 1 && what <= 3, "cnf_train_step_prepare: what = %d (1 operands, 2 inputs, 3 both)", what);
  CNF_REQUIRE(params && pe_B && pe_scale && shape_codes && tex_codes && n_obj && ((what & 2) == 0 || inst) && ws,
              "cnf_train_step_prepare: null pointer");
  const FusedWs wl(K, R, n_obj_max);
  if (ws_bytes < wl.total) { set_error("cnf_train_step_prepare: workspace too small (%zu < %zu)", ws_bytes, wl.total); return CNF_ERR_WORKSPACE; }
  return train_step_prepare(K, R, S, L, n_obj_max, params, pe_B, pe_scale, shape_codes, tex_codes, n_obj, inst,
                            reinterpret_cast<uint8_t*>(ws), wl, as_stream(stream), what);
}

// fused kernel + finalize (the prepare half has run): shared by cnf_train_step and cnf_train_step_sampled
static int run_fused(int K, int R, int S, int L, int n_obj_max, const float* params, const float* pe_scale, const float* shape_codes,
                     const float* tex_codes, const int32_t* n_obj, const float* origins, const float* dirs, const float* z,
                     const float* gt_depth, const uint8_t* gt_rgb, const uint8_t* mask_depth, const int32_t* counts, float cs, float os,
                     float reg_scaling, int flags, float* g_params, float* g_B, float* g_shape, float* g_tex, float* out_loss,
                     uint8_t* w8, const FusedWs& wl, cudaStream_t st, const SampleIn* si) {
  const int nom = n_obj_max;
  float* raw = reinterpret_cast<float*>(w8 + wl.raw);
  float* rawi = reinterpret_cast<float*>(w8 + wl.rawi);
  float* reg = reinterpret_cast<float*>(w8 + wl.reg);
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(fused_step_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    CNF_CUDA(cudaFuncSetAttribute(fused_step_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
#ifdef CNF_DEV
    CNF_CUDA(cudaFuncSetAttribute(fused_step_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    CNF_CUDA(cudaFuncSetAttribute(fused_step_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
#endif
    attr_set.done();
  }
  int dev = 0, n_sm = 148;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int rpw = 32 / S, RPT = 4 * rpw, n_tiles_ub = (R + RPT - 1) / RPT + n_obj_max;
  int cpc = n_sm / K; if (cpc < 1) cpc = 1; if (cpc > n_tiles_ub) cpc = n_tiles_ub;
  FusedArgs fa;
  fa.K = K; fa.R = R; fa.S = S; fa.nom = nom;
  fa.prep = w8 + wl.prep; fa.ctab = reinterpret_cast<float*>(w8 + wl.ctab);
  fa.origins = origins; fa.dirs = dirs; fa.zs = z; fa.gt_depth = gt_depth; fa.gt_rgb = gt_rgb; fa.mask_depth = mask_depth;
  fa.perm = reinterpret_cast<int32_t*>(w8 + wl.perm); fa.seg_ray0 = reinterpret_cast<int32_t*>(w8 + wl.seg_ray0);
  fa.seg_tile0 = reinterpret_cast<int32_t*>(w8 + wl.seg_tile0); fa.n_obj = n_obj; fa.counts = counts;
  const int cstride = ((flags >> 8) & 0xFF) ? ((flags >> 8) & 0xFF) : 4;
  fa.cstride = cstride;
  fa.cs = cs; fa.os = os; fa.raw = raw; fa.rawi = rawi; fa.dbg = g_dbg;
  fa.si = si ? *si : SampleIn{}; fa.rec = reinterpret_cast<const int4*>(w8 + wl.rec);
  if (g_ev_start && g_ev_stop) CNF_CUDA(cudaEventRecord(g_ev_start, st));
#ifdef CNF_DEV
  if (g_dbg) { if (si) fused_step_kernel<true, true><<<dim3(cpc, K), NT, SM_TOTAL, st>>>(fa); else fused_step_kernel<true, false><<<dim3(cpc, K), NT, SM_TOTAL, st>>>(fa); }
  else
#endif
  if (si) fused_step_kernel<false, true><<<dim3(cpc, K), NT, SM_TOTAL, st>>>(fa);
  else fused_step_kernel<false, false><<<dim3(cpc, K), NT, SM_TOTAL, st>>>(fa);
  if (g_ev_start && g_ev_stop) { CNF_CUDA(cudaEventRecord(g_ev_stop, st)); g_ev_start = g_ev_stop = nullptr; }
  int pm = (2 * n_sm) / (K * FIN_JOBS);                 // part multiplier: fill the machine when there are few categories
  if (pm < 1) pm = 1;
  if (pm > 8) pm = 8;
  fused_finalize_kernel<<<dim3(K, fin_jobs(pm)), 256, 0, st>>>(L, nom, params, pe_scale, shape_codes, tex_codes, n_obj, raw, rawi,
                                                            reinterpret_cast<float*>(w8 + wl.zlat), reg_scaling, flags & 1, (flags >> 2) & 1, g_params, g_B,
                                                            g_shape, g_tex, reg, counts, cstride, cs, os, reinterpret_cast<int*>(w8 + wl.done), out_loss,
                                                            g_dbg);
  return check_launch("cnf_train_step");
}

extern "C" int cnf_train_step(int K, int R, int S, int L, int n_obj_max, const float* params, const float* pe_B,
                              const float* pe_scale, const float* shape_codes, const float* tex_codes,
                              const int32_t* n_obj, const float* origins, const float* dirs, const float* z,
                              const float* gt_depth, const uint8_t* gt_rgb, const uint8_t* mask_depth,
                              const int32_t* inst, const int32_t* counts, float cs, float os, float reg_scaling,
                              int flags, float* g_params, float* g_B, float* g_shape, float* g_tex, float* out_loss,
                              void* ws, size_t ws_bytes, cnf_stream_t stream) {
  if (int e = check_step_sizes("cnf_train_step", K, R, S, L, n_obj_max)) return e;
  CNF_REQUIRE(params && pe_B && pe_scale && shape_codes && tex_codes && n_obj && origins && dirs && z && gt_depth && gt_rgb &&
                  mask_depth && inst && counts && g_params && g_B && g_shape && g_tex && out_loss && ws,
              "cnf_train_step: null pointer");
  const FusedWs wl(K, R, n_obj_max);
  if (ws_bytes < wl.total) { set_error("cnf_train_step: workspace too small (%zu < %zu)", ws_bytes, wl.total); return CNF_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  if (!(flags & 2)) {          // not prepared by cnf_train_step_prepare (e.g. on a side stream, overlapping the sampler)
    if (int e = train_step_prepare(K, R, S, L, n_obj_max, params, pe_B, pe_scale, shape_codes, tex_codes, n_obj, inst, w8, wl, st)) return e;
  }
  return run_fused(K, R, S, L, n_obj_max, params, pe_scale, shape_codes, tex_codes, n_obj, origins, dirs, z, gt_depth, gt_rgb, mask_depth,
                   counts, cs, os, reg_scaling, flags, g_params, g_B, g_shape, g_tex, out_loss, w8, wl, st, nullptr);
}

extern "C" int cnf_train_step_sampled_max_rays(void) { return SORT_SMALL_BLOCKS * 256; }

extern "C" int cnf_train_step_sampled(int K, int R, int N, int M, int L, int n_obj_max, const float* params, const float* pe_B,
                                      const float* pe_scale, const float* shape_codes, const float* tex_codes, const int32_t* n_obj,
                                      const uint8_t* rgbs, const float* depth, const float* origins, const float* dirs, const int32_t* inst,
                                      const float* u_inv, const float* u_val, const float* n_draws, const float* u_oth,
                                      const float* lin_S, const float* lin_N, const float* lin_M, float min_bound, float eps, float stop_eps,
                                      int this_obj, const float* max_depth_override, float cs, float os, float reg_scaling, int flags,
                                      float* g_params, float* g_B, float* g_shape, float* g_tex, float* out_loss, int32_t* stats,
                                      float* z_out, void* ws, size_t ws_bytes, cnf_stream_t stream) {
  const int S = N + M;
  if (int e = check_step_sizes("cnf_train_step_sampled", K, R, S, L, n_obj_max)) return e;
  CNF_REQUIRE(N >= 0 && M >= 1, "cnf_train_step_sampled: bad bin counts N=%d M=%d", N, M);
  CNF_REQUIRE(params && pe_B && pe_scale && shape_codes && tex_codes && n_obj && rgbs && depth && origins && dirs && inst && u_inv &&
                  u_val && n_draws && u_oth && lin_S && lin_M && (N == 0 || lin_N) && g_params && g_B && g_shape && g_tex && out_loss &&
                  stats && ws,
              "cnf_train_step_sampled: null pointer");
  const FusedWs wl(K, R, n_obj_max);
  if (ws_bytes < wl.total) { set_error("cnf_train_step_sampled: workspace too small (%zu < %zu)", ws_bytes, wl.total); return CNF_ERR_WORKSPACE; }
  if (wl.nblk > SORT_SMALL_BLOCKS) {
    set_error("cnf_train_step_sampled: R=%d > %d rays per category (use cnf_sample_fwd + cnf_train_step)", R, SORT_SMALL_BLOCKS * 256);
    return CNF_ERR_UNSUPPORTED;
  }
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  SampleIn si;
  si.rgbs = rgbs; si.depth = depth; si.u_inv = u_inv; si.u_val = u_val; si.n_obj = n_draws; si.u_oth = u_oth;
  si.lin_S = lin_S; si.lin_N = N ? lin_N : lin_S; si.lin_M = lin_M; si.max_override = max_depth_override;
  si.min_bound = min_bound; si.eps = eps; si.stop_eps = stop_eps; si.this_obj = this_obj; si.N = N; si.M = M; si.z_out = z_out;
  // flags bit3: the operand half (weight images, latent biases) is up to date (cnf_train_step_prepare(what = 1) after the optimizer)
  if (int e = train_step_prepare(K, R, S, L, n_obj_max, params, pe_B, pe_scale, shape_codes, tex_codes, n_obj, inst, w8, wl, st,
                                 (flags & 8) ? 2 : 3, &si, stats)) return e;
  const int f2 = (flags & 5) | (8 << 8);                  // counts = stats + 4 with row stride 8 (stats[k][7] = max depth bits)
  return run_fused(K, R, S, L, n_obj_max, params, pe_scale, shape_codes, tex_codes, n_obj, origins, dirs, nullptr, depth, rgbs, nullptr,
                   stats + 4, cs, os, reg_scaling, f2, g_params, g_B, g_shape, g_tex, out_loss, w8, wl, st, &si);
}

static int eval_set_attrs() {
  static DeviceOnce done;
  if (done.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(eval_points_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EV_SMEM));
    CNF_CUDA(cudaFuncSetAttribute(eval_points_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EV_SMEM));
    done.done();
  }
  return CNF_OK;
}

// ---- forward-only evaluation at arbitrary points with one instance's codes (trainer.py:125-151)
static size_t eval_ws_prep() { return 0; }
static size_t eval_ws_ctab() { return align_up(PREP_BYTES, 256); }
static size_t eval_ws_zlat() { return eval_ws_ctab() + 512; }
static size_t eval_ws_nobj() { return eval_ws_zlat() + 512; }
extern "C" size_t cnf_eval_points_workspace_bytes(void) { return eval_ws_nobj() + 256; }

extern "C" int cnf_eval_points(int P, int L, const float* params, const float* pe_B, const float* pe_scale, const float* shape_code,
                               const float* tex_code, const float* points, float* occ, float* rgb, void* ws, size_t ws_bytes,
                               cnf_stream_t stream) {
  CNF_REQUIRE(P >= 0 && L > 0, "cnf_eval_points: bad sizes P=%d L=%d", P, L);
  CNF_REQUIRE(params && pe_B && pe_scale && shape_code && tex_code && ws && (P == 0 || (points && occ && rgb)), "cnf_eval_points: null pointer");
  if (ws_bytes < cnf_eval_points_workspace_bytes()) { set_error("cnf_eval_points: workspace too small"); return CNF_ERR_WORKSPACE; }
  if (P == 0) return CNF_OK;
  if (int e = eval_set_attrs()) return e;
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  int32_t* n_obj = reinterpret_cast<int32_t*>(w8 + eval_ws_nobj());
  CNF_CUDA(cudaMemsetAsync(n_obj, 1, 4, st));                      // any value >= 1: one instance (clamped to n_obj_max = 1)
  fused_prep_kernel<<<dim3(1, PREP_IMG_PARTS + 4), 256, 0, st>>>(L, 1, params, pe_B, pe_scale, shape_code, tex_code, n_obj, w8 + eval_ws_prep(),
                                                                  reinterpret_cast<float*>(w8 + eval_ws_zlat()), reinterpret_cast<float*>(w8 + eval_ws_ctab()));
  int dev = 0, n_sm = 148;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int n_tiles = (P + TP - 1) / TP;
  int grid = (n_tiles + EV_CHAINS - 1) / EV_CHAINS; if (grid > n_sm) grid = n_sm;
  eval_points_kernel<false><<<grid, EV_NT, EV_SMEM, st>>>(P, w8 + eval_ws_prep(), reinterpret_cast<const float*>(w8 + eval_ws_ctab()), points, occ, rgb,
                                                           1, nullptr, nullptr);
  return check_launch("cnf_eval_points");
}

// ---- forward-only rendering of rays with one instance's codes (BASELINE config 5: dense views; category_registration.py:146-152
// uncertainty fields): tensor-core point evaluation in chunks whose per-point results stay in L2, then per-ray compositing.
constexpr int RENDER_CHUNK_POINTS = 4 << 20;                      // 4 Mi points x 16 B = 64 MiB of {occ, rgb} per chunk (L2: 126 MB)
static size_t render_chunk_rays(int R, int S) { size_t rc = (size_t)RENDER_CHUNK_POINTS / S; if (rc < 1) rc = 1; return rc < (size_t)R ? rc : (size_t)R; }
static size_t render_ws_pts() { return align_up(cnf_eval_points_workspace_bytes(), 256); }
extern "C" size_t cnf_render_rays_workspace_bytes(int R, int S) {
  if (R < 1 || S < 1) return render_ws_pts();
  return render_ws_pts() + render_chunk_rays(R, S) * (size_t)S * 16;
}

extern "C" int cnf_render_rays(int R, int S, int L, const float* params, const float* pe_B, const float* pe_scale,
                               const float* shape_code, const float* tex_code, const float* origins, const float* dirs,
                               const float* z, float* out_rays, float* out_w, void* ws, size_t ws_bytes, cnf_stream_t stream) {
  CNF_REQUIRE(R >= 0 && S >= 1 && L > 0, "cnf_render_rays: bad sizes R=%d S=%d L=%d", R, S, L);
  CNF_REQUIRE(params && pe_B && pe_scale && shape_code && tex_code && ws && (R == 0 || (origins && dirs && z && out_rays)),
              "cnf_render_rays: null pointer");
  if (ws_bytes < cnf_render_rays_workspace_bytes(R, S)) { set_error("cnf_render_rays: workspace too small"); return CNF_ERR_WORKSPACE; }
  if (R == 0) return CNF_OK;
  if (int e = eval_set_attrs()) return e;
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  int32_t* n_obj = reinterpret_cast<int32_t*>(w8 + eval_ws_nobj());
  CNF_CUDA(cudaMemsetAsync(n_obj, 1, 4, st));
  fused_prep_kernel<<<dim3(1, PREP_IMG_PARTS + 4), 256, 0, st>>>(L, 1, params, pe_B, pe_scale, shape_code, tex_code, n_obj, w8 + eval_ws_prep(),
                                                                  reinterpret_cast<float*>(w8 + eval_ws_zlat()), reinterpret_cast<float*>(w8 + eval_ws_ctab()));
  int dev = 0, n_sm = 148;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  float* pts = reinterpret_cast<float*>(w8 + render_ws_pts());
  const size_t rc = render_chunk_rays(R, S);
  for (size_t r0 = 0; r0 < (size_t)R; r0 += rc) {
    const int nr = (int)((size_t)R - r0 < rc ? (size_t)R - r0 : rc);
    const long long P = (long long)nr * S;
    CNF_REQUIRE(P < (1ll << 31), "cnf_render_rays: chunk too large");
    const int n_tiles = (int)((P + TP - 1) / TP);
    int grid = (n_tiles + EV_CHAINS - 1) / EV_CHAINS; if (grid > n_sm) grid = n_sm;
    eval_points_kernel<true><<<grid, EV_NT, EV_SMEM, st>>>((int)P, w8 + eval_ws_prep(), reinterpret_cast<const float*>(w8 + eval_ws_ctab()),
                                                            origins + r0 * 3, pts, nullptr, S, dirs + r0 * 3, z + r0 * S);
    if (int e = render_composite_launch(nr, S, pts, z + r0 * S, out_rays + r0 * 6, out_w ? out_w + r0 * S : nullptr, st)) return e;
  }
  return check_launch("cnf_render_rays");
}
