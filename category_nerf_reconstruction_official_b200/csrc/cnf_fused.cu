// Fused training step on Blackwell tensor cores (tcgen05 / TMEM), bf16 operands, fp32 accumulation.
//
//   embed -> CodeNeRF (W=32) -> occupancy compositing -> depth/colour/opacity losses, forward AND backward,
//   for K categories in ONE launch (reference train.py:136-182 minus the optimizer; math: src/embedding.py:82-92,
//   src/model.py:56-84, src/render_rays.py:25-95, src/loss.py:18-74).
//
// One CTA = 288 threads walks tiles of 128 sample points (whole rays, warp-aligned):
//   * warps 0-7 are the epilogue warps: TWO threads per tile row (thread pair (row, half) splits every layer's columns,
//     the embedding directions and the embedding backward), row r <-> TMEM lane r;
//   * warp 8 only issues tcgen05.mma: it waits on a named barrier until the operand tiles of a stage are published,
//     issues the stage's data-path MMAs, commits them to an mbarrier, then issues the weight-gradient MMAs of the
//     stage, which run on the tensor pipe while the epilogue warps already work on the data-path result.
// Per tile the layers run as UMMA tiles M=128 x N<=128 x K<=128:
//   * operands live in shared memory in the canonical no-swizzle 8x16B core-matrix layout; the SAME bytes are
//     read K-major by the forward GEMM, MN-major by the data-gradient GEMM (weights) and MN-major by the
//     weight-gradient GEMM (activations, upstream gradients) -- no transposes are ever materialised;
//   * weight gradients, and the per-instance bias gradients of the latent-conditioned layers (via a one-hot
//     instance operand), accumulate in TMEM across ALL tiles of the CTA and are flushed once at the end;
//   * biases of layer 1 / 6 ride in a constant-one embedding column (free in the GEMM, and its weight-gradient
//     row is the bias gradient); latent adds are folded into per-instance biases c_l[inst] = W_l z_l[inst] + b_l;
//   * encoding_shape (linear, no activation) is folded algebraically into its two linear consumers
//     (W6' = Wvd[:, :32] W5, wsig' = wsig W5); finalize un-folds the gradients by the chain rule;
//   * the Fourier embedding is computed in registers (one MUFU sin/cos per direction + double-angle recurrence)
//     straight into the layer-1 / layer-3 / layer-6 operand tiles, in a column order chosen for the thread pairs
//     (weight images and gradient read-back use the same permutation); compositing is a warp-shuffle segmented scan.
// A prep kernel builds the bf16 operand images and per-instance biases; a finalize kernel maps the raw TMEM
// accumulators back to the reference's nn.Linear-shaped gradients (incl. latent layers, code tables, regulariser).
#include "cnf_common.cuh"
#include "cnf_umma.cuh"

namespace cnf {
using namespace umma;

// ---------------------------------------------------------------------------------------------------------
// static layout
// ---------------------------------------------------------------------------------------------------------
constexpr int TP = 128;           // tile rows (points)
constexpr int MAXOBJ = 64;        // instances per category supported by the one-hot operand
constexpr int NL = 32;            // layer width W
constexpr int NT = 288;           // + the MMA-issuing warp

// Column order of the embedding operand tiles (per thread-pair half h: directions d < 11 / d >= 11, d-major, f-minor):
//   xg tile (96): [t0 t1 t2 ONE 0 0 0 0 | h0: 11 dirs x f0..3 (44) + 4 pad | h1: 10 dirs x f0..3 (40)]
//   xc tile (48): [h0: 11 dirs x f4,5 (22), ONE, 0 | h1: 10 dirs x f4,5 (20) + 4 pad]
constexpr int SRC_ZERO = -1, SRC_BIAS = -2;
__host__ __device__ constexpr int xg_src(int c) {          // tile column -> reference embedding index (0..86)
  if (c < 3) return c;
  if (c == 3) return SRC_BIAS;
  if (c < 8) return SRC_ZERO;
  if (c < 56) { const int j = c - 8, d = j / 4, f = j % 4; return d < 11 ? 3 + f * 21 + d : SRC_ZERO; }
  { const int j = c - 56, d = 11 + j / 4, f = j % 4; return 3 + f * 21 + d; }
}
__host__ __device__ constexpr int xc_src(int c) {          // tile column -> reference embedding index (87..128)
  if (c < 24) { const int d = c / 2, f = c % 2; return d < 11 ? 87 + f * 21 + d : (c == 22 ? SRC_BIAS : SRC_ZERO); }
  { const int j = c - 24, d = 11 + j / 2, f = j % 2; return d < 21 ? 87 + f * 21 + d : SRC_ZERO; }
}
__host__ __device__ constexpr int xg_col(int e) {          // inverse: embedding index (0..86) -> xg tile column
  if (e < 3) return e;
  const int f = (e - 3) / 21, d = (e - 3) % 21;
  return d < 11 ? 8 + d * 4 + f : 56 + (d - 11) * 4 + f;
}
__host__ __device__ constexpr int xc_col(int e) {          // embedding index (87..128) -> xc tile column
  const int f = (e - 87) / 21, d = (e - 87) % 21;
  return d < 11 ? d * 2 + f : 24 + (d - 11) * 2 + f;
}
constexpr int XG_BIAS_COL = 3, XC_BIAS_COL = 22;

// bf16 weight images (blocked layout, rows x cols), byte offsets in the weight blob
constexpr int W1_R = 32, W1_C = 96;     // encoding_xyz      (xg column order; col 3 = bias via the ones column)
constexpr int W2_R = 32, W2_C = 32;     // shape_layer_1
constexpr int W3_R = 32, W3_C = 128;    // cat_layer         cols [r2 32 | xg 96]
constexpr int W4_R = 32, W4_C = 32;     // shape_layer_2
constexpr int W6_R = 48, W6_C = 80;     // encoding_viewdir o encoding_shape (+ row 32 = sigma head): cols [r4 32 | xc 48]
constexpr int W7_R = 32, W7_C = 32;     // texture_layer_1
constexpr int W8_R = 16, W8_C = 32;     // rgb.0
constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + W1_R * W1_C * 2;
constexpr int OFF_W3 = OFF_W2 + W2_R * W2_C * 2;
constexpr int OFF_W4 = OFF_W3 + W3_R * W3_C * 2;
constexpr int OFF_W6 = OFF_W4 + W4_R * W4_C * 2;
constexpr int OFF_W7 = OFF_W6 + W6_R * W6_C * 2;
constexpr int OFF_W8 = OFF_W7 + W7_R * W7_C * 2;
constexpr int WBLOB_BYTES = OFF_W8 + W8_R * W8_C * 2;     // 29184
// fp32 constants block (floats)
constexpr int FC_B8 = 0;            // [16]
constexpr int FC_W9 = 16;           // [3][16]
constexpr int FC_B9 = 64;           // [3] (+1 pad)
constexpr int FC_PEB = 68;          // [21][3] (+1 pad)
constexpr int FC_SCALE = 132;       // inv scale (+3 pad)
constexpr int FC_FLOATS = 136;
constexpr int CTAB_FLOATS = 4 * MAXOBJ * NL;   // c2, c3, c4, c7 per instance
constexpr int PREP_BYTES = WBLOB_BYTES + (FC_FLOATS + CTAB_FLOATS) * 4;   // per category image copied to smem

// activation / gradient tiles in shared memory (bf16 blocked [128 x cols])
constexpr int SM_W = 0;
constexpr int SM_FC = SM_W + WBLOB_BYTES;
constexpr int SM_CTAB = SM_FC + FC_FLOATS * 4;
constexpr int SM_R2XG = SM_CTAB + CTAB_FLOATS * 4;          // [128 x 128]: r2 | xg
constexpr int SM_R4XC = SM_R2XG + TP * 128 * 2;             // [128 x 80] : r4 | xc
constexpr int SM_R1 = SM_R4XC + TP * 80 * 2;
constexpr int SM_R3 = SM_R1 + TP * 32 * 2;
constexpr int SM_H1 = SM_R3 + TP * 32 * 2;
constexpr int SM_H2Z3 = SM_H1 + TP * 32 * 2;                // [128 x 80] : h2 | zeros 32 | h3 16  (later: parked d xc)
constexpr int SM_OH = SM_H2Z3 + TP * 80 * 2;                // [128 x 64] one-hot(instance)
constexpr int SM_GX = SM_OH + TP * 64 * 2;                  // 12 KB gradient tile (G8+G9, G6, G3, G1)
constexpr int SM_GY = SM_GX + TP * 48 * 2;                  // 8 KB gradient tile (G7, G4, G2)
constexpr int SM_PAD = SM_GY + TP * 32 * 2;                 // over-read slack for M=128 views of narrow tiles
constexpr int SM_TOTAL = SM_PAD + 4096;
static_assert(SM_R2XG % 16 == 0 && PREP_BYTES % 16 == 0, "tile alignment");

// TMEM columns
constexpr int TM_WG1 = 0, TM_WG2 = 32, TM_WG3 = 64, TM_WG4 = 96, TM_WG6 = 160, TM_WG7 = 208, TM_WG89 = 240;
constexpr int TM_OH2 = 256, TM_OH3 = 288, TM_OH4 = 320, TM_OH7 = 352, TM_ACC = 384;

// raw accumulator layout flushed to global (floats, [lane][col] per region)
constexpr int RW1 = 0;                      // [96][32]   lane = xg tile column
constexpr int RW2 = RW1 + 96 * 32;          // [32][32]
constexpr int RW3 = RW2 + 32 * 32;          // [128][32]  lane = r2 feature | 32 + xg tile column
constexpr int RW4 = RW3 + 128 * 32;         // [32][32]
constexpr int RW6 = RW4 + 32 * 32;          // [80][48]   lane = r4 feature | 32 + xc tile column ; col 32 = sigma
constexpr int RW7 = RW6 + 80 * 48;          // [32][32]
constexpr int RW89 = RW7 + 32 * 32;         // [48][16]
constexpr int ROH2 = RW89 + 48 * 16;        // [64][32] x 4
constexpr int ROH3 = ROH2 + 64 * 32;
constexpr int ROH4 = ROH3 + 64 * 32;
constexpr int ROH7 = ROH4 + 64 * 32;
constexpr int RDB8 = ROH7 + 64 * 32;        // [16]
constexpr int RDB9 = RDB8 + 16;             // [3] (+1)
constexpr int RLOSS = RDB9 + 4;             // [3] raw loss sums (+1)
constexpr int RAW_FLOATS = RLOSS + 4;

// parameter indices (SURVEY appendix D order)
struct PLay {
  size_t o[28]; size_t total; int L;
  __host__ __device__ explicit PLay(int L_) : L(L_) {
    const size_t w = 32, l = L_, E1 = CNF_EMB1, E2 = CNF_EMB2;
    const size_t sz[28] = {w * E1, w, w * l, w, w * w, w, w * l, w, w * w, w, w * (w + E1), w, w * l, w, w * w, w,
                           w, 1, w * (w + E2), w, w * w, w, w * l, w, (w / 2) * w, w / 2, 3 * (w / 2), 3};
    size_t off = 0;
    for (int i = 0; i < 28; ++i) { o[i] = off; off += sz[i]; }
    total = off;
  }
};

// ---------------------------------------------------------------------------------------------------------
// prep: per category -> {bf16 weight images, fp32 constants, per-instance biases} + latent activations z
// grid = (K, 5), block = 256.  y = 0: weight images (one 16-byte core-matrix row per thread-iteration, coalesced)
// + constants; y = 1..4: latent slot s = y-1 (zsl1, zcl, zsl2, ztl): z = relu(W_lat code + b) for the category's
// instances (warp per output, coalesced rows), then the per-instance bias c = W_cons[:, :32] z + b_cons.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float w_img_value(const float* __restrict__ P, const PLay& lay, int img, int r, int c) {
  // value of element (r, c) of weight image `img` (0..6 = W1,W2,W3,W4,W6',W7,W8), zero outside the valid region
  switch (img) {
    case 0: {
      const int e = xg_src(c);
      return e >= 0 ? P[lay.o[0] + r * 87 + e] : (e == SRC_BIAS ? P[lay.o[1] + r] : 0.f);
    }
    case 1: return P[lay.o[4] + r * 32 + c];
    case 2: {
      if (c < 32) return P[lay.o[10] + r * 119 + c];
      const int e = xg_src(c - 32);
      return e >= 0 ? P[lay.o[10] + r * 119 + 32 + e] : 0.f;     // bias of the cat layer is per-instance (c3)
    }
    case 3: return P[lay.o[8] + r * 32 + c];
    case 4: {
      // rows 0..31: encoding_viewdir folded with encoding_shape; row 32: sigma head folded likewise; rows 33..47 zero
      if (r > 32) return 0.f;
      const float* wrow = r < 32 ? P + lay.o[18] + r * 74 : P + lay.o[16];      // [.., :32] part of the consumer row
      if (c < 32) {                                                             // (W_cons[:, :32] W5)[r][c]
        float a = 0.f;
        for (int j = 0; j < 32; ++j) a = fmaf(wrow[j], P[lay.o[14] + j * 32 + c], a);
        return a;
      }
      const int e = xc_src(c - 32);
      if (e >= 0) return r < 32 ? P[lay.o[18] + r * 74 + 32 + (e - 87)] : 0.f;
      if (e == SRC_BIAS) {                                                      // W_cons[:, :32] b5 + b_cons
        float a = r < 32 ? P[lay.o[19] + r] : P[lay.o[17]];
        for (int j = 0; j < 32; ++j) a = fmaf(wrow[j], P[lay.o[15] + j], a);
        return a;
      }
      return 0.f;
    }
    case 5: return P[lay.o[20] + r * 32 + c];
    default: return P[lay.o[24] + r * 32 + c];
  }
}

__global__ void __launch_bounds__(256)
fused_prep_kernel(int L, int n_obj_max, const float* __restrict__ params, const float* __restrict__ pe_B,
                  const float* __restrict__ pe_scale, const float* __restrict__ shape_codes,
                  const float* __restrict__ tex_codes, const int32_t* __restrict__ n_obj_arr,
                  uint8_t* __restrict__ prep, float* __restrict__ zlat) {
  const int k = blockIdx.x, tid = threadIdx.x;
  const PLay lay(L);
  const float* P = params + (size_t)k * lay.total;
  uint8_t* img = prep + (size_t)k * PREP_BYTES;
  float* fc = reinterpret_cast<float*>(img + WBLOB_BYTES);
  float* ctab = fc + FC_FLOATS;
  float* z = zlat + (size_t)k * 4 * MAXOBJ * NL;      // [4][MAXOBJ][32]: zsl1, zcl, zsl2, ztl
  const int n_obj = min(n_obj_arr[k], MAXOBJ);
  if (blockIdx.y == 0) {
    const int offs[8] = {OFF_W1, OFF_W2, OFF_W3, OFF_W4, OFF_W6, OFF_W7, OFF_W8, WBLOB_BYTES};
    const int colsv[7] = {W1_C, W2_C, W3_C, W4_C, W6_C, W7_C, W8_C};
    for (int q = tid; q < WBLOB_BYTES / 16; q += 256) {
      int im = 0;
      while (q * 16 >= offs[im + 1]) ++im;
      const int cols = colsv[im];
      const int ql = q - offs[im] / 16;                  // chunk inside the image: rg*cols + cb*8 + r8
      const int rg = ql / cols, rem = ql - rg * cols, cb = rem >> 3, r8 = rem & 7;
      const int r = rg * 8 + r8, c0 = cb * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = w_img_value(P, lay, im, r, c0 + e);
      uint4 o;
      __nv_bfloat162 h;
      h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
      reinterpret_cast<uint4*>(img)[q] = o;
    }
    for (int i = tid; i < FC_FLOATS; i += 256) {
      float v = 0.f;
      if (i < FC_W9) v = P[lay.o[25] + i];
      else if (i < FC_B9) v = P[lay.o[26] + (i - FC_W9)];
      else if (i < FC_B9 + 3) v = P[lay.o[27] + (i - FC_B9)];
      else if (i >= FC_PEB && i < FC_PEB + 63) v = pe_B[(size_t)k * 63 + (i - FC_PEB)];
      else if (i == FC_SCALE) v = 1.0f / pe_scale[k];
      fc[i] = v;
    }
    return;
  }
  const int s = blockIdx.y - 1;
  const int lat_w[4] = {2, 12, 6, 22};
  const int cons_w[4] = {4, 10, 8, 20};
  const int cons_ld[4] = {32, 119, 32, 32};
  __shared__ float s_z[MAXOBJ * NL];
  const int warp = tid >> 5, lane = tid & 31;
  const float* codes = (s == 3 ? tex_codes : shape_codes) + (size_t)k * n_obj_max * L;
  for (int o = warp; o < n_obj * NL; o += 8) {           // warp per (inst, n): coalesced dot product over L
    const int inst = o / NL, n = o % NL;
    const float* w = P + lay.o[lat_w[s]] + (size_t)n * L;
    const float* code = codes + (size_t)inst * L;
    float acc = 0.f;
    for (int j = lane; j < L; j += 32) acc = fmaf(w[j], code[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_z[o] = fmaxf(acc + P[lay.o[lat_w[s] + 1] + n], 0.f);
  }
  __syncthreads();
  for (int i = tid; i < MAXOBJ * NL; i += 256) {
    const int inst = i / NL, n = i % NL;
    z[s * MAXOBJ * NL + i] = inst < n_obj ? s_z[i] : 0.f;
    float acc = P[lay.o[cons_w[s] + 1] + n];             // instances beyond n_obj get the plain bias (z = 0)
    if (inst < n_obj) {
      const float* w = P + lay.o[cons_w[s]] + (size_t)n * cons_ld[s];
      const float* zz = s_z + inst * NL;
#pragma unroll 8
      for (int j = 0; j < NL; ++j) acc = fmaf(w[j], zz[j], acc);
    }
    ctab[s * MAXOBJ * NL + i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// device helpers for the fused kernel
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st16(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}
// address of (row, col-block) in a blocked [128 x cols] tile
__device__ __forceinline__ uint8_t* rowblk(uint8_t* base, int row, int cb, int cols) {
  return base + (row >> 3) * (16 * cols) + cb * 128 + (row & 7) * 16;
}
// write 16 fp32 values as bf16 into col-blocks cb0, cb0+1 of a blocked tile row
__device__ __forceinline__ void store_row16(uint8_t* base, int row, int cols, int cb0, const float* v) {
  st16(rowblk(base, row, cb0, cols), pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
  st16(rowblk(base, row, cb0 + 1, cols), pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
}
__device__ __forceinline__ void store_blk8(uint8_t* base, int row, int cols, int cb, const float* v) {
  st16(rowblk(base, row, cb, cols), pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

struct Issue {   // MMA issue helpers (executed by one thread)
  uint32_t tmem;
  // forward: D[128 x N] (+)= A[128 x 16*nK] (K-major) * W[N x .]^T (K-major)
  __device__ __forceinline__ void fwd(int col, uint32_t aAddr, int aCols, int nK, uint32_t wAddr, int wCols, int N) const {
    const uint32_t id = idesc_bf16(128, N, 0, 0);
#pragma unroll
    for (int ks = 0; ks < nK; ++ks)
      mma_ss(tmem + col, smem_desc(aAddr + ks * 256, 128, 16 * aCols), smem_desc(wAddr + ks * 256, 128, 16 * wCols), id, ks > 0);
  }
  // data gradient: D[128 x Nout] (+)= G[128 x Nl] (K-major) * W[Nl x Nout] (MN-major view of the weight image)
  __device__ __forceinline__ void dgrad(int col, uint32_t gAddr, int gCols, int Nl, uint32_t wAddr, int wCols, int Nout, bool acc) const {
    const uint32_t id = idesc_bf16(128, Nout, 0, 1);
#pragma unroll
    for (int ks = 0; ks < Nl / 16; ++ks)
      mma_ss(tmem + col, smem_desc(gAddr + ks * 256, 128, 16 * gCols), smem_desc(wAddr + ks * 2 * (16 * wCols), 16 * wCols, 128), id,
             acc || ks > 0);
  }
  // weight gradient: D[128(feature rows) x Nl] += A^T (MN-major view of the activation tile) * G[128 pts x Nl] (MN-major)
  __device__ __forceinline__ void wgrad(int col, uint32_t aAddr, int aCols, uint32_t gAddr, int gCols, int Nl, bool acc) const {
    const uint32_t id = idesc_bf16(128, Nl, 1, 1);
#pragma unroll
    for (int ks = 0; ks < TP / 16; ++ks)
      mma_ss(tmem + col, smem_desc(aAddr + ks * 2 * (16 * aCols), 16 * aCols, 128), smem_desc(gAddr + ks * 2 * (16 * gCols), 16 * gCols, 128), id,
             acc || ks > 0);
  }
};

constexpr float PI_F = 3.14159265358979f;

__device__ __forceinline__ void pe_double(float& s, float& c) {
  const float s2 = 2.0f * s * c;
  c = fmaf(-2.0f * s, s, 1.0f);
  s = s2;
}
// sin/cos(pi * proj) of one direction, fast path (|r| <= 1 after range reduction)
__device__ __forceinline__ void pe_dir(const float* __restrict__ b3, const float t[3], float& s, float& c) {
  const float proj = fmaf(t[2], b3[2], fmaf(t[1], b3[1], t[0] * b3[0]));
  const float r = proj - 2.0f * rintf(0.5f * proj);
  __sincosf(PI_F * r, &s, &c);
}

// ---------------------------------------------------------------------------------------------------------
// the fused kernel.  grid = (ctas_per_cat, K), block = 288
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
fused_step_kernel(int K, int R, int S, const uint8_t* __restrict__ prep,
                  const float* __restrict__ origins, const float* __restrict__ dirs, const float* __restrict__ zs,
                  const float* __restrict__ gt_depth, const uint8_t* __restrict__ gt_rgb,
                  const uint8_t* __restrict__ mask_depth, const int32_t* __restrict__ inst_idx,
                  const int32_t* __restrict__ counts, float cs, float os,
                  float* __restrict__ raw, float* __restrict__ dp_scratch, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar, bar2;      // bar: per-stage commit; bar2: end-of-tile commit covering the trailing weight-gradient MMAs
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_red[4][24];

  const int k = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rpw = 32 / S, RPT = 4 * rpw;
  const int n_tiles = (R + RPT - 1) / RPT;

  // ---- one-time setup: prep image -> smem, zero the constant parts of the tiles, barriers, TMEM
  {
    const uint4* src = reinterpret_cast<const uint4*>(prep + (size_t)k * PREP_BYTES);
    uint4* dst = reinterpret_cast<uint4*>(sm);
    for (int i = tid; i < PREP_BYTES / 16; i += NT) dst[i] = src[i];
    const uint4 zero = make_uint4(0, 0, 0, 0);
    uint4* t0 = reinterpret_cast<uint4*>(sm + SM_R2XG);
    for (int i = tid; i < (SM_TOTAL - SM_R2XG) / 16; i += NT) t0[i] = zero;
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sbase = smem_u32(sm);

  // =========================================================================================================
  // warp 8: MMA issuer
  // =========================================================================================================
  if (warp == 8) {
    const Issue mma{tmem};
    bool first = true;
    const bool leader = lane == 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const bool accf = !first;
      // ---- forward
      nbar_sync(1, NT); if (leader) { tc_fence_after(); mma.fwd(TM_ACC, sbase + SM_R2XG + 4 * 128, 128, 6, sbase + SM_W + OFF_W1, W1_C, 32); mma_commit(&bar); }
      nbar_sync(1, NT); if (leader) { tc_fence_after(); mma.fwd(TM_ACC, sbase + SM_R1, 32, 2, sbase + SM_W + OFF_W2, W2_C, 32); mma_commit(&bar); }
      nbar_sync(1, NT); if (leader) { tc_fence_after(); mma.fwd(TM_ACC, sbase + SM_R2XG, 128, 8, sbase + SM_W + OFF_W3, W3_C, 32); mma_commit(&bar); }
      nbar_sync(1, NT); if (leader) { tc_fence_after(); mma.fwd(TM_ACC, sbase + SM_R3, 32, 2, sbase + SM_W + OFF_W4, W4_C, 32); mma_commit(&bar); }
      nbar_sync(1, NT); if (leader) { tc_fence_after(); mma.fwd(TM_ACC, sbase + SM_R4XC, 80, 5, sbase + SM_W + OFF_W6, W6_C, 48); mma_commit(&bar); }
      nbar_sync(1, NT); if (leader) { tc_fence_after(); mma.fwd(TM_ACC, sbase + SM_H1, 32, 2, sbase + SM_W + OFF_W7, W7_C, 32); mma_commit(&bar); }
      nbar_sync(1, NT); if (leader) { tc_fence_after(); mma.fwd(TM_ACC, sbase + SM_H2Z3, 80, 2, sbase + SM_W + OFF_W8, W8_C, 16); mma_commit(&bar); }
      // ---- backward: data gradient first (committed), weight gradients trail on the tensor pipe
      nbar_sync(1, NT);                                                                       // B8: G8 | G9 in GX
      if (leader) {
        tc_fence_after();
        mma.dgrad(TM_ACC, sbase + SM_GX, 16, 16, sbase + SM_W + OFF_W8, W8_C, 32, false);
        mma_commit(&bar);
        mma.wgrad(TM_WG89, sbase + SM_H2Z3, 80, sbase + SM_GX, 16, 16, accf);                          // rows 0..31: dW8^T
        mma.wgrad(TM_WG89, sbase + SM_H2Z3 + 4 * 128, 80, sbase + SM_GX + TP * 16 * 2, 16, 16, true); // rows 32..47: dW9^T
      }
      nbar_sync(1, NT);                                                                       // B7: G7 in GY
      if (leader) {
        tc_fence_after();
        mma.dgrad(TM_ACC, sbase + SM_GY, 32, 32, sbase + SM_W + OFF_W7, W7_C, 32, false);
        mma_commit(&bar);
        mma.wgrad(TM_WG7, sbase + SM_H1, 32, sbase + SM_GY, 32, 32, accf);
        mma.wgrad(TM_OH7, sbase + SM_OH, 64, sbase + SM_GY, 32, 32, accf);
      }
      nbar_sync(1, NT);                                                                       // B6': G6 (48) in GX
      if (leader) {
        tc_fence_after();
        mma.dgrad(TM_ACC, sbase + SM_GX, 48, 48, sbase + SM_W + OFF_W6, W6_C, 80, false);     // [d r4 | d xc]
        mma_commit(&bar);
        mma.wgrad(TM_WG6, sbase + SM_R4XC, 80, sbase + SM_GX, 48, 48, accf);
      }
      nbar_sync(1, NT);                                                                       // B4: G4 in GY
      if (leader) {
        tc_fence_after();
        mma.dgrad(TM_ACC, sbase + SM_GY, 32, 32, sbase + SM_W + OFF_W4, W4_C, 32, false);
        mma_commit(&bar);
        mma.wgrad(TM_WG4, sbase + SM_R3, 32, sbase + SM_GY, 32, 32, accf);
        mma.wgrad(TM_OH4, sbase + SM_OH, 64, sbase + SM_GY, 32, 32, accf);
      }
      nbar_sync(1, NT);                                                                       // B3: G3 in GX
      if (leader) {
        tc_fence_after();
        mma.dgrad(TM_ACC, sbase + SM_GX, 32, 32, sbase + SM_W + OFF_W3, W3_C, 128, false);    // cols 0..31 d r2, 32..127 d xg
        mma_commit(&bar);
        mma.wgrad(TM_WG3, sbase + SM_R2XG, 128, sbase + SM_GX, 32, 32, accf);
        mma.wgrad(TM_OH3, sbase + SM_OH, 64, sbase + SM_GX, 32, 32, accf);
      }
      nbar_sync(1, NT);                                                                       // B2: G2 in GY
      if (leader) {
        tc_fence_after();
        mma.dgrad(TM_ACC, sbase + SM_GY, 32, 32, sbase + SM_W + OFF_W2, W2_C, 32, false);
        mma_commit(&bar);
        mma.wgrad(TM_WG2, sbase + SM_R1, 32, sbase + SM_GY, 32, 32, accf);
        mma.wgrad(TM_OH2, sbase + SM_OH, 64, sbase + SM_GY, 32, 32, accf);
      }
      nbar_sync(1, NT);                                                                       // B1: G1 in GX
      if (leader) {
        tc_fence_after();
        mma.dgrad(TM_ACC + 32, sbase + SM_GX, 32, 32, sbase + SM_W + OFF_W1, W1_C, 96, true); // d xg +=
        mma_commit(&bar);
        mma.wgrad(TM_WG1, sbase + SM_R2XG + 4 * 128, 128, sbase + SM_GX, 32, 32, accf);
        mma_commit(&bar2);                        // everything of this tile (incl. all trailing weight-gradient MMAs) done
      }
      first = false;
      __syncwarp();
    }
    // join the epilogue warps for the flush / teardown barriers
    __syncthreads();      // (A) flush done
    return;
  }

  // =========================================================================================================
  // warps 0-7: epilogue threads, 2 per row
  // =========================================================================================================
  const int row = tid & 127, half = tid >> 7;
  const uint32_t tl = tmem_lane_base(tmem);
  const float* fcs = reinterpret_cast<const float*>(sm + SM_FC);
  const float* ctab = reinterpret_cast<const float*>(sm + SM_CTAB);
  uint32_t phase = 0, phase2 = 0;
  auto wait_mma = [&]() { mbar_wait(&bar, phase); phase ^= 1; tc_fence_after(); };
  auto publish = [&]() { fence_async_smem(); tc_fence_before(); nbar_arrive(1, NT); };   // operand tiles written + TMEM reads done
  // optional stage timestamps (development aid; dbg == nullptr in production): CTA 0, second tile, threads 0 and 128
  int dbg_i = 0;
  auto stamp = [&](int tile) {
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && tile == (int)gridDim.x && (tid == 0 || tid == 128) && dbg_i < 64)
      dbg[(tid >> 7) * 64 + dbg_i++] = clock64();
  };

  // loss normalisers; a term with an empty mask in ANY category is zero everywhere
  bool e0 = false, e1 = false, e2 = false;
  for (int kk = 0; kk < K; ++kk) { e0 |= counts[kk * 4 + 0] == 0; e1 |= counts[kk * 4 + 1] == 0; e2 |= counts[kk * 4 + 2] == 0; }
  const float iNd = e0 ? 0.f : 1.0f / ((float)counts[k * 4 + 0] + 1e-10f);
  const float iNo = e1 ? 0.f : cs / ((float)counts[k * 4 + 1] + 1e-10f);
  const float iNs = e2 ? 0.f : os / ((float)counts[k * 4 + 2] + 1e-10f);

  float l_d = 0.f, l_c = 0.f, l_o = 0.f;          // raw loss sums (this thread's rays; half 0 only)
  float db8[16], db9[3];
#pragma unroll
  for (int j = 0; j < 16; ++j) db8[j] = 0.f;
  db9[0] = db9[1] = db9[2] = 0.f;
  bool first = true;

  const int wq = warp & 3;                         // warp quadrant = 32-row group
  const int seg = lane / S, sl = lane - seg * S;
  const bool lane_used = lane < rpw * S;
  const int d0 = half * 11, nd = half ? 10 : 11;   // this thread's embedding directions
  const int hc = half * 16;                        // this thread's column offset inside a 32-wide layer

  // one 16-column epilogue: acc + bias -> relu -> mask bits + values
  auto relu16 = [&](const uint32_t* u, const float* bias, float* v, uint32_t& m) {
    m = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x = fmaxf(__uint_as_float(u[j]) + (bias ? bias[j] : 0.f), 0.f);
      v[j] = x; m |= (x > 0.f ? 1u : 0u) << j;
    }
  };
  auto load_bias16 = [&](const float* src, float* b) {      // issued BEFORE the mbarrier wait: shared-memory latency hides behind the MMA
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 x = s4[q]; b[4 * q] = x.x; b[4 * q + 1] = x.y; b[4 * q + 2] = x.z; b[4 * q + 3] = x.w; }
  };
  float b16[16];

  // ray / point data of a tile, loaded one tile ahead so the global-memory latency hides behind the previous tile
  struct RayData { float pz, o[3], d[3], gtd; int inst; uint32_t px; uint32_t md; };
  auto load_rays = [&](int tile, RayData& rd) {
    const int ray = tile * RPT + wq * rpw + seg;
    const bool a = lane_used && ray < R && tile < n_tiles;
    rd.pz = 0.f; rd.inst = 0; rd.px = 0; rd.md = 0; rd.gtd = 0.f;
    rd.o[0] = rd.o[1] = rd.o[2] = rd.d[0] = rd.d[1] = rd.d[2] = 0.f;
    if (a) {
      const size_t g = (size_t)k * R + ray;
      rd.pz = zs[g * S + sl];
#pragma unroll
      for (int i = 0; i < 3; ++i) { rd.o[i] = origins[g * 3 + i]; rd.d[i] = dirs[g * 3 + i]; }
      rd.inst = inst_idx[g];
      if (half == 0) { rd.px = *reinterpret_cast<const uint32_t*>(gt_rgb + g * 4); rd.md = mask_depth[g]; rd.gtd = gt_depth[g]; }
    }
  };
  RayData nxt;
  load_rays(blockIdx.x, nxt);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (!first) { mbar_wait(&bar2, phase2); phase2 ^= 1; tc_fence_after(); }   // previous tile's weight-gradient MMAs have read their tiles
    stamp(tile);
    // ------------------------------------------------------------------ ray / point data (prefetched)
    const int ray = tile * RPT + wq * rpw + seg;
    const bool act = lane_used && ray < R;
    const size_t gr = (size_t)k * R + (act ? ray : 0);
    const RayData cur = nxt;
    load_rays(tile + gridDim.x, nxt);
    const float pz = cur.pz;
    float t[3];
    {
      const float inv_scale = fcs[FC_SCALE];
#pragma unroll
      for (int a = 0; a < 3; ++a) t[a] = __fadd_rn(cur.o[a], __fmul_rn(cur.d[a], pz)) * inv_scale;
    }
    const int inst = min(max(cur.inst, 0), MAXOBJ - 1);  // never index outside the per-instance tables on bad input
    // ------------------------------------------------------------------ embedding -> xg (R2XG cols 32..127), xc (R4XC cols 32..79)
    {
      const float* peB = fcs + FC_PEB + d0 * 3;
      float xc_v[8];
      uint8_t* xg_base = sm + SM_R2XG;
      uint8_t* xc_base = sm + SM_R4XC;
      if (half == 0) st16(rowblk(xg_base, row, 4, 128), pack2(t[0], t[1]), pack2(t[2], 1.0f), 0u, 0u);   // t, ONE, pad
      float prev[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int dd = 0; dd < 11; ++dd) {
        float s = 0.f, c = 1.f, f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f, f5 = 0.f;
        if (dd < nd) {
          pe_dir(peB + dd * 3, t, s, c);
          f0 = s; pe_double(s, c); f1 = s; pe_double(s, c); f2 = s; pe_double(s, c); f3 = s;
          pe_double(s, c); f4 = s; pe_double(s, c); f5 = s;
        }
        if (dd & 1) {        // two directions x f0..3 = one 16-byte chunk of the xg tile
          if (half == 0 || dd < 10)
            st16(rowblk(xg_base, row, 4 + (half ? 7 : 1) + (dd >> 1), 128), pack2(prev[0], prev[1]), pack2(prev[2], prev[3]), pack2(f0, f1), pack2(f2, f3));
        } else {
          prev[0] = f0; prev[1] = f1; prev[2] = f2; prev[3] = f3;
          if (dd == 10 && half == 0)   // last (odd-one-out) direction of half 0 + 4 pad columns
            st16(rowblk(xg_base, row, 4 + 6, 128), pack2(f0, f1), pack2(f2, f3), 0u, 0u);
        }
        xc_v[(dd & 3) * 2] = f4; xc_v[(dd & 3) * 2 + 1] = f5;
        if ((dd & 3) == 3) store_blk8(xc_base, row, 80, 4 + half * 3 + (dd >> 2), xc_v);
        if (dd == 10) {      // chunk 2 of this half: dirs 8, 9, 10 (half 1: 10 is absent -> zeros), then ONE (half 0) / 0, 0
          xc_v[6] = half == 0 ? 1.0f : 0.0f; xc_v[7] = 0.f;
          store_blk8(xc_base, row, 80, 4 + half * 3 + 2, xc_v);
        }
      }
    }
    // one-hot(instance) row: 64 cols = 8 col-blocks, 4 per half
    {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cb = half * 4 + q;
        uint32_t w4[4] = {0, 0, 0, 0};
        if (act && (inst >> 3) == cb) {
          const int e = inst & 7;
          const uint32_t one = (e & 1) ? 0x3F800000u : 0x00003F80u;     // bf16(1.0) in the high / low half
          w4[0] = (e >> 1) == 0 ? one : 0u; w4[1] = (e >> 1) == 1 ? one : 0u;
          w4[2] = (e >> 1) == 2 ? one : 0u; w4[3] = (e >> 1) == 3 ? one : 0u;
        }
        st16(rowblk(sm + SM_OH, row, cb, 64), w4[0], w4[1], w4[2], w4[3]);
      }
    }
    stamp(tile); publish();

    uint32_t m1, m2, m3, m4, m6, m7, m8 = 0;     // relu masks (16 bits: this thread's columns)
    uint32_t u16[16];
    float v16[16];

    // ------------------------------------------------------------------ L1: r1 = relu(W1 [xg,1])
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
    relu16(u16, nullptr, v16, m1);
    store_row16(sm + SM_R1, row, 32, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ L2: r2 = relu(W2 r1 + c2[inst])
    load_bias16(ctab + (0 * MAXOBJ + inst) * NL + hc, b16);
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
    relu16(u16, b16, v16, m2);
    store_row16(sm + SM_R2XG, row, 128, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ L3: r3 = relu(W3 [r2, xg] + c3[inst])
    load_bias16(ctab + (1 * MAXOBJ + inst) * NL + hc, b16);
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
    relu16(u16, b16, v16, m3);
    store_row16(sm + SM_R3, row, 32, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ L4: r4 = relu(W4 r3 + c4[inst])
    load_bias16(ctab + (2 * MAXOBJ + inst) * NL + hc, b16);
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
    relu16(u16, b16, v16, m4);
    store_row16(sm + SM_R4XC, row, 80, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ L6': h1 = relu(W6' [r4, xc, 1]); sigma = 10 * col 32
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16);
    uint32_t sg_u[8];
    if (half == 0) tmem_ld8(tl + TM_ACC + 32, sg_u);
    tmem_ld_wait();
    const float sigma = half == 0 ? 10.0f * __uint_as_float(sg_u[0]) : 0.f;     // model.py:71,75
    relu16(u16, nullptr, v16, m6);
    store_row16(sm + SM_H1, row, 32, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ L7: h2 = relu(W7 h1 + c7[inst])
    load_bias16(ctab + (3 * MAXOBJ + inst) * NL + hc, b16);
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
    relu16(u16, b16, v16, m7);
    store_row16(sm + SM_H2Z3, row, 80, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ L8 + L9 + compositing + B9: half 0 only (16-wide layers)
    if (half == 0) load_bias16(fcs + FC_B8, b16);
    stamp(tile); wait_mma(); stamp(tile);
    float d_sig = 0.f;
    if (half == 0) {
      tmem_ld16(tl + TM_ACC, u16); tmem_ld_wait();
      float h3[16];
      relu16(u16, b16, h3, m8);
      store_row16(sm + SM_H2Z3, row, 80, 8, h3);
      stamp(tile);
      float col[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float a = fcs[FC_B9 + j];
#pragma unroll
        for (int q = 0; q < 16; ++q) a = fmaf(fcs[FC_W9 + j * 16 + q], h3[q], a);
        col[j] = __frcp_rn(1.0f + __expf(-a));                                // model.py:54 (sigmoid)
      }
      stamp(tile);
      // ---------------- compositing + losses + their gradients (warp segmented scans)
      float d_rgbpre[3] = {0.f, 0.f, 0.f};
      {
        const float occ = act ? __frcp_rn(1.0f + __expf(-sigma)) : 0.f;       // render_rays.py:3-7
        const float fr = act ? __fadd_rn(__fsub_rn(1.0f, occ), 1e-10f) : 1.0f;   // render_rays.py:28
        float incl = fr;
        for (int o = 1; o < S; o <<= 1) { const float n = __shfl_up_sync(0xffffffffu, incl, o); if (sl >= o) incl *= n; }
        float T = __shfl_up_sync(0xffffffffu, incl, 1);
        if (sl == 0) T = 1.0f;
        const float w = occ * T;                                             // render_rays.py:33
        const int last = min(seg * S + S - 1, 31);
        // five segment sums scanned together (independent shuffle chains interleave)
        float a0 = w * pz, a1 = w, a2 = w * col[0], a3 = w * col[1], a4 = w * col[2];
        for (int o = 1; o < S; o <<= 1) {
          const float n0 = __shfl_up_sync(0xffffffffu, a0, o), n1 = __shfl_up_sync(0xffffffffu, a1, o);
          const float n2 = __shfl_up_sync(0xffffffffu, a2, o), n3 = __shfl_up_sync(0xffffffffu, a3, o);
          const float n4 = __shfl_up_sync(0xffffffffu, a4, o);
          if (sl >= o) { a0 += n0; a1 += n1; a2 += n2; a3 += n3; a4 += n4; }
        }
        const float D = __shfl_sync(0xffffffffu, a0, last), O = __shfl_sync(0xffffffffu, a1, last);
        const float C0 = __shfl_sync(0xffffffffu, a2, last), C1 = __shfl_sync(0xffffffffu, a3, last);
        const float C2 = __shfl_sync(0xffffffffu, a4, last);
        const float dz = pz - D;
        float av = w * dz * dz;                                              // loss.py:45-46 (detached)
        for (int o = 1; o < S; o <<= 1) { const float n = __shfl_up_sync(0xffffffffu, av, o); if (sl >= o) av += n; }
        const float var = __shfl_sync(0xffffffffu, av, last);
        float gD = 0.f, gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gO = 0.f;
        if (act) {
          const uint32_t px = cur.px;
          const uint32_t lab = px >> 24;
          const float mo = (lab != 0) ? 1.f : 0.f, ms = (lab != 2) ? 1.f : 0.f;
          const float md = (cur.md != 0 && lab != 0) ? 1.f : 0.f;
          const float g0 = (float)(px & 255u) * (1.0f / 255.0f), g1 = (float)((px >> 8) & 255u) * (1.0f / 255.0f),
                      g2 = (float)((px >> 16) & 255u) * (1.0f / 255.0f);
          const float info = __frcp_rn(sqrtf(var) + 1e-4f);
          const float rD = D - cur.gtd, r0 = C0 - g0, r1 = C1 - g1, r2 = C2 - g2, rO = O - mo;
          if (sl == 0) {
            l_d += fabsf(rD) * md * info;
            l_c += (fabsf(r0) + fabsf(r1) + fabsf(r2)) * mo;
            l_o += fabsf(rO) * ms;
          }
          auto sgn = [](float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); };
          gD = sgn(rD) * md * info * iNd;
          gC0 = sgn(r0) * mo * iNo; gC1 = sgn(r1) * mo * iNo; gC2 = sgn(r2) * mo * iNo;
          gO = sgn(rO) * ms * iNs;
        }
        const float Gi = gD * pz + gC0 * col[0] + gC1 * col[1] + gC2 * col[2] + gO;
        const float Gw = act ? Gi * w : 0.f;
        float sfx = Gw;
        for (int o = 1; o < S; o <<= 1) { const float n = __shfl_down_sync(0xffffffffu, sfx, o); if (sl + o < S) sfx += n; }
        if (act) {
          const float docc = Gi * T - (sfx - Gw) * __frcp_rn(fr);
          d_sig = 10.0f * docc * occ * (1.0f - occ);                          // into the raw sigma column (x10 scale)
          d_rgbpre[0] = w * gC0 * col[0] * (1.0f - col[0]);
          d_rgbpre[1] = w * gC1 * col[1] * (1.0f - col[1]);
          d_rgbpre[2] = w * gC2 * col[2] * (1.0f - col[2]);
        }
      }
      stamp(tile);
      // ---------------- B9 (CUDA cores): d h3 -> G8 ; G9 = d rgb_pre
      float g8[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        float a = fcs[FC_W9 + q] * d_rgbpre[0];
        a = fmaf(fcs[FC_W9 + 16 + q], d_rgbpre[1], a);
        a = fmaf(fcs[FC_W9 + 32 + q], d_rgbpre[2], a);
        g8[q] = ((m8 >> q) & 1u) ? a : 0.f;
        db8[q] += g8[q];
      }
      db9[0] += d_rgbpre[0]; db9[1] += d_rgbpre[1]; db9[2] += d_rgbpre[2];
      uint8_t* G8 = sm + SM_GX; uint8_t* G9 = sm + SM_GX + TP * 16 * 2;
      store_row16(G8, row, 16, 0, g8);
      st16(rowblk(G9, row, 0, 16), pack2(d_rgbpre[0], d_rgbpre[1]), pack2(d_rgbpre[2], 0.f), 0u, 0u);
      st16(rowblk(G9, row, 1, 16), 0u, 0u, 0u, 0u);
    }
    stamp(tile); publish();
    // ------------------------------------------------------------------ B8: d h2 = G8 W8 -> G7 (GY)
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v16[j] = ((m7 >> j) & 1u) ? __uint_as_float(u16[j]) : 0.f;
    store_row16(sm + SM_GY, row, 32, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ B7: d h1 = G7 W7 -> G6 (GX, 48 cols; col 32 = d raw sigma)
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v16[j] = ((m6 >> j) & 1u) ? __uint_as_float(u16[j]) : 0.f;
    store_row16(sm + SM_GX, row, 48, half * 2, v16);
    if (half == 0) st16(rowblk(sm + SM_GX, row, 4, 48), pack2(d_sig, 0.f), 0u, 0u, 0u);
    else st16(rowblk(sm + SM_GX, row, 5, 48), 0u, 0u, 0u, 0u);
    stamp(tile); publish();
    // ------------------------------------------------------------------ B6': [d r4 | d xc] = G6 W6' -> G4 (GY) ; park d xc
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v16[j] = ((m4 >> j) & 1u) ? __uint_as_float(u16[j]) : 0.f;
    store_row16(sm + SM_GY, row, 32, half * 2, v16);
    {
      // d xc: this half's 24 columns (tile cols 24*half .. +24) -> bf16 into the dead h2/h3 tile:
      // half 0 -> col-blocks 0,1,2 ; half 1 -> col-blocks 3, 8, 9.  The zero block (cols 32..63) stays untouched.
      uint32_t x8[8];
      tmem_ld16(tl + TM_ACC + 32 + 24 * half, u16);
      tmem_ld8(tl + TM_ACC + 32 + 24 * half + 16, x8);
      tmem_ld_wait();
      const float* xa = reinterpret_cast<const float*>(u16);
      const float* xb = reinterpret_cast<const float*>(x8);
      const int b0 = half ? 3 : 0, b1 = half ? 8 : 1, b2 = half ? 9 : 2;
      store_blk8(sm + SM_H2Z3, row, 80, b0, xa);
      store_blk8(sm + SM_H2Z3, row, 80, b1, xa + 8);
      store_blk8(sm + SM_H2Z3, row, 80, b2, xb);
    }
    stamp(tile); publish();
    // ------------------------------------------------------------------ B4: d r3 = G4 W4 -> G3 (GX)
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v16[j] = ((m3 >> j) & 1u) ? __uint_as_float(u16[j]) : 0.f;
    store_row16(sm + SM_GX, row, 32, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ B3: [d r2 | d xg] = G3 W3 -> G2 (GY)
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v16[j] = ((m2 >> j) & 1u) ? __uint_as_float(u16[j]) : 0.f;
    store_row16(sm + SM_GY, row, 32, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ B2: d r1 = G2 W2 -> G1 (GX)
    stamp(tile); wait_mma(); stamp(tile);
    tmem_ld16(tl + TM_ACC + hc, u16); tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v16[j] = ((m1 >> j) & 1u) ? __uint_as_float(u16[j]) : 0.f;
    store_row16(sm + SM_GX, row, 32, half * 2, v16);
    stamp(tile); publish();
    // ------------------------------------------------------------------ B1: d xg += G1 W1 ; embedding backward
    stamp(tile); wait_mma(); stamp(tile);
    {
      // this half's xg gradient columns: half 0 -> xg cols 8..55 (48), half 1 -> 56..95 (40); TMEM col = TM_ACC + 32 + xg col
      uint32_t ga[32], gb[16];
      const int c0 = half ? 56 : 8;
      tmem_ld32(tl + TM_ACC + 32 + c0, ga);
      if (half == 0) tmem_ld16(tl + TM_ACC + 32 + c0 + 32, gb); else tmem_ld8(tl + TM_ACC + 32 + c0 + 32, gb);
      tmem_ld_wait();
      // parked d xc of this half: 3 col-blocks
      const int b0 = half ? 3 : 0, b1 = half ? 8 : 1, b2 = half ? 9 : 2;
      const uint4 q0 = *reinterpret_cast<const uint4*>(rowblk(sm + SM_H2Z3, row, b0, 80));
      const uint4 q1 = *reinterpret_cast<const uint4*>(rowblk(sm + SM_H2Z3, row, b1, 80));
      const uint4 q2 = *reinterpret_cast<const uint4*>(rowblk(sm + SM_H2Z3, row, b2, 80));
      const uint32_t xq[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
      const float* peB = fcs + FC_PEB + d0 * 3;
      float dp[11];
#pragma unroll
      for (int dd = 0; dd < 11; ++dd) {
        float acc = 0.f;
        if (dd < nd) {
          float s, c;
          pe_dir(peB + dd * 3, t, s, c);
          // xg columns of direction dd: 4*dd + f (f = 0..3) relative to c0 ; xc columns: 2*dd + (f-4) relative to 24*half
#pragma unroll
          for (int f = 0; f < 6; ++f) {
            if (f > 0) pe_double(s, c);
            float g;
            if (f < 4) {
              const int j = 4 * dd + f;
              g = __uint_as_float(j < 32 ? ga[j] : gb[j - 32]);
            } else {
              const int j = 2 * dd + (f - 4);
              g = __uint_as_float((j & 1) ? (xq[j >> 1] & 0xFFFF0000u) : (xq[j >> 1] << 16));
            }
            acc = fmaf(g * (PI_F * (float)(1 << f)), c, acc);
          }
        }
        dp[dd] = acc;
      }
      if (act) {
        // scratch row (24 floats): [0..10] dp of dirs 0..10, [11] t0, [12..21] dp of dirs 11..20, [22] t1, [23] t2
        float* out = dp_scratch + (gr * S + sl) * 24 + half * 12;
        float4* o4 = reinterpret_cast<float4*>(out);
        if (half == 0) {
          o4[0] = make_float4(dp[0], dp[1], dp[2], dp[3]);
          o4[1] = make_float4(dp[4], dp[5], dp[6], dp[7]);
          o4[2] = make_float4(dp[8], dp[9], dp[10], t[0]);
        } else {
          o4[0] = make_float4(dp[0], dp[1], dp[2], dp[3]);
          o4[1] = make_float4(dp[4], dp[5], dp[6], dp[7]);
          o4[2] = make_float4(dp[8], dp[9], t[1], t[2]);
        }
      }
    }
    stamp(tile);
    tc_fence_before();      // TMEM reads of this tile are ordered before the next tile's first publish/arrive
    first = false;
  }

  // ---- flush: TMEM accumulators -> raw[k] (fp32 vector atomics; several CTAs share a category). half 0 threads only.
  float* rk = raw + (size_t)k * RAW_FLOATS;
  if (!first && half == 0) {
    mbar_wait(&bar2, phase2);
    tc_fence_after();
    auto flush32 = [&](int col, int roff, int rows, int ld, int c0) {
      uint32_t u[32];
      tmem_ld32(tl + col + c0, u); tmem_ld_wait();
      if (row < rows) {
        float4* dst = reinterpret_cast<float4*>(rk + roff + row * ld + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          atomicAdd(dst + j, make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                                         __uint_as_float(u[4 * j + 3])));
      }
    };
    auto flush16 = [&](int col, int roff, int rows, int ld, int c0) {
      uint32_t u[16];
      tmem_ld16(tl + col + c0, u); tmem_ld_wait();
      if (row < rows) {
        float4* dst = reinterpret_cast<float4*>(rk + roff + row * ld + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          atomicAdd(dst + j, make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                                         __uint_as_float(u[4 * j + 3])));
      }
    };
    flush32(TM_WG1, RW1, 96, 32, 0);
    flush32(TM_WG2, RW2, 32, 32, 0);
    flush32(TM_WG3, RW3, 128, 32, 0);
    flush32(TM_WG4, RW4, 32, 32, 0);
    flush32(TM_WG6, RW6, 80, 48, 0);
    flush16(TM_WG6, RW6, 80, 48, 32);
    flush32(TM_WG7, RW7, 32, 32, 0);
    flush16(TM_WG89, RW89, 48, 16, 0);
    flush32(TM_OH2, ROH2, 64, 32, 0);
    flush32(TM_OH3, ROH3, 64, 32, 0);
    flush32(TM_OH4, ROH4, 64, 32, 0);
    flush32(TM_OH7, ROH7, 64, 32, 0);
  }
  // per-thread register accumulators (half 0): block reduce then one atomic each
  if (half == 0) {
    float vals[22];
#pragma unroll
    for (int j = 0; j < 16; ++j) vals[j] = db8[j];
    vals[16] = db9[0]; vals[17] = db9[1]; vals[18] = db9[2]; vals[19] = l_d; vals[20] = l_c; vals[21] = l_o;
#pragma unroll
    for (int j = 0; j < 22; ++j) { const float s = warp_sum(vals[j]); if (lane == 0) s_red[warp][j] = s; }
  }
  tc_fence_before();
  __syncthreads();        // (A) matches the MMA warp's barrier: flush + partial sums done
  if (tid < 22) {
    const float s = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];
    const int dst = tid < 16 ? RDB8 + tid : (tid < 19 ? RDB9 + (tid - 16) : RLOSS + (tid - 19));
    atomicAdd(rk + dst, s);
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// dB[d][a] = sum_p dproj[p][d] * t[p][a]  from the scratch rows {dp[21], t[3]}
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fused_pe_grad_kernel(size_t P, const float* __restrict__ scratch, float* __restrict__ g_B) {
  const int k = blockIdx.y;
  float acc[63];
#pragma unroll
  for (int j = 0; j < 63; ++j) acc[j] = 0.f;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float4* r = reinterpret_cast<const float4*>(scratch + ((size_t)k * P + p) * 24);
    float v[24];
#pragma unroll
    for (int q = 0; q < 6; ++q) { const float4 x = r[q]; v[q * 4] = x.x; v[q * 4 + 1] = x.y; v[q * 4 + 2] = x.z; v[q * 4 + 3] = x.w; }
#pragma unroll
    for (int d = 0; d < 21; ++d) {
      const float g = v[d < 11 ? d : d + 1];          // row layout: [0..10] dirs 0..10, [11] t0, [12..21] dirs 11..20, [22] t1, [23] t2
      acc[d * 3 + 0] = fmaf(g, v[11], acc[d * 3 + 0]);
      acc[d * 3 + 1] = fmaf(g, v[22], acc[d * 3 + 1]);
      acc[d * 3 + 2] = fmaf(g, v[23], acc[d * 3 + 2]);
    }
  }
  __shared__ float s[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 63; ++j) { const float x = warp_sum(acc[j]); if (lane == 0) s[warp][j] = x; }
  __syncthreads();
  if (threadIdx.x < 63) {
    float x = 0.f;
    for (int w = 0; w < 8; ++w) x += s[w][threadIdx.x];
    atomicAdd(&g_B[(size_t)k * 63 + threadIdx.x], x);
  }
}

// ---------------------------------------------------------------------------------------------------------
// finalize: raw accumulators -> reference-shaped gradients (one CTA per category) + code regulariser
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fused_finalize_kernel(int L, int n_obj_max, const float* __restrict__ params, const float* __restrict__ pe_scale,
                      const float* __restrict__ shape_codes, const float* __restrict__ tex_codes,
                      const int32_t* __restrict__ n_obj_arr, const float* __restrict__ raw, const float* __restrict__ zlat,
                      float reg_scaling, int add_reg, float* __restrict__ g_params, float* __restrict__ g_B,
                      float* __restrict__ g_shape, float* __restrict__ g_tex, float* __restrict__ reg_out) {
  // grid = (K, FIN_PARTS): every CTA recomputes the small per-instance terms, then takes a slice of the outputs
  const int k = blockIdx.x, tid = threadIdx.x;
  const int gtid = blockIdx.y * 256 + tid, gstride = 256 * gridDim.y;
  const PLay lay(L);
  const float* P = params + (size_t)k * lay.total;
  float* G = g_params + (size_t)k * lay.total;
  const float* rk = raw + (size_t)k * RAW_FLOATS;
  const float* z = zlat + (size_t)k * 4 * MAXOBJ * NL;
  const int n_obj = min(n_obj_arr[k], MAXOBJ);
  __shared__ float s_dpre[4][MAXOBJ][NL];      // relu-masked gradient wrt the latent pre-activations
  __shared__ float s_norm[2][MAXOBJ];
  (void)pe_scale; (void)g_B;                   // g_B is produced by fused_pe_grad_kernel (t = x/scale already applied)
  const int cons_w[4] = {4, 10, 8, 20};
  const int cons_ld[4] = {32, 119, 32, 32};
  const int rw[4] = {RW2, RW3, RW4, RW7};
  const int roh[4] = {ROH2, ROH3, ROH4, ROH7};
  const int lat_w[4] = {2, 12, 6, 22};
  const int warp = tid >> 5, lane = tid & 31;

  // --- per-instance terms: d z[inst][c] = sum_n dC[inst][n] W_cons[n][c], masked by relu(z) ; code norms
  for (int i = tid; i < 4 * n_obj * NL; i += 256) {
    const int s = i / (n_obj * NL), inst = (i / NL) % n_obj, c = i % NL;
    float a = 0.f;
    if (z[(s * MAXOBJ + inst) * NL + c] > 0.f) {
      const int ld = cons_ld[s];
#pragma unroll 8
      for (int n = 0; n < 32; ++n) a = fmaf(rk[roh[s] + inst * 32 + n], P[lay.o[cons_w[s]] + (size_t)n * ld + c], a);
    }
    s_dpre[s][inst][c] = a;
  }
  for (int o = warp; o < 2 * n_obj; o += 8) {
    const int which = o / n_obj, inst = o % n_obj;
    const float* code = (which ? tex_codes : shape_codes) + ((size_t)k * n_obj_max + inst) * L;
    float a = 0.f;
    for (int j = lane; j < L; j += 32) a = fmaf(code[j], code[j], a);
    a = warp_sum(a);
    if (lane == 0) s_norm[which][inst] = sqrtf(a);
  }
  __syncthreads();

  // --- plain layers (transposed copies out of the raw accumulators; xg / xc lanes follow the tile column order)
  for (int i = gtid; i < 32 * 87; i += gstride) { const int n = i / 87, e = i % 87; G[lay.o[0] + i] = rk[RW1 + xg_col(e) * 32 + n]; }
  for (int i = gtid; i < 16 * 32; i += gstride) { const int n = i / 32, c = i % 32; G[lay.o[24] + i] = rk[RW89 + c * 16 + n]; }
  for (int i = gtid; i < 3 * 16; i += gstride) { const int j = i / 16, c = i % 16; G[lay.o[26] + i] = rk[RW89 + (32 + c) * 16 + j]; }
  // encoding_viewdir / encoding_shape / sigma: the kernel trained the folded weights
  //   W6a' = Wvd[:, :32] W5, b6' = Wvd[:, :32] b5 + b_vd, wsig' = wsig W5, bsig' = wsig b5 + bsig.
  // With Ga'[n][k] = dL/dW6a'[n][k], gb6'[n], gws'[k], gbs':
  //   dWvd[n][j<32] = sum_k Ga'[n][k] W5[j][k] + gb6'[n] b5[j]         dWvd[n][32+e] direct
  //   dW5[j][k]     = sum_n Wvd[n][j] Ga'[n][k] + wsig[j] gws'[k]      db5[j] = sum_n Wvd[n][j] gb6'[n] + wsig[j] gbs'
  //   dwsig[j]      = sum_k gws'[k] W5[j][k] + gbs' b5[j]              db_vd = gb6', dbsig = gbs'
  {
    const int BL = 32 + XC_BIAS_COL;                                   // lane of the ones column in the [r4 | xc] tile
    const float gbs = rk[RW6 + BL * 48 + 32];
    for (int i = gtid; i < 32 * 74; i += gstride) {
      const int n = i / 74, j = i % 74;
      float a;
      if (j < 32) {
        a = rk[RW6 + BL * 48 + n] * P[lay.o[15] + j];
        for (int kk = 0; kk < 32; ++kk) a = fmaf(rk[RW6 + kk * 48 + n], P[lay.o[14] + j * 32 + kk], a);
      } else {
        a = rk[RW6 + (32 + xc_col(87 + (j - 32))) * 48 + n];
      }
      G[lay.o[18] + i] = a;
    }
    for (int i = gtid; i < 32 * 32; i += gstride) {                    // dW5[j][k]
      const int j = i / 32, kk = i % 32;
      float a = P[lay.o[16] + j] * rk[RW6 + kk * 48 + 32];
      for (int n = 0; n < 32; ++n) a = fmaf(P[lay.o[18] + (size_t)n * 74 + j], rk[RW6 + kk * 48 + n], a);
      G[lay.o[14] + i] = a;
    }
    if (blockIdx.y == 0) {
      for (int j = tid; j < 32; j += 256) {
        float b5 = P[lay.o[16] + j] * gbs;                             // db5[j]
        for (int n = 0; n < 32; ++n) b5 = fmaf(P[lay.o[18] + (size_t)n * 74 + j], rk[RW6 + BL * 48 + n], b5);
        G[lay.o[15] + j] = b5;
        float ws = gbs * P[lay.o[15] + j];                             // dwsig[j]
        for (int kk = 0; kk < 32; ++kk) ws = fmaf(rk[RW6 + kk * 48 + 32], P[lay.o[14] + j * 32 + kk], ws);
        G[lay.o[16] + j] = ws;
        G[lay.o[19] + j] = rk[RW6 + BL * 48 + j];                      // b_vd
        G[lay.o[1] + j] = rk[RW1 + XG_BIAS_COL * 32 + j];              // b_xyz (row of the ones column)
      }
      if (tid == 0) G[lay.o[17]] = gbs;                                // b_sigma
      for (int n = tid; n < 16; n += 256) G[lay.o[25] + n] = rk[RDB8 + n];
      for (int j = tid; j < 3; j += 256) G[lay.o[27] + j] = rk[RDB9 + j];
    }
  }
  // --- layers with per-instance bias c = W[:, :32] z + b : slots (zsl1->s1), (zcl->cat), (zsl2->s2), (ztl->t1)
  for (int s = 0; s < 4; ++s) {
    const int ld = cons_ld[s];
    for (int i = gtid; i < 32 * ld; i += gstride) {
      const int n = i / ld, c = i % ld;
      float a = rk[rw[s] + (c < 32 ? c : 32 + xg_col(c - 32)) * 32 + n];     // only the cat layer (s = 1) has c >= 32
      if (c < 32)
        for (int inst = 0; inst < n_obj; ++inst) a = fmaf(rk[roh[s] + inst * 32 + n], z[(s * MAXOBJ + inst) * NL + c], a);
      G[lay.o[cons_w[s]] + i] = a;
    }
    if (blockIdx.y == 0)
      for (int n = tid; n < 32; n += 256) {
        float a = 0.f;
        for (int inst = 0; inst < n_obj; ++inst) a += rk[roh[s] + inst * 32 + n];
        G[lay.o[cons_w[s] + 1] + n] = a;
      }
    // latent layer: dW_lat[c][j] = sum_inst dpre[inst][c] code[inst][j]; db_lat[c] = sum_inst dpre[inst][c]
    const float* codes = (s == 3 ? tex_codes : shape_codes) + (size_t)k * n_obj_max * L;
    for (int i = gtid; i < 32 * L; i += gstride) {
      const int c = i / L, j = i % L;
      float a = 0.f;
      for (int inst = 0; inst < n_obj; ++inst) a = fmaf(s_dpre[s][inst][c], codes[(size_t)inst * L + j], a);
      G[lay.o[lat_w[s]] + i] = a;
    }
    if (blockIdx.y == 0)
      for (int c = tid; c < 32; c += 256) {
        float a = 0.f;
        for (int inst = 0; inst < n_obj; ++inst) a += s_dpre[s][inst][c];
        G[lay.o[lat_w[s] + 1] + c] = a;
      }
  }
  // --- code tables: d code[inst][j] = sum_s sum_c dpre[s][inst][c] W_lat_s[c][j]  (+ regulariser)
  for (int i = gtid; i < n_obj_max * L; i += gstride) {
    const int inst = i / L, j = i % L;
    float gs = 0.f, gt = 0.f;
    if (inst < n_obj) {
#pragma unroll 4
      for (int c = 0; c < 32; ++c) {
        gs = fmaf(s_dpre[0][inst][c], P[lay.o[2] + (size_t)c * L + j], gs);
        gs = fmaf(s_dpre[1][inst][c], P[lay.o[12] + (size_t)c * L + j], gs);
        gs = fmaf(s_dpre[2][inst][c], P[lay.o[6] + (size_t)c * L + j], gs);
        gt = fmaf(s_dpre[3][inst][c], P[lay.o[22] + (size_t)c * L + j], gt);
      }
      if (add_reg && n_obj > 1) {                                            // loss.py:12, train.py:165-167
        const float ns = s_norm[0][inst], nt = s_norm[1][inst];
        if (ns > 0.f) gs += reg_scaling * shape_codes[((size_t)k * n_obj_max + inst) * L + j] / ns;
        if (nt > 0.f) gt += reg_scaling * tex_codes[((size_t)k * n_obj_max + inst) * L + j] / nt;
      }
    }
    g_shape[(size_t)k * n_obj_max * L + i] = gs;
    g_tex[(size_t)k * n_obj_max * L + i] = gt;
  }
  if (blockIdx.y == 0 && tid == 0) {
    float rs = 0.f, rt = 0.f;
    if (n_obj > 1) for (int inst = 0; inst < n_obj; ++inst) { rs += s_norm[0][inst]; rt += s_norm[1][inst]; }
    reg_out[k * 2 + 0] = rs; reg_out[k * 2 + 1] = rt;
  }
}

// out_loss = {loss, e_depth, e_obj, e_sem, depth[K], color[K], opacity[K], reg_shape[K], reg_tex[K]}
__global__ void fused_loss_kernel(int K, const float* __restrict__ raw, const int32_t* __restrict__ counts,
                                  const float* __restrict__ reg, float cs, float os, float reg_scaling, int add_reg,
                                  float* __restrict__ out_loss) {
  // one warp; lane per category (strided)
  const int lane = threadIdx.x;
  int f0 = 0, f1 = 0, f2 = 0;
  for (int k = lane; k < K; k += 32) { f0 |= counts[k * 4 + 0] == 0; f1 |= counts[k * 4 + 1] == 0; f2 |= counts[k * 4 + 2] == 0; }
  const bool e0 = __any_sync(0xffffffffu, f0), e1 = __any_sync(0xffffffffu, f1), e2 = __any_sync(0xffffffffu, f2);
  float total = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float* l = raw + (size_t)k * RAW_FLOATS + RLOSS;
    const float d = e0 ? 0.f : l[0] / ((float)counts[k * 4 + 0] + 1e-10f);
    const float c = e1 ? 0.f : l[1] / ((float)counts[k * 4 + 1] + 1e-10f);
    const float o = e2 ? 0.f : l[2] / ((float)counts[k * 4 + 2] + 1e-10f);
    out_loss[4 + k] = d; out_loss[4 + K + k] = c; out_loss[4 + 2 * K + k] = o;
    out_loss[4 + 3 * K + k] = reg[k * 2]; out_loss[4 + 4 * K + k] = reg[k * 2 + 1];
    total += d + c * cs + o * os;
    if (add_reg) total += reg_scaling * (reg[k * 2] + reg[k * 2 + 1]);
  }
  total = warp_sum(total);
  if (lane == 0) { out_loss[0] = total; out_loss[1] = e0; out_loss[2] = e1; out_loss[3] = e2; }
}

struct FusedWs {
  size_t prep, zlat, raw, reg, scratch, total;
  FusedWs(size_t K, size_t R, size_t S) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    prep = take(K * PREP_BYTES);
    zlat = take(K * 4 * MAXOBJ * NL * 4);
    raw = take(K * RAW_FLOATS * 4);
    reg = take(K * 2 * 4);
    scratch = take(K * R * S * 24 * 4);
    total = o;
  }
};

}  // namespace cnf

using namespace cnf;

static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
static thread_local long long* g_dbg = nullptr;
extern "C" int cnf_set_debug_buffer(void* p) { g_dbg = reinterpret_cast<long long*>(p); return CNF_OK; }
extern "C" int cnf_set_kernel_events(void* ev_start, void* ev_stop) {
  g_ev_start = reinterpret_cast<cudaEvent_t>(ev_start);
  g_ev_stop = reinterpret_cast<cudaEvent_t>(ev_stop);
  return CNF_OK;
}

extern "C" size_t cnf_train_step_workspace_bytes(int K, int R, int S, int L, int n_obj_max) {
  (void)L; (void)n_obj_max;
  return FusedWs(K, R, S).total;
}

extern "C" int cnf_train_step(int K, int R, int S, int L, int n_obj_max, const float* params, const float* pe_B,
                              const float* pe_scale, const float* shape_codes, const float* tex_codes,
                              const int32_t* n_obj, const float* origins, const float* dirs, const float* z,
                              const float* gt_depth, const uint8_t* gt_rgb, const uint8_t* mask_depth,
                              const int32_t* inst, const int32_t* counts, float cs, float os, float reg_scaling,
                              int flags, float* g_params, float* g_B, float* g_shape, float* g_tex, float* out_loss,
                              void* ws, size_t ws_bytes, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && R > 0 && L > 0, "cnf_train_step: bad sizes K=%d R=%d L=%d", K, R, L);
  if (S < 1 || S > 32) { set_error("cnf_train_step: samples per ray S=%d outside 1..32 (one ray per warp segment)", S); return CNF_ERR_UNSUPPORTED; }
  if (n_obj_max < 1 || n_obj_max > MAXOBJ) { set_error("cnf_train_step: n_obj_max=%d outside 1..%d", n_obj_max, MAXOBJ); return CNF_ERR_UNSUPPORTED; }
  CNF_REQUIRE(params && pe_B && pe_scale && shape_codes && tex_codes && n_obj && origins && dirs && z && gt_depth && gt_rgb &&
                  mask_depth && inst && counts && g_params && g_B && g_shape && g_tex && out_loss && ws,
              "cnf_train_step: null pointer");
  const FusedWs wl(K, R, S);
  if (ws_bytes < wl.total) { set_error("cnf_train_step: workspace too small (%zu < %zu)", ws_bytes, wl.total); return CNF_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  uint8_t* prep = w8 + wl.prep;
  float* zlat = reinterpret_cast<float*>(w8 + wl.zlat);
  float* raw = reinterpret_cast<float*>(w8 + wl.raw);
  float* reg = reinterpret_cast<float*>(w8 + wl.reg);
  float* scratch = reinterpret_cast<float*>(w8 + wl.scratch);
  static bool attr_set = false;
  if (!attr_set) {
    CNF_CUDA(cudaFuncSetAttribute(fused_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    attr_set = true;
  }
  int dev = 0, n_sm = 148;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int rpw = 32 / S, RPT = 4 * rpw, n_tiles = (R + RPT - 1) / RPT;
  int cpc = n_sm / K; if (cpc < 1) cpc = 1; if (cpc > n_tiles) cpc = n_tiles;
  CNF_CUDA(cudaMemsetAsync(raw, 0, (size_t)K * RAW_FLOATS * 4, st));
  CNF_CUDA(cudaMemsetAsync(g_B, 0, (size_t)K * 63 * 4, st));
  fused_prep_kernel<<<dim3(K, 5), 256, 0, st>>>(L, n_obj_max, params, pe_B, pe_scale, shape_codes, tex_codes, n_obj, prep, zlat);
  if (g_ev_start && g_ev_stop) CNF_CUDA(cudaEventRecord(g_ev_start, st));
  fused_step_kernel<<<dim3(cpc, K), NT, SM_TOTAL, st>>>(K, R, S, prep, origins, dirs, z, gt_depth, gt_rgb, mask_depth, inst,
                                                         counts, cs, os, raw, scratch, g_dbg);
  if (g_ev_start && g_ev_stop) { CNF_CUDA(cudaEventRecord(g_ev_stop, st)); g_ev_start = g_ev_stop = nullptr; }
  const size_t P = (size_t)R * S;
  int pb = (int)((P + 255) / 256); if (pb > 64) pb = 64;
  fused_pe_grad_kernel<<<dim3(pb, K), 256, 0, st>>>(P, scratch, g_B);
  fused_finalize_kernel<<<dim3(K, 8), 256, 0, st>>>(L, n_obj_max, params, pe_scale, shape_codes, tex_codes, n_obj, raw, zlat, reg_scaling,
                                           flags & 1, g_params, g_B, g_shape, g_tex, reg);
  fused_loss_kernel<<<1, 32, 0, st>>>(K, raw, counts, reg, cs, os, reg_scaling, flags & 1, out_loss);
  return check_launch("cnf_train_step");
}
