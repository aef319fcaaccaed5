// OccupancyMap (hidden size 128: the background model of the reference) on Blackwell tensor cores, forward and backward.
// Reference: src/model.py:86-154 (fc_block :8-12); called from train.py:174 on 1200 rays x (5+9) samples per step.
//
//   fc1 = relu(W_in xg + b)        fc2 = relu(W_mid1 fc1 + b)      fc3 = relu(W_cat [fc2 | xg] + b)     fc4 = relu(W_mid2 fc3 + b)
//   alpha = 10 (w_a . fc4 + b_a)   h = relu(W_col [fc4 | xc] + b)  rgb = sigmoid(W_oc h + b_oc)         (xg = emb[:87], xc = emb[87:])
//
// One CTA works on tiles of 128 points.  All GEMM operands are bf16 "core-matrix" images in shared memory (cnf_umma.cuh blk_off):
// the SAME bytes serve as the K-major operand of the forward GEMM, the MN-major operand of the data-gradient GEMM (weights) and
// the MN-major operands of the weight-gradient GEMM (activations / deltas, contraction over the tile's 128 points).  fp32
// accumulators live in tensor memory.  Weight images (192.5 KB, built by a prep kernel) do not fit next to the activations, so
// they stream through a 3-slot ring by 1-D bulk copies (TMA engine, mbarrier complete_tx) issued two layers ahead.
// Biases of in_layer / cat_layer / color_linear ride in a constant-one column of xg / xc (their weight-gradient column is the
// bias gradient).  The forward spills every activation image to a workspace (bulk shared->global copies; 1.5 KB/point, L2
// resident at the reference's batch size); the backward streams them back while the previous layer's GEMMs run.
// Weight gradients: every layer's accumulator is read out of tensor memory once per tile and added (coalesced RED.ADD.F32, stored
// transposed so that a warp's 32 lanes hit 128 contiguous bytes) into ONE pre-zeroed fp32 buffer; a small kernel scatters it
// into nn.Linear order.  (Per-CTA partial buffers + a reduce pass measured 26 % slower: 53 MB of extra L2 traffic.)
#include <cuda_bf16.h>

#include "cnf_common.cuh"
#include "cnf_umma.cuh"

namespace cnf {
using namespace umma;
namespace otc {

constexpr int H = 128, TPT = 128;
constexpr int E1 = 87, E2 = 42, EMB = 129;
constexpr int XG_C = 96, XC_C = 48, XG_ONE = 87, XC_ONE = 42;
constexpr int NT = 160;                         // 4 epilogue warps (thread = tile row = TMEM lane) + 1 control warp
constexpr int NT_B = 288;                       // backward: + 4 read-out warps (weight-gradient accumulators -> RED), control warp last

// weight blocks (bf16 images [128 x cols]): in | mid1 | cat(fc2 part) | cat(xg part) | mid2 | color(fc4 part) | color(xc part)
constexpr int NBLK = 7;
__host__ __device__ constexpr int blk_cols(int b) { return (b == 0 || b == 3) ? 96 : (b == 6 ? 48 : 128); }
__host__ __device__ constexpr int blk_byte(int b) {
  int o = 0;
  for (int i = 0; i < b; ++i) o += blk_cols(i) * 256;
  return o;
}
constexpr int WIMG_BYTES = 752 * 256;           // 192512
// fp32 constants behind the images
constexpr int C_BMID1 = 0, C_BMID2 = 128, C_WA = 256, C_WOC = 384, C_BA = 768, C_BOC = 769, C_PEB = 776 /* [21][3] */, C_ISCALE = 839,
              C_FLOATS = 840;
constexpr int PREP_BYTES = WIMG_BYTES + C_FLOATS * 4;
// flat parameter offsets (named_parameters() order, H = 128)
constexpr int P_INW = 0, P_INB = 11136, P_M1W = 11264, P_M1B = 27648, P_CATW = 27776, P_CATB = 55296, P_M2W = 55424,
              P_M2B = 71808, P_AW = 71936, P_AB = 72064, P_COLW = 72065, P_COLB = 93825, P_OCW = 93953, P_OCB = 94337,
              P_TOTAL = 94340;
// activation images of one tile in the workspace
constexpr int IMG_XG = 0, IMG_XC = 24576, IMG_FC1 = 36864, IMG_FC2 = 69632, IMG_FC3 = 102400, IMG_FC4 = 135168, IMG_H = 167936,
              IMG_DXC = 200704 /* backward scratch: d xc, fp32 [48][128] */, TILE_IMG_BYTES = 225280;
constexpr int ACT_BYTES = 32768;
constexpr int EMB_TILE_BYTES = TPT * EMB * 4;    // 66048: one tile's contiguous fp32 [128 x 129] block of emb / g_emb

// development aid: SM-clock stamps of CTA 0 (epilogue thread 0: slots 0..31, control lane 0: slots 32..63); NULL = off
// (compiled only into the development build, -DCNF_DEV: include/cnf_dev.h)
#ifdef CNF_DEV
__device__ long long* g_stamp = nullptr;
__device__ __forceinline__ void stamp(int slot) {
  if (g_stamp != nullptr && blockIdx.x == 0) g_stamp[slot] = clock64();
}
#else
__device__ __forceinline__ void stamp(int) {}
#endif

// packed operand pairs / single operand elements (fp16 by default, bf16 with -DCNF_OPERAND_FP16=0: cnf_umma.cuh)
__device__ __forceinline__ uint32_t pack2(float a, float b) { return op_pack2(a, b); }
__device__ __forceinline__ unsigned short op_elem(float a) { return (unsigned short)(op_pack2(a, 0.f) & 0xFFFFu); }
__device__ __forceinline__ uint8_t* rowblk(uint8_t* base, int row, int cb, int cols) {
  return base + (row >> 3) * (16 * cols) + cb * 128 + (row & 7) * 16;
}

__device__ __forceinline__ float wval(const float* __restrict__ P, int b, int n, int k) {
  switch (b) {
    case 0: return k < E1 ? P[P_INW + n * E1 + k] : (k == XG_ONE ? P[P_INB + n] : 0.f);
    case 1: return P[P_M1W + n * H + k];
    case 2: return P[P_CATW + n * (H + E1) + k];
    case 3: return k < E1 ? P[P_CATW + n * (H + E1) + H + k] : (k == XG_ONE ? P[P_CATB + n] : 0.f);
    case 4: return P[P_M2W + n * H + k];
    case 5: return P[P_COLW + n * (H + E2) + k];
    default: return k < E2 ? P[P_COLW + n * (H + E2) + H + k] : (k == XC_ONE ? P[P_COLB + n] : 0.f);
  }
}

// fp32 nn.Linear parameters -> bf16 operand images + fp32 constants.  One thread per 16-byte chunk.
__global__ void prep_kernel(const float* __restrict__ P, const float* __restrict__ pe_B /* may be NULL */,
                            const float* __restrict__ pe_scale, uint8_t* __restrict__ prep) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < WIMG_BYTES / 16) {
    const int byte = i * 16;
    int b = 0;
    while (b + 1 < NBLK && byte >= blk_byte(b + 1)) ++b;
    const int cols = blk_cols(b), local = byte - blk_byte(b);
    const int rg = local / (16 * cols), rem = local - rg * (16 * cols), cb = rem >> 7, r8 = (rem & 127) >> 4;
    const int n = rg * 8 + r8, k0 = cb * 8;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = pack2(wval(P, b, n, k0 + 2 * j), wval(P, b, n, k0 + 2 * j + 1));
    *reinterpret_cast<uint4*>(prep + byte) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (i < C_FLOATS) {
    float v = 0.f;
    if (i < C_BMID2) v = P[P_M1B + i];
    else if (i < C_WA) v = P[P_M2B + i - C_BMID2];
    else if (i < C_WOC) v = P[P_AW + i - C_WA];
    else if (i < C_BA) v = P[P_OCW + i - C_WOC];
    else if (i == C_BA) v = P[P_AB];
    else if (i < C_BOC + 3) v = P[P_OCB + i - C_BOC];
    else if (i >= C_PEB && i < C_PEB + 63 && pe_B != nullptr) v = pe_B[i - C_PEB];
    else if (i == C_ISCALE && pe_scale != nullptr) v = 1.0f / pe_scale[0];
    reinterpret_cast<float*>(prep + WIMG_BYTES)[i] = v;
  }
}

constexpr float PI_F = 3.14159265358979f;
// sin / cos of pi * proj (one MUFU pair after an exact range reduction), then the double-angle recurrence for the octaves
__device__ __forceinline__ void pe_sincos(const float* __restrict__ b3, const float t[3], float& s, float& c) {
  const float proj = fmaf(t[2], b3[2], fmaf(t[1], b3[1], t[0] * b3[0]));
  const float r = proj - 2.0f * rintf(0.5f * proj);
  __sincosf(PI_F * r, &s, &c);
}
__device__ __forceinline__ void pe_double(float& s, float& c) {
  const float s2 = 2.0f * s * c;
  c = fmaf(-2.0f * s, s, 1.0f);
  s = s2;
}

// forward issue: D[128 x 128] (+)= A[128 x aCols] (K-major image) * W[128 x aCols]^T (K-major image)
__device__ __forceinline__ void issue_fwd(uint32_t d, uint32_t aAddr, int aCols, uint32_t wAddr, bool acc) {
  const uint32_t id = idesc_op(128, 128, 0, 0);
  for (int ks = 0; ks < aCols / 16; ++ks)
    mma_ss(d, smem_desc(aAddr + ks * 256, 128, 16 * aCols), smem_desc(wAddr + ks * 256, 128, 16 * aCols), id, (acc || ks > 0) ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
// The fp32 embedding tile (128 x 129 floats = 66048 B, contiguous in global memory) is staged over A0 | A1 (+ 512 B of slack)
constexpr int F_XG = 0, F_XC = 24576, F_A0 = 36864, F_A1 = 69632, F_W = 103424, F_C = F_W + 3 * 32768, F_SMEM = F_C + C_FLOATS * 4;
static_assert(F_A0 + EMB_TILE_BYTES <= F_W, "embedding staging overlaps the weight ring");

// PTS = false: emb is the fp32 embedding [P,129] (module-level drop-in: OccupancyMap.forward takes the embedding);
// PTS = true: emb is the POINTS [P,3] and the embedding (UniDirsEmbed, src/embedding.py:82-92) is computed here, straight into the
// bf16 operand images (no [P,129] tensor exists).
template <bool PTS>
__global__ void __launch_bounds__(NT, 1)
fwd_kernel(int P, const uint8_t* __restrict__ prep, const float* __restrict__ emb, uint8_t* __restrict__ imgs /* NULL: inference */,
           float* __restrict__ alpha_out, float* __restrict__ rgb_out, int occ_act /* alpha_out = sigmoid(alpha) */) {
  extern __shared__ __align__(128) uint8_t sm[];   // no-swizzle operand images need 16-byte alignment only
  __shared__ uint64_t full[3];
  __shared__ uint64_t bar_mma, full_emb;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < C_FLOATS; i += NT) reinterpret_cast<float*>(sm + F_C)[i] = reinterpret_cast<const float*>(prep + WIMG_BYTES)[i];
  if (tid == 0) {
    for (int s = 0; s < 3; ++s) mbar_init(&full[s], 1);
    mbar_init(&bar_mma, 1);
    mbar_init(&full_emb, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int n_tiles = (P + TPT - 1) / TPT;
  const int nt = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t sbase = smem_u32(sm);

  if (warp == 4) {
    // ------------------------------------------------------------------ control warp: weight ring, spills, MMA issue
    int loaded = 0, consumed = 0, stage = 0;
    uint32_t mph = 0;
    const int total_blocks = nt * NBLK;
    auto load_more = [&]() {
      while (loaded < consumed + 3 && loaded < total_blocks) {
        const int b = loaded % NBLK, slot = loaded % 3;
        if (lane == 0) {
          const uint32_t bytes = blk_cols(b) * 256;
          mbar_expect_tx(&full[slot], bytes);
          bulk_g2s(sm + F_W + slot * ACT_BYTES, prep + blk_byte(b), bytes, &full[slot]);
        }
        ++loaded;
      }
      __syncwarp();
    };
    load_more();
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      uint8_t* img = imgs + (size_t)tile * TILE_IMG_BYTES;
      if (!PTS && (tile + 1) * TPT <= P) {               // full tile: one bulk copy of its contiguous fp32 embedding block
        if (lane == 0) {
          bulk_wait_read<0>();                           // the previous tile's spills out of A0 / A1 have been read
          mbar_expect_tx(&full_emb, EMB_TILE_BYTES);
          bulk_g2s(sm + F_A0, emb + (size_t)tile * TPT * EMB, EMB_TILE_BYTES, &full_emb);
        }
        __syncwarp();
      }
      for (int l = 0; l < 5; ++l) {
        const int a1 = (l == 0) ? F_XG : ((l & 1) ? F_A0 : F_A1);                 // input activation of layer l
        const int a1c = (l == 0) ? XG_C : H;
        const int a2 = (l == 2) ? F_XG : F_XC, a2c = (l == 2) ? XG_C : XC_C;     // second operand of the concatenating layers
        const int nb = (l == 2 || l == 4) ? 2 : 1;
        const int sp_src = (l == 0) ? F_XG : a1, sp_bytes = (l == 0) ? (XG_C + XC_C) * 256 : ACT_BYTES;
        const int sp_dst = (l == 0) ? IMG_XG : IMG_FC1 + (l - 1) * ACT_BYTES;
        nbar_sync(1 + (stage & 1), NT); ++stage;                                  // the layer's input image is complete
        for (int q = 0; q < nb; ++q) mbar_wait(&full[(consumed + q) % 3], ((consumed + q) / 3) & 1);
        if (lane == 0 && imgs != nullptr) {
          bulk_s2g(img + sp_dst, sm + sp_src, sp_bytes);                         // spill for the backward
          bulk_commit();
          bulk_wait_read<1>();                                                   // the image spilled one layer ago is free again
        }
        __syncwarp();
        if (elect_one()) {
          tc_fence_after();
          issue_fwd(tmem, sbase + a1, a1c, sbase + F_W + (consumed % 3) * ACT_BYTES, false);
          if (nb == 2) issue_fwd(tmem, sbase + a2, a2c, sbase + F_W + ((consumed + 1) % 3) * ACT_BYTES, true);
          mma_commit(&bar_mma);
        }
        __syncwarp();
        mbar_wait(&bar_mma, mph); mph ^= 1;
        consumed += nb;
        load_more();
      }
      nbar_sync(1 + (stage & 1), NT); ++stage;                                    // h is complete
      if (lane == 0 && imgs != nullptr) { bulk_s2g(img + IMG_H, sm + F_A0, ACT_BYTES); bulk_commit(); }
      __syncwarp();
    }
    if (lane == 0) bulk_wait<0>();
  } else {
    // ------------------------------------------------------------------ epilogue warps: thread = tile row
    const int row = tid;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const float* cs = reinterpret_cast<const float*>(sm + F_C);
    uint32_t mph = 0, eph = 0;
    int stage = 0, si = 0;
    auto mark = [&]() { if (tid == 0) stamp(si); ++si; };
    auto publish = [&]() { fence_async_smem(); tc_fence_before(); nbar_arrive(1 + (stage & 1), NT); ++stage; mark(); };
    // one layer's epilogue: D (+ bias) -> relu -> bf16 image; optional head dot products on the fp32 activations
    auto layer = [&](uint8_t* dst, const float* bias, const float* hw, int nh, float* hacc) {
      mbar_wait(&bar_mma, mph); mph ^= 1;
      tc_fence_after();
      mark();
      uint32_t uu[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld32(tl + ch * 32, uu[ch]);
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint32_t* u = uu[ch];
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(u[2 * j]), b = __uint_as_float(u[2 * j + 1]);
          if (bias != nullptr) { a += bias[ch * 32 + 2 * j]; b += bias[ch * 32 + 2 * j + 1]; }
          a = fmaxf(a, 0.f); b = fmaxf(b, 0.f);
          for (int q = 0; q < nh; ++q) hacc[q] = fmaf(hw[q * H + ch * 32 + 2 * j + 1], b, fmaf(hw[q * H + ch * 32 + 2 * j], a, hacc[q]));
          w[j] = pack2(a, b);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(rowblk(dst, row, ch * 4 + q, H)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      }
    };
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int p0 = tile * TPT, p = p0 + row;
      mark();
      if (PTS) {
        // embedding in registers: t = x / scale, sin(pi 2^f t.B_d) for 21 directions x 6 octaves, column order of the reference
        float t[3] = {0.f, 0.f, 0.f};
        if (p < P) {
#pragma unroll
          for (int q = 0; q < 3; ++q) t[q] = emb[(size_t)p * 3 + q] * cs[C_ISCALE];
        }
        float v[EMB + 7];
        v[0] = t[0]; v[1] = t[1]; v[2] = t[2];
#pragma unroll
        for (int d = 0; d < 21; ++d) {
          float sn, cn;
          pe_sincos(cs + C_PEB + d * 3, t, sn, cn);
          v[3 + d] = sn;
#pragma unroll
          for (int f = 1; f < 6; ++f) { pe_double(sn, cn); v[3 + f * 21 + d] = sn; }
        }
        v[EMB] = 1.0f;                                   // the ones column follows xc in this array (xc column 42)
#pragma unroll
        for (int j = 1; j < 7; ++j) v[EMB + j] = 0.f;
#pragma unroll
        for (int cb = 0; cb < 11; ++cb)
          *reinterpret_cast<uint4*>(rowblk(sm + F_XG, row, cb, XG_C)) =
              make_uint4(pack2(v[8 * cb], v[8 * cb + 1]), pack2(v[8 * cb + 2], v[8 * cb + 3]), pack2(v[8 * cb + 4], v[8 * cb + 5]),
                         pack2(v[8 * cb + 6], cb == 10 ? 1.0f : v[8 * cb + 7]));
        *reinterpret_cast<uint4*>(rowblk(sm + F_XG, row, 11, XG_C)) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int cb = 0; cb < 6; ++cb)
          *reinterpret_cast<uint4*>(rowblk(sm + F_XC, row, cb, XC_C)) =
              make_uint4(pack2(v[E1 + 8 * cb], v[E1 + 8 * cb + 1]), pack2(v[E1 + 8 * cb + 2], v[E1 + 8 * cb + 3]),
                         pack2(v[E1 + 8 * cb + 4], v[E1 + 8 * cb + 5]), pack2(v[E1 + 8 * cb + 6], v[E1 + 8 * cb + 7]));
      } else if (p0 + TPT <= P) {
        // embedding tile -> xg | xc images: staged by one bulk copy (the tile's fp32 [128 x 129] block is contiguous)
        // staged tile: thread = row converts its 129 floats (row stride 129 words: conflict-free) into the two images
        mbar_wait(&full_emb, eph); eph ^= 1;
        const float* src = reinterpret_cast<const float*>(sm + F_A0) + row * EMB;
#pragma unroll
        for (int cb = 0; cb < 11; ++cb) {
          const float last = (cb == 10) ? 1.0f : 0.f;    // column 87 = constant one
          *reinterpret_cast<uint4*>(rowblk(sm + F_XG, row, cb, XG_C)) =
              make_uint4(pack2(src[8 * cb], src[8 * cb + 1]), pack2(src[8 * cb + 2], src[8 * cb + 3]), pack2(src[8 * cb + 4], src[8 * cb + 5]),
                         pack2(src[8 * cb + 6], cb == 10 ? last : src[8 * cb + 7]));
        }
        *reinterpret_cast<uint4*>(rowblk(sm + F_XG, row, 11, XG_C)) = make_uint4(0u, 0u, 0u, 0u);
        const float* sc = src + E1;
#pragma unroll
        for (int cb = 0; cb < 5; ++cb)
          *reinterpret_cast<uint4*>(rowblk(sm + F_XC, row, cb, XC_C)) =
              make_uint4(pack2(sc[8 * cb], sc[8 * cb + 1]), pack2(sc[8 * cb + 2], sc[8 * cb + 3]), pack2(sc[8 * cb + 4], sc[8 * cb + 5]),
                         pack2(sc[8 * cb + 6], sc[8 * cb + 7]));
        *reinterpret_cast<uint4*>(rowblk(sm + F_XC, row, 5, XC_C)) = make_uint4(pack2(sc[40], sc[41]), pack2(1.0f, 0.f), 0u, 0u);
      } else {
        // ragged last tile: cooperative coalesced read of the valid rows
        const float* e = emb + (size_t)p0 * EMB;
        const int nvalid = min(TPT, P - p0) * EMB;
        for (int i = tid; i < TPT * EMB; i += 128) {
          const int r = i / EMB, c = i - r * EMB;
          const unsigned short hv = op_elem(i < nvalid ? e[i] : 0.f);
          if (c < E1) *reinterpret_cast<unsigned short*>(sm + F_XG + blk_off(r, c, XG_C)) = hv;
          else *reinterpret_cast<unsigned short*>(sm + F_XC + blk_off(r, c - E1, XC_C)) = hv;
        }
        const unsigned short one = (unsigned short)OP_ONE, zero = 0;
        *reinterpret_cast<unsigned short*>(sm + F_XG + blk_off(row, XG_ONE, XG_C)) = one;
        *reinterpret_cast<uint4*>(rowblk(sm + F_XG, row, 11, XG_C)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<unsigned short*>(sm + F_XC + blk_off(row, XC_ONE, XC_C)) = one;
        for (int c = XC_ONE + 1; c < XC_C; ++c) *reinterpret_cast<unsigned short*>(sm + F_XC + blk_off(row, c, XC_C)) = zero;
      }
      publish();
      layer(sm + F_A0, nullptr, nullptr, 0, nullptr); publish();                  // fc1
      layer(sm + F_A1, cs + C_BMID1, nullptr, 0, nullptr); publish();             // fc2
      layer(sm + F_A0, nullptr, nullptr, 0, nullptr); publish();                  // fc3
      float ha[1] = {0.f};
      layer(sm + F_A1, cs + C_BMID2, cs + C_WA, 1, ha); publish();                // fc4 + alpha head
      float hc[3] = {0.f, 0.f, 0.f};
      layer(sm + F_A0, nullptr, cs + C_WOC, 3, hc); publish();                    // h + colour head
      if (p < P) {
        const float al = 10.0f * (ha[0] + cs[C_BA]);                              // model.py:142-146
        alpha_out[p] = occ_act ? 1.0f / (1.0f + __expf(-al)) : al;                // render_rays.py:3-7 (occupancy_activation)
#pragma unroll
        for (int q = 0; q < 3; ++q) rgb_out[(size_t)p * 3 + q] = 1.0f / (1.0f + __expf(-(hc[q] + cs[C_BOC + q])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------------------
// forward-only evaluation (Trainer.eval_points of the background model): TWO CTAs per SM
// ---------------------------------------------------------------------------------------------------------
// Nothing is kept for a backward, so a CTA needs ONE activation image (a layer's epilogue overwrites the image its own GEMM has
// finished reading), the xg / xc operands live in tensor memory (A from TMEM), and the weight ring shrinks to two slots:
// 100 KB of shared memory and 256 TMEM columns per CTA.  Two CTAs share an SM: while one waits for a GEMM round trip or a weight
// block the other runs its epilogue -- the overlap the training kernels cannot have at one tile per SM.
constexpr int E_ACT = 0, E_W = 32768, E_C = E_W + 2 * 32768, E_SMEM = E_C + C_FLOATS * 4;
constexpr int ET_D = 0, ET_XG = 128, ET_XC = 176;     // TMEM columns: accumulator 128 | xg 96 bf16 | xc 48 bf16

__device__ __forceinline__ void issue_fwd_ts(uint32_t d, uint32_t aTmem, int aCols, uint32_t wAddr, bool acc) {
  const uint32_t id = idesc_op(128, 128, 0, 0);
  for (int ks = 0; ks < aCols / 16; ++ks)
    mma_ts(d, aTmem + ks * 8, smem_desc(wAddr + ks * 256, 128, 16 * aCols), id, (acc || ks > 0) ? 1u : 0u);
}

// RAYS = false: point p = points[p]; outputs occ_out [P], rgb_out [P,3].
// RAYS = true: point p = sample p % S of ray p / S: origin + dir * z (points = origins, dirs / zs alongside); the result goes out as
// one float4 {occupancy, r, g, b} per point (occ_out viewed as float4[P]) for the per-ray compositing kernel.
template <bool RAYS>
__global__ void __launch_bounds__(NT, 2)
eval_kernel(int P, const uint8_t* __restrict__ prep, const float* __restrict__ points, float* __restrict__ occ_out,
            float* __restrict__ rgb_out, int S, const float* __restrict__ dirs, const float* __restrict__ zs) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[2];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < C_FLOATS; i += NT) reinterpret_cast<float*>(sm + E_C)[i] = reinterpret_cast<const float*>(prep + WIMG_BYTES)[i];
  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int n_tiles = (P + TPT - 1) / TPT;
  const int nt = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t sbase = smem_u32(sm);

  if (warp == 4) {
    // ------------------------------------------------------------------ control warp: two-slot weight ring, MMA issue
    int loaded = 0, consumed = 0, stage = 0;
    uint32_t mph = 0;
    const int total_blocks = nt * NBLK;
    auto load_more = [&]() {
      while (loaded < consumed + 2 && loaded < total_blocks) {
        const int b = loaded % NBLK, slot = loaded & 1;
        if (lane == 0) {
          const uint32_t bytes = blk_cols(b) * 256;
          mbar_expect_tx(&full[slot], bytes);
          bulk_g2s(sm + E_W + slot * ACT_BYTES, prep + blk_byte(b), bytes, &full[slot]);
        }
        ++loaded;
      }
      __syncwarp();
    };
    load_more();
    for (int it = 0; it < nt; ++it) {
      for (int l = 0; l < 5; ++l) {
        const int nb = (l == 2 || l == 4) ? 2 : 1;
        nbar_sync(1 + (stage & 1), NT); ++stage;                                  // the layer's input operand is complete
        for (int q = 0; q < nb; ++q) mbar_wait(&full[(consumed + q) & 1], ((consumed + q) >> 1) & 1);
        if (elect_one()) {
          tc_fence_after();
          const uint32_t w0 = sbase + E_W + (consumed & 1) * ACT_BYTES, w1 = sbase + E_W + ((consumed + 1) & 1) * ACT_BYTES;
          if (l == 0) {
            issue_fwd_ts(tmem + ET_D, tmem + ET_XG, XG_C, w0, false);
          } else {
            issue_fwd(tmem + ET_D, sbase + E_ACT, H, w0, false);
            if (l == 2) issue_fwd_ts(tmem + ET_D, tmem + ET_XG, XG_C, w1, true);
            if (l == 4) issue_fwd_ts(tmem + ET_D, tmem + ET_XC, XC_C, w1, true);
          }
          mma_commit(&bar_mma);
        }
        __syncwarp();
        mbar_wait(&bar_mma, mph); mph ^= 1;
        consumed += nb;
        load_more();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: thread = tile row
    const int row = tid;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const float* cs = reinterpret_cast<const float*>(sm + E_C);
    uint32_t mph = 0;
    int stage = 0;
    auto publish = [&]() { fence_async_smem(); tmem_st_wait(); tc_fence_before(); nbar_arrive(1 + (stage & 1), NT); ++stage; };
    auto layer = [&](uint8_t* dst, const float* bias, const float* hw, int nh, float* hacc) {
      mbar_wait(&bar_mma, mph); mph ^= 1;
      tc_fence_after();
      uint32_t uu[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld32(tl + ET_D + ch * 32, uu[ch]);
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint32_t* u = uu[ch];
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(u[2 * j]), b = __uint_as_float(u[2 * j + 1]);
          if (bias != nullptr) { a += bias[ch * 32 + 2 * j]; b += bias[ch * 32 + 2 * j + 1]; }
          a = fmaxf(a, 0.f); b = fmaxf(b, 0.f);
          for (int q = 0; q < nh; ++q) hacc[q] = fmaf(hw[q * H + ch * 32 + 2 * j + 1], b, fmaf(hw[q * H + ch * 32 + 2 * j], a, hacc[q]));
          w[j] = pack2(a, b);
        }
        if (dst != nullptr) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(rowblk(dst, row, ch * 4 + q, H)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
      }
    };
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int p = tile * TPT + row;
      float t[3] = {0.f, 0.f, 0.f};
      if (p < P) {
        if (RAYS) {
          const int ray = p / S;
          const float zz = zs[p];
#pragma unroll
          for (int q = 0; q < 3; ++q)                                             // scene_cateogries.py:543-544 (mul, then add)
            t[q] = __fadd_rn(points[(size_t)ray * 3 + q], __fmul_rn(dirs[(size_t)ray * 3 + q], zz)) * cs[C_ISCALE];
        } else {
#pragma unroll
          for (int q = 0; q < 3; ++q) t[q] = points[(size_t)p * 3 + q] * cs[C_ISCALE];
        }
      }
      float v[EMB + 7];
      v[0] = t[0]; v[1] = t[1]; v[2] = t[2];
#pragma unroll
      for (int d = 0; d < 21; ++d) {
        float sn, cn;
        pe_sincos(cs + C_PEB + d * 3, t, sn, cn);
        v[3 + d] = sn;
#pragma unroll
        for (int f = 1; f < 6; ++f) { pe_double(sn, cn); v[3 + f * 21 + d] = sn; }
      }
      v[EMB] = 1.0f;
#pragma unroll
      for (int j = 1; j < 7; ++j) v[EMB + j] = 0.f;
#pragma unroll
      for (int cb = 0; cb < 11; ++cb)
        tmem_st4(tl + ET_XG + 4 * cb, pack2(v[8 * cb], v[8 * cb + 1]), pack2(v[8 * cb + 2], v[8 * cb + 3]), pack2(v[8 * cb + 4], v[8 * cb + 5]),
                 pack2(v[8 * cb + 6], cb == 10 ? 1.0f : v[8 * cb + 7]));
      tmem_st4(tl + ET_XG + 44, 0u, 0u, 0u, 0u);
#pragma unroll
      for (int cb = 0; cb < 6; ++cb)
        tmem_st4(tl + ET_XC + 4 * cb, pack2(v[E1 + 8 * cb], v[E1 + 8 * cb + 1]), pack2(v[E1 + 8 * cb + 2], v[E1 + 8 * cb + 3]),
                 pack2(v[E1 + 8 * cb + 4], v[E1 + 8 * cb + 5]), pack2(v[E1 + 8 * cb + 6], v[E1 + 8 * cb + 7]));
      publish();
      layer(sm + E_ACT, nullptr, nullptr, 0, nullptr); publish();                 // fc1
      layer(sm + E_ACT, cs + C_BMID1, nullptr, 0, nullptr); publish();            // fc2
      layer(sm + E_ACT, nullptr, nullptr, 0, nullptr); publish();                 // fc3
      float ha[1] = {0.f};
      layer(sm + E_ACT, cs + C_BMID2, cs + C_WA, 1, ha); publish();               // fc4 + alpha head
      float hc[3] = {0.f, 0.f, 0.f};
      layer(nullptr, nullptr, cs + C_WOC, 3, hc);                                 // h (not stored) + colour head
      tc_fence_before();
      if (p < P) {
        const float occ = 1.0f / (1.0f + __expf(-10.0f * (ha[0] + cs[C_BA])));    // model.py:142-146, render_rays.py:3-7
        float col[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) col[q] = 1.0f / (1.0f + __expf(-(hc[q] + cs[C_BOC + q])));
        if (RAYS) {
          reinterpret_cast<float4*>(occ_out)[p] = make_float4(occ, col[0], col[1], col[2]);
        } else {
          occ_out[p] = occ;
          rgb_out[(size_t)p * 3 + 0] = col[0]; rgb_out[(size_t)p * 3 + 1] = col[1]; rgb_out[(size_t)p * 3 + 2] = col[2];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// data gradient: D[128 x Nout] (+)= G[128 x 128] (K-major image) * W[128 x Nout] (MN-major view of the weight image)
__device__ __forceinline__ void issue_dgrad(uint32_t d, uint32_t gAddr, uint32_t wAddr, int Nout, bool acc) {
  const uint32_t id = idesc_op(128, Nout, 0, 1);
  for (int ks = 0; ks < H / 16; ++ks)
    mma_ss(d, smem_desc(gAddr + ks * 256, 128, 16 * H), smem_desc(wAddr + ks * 2 * (16 * Nout), 16 * Nout, 128), id, (acc || ks > 0) ? 1u : 0u);
}
// weight gradient: D[aCols (rows) x bCols] = A^T * B, both MN-major views of [128 points x cols] images (contraction over the points)
__device__ __forceinline__ void issue_wgrad(uint32_t d, uint32_t aAddr, int aCols, uint32_t bAddr, int bCols) {
  const uint32_t id = idesc_op(128, bCols, 1, 1);
  for (int ks = 0; ks < TPT / 16; ++ks)
    mma_ss(d, smem_desc(aAddr + ks * 2 * (16 * aCols), 16 * aCols, 128), smem_desc(bAddr + ks * 2 * (16 * bCols), 16 * bCols, 128), id,
           ks > 0 ? 1u : 0u);
}

// A0 | A1 | G1 are contiguous: in the last stage they are free and stage the tile's fp32 g_emb block (128 x 129 floats)
constexpr int B_A0 = 0, B_A1 = 32768, B_G1 = 65536, B_G0 = 98304, B_AX = 131072, B_W0 = 155648, B_W1 = 188416, B_ONES = 221184,
              B_HD = 225280, B_C = 229376, B_CPE = 128 + 384 /* [21][3] PE matrix, 1/scale at +63 */, B_SMEM = B_C + (128 + 384 + 64) * 4;
static_assert(B_A0 + EMB_TILE_BYTES <= B_G0, "g_emb staging must fit A0 | A1 | G1");
constexpr int T_DA = 0, T_DB = 128, T_DW = 224, T_DO = 448, T_DH = 464;
// per-CTA partial weight gradients (fp32): rows = output feature (TMEM lane), row lengths 96 / 128 / 224 / 128 / 176 / 1 / 1 / 32
constexpr int PR_IN = 0, PR_MID1 = 12288, PR_CAT = 28672, PR_MID2 = 57344, PR_COL = 73728, PR_DB1 = 96256, PR_DB2 = 96384,
              PR_HEAD = 96512, PR_HB = 100608, PR_PE = 100612 /* [21][3] d B_layer */, PART_FLOATS = 100676;

// PTS = false: writes g_emb [P,129] (may be NULL).  PTS = true: g_emb is the POINTS [P,3] (read), the embedding backward runs here
// (d proj_d = sum_f d emb[3 + 21 f + d] pi 2^f cos(pi 2^f proj_d); d B_d += d proj_d t) and no [P,129] gradient exists.
template <bool PTS>
__global__ void __launch_bounds__(NT_B, 1)
bwd_kernel(int P, const uint8_t* __restrict__ prep, uint8_t* __restrict__ imgs, const float* __restrict__ rgb,
           const float* __restrict__ g_alpha, const float* __restrict__ g_rgb, float* __restrict__ part /* [PART_FLOATS], zeroed */,
           float* __restrict__ g_emb) {
  extern __shared__ __align__(128) uint8_t sm[];   // no-swizzle operand images need 16-byte alignment only
  __shared__ uint64_t fullA0, fullA1, fullAX, fullH, fullW0, fullW1, bar_d, bar_w;
  __shared__ uint32_t tmem_base_s;
  // fp16 operands: every tile's incoming loss gradients (O(1 / rays): far below fp16's normal range at large batches) are scaled by
  // a power of two chosen from the tile's own largest |gradient| (mapped to [64, 128): 9 binades of head room up the data-gradient
  // chain, conversions saturate); everything downstream is linear in it, and the inverse is applied where fp32 results leave the
  // tile (accumulator read-out, PE / head-bias sums, g_emb).  s_gmax: per-warp maxima; s_inv[it & 1]: the tile's 1 / scale.
  __shared__ float s_gmax[4], s_inv[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    float* c = reinterpret_cast<float*>(sm + B_C);
    const float* src = reinterpret_cast<const float*>(prep + WIMG_BYTES);
    for (int i = tid; i < 128 + 384; i += NT_B) c[i] = src[C_WA + i];           // w_alpha | W_oc
    if (tid < 64) c[B_CPE + tid] = src[C_PEB + tid];                            // PE matrix | 1/scale
    // constant-one image [128 points x 16]: column 0 = 1
    for (int i = tid; i < 4096 / 16; i += NT_B) {
      // chunk i: row-group i / 2, column block i % 2 ... rows x 16 B; column 0 lives in column block 0
      *reinterpret_cast<uint4*>(sm + B_ONES + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  if (tid < TPT) *reinterpret_cast<unsigned short*>(sm + B_ONES + blk_off(tid, 0, 16)) = (unsigned short)OP_ONE;
  if (tid == 0) {
    mbar_init(&fullA0, 1); mbar_init(&fullA1, 1); mbar_init(&fullAX, 1); mbar_init(&fullH, 1); mbar_init(&fullW0, 1); mbar_init(&fullW1, 1);
    mbar_init(&bar_d, 1); mbar_init(&bar_w, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int n_tiles = (P + TPT - 1) / TPT;
  const int nt = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t sb = smem_u32(sm);

  if (warp == 8) {
    // ------------------------------------------------------------------ control warp
    uint32_t pA0 = 0, pA1 = 0, pAX = 0, pH = 0, pW0 = 0, pW1 = 0, pD = 0;
    int sg = 0, sd = 0;                                      // named-barrier stage counters: G-ready (ids 1,2), drained (ids 3,4)
    auto load = [&](int dst, const uint8_t* src, uint32_t bytes, uint64_t* bar) {
      if (lane == 0) { mbar_expect_tx(bar, bytes); bulk_g2s(sm + dst, src, bytes, bar); }
    };
    auto g_ready = [&]() { nbar_sync(1 + (sg & 1), 160); ++sg; };
    auto drained = [&]() { nbar_sync(3 + (sd & 1), 160); ++sd; };
    auto wait = [&](uint64_t* bar, uint32_t& ph) { mbar_wait(bar, ph); ph ^= 1; };
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const uint8_t* img = imgs + (size_t)tile * TILE_IMG_BYTES;
      if (it > 0) { drained(); nbar_sync(6, 160); }          // previous tile: accumulators read out, epilogue done with the buffers
      load(B_G1, img + IMG_H, ACT_BYTES, &fullH);
      load(B_A1, img + IMG_FC4, ACT_BYTES, &fullA1);
      load(B_AX, img + IMG_XC, XC_C * 256, &fullAX);
      load(B_A0, img + IMG_FC3, ACT_BYTES, &fullA0);
      load(B_W0, prep + blk_byte(5), 32768, &fullW0);
      load(B_W1, prep + blk_byte(6), XC_C * 256, &fullW1);
      __syncwarp();
      // ---- S1: color_linear (delta = d h in G0; h in G1)
      g_ready();
      wait(&fullH, pH); wait(&fullW0, pW0); wait(&fullW1, pW1);
      if (elect_one()) {
        tc_fence_after();
        issue_wgrad(tmem + T_DH, sb + B_G1, H, sb + B_HD, 16);                   // out_color weights: h^T x [d rgb_pre | d raw]
        issue_dgrad(tmem + T_DA, sb + B_G0, sb + B_W0, H, false);                // d fc4 (colour branch)
        issue_dgrad(tmem + T_DB, sb + B_G0, sb + B_W1, XC_C, false);             // d xc
        mma_commit(&bar_d);
      }
      __syncwarp();
      wait(&fullA1, pA1); wait(&fullAX, pAX);
      if (elect_one()) {
        issue_wgrad(tmem + T_DW, sb + B_G0, H, sb + B_A1, H);                    // color_linear weights: fc4 part
        issue_wgrad(tmem + T_DW + H, sb + B_G0, H, sb + B_AX, XC_C);             //                       xc part (+ bias column)
        issue_wgrad(tmem + T_DH + 16, sb + B_A1, H, sb + B_HD, 16);              // out_alpha weights: fc4^T x [.. | d raw]
        mma_commit(&bar_w);
      }
      __syncwarp();
      wait(&bar_d, pD);
      load(B_W0, prep + blk_byte(4), 32768, &fullW0);
      load(B_W1, prep + blk_byte(2), 32768, &fullW1);
      __syncwarp();
      // ---- S2: mid2 (delta = d fc4 in G1)
      g_ready();
      wait(&fullW0, pW0);
      if (elect_one()) { tc_fence_after(); issue_dgrad(tmem + T_DA, sb + B_G1, sb + B_W0, H, false); mma_commit(&bar_d); }
      __syncwarp();
      drained();                                             // S1's weight-gradient GEMMs are complete and read out: A1, AX free
      load(B_A1, img + IMG_FC2, ACT_BYTES, &fullA1);
      load(B_AX, img + IMG_XG, XG_C * 256, &fullAX);
      __syncwarp();
      wait(&fullA0, pA0);
      if (elect_one()) {
        tc_fence_after();
        issue_wgrad(tmem + T_DW, sb + B_G1, H, sb + B_A0, H);                    // mid2 weights
        issue_wgrad(tmem + T_DO, sb + B_G1, H, sb + B_ONES, 16);                 // mid2 bias
        mma_commit(&bar_w);
      }
      __syncwarp();
      wait(&bar_d, pD);
      load(B_W0, prep + blk_byte(3), XG_C * 256, &fullW0);
      __syncwarp();
      // ---- S3: cat_layer (delta = d fc3 in G0)
      g_ready();
      wait(&fullW1, pW1); wait(&fullW0, pW0);
      if (elect_one()) {
        tc_fence_after();
        issue_dgrad(tmem + T_DA, sb + B_G0, sb + B_W1, H, false);                // d fc2
        issue_dgrad(tmem + T_DB, sb + B_G0, sb + B_W0, XG_C, false);             // d xg (cat part)
        mma_commit(&bar_d);
      }
      __syncwarp();
      drained();
      load(B_A0, img + IMG_FC1, ACT_BYTES, &fullA0);
      __syncwarp();
      wait(&fullA1, pA1); wait(&fullAX, pAX);
      if (elect_one()) {
        tc_fence_after();
        issue_wgrad(tmem + T_DW, sb + B_G0, H, sb + B_A1, H);                    // cat weights: fc2 part
        issue_wgrad(tmem + T_DW + H, sb + B_G0, H, sb + B_AX, XG_C);             //              xg part (+ bias column)
        mma_commit(&bar_w);
      }
      __syncwarp();
      wait(&bar_d, pD);
      load(B_W1, prep + blk_byte(1), 32768, &fullW1);
      load(B_W0, prep + blk_byte(0), XG_C * 256, &fullW0);
      __syncwarp();
      // ---- S4: mid1 (delta = d fc2 in G1)
      g_ready();
      wait(&fullW1, pW1);
      if (elect_one()) { tc_fence_after(); issue_dgrad(tmem + T_DA, sb + B_G1, sb + B_W1, H, false); mma_commit(&bar_d); }
      __syncwarp();
      drained();
      wait(&fullA0, pA0);
      if (elect_one()) {
        tc_fence_after();
        issue_wgrad(tmem + T_DW, sb + B_G1, H, sb + B_A0, H);                    // mid1 weights
        issue_wgrad(tmem + T_DO, sb + B_G1, H, sb + B_ONES, 16);                 // mid1 bias
        mma_commit(&bar_w);
      }
      __syncwarp();
      wait(&bar_d, pD);
      // ---- S5: in_layer (delta = d fc1 in G0)
      g_ready();
      wait(&fullW0, pW0);
      if (elect_one()) { tc_fence_after(); issue_dgrad(tmem + T_DB, sb + B_G0, sb + B_W0, XG_C, true); mma_commit(&bar_d); }   // += onto cat's d xg
      __syncwarp();
      drained();
      if (elect_one()) { tc_fence_after(); issue_wgrad(tmem + T_DW, sb + B_G0, H, sb + B_AX, XG_C); mma_commit(&bar_w); }       // in_layer weights (+ bias column)
      __syncwarp();
      wait(&bar_d, pD);
    }
    if (nt > 0) { drained(); nbar_sync(6, 160); }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ read-out warps: thread = accumulator lane (output row).
    // Every stage's weight-gradient accumulators are added into the shared gradient buffer (stored TRANSPOSED, [column][row], so the
    // 32 lanes of a warp hit 128 contiguous bytes per column: coalesced RED.ADD at the L2) off the delta epilogues' critical path.
    const int row = tid - 128;
    const uint32_t tl = tmem + ((uint32_t)((warp - 4) * 32) << 16);
    uint32_t pW = 0;
    int sd = 0;
    float inv = 1.0f;                                      // this tile's 1 / gradient scale
    auto drain = [&](int tcol, int ncols, float* dst) {
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t u[32];
        const int n = min(32, ncols - c0);
        if (n == 32) tmem_ld32(tl + tcol + c0, u); else tmem_ld16(tl + tcol + c0, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < n) atomicAdd(dst + (size_t)(c0 + j) * 128, __uint_as_float(u[j]) * inv);
      }
    };
    auto bias = [&](float* d) { uint32_t u[4]; tmem_ld4(tl + T_DO, u); tmem_ld_wait(); atomicAdd(d, __uint_as_float(u[0]) * inv); };
    auto begin = [&]() { mbar_wait(&bar_w, pW); pW ^= 1; tc_fence_after(); };
    auto done = [&]() { tc_fence_before(); nbar_arrive(3 + (sd & 1), 160); ++sd; };
    for (int it = 0; it < nt; ++it) {
      // (s_inv[it & 1] was written before the epilogue published stage S1, which precedes the first bar_w completion of this tile)
      begin(); inv = *reinterpret_cast<volatile float*>(&s_inv[it & 1]);
      drain(T_DW, H + XC_C, part + PR_COL + row); drain(T_DH, 32, part + PR_HEAD + row); done();               // S1 color_linear + heads
      begin(); drain(T_DW, H, part + PR_MID2 + row); bias(part + PR_DB2 + row); done();                        // S2 mid2
      begin(); drain(T_DW, H + XG_C, part + PR_CAT + row); done();                                             // S3 cat_layer
      begin(); drain(T_DW, H, part + PR_MID1 + row); bias(part + PR_DB1 + row); done();                        // S4 mid1
      begin(); drain(T_DW, XG_C, part + PR_IN + row); done();                                                  // S5 in_layer
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: thread = tile row
    const int row = tid;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const float* wa = reinterpret_cast<const float*>(sm + B_C);
    const float* woc = wa + 128;
    uint32_t pA0 = 0, pA1 = 0, pH = 0, pD = 0;
    int sg = 0, si = 0;
    float acc_hb = 0.f, acc_pe0 = 0.f, acc_pe1 = 0.f;    // per-lane partial sums of the head biases / the PE matrix gradient
    auto mark = [&]() { if (tid == 0) stamp(si); ++si; };
    auto wait = [&](uint64_t* bar, uint32_t& ph) { mbar_wait(bar, ph); ph ^= 1; mark(); };
    auto publish = [&]() { fence_async_smem(); tc_fence_before(); nbar_arrive(1 + (sg & 1), 160); ++sg; mark(); };
    for (int it = 0; it < nt; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int p = tile * TPT + row;
      const bool act = p < P;
      mark();
      // delta epilogue: D (+ extra * xw) masked by the layer's own activation (> 0) -> bf16 image
      auto delta = [&](const uint8_t* actImg, uint8_t* dst, float extra, const float* xw) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {             // two rounds of 64 columns (register budget of the 288-thread kernel)
        uint32_t uu[2][32];
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) tmem_ld32(tl + T_DA + (half * 2 + c2) * 32, uu[c2]);
        tmem_ld_wait();
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int ch = half * 2 + c2;
          const uint32_t* u = uu[c2];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 a = *reinterpret_cast<const uint4*>(actImg + (row >> 3) * (16 * H) + (ch * 4 + q) * 128 + (row & 7) * 16);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = ch * 32 + q * 8 + e * 2;
              float v0 = __uint_as_float(u[q * 8 + e * 2]), v1 = __uint_as_float(u[q * 8 + e * 2 + 1]);
              if (xw != nullptr) { v0 = fmaf(extra, xw[c], v0); v1 = fmaf(extra, xw[c + 1], v1); }
              // bf16 activations are >= 0 after relu: positive <=> non-zero bits (no negative zero is ever produced by fmaxf(x, 0))
              if ((aw[e] & 0xffffu) == 0u) v0 = 0.f;
              if ((aw[e] >> 16) == 0u) v1 = 0.f;
              w[e] = pack2(v0, v1);
            }
            *reinterpret_cast<uint4*>(rowblk(dst, row, ch * 4 + q, H)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        }
      };
      // ---- S0: loss gradients of the heads -> d h (G0) and the head-delta image
      float dp[3] = {0.f, 0.f, 0.f}, draw = 0.f;
      if (act) {
#pragma unroll
        for (int q = 0; q < 3; ++q) { const float c = rgb[(size_t)p * 3 + q]; dp[q] = g_rgb[(size_t)p * 3 + q] * c * (1.0f - c); }   // sigmoid head
        draw = 10.0f * g_alpha[p];                                                                                                     // alpha = 10 raw
      }
      {   // head-bias gradients: lane q of each warp accumulates sum q in a register (shared-memory float atomics are CAS loops)
        const float s0 = warp_sum(dp[0]), s1 = warp_sum(dp[1]), s2 = warp_sum(dp[2]), s3 = warp_sum(draw);
        acc_hb += lane == 0 ? s0 : (lane == 1 ? s1 : (lane == 2 ? s2 : (lane == 3 ? s3 : 0.f)));
      }
      float tinv = 1.0f;                                                         // 1 / this tile's gradient scale
#if CNF_OPERAND_FP16
      {
        float mx = fmaxf(fmaxf(fabsf(dp[0]), fabsf(dp[1])), fmaxf(fabsf(dp[2]), fabsf(draw)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_gmax[warp] = mx;
        nbar_sync(5, TPT);
        mx = fmaxf(fmaxf(s_gmax[0], s_gmax[1]), fmaxf(s_gmax[2], s_gmax[3]));
        // finite non-zero maximum 2^e (1 + f) -> scale 2^(6 - e); zero / non-finite gradients are passed through unscaled
        const int e = (int)((__float_as_uint(mx) >> 23) & 0xFFu);
        float gs = 1.0f;
        if (e > 0 && e < 255) { const int se = min(max(127 + 6 - (e - 127), 1), 254); gs = __uint_as_float((uint32_t)se << 23); }
        tinv = 1.0f / gs;
        if (tid == 0) s_inv[it & 1] = tinv;
        dp[0] *= gs; dp[1] *= gs; dp[2] *= gs; draw *= gs;
        nbar_sync(5, TPT);                                                       // s_gmax may be rewritten by the next tile
      }
#endif
      *reinterpret_cast<uint4*>(rowblk(sm + B_HD, row, 0, 16)) = make_uint4(pack2(dp[0], dp[1]), pack2(dp[2], draw), 0u, 0u);
      *reinterpret_cast<uint4*>(rowblk(sm + B_HD, row, 1, 16)) = make_uint4(0u, 0u, 0u, 0u);
      wait(&fullH, pH);
#pragma unroll 1
      for (int cb = 0; cb < 16; ++cb) {
        const uint4 a = *reinterpret_cast<const uint4*>(sm + B_G1 + (row >> 3) * (16 * H) + cb * 128 + (row & 7) * 16);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = cb * 8 + e * 2;
          float v0 = fmaf(woc[2 * H + c], dp[2], fmaf(woc[H + c], dp[1], woc[c] * dp[0]));
          float v1 = fmaf(woc[2 * H + c + 1], dp[2], fmaf(woc[H + c + 1], dp[1], woc[c + 1] * dp[0]));
          if ((aw[e] & 0xffffu) == 0u) v0 = 0.f;
          if ((aw[e] >> 16) == 0u) v1 = 0.f;
          w[e] = pack2(v0, v1);
        }
        *reinterpret_cast<uint4*>(rowblk(sm + B_G0, row, cb, H)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      publish();
      // ---- S1: color_linear
      wait(&bar_d, pD); tc_fence_after();
      wait(&fullA1, pA1);
      delta(sm + B_A1, sm + B_G1, draw, wa);                                     // d fc4 = mask(fc4) (colour branch + alpha head)
      float* dxc = reinterpret_cast<float*>(imgs + (size_t)tile * TILE_IMG_BYTES + IMG_DXC) + row;      // [column][row]: coalesced
      const float* peB = wa + B_CPE;
      float tpt[3] = {0.f, 0.f, 0.f};                                            // PTS: this row's t = x / scale
      if (PTS && act) {
#pragma unroll
        for (int q = 0; q < 3; ++q) tpt[q] = g_emb[(size_t)p * 3 + q] * peB[63];
      }
      if (PTS) {
        // octaves 4, 5 of d proj from d xc; parked (as d xc is in the other mode) until the last stage
        uint32_t u[32], v[16];
        tmem_ld32(tl + T_DB, u); tmem_ld16(tl + T_DB + 32, v); tmem_ld_wait();
        float gx[48];
#pragma unroll
        for (int c = 0; c < 32; ++c) gx[c] = __uint_as_float(u[c]);
#pragma unroll
        for (int c = 0; c < 16; ++c) gx[32 + c] = __uint_as_float(v[c]);
#pragma unroll
        for (int d = 0; d < 21; ++d) {
          float sn, cn;
          pe_sincos(peB + d * 3, tpt, sn, cn);
          pe_double(sn, cn); pe_double(sn, cn); pe_double(sn, cn); pe_double(sn, cn);        // octave 4
          float gp = gx[d] * (16.0f * PI_F) * cn;
          pe_double(sn, cn);
          gp = fmaf(gx[21 + d] * (32.0f * PI_F), cn, gp);
          dxc[d * 128] = gp;
        }
      } else if (g_emb != nullptr) {                                             // d xc: parked until the last stage
        uint32_t u[32], v[16];
        tmem_ld32(tl + T_DB, u); tmem_ld16(tl + T_DB + 32, v); tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) dxc[c * 128] = __uint_as_float(u[c]);
#pragma unroll
        for (int c = 0; c < E2 - 32; ++c) dxc[(32 + c) * 128] = __uint_as_float(v[c]);
      }
      publish();
      // ---- S2: mid2
      wait(&bar_d, pD); tc_fence_after();
      wait(&fullA0, pA0);
      delta(sm + B_A0, sm + B_G0, 0.f, nullptr);                                 // d fc3
      publish();
      // ---- S3: cat_layer
      wait(&bar_d, pD); tc_fence_after();
      wait(&fullA1, pA1);
      delta(sm + B_A1, sm + B_G1, 0.f, nullptr);                                 // d fc2
      publish();
      // ---- S4: mid1
      wait(&bar_d, pD); tc_fence_after();
      wait(&fullA0, pA0);
      delta(sm + B_A0, sm + B_G0, 0.f, nullptr);                                 // d fc1
      publish();
      // ---- S5: in_layer
      wait(&bar_d, pD); tc_fence_after();
      if (PTS) {
        // embedding backward: octaves 0..3 from d xg, plus the parked octaves 4, 5; d B_d += d proj_d * t, summed over the warp's
        // 32 rows by shuffles, then over the CTA in shared memory (flushed with 63 REDs per CTA at the end)
        uint32_t uu[3][32];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) tmem_ld32(tl + T_DB + ch * 32, uu[ch]);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 21; ++d) {
          float sn, cn;
          pe_sincos(peB + d * 3, tpt, sn, cn);
          float gp = dxc[d * 128];
          float w = PI_F;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int col = 3 + f * 21 + d;
            gp = fmaf(__uint_as_float(uu[col >> 5][col & 31]) * w, cn, gp);
            pe_double(sn, cn);
            w *= 2.0f;
          }
          gp = act ? gp * tinv : 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float r = warp_sum(gp * tpt[q]);
            const int i = d * 3 + q;                     // lane i % 32 keeps sum i in a register
            if (lane == (i & 31)) { if (i < 32) acc_pe0 += r; else acc_pe1 += r; }
          }
        }
      } else if (g_emb != nullptr) {
        // g_emb rows of the tile are one contiguous block: stage [128][129] fp32 in shared memory (row stride 129 words:
        // conflict-free), then ONE bulk copy (ragged last tile: coalesced cooperative stores)
        float* stg = reinterpret_cast<float*>(sm + B_A0) + row * EMB;
        uint32_t uu[3][32];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) tmem_ld32(tl + T_DB + ch * 32, uu[ch]);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
          for (int c = 0; c < 32; ++c) if (ch * 32 + c < E1) stg[ch * 32 + c] = __uint_as_float(uu[ch][c]) * tinv;   // d xg (cat part + in part)
#pragma unroll
        for (int c = 0; c < E2; ++c) stg[E1 + c] = dxc[c * 128] * tinv;
        fence_async_smem();
        nbar_sync(5, TPT);
        const int p0 = tile * TPT;
        if (p0 + TPT <= P) {
          if (tid == 0) { bulk_s2g(g_emb + (size_t)p0 * EMB, sm + B_A0, EMB_TILE_BYTES); bulk_commit(); bulk_wait_read<0>(); }
        } else {
          const float* all = reinterpret_cast<const float*>(sm + B_A0);
          const int nvalid = (P - p0) * EMB;
          for (int i = tid; i < nvalid; i += TPT) g_emb[(size_t)p0 * EMB + i] = all[i];
        }
      }
      nbar_arrive(6, 160);                                   // this tile's shared-memory buffers may be reloaded
    }
    if (lane < 4) atomicAdd(&part[PR_HB + lane], acc_hb);
    if (PTS) {
      atomicAdd(&part[PR_PE + lane], acc_pe0);
      if (lane < 31) atomicAdd(&part[PR_PE + 32 + lane], acc_pe1);
    }
  }
  if (tid == 0) bulk_wait<0>();                  // g_emb bulk stores have landed
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// accumulation buffer -> flat nn.Linear-ordered gradient (OVERWRITES g_params); one thread per slot, slot (column k, row o) of a
// region lives at region + k * 128 + o.
__global__ void reduce_kernel(int n_part, const float* __restrict__ partials, float* __restrict__ g_params,
                              float* __restrict__ g_B /* [21][3], may be NULL */) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= PART_FLOATS) return;
  if (s >= PR_PE) {
    if (g_B != nullptr && s < PR_PE + 63) g_B[s - PR_PE] = partials[s];
    return;
  }
  int dst = -1;
  if (s < PR_MID1) { const int k = s >> 7, o = s & 127; dst = k < E1 ? P_INW + o * E1 + k : (k == XG_ONE ? P_INB + o : -1); }
  else if (s < PR_CAT) { const int t = s - PR_MID1, k = t >> 7, o = t & 127; dst = P_M1W + o * H + k; }
  else if (s < PR_MID2) { const int t = s - PR_CAT, k = t >> 7, o = t & 127; dst = k < H + E1 ? P_CATW + o * (H + E1) + k : (k == H + XG_ONE ? P_CATB + o : -1); }
  else if (s < PR_COL) { const int t = s - PR_MID2, k = t >> 7, o = t & 127; dst = P_M2W + o * H + k; }
  else if (s < PR_DB1) { const int t = s - PR_COL, k = t >> 7, o = t & 127; dst = k < H + E2 ? P_COLW + o * (H + E2) + k : (k == H + XC_ONE ? P_COLB + o : -1); }
  else if (s < PR_DB2) dst = P_M1B + (s - PR_DB1);
  else if (s < PR_HEAD) dst = P_M2B + (s - PR_DB2);
  else if (s < PR_HB) { const int t = s - PR_HEAD, j = t >> 7, c = t & 127; dst = j < 3 ? P_OCW + j * H + c : (j == 19 ? P_AW + c : -1); }
  else { const int t = s - PR_HB; dst = t < 3 ? P_OCB + t : (t == 3 ? P_AB : -1); }
  if (dst < 0) return;
  float a = 0.f;
#pragma unroll 4
  for (int c = 0; c < n_part; ++c) a += partials[(size_t)c * PART_FLOATS + s];
  g_params[dst] = a;
}

}  // namespace otc
}  // namespace cnf

using namespace cnf;
using namespace cnf::otc;

static size_t ws_prep() { return 0; }
static size_t ws_imgs() { return align_up(PREP_BYTES, 1024); }
static size_t ws_imgs_bytes(int P) { return (size_t)((P + TPT - 1) / TPT) * TILE_IMG_BYTES; }

static size_t ws_parts(int P) { return ws_imgs() + align_up(ws_imgs_bytes(P), 1024); }

extern "C" size_t cnf_occmap_tc_workspace_bytes(int n_pts) {
  if (n_pts < 0) n_pts = 0;
  return ws_parts(n_pts) + (size_t)PART_FLOATS * sizeof(float);
}

static int otc_forward(bool pts, const char* who, int P, const float* params, const float* pe_B, const float* pe_scale, const float* in,
                       float* alpha, float* rgb, void* ws, size_t ws_bytes, cnf_stream_t stream, bool prepared = false) {
  if (P < 0) { set_error("%s: bad size P=%d", who, P); return CNF_ERR_ARG; }
  if (!(params && ws && (P == 0 || (in && alpha && rgb)) && (!pts || (pe_B && pe_scale)))) { set_error("%s: null pointer", who); return CNF_ERR_ARG; }
  if (ws_bytes < cnf_occmap_tc_workspace_bytes(P)) { set_error("%s: workspace too small", who); return CNF_ERR_WORKSPACE; }
  if (P == 0) return CNF_OK;
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    CNF_CUDA(cudaFuncSetAttribute(fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    attr_set.done();
  }
  int dev = 0, n_sm = 148;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  if (!prepared) prep_kernel<<<(WIMG_BYTES / 16 + 255) / 256, 256, 0, st>>>(params, pts ? pe_B : nullptr, pts ? pe_scale : nullptr, w8 + ws_prep());
  const int n_tiles = (P + TPT - 1) / TPT, grid = n_tiles < n_sm ? n_tiles : n_sm;
  if (pts) fwd_kernel<true><<<grid, NT, F_SMEM, st>>>(P, w8 + ws_prep(), in, w8 + ws_imgs(), alpha, rgb, 0);
  else fwd_kernel<false><<<grid, NT, F_SMEM, st>>>(P, w8 + ws_prep(), in, w8 + ws_imgs(), alpha, rgb, 0);
  return check_launch(who);
}

// forward-only evaluation at arbitrary points (Trainer.eval_points of the background model, trainer.py:125-151): no activation
// images are kept; occ = sigmoid(alpha)
extern "C" size_t cnf_occmap_tc_eval_workspace_bytes(void) { return align_up(PREP_BYTES, 1024); }
static int otc_eval(const char* who, int P, const float* params, const float* pe_B, const float* pe_scale, const float* points, float* occ,
                    float* rgb, int S, const float* dirs, const float* zs, void* ws, cnf_stream_t stream, bool do_prep = true) {
  if (P == 0) return CNF_OK;
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM));
    CNF_CUDA(cudaFuncSetAttribute(eval_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CNF_CUDA(cudaFuncSetAttribute(eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, E_SMEM));
    CNF_CUDA(cudaFuncSetAttribute(eval_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set.done();
  }
  int dev = 0, n_sm = 148;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  if (do_prep) prep_kernel<<<(WIMG_BYTES / 16 + 255) / 256, 256, 0, st>>>(params, pe_B, pe_scale, w8);
  const int n_tiles = (P + TPT - 1) / TPT, grid = n_tiles < 2 * n_sm ? n_tiles : 2 * n_sm;       // two CTAs per SM
  if (dirs == nullptr) {
    eval_kernel<false><<<grid, NT, E_SMEM, st>>>(P, w8, points, occ, rgb, 1, nullptr, nullptr);
  } else {
    eval_kernel<true><<<grid, NT, E_SMEM, st>>>(P, w8, points, occ, nullptr, S, dirs, zs);
  }
  return check_launch(who);
}

extern "C" int cnf_occmap_tc_eval_points(int P, const float* params, const float* pe_B, const float* pe_scale, const float* points,
                                         float* occ, float* rgb, void* ws, size_t ws_bytes, cnf_stream_t stream) {
  CNF_REQUIRE(P >= 0 && params && pe_B && pe_scale && ws && (P == 0 || (points && occ && rgb)), "cnf_occmap_tc_eval_points: bad arguments");
  if (ws_bytes < cnf_occmap_tc_eval_workspace_bytes()) { set_error("cnf_occmap_tc_eval_points: workspace too small"); return CNF_ERR_WORKSPACE; }
  return otc_eval("cnf_occmap_tc_eval_points", P, params, pe_B, pe_scale, points, occ, rgb, 1, nullptr, nullptr, ws, stream);
}

// forward-only rendering of rays through the background model (dense views; any S): point kernel in chunks whose {occ, rgb} stay in
// L2, then per-ray compositing -- the OccupancyMap counterpart of cnf_render_rays
constexpr int OTC_RENDER_CHUNK_POINTS = 4 << 20;
static size_t otc_render_chunk_rays(int R, int S) { size_t rc = (size_t)OTC_RENDER_CHUNK_POINTS / S; if (rc < 1) rc = 1; return rc < (size_t)R ? rc : (size_t)R; }
extern "C" size_t cnf_occmap_tc_render_workspace_bytes(int R, int S) {
  const size_t base = cnf_occmap_tc_eval_workspace_bytes();
  if (R < 1 || S < 1) return base;
  return base + otc_render_chunk_rays(R, S) * (size_t)S * 16;
}
extern "C" int cnf_occmap_tc_render_rays(int R, int S, const float* params, const float* pe_B, const float* pe_scale, const float* origins,
                                         const float* dirs, const float* z, float* out_rays, float* out_w, void* ws, size_t ws_bytes,
                                         cnf_stream_t stream) {
  CNF_REQUIRE(R >= 0 && S >= 1 && params && pe_B && pe_scale && ws && (R == 0 || (origins && dirs && z && out_rays)),
              "cnf_occmap_tc_render_rays: bad arguments");
  if (ws_bytes < cnf_occmap_tc_render_workspace_bytes(R, S)) { set_error("cnf_occmap_tc_render_rays: workspace too small"); return CNF_ERR_WORKSPACE; }
  if (R == 0) return CNF_OK;
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  float* pts = reinterpret_cast<float*>(w8 + cnf_occmap_tc_eval_workspace_bytes());
  const size_t rc = otc_render_chunk_rays(R, S);
  for (size_t r0 = 0; r0 < (size_t)R; r0 += rc) {
    const int nr = (int)((size_t)R - r0 < rc ? (size_t)R - r0 : rc);
    if (int e = otc_eval("cnf_occmap_tc_render_rays", nr * S, params, pe_B, pe_scale, origins + r0 * 3, pts, nullptr, S, dirs + r0 * 3,
                         z + r0 * S, ws, stream, r0 == 0))
      return e;
    if (int e = render_composite_launch(nr, S, pts, z + r0 * S, out_rays + r0 * 6, out_w ? out_w + r0 * S : nullptr, as_stream(stream))) return e;
  }
  return CNF_OK;
}

extern "C" int cnf_occmap_tc_fwd(int P, const float* params, const float* emb, float* alpha, float* rgb, void* ws, size_t ws_bytes,
                                 cnf_stream_t stream) {
  return otc_forward(false, "cnf_occmap_tc_fwd", P, params, nullptr, nullptr, emb, alpha, rgb, ws, ws_bytes, stream);
}

extern "C" int cnf_occmap_tc_fwd_points(int P, const float* params, const float* pe_B, const float* pe_scale, const float* points,
                                        float* alpha, float* rgb, void* ws, size_t ws_bytes, int flags, cnf_stream_t stream) {
  return otc_forward(true, "cnf_occmap_tc_fwd_points", P, params, pe_B, pe_scale, points, alpha, rgb, ws, ws_bytes, stream, (flags & 1) != 0);
}

// the parameter-only half of a step (operand images, constants, zeroed gradient buffer): may run on another stream next to the sampler
extern "C" int cnf_occmap_tc_prepare(int P, const float* params, const float* pe_B, const float* pe_scale, void* ws, size_t ws_bytes,
                                     cnf_stream_t stream) {
  CNF_REQUIRE(P >= 0 && params && ws, "cnf_occmap_tc_prepare: bad arguments");
  if (ws_bytes < cnf_occmap_tc_workspace_bytes(P)) { set_error("cnf_occmap_tc_prepare: workspace too small"); return CNF_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  CNF_CUDA(cudaMemsetAsync(w8 + ws_parts(P), 0, PART_FLOATS * sizeof(float), st));
  prep_kernel<<<(WIMG_BYTES / 16 + 255) / 256, 256, 0, st>>>(params, pe_B, pe_B ? pe_scale : nullptr, w8 + ws_prep());
  return check_launch("cnf_occmap_tc_prepare");
}

static int otc_backward(bool pts, const char* who, int P, const float* params, const float* points, const float* rgb, const float* g_alpha,
                        const float* g_rgb, void* ws, size_t ws_bytes, float* g_params, float* g_emb, float* g_B, cnf_stream_t stream,
                        bool prepared = false) {
  if (P <= 0) { set_error("%s: bad size P=%d", who, P); return CNF_ERR_ARG; }
  if (!(params && rgb && g_alpha && g_rgb && ws && g_params && (!pts || (points && g_B)))) { set_error("%s: null pointer", who); return CNF_ERR_ARG; }
  if (ws_bytes < cnf_occmap_tc_workspace_bytes(P)) { set_error("%s: workspace too small", who); return CNF_ERR_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
  static DeviceOnce attr_set;
  if (attr_set.pending()) {
    CNF_CUDA(cudaFuncSetAttribute(bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    CNF_CUDA(cudaFuncSetAttribute(bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    attr_set.done();
  }
  int dev = 0, n_sm = 148;
  CNF_CUDA(cudaGetDevice(&dev));
  CNF_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int n_tiles = (P + TPT - 1) / TPT;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  float* parts = reinterpret_cast<float*>(w8 + ws_parts(P));
  if (!prepared) CNF_CUDA(cudaMemsetAsync(parts, 0, PART_FLOATS * sizeof(float), st));
  if (pts) bwd_kernel<true><<<grid, NT_B, B_SMEM, st>>>(P, w8 + ws_prep(), w8 + ws_imgs(), rgb, g_alpha, g_rgb, parts, const_cast<float*>(points));
  else bwd_kernel<false><<<grid, NT_B, B_SMEM, st>>>(P, w8 + ws_prep(), w8 + ws_imgs(), rgb, g_alpha, g_rgb, parts, g_emb);
  reduce_kernel<<<(PART_FLOATS + 255) / 256, 256, 0, st>>>(1, parts, g_params, pts ? g_B : nullptr);
  return check_launch(who);
}

extern "C" int cnf_occmap_tc_bwd(int P, const float* params, const float* rgb, const float* g_alpha, const float* g_rgb, void* ws,
                                 size_t ws_bytes, float* g_params, float* g_emb, cnf_stream_t stream) {
  return otc_backward(false, "cnf_occmap_tc_bwd", P, params, nullptr, rgb, g_alpha, g_rgb, ws, ws_bytes, g_params, g_emb, nullptr, stream);
}

extern "C" int cnf_occmap_tc_bwd_points(int P, const float* params, const float* points, const float* rgb, const float* g_alpha,
                                        const float* g_rgb, void* ws, size_t ws_bytes, float* g_params, float* g_B, int flags,
                                        cnf_stream_t stream) {
  return otc_backward(true, "cnf_occmap_tc_bwd_points", P, params, points, rgb, g_alpha, g_rgb, ws, ws_bytes, g_params, nullptr, g_B, stream,
                      (flags & 1) != 0);
}

#ifdef CNF_DEV
extern "C" int cnf_occmap_tc_set_debug(void* stamps /* device buffer of 64 int64, or NULL */) {
  long long* p = reinterpret_cast<long long*>(stamps);
  CNF_CUDA(cudaMemcpyToSymbol(cnf::otc::g_stamp, &p, sizeof(p)));
  return CNF_OK;
}
#endif
