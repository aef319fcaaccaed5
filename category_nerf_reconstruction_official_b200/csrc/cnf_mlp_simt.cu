// fp32 CUDA-core MLP path: CodeNeRF.forward / OccupancyMap.forward (reference src/model.py:56-84, 124-154)
// and their backward, for the module-level drop-in API and as the exact-precision device reference the
// tcgen05 fused step (cnf_fused.cu) is checked against.  Any width W, any latent size L, category-batched
// (blockIdx.z = category, like functorch.vmap over stacked parameters, reference train.py:154-155).
//
// Three generic kernels -- linear (forward and data-gradient via a transposed weight view), weight-gradient,
// segment row-sum -- are sequenced by host code below.  Concatenated inputs ([y, xg], [y, xc]) are read
// as two K-segments, never materialised; the per-ray latent add (y + z) is applied while loading.
#include "cnf_common.cuh"

namespace cnf {

struct LinSeg {
  const float* X; long long sX; int ld; int K;                 // [P, ld], first K columns
  const float* Z; long long sZ; int ldz; int rows_per;         // optional: x += Z[p / rows_per][k]
};
enum { PRO_NONE = 0, PRO_RELU = 1, PRO_SIGMOID = 2 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };
struct LinArgs {
  LinSeg s0, s1;
  const float* Mk; long long sMk; int ldm; int pro; float in_scale;  // segment-0 prologue: x *= in_scale * mask(Mk)
  const float* W; long long sW; int ldw; int wt;                // wt=0: w(n,k)=W[n*ldw+k]; wt=1: w(n,k)=W[k*ldw+n]
  const float* b; long long sb;
  float* Y; long long sY; int ldy; int N; int accumulate; int act; float out_scale;
  int P;
};
struct WgradArgs {
  LinSeg s0, s1;                                                // layer input (two K segments)
  const float* dY; long long sdY; int lddy; int N;              // upstream gradient [P, lddy]
  const float* Mk; long long sMk; int ldm; int pro; float in_scale;  // dY prologue (activation derivative)
  float* dW; long long sdW; int ldw;                            // [N, ldw] accumulated (atomics)
  float* db; long long sdb;                                     // [N] accumulated, may be null
  int P; int pts_per_block;
};

__device__ __forceinline__ float seg_load(const LinSeg& s, int k, int p, int kk) {
  float v = s.X[(long long)k * s.sX + (long long)p * s.ld + kk];
  if (s.Z) v += s.Z[(long long)k * s.sZ + (long long)(p / s.rows_per) * s.ldz + kk];
  return v;
}
__device__ __forceinline__ float pro_apply(float v, int pro, float scale, const float* Mk, long long off) {
  v *= scale;
  if (pro == PRO_RELU) v = (Mk[off] > 0.f) ? v : 0.f;
  else if (pro == PRO_SIGMOID) { const float c = Mk[off]; v *= c * (1.f - c); }
  return v;
}

constexpr int LT_P = 128, LT_N = 32, LT_K = 32;

__global__ void __launch_bounds__(LT_P)
linear_kernel(const LinArgs a) {
  const int k = blockIdx.z, p0 = blockIdx.x * LT_P, n0 = blockIdx.y * LT_N;
  __shared__ float xs[LT_K][LT_P + 1];
  __shared__ __align__(16) float ws[LT_K][LT_N];
  const int tid = threadIdx.x;
  const int K0 = a.s0.K, Kt = K0 + a.s1.K;
  float acc[LT_N];
#pragma unroll
  for (int n = 0; n < LT_N; ++n) acc[n] = 0.f;
  for (int kc = 0; kc < Kt; kc += LT_K) {
    for (int it = 0; it < LT_K; ++it) {
      const int idx = it * LT_P + tid;
      const int pp = idx / LT_K, kk = idx % LT_K;
      const int p = p0 + pp, kg = kc + kk;
      float v = 0.f;
      if (p < a.P && kg < Kt) {
        if (kg < K0) {
          v = seg_load(a.s0, k, p, kg);
          if (a.pro != PRO_NONE || a.in_scale != 1.0f)
            v = pro_apply(v, a.pro, a.in_scale, a.Mk, (long long)k * a.sMk + (long long)p * a.ldm + kg);
        } else {
          v = seg_load(a.s1, k, p, kg - K0);
        }
      }
      xs[kk][pp] = v;
    }
    for (int it = 0; it < (LT_K * LT_N) / LT_P; ++it) {
      const int idx = it * LT_P + tid;
      int n, kk;
      if (a.wt == 0) { n = idx / LT_K; kk = idx % LT_K; } else { kk = idx / LT_N; n = idx % LT_N; }
      const int ng = n0 + n, kg = kc + kk;
      float w = 0.f;
      if (ng < a.N && kg < Kt)
        w = a.wt == 0 ? a.W[(long long)k * a.sW + (long long)ng * a.ldw + kg]
                      : a.W[(long long)k * a.sW + (long long)kg * a.ldw + ng];
      ws[kk][n] = w;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < LT_K; ++kk) {
      const float x = xs[kk][tid];
      const float4* wr = reinterpret_cast<const float4*>(&ws[kk][0]);
#pragma unroll
      for (int q = 0; q < LT_N / 4; ++q) {
        const float4 w4 = wr[q];
        acc[q * 4 + 0] = fmaf(x, w4.x, acc[q * 4 + 0]);
        acc[q * 4 + 1] = fmaf(x, w4.y, acc[q * 4 + 1]);
        acc[q * 4 + 2] = fmaf(x, w4.z, acc[q * 4 + 2]);
        acc[q * 4 + 3] = fmaf(x, w4.w, acc[q * 4 + 3]);
      }
    }
    __syncthreads();
  }
  const int p = p0 + tid;
  if (p >= a.P) return;
  float* yrow = a.Y + (long long)k * a.sY + (long long)p * a.ldy;
#pragma unroll
  for (int n = 0; n < LT_N; ++n) {
    const int ng = n0 + n;
    if (ng < a.N) {
      float v = acc[n];
      if (a.b) v += a.b[(long long)k * a.sb + ng];
      if (a.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (a.act == ACT_SIGMOID) v = 1.0f / (1.0f + expf(-v));
      v *= a.out_scale;
      if (a.accumulate) yrow[ng] += v; else yrow[ng] = v;
    }
  }
}

constexpr int WG_T = 128;

__global__ void __launch_bounds__(WG_T)
wgrad_kernel(const WgradArgs a, int n_tiles) {
  const int k = blockIdx.z / n_tiles, nt = blockIdx.z % n_tiles;
  const int n0 = nt * 32, k0 = blockIdx.y * 32;
  const int pbeg = blockIdx.x * a.pts_per_block, pend = min(a.P, pbeg + a.pts_per_block);
  __shared__ float xs[32][33];
  __shared__ __align__(16) float dys[32][32];
  const int tid = threadIdx.x, kk = tid & 31, ng = tid >> 5;
  const int K0 = a.s0.K, Kt = K0 + a.s1.K;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  float accb = 0.f;
  const bool do_bias = (a.db != nullptr) && blockIdx.y == 0 && ng == 0;
  for (int pc = pbeg; pc < pend; pc += 32) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * WG_T + tid;
      const int pp = idx >> 5, c = idx & 31;
      const int p = pc + pp;
      float xv = 0.f, dv = 0.f;
      if (p < pend) {
        const int kg = k0 + c;
        if (kg < K0) xv = seg_load(a.s0, k, p, kg);
        else if (kg < Kt) xv = seg_load(a.s1, k, p, kg - K0);
        const int n = n0 + c;
        if (n < a.N) {
          dv = a.dY[(long long)k * a.sdY + (long long)p * a.lddy + n];
          dv = pro_apply(dv, a.pro, a.in_scale, a.Mk, (long long)k * a.sMk + (long long)p * a.ldm + n);
        }
      }
      xs[pp][c] = xv;
      dys[pp][c] = dv;
    }
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < 32; ++pp) {
      const float x = xs[pp][kk];
      const float4 d0 = *reinterpret_cast<const float4*>(&dys[pp][ng * 8]);
      const float4 d1 = *reinterpret_cast<const float4*>(&dys[pp][ng * 8 + 4]);
      acc[0] = fmaf(x, d0.x, acc[0]); acc[1] = fmaf(x, d0.y, acc[1]);
      acc[2] = fmaf(x, d0.z, acc[2]); acc[3] = fmaf(x, d0.w, acc[3]);
      acc[4] = fmaf(x, d1.x, acc[4]); acc[5] = fmaf(x, d1.y, acc[5]);
      acc[6] = fmaf(x, d1.z, acc[6]); acc[7] = fmaf(x, d1.w, acc[7]);
      if (do_bias) accb += dys[pp][kk];
    }
    __syncthreads();
  }
  const int kg = k0 + kk;
  if (kg < Kt) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + ng * 8 + j;
      if (n < a.N) atomicAdd(&a.dW[(long long)k * a.sdW + (long long)n * a.ldw + kg], acc[j]);
    }
  }
  if (do_bias && n0 + kk < a.N) atomicAdd(&a.db[(long long)k * a.sdb + n0 + kk], accb);
}

// out[k][r][c] = sum_{j < rows_per} in[k][(r*rows_per + j)][c]
__global__ void rowsum_kernel(int n_rows, int rows_per, int C, const float* __restrict__ in, long long s_in, int ld_in,
                              float* __restrict__ out, long long s_out, int ld_out, int P) {
  const int k = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n_rows * C) return;
  const int r = (int)(e / C), c = (int)(e % C);
  float acc = 0.f;
  const long long pb = (long long)r * rows_per;
  for (int j = 0; j < rows_per && pb + j < P; ++j) acc += in[(long long)k * s_in + (pb + j) * ld_in + c];
  out[(long long)k * s_out + (long long)r * ld_out + c] = acc;
}

// ------------------------------------------------------------------------------------------------
// host-side sequencing
// ------------------------------------------------------------------------------------------------
struct Launcher {
  cudaStream_t st; int K; int P;
  void lin(LinArgs a) const {
    a.P = a.P ? a.P : P;
    dim3 grid((a.P + LT_P - 1) / LT_P, (a.N + LT_N - 1) / LT_N, K);
    linear_kernel<<<grid, LT_P, 0, st>>>(a);
  }
  void wgrad(WgradArgs a) const {
    a.P = a.P ? a.P : P;
    const int Kt = a.s0.K + a.s1.K;
    int ppb = 2048;
    while (ppb > 64 && (a.P + ppb - 1) / ppb * ((Kt + 31) / 32) * ((a.N + 31) / 32) * K < 296) ppb >>= 1;
    a.pts_per_block = ppb;
    const int n_tiles = (a.N + 31) / 32;
    dim3 grid((a.P + ppb - 1) / ppb, (Kt + 31) / 32, n_tiles * K);
    wgrad_kernel<<<grid, WG_T, 0, st>>>(a, n_tiles);
  }
  void rowsum(int n_rows, int rows_per, int C, const float* in, long long s_in, int ld_in, float* out, long long s_out,
              int ld_out, int Ptot) const {
    const long long n = (long long)n_rows * C;
    rowsum_kernel<<<dim3((unsigned)((n + 255) / 256), K), 256, 0, st>>>(n_rows, rows_per, C, in, s_in, ld_in, out, s_out,
                                                                          ld_out, Ptot);
  }
};

static LinSeg seg(const float* X, long long sX, int ld, int K) { return LinSeg{X, sX, ld, K, nullptr, 0, 0, 1}; }
static LinSeg segz(const float* X, long long sX, int ld, int K, const float* Z, long long sZ, int ldz, int rows_per) {
  return LinSeg{X, sX, ld, K, Z, sZ, ldz, rows_per};
}
static LinSeg noseg() { return LinSeg{nullptr, 0, 0, 0, nullptr, 0, 0, 1}; }

static LinArgs mk_lin(LinSeg s0, LinSeg s1, const float* W, long long sW, int ldw, int wt, const float* b, float* Y,
                      long long sY, int ldy, int N, int act) {
  LinArgs a{};
  a.s0 = s0; a.s1 = s1; a.Mk = nullptr; a.sMk = 0; a.ldm = 0; a.pro = PRO_NONE; a.in_scale = 1.f;
  a.W = W; a.sW = sW; a.ldw = ldw; a.wt = wt; a.b = b; a.sb = sW; a.Y = Y; a.sY = sY; a.ldy = ldy; a.N = N;
  a.accumulate = 0; a.act = act; a.out_scale = 1.f; a.P = 0;
  return a;
}

struct CodeNerfLayout {
  int W, L; size_t o[28]; size_t total;
  CodeNerfLayout(int W_, int L_) : W(W_), L(L_) {
    const size_t E1 = CNF_EMB1, E2 = CNF_EMB2, w = W_, l = L_;
    const size_t sz[28] = {w * E1, w, w * l, w, w * w, w, w * l, w, w * w, w, w * (w + E1), w, w * l, w, w * w, w,
                           w, 1, w * (w + E2), w, w * w, w, w * l, w, (w / 2) * w, w / 2, 3 * (w / 2), 3};
    size_t off = 0;
    for (int i = 0; i < 28; ++i) { o[i] = off; off += sz[i]; }
    total = off;
  }
};

struct OccMapLayout {
  int H; size_t o[14]; size_t total;
  explicit OccMapLayout(int H_) : H(H_) {
    const size_t E1 = CNF_EMB1, E2 = CNF_EMB2, h = H_;
    const size_t sz[14] = {h * E1, h, h * h, h, h * (h + E1), h, h * h, h, h, 1, h * (h + E2), h, 3 * h, 3};
    size_t off = 0;
    for (int i = 0; i < 14; ++i) { o[i] = off; off += sz[i]; }
    total = off;
  }
};

// per-category workspace layout (floats). Saved activations first, backward scratch after.
struct CodeNerfWs {
  size_t zl1, zcl, zl2, ztl, r1, r2, r3, r4, y5, h1, h2, h3, gA, gB, gzl, per_cat;
  CodeNerfWs(size_t P, size_t NR, size_t W) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };
    zl1 = take(NR * W); zcl = take(NR * W); zl2 = take(NR * W); ztl = take(NR * W);
    r1 = take(P * W); r2 = take(P * W); r3 = take(P * W); r4 = take(P * W); y5 = take(P * W);
    h1 = take(P * W); h2 = take(P * W); h3 = take(P * (W / 2));
    gA = take(P * W); gB = take(P * W); gzl = take(NR * W);
    per_cat = o;
  }
};
struct OccMapWs {
  size_t f1, f2, f3, f4, hc, gA, gB, per_cat;
  OccMapWs(size_t P, size_t H) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };
    f1 = take(P * H); f2 = take(P * H); f3 = take(P * H); f4 = take(P * H); hc = take(P * H);
    gA = take(P * H); gB = take(P * H);
    per_cat = o;
  }
};

}  // namespace cnf

using namespace cnf;

extern "C" size_t cnf_codenerf_param_count(int W, int L) { return CodeNerfLayout(W, L).total; }
extern "C" size_t cnf_codenerf_workspace_floats(int P, int NR, int W) { return CodeNerfWs(P, NR, W).per_cat; }
extern "C" size_t cnf_occmap_param_count(int H) { return OccMapLayout(H).total; }
extern "C" size_t cnf_occmap_workspace_floats(int P, int H) { return OccMapWs(P, H).per_cat; }

extern "C" int cnf_codenerf_fwd(int K, int P, int NR, int ppr, int W, int L, const float* params, const float* emb,
                                const float* shape_lat, const float* tex_lat, float* sigma, float* rgb, float* ws,
                                size_t ws_floats, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && P > 0 && NR > 0 && ppr > 0 && W >= 2 && W % 2 == 0 && L > 0, "cnf_codenerf_fwd: bad sizes");
  CNF_REQUIRE((long long)NR * ppr >= P, "cnf_codenerf_fwd: n_rows*pts_per_row < n_pts");
  CNF_REQUIRE(params && emb && shape_lat && tex_lat && sigma && rgb && ws, "cnf_codenerf_fwd: null pointer");
  const CodeNerfLayout lay(W, L);
  const CodeNerfWs wl(P, NR, W);
  if (ws_floats < wl.per_cat * (size_t)K) { set_error("cnf_codenerf_fwd: workspace too small"); return CNF_ERR_WORKSPACE; }
  const long long sP = (long long)lay.total, sE = (long long)P * CNF_EMB_DIM, sWs = (long long)wl.per_cat,
                  sLat = (long long)NR * L;
  Launcher ln{as_stream(stream), K, P};
  auto prm = [&](int i) { return params + lay.o[i]; };
  const float* xg = emb; const float* xc = emb + CNF_EMB1;
  const int E1 = CNF_EMB1, E2 = CNF_EMB2, ED = CNF_EMB_DIM;
  // latent projections (per latent row): relu(W s + b)   model.py:62,66,80
  { LinArgs a = mk_lin(seg(shape_lat, sLat, L, L), noseg(), prm(2), sP, L, 0, prm(3), ws + wl.zl1, sWs, W, W, ACT_RELU); a.P = NR; ln.lin(a); }
  { LinArgs a = mk_lin(seg(shape_lat, sLat, L, L), noseg(), prm(12), sP, L, 0, prm(13), ws + wl.zcl, sWs, W, W, ACT_RELU); a.P = NR; ln.lin(a); }
  { LinArgs a = mk_lin(seg(shape_lat, sLat, L, L), noseg(), prm(6), sP, L, 0, prm(7), ws + wl.zl2, sWs, W, W, ACT_RELU); a.P = NR; ln.lin(a); }
  { LinArgs a = mk_lin(seg(tex_lat, sLat, L, L), noseg(), prm(22), sP, L, 0, prm(23), ws + wl.ztl, sWs, W, W, ACT_RELU); a.P = NR; ln.lin(a); }
  // r1 = relu(Wxyz xg + b)                                 model.py:59
  ln.lin(mk_lin(seg(xg, sE, ED, E1), noseg(), prm(0), sP, E1, 0, prm(1), ws + wl.r1, sWs, W, W, ACT_RELU));
  // r2 = relu(Ws1 (r1 + zl1) + b)                          model.py:66-68
  ln.lin(mk_lin(segz(ws + wl.r1, sWs, W, W, ws + wl.zl1, sWs, W, ppr), noseg(), prm(4), sP, W, 0, prm(5), ws + wl.r2, sWs, W, W, ACT_RELU));
  // r3 = relu(Wcat [r2 + zcl, xg] + b)                     model.py:62-65
  ln.lin(mk_lin(segz(ws + wl.r2, sWs, W, W, ws + wl.zcl, sWs, W, ppr), seg(xg, sE, ED, E1), prm(10), sP, W + E1, 0, prm(11), ws + wl.r3, sWs, W, W, ACT_RELU));
  // r4 = relu(Ws2 (r3 + zl2) + b)
  ln.lin(mk_lin(segz(ws + wl.r3, sWs, W, W, ws + wl.zl2, sWs, W, ppr), noseg(), prm(8), sP, W, 0, prm(9), ws + wl.r4, sWs, W, W, ACT_RELU));
  // y5 = Wes r4 + b (no activation)                        model.py:70
  ln.lin(mk_lin(seg(ws + wl.r4, sWs, W, W), noseg(), prm(14), sP, W, 0, prm(15), ws + wl.y5, sWs, W, W, ACT_NONE));
  // sigma = 10 (Wsig y5 + b)                               model.py:71,75
  { LinArgs a = mk_lin(seg(ws + wl.y5, sWs, W, W), noseg(), prm(16), sP, W, 0, prm(17), sigma, (long long)P, 1, 1, ACT_NONE); a.out_scale = 10.f; ln.lin(a); }
  // h1 = relu(Wvd [y5, xc] + b)                            model.py:77-78
  ln.lin(mk_lin(seg(ws + wl.y5, sWs, W, W), seg(xc, sE, ED, E2), prm(18), sP, W + E2, 0, prm(19), ws + wl.h1, sWs, W, W, ACT_RELU));
  // h2 = relu(Wt1 (h1 + ztl) + b)                          model.py:80-82
  ln.lin(mk_lin(segz(ws + wl.h1, sWs, W, W, ws + wl.ztl, sWs, W, ppr), noseg(), prm(20), sP, W, 0, prm(21), ws + wl.h2, sWs, W, W, ACT_RELU));
  // h3 = relu(Wr1 h2 + b); rgb = sigmoid(Wr2 h3 + b)       model.py:54,83
  ln.lin(mk_lin(seg(ws + wl.h2, sWs, W, W), noseg(), prm(24), sP, W, 0, prm(25), ws + wl.h3, sWs, W / 2, W / 2, ACT_RELU));
  ln.lin(mk_lin(seg(ws + wl.h3, sWs, W / 2, W / 2), noseg(), prm(26), sP, W / 2, 0, prm(27), rgb, (long long)P * 3, 3, 3, ACT_SIGMOID));
  return check_launch("cnf_codenerf_fwd");
}

namespace cnf {
// one dense layer's backward: weight/bias gradient + data gradient(s)
struct BwdCtx {
  Launcher ln; long long sP, sWs;
  const float* params; float* g_params;
  // dY = dA * in_scale * mask(Mk)   ;  layer y = act(W [s0, s1] + b)
  void layer(int wi, int N, LinSeg s0, LinSeg s1, const float* dA, long long sdA, int lddA, const float* Mk, long long sMk,
             int ldm, int pro, float in_scale,
             float* d0, long long sd0, int ldd0, int acc0,      // gradient wrt segment 0 (nullable)
             float* d1, long long sd1, int ldd1, int acc1,      // gradient wrt segment 1 (nullable)
             size_t w_off, size_t b_off, int Pn = 0) const {
    const int Kt = s0.K + s1.K;
    WgradArgs w{};
    w.s0 = s0; w.s1 = s1; w.dY = dA; w.sdY = sdA; w.lddy = lddA; w.N = N; w.Mk = Mk; w.sMk = sMk; w.ldm = ldm;
    w.pro = pro; w.in_scale = in_scale; w.dW = g_params + w_off; w.sdW = sP; w.ldw = Kt; w.db = g_params + b_off;
    w.sdb = sP; w.P = Pn;
    ln.wgrad(w);
    (void)wi;
    if (d0) {
      LinArgs a = mk_lin(seg(dA, sdA, lddA, N), noseg(), params + w_off, sP, Kt, 1, nullptr, d0, sd0, ldd0, s0.K, ACT_NONE);
      a.Mk = Mk; a.sMk = sMk; a.ldm = ldm; a.pro = pro; a.in_scale = in_scale; a.accumulate = acc0; a.P = Pn;
      ln.lin(a);
    }
    if (d1) {
      LinArgs a = mk_lin(seg(dA, sdA, lddA, N), noseg(), params + w_off + s0.K, sP, Kt, 1, nullptr, d1, sd1, ldd1, s1.K, ACT_NONE);
      a.Mk = Mk; a.sMk = sMk; a.ldm = ldm; a.pro = pro; a.in_scale = in_scale; a.accumulate = acc1; a.P = Pn;
      ln.lin(a);
    }
  }
};
}  // namespace cnf

extern "C" int cnf_codenerf_bwd(int K, int P, int NR, int ppr, int W, int L, const float* params, const float* emb,
                                const float* shape_lat, const float* tex_lat, const float* rgb, const float* g_sigma,
                                const float* g_rgb, float* ws, size_t ws_floats, float* g_params, float* g_emb,
                                float* g_shape_lat, float* g_tex_lat, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && P > 0 && NR > 0 && ppr > 0 && W >= 2 && W % 2 == 0 && L > 0, "cnf_codenerf_bwd: bad sizes");
  CNF_REQUIRE(params && emb && shape_lat && tex_lat && rgb && g_sigma && g_rgb && ws && g_params, "cnf_codenerf_bwd: null pointer");
  const CodeNerfLayout lay(W, L);
  const CodeNerfWs wl(P, NR, W);
  if (ws_floats < wl.per_cat * (size_t)K) { set_error("cnf_codenerf_bwd: workspace too small"); return CNF_ERR_WORKSPACE; }
  const long long sP = (long long)lay.total, sE = (long long)P * CNF_EMB_DIM, sWs = (long long)wl.per_cat,
                  sLat = (long long)NR * L;
  cudaStream_t st = as_stream(stream);
  BwdCtx c{Launcher{st, K, P}, sP, sWs, params, g_params};
  const float* xg = emb; const float* xc = emb + CNF_EMB1;
  const int E1 = CNF_EMB1, E2 = CNF_EMB2, ED = CNF_EMB_DIM, Wh = W / 2;
  float* gA = ws + wl.gA; float* gB = ws + wl.gB; float* gz = ws + wl.gzl;
  float* ge_g = g_emb; float* ge_c = g_emb ? g_emb + E1 : nullptr;
  // rgb = sigmoid(Wr2 h3 + b): dY = g_rgb * c(1-c)  -> d h3 (gA, width Wh)
  c.layer(26, 3, seg(ws + wl.h3, sWs, Wh, Wh), noseg(), g_rgb, (long long)P * 3, 3, rgb, (long long)P * 3, 3, PRO_SIGMOID, 1.f,
          gA, sWs, Wh, 0, nullptr, 0, 0, 0, lay.o[26], lay.o[27]);
  // h3 = relu(Wr1 h2 + b) -> d h2 (gB)
  c.layer(24, Wh, seg(ws + wl.h2, sWs, W, W), noseg(), gA, sWs, Wh, ws + wl.h3, sWs, Wh, PRO_RELU, 1.f,
          gB, sWs, W, 0, nullptr, 0, 0, 0, lay.o[24], lay.o[25]);
  // h2 = relu(Wt1 (h1 + ztl) + b) -> d(h1+ztl) (gA); d ztl = rowsum
  c.layer(20, W, segz(ws + wl.h1, sWs, W, W, ws + wl.ztl, sWs, W, ppr), noseg(), gB, sWs, W, ws + wl.h2, sWs, W, PRO_RELU, 1.f,
          gA, sWs, W, 0, nullptr, 0, 0, 0, lay.o[20], lay.o[21]);
  c.ln.rowsum(NR, ppr, W, gA, sWs, W, gz, sWs, W, P);
  // texture latent layer: ztl = relu(Wtl1 t + b)
  c.layer(22, W, seg(tex_lat, sLat, L, L), noseg(), gz, sWs, W, ws + wl.ztl, sWs, W, PRO_RELU, 1.f,
          g_tex_lat, sLat, L, 0, nullptr, 0, 0, 0, lay.o[22], lay.o[23], NR);
  // h1 = relu(Wvd [y5, xc] + b) -> d y5 (gB), d xc (g_emb[:,87:])
  c.layer(18, W, seg(ws + wl.y5, sWs, W, W), seg(xc, sE, ED, E2), gA, sWs, W, ws + wl.h1, sWs, W, PRO_RELU, 1.f,
          gB, sWs, W, 0, ge_c, sE, ED, 0, lay.o[18], lay.o[19]);
  // sigma = 10 (Wsig y5 + b): dY = 10 g_sigma -> accumulate into d y5 (gB)
  c.layer(16, 1, seg(ws + wl.y5, sWs, W, W), noseg(), g_sigma, (long long)P, 1, nullptr, 0, 0, PRO_NONE, 10.f,
          gB, sWs, W, 1, nullptr, 0, 0, 0, lay.o[16], lay.o[17]);
  // y5 = Wes r4 + b (linear) -> d r4 (gA)
  c.layer(14, W, seg(ws + wl.r4, sWs, W, W), noseg(), gB, sWs, W, nullptr, 0, 0, PRO_NONE, 1.f,
          gA, sWs, W, 0, nullptr, 0, 0, 0, lay.o[14], lay.o[15]);
  // r4 = relu(Ws2 (r3 + zl2) + b) -> d(r3+zl2) (gB); d zl2 = rowsum -> shape latent layer 2
  c.layer(8, W, segz(ws + wl.r3, sWs, W, W, ws + wl.zl2, sWs, W, ppr), noseg(), gA, sWs, W, ws + wl.r4, sWs, W, PRO_RELU, 1.f,
          gB, sWs, W, 0, nullptr, 0, 0, 0, lay.o[8], lay.o[9]);
  c.ln.rowsum(NR, ppr, W, gB, sWs, W, gz, sWs, W, P);
  c.layer(6, W, seg(shape_lat, sLat, L, L), noseg(), gz, sWs, W, ws + wl.zl2, sWs, W, PRO_RELU, 1.f,
          g_shape_lat, sLat, L, 0, nullptr, 0, 0, 0, lay.o[6], lay.o[7], NR);
  // r3 = relu(Wcat [r2 + zcl, xg] + b) -> d(r2+zcl) (gA), d xg (g_emb[:, :87], first writer)
  c.layer(10, W, segz(ws + wl.r2, sWs, W, W, ws + wl.zcl, sWs, W, ppr), seg(xg, sE, ED, E1), gB, sWs, W, ws + wl.r3, sWs, W, PRO_RELU, 1.f,
          gA, sWs, W, 0, ge_g, sE, ED, 0, lay.o[10], lay.o[11]);
  c.ln.rowsum(NR, ppr, W, gA, sWs, W, gz, sWs, W, P);
  c.layer(12, W, seg(shape_lat, sLat, L, L), noseg(), gz, sWs, W, ws + wl.zcl, sWs, W, PRO_RELU, 1.f,
          g_shape_lat, sLat, L, 1, nullptr, 0, 0, 0, lay.o[12], lay.o[13], NR);
  // r2 = relu(Ws1 (r1 + zl1) + b) -> d(r1+zl1) (gB)
  c.layer(4, W, segz(ws + wl.r1, sWs, W, W, ws + wl.zl1, sWs, W, ppr), noseg(), gA, sWs, W, ws + wl.r2, sWs, W, PRO_RELU, 1.f,
          gB, sWs, W, 0, nullptr, 0, 0, 0, lay.o[4], lay.o[5]);
  c.ln.rowsum(NR, ppr, W, gB, sWs, W, gz, sWs, W, P);
  c.layer(2, W, seg(shape_lat, sLat, L, L), noseg(), gz, sWs, W, ws + wl.zl1, sWs, W, PRO_RELU, 1.f,
          g_shape_lat, sLat, L, 1, nullptr, 0, 0, 0, lay.o[2], lay.o[3], NR);
  // r1 = relu(Wxyz xg + b) -> d xg accumulated into g_emb[:, :87]
  c.layer(0, W, seg(xg, sE, ED, E1), noseg(), gB, sWs, W, ws + wl.r1, sWs, W, PRO_RELU, 1.f,
          ge_g, sE, ED, 1, nullptr, 0, 0, 0, lay.o[0], lay.o[1]);
  return check_launch("cnf_codenerf_bwd");
}

extern "C" int cnf_occmap_fwd(int K, int P, int H, const float* params, const float* emb, float* alpha, float* rgb,
                              float* ws, size_t ws_floats, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && P > 0 && H > 0, "cnf_occmap_fwd: bad sizes");
  CNF_REQUIRE(params && emb && alpha && rgb && ws, "cnf_occmap_fwd: null pointer");
  const OccMapLayout lay(H);
  const OccMapWs wl(P, H);
  if (ws_floats < wl.per_cat * (size_t)K) { set_error("cnf_occmap_fwd: workspace too small"); return CNF_ERR_WORKSPACE; }
  const long long sP = (long long)lay.total, sE = (long long)P * CNF_EMB_DIM, sWs = (long long)wl.per_cat;
  Launcher ln{as_stream(stream), K, P};
  auto prm = [&](int i) { return params + lay.o[i]; };
  const float* xg = emb; const float* xc = emb + CNF_EMB1;
  const int E1 = CNF_EMB1, E2 = CNF_EMB2, ED = CNF_EMB_DIM;
  ln.lin(mk_lin(seg(xg, sE, ED, E1), noseg(), prm(0), sP, E1, 0, prm(1), ws + wl.f1, sWs, H, H, ACT_RELU));              // model.py:129
  ln.lin(mk_lin(seg(ws + wl.f1, sWs, H, H), noseg(), prm(2), sP, H, 0, prm(3), ws + wl.f2, sWs, H, H, ACT_RELU));         // :130
  ln.lin(mk_lin(seg(ws + wl.f2, sWs, H, H), seg(xg, sE, ED, E1), prm(4), sP, H + E1, 0, prm(5), ws + wl.f3, sWs, H, H, ACT_RELU));  // :133-134
  ln.lin(mk_lin(seg(ws + wl.f3, sWs, H, H), noseg(), prm(6), sP, H, 0, prm(7), ws + wl.f4, sWs, H, H, ACT_RELU));         // :137
  { LinArgs a = mk_lin(seg(ws + wl.f4, sWs, H, H), noseg(), prm(8), sP, H, 0, prm(9), alpha, (long long)P, 1, 1, ACT_NONE); a.out_scale = 10.f; ln.lin(a); }  // :141,147
  ln.lin(mk_lin(seg(ws + wl.f4, sWs, H, H), seg(xc, sE, ED, E2), prm(10), sP, H + E2, 0, prm(11), ws + wl.hc, sWs, H, H, ACT_RELU));  // :151
  ln.lin(mk_lin(seg(ws + wl.hc, sWs, H, H), noseg(), prm(12), sP, H, 0, prm(13), rgb, (long long)P * 3, 3, 3, ACT_SIGMOID));           // :152-153
  return check_launch("cnf_occmap_fwd");
}

extern "C" int cnf_occmap_bwd(int K, int P, int H, const float* params, const float* emb, const float* rgb,
                              const float* g_alpha, const float* g_rgb, float* ws, size_t ws_floats, float* g_params,
                              float* g_emb, cnf_stream_t stream) {
  CNF_REQUIRE(K > 0 && P > 0 && H > 0, "cnf_occmap_bwd: bad sizes");
  CNF_REQUIRE(params && emb && rgb && g_alpha && g_rgb && ws && g_params, "cnf_occmap_bwd: null pointer");
  const OccMapLayout lay(H);
  const OccMapWs wl(P, H);
  if (ws_floats < wl.per_cat * (size_t)K) { set_error("cnf_occmap_bwd: workspace too small"); return CNF_ERR_WORKSPACE; }
  const long long sP = (long long)lay.total, sE = (long long)P * CNF_EMB_DIM, sWs = (long long)wl.per_cat;
  BwdCtx c{Launcher{as_stream(stream), K, P}, sP, sWs, params, g_params};
  const float* xg = emb; const float* xc = emb + CNF_EMB1;
  const int E1 = CNF_EMB1, E2 = CNF_EMB2, ED = CNF_EMB_DIM;
  float* gA = ws + wl.gA; float* gB = ws + wl.gB;
  float* ge_g = g_emb; float* ge_c = g_emb ? g_emb + E1 : nullptr;
  // color = sigmoid(Woc hc + b) -> d hc (gA)
  c.layer(12, 3, seg(ws + wl.hc, sWs, H, H), noseg(), g_rgb, (long long)P * 3, 3, rgb, (long long)P * 3, 3, PRO_SIGMOID, 1.f,
          gA, sWs, H, 0, nullptr, 0, 0, 0, lay.o[12], lay.o[13]);
  // hc = relu(Wc [f4, xc] + b) -> d f4 (gB), d xc
  c.layer(10, H, seg(ws + wl.f4, sWs, H, H), seg(xc, sE, ED, E2), gA, sWs, H, ws + wl.hc, sWs, H, PRO_RELU, 1.f,
          gB, sWs, H, 0, ge_c, sE, ED, 0, lay.o[10], lay.o[11]);
  // alpha = 10 (Wa f4 + b) -> accumulate into d f4
  c.layer(8, 1, seg(ws + wl.f4, sWs, H, H), noseg(), g_alpha, (long long)P, 1, nullptr, 0, 0, PRO_NONE, 10.f,
          gB, sWs, H, 1, nullptr, 0, 0, 0, lay.o[8], lay.o[9]);
  // f4 = relu(Wm2 f3 + b) -> d f3 (gA)
  c.layer(6, H, seg(ws + wl.f3, sWs, H, H), noseg(), gB, sWs, H, ws + wl.f4, sWs, H, PRO_RELU, 1.f,
          gA, sWs, H, 0, nullptr, 0, 0, 0, lay.o[6], lay.o[7]);
  // f3 = relu(Wcat [f2, xg] + b) -> d f2 (gB), d xg (first writer)
  c.layer(4, H, seg(ws + wl.f2, sWs, H, H), seg(xg, sE, ED, E1), gA, sWs, H, ws + wl.f3, sWs, H, PRO_RELU, 1.f,
          gB, sWs, H, 0, ge_g, sE, ED, 0, lay.o[4], lay.o[5]);
  // f2 = relu(Wm1 f1 + b) -> d f1 (gA)
  c.layer(2, H, seg(ws + wl.f1, sWs, H, H), noseg(), gB, sWs, H, ws + wl.f2, sWs, H, PRO_RELU, 1.f,
          gA, sWs, H, 0, nullptr, 0, 0, 0, lay.o[2], lay.o[3]);
  // f1 = relu(Win xg + b) -> d xg accumulate
  c.layer(0, H, seg(xg, sE, ED, E1), noseg(), gA, sWs, H, ws + wl.f1, sWs, H, PRO_RELU, 1.f,
          ge_g, sE, ED, 1, nullptr, 0, 0, 0, lay.o[0], lay.o[1]);
  return check_launch("cnf_occmap_bwd");
}
