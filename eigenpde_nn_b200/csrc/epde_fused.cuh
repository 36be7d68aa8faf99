// Fused FFMA2 kernel family for nets  d -> H -> H -> H -> 1  with Tanh (the shape of both
// shipped examples, arch_size_list = 20,20,20).  See DESIGN.md "Kernels" for the derivation.
//
//   k_pack   : flat parameters -> per-net packed shared-memory image (+ K = W1 diag(c) W1^T)
//   k_pass1  : forward + input-Jacobian chain + energy -> per-CTA fp64 partial sums, y[B][k]
//   k_pass2  : recompute forward/Jacobian, reverse sweeps, weight-gradient outer products
//   k_fin2   : fixed-order fp64 reduction of the per-CTA gradient partials -> flat gradient
//
// Thread = one PAIR of samples (the two lanes of every FFMA2); weights are broadcast from
// shared memory (one LDS.128 per four FFMA2).
#pragma once
#include "epde_common.cuh"
#include "epde_coef.cuh"

namespace epde {

// ---- packed per-net weight image (floats) ---------------------------------------------------
template <int H>
struct PackLayout {
  int d_pad;
  __host__ __device__ explicit PackLayout(int dp) : d_pad(dp) {}
  // W1T[j][o] = W1[o][j]; rows j >= d are zero and the row count is padded to a multiple of 16 so
  // that the unguarded 8- / 16-column blocks of layer 1 never leave the array
  __host__ __device__ int w1rows() const { return (d_pad + 15) / 16 * 16; }
  __host__ __device__ int W1c() const { return 0; }
  __host__ __device__ int b1() const { return w1rows() * H; }
  __host__ __device__ int W2T() const { return b1() + H; }          // [i][o] = W2[o][i]
  __host__ __device__ int W2() const { return W2T() + H * H; }      // [o][i]
  __host__ __device__ int b2() const { return W2() + H * H; }
  __host__ __device__ int W3T() const { return b2() + H; }
  __host__ __device__ int W3() const { return W3T() + H * H; }
  __host__ __device__ int b3() const { return W3() + H * H; }
  __host__ __device__ int w4() const { return b3() + H; }
  __host__ __device__ int b4() const { return w4() + H; }           // 4 floats, [0] used
  __host__ __device__ int Km() const { return b4() + 4; }           // [H][H] symmetric
  __host__ __device__ int total() const { return Km() + H * H; }
};

// ---- per-net gradient accumulator image (floats) --------------------------------------------
template <int H>
struct GradLayout {
  int d_pad;
  __host__ __device__ explicit GradLayout(int dp) : d_pad(dp) {}
  __host__ __device__ int gW1() const { return 0; }                 // [H][d_pad]  (z-part only)
  __host__ __device__ int gb1() const { return H * d_pad; }
  __host__ __device__ int gW2() const { return gb1() + H; }         // [H][H]
  __host__ __device__ int gb2() const { return gW2() + H * H; }
  __host__ __device__ int gW3() const { return gb2() + H; }
  __host__ __device__ int gb3() const { return gW3() + H * H; }
  __host__ __device__ int gw4() const { return gb3() + H; }
  __host__ __device__ int gb4() const { return gw4() + H; }         // 4 floats, [0] used
  __host__ __device__ int Gam() const { return gb4() + 4; }         // [H][H] = sum ebar g1 g1^T
  __host__ __device__ int total() const { return Gam() + H * H; }
};

// flat (model.parameters()) size of one net
template <int H>
__host__ __device__ inline int flat_net_size(int d) { return H * d + H + 2 * (H * H + H) + H + 1; }

// ---------------------------------------------------------------------------------------------
// k_pack: one CTA per net
// ---------------------------------------------------------------------------------------------
template <int H>
__global__ void k_pack(const float* __restrict__ params, const float* __restrict__ diag_coeff, int d, int d_pad,
                       float* __restrict__ pw) {
  const PackLayout<H> L(d_pad);
  const int net = blockIdx.x;
  const float* p = params + (size_t)net * flat_net_size<H>(d);
  float* o = pw + (size_t)net * L.total();
  const float* W1 = p;
  const float* b1 = W1 + H * d;
  const float* W2 = b1 + H;
  const float* b2 = W2 + H * H;
  const float* W3 = b2 + H;
  const float* b3 = W3 + H * H;
  const float* W4 = b3 + H;
  const float* b4 = W4 + H;
  for (int e = threadIdx.x; e < L.w1rows() * H; e += blockDim.x) {
    const int oo = e % H, j = e / H;
    o[L.W1c() + e] = (j < d) ? W1[oo * d + j] : 0.f;
  }
  for (int e = threadIdx.x; e < H * H; e += blockDim.x) {
    const int a = e / H, b = e % H;
    o[L.W2T() + e] = W2[b * H + a];
    o[L.W2() + e] = W2[e];
    o[L.W3T() + e] = W3[b * H + a];
    o[L.W3() + e] = W3[e];
    double s = 0.0;
    for (int j = 0; j < d; j++) s += (double)diag_coeff[j] * (double)W1[a * d + j] * (double)W1[b * d + j];
    o[L.Km() + e] = (float)s;
  }
  for (int e = threadIdx.x; e < H; e += blockDim.x) {
    o[L.b1() + e] = b1[e]; o[L.b2() + e] = b2[e]; o[L.b3() + e] = b3[e]; o[L.w4() + e] = W4[e];
  }
  if (threadIdx.x < 4) o[L.b4() + threadIdx.x] = (threadIdx.x == 0) ? b4[0] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// shared pieces of the per-sample-pair computation
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// k_gather: X[idx[n]][:] -> tile-major TRANSPOSED copy  xt[tile][j][256]  (+ wt[n] = w[idx[n]])
// ---------------------------------------------------------------------------------------------
// Row gathers by random indices are expensive inside the compute kernels (a warp-wide LDG.128
// over 32 different rows touches 32 cache lines); here each row is read once, cooperatively and
// coalesced, transposed through shared memory, and written so that a sample pair's x_j is ONE
// coalesced 8-byte load and a tile's row j is 1 KB contiguous (the layout of the k_pass2 rows).
constexpr int GT = 256;                       // samples per tile (= P2_T)
template <int DUMMY = 0>
__global__ void __launch_bounds__(256)
k_gather(const float* __restrict__ X, const float* __restrict__ w, const int64_t* __restrict__ idx, int64_t B,
         int d_pad, float* __restrict__ xt, float* __restrict__ wt, int tile128) {
  extern __shared__ float4 smem4[];
  float* ts = reinterpret_cast<float*>(smem4);           // [256][d_pad + 1]
  __shared__ int64_t srow[GT];
  const int ld = d_pad + 1, nq = d_pad >> 2, t = threadIdx.x;
  const int64_t tile = blockIdx.x, n = tile * GT + t;
  const bool v = n < B;
  const int64_t r = v ? (idx ? idx[n] : n) : 0;
  srow[t] = v ? r : -1;
  wt[n] = v ? (w ? w[r] : 1.f) : 0.f;
  __syncthreads();
#pragma unroll 4
  for (int c = t; c < GT * nq; c += 256) {
    const int row = c / nq, q = c - row * nq;
    const int64_t rr = srow[row];
    const float4 val = (rr >= 0) ? __ldg(reinterpret_cast<const float4*>(X + rr * d_pad) + q)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    float* dst = ts + row * ld + 4 * q;
    dst[0] = val.x; dst[1] = val.y; dst[2] = val.z; dst[3] = val.w;
  }
  __syncthreads();
  if (tile128) {      // tensor-core family: 128-sample tiles, xt[tile128][j][128] (one contiguous block per tile)
    float* out = xt + (size_t)(tile * 2 + (t >> 7)) * d_pad * 128 + (t & 127);
    for (int j = 0; j < d_pad; j++) out[(size_t)j * 128] = ts[t * ld + j];
    return;
  }
  float* out = xt + (size_t)tile * d_pad * GT;
  for (int j = 0; j < d_pad; j++) out[(size_t)j * GT + t] = ts[t * ld + j];
}

// k_gather128: the same for the tensor-core family: one CTA per 128-sample tile, 256 threads, 27 KB of shared memory
// -> 8 CTAs = 64 warps per SM (k_gather: 4 CTAs = 32 warps), more row loads in flight for the DRAM-bound gather.
// Output: xt[tile128][j][128] (one contiguous block per tile), wt[n].
template <int DUMMY = 0>
__device__ __forceinline__ void gather128_body(const float* __restrict__ X, const float* __restrict__ w,
                                               const int64_t* __restrict__ idx, int64_t B, int d_pad, float* __restrict__ xt,
                                               float* __restrict__ wt, float* __restrict__ mx, int64_t tile) {
  extern __shared__ float4 smem4[];
  float* ts = reinterpret_cast<float*>(smem4);           // [128][d_pad + 1]
  __shared__ int64_t srow[128];
  const int ld = d_pad + 1, nq = d_pad >> 2, t = threadIdx.x;
  float mw = 0.f, mxv = 0.f;                               // max w, max |x| of the minibatch (bounds of phase 2's fp16 staging)
  if (t < 128) {
    const int64_t n = tile * 128 + t;
    const bool v = n < B;
    const int64_t r = v ? (idx ? idx[n] : n) : 0;
    srow[t] = v ? r : -1;
    const float wv = v ? (w ? w[r] : 1.f) : 0.f;
    wt[n] = wv;
    mw = fabsf(wv);
  }
  __syncthreads();
#pragma unroll 4
  for (int c = t; c < 128 * nq; c += 256) {
    const int row = c / nq, q = c - row * nq;
    const int64_t rr = srow[row];
    const float4 val = (rr >= 0) ? __ldg(reinterpret_cast<const float4*>(X + rr * d_pad) + q)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    float* dst = ts + row * ld + 4 * q;
    dst[0] = val.x; dst[1] = val.y; dst[2] = val.z; dst[3] = val.w;
    mxv = fmaxf(fmaxf(mxv, fmaxf(fabsf(val.x), fabsf(val.y))), fmaxf(fabsf(val.z), fabsf(val.w)));
  }
  __shared__ uint32_t smx[8][2];
  if (mx) {
    const uint32_t r0 = __reduce_max_sync(0xffffffffu, __float_as_uint(mw)), r1 = __reduce_max_sync(0xffffffffu, __float_as_uint(mxv));
    if ((t & 31) == 0) { smx[t >> 5][0] = r0; smx[t >> 5][1] = r1; }
  }
  __syncthreads();
  if (mx && t < 2) {      // one atomic per CTA and quantity, and only while the running maximum still grows (same-address atomics serialise in L2)
    uint32_t v = 0;
    for (int q = 0; q < 8; q++) v = max(v, smx[q][t]);
    uint32_t* m = reinterpret_cast<uint32_t*>(mx) + t;
    if (v > *reinterpret_cast<volatile uint32_t*>(m)) atomicMax(m, v);
  }
  float* out = xt + (size_t)tile * d_pad * 128 + (t & 127);
  for (int j = (t >> 7); j < d_pad; j += 2) out[(size_t)j * 128] = ts[(t & 127) * ld + j];
  pdl_wait();      // launched alongside k_pack_tc (PDL): this grid's completion must imply the packing's
}
template <int DUMMY = 0>
__global__ void __launch_bounds__(256)
k_gather128(const float* __restrict__ X, const float* __restrict__ w, const int64_t* __restrict__ idx, int64_t B,
            int d_pad, float* __restrict__ xt, float* __restrict__ wt, float* __restrict__ mx) {
  gather128_body<DUMMY>(X, w, idx, B, d_pad, xt, wt, mx, (int64_t)blockIdx.x);
}

// ---- activation (phi, phi', phi'' as a parameter of the fused kernels) --------------------------------
// The kernels keep only a = phi(z) of every hidden unit, so the derivatives are written as functions of a:
//   d1(a) = phi'(z),   r(a) = phi''(z) / phi'(z)   (phi'' u = r(a) * g with g = phi' u: what the reverse sweeps use)
// GA = false: Tanh, the expressions the kernels were written with (d1 = 1 - a^2, r = -2 a).  GA = true: the activation is a
// run-time id (CTA-uniform branch) among those invertible this way - Sigmoid, Softplus, LogSigmoid, ELU / CELU, ReLU
// (torch.nn defaults, src/network_arch.py:27-34); Tanhshrink is not (z - tanh z does not determine tanh z cheaply) and stays
// on the layer-wise family.
template <bool GA>
struct Act {
  int id;
  __device__ __forceinline__ float f1(float z) const {
    switch (id) {
      case EPDE_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
      case EPDE_ACT_SOFTPLUS: return z > 20.f ? z : (z > 0.f ? z + log1pf(expf(-z)) : log1pf(expf(z)));
      case EPDE_ACT_LOGSIGMOID: return z < 0.f ? z - log1pf(expf(z)) : -log1pf(expf(-z));
      case EPDE_ACT_ELU: return z > 0.f ? z : expm1f(z);
      default: return z > 0.f ? z : 0.f;                               // ReLU
    }
  }
  __device__ __forceinline__ float d1s(float a) const {
    switch (id) {
      case EPDE_ACT_SIGMOID: return a * (1.f - a);
      case EPDE_ACT_SOFTPLUS: return -expm1f(-a);                      // sigmoid(z) = 1 - exp(-softplus(z))
      case EPDE_ACT_LOGSIGMOID: return -expm1f(a);                     // 1 - sigmoid(z) = 1 - exp(logsigmoid(z))
      case EPDE_ACT_ELU: return a > 0.f ? 1.f : a + 1.f;
      default: return a > 0.f ? 1.f : 0.f;
    }
  }
  __device__ __forceinline__ float rs(float a) const {
    switch (id) {
      case EPDE_ACT_SIGMOID: return 1.f - 2.f * a;
      case EPDE_ACT_SOFTPLUS: return expf(-a);                         // 1 - sigmoid(z)
      case EPDE_ACT_LOGSIGMOID: return -expf(a);                       // -sigmoid(z)
      case EPDE_ACT_ELU: return a > 0.f ? 0.f : 1.f;
      default: return 0.f;
    }
  }
  __device__ __forceinline__ float2 f(float2 z) const {
    if (!GA) return tanh2(z);
    return make_float2(f1(z.x), f1(z.y));
  }
  __device__ __forceinline__ float2 d1(float2 a) const {
    if (!GA) return ffma2(make_float2(-a.x, -a.y), a, splat(1.0f));
    return make_float2(d1s(a.x), d1s(a.y));
  }
  __device__ __forceinline__ float2 r(float2 a) const {
    if (!GA) return make_float2(-2.f * a.x, -2.f * a.y);
    return make_float2(rs(a.x), rs(a.y));
  }
  // Whole-vector forms: the (CTA-uniform) switch sits OUTSIDE the unrolled loop over the H units, so that the H independent
  // evaluations of a case overlap (with the switch per element they ran one after the other: 5x slower kernels).
#define EPDE_ACT_CASES(EXPR_SIG, EXPR_SP, EXPR_LS, EXPR_ELU, EXPR_RELU)                         \
    switch (id) {                                                                               \
      case EPDE_ACT_SIGMOID:    { _Pragma("unroll") for (int o = 0; o < H; o++) { EXPR_SIG; } break; }   \
      case EPDE_ACT_SOFTPLUS:   { _Pragma("unroll") for (int o = 0; o < H; o++) { EXPR_SP; } break; }    \
      case EPDE_ACT_LOGSIGMOID: { _Pragma("unroll") for (int o = 0; o < H; o++) { EXPR_LS; } break; }    \
      case EPDE_ACT_ELU:        { _Pragma("unroll") for (int o = 0; o < H; o++) { EXPR_ELU; } break; }   \
      default:                  { _Pragma("unroll") for (int o = 0; o < H; o++) { EXPR_RELU; } break; }  \
    }
  template <int H>
  __device__ __forceinline__ void f_all(float2 (&a)[H]) const {                 // a <- phi(a)
    if (!GA) {
#pragma unroll
      for (int o = 0; o < H; o++) a[o] = tanh2(a[o]);
      return;
    }
#define EPDE_F2(ID) a[o] = make_float2(Act<true>{ID}.f1(a[o].x), Act<true>{ID}.f1(a[o].y))
    EPDE_ACT_CASES(EPDE_F2(EPDE_ACT_SIGMOID), EPDE_F2(EPDE_ACT_SOFTPLUS), EPDE_F2(EPDE_ACT_LOGSIGMOID), EPDE_F2(EPDE_ACT_ELU),
                   EPDE_F2(EPDE_ACT_RELU))
#undef EPDE_F2
  }
  template <int H>
  __device__ __forceinline__ void d1_all(const float2 (&a)[H], float2 (&dd)[H]) const {      // dd = phi'(z)
    if (!GA) {
#pragma unroll
      for (int o = 0; o < H; o++) dd[o] = ffma2(make_float2(-a[o].x, -a[o].y), a[o], splat(1.0f));
      return;
    }
#define EPDE_D2(ID) dd[o] = make_float2(Act<true>{ID}.d1s(a[o].x), Act<true>{ID}.d1s(a[o].y))
    EPDE_ACT_CASES(EPDE_D2(EPDE_ACT_SIGMOID), EPDE_D2(EPDE_ACT_SOFTPLUS), EPDE_D2(EPDE_ACT_LOGSIGMOID), EPDE_D2(EPDE_ACT_ELU),
                   EPDE_D2(EPDE_ACT_RELU))
#undef EPDE_D2
  }
  template <int H>
  __device__ __forceinline__ void r_all(const float2 (&a)[H], float2 (&rr)[H]) const {       // rr = phi''(z) / phi'(z)
    if (!GA) {
#pragma unroll
      for (int o = 0; o < H; o++) rr[o] = make_float2(-2.f * a[o].x, -2.f * a[o].y);
      return;
    }
#define EPDE_R2(ID) rr[o] = make_float2(Act<true>{ID}.rs(a[o].x), Act<true>{ID}.rs(a[o].y))
    EPDE_ACT_CASES(EPDE_R2(EPDE_ACT_SIGMOID), EPDE_R2(EPDE_ACT_SOFTPLUS), EPDE_R2(EPDE_ACT_LOGSIGMOID), EPDE_R2(EPDE_ACT_ELU),
                   EPDE_R2(EPDE_ACT_RELU))
#undef EPDE_R2
  }
#undef EPDE_ACT_CASES
};

// layer 1: a1 = tanh(W1 x + b1); x_j of the sample pair is one coalesced float2 of the transposed
// tile, software-pipelined PF columns ahead of the FFMA2 stream.
template <int H, int PF = 8, bool GA = false>
__device__ __forceinline__ void fwd_layer1(const float* __restrict__ sW1T, const float* __restrict__ sb1,
                                           const float* __restrict__ xcol, int d, int d_pad, float2 (&a1)[H],
                                           const Act<GA> act = Act<GA>{0}) {
  bias2<H>(sb1, a1);
  float2 cur[PF];
#pragma unroll
  for (int q = 0; q < PF; q++)
    cur[q] = __ldg(reinterpret_cast<const float2*>(xcol + (size_t)(q < d_pad ? q : d_pad - 1) * GT));
#pragma unroll 1
  for (int j0 = 0; j0 < d; j0 += PF) {
    float2 nxt[PF];
#pragma unroll
    for (int q = 0; q < PF; q++) {     // clamped address: columns >= d meet zero weight rows
      const int j = j0 + PF + q;
      nxt[q] = __ldg(reinterpret_cast<const float2*>(xcol + (size_t)(j < d_pad ? j : d_pad - 1) * GT));
    }
#pragma unroll
    for (int q = 0; q < PF; q++) {
      const float* wr = sW1T + (j0 + q) * H;
#pragma unroll
      for (int c = 0; c < H / 4; c++) {
        const float4 wv = *reinterpret_cast<const float4*>(wr + 4 * c);
        a1[4 * c + 0] = ffma2(splat(wv.x), cur[q], a1[4 * c + 0]);
        a1[4 * c + 1] = ffma2(splat(wv.y), cur[q], a1[4 * c + 1]);
        a1[4 * c + 2] = ffma2(splat(wv.z), cur[q], a1[4 * c + 2]);
        a1[4 * c + 3] = ffma2(splat(wv.w), cur[q], a1[4 * c + 3]);
      }
    }
#pragma unroll
    for (int q = 0; q < PF; q++) cur[q] = nxt[q];
  }
  act.f_all(a1);
}

// hidden layer: a_out = tanh(W a_in + b), weights in [in][out] layout
template <int H, bool GA = false>
__device__ __forceinline__ void fwd_hidden(const float* __restrict__ sWT, const float* __restrict__ sb,
                                           const float2 (&ain)[H], float2 (&aout)[H], const Act<GA> act = Act<GA>{0}) {
  bias2<H>(sb, aout);
  mv_acc<H, H>(sWT, ain, aout);
  act.f_all(aout);
}

// y = b4 + w4 . a3
template <int H>
__device__ __forceinline__ float2 fwd_out(const float* __restrict__ sw4, const float* __restrict__ sb4,
                                          const float2 (&a3)[H]) {
  float2 y0 = splat(sb4[0]), y1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < H / 4; q++) {
    const float4 w = *reinterpret_cast<const float4*>(sw4 + 4 * q);
    y0 = ffma2(splat(w.x), a3[4 * q + 0], y0);
    y1 = ffma2(splat(w.y), a3[4 * q + 1], y1);
    y0 = ffma2(splat(w.z), a3[4 * q + 2], y0);
    y1 = ffma2(splat(w.w), a3[4 * q + 3], y1);
  }
  return fadd2(y0, y1);
}

// g = phi'(z) * u   (Tanh: phi'(z) = 1 - a^2)
template <int H, bool GA = false>
__device__ __forceinline__ void dtanh_mul(const float2 (&a)[H], const float2 (&u)[H], float2 (&g)[H],
                                          const Act<GA> act = Act<GA>{0}) {
  float2 dd[H];
  act.d1_all(a, dd);
#pragma unroll
  for (int o = 0; o < H; o++) g[o] = fmul2(dd[o], u[o]);
}

// g3 = phi'(z3) * w4
template <int H, bool GA = false>
__device__ __forceinline__ void jac_top(const float* __restrict__ sw4, const float2 (&a3)[H], float2 (&g3)[H],
                                        const Act<GA> act = Act<GA>{0}) {
  float2 dd[H];
  act.d1_all(a3, dd);
#pragma unroll
  for (int q = 0; q < H / 4; q++) {
    const float4 w = *reinterpret_cast<const float4*>(sw4 + 4 * q);
    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int o = 4 * q + r;
      g3[o] = fmul2(dd[o], splat(ww[r]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-row metric (MD alignment layer):  e_n = u0^T M_r u0,  ubar0 = 2 ebar M_r u0  with r = idx[n]
// ---------------------------------------------------------------------------------------------
// The energy of a sample is taken in the metric M_r = J diag(c) J^T of the alignment layer's Jacobian at data-set row r
// (epde_md_align), a dense symmetric d x d matrix per ROW: the K = W1 diag(c) W1^T shortcut of the plain path does not
// exist, so u0 = W1^T g1 is formed explicitly.  MD batches are small (1e4 samples = 20 CTAs of sample-pair threads), and a
// matrix product per sample inside those few CTAs is bound by memory latency (measured: 0.32 ms per step at B = 10000,
// 0.20 of it in phase 1).  So the product is a kernel of its own with one WARP per sample - the whole GPU works on it:
//   k_pass1<MD>  writes u0 of every net to  U[net][tile][j][256]      (the layout of the gathered x tiles)
//   k_metric     p = M_r u0 in place (lane i owns p_i and walks down column i of the symmetric M: every load is one
//                coalesced row; the matrix is read once for all k nets), e = u0 . p, fp64 partial sums of w e per CTA
//   k_pass2<MD>  reads p back: t = W1 p, and the extra outer-product phase GM0: dW1 += (2 ebar g1) (x) p
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// u0_j = sum_o W1[o][j] g1[o] for j < d -> dst[j * stride] (float2: the thread's two samples)
template <int H>
__device__ __forceinline__ void u0_store(const float* __restrict__ sW1T, const float2 (&g)[H], int d, float* __restrict__ dst,
                                         int stride) {
#pragma unroll 2
  for (int j = 0; j < d; j++) {
    const float* wr = sW1T + j * H;
    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < H / 4; q++) {
      const float4 wv = *reinterpret_cast<const float4*>(wr + 4 * q);
      s0 = ffma2(splat(wv.x), g[4 * q + 0], s0); s1 = ffma2(splat(wv.y), g[4 * q + 1], s1);
      s0 = ffma2(splat(wv.z), g[4 * q + 2], s0); s1 = ffma2(splat(wv.w), g[4 * q + 3], s1);
    }
    *reinterpret_cast<float2*>(dst + (size_t)j * stride) = fadd2(s0, s1);
  }
}
constexpr int MET_THREADS = 256;            // k_metric: 8 warps = 8 samples in flight per CTA
template <int DUMMY = 0>
__global__ void __launch_bounds__(MET_THREADS)
k_metric(const float* __restrict__ metric, const int64_t* __restrict__ idx, const float* __restrict__ wt, int64_t B, int d,
         int K, float* __restrict__ U, int64_t ntiles, double* __restrict__ epart) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = lane, i1 = lane + 32;
  const bool h0 = i0 < d, h1 = i1 < d;                   // d <= 64
  const size_t dd = (size_t)d * d;
  double acc[EPDE_MAX_K];
#pragma unroll
  for (int q = 0; q < EPDE_MAX_K; q++) acc[q] = 0.0;
  const int64_t nwarps = (int64_t)gridDim.x * (MET_THREADS / 32);
#pragma unroll 1
  for (int64_t n = (int64_t)blockIdx.x * (MET_THREADS / 32) + warp; n < B; n += nwarps) {
    const float* M = metric + (size_t)(idx ? __ldg(idx + n) : n) * dd;
    const float wn = __ldg(wt + n);
    const int64_t tile = n / GT;
    const int c = (int)(n % GT);
    float u0[EPDE_MAX_K], u1[EPDE_MAX_K], p0[EPDE_MAX_K], p1[EPDE_MAX_K];
#pragma unroll
    for (int q = 0; q < EPDE_MAX_K; q++) {
      u0[q] = 0.f; u1[q] = 0.f; p0[q] = 0.f; p1[q] = 0.f;
      if (q < K) {
        const float* Uq = U + ((size_t)(q * ntiles + tile) * d) * GT + c;
        if (h0) u0[q] = Uq[(size_t)i0 * GT];
        if (h1) u1[q] = Uq[(size_t)i1 * GT];
      }
    }
    const int c0 = h0 ? i0 : d - 1, c1 = h1 ? i1 : d - 1;          // idle lanes repeat the last column
    if (d <= 32) {      // all d loads of the column in flight at once (a rolled loop is bound by the load latency)
      float m[32];
#pragma unroll
      for (int j = 0; j < 32; j++) if (j < d) m[j] = __ldg(M + (size_t)j * d + c0);
#pragma unroll
      for (int j = 0; j < 32; j++) {
        if (j < d) {
#pragma unroll
          for (int q = 0; q < EPDE_MAX_K; q++)
            if (q < K) p0[q] = fmaf(m[j], __shfl_sync(0xffffffffu, u0[q], j), p0[q]);
        }
      }
    } else {
#pragma unroll 8
      for (int j = 0; j < d; j++) {
        const float m0 = __ldg(M + (size_t)j * d + c0);
        const float m1 = __ldg(M + (size_t)j * d + c1);
#pragma unroll
        for (int q = 0; q < EPDE_MAX_K; q++) {
          if (q < K) {
            const float uj = __shfl_sync(0xffffffffu, j < 32 ? u0[q] : u1[q], j & 31);
            p0[q] = fmaf(m0, uj, p0[q]); p1[q] = fmaf(m1, uj, p1[q]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < EPDE_MAX_K; q++) {
      if (q < K) {
        const float es = warp_sum_f((h0 ? u0[q] * p0[q] : 0.f) + (h1 ? u1[q] * p1[q] : 0.f));
        acc[q] += (double)(wn * es);
        float* Uq = U + ((size_t)(q * ntiles + tile) * d) * GT + c;
        if (h0) Uq[(size_t)i0 * GT] = p0[q];
        if (h1) Uq[(size_t)i1 * GT] = p1[q];
      }
    }
  }
  __shared__ double red[MET_THREADS / 32][EPDE_MAX_K];
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < EPDE_MAX_K; q++) red[warp][q] = acc[q];
  }
  __syncthreads();
  if ((int)threadIdx.x < K) {
    double v = 0.0;
    for (int w = 0; w < MET_THREADS / 32; w++) v += red[w][threadIdx.x];
    epart[(size_t)blockIdx.x * K + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// k_pass1
// ---------------------------------------------------------------------------------------------
constexpr int P1_THREADS = 256;

// MD: per-row metric - u0 = W1^T g1 of every net goes to U[net][tile][j][256] for k_metric, which also produces the energy
// sums (this kernel then leaves the E slots of its partial sums at zero)
// GA: general activation (run-time id `act_id`, see Act) instead of Tanh
template <int H, int K, bool MD = false, bool GA = false>
__global__ void __launch_bounds__(P1_THREADS, 1)
k_pass1(const float* __restrict__ xt, const float* __restrict__ wt, int64_t B, int d, int d_pad,
        const float* __restrict__ pw, float* __restrict__ ybuf, double* __restrict__ part,
        float* __restrict__ U = nullptr, int act_id = 0) {
  constexpr int NS = 1 + 2 * K + K * (K + 1) / 2;
  const Act<GA> act{act_id};
  extern __shared__ float4 smem4[];
  float* sw = reinterpret_cast<float*>(smem4);
  const PackLayout<H> L(d_pad);
  const int pwn = L.total();
  for (int i = threadIdx.x; i < K * pwn; i += P1_THREADS) sw[i] = pw[i];
  __syncthreads();

  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) acc[s] = 0.0;

  const int64_t npairs = (B + 1) >> 1;
  const int64_t ntiles = (B + GT - 1) / GT;
  for (int64_t p = (int64_t)blockIdx.x * P1_THREADS + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * P1_THREADS) {
    const int64_t n0 = 2 * p, n1 = 2 * p + 1;
    const bool v1 = n1 < B;
    const float2 w01 = __ldg(reinterpret_cast<const float2*>(wt + n0));     // wt is padded to whole tiles
    const float w0 = w01.x, w1 = w01.y;
    const float* xcol = xt + (size_t)(n0 / GT) * d_pad * GT + (n0 % GT);
    float2 yv[K], ev[K];
#pragma unroll 1
    for (int net = 0; net < K; net++) {
      const float* s = sw + net * pwn;
      float2 a1[H], a2[H], a3[H], g[H], u[H];
      fwd_layer1<H, 8, GA>(s + L.W1c(), s + L.b1(), xcol, d, d_pad, a1, act);
      fwd_hidden<H, GA>(s + L.W2T(), s + L.b2(), a1, a2, act);
      fwd_hidden<H, GA>(s + L.W3T(), s + L.b3(), a2, a3, act);
      const float2 y = fwd_out<H>(s + L.w4(), s + L.b4(), a3);
      jac_top<H, GA>(s + L.w4(), a3, g, act);           // g3
      zero2(u); mv_acc<H, H>(s + L.W3(), g, u);         // u2 = W3^T g3
      dtanh_mul<H, GA>(a2, u, g, act);                  // g2
      zero2(u); mv_acc<H, H>(s + L.W2(), g, u);         // u1 = W2^T g2
      dtanh_mul<H, GA>(a1, u, g, act);                  // g1
      float2 e;
      if (MD) {                                         // u0 = W1^T g1 -> U; e = u0^T M_r u0 is k_metric's
        u0_store<H>(s + L.W1c(), g, d, U + ((size_t)(net * ntiles + n0 / GT) * d) * GT + (n0 % GT), GT);
        e = make_float2(0.f, 0.f);
      } else {
        zero2(u); mv_acc<H, H>(s + L.Km(), g, u);       // t = K g1
        float2 e0 = make_float2(0.f, 0.f), e1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int o = 0; o < H; o += 2) { e0 = ffma2(g[o], u[o], e0); e1 = ffma2(g[o + 1], u[o + 1], e1); }
        e = fadd2(e0, e1);                              // sum_j c_j (d_j f)^2 = g1^T K g1
      }
      ybuf[n0 * K + net] = y.x;
      if (v1) ybuf[n1 * K + net] = y.y;
#pragma unroll
      for (int q = 0; q < K; q++) if (q == net) { yv[q] = y; ev[q] = e; }
    }
    // fp64 accumulation of the batch sums
    acc[0] += (double)w0 + (double)w1;
    int s2 = 1 + 2 * K;
#pragma unroll
    for (int i = 0; i < K; i++) {
      const float wy0 = w0 * yv[i].x, wy1 = w1 * yv[i].y;
      acc[1 + i] += (double)wy0 + (double)wy1;
      acc[1 + K + i] += (double)(w0 * ev[i].x) + (double)(w1 * ev[i].y);
#pragma unroll
      for (int j = 0; j <= i; j++) { acc[s2] += (double)(wy0 * yv[j].x) + (double)(wy1 * yv[j].y); s2++; }
    }
  }
  // block reduction in a fixed order
  __shared__ double red[P1_THREADS / 32][NS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    const double v = warp_sum(acc[s]);
    if (lane == 0) red[wid][s] = v;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double v = 0.0;
    for (int q = 0; q < P1_THREADS / 32; q++) v += red[q][threadIdx.x];
    part[(size_t)blockIdx.x * NS + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// k_pass2
// ---------------------------------------------------------------------------------------------
#ifdef EPDE_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8];
#define PHASE_T0() long long pt_ = clock64()
#define PHASE_ADD(slot) do { long long n_ = clock64(); if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[slot], (unsigned long long)(n_ - pt_)); pt_ = n_; } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_ADD(slot) do {} while (0)
#endif

constexpr int P2_THREADS = 128;            // 4 warps, one per SM sub-partition
constexpr int P2_T = 2 * P2_THREADS;       // samples per tile
constexpr int P2_TS = P2_T + 4;            // row stride (floats): rows start 4 banks apart
constexpr int P2_NG = 2;                   // sample groups of the outer-product phases

// shared-memory row map (rows of P2_TS floats, [feature][sample])
template <int H>
struct RowMap {
  int xr;   // rows reserved for the transposed x tile = 10 * ceil(d / 10)
  __host__ __device__ explicit RowMap(int d) : xr(10 * ((d + 9) / 10)) {}
  // region 1
  __host__ __device__ int A1() const { return 0; }
  __host__ __device__ int A2() const { return H; }
  __host__ __device__ int A3() const { return 2 * H; }      // later Z3
  __host__ __device__ int Z3() const { return 2 * H; }
  // region 2 (phase 1 use)
  __host__ __device__ int R2() const { return 3 * H; }
  __host__ __device__ int G3() const { return R2(); }
  __host__ __device__ int G2() const { return R2() + H; }
  __host__ __device__ int G1() const { return R2() + 2 * H; }
  __host__ __device__ int xtop() const { return xr > 3 * H ? xr : 3 * H; }
  __host__ __device__ int U1() const { return R2() + xtop(); }
  __host__ __device__ int U2() const { return U1() + H; }
  __host__ __device__ int EB() const { return U2() + H; }
  // region 2 (phase 2 use): X over the G rows, Z2/Z1 over U1/U2
  __host__ __device__ int XT() const { return R2(); }
  __host__ __device__ int Z2() const { return U1(); }
  __host__ __device__ int Z1() const { return U2(); }
  // region 3: rows whose sums are gw4 / gb4
  __host__ __device__ int S4() const { return EB() + 1; }
  __host__ __device__ int YB() const { return S4() + H; }
  __host__ __device__ int ONE() const { return YB() + 1; }   // a row of ones (unit scale)
  __host__ __device__ int total() const { return ONE() + 1; }
};

template <int H>
__host__ __device__ inline size_t pass2_smem_bytes(int d, int d_pad) {
  return (size_t)RowMap<H>(d).total() * P2_TS * 4 + (size_t)PackLayout<H>(d_pad).total() * 4 +
         (size_t)P2_NG * GradLayout<H>(d_pad).total() * 4;
}

// one 4x5 register tile of an outer product over the samples [n0, n0 + cnt):
//   out[o][i] += sum_n P[o][n] * Q[i][n],  o in {ot + 5a}, i in {5*it + b}
// FFMA2 lanes pair neighbouring samples; the two partial sums are added at the end.
template <bool SCALE, int RI>
struct OuterOps {
  float4 p[4], q[RI], s;
  __device__ __forceinline__ void load(const float* __restrict__ p0, const float* __restrict__ q0,
                                       const float* __restrict__ e0, int n) {
#pragma unroll
    for (int a = 0; a < 4; a++) p[a] = *reinterpret_cast<const float4*>(p0 + (5 * a) * P2_TS + n);
#pragma unroll
    for (int b = 0; b < RI; b++) q[b] = *reinterpret_cast<const float4*>(q0 + b * P2_TS + n);
    if (SCALE) s = *reinterpret_cast<const float4*>(e0 + n);
  }
  __device__ __forceinline__ void mac(float2 (&acc)[4][RI]) {
    if (SCALE) {
#pragma unroll
      for (int a = 0; a < 4; a++) { p[a].x *= s.x; p[a].y *= s.y; p[a].z *= s.z; p[a].w *= s.w; }
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < RI; b++) {
        acc[a][b] = ffma2(make_float2(p[a].x, p[a].y), make_float2(q[b].x, q[b].y), acc[a][b]);
        acc[a][b] = ffma2(make_float2(p[a].z, p[a].w), make_float2(q[b].z, q[b].w), acc[a][b]);
      }
  }
};

// one 4 x RI register tile of an outer product over the samples [n0, n0 + cnt), cnt % 8 == 0:
//   out[o][i] += sum_n P[o][n] * Q[i][n],  o in {ot + 5a}, i in {RI*it + b}
// FFMA2 lanes pair neighbouring samples (the two partial sums are added at the end).  Two operand
// buffers alternate so that the loads of step n+4 are in flight while step n is multiplied (one warp
// per scheduler: nobody else hides the shared-memory latency).
template <bool SCALE, int RI>
__device__ __forceinline__ void outer_unit(const float* __restrict__ prow, const float* __restrict__ qrow,
                                           const float* __restrict__ eb, int n0, int cnt,
                                           float* __restrict__ out, int ld, int ot, int it, int imax) {
  float2 acc[4][RI];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < RI; b++) acc[a][b] = make_float2(0.f, 0.f);
  const float* p0 = prow + ot * P2_TS + n0;
  const float* q0 = qrow + (RI * it) * P2_TS + n0;
  const float* e0 = SCALE ? eb + n0 : nullptr;
  OuterOps<SCALE, RI> A, B;
  A.load(p0, q0, e0, 0);
#pragma unroll 1
  for (int n = 0; n < cnt; n += 8) {
    B.load(p0, q0, e0, n + 4);
    A.mac(acc);
    A.load(p0, q0, e0, (n + 8 < cnt) ? n + 8 : n);     // last iteration: harmless reload
    B.mac(acc);
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < RI; b++) {
      const int i = RI * it + b;
      if (i < imax) out[(ot + 5 * a) * ld + i] += acc[a][b].x + acc[a][b].y;
    }
}

__device__ __forceinline__ void prefetch_l2g(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// sum of one row over the tile -> out
__device__ __forceinline__ void row_sum(const float* __restrict__ row, float* __restrict__ out) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int n = 0; n < P2_T; n += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + n);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *out += (s.x + s.y) + (s.z + s.w);
}

template <int H, bool MD = false, bool GA = false>
__global__ void __launch_bounds__(P2_THREADS, 1)
k_pass2(const float* __restrict__ xt, const float* __restrict__ wt, int64_t B,
        int d, int d_pad, int K, const float* __restrict__ pw, const float* __restrict__ ybuf,
        const Coef* __restrict__ coef, float* __restrict__ pg, const float* __restrict__ U = nullptr, int act_id = 0) {
  const Act<GA> act{act_id};
  static_assert(H == 20, "outer-product tiling (4x5 tiles, row stride 5) is written for H = 20");
  extern __shared__ float4 smem4[];
  const RowMap<H> R(d);
  const PackLayout<H> L(d_pad);
  const GradLayout<H> G(d_pad);
  float* rows = reinterpret_cast<float*>(smem4);
  float* sw = rows + (size_t)R.total() * P2_TS;
  float* accS = sw + L.total();
  const int gn = G.total();
  const int net = blockIdx.y;
  const int tid = threadIdx.x;

  for (int i = tid; i < L.total(); i += P2_THREADS) sw[i] = pw[(size_t)net * L.total() + i];
  for (int i = tid; i < P2_NG * gn; i += P2_THREADS) accS[i] = 0.f;
  for (int i = tid; i < P2_TS; i += P2_THREADS) rows[R.ONE() * P2_TS + i] = 1.0f;
  const float ecoef = coef->ecoef[net];
  const float cm = coef->cm[net];
  float cw[8];
#pragma unroll
  for (int j = 0; j < 8; j++) cw[j] = (j < K) ? coef->Cw[net][j] : 0.f;
  __syncthreads();

  const int col = 2 * tid;
  const int nj4 = d_pad >> 2;
  const int64_t ntiles = (B + P2_T - 1) / P2_T;
#define ROW2(r) (*reinterpret_cast<float2*>(rows + (r) * P2_TS + col))

  // Per-tile scalars (sample weights, ybar seed): loaded for tile t+1 at the start of GM2 of tile t,
  // combined at its end -- no global-load latency at the top of a tile.
  float2 c_wn = make_float2(0.f, 0.f), c_ybar = make_float2(0.f, 0.f);
  float2 p_w = make_float2(0.f, 0.f);
  float p_y0[8], p_y1[8];
  auto stage_b = [&](int64_t tile) {
    const int64_t n0 = tile * P2_T + col, n1 = n0 + 1;
    const bool v0 = n0 < B, v1 = n1 < B;
    p_w = __ldg(reinterpret_cast<const float2*>(wt + n0));          // 0 for samples beyond B
#pragma unroll
    for (int j = 0; j < 8; j++) {
      p_y0[j] = (j < K && v0) ? ybuf[n0 * K + j] : 0.f;
      p_y1[j] = (j < K && v1) ? ybuf[n1 * K + j] : 0.f;
    }
  };
  auto stage_c = [&]() {
    float s0 = -cm, s1 = -cm;                       // ybar_n = w_n (sum_j Cw[j] y_{n,j} - cm)
#pragma unroll
    for (int j = 0; j < 8; j++) { s0 = fmaf(cw[j], p_y0[j], s0); s1 = fmaf(cw[j], p_y1[j], s1); }
    c_wn = p_w;
    c_ybar = make_float2(p_w.x * s0, p_w.y * s1);
  };
  if ((int64_t)blockIdx.x < ntiles) { stage_b(blockIdx.x); stage_c(); }

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const float2 wn = c_wn;
    const float* xtile = xt + (size_t)tile * d_pad * GT;
    const int64_t next_tile = tile + gridDim.x;
    const bool has_next = next_tile < ntiles;
    float2 ze1[H], ze2[H];

    PHASE_T0();
    // ================= MV1: forward, Jacobian chain, first reverse sweep ====================
    {
      const float2 ybar = c_ybar;
      const float2 ebar = make_float2(wn.x * ecoef, wn.y * ecoef);
      float2 a[H], g[H], u[H];
      // layer 1: x columns software-pipelined 16 ahead (one warp per scheduler: deeper than k_pass1)
      fwd_layer1<H, 16, GA>(sw + L.W1c(), sw + L.b1(), xtile + col, d, d_pad, a, act);
#pragma unroll
      for (int o = 0; o < H; o++) ROW2(R.A1() + o) = a[o];
      fwd_hidden<H, GA>(sw + L.W2T(), sw + L.b2(), a, g, act);         // g := a2
#pragma unroll
      for (int o = 0; o < H; o++) ROW2(R.A2() + o) = g[o];
      fwd_hidden<H, GA>(sw + L.W3T(), sw + L.b3(), g, a, act);         // a := a3
#pragma unroll
      for (int o = 0; o < H; o++) ROW2(R.A3() + o) = a[o];
      jac_top<H, GA>(sw + L.w4(), a, g, act);                          // g := g3
#pragma unroll
      for (int o = 0; o < H; o++) ROW2(R.G3() + o) = g[o];
      zero2(u); mv_acc<H, H>(sw + L.W3(), g, u);                       // u2 = W3^T g3
#pragma unroll
      for (int o = 0; o < H; o++) a[o] = ROW2(R.A2() + o);             // a := a2
      dtanh_mul<H, GA>(a, u, g, act);                                  // g := g2
#pragma unroll
      for (int o = 0; o < H; o++) ROW2(R.G2() + o) = g[o];
      zero2(u); mv_acc<H, H>(sw + L.W2(), g, u);                       // u1 = W2^T g2
#pragma unroll
      for (int o = 0; o < H; o++) a[o] = ROW2(R.A1() + o);             // a := a1
      dtanh_mul<H, GA>(a, u, g, act);                                  // g := g1
#pragma unroll
      for (int o = 0; o < H; o++) ROW2(R.G1() + o) = g[o];
      const float2 e2 = make_float2(2.f * ebar.x, 2.f * ebar.y);
      ROW2(R.EB()) = MD ? e2 : ebar;
      // ---- E-sweep, layer 1:  gbar1 = 2 ebar K g1
      if (MD) {
        // per-row metric: gbar1 = 2 ebar W1 p with p = M_r u0 (k_metric left it in U), and dW1 += (2 ebar g1) (x) p (there
        // is no Gamma shortcut).  p goes into the rows of U1 / U2 (free until ubar1 is written; d <= 2 H) for the extra
        // outer-product phase GM0 over the whole tile, with t = W1 p parked in the (still free) S4 rows.
        const float* Pn = U + ((size_t)(net * ntiles + tile) * d) * GT + col;
        const bool pv0 = tile * P2_T + col < B, pv1 = tile * P2_T + col + 1 < B;
        zero2(u);
#pragma unroll 2
        for (int j = 0; j < d; j++) {                                  // t = W1 p
          float2 pj = __ldg(reinterpret_cast<const float2*>(Pn + (size_t)j * GT));
          if (!pv0) pj.x = 0.f;                                        // samples beyond B: U holds nothing for them
          if (!pv1) pj.y = 0.f;
          ROW2(R.U1() + j) = pj;
          const float* wr = sw + L.W1c() + j * H;
#pragma unroll
          for (int q = 0; q < H / 4; q++) {
            const float4 wv = *reinterpret_cast<const float4*>(wr + 4 * q);
            u[4 * q + 0] = ffma2(splat(wv.x), pj, u[4 * q + 0]); u[4 * q + 1] = ffma2(splat(wv.y), pj, u[4 * q + 1]);
            u[4 * q + 2] = ffma2(splat(wv.z), pj, u[4 * q + 2]); u[4 * q + 3] = ffma2(splat(wv.w), pj, u[4 * q + 3]);
          }
        }
#pragma unroll
        for (int o = 0; o < H; o++) ROW2(R.S4() + o) = u[o];
        __syncthreads();
        {  // GM0: dW1 += sum_n (2 ebar g1)_n (x) p_n
          constexpr int SG = P2_T / P2_NG;
          const int NU = 5 * (R.xr / 5);                               // (H / 4) x (columns / 5) units per sample group
          for (int wk = tid; wk < NU * P2_NG; wk += P2_THREADS) {
            const int grp = wk / NU, un = wk % NU;
            outer_unit<true, 5>(rows + R.G1() * P2_TS, rows + R.U1() * P2_TS, rows + R.EB() * P2_TS, grp * SG, SG,
                                accS + grp * gn + G.gW1(), d_pad, un % 5, un / 5, d);
          }
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < H; o++) { u[o] = ROW2(R.S4() + o); a[o] = ROW2(R.A1() + o); g[o] = ROW2(R.G1() + o); }
      } else {
        zero2(u); mv_acc<H, H>(sw + L.Km(), g, u);                     // t = K g1
      }
      {
        float2 dv[H], rv[H];
        act.d1_all(a, dv); act.r_all(a, rv);
#pragma unroll
        for (int o = 0; o < H; o++) {
          const float2 gb = fmul2(e2, u[o]);                           // gbar1
          u[o] = fmul2(dv[o], gb);                                     // ubar1
          ROW2(R.U1() + o) = u[o];
          ze1[o] = fmul2(fmul2(rv[o], g[o]), gb);                      // phi'' u1 gbar1
        }
      }
      // ---- layer 2: gbar2 = W2 ubar1
      zero2(g); mv_acc<H, H>(sw + L.W2T(), u, g);                      // g := gbar2
      {
        float2 av[H], dv[H], rv[H];
#pragma unroll
        for (int o = 0; o < H; o++) av[o] = ROW2(R.A2() + o);
        act.d1_all(av, dv); act.r_all(av, rv);
#pragma unroll
        for (int o = 0; o < H; o++) {
          const float2 g2 = ROW2(R.G2() + o);
          u[o] = fmul2(dv[o], g[o]);                                   // ubar2
          ROW2(R.U2() + o) = u[o];
          ze2[o] = fmul2(fmul2(rv[o], g2), g[o]);
        }
      }
      // ---- layer 3: gbar3 = W3 ubar2
      zero2(g); mv_acc<H, H>(sw + L.W3T(), u, g);                      // g := gbar3
      {
        float2 av[H], dv[H], rv[H];
#pragma unroll
        for (int o = 0; o < H; o++) av[o] = ROW2(R.A3() + o);
        act.d1_all(av, dv); act.r_all(av, rv);
#pragma unroll
        for (int o = 0; o < H; o++) {
          const float2 g3 = ROW2(R.G3() + o);
          const float2 ub3 = fmul2(dv[o], g[o]);                       // ubar3 -> d/d w4 (g_4 = 1)
          ROW2(R.S4() + o) = ffma2(ybar, av[o], ub3);                  // ubar3 + ybar a3
          // zbar3 = phi'' u3 gbar3 + phi' w4 ybar = g3 (r(a3) gbar3 + ybar),  r = phi'' / phi' (Tanh: -2 a3)
          const float2 t = ffma2(rv[o], g[o], ybar);
          ROW2(R.Z3() + o) = fmul2(g3, t);
        }
      }
      ROW2(R.YB()) = ybar;
    }
    PHASE_ADD(0);
    __syncthreads();
    PHASE_ADD(1);

    // ================= GM1: Gamma, and the g (x) ubar parts of dW2, dW3 ======================
    // Every work item runs the same code (no divergence inside a warp): only pointers differ.
    {
      constexpr int UPJ = (H / 4) * (H / 5);     // 4x5 tiles of an H x H block
      constexpr int NU = 3 * UPJ;
      constexpr int SG = P2_T / P2_NG;
      for (int wk = tid; wk < NU * P2_NG; wk += P2_THREADS) {
        const int grp = wk / NU, un = wk % NU, job = un / UPJ, t = un % UPJ;
        if (MD && job == 0) continue;              // no Gamma with a per-row metric (GM0 did the g1 (x) ubar0 term)
        const int prow = (job == 0) ? R.G1() : (job == 1) ? R.G2() : R.G3();
        const int qrow = (job == 0) ? R.G1() : (job == 1) ? R.U1() : R.U2();
        const int srow = (job == 0) ? R.EB() : R.ONE();
        const int oo = (job == 0) ? G.Gam() : (job == 1) ? G.gW2() : G.gW3();
        outer_unit<true, 5>(rows + prow * P2_TS, rows + qrow * P2_TS, rows + srow * P2_TS, grp * SG, SG,
                            accS + grp * gn + oo, H, t % 5, t / 5, H);
      }
    }
    PHASE_ADD(2);
    __syncthreads();
    PHASE_ADD(3);

    // ================= MV2: second reverse sweep, x tile (transposed) ========================
    {
      // The x tile is the right operand of dW1 = zbar1 (x) x: its rows (1 KB contiguous in the
      // transposed copy) go straight from global to the shared-memory rows freed by GM1 with
      // cp.async (no registers); the copies complete behind the two mat-vecs of this phase.
      {
        const int nchunk = d * (GT / 4);                               // 16-byte chunks of the tile
        for (int c = tid; c < nchunk; c += P2_THREADS) {
          const unsigned dst = (unsigned)__cvta_generic_to_shared(rows + (R.XT() + (c >> 6)) * P2_TS + 4 * (c & 63));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const float4*>(xtile) + c));
        }
        asm volatile("cp.async.commit_group;");
      }
      float2 z[H], t[H];
#pragma unroll
      for (int o = 0; o < H; o++) z[o] = ROW2(R.Z3() + o);
      zero2(t); mv_acc<H, H>(sw + L.W3(), z, t);                       // W3^T zbar3
      {
        float2 av[H], dv[H];
#pragma unroll
        for (int o = 0; o < H; o++) av[o] = ROW2(R.A2() + o);
        act.d1_all(av, dv);
#pragma unroll
        for (int o = 0; o < H; o++) {
          z[o] = ffma2(dv[o], t[o], ze2[o]);                           // zbar2
          ROW2(R.Z2() + o) = z[o];
        }
      }
      zero2(t); mv_acc<H, H>(sw + L.W2(), z, t);                       // W2^T zbar2
      {
        float2 av[H], dv[H];
#pragma unroll
        for (int o = 0; o < H; o++) av[o] = ROW2(R.A1() + o);
        act.d1_all(av, dv);
#pragma unroll
        for (int o = 0; o < H; o++) ROW2(R.Z1() + o) = ffma2(dv[o], t[o], ze1[o]);      // zbar1
      }
      for (int j = d; j < R.xr; j++) ROW2(R.XT() + j) = make_float2(0.f, 0.f);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    PHASE_ADD(4);
    __syncthreads();
    PHASE_ADD(5);

    // ================= GM2: zbar (x) [x | a1 | a2], bias and w4 row sums =====================
    if (has_next) stage_b(next_tile);
    {
      // 4x10 register tiles (14 LDS.128 per 80 FFMA2), two sample groups of 128
      constexpr int UPJ = (H / 4) * (H / 10);
      constexpr int SG = P2_T / P2_NG;
      const int njt = R.xr / 10;
      const int nu1 = 5 * njt;                   // (H/4) * njt units for dW1
      const int NU = (nu1 + 2 * UPJ) * P2_NG;
      for (int wk = tid; wk < NU; wk += P2_THREADS) {
        const int grp = wk / (NU / P2_NG), un = wk % (NU / P2_NG);
        const int job = (un < nu1) ? 0 : (un < nu1 + UPJ) ? 1 : 2;
        const int t = (job == 0) ? un : (job == 1) ? un - nu1 : un - nu1 - UPJ;
        const int prow = (job == 0) ? R.Z1() : (job == 1) ? R.Z2() : R.Z3();
        const int qrow = (job == 0) ? R.XT() : (job == 1) ? R.A1() : R.A2();
        const int oo = (job == 0) ? G.gW1() : (job == 1) ? G.gW2() : G.gW3();
        const int ld = (job == 0) ? d_pad : H, imax = (job == 0) ? d : H;
        outer_unit<false, 10>(rows + prow * P2_TS, rows + qrow * P2_TS, nullptr, grp * SG, SG,
                              accS + grp * gn + oo, ld, t % 5, t / 5, imax);
      }
      const int nrows = 4 * H + 1;               // bias / w4 row sums on the threads without a unit
      const int first = (NU < P2_THREADS) ? NU : 0, nfree = P2_THREADS - first;
      for (int r = tid - first; r >= 0 && r < nrows; r += nfree) {
        const int q = r / H, o = r % H;
        const int srow = (q == 0) ? R.Z1() + o : (q == 1) ? R.Z2() + o : (q == 2) ? R.Z3() + o
                       : (q == 3) ? R.S4() + o : R.YB();
        const int oo = (q == 0) ? G.gb1() + o : (q == 1) ? G.gb2() + o : (q == 2) ? G.gb3() + o
                     : (q == 3) ? G.gw4() + o : G.gb4();
        row_sum(rows + srow * P2_TS, accS + oo);
      }
    }
    if (has_next) stage_c();
    PHASE_ADD(6);
    __syncthreads();
    PHASE_ADD(7);
  }
#undef ROW2
  float* dst = pg + ((size_t)blockIdx.x * gridDim.y + net) * gn;
  for (int i = tid; i < gn; i += P2_THREADS) dst[i] = accS[i] + accS[gn + i];
}

// ---------------------------------------------------------------------------------------------
// k_fin2: fixed-order fp64 reduction of the per-CTA gradient images of phase 2, plus
//         dW1 += 2 Gamma W1 diag(c)   (the g1 (x) ubar_0 term, via u0 = W1^T g1)
// ---------------------------------------------------------------------------------------------
// grid = (K, S), 1024 threads.  Work items of a net: H "row" items (row o of W1: needs row o of Gamma only - 20 elements
// instead of the whole matrix every CTA used to reduce - and produces the d elements dW1[o][:]) and chunks of
// FIN2_CHUNK of the remaining elements; CTA (net, s) takes items s, s + S, ...  (fin2_items(d) CTAs per net = one item
// each).  One WARP per output element: the lanes stride over the phase-2 CTAs (all loads of an element in flight at
// once) and a fixed shuffle tree adds the lane sums - the same order in every run.
constexpr int FIN2_CHUNK = 128;
template <int H>
__host__ __device__ inline int fin2_items(int d) { return H + (flat_net_size<H>(d) - H * d + FIN2_CHUNK - 1) / FIN2_CHUNK; }
template <int H>
__global__ void __launch_bounds__(1024)
k_fin2(const float* __restrict__ pg, int ncta, int K, int d, int d_pad,
       const float* __restrict__ params, const float* __restrict__ diag_coeff, float* __restrict__ grad,
       Peer peer, uint32_t* __restrict__ ticket) {
  static_assert(H <= 32, "one lane per Gamma column");
  __shared__ double sG[H];                       // row o of Gamma
  pdl_wait();                                    // the tensor-core family launches this kernel programmatically (PDL) behind k_pass2_tc
  const GradLayout<H> G(d_pad);
  const int gn = G.total();
  const int net = blockIdx.x, S = gridDim.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* base = pg + (size_t)net * gn;
  auto red = [&](int e) -> double {              // this lane's share of element e: CTAs lane, lane + 32, ...
    double s = 0.0;
    for (int c = lane; c < ncta; c += 32) s += (double)__ldg(base + (size_t)c * K * gn + e);
    return s;
  };
  const int fn = flat_net_size<H>(d), nrest = fn - H * d;
  const int nitems = fin2_items<H>(d);
  const float* W1 = params + (size_t)net * fn;
  float* g = grad + (size_t)net * fn;
  for (int item = blockIdx.y; item < nitems; item += S) {
    if (item < H) {
      const int o = item;
      __syncthreads();                           // sG is re-used when a CTA takes several items
      for (int b = warp; b < H; b += nw) {
        const double v = warp_sum(red(G.Gam() + o * H + b));
        if (lane == 0) sG[b] = v;
      }
      __syncthreads();
      for (int j = warp; j < d; j += nw) {
        double v = red(G.gW1() + o * d_pad + j);
        if (lane < H) v += 2.0 * (double)diag_coeff[j] * sG[lane] * (double)W1[lane * d + j];
        v = warp_sum(v);
        if (lane == 0) g[o * d + j] = (float)v;
      }
    } else {
      const int r0 = (item - H) * FIN2_CHUNK, r1 = min(nrest, r0 + FIN2_CHUNK);
      for (int rr = r0 + warp; rr < r1; rr += nw) {
        int r = rr, e;
        if (r < H) e = G.gb1() + r;
        else if ((r -= H) < H * H) e = G.gW2() + r;
        else if ((r -= H * H) < H) e = G.gb2() + r;
        else if ((r -= H) < H * H) e = G.gW3() + r;
        else if ((r -= H * H) < H) e = G.gb3() + r;
        else if ((r -= H) < H) e = G.gw4() + r;
        else e = G.gb4();
        const double v = warp_sum(red(e));
        if (lane == 0) g[H * d + rr] = (float)v;
      }
    }
  }
  if (peer.bufs) {
    // sharded step: the CTA that finishes last all-reduces the whole gradient over NVLink peer memory - no separate
    // exchange launch.  *ticket is zero at launch (set with the maxima at the start of the step).
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1u;
    __syncthreads();
    if (s_last) {
      __threadfence();
      const uint32_t epoch = *peer.epoch + 1u;
      peer_allreduce_block<float>(peer, epoch, grad, grad, (float*)nullptr, K * fn);
      __syncthreads();
      if (threadIdx.x == 0) *peer.epoch = epoch;
    }
  }
}

// forward only: y[B][K]
template <int H, bool GA = false>
__global__ void __launch_bounds__(P1_THREADS, 1)
k_forward(const float* __restrict__ xt, int64_t B, int d, int d_pad, int K,
          const float* __restrict__ pw, float* __restrict__ y, int act_id = 0) {
  const Act<GA> act{act_id};
  extern __shared__ float4 smem4[];
  float* sw = reinterpret_cast<float*>(smem4);
  const PackLayout<H> L(d_pad);
  const int pwn = L.total();
  for (int i = threadIdx.x; i < K * pwn; i += P1_THREADS) sw[i] = pw[i];
  __syncthreads();
  const int64_t npairs = (B + 1) >> 1;
  for (int64_t p = (int64_t)blockIdx.x * P1_THREADS + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * P1_THREADS) {
    const int64_t n0 = 2 * p, n1 = 2 * p + 1;
    const bool v1 = n1 < B;
    const float* xcol = xt + (size_t)(n0 / GT) * d_pad * GT + (n0 % GT);
#pragma unroll 1
    for (int net = 0; net < K; net++) {
      const float* s = sw + net * pwn;
      float2 a1[H], a2[H];
      fwd_layer1<H, 8, GA>(s + L.W1c(), s + L.b1(), xcol, d, d_pad, a1, act);
      fwd_hidden<H, GA>(s + L.W2T(), s + L.b2(), a1, a2, act);
      fwd_hidden<H, GA>(s + L.W3T(), s + L.b3(), a2, a1, act);
      const float2 yy = fwd_out<H>(s + L.w4(), s + L.b4(), a1);
      y[n0 * K + net] = yy.x;
      if (v1) y[n1 * K + net] = yy.y;
    }
  }
}

}  // namespace epde
