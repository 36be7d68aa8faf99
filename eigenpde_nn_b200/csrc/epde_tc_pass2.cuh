// k_pass2_tc: phase 2 of the fused step on the tensor cores (see epde_tc_fused.cuh for the conventions).
//
// Per 128-sample tile and net, thread = sample:
//   mat-vec chain (TS-form MMAs, A in TMEM):  x -> a1 -> a2 -> a3 | g3 -> u2 | g2 -> u1 | g1 -> K g1 |
//       ubar1 -> gbar2 | ubar2 -> gbar3 | zbar3 -> W3^T zbar3 | zbar2 -> W2^T zbar2 -> zbar1
//   weight gradients = contractions over the SAMPLES:  G[p][q] = sum_n P[n][p] Q[n][q]  (SS-form MMAs, M = 64):
//       the P / Q vectors of a "round" are written by their threads into a shared staging buffer as K-major
//       [feature][sample] images (hi / lo), several vectors stacked along M and along N:
//         round 1:  M = [g2 | g3 | ebar g1]        N = [ubar1 | ubar2 | g1]      -> dW2 (g part), dW3 (g part), Gamma
//         round 2:  M = [zbar2 | zbar3 | s4 | ybar] N = [a1 | 1 | a2 | 1]         -> dW2, db2, dW3, db3, dw4, db4
//         round 3:  M = [zbar1]                     N = [x | 1]                   -> dW1 (z part), db1
//   The two groups of a CTA share ONE staging buffer: rounds are serialised by a ticket, round T commits to
//   ring[T % 4] and the holder of ticket T waits for round T-1 before it touches the buffer.  The accumulators of a
//   round live in TMEM only for that round (the fp32 accumulator truncates, so nothing is accumulated across
//   tiles there): the next holder of the same round type first adds them (round-to-nearest) into the per-CTA
//   shared-memory gradient image.
#pragma once
#include "epde_tc_fused.cuh"

namespace epde {
namespace tc {

constexpr int P2G = 2;                          // groups per CTA (long-lived per-sample state stays in registers)
constexpr int P2THREADS = 128 * P2G;
constexpr uint32_t SG_LBO = 144;                // bytes between 4-sample chunks of a staging row group (conflict-free STS.32)
constexpr uint32_t SG_SBO = 32 * SG_LBO;        // bytes between 8-row groups (128 samples)
constexpr int SG_ROWS = 64;                     // rows per image
constexpr uint32_t SG_IMG = (SG_ROWS / 8) * SG_SBO;   // one image (hi or lo) = 36864 B
// staging buffer: [M hi][M lo][N hi][N lo]
constexpr uint32_t SG_BYTES = 4 * SG_IMG;
constexpr int P2_S1 = 65, P2_S2 = 49, P2_S3 = 65;   // row strides of the raw accumulator images (odd: conflict-free)
constexpr int P2_ACC2 = 64 * P2_S1 + 64 * P2_S2 + 20 * P2_S3;   // raw per-CTA accumulator images of the three round types

__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}
// write value v of staging row r (this thread's sample column) into the hi and lo images starting at img
// (truncating split: hi = top 19 bits exactly as the MMA would read them, lo = x - hi exact; one instruction less
//  per value than the round-to-nearest split of the mat-vec operands, same 3xTF32 accuracy to first order)
__device__ __forceinline__ void stage_val(uint32_t img_hi, int r, float v) {
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u), lo = v - hi;
  const uint32_t o = (uint32_t)(r >> 3) * SG_SBO + (uint32_t)(r & 7) * 16u;
  sts32(img_hi + o, hi); sts32(img_hi + SG_IMG + o, lo);
}

// one gradient round: D[64 x N] = sum over 128 samples, 3xTF32, small terms first; fresh accumulator
template <int N>
__device__ __forceinline__ void issue_round(uint32_t d_t, uint32_t sg_base) {
  const uint32_t idesc = idesc_tf32(64, N);
  const uint64_t mh = kdesc(sg_base, SG_LBO, SG_SBO), ml = kdesc(sg_base + SG_IMG, SG_LBO, SG_SBO);
  const uint64_t nh = kdesc(sg_base + 2 * SG_IMG, SG_LBO, SG_SBO), nl = kdesc(sg_base + 3 * SG_IMG, SG_LBO, SG_SBO);
#pragma unroll
  for (int ks = 0; ks < 16; ks++) {               // 8 samples per MMA = two 4-sample chunks = 288 B = 18 address units
    mma_ss(d_t, ml + 18 * ks, nh + 18 * ks, idesc, ks ? 1u : 0u);
    mma_ss(d_t, mh + 18 * ks, nl + 18 * ks, idesc, 1u);
  }
#pragma unroll
  for (int ks = 0; ks < 16; ks++) mma_ss(d_t, mh + 18 * ks, nh + 18 * ks, idesc, 1u);
}
__device__ __forceinline__ void issue_round_rt(int n8, uint32_t d_t, uint32_t sg_base) {
  switch (n8) {
    case 1: issue_round<8>(d_t, sg_base); break;
    case 2: issue_round<16>(d_t, sg_base); break;
    case 3: issue_round<24>(d_t, sg_base); break;
    case 4: issue_round<32>(d_t, sg_base); break;
    case 5: issue_round<40>(d_t, sg_base); break;
    case 6: issue_round<48>(d_t, sg_base); break;
    case 7: issue_round<56>(d_t, sg_base); break;
    default: issue_round<64>(d_t, sg_base); break;
  }
}

// KSC = number of layer-1 k-steps, 8 KSC >= d + 1 (1..8): a template parameter because every TMEM column offset, weight-
// image offset and MMA count of the kernel derives from it - computed per use they cost 16 % of the kernel's time.
template <int KSC, int DC = 0>
__global__ void __launch_bounds__(P2THREADS, 1)
k_pass2_tc(const float* __restrict__ xt, const float* __restrict__ wt, int64_t B, int d_rt, int d_pad_rt, int K,
           const float* __restrict__ pwt, const float* __restrict__ ybuf, const Coef* __restrict__ coef,
           float* __restrict__ pg, int fixed_order) {
  constexpr int H = TH;
  const int d = DC > 0 ? DC : d_rt;                  // DC > 0: the exact input dimension is a compile-time constant as well
  const int d_pad = DC > 0 ? (DC + 3) / 4 * 4 : d_pad_rt;
  extern __shared__ __align__(1024) unsigned char tsm[];
  const TcPack L(d, KSC);
  const GradLayout<H> G(d_pad);
  const int gn = G.total();
  unsigned char* sg = tsm;                                          // staging buffer (SG_BYTES)
  float* sw = reinterpret_cast<float*>(tsm + SG_BYTES);             // packed weights
  float* acc = reinterpret_cast<float*>(sg);                        // per-CTA gradient image (GradLayout): built at the END in the (then idle) staging buffer
  float* acc2 = sw + L.total();                                     // raw accumulator images (see flush*)
  uint64_t* bars = reinterpret_cast<uint64_t*>(acc2 + P2_ACC2);     // [P2G] mat-vec, [4] round ring
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + P2G + 4);     // [0] tmem base, [1] ticket, [2..3] ticket of group, [4..6] turn of round type
  const int tid = threadIdx.x, warp = tid >> 5, g = tid >> 7, wq = warp & 3, lane = tid & 31, net = blockIdx.y;
  const int sn = tid & 127;                                         // sample within the tile = TMEM lane
  {
    const float* src = pwt + (size_t)net * L.total();
    for (int i = tid; i < L.total(); i += P2THREADS) sw[i] = src[i];
    for (int i = tid; i < P2_ACC2; i += P2THREADS) acc2[i] = 0.f;
    for (uint32_t i = tid; i < SG_BYTES / 4; i += P2THREADS) reinterpret_cast<float*>(sg)[i] = 0.f;
  }
  if (tid == 0) { for (int i = 0; i < P2G + 4; i++) mbar_init(bars + i, 1); slot[1] = 0; slot[4] = 0; slot[5] = 0; slot[6] = 0; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tm = slot[0];
  const int K1 = L.K1(), ks1 = L.ks1;
  const int KA = K1 > 56 ? K1 : 56;                                  // A region: x (K1 columns); hidden vectors use 24 of them,
  const uint32_t gc = (uint32_t)g * (2 * KA + TN);                   // columns 24..43 of the hi and lo halves park long-lived vectors
  const uint32_t tl = tm + ((uint32_t)(wq * 32) << 16);
  const uint32_t ahi = gc, alo = gc + KA, dcol = gc + 2 * KA;
  const uint32_t pkA = tl + ahi + 24, pkB = tl + alo + 24;           // parking slots (20 columns each)
  const uint32_t gbase = (uint32_t)P2G * (2 * KA + TN);             // gradient accumulators: D1 (64), D2 (48), D3 (K1)
  const uint32_t D1 = gbase, D2 = gbase + 64, D3 = gbase + 112;
  {  // zero the accumulators once (the first flush of every round type adds them)
    float zr[16];
#pragma unroll
    for (int i = 0; i < 16; i++) zr[i] = 0.f;
    if (g == 0) { for (uint32_t c = 0; c < 112 + (uint32_t)K1; c += 16) tm_st16(tl + gbase + c, zr); wait_st(); }
  }
  const uint32_t idesc = idesc_tf32(128, TN);
  const uint32_t sbase = smem_u32(sw), sgb = smem_u32(sg);
  const float b4r = sw[L.b4()];
  const float ecoef = coef->ecoef[net], cm = coef->cm[net];
  uint64_t* bar = bars + g;
  uint64_t* ring = bars + P2G;
  uint32_t ph = 0;
  TC_DECL();
  // this thread's column inside a staging image
  const uint32_t tcol = (uint32_t)(sn >> 2) * SG_LBO + (uint32_t)(sn & 3) * 4u;
  const uint32_t sM = sgb + tcol, sN = sgb + 2 * SG_IMG + tcol;
  fence_before(); __syncthreads(); fence_after();

  // ---- flush of a round's accumulator (M = 64 layout: row r in lane (r / 16) * 32 + r % 16) into acc ----------
  // Only the holder of the staging buffer runs this, so plain read-modify-write is safe; every chunk is
  // load-all / add-all / store-all (a chain of dependent shared-memory RMWs costs ~60 cycles per element).
  auto rmw16 = [&](float* p, const float* v, bool on) {      // p[0..15] += v[0..15]
    if (on) {
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; i++) o[i] = p[i];
#pragma unroll
      for (int i = 0; i < 16; i++) p[i] = o[i] + v[i];
    }
  };
  // acc image used while the kernel runs: the three accumulators verbatim, [row][col] with 64 / 48 / 64 columns
  //   A1 = acc2 + 0, A2 = acc2 + 64 * 64, A3 = acc2 + 64 * 64 + 64 * 48
  // two 16-column chunks per tcgen05.wait::ld
  auto flush_rows = [&](uint32_t dcol0, float* base, int stride, int ncol, bool on) {
    const int r = 16 * wq + lane;
    float v[32];
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      if (c0 < ncol) {
        tm_ld16(tl + dcol0 + c0, v);
        if (c0 + 16 < ncol) tm_ld16(tl + dcol0 + c0 + 16, v + 16);
        wait_ld();
        rmw16(base + r * stride + c0, v, on);
        if (c0 + 16 < ncol) rmw16(base + r * stride + c0 + 16, v + 16, on);
      }
    }
  };
  auto flush1 = [&]() { flush_rows(D1, acc2, P2_S1, 64, lane < 16); };                       // rows g2 | g3 | ebar g1, cols ubar1 | ubar2 | g1
  auto flush2 = [&]() { flush_rows(D2, acc2 + 64 * P2_S1, P2_S2, 48, lane < 16); };          // rows zbar2 | zbar3 | s4 | ybar, cols a1 | 1 | a2 | 1
  auto flush3 = [&]() { flush_rows(D3, acc2 + 64 * P2_S1 + 64 * P2_S2, P2_S3, K1, lane < 16 && 16 * wq + lane < 20); };   // rows zbar1, cols x | 1
  // take the staging buffer: ticket T, wait for round T-1 (its MMAs have read the buffer and written their D)
  // With fixed_order (EPDE_FLAG_DETERMINISTIC) rounds of one type are taken in a FIXED order (group 0 tile i, group 1 tile i, group 0 tile i+1, ...), so the
  // order in which tile contributions are added to the accumulator images - and with it every bit of the
  // gradient - does not depend on how the two groups happen to interleave.
  auto acquire = [&](int r, uint32_t want) -> uint32_t {
    if (sn == 0) {
      uint32_t cur = want; int spin = 0;
      if (fixed_order) do {
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(cur) : "r"(smem_u32(slot + 4 + r)) : "memory");
        if (cur != want) { __nanosleep(32); if (++spin > (1 << 24)) __trap(); }
      } while (cur != want);
      slot[2 + g] = atomicAdd(slot + 1, 1u);
    }
    group_bar(g);
    const uint32_t T = slot[2 + g];
    TC_ADD(0);
    if (T > 0) mbar_wait(ring + ((T - 1) & 3), ((T - 1) >> 2) & 1);
    fence_after();
    TC_ADD(5);
    return T;
  };
#define EPDE_P2_ROUND_ISSUE(T, RT, ISSUE)                                                                  \
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                          \
    fence_before(); group_bar(g); TC_ADD(1);                                                         \
    if (wq == 0) { if (elect_one()) { fence_after(); ISSUE; commit(ring + ((T) & 3)); } __syncwarp(); }          \
    if (fixed_order && sn == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(smem_u32(slot + 4 + RT)), "r"(turn_next) : "memory"); \
    TC_ADD(8);

#define EPDE_P2_ISSUE(ISSUE)                                                                         \
    TC_ADD(0); fence_before(); group_bar(g); TC_ADD(1);                                              \
    if (wq == 0) { if (elect_one()) { fence_after(); ISSUE; commit(bar); } __syncwarp(); } TC_ADD(2);
#define EPDE_P2_WAIT() mbar_wait(bar, ph); ph ^= 1; fence_after(); TC_ADD(3);
#define EPDE_P2_MV(m) issue_matvec<3>(tm + dcol, tm + ahi, tm + alo, kdesc(sbase + L.mat(m, 0) * 4, 128, TKH * 32), \
                                      kdesc(sbase + L.mat(m, 1) * 4, 128, TKH * 32), idesc)

  const int64_t nt128 = (B + 127) / 128;
  const int64_t tstep = (int64_t)gridDim.x * P2G;
#pragma unroll 1
  for (int64_t t = (int64_t)blockIdx.x * P2G + g; t < nt128; t += tstep) {
    const int64_t n = t * 128 + sn;
    const float* xcol = xt + (size_t)t * d_pad * 128 + sn;
    // turn numbers of this tile's rounds: 2 * (local tile index) + group; group 0 skips the other group's turn when
    // that group has no tile of the same index (only possible for the last one)
    const uint32_t turn_want = 2u * (uint32_t)((t - ((int64_t)blockIdx.x * P2G + g)) / tstep) + (uint32_t)g;
    const uint32_t turn_next = turn_want + ((g == 0 && t + 1 >= nt128) ? 2u : 1u);
    const float wn = __ldg(wt + n);
    float ybar = -cm;
    if (n < B) {
#pragma unroll
      for (int j = 0; j < 8; j++) if (j < K) ybar = fmaf(coef->Cw[net][j], __ldg(ybuf + n * K + j), ybar);
    }
    ybar *= wn;                                     // ybar_n = w_n (sum_j Cw[j] y_{n,j} - cm)
    const float ebar = wn * ecoef;
    TC_T0();
    // ---- S0: x -> A -------------------------------------------------------------------------------
    {
      float xv[64];
      load_x_cols(xv, xcol, 128, d, ks1);
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (c < ks1) {
          float hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; i++) split_tf32(xv[8 * c + i], hi[i], lo[i]);
          tm_st8(tl + ahi + 8 * c, hi); tm_st8(tl + alo + 8 * c, lo);
        }
      }
      wait_st();
    }
    TC_ADD(4);
    EPDE_P2_ISSUE(issue_matvec_rt(ks1, tm + dcol, tm + ahi, tm + alo, kdesc(sbase + L.W1f(0) * 4, 128, K1 * 32),
                                  kdesc(sbase + L.W1f(1) * 4, 128, K1 * 32), idesc))
    EPDE_P2_WAIT()
    float z[H], v[TKH], a1[H], a2[H], a3[H], g1[H], g2[H], ze1[H], ze2[H];
    v[20] = 1.f; v[21] = 0.f; v[22] = 0.f; v[23] = 0.f;
    // ---- S1 .. S2: a1, a2 ---------------------------------------------------------------------------
    tm_ld20(tl + dcol, z);
#pragma unroll
    for (int i = 0; i < H; i++) { a1[i] = tanh1(z[i]); v[i] = a1[i]; }
    tm_st16(pkA, a1); tm_st4(pkA + 16, a1 + 16);        // a1 is parked in TMEM until S5
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(0))
    EPDE_P2_WAIT()
    tm_ld20(tl + dcol, z);
#pragma unroll
    for (int i = 0; i < H; i++) { a2[i] = tanh1(z[i]); v[i] = a2[i]; }
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(1))
    EPDE_P2_WAIT()
    // ---- S3: a3, g3 -> u2 ---------------------------------------------------------------------------
    tm_ld20(tl + dcol, z);
    v[20] = 0.f;                                      // no bias row from here on
#pragma unroll
    for (int i = 0; i < H; i++) { a3[i] = tanh1(z[i]); v[i] = fmaf(-a3[i], a3[i], 1.f) * sw[L.w4() + i]; }   // g3
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(3))
    EPDE_P2_WAIT()
    // ---- S4: g2 -> u1 -------------------------------------------------------------------------------
    tm_ld20(tl + dcol, z);
#pragma unroll
    for (int i = 0; i < H; i++) { g2[i] = fmaf(-a2[i], a2[i], 1.f) * z[i]; v[i] = g2[i]; }
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(2))
    EPDE_P2_WAIT()
    // ---- S5: g1 -> K g1 -----------------------------------------------------------------------------
    tm_ld16(pkA, a1); tm_ld4(pkA + 16, a1 + 16);
    tm_ld20(tl + dcol, z);
#pragma unroll
    for (int i = 0; i < H; i++) { g1[i] = fmaf(-a1[i], a1[i], 1.f) * z[i]; v[i] = g1[i]; }
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(4))
    EPDE_P2_WAIT()
    // ---- S6: gbar1 = 2 ebar K g1, ubar1 -> gbar2 = W2 ubar1 --------------------------------------------
    tm_ld20(tl + dcol, z);
    float ub1[H];
    {
      const float e2 = 2.f * ebar;
#pragma unroll
      for (int i = 0; i < H; i++) {
        const float gb = e2 * z[i];
        ub1[i] = fmaf(-a1[i], a1[i], 1.f) * gb;
        ze1[i] = (-2.f * a1[i]) * g1[i] * gb;          // phi'' u1 gbar1
        v[i] = ub1[i];
      }
    }
    tm_st16(pkB, ze1); tm_st4(pkB + 16, ze1 + 16);      // phi'' term of zbar1: parked until S10 (a1 stays parked in pkA)
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(0))
    EPDE_P2_WAIT()
    // ---- S7: ubar2 -> gbar3 = W3 ubar2;  round 1 is staged while that MMA runs ---------------------------
    tm_ld20(tl + dcol, z);
    float ub2[H];
#pragma unroll
    for (int i = 0; i < H; i++) {
      ub2[i] = fmaf(-a2[i], a2[i], 1.f) * z[i];
      ze2[i] = (-2.f * a2[i]) * g2[i] * z[i];
      v[i] = ub2[i];
    }
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(1))
    {
      const uint32_t T = acquire(0, turn_want);
#pragma unroll
      for (int i = 0; i < H; i++) {
        stage_val(sM, i, g2[i]);
        stage_val(sM, 20 + i, fmaf(-a3[i], a3[i], 1.f) * sw[L.w4() + i]);
        stage_val(sM, 40 + i, ebar * g1[i]);
        stage_val(sN, i, ub1[i]);
        stage_val(sN, 20 + i, ub2[i]);
        stage_val(sN, 40 + i, g1[i]);
      }
      TC_ADD(7);
      flush1(); TC_ADD(6);
      EPDE_P2_ROUND_ISSUE(T, 0, issue_round<64>(tm + D1, sgb))
    }
    EPDE_P2_WAIT()
    // ---- S8: zbar3, s4 -> W3^T zbar3 -------------------------------------------------------------------
    tm_ld20(tl + dcol, z);
    float zb3[H], s4[H];
#pragma unroll
    for (int i = 0; i < H; i++) {
      const float t3 = fmaf(-a3[i], a3[i], 1.f);
      s4[i] = fmaf(ybar, a3[i], t3 * z[i]);                              // ubar3 + ybar a3  (d / d w4)
      zb3[i] = (t3 * sw[L.w4() + i]) * fmaf(-2.f * a3[i], z[i], ybar);   // g3 (-2 a3 gbar3 + ybar)
      v[i] = zb3[i];
    }
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(3))
    EPDE_P2_WAIT()
    // ---- S9: zbar2 -> W2^T zbar2;  round 2 is staged while that MMA runs ----------------------------------
    tm_ld16(pkA, a1); tm_ld4(pkA + 16, a1 + 16);
    tm_ld20(tl + dcol, z);
    float zb2[H];
#pragma unroll
    for (int i = 0; i < H; i++) { zb2[i] = fmaf(fmaf(-a2[i], a2[i], 1.f), z[i], ze2[i]); v[i] = zb2[i]; }
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_ISSUE(EPDE_P2_MV(2))
    {
      const uint32_t T = acquire(1, turn_want);
#pragma unroll
      for (int i = 0; i < H; i++) {
        stage_val(sM, i, zb2[i]);
        stage_val(sM, 20 + i, zb3[i]);
        stage_val(sM, 40 + i, s4[i]);
        stage_val(sN, i, a1[i]);
        stage_val(sN, 21 + i, a2[i]);
      }
      stage_val(sM, 60, ybar);                          // rows 61..63 / columns 42..47 hold stale finite values: ignored
      stage_val(sN, 20, 1.f); stage_val(sN, 41, 1.f);
      TC_ADD(7);
      flush2(); TC_ADD(6);
      EPDE_P2_ROUND_ISSUE(T, 1, issue_round<48>(tm + D2, sgb))
    }
    EPDE_P2_WAIT()
    // ---- S10: zbar1;  round 3 ----------------------------------------------------------------------------
    tm_ld16(pkB, ze1); tm_ld4(pkB + 16, ze1 + 16);
    tm_ld20(tl + dcol, z);
    {
      float xv[64];
      load_x_cols(xv, xcol, 128, d, ks1);
      TC_ADD(9);
      const uint32_t T = acquire(2, turn_want);
#pragma unroll
      for (int i = 0; i < H; i++) stage_val(sM, i, fmaf(fmaf(-a1[i], a1[i], 1.f), z[i], ze1[i]));   // zbar1
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (c < ks1) {
#pragma unroll
          for (int i = 0; i < 8; i++) stage_val(sN, 8 * c + i, xv[8 * c + i]);
        }
      }
      TC_ADD(7);
      flush3(); TC_ADD(6);
      EPDE_P2_ROUND_ISSUE(T, 2, issue_round_rt(ks1, tm + D3, sgb))
    }
  }
#undef EPDE_P2_ROUND_ISSUE
#undef EPDE_P2_ISSUE
#undef EPDE_P2_WAIT
#undef EPDE_P2_MV
  // ---- all rounds done: last flush, per-CTA gradient image -> global ----------------------------------------
  fence_before();
  __syncthreads();
  fence_after();
  {
    const uint32_t total = slot[1];
    if (g == 0) {
      if (total > 0) mbar_wait(ring + ((total - 1) & 3), ((total - 1) >> 2) & 1);
      fence_after();
      flush1(); flush2(); flush3();
    }
  }
  fence_before();
  __syncthreads();
  {  // raw accumulator images -> GradLayout image of this CTA
    const float* A1 = acc2; const float* A2 = acc2 + 64 * P2_S1; const float* A3 = A2 + 64 * P2_S2;
    for (int e = tid; e < H * H; e += P2THREADS) {
      const int r = e / H, c = e % H;
      acc[G.gW2() + e] = A1[r * P2_S1 + c] + A2[r * P2_S2 + c];
      acc[G.gW3() + e] = A1[(20 + r) * P2_S1 + 20 + c] + A2[(20 + r) * P2_S2 + 21 + c];
      acc[G.Gam() + e] = A1[(40 + r) * P2_S1 + 40 + c];
    }
    for (int e = tid; e < H * d_pad; e += P2THREADS) {
      const int r = e / d_pad, c = e % d_pad;
      acc[G.gW1() + e] = (c < d) ? A3[r * P2_S3 + c] : 0.f;
    }
    for (int r = tid; r < H; r += P2THREADS) {
      acc[G.gb1() + r] = A3[r * P2_S3 + d];
      acc[G.gb2() + r] = A2[r * P2_S2 + 20];
      acc[G.gb3() + r] = A2[(20 + r) * P2_S2 + 41];
      acc[G.gw4() + r] = A2[(40 + r) * P2_S2 + 20];
    }
    if (tid == 0) acc[G.gb4()] = A2[60 * P2_S2 + 20];
  }
  __syncthreads();
  float* dst = pg + ((size_t)blockIdx.x * gridDim.y + net) * gn;
  for (int i = tid; i < gn; i += P2THREADS) dst[i] = acc[i];
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tm));
}

}  // namespace tc
}  // namespace epde
