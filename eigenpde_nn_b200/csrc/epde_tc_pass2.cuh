// k_pass2_tc: phase 2 of the fused step on the tensor cores (see epde_tc_fused.cuh for the conventions).
//
// Per 128-sample tile and net, thread = sample:
//   mat-vec chain (TS-form tf32 MMAs, 3xTF32 split, A in TMEM):  x -> a1 -> a2 -> a3 | g3 -> u2 | g2 -> u1 | g1 -> K g1 |
//       ubar1 -> gbar2 | ubar2 -> gbar3 | zbar3 -> W3^T zbar3 | zbar2 -> W2^T zbar2 -> zbar1
//   weight gradients = contractions over the SAMPLES:  G[p][q] = sum_n P[n][p] Q[n][q]  (SS-form MMAs, M = 64):
//       the P / Q vectors of a "round" are written by their threads into the group's staging buffer as K-major
//       [feature][sample] images, several vectors stacked along M and along N:
//         round 1:  M = [g2 | g3 | ebar g1]        N = [ubar1 | ubar2 | g1]      -> dW2 (g part), dW3 (g part), Gamma
//         round 2:  M = [zbar2 | zbar3 | s4 | ybar] N = [a1 | 1 | a2 | 1]         -> dW2, db2, dW3, db3, dw4, db4
//         round 3:  M = [zbar1]                     N = [x | 1]                   -> dW1 (z part), db1
//
// Round 2 of the build (measured on B200, tools/tc_test/f16_probe.cu, profiles/f16_probe_r2.txt):
//   * the rounds run in kind::f16 with a TWO-term fp16 split  v * s = hi + lo  (s = a power of two per vector type that
//     maps the vector's a-priori bound to 2^14, computed per step by k_coef from maxima measured in phase 1): hi * hi +
//     hi * lo + lo * hi is as accurate as 3xTF32 (both 11-bit significands; measured 4.8e-5 vs 7.1e-5 abs. on sums of
//     ~100), but one MMA covers 16 samples instead of 8 and an image is half the size: half the MMAs, half the
//     shared-memory operand traffic, and TWO staging buffers fit - one per group;
//   * an M = 64 accumulator occupies lanes 0-15 of every 32-lane quarter; a lane offset of 16 in the D address puts a
//     second accumulator into the other half of the same columns, so each group owns private accumulators.
//   Together: no ticket, no shared buffer, no cross-group synchronisation at all.  A group accumulates P2F tiles in
//   TMEM, then adds the needed blocks (round-to-nearest) into its private shared-memory image; the two images are summed
//   in a fixed order at the end, so the result is bit-identical run to run without a special mode.
#pragma once
#include <cuda_fp16.h>
#include "epde_tc_fused.cuh"

namespace epde {
namespace tc {

constexpr int P2G = 2;                          // groups per CTA (long-lived per-sample state stays in registers)
constexpr int P2THREADS = 128 * P2G + 128;      // + a third warpgroup: one of its warps does nothing but issue the rounds' MMAs, the
                                                // other three only donate their registers (setmaxnreg works per warpgroup)
constexpr int P2CTHREADS = 128 * P2G;           // compute threads
constexpr int P2F = 1;                          // tiles accumulated in TMEM between flushes; must stay 1 while some staging scales are per-tile
// staging images are MN-major, no swizzle: element (feature f, sample n) of an image at
//   (f / 8) * SG_SMN + (n / 8) * 128 + (n % 8) * 16 + (f % 8) * 2   bytes,
// a core matrix = 8 samples (K) x 8 features (contiguous 16 bytes); core matrices are contiguous along K, so a thread
// (= sample) writes 8 features with ONE 16-byte store and a warp's stores are 512 contiguous bytes (conflict-free).
// Descriptor: LBO = 128 (K direction), SBO = SG_SMN (measured: tools/tc_test/mn_probe.cu; the K-major layout of round 1
// needed one 2-byte store per value and was bound by the LSU instruction rate).
constexpr uint32_t SG_SMN = 16 * 128;           // bytes between 8-feature groups (128 samples)
constexpr int SG_ROWS = 64;                     // features per image
constexpr uint32_t SG_IMG = (SG_ROWS / 8) * SG_SMN;   // one image (hi or lo) = 16384 B
constexpr uint32_t SG_BYTES = 4 * SG_IMG;       // one group's staging buffer: [M hi][M lo][N hi][N lo]
// group-private gradient image (floats): W2 | b2, W3 | b3, Gamma with row stride 21, w4 | b4, W1 | b1 with row stride S3
constexpr int IM_S = 21;
constexpr int IM_W2 = 0, IM_W3 = 20 * IM_S, IM_G = 40 * IM_S, IM_W4 = 60 * IM_S, IM_W1 = IM_W4 + 24;
__host__ __device__ constexpr int im_s3(int K1) { return K1 + 1; }                    // odd: conflict-free row-wise RMW
__host__ __device__ constexpr int im_total(int K1) { return IM_W1 + 20 * im_s3(K1) + 4; }

__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&v)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// Stage NCH chunks of 8 features of this thread's sample: get(r) returns the value of feature row r, sc(r) its power-of-
// two scale (r is a compile-time constant after unrolling, so the selections inside cost nothing; rows 2j, 2j + 1 always
// belong to the same vector).  v * s = hi + lo with fp16 hi, lo; packed f32x2 arithmetic: 6 instructions per pair.
template <int NCH, typename F, typename S>
__device__ __forceinline__ void stage_rows(uint32_t img_hi, F get, S sc) {
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float2 x = fmul2(make_float2(get(8 * c + 2 * j), get(8 * c + 2 * j + 1)), splat(sc(8 * c + 2 * j)));
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(x.y), "f"(x.x));       // x.x -> low half
      const float2 r = ffma2(splat(-1.f), __half22float2(*reinterpret_cast<const __half2*>(&h[j])), x);   // exact residual
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l[j]) : "f"(r.y), "f"(r.x));
    }
    sts128(img_hi + (uint32_t)c * SG_SMN, h); sts128(img_hi + SG_IMG + (uint32_t)c * SG_SMN, l);
  }
}

__device__ __forceinline__ constexpr uint32_t idesc_f16(int M, int N) {     // D fp32, A/B fp16, both MN-major (bits 15, 16)
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// one gradient round: D[64 x N] (+)= sum over 128 samples of (hi + lo)(hi + lo) without lo * lo, small terms first
template <int N>
__device__ __forceinline__ void issue_round(uint32_t d_t, uint32_t sg_base, uint32_t accumulate) {
  const uint32_t idesc = idesc_f16(64, N);
  const uint64_t mh = kdesc(sg_base, 128, SG_SMN), ml = kdesc(sg_base + SG_IMG, 128, SG_SMN);
  const uint64_t nh = kdesc(sg_base + 2 * SG_IMG, 128, SG_SMN), nl = kdesc(sg_base + 3 * SG_IMG, 128, SG_SMN);
#pragma unroll
  for (int ks = 0; ks < 8; ks++) {                // 16 samples per MMA = two core matrices along K = 256 B = 16 address units
    mma_ss16(d_t, ml + 16 * ks, nh + 16 * ks, idesc, ks ? 1u : accumulate);
    mma_ss16(d_t, mh + 16 * ks, nl + 16 * ks, idesc, 1u);
  }
#pragma unroll
  for (int ks = 0; ks < 8; ks++) mma_ss16(d_t, mh + 16 * ks, nh + 16 * ks, idesc, 1u);
}
// the same with every operand a run-time value (one copy of the code serves both groups and the three round types: the
// issuing warp lives on 40 registers)
__device__ __forceinline__ void issue_round_rt(uint32_t d_t, uint32_t sg_base, uint32_t idesc) {
  const uint64_t mh = kdesc(sg_base, 128, SG_SMN), ml = kdesc(sg_base + SG_IMG, 128, SG_SMN);
  const uint64_t nh = kdesc(sg_base + 2 * SG_IMG, 128, SG_SMN), nl = kdesc(sg_base + 3 * SG_IMG, 128, SG_SMN);
#pragma unroll
  for (int ks = 0; ks < 8; ks++) {
    mma_ss16(d_t, ml + 16 * ks, nh + 16 * ks, idesc, ks ? 1u : 0u);
    mma_ss16(d_t, mh + 16 * ks, nl + 16 * ks, idesc, 1u);
  }
#pragma unroll
  for (int ks = 0; ks < 8; ks++) mma_ss16(d_t, mh + 16 * ks, nh + 16 * ks, idesc, 1u);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {       // non-blocking
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void tm_ld8(uint32_t t, float* r) {
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(t));
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = __uint_as_float(v[i]);
}

// KSC = number of layer-1 k-steps, 8 KSC >= d + 1 (1..8): a template parameter because every TMEM column offset, weight-
// image offset and MMA count of the kernel derives from it - computed per use they cost 16 % of the kernel's time.
template <int KSC, int DC = 0>
__global__ void __launch_bounds__(P2THREADS, 1)
k_pass2_tc(const float* __restrict__ xt, const float* __restrict__ wt, int64_t B, int d_rt, int d_pad_rt, int K,
           const float* __restrict__ pwt, const float* __restrict__ ybuf, const Coef* __restrict__ coef,
           const float* __restrict__ stash, float* __restrict__ pg) {
  constexpr int H = TH;
  const int d = DC > 0 ? DC : d_rt;                  // DC > 0: the exact input dimension is a compile-time constant as well
  const int d_pad = DC > 0 ? (DC + 3) / 4 * 4 : d_pad_rt;
  extern __shared__ __align__(1024) unsigned char tsm[];
  const TcPack L(d, KSC);
  const GradLayout<H> G(d_pad);
  const int gn = G.total();
  constexpr int K1 = 8 * KSC, ks1 = KSC;
  constexpr int S3 = im_s3(K1), IMT = im_total(K1);
  unsigned char* sg = tsm;                                          // staging buffers (P2G * SG_BYTES)
  float* sw = reinterpret_cast<float*>(tsm + P2G * SG_BYTES);       // packed weights
  float* acc = reinterpret_cast<float*>(sg);                        // per-CTA gradient image (GradLayout): built at the END in the (then idle) staging buffers
  float* imgs = sw + L.total();                                     // group-private gradient images
  uint64_t* bars = reinterpret_cast<uint64_t*>(imgs + P2G * IMT);   // [P2G] mat-vec done, [P2G] round done, [P2G] mat-vec requests, [P2G] round requests
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4 * P2G);     // [0] tmem base
  float* scs = reinterpret_cast<float*>(slot + 4);                  // [16] staging scales, [8] flush factors of this net
  uint32_t* dynm = reinterpret_cast<uint32_t*>(scs + 32) + ((threadIdx.x >> 7) & 1) * 32;   // this group's [8 vector types][4 warps] tile maxima (float bits)
  const int tid = threadIdx.x, warp = tid >> 5, g = tid >> 7, wq = warp & 3, lane = tid & 31, net = blockIdx.y;
  const int sn = tid & 127;                                         // sample within the tile = TMEM lane
  {
    const float* src = pwt + (size_t)net * L.total();
    for (int i = tid; i < L.total(); i += P2THREADS) sw[i] = src[i];
    for (int i = tid; i < P2G * IMT; i += P2THREADS) imgs[i] = 0.f;
    for (uint32_t i = tid; i < P2G * SG_BYTES / 4; i += P2THREADS) reinterpret_cast<float*>(sg)[i] = 0.f;
  }
  if (tid == 0) {
    for (int i = 0; i < 2 * P2G; i++) mbar_init(bars + i, 1);                   // completed by tcgen05.commit
    for (int i = 2 * P2G; i < 4 * P2G; i++) mbar_init(bars + i, 128);           // completed by the 128 threads of a group
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // Everything above is independent of the coefficient kernel, which this launch follows programmatically (PDL): the CTA
  // is resident and set up while k_reduce_coef_tc / k_coef still runs.  Phase 1's outputs (pwt, stash, ybuf, xt, wt) were
  // complete before that kernel started.
  pdl_wait();
  {
    if (tid < 16) scs[tid] = coef->sc[net][tid];
    if (tid >= 32 && tid < 40) scs[16 + tid - 32] = coef->fl[net][tid - 32];
    if (tid == 40) { scs[26] = 1.f / coef->sc[net][1]; scs[27] = 1.f / coef->sc[net][12]; }      // 1 / s(g3), 1 / s(x)
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tm = slot[0];
  constexpr int KA = K1 > 56 ? K1 : 56;                              // A region: x (K1 columns); hidden vectors use 24 of them,
  const uint32_t gc = (uint32_t)g * (2 * KA + TN);                   // columns 24..43 of the hi and lo halves park long-lived vectors
  const uint32_t tl = tm + ((uint32_t)(wq * 32) << 16);
  const uint32_t ahi = gc, alo = gc + KA, dcol = gc + 2 * KA;
  const uint32_t pkA = tl + ahi + 24, pkB = tl + alo + 24;           // parking slots (20 columns each)
  constexpr uint32_t gbase = (uint32_t)P2G * (2 * KA + TN);          // gradient accumulators: D1 (64), D2 (48), D3 (K1); group g in lanes 16 g .. 16 g + 15
  constexpr uint32_t D1 = gbase, D2 = gbase + 64, D3 = gbase + 112;
  const uint32_t dlane = (uint32_t)(16 * g) << 16;                   // lane offset of this group's accumulators (MMA D address)
  const uint32_t idesc = idesc_tf32(128, TN);
  const uint32_t sbase = smem_u32(sw), sgb = smem_u32(sg) + (uint32_t)g * SG_BYTES;
  const float b4r = sw[L.b4()];
  const float ecoef = coef->ecoef[net], cm = coef->cm[net];
  uint64_t* bar = bars + g;
  uint64_t* rbar = bars + P2G + g;
  uint32_t ph = 0, rph = 0;
  float* img = imgs + g * IMT;
  TC_DECL();
  // this thread's 16-byte slot inside a feature group of a staging image
  const uint32_t sM = sgb + (uint32_t)sn * 16u, sN = sgb + 2 * SG_IMG + (uint32_t)sn * 16u;
  // accumulator rows this thread reads back (M = 64 layout: row r in lane (r / 16) * 32 + r % 16, + 16 for group 1)
  const bool frow_on = (lane >> 4) == g;
  const int frow = 16 * wq + (lane & 15);
  const int fblk = frow / 20, fi = frow - 20 * fblk;

  // ---- dynamic staging scales ------------------------------------------------------------------------------------
  // The backward vectors (ubar2, zbar3, zbar2, s4, zbar1) cancel heavily, so a-priori bounds overestimate them by up to
  // 2^18 (MD-style weights, tests/golden/cfg3_md_30d_k2): their scale is taken from the tile's ACTUAL maximum instead.
  // Every warp publishes max |v| over its 32 samples (float bits) before a group barrier the stage needs anyway; after
  // it every thread derives the same power of two s with  2^13 <= s * max < 2^14.  fd[] keeps the factors that undo the
  // scales of the tile whose accumulators are flushed next: [0] g3 x ubar2, [1] zbar2 x a, [2] zbar3 x a, [3] s4 x 1,
  // [4] zbar1 x x.
  float fd[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  auto publish_max = [&](int slot_id, const float (&v)[H]) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < H; i += 2) m = fmaxf(m, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
    const uint32_t r = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if (lane == 0) dynm[slot_id * 4 + wq] = r;
  };
  auto dyn_scale = [&](int slot_id, float& inv) -> float {          // after the group barrier that follows publish_max
    const uint4 q = *reinterpret_cast<const uint4*>(dynm + slot_id * 4);
    uint32_t e = max(max(q.x, q.y), max(q.z, q.w)) >> 23;            // biased exponent of the group maximum
    e = e < 87u ? 87u : (e > 167u ? 167u : e);                       // scales within 2^-26 .. 2^54
    inv = __uint_as_float((e - 13u) << 23);                          // 2^(e - 127 + 1 - 14)
    return __uint_as_float((267u - e) << 23);                        // 2^(14 - (e - 127 + 1))
  };
  // ---- flush: add the needed blocks of a round's accumulator, un-scaled, into this group's image ---------------
  auto rmw = [&](float* p, const float* v, int n, float f, bool on) {      // p[0..n) += f * v[0..n)   (n <= 21, compile-time after unrolling)
    if (on) {
      float o[21];
#pragma unroll
      for (int i = 0; i < 21; i++) if (i < n) o[i] = p[i];
#pragma unroll
      for (int i = 0; i < 21; i++) if (i < n) p[i] = fmaf(v[i], f, o[i]);
    }
  };
  auto flush1 = [&]() {      // rows g2 | g3 | ebar g1 against cols ubar1 | ubar2 | g1: the three diagonal 20 x 20 blocks
    float w[40], x[20];
    const uint32_t c0 = wq < 2 ? 0u : 20u;                           // warp-uniform 40-column window
    tm_ld16(tl + D1 + c0, w); tm_ld16(tl + D1 + c0 + 16, w + 16); tm_ld8(tl + D1 + c0 + 32, w + 32); wait_ld();
    const bool hi = (uint32_t)(20 * fblk) != c0;
#pragma unroll
    for (int i = 0; i < 20; i++) x[i] = hi ? w[20 + i] : w[i];
    float* p = img + (fblk == 0 ? IM_W2 : (fblk == 1 ? IM_W3 : IM_G)) + fi * IM_S;
    rmw(p, x, 20, fblk == 1 ? fd[0] : scs[16 + (fblk < 3 ? fblk : 0)], frow_on && fblk < 3);
  };
  auto flush2 = [&]() {      // rows zbar2 | zbar3 | s4 | ybar against cols a1 | 1 | a2 | 1
    float w[48], x[21];
    tm_ld16(tl + D2, w); tm_ld16(tl + D2 + 16, w + 16); tm_ld16(tl + D2 + 32, w + 32); wait_ld();
#pragma unroll
    for (int i = 0; i < 21; i++) x[i] = fblk == 1 ? w[21 + i] : w[i];
    if (fblk < 2) {
      rmw(img + (fblk == 0 ? IM_W2 : IM_W3) + fi * IM_S, x, 21, fblk == 0 ? fd[1] : fd[2], frow_on);
    } else if (frow_on && frow <= 60) {
      img[IM_W4 + frow - 40] = fmaf(w[20], frow < 60 ? fd[3] : scs[22], img[IM_W4 + frow - 40]);     // dw4[0..19], db4
    }
  };
  auto flush3 = [&]() {      // rows zbar1 against cols x | 1
    if (wq < 2) {
      const bool on = frow_on && frow < 20;
      const float f = fd[4];
#pragma unroll
      for (int c0 = 0; c0 < K1; c0 += 16) {
        float w[16];
        if (c0 + 16 <= K1) { tm_ld16(tl + D3 + c0, w); wait_ld(); rmw(img + IM_W1 + frow * S3 + c0, w, 16, f, on); }
        else { tm_ld8(tl + D3 + c0, w); wait_ld(); rmw(img + IM_W1 + frow * S3 + c0, w, 8, f, on); }
      }
    }
  };
  // start of a round: the group's previous round has read the staging buffer and written its accumulator
  bool pend = false;                                 // a round of this group is committed to rbar and not yet waited for
  auto round_wait = [&]() {
    TC_ADD(0);
    if (pend) { mbar_wait(rbar, rph); rph ^= 1; pend = false; }
    fence_after();
    TC_ADD(5);
  };
  uint64_t* rreq = bars + 3 * P2G + (g & 1);

  // Mat-vecs (nine MMAs, latency-critical) are issued by the group's own warp 0 right after the group barrier.  A round's
  // 24 MMAs block their issuing thread for ~700 cycles (the MMA queue is 4 deep): issued by a compute warp, that warp
  // arrives late at the next group barrier and the other three wait for it.  So a compute thread only publishes "my part
  // of the round's operands is in place" on the group's request barrier (128 arrivals) and a NINTH warp, which has
  // nothing else to do, issues the round and commits it to the group's round barrier.
#define EPDE_P2_ROUND_REQ()                                                                              \
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                         \
    fence_before(); mbar_arrive(rreq); pend = true; TC_ADD(8);
#define EPDE_P2_MV_REQ(m)                                                                            \
    TC_ADD(0); fence_before(); group_bar(g); TC_ADD(1);                                              \
    if (wq == 0) { if (elect_one()) { fence_after(); EPDE_P2_MV(m); commit(bar); } __syncwarp(); } TC_ADD(2);
#define EPDE_P2_MV(m) issue_matvec<3>(tm + dcol, tm + ahi, tm + alo, kdesc(sbase + L.mat(m, 0) * 4, 128, TKH * 32), \
                                      kdesc(sbase + L.mat(m, 1) * 4, 128, TKH * 32), idesc)
#define EPDE_P2_WAIT() mbar_wait(bar, ph); ph ^= 1; fence_after(); TC_ADD(3);

  const int64_t nt128 = (B + 127) / 128;
  const int64_t tstep = (int64_t)gridDim.x * P2G;
  uint32_t lt = 0;                                   // tiles this group has started
  if (warp < 4 * P2G) {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");         // 384 x 168 registers at launch = 256 x 232 + 128 x 40
#pragma unroll 1
  for (int64_t t = (int64_t)blockIdx.x * P2G + g; t < nt128; t += tstep, lt++) {
    const int64_t n = t * 128 + sn;
    const float* xcol = xt + (size_t)t * d_pad * 128 + sn;
    const bool window_start = (lt % P2F) == 0;       // fresh accumulators; the previous window is flushed first
    const float wn = __ldg(wt + n);
    float ybar = -cm;
    if (n < B) {
#pragma unroll
      for (int j = 0; j < 8; j++) if (j < K) ybar = fmaf(coef->Cw[net][j], __ldg(ybuf + n * K + j), ybar);
    }
    ybar *= wn;                                     // ybar_n = w_n (sum_j Cw[j] y_{n,j} - cm)
    const float ebar = wn * ecoef;
    TC_T0();
    // ---- phase 1's per-sample intermediates come back from the stash (no forward / Jacobian recompute) ----------
    // t, a1, g1 (needed first) are requested before the previous tile's accumulators are flushed, the rest after it:
    // the flush hides the DRAM latency of the first half and bounds the registers in flight.
    const float* sp = stash + ((size_t)t * K + net) * ST_TILE + sn;
    float z[H], v[TKH], a1[H], a2[H], a3[H], g1[H], g2[H], ze1[H], ze2[H];
    // (requesting z, a1, g1 one tile ahead - across the loop back-edge - was measured and is SLOWER: 0.81 vs 0.77 ms;
    //  60 more live registers through the flush)
    stash_ld(sp, 4, g1); stash_ld(sp, 0, a1);
    if (t + tstep < nt128) {      // the next tile's stash rows (100 rows of 512 B) towards L2: 400 lines, up to 4 per thread
      const char* nx = reinterpret_cast<const char*>(stash + ((size_t)(t + tstep) * K + net) * ST_TILE);
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (q * 128 + sn < (int)(ST_TILE * 4 / 128)) asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + ((size_t)(q * 128 + sn) << 7)));
    }
    // ---- S5: t = K g1 (not stashed: phase 1 is bound by the stash write); the mat-vec runs while the previous tile's
    //      accumulators are flushed
    v[20] = 0.f; v[21] = 0.f; v[22] = 0.f; v[23] = 0.f;       // no bias rows in the reverse sweeps
#pragma unroll
    for (int i = 0; i < H; i++) v[i] = g1[i];
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_MV_REQ(4)                                   // K g1
    if (window_start && lt > 0) { round_wait(); flush1(); flush2(); flush3(); TC_ADD(6); }
    stash_ld(sp, 1, a2); stash_ld(sp, 3, g2); stash_ld(sp, 2, a3);
    tm_st16(pkA, a1); tm_st4(pkA + 16, a1 + 16);              // a1 is parked in TMEM until S9 / S10
    EPDE_P2_WAIT()
    tm_ld20(tl + dcol, z);                                    // t
    TC_ADD(4);
    // ---- S6: gbar1 = 2 ebar K g1, ubar1 -> gbar2 = W2 ubar1 --------------------------------------------
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
    EPDE_P2_MV_REQ(0)                                   // W2 ubar1
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
    publish_max(0, ub2);
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_MV_REQ(1)                                   // W3 ubar2
    float fn0;                                          // factor of this tile's g3 x ubar2 block (into fd[0] once the previous tile is flushed)
    {
      // round 1: the M side (static scales) is staged while the mat-vec runs; its completion also means that every thread
      // of the group has published its maximum, so the N side (ubar2: per-tile scale) follows the wait
      round_wait();
      const float s0 = scs[0], s1 = scs[1], s2 = scs[2] * ebar, s3 = scs[3], s5 = scs[5];
      stage_rows<8>(sM, [&](int r) -> float {
        return r < 20 ? g2[r < 20 ? r : 0]
             : r < 40 ? fmaf(-a3[(r >= 20 && r < 40) ? r - 20 : 0], a3[(r >= 20 && r < 40) ? r - 20 : 0], 1.f) * sw[L.w4() + ((r >= 20 && r < 40) ? r - 20 : 0)]
             : r < 60 ? g1[(r >= 40 && r < 60) ? r - 40 : 0] : 0.f; },
        [&](int r) -> float { return r < 20 ? s0 : r < 40 ? s1 : s2; });
      TC_ADD(7);
      EPDE_P2_WAIT()
      float iub2;
      const float s4s = dyn_scale(0, iub2);
      fn0 = scs[26] * iub2;
      stage_rows<8>(sN, [&](int r) -> float {
        return r < 20 ? ub1[r < 20 ? r : 0] : r < 40 ? ub2[(r >= 20 && r < 40) ? r - 20 : 0]
             : r < 60 ? g1[(r >= 40 && r < 60) ? r - 40 : 0] : 0.f; },
        [&](int r) -> float { return r < 20 ? s3 : r < 40 ? s4s : s5; });
      TC_ADD(7);
      EPDE_P2_ROUND_REQ()
    }
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
    EPDE_P2_MV_REQ(3)                                   // W3^T zbar3
    EPDE_P2_WAIT()
    // ---- S9: zbar2 -> W2^T zbar2;  round 2 is staged while that MMA runs ----------------------------------
    tm_ld16(pkA, a1); tm_ld4(pkA + 16, a1 + 16);
    tm_ld20(tl + dcol, z);
    float zb2[H];
#pragma unroll
    for (int i = 0; i < H; i++) { zb2[i] = fmaf(fmaf(-a2[i], a2[i], 1.f), z[i], ze2[i]); v[i] = zb2[i]; }
    publish_max(1, zb2); publish_max(2, zb3); publish_max(3, s4);
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_P2_MV_REQ(2)                                   // W2^T zbar2
    float fn1, fn2, fn3;
    {
      // round 2: the N side (a1 | 1 | a2 | 1, static scale) first, the M side (per-tile scales) after the mat-vec wait
      round_wait();
      const float s9 = scs[9], sa = scs[10];
      stage_rows<6>(sN, [&](int r) -> float {          // rows 42..47 zero
        return r < 20 ? a1[r < 20 ? r : 0] : r == 20 ? 1.f : r < 41 ? a2[(r >= 21 && r < 41) ? r - 21 : 0]
             : r == 41 ? 1.f : 0.f; },
        [&](int) -> float { return sa; });
      TC_ADD(7);
      EPDE_P2_WAIT()
      float i6, i7, i8;
      const float s6 = dyn_scale(1, i6), s7 = dyn_scale(2, i7), s8 = dyn_scale(3, i8);
      fn1 = i6 * (1.f / 16384.f); fn2 = i7 * (1.f / 16384.f); fn3 = i8 * (1.f / 16384.f);
      stage_rows<8>(sM, [&](int r) -> float {
        return r < 20 ? zb2[r < 20 ? r : 0] : r < 40 ? zb3[(r >= 20 && r < 40) ? r - 20 : 0]
             : r < 60 ? s4[(r >= 40 && r < 60) ? r - 40 : 0] : r == 60 ? ybar : 0.f; },
        [&](int r) -> float { return r < 20 ? s6 : r < 40 ? s7 : r < 60 ? s8 : s9; });
      TC_ADD(7);
      EPDE_P2_ROUND_REQ()
    }
    // ---- S10: zbar1;  round 3 ----------------------------------------------------------------------------
    tm_ld16(pkB, ze1); tm_ld4(pkB + 16, ze1 + 16);
    tm_ld20(tl + dcol, z);
    {
      float xv[64];
      load_x_cols(xv, xcol, 128, d, ks1);
      TC_ADD(9);
#pragma unroll
      for (int i = 0; i < H; i++) z[i] = fmaf(fmaf(-a1[i], a1[i], 1.f), z[i], ze1[i]);      // zbar1
      publish_max(4, z);
      round_wait();
      const float s12 = scs[12];
      stage_rows<KSC>(sN, [&](int r) -> float { return xv[r < 64 ? r : 0]; }, [&](int) -> float { return s12; });      // x | 1 | 0...
      TC_ADD(7);
      fence_before(); group_bar(g);                     // every warp's maximum of zbar1 is published
      TC_ADD(1);
      float i11;
      const float s11 = dyn_scale(4, i11);
      fd[4] = i11 * scs[27];
      stage_rows<3>(sM, [&](int r) -> float { return r < 20 ? z[r < 20 ? r : 0] : 0.f; },   // rows 20..23 zero (24..63: stale finite values, ignored)
                    [&](int) -> float { return s11; });
      TC_ADD(7);
      EPDE_P2_ROUND_REQ()
    }
    fd[0] = fn0; fd[1] = fn1; fd[2] = fn2; fd[3] = fn3;       // this tile's accumulators are flushed at the start of the next one
  }
#undef EPDE_P2_ROUND_REQ
#undef EPDE_P2_MV_REQ
#undef EPDE_P2_MV
#undef EPDE_P2_WAIT
  // ---- all rounds requested: last flush of this group's accumulators -------------------------------------------
  if (lt > 0) { round_wait(); flush1(); flush2(); flush3(); }
  } else {
    // ---- the issuing warp: serves the round requests of both groups (fixed order per group: rounds 1, 2, 3 per tile) ----
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");        // this warpgroup's registers go to the compute warps
    if (warp == 4 * P2G && elect_one()) {
      uint32_t rn[P2G], ri[P2G];
#pragma unroll
      for (int q = 0; q < P2G; q++) {
        const int64_t first = (int64_t)blockIdx.x * P2G + q;
        rn[q] = first < nt128 ? 3u * (uint32_t)((nt128 - 1 - first) / tstep + 1) : 0u;
        ri[q] = 0u;
      }
      uint32_t idle = 0;
      int q = 0;
      while (ri[0] < rn[0] || ri[1] < rn[1]) {
        const uint32_t done = q ? ri[1] : ri[0], want = q ? rn[1] : rn[0];
        if (done < want && mbar_test(bars + 3 * P2G + q, done & 1u)) {
          fence_after();
          const uint32_t ty = done % 3u;
          const uint32_t dcolq = ty == 0 ? D1 : (ty == 1 ? D2 : D3);
          const uint32_t idq = ty == 0 ? idesc_f16(64, 64) : (ty == 1 ? idesc_f16(64, 48) : idesc_f16(64, K1));
          issue_round_rt(tm + dcolq + ((uint32_t)(16 * q) << 16), smem_u32(sg) + (uint32_t)q * SG_BYTES, idq);
          commit(bars + P2G + q);
          if (q) ri[1]++; else ri[0]++;
          idle = 0;
        } else if (++idle > (1u << 26)) __trap();           // a protocol error traps instead of hanging the GPU
        q ^= 1;
      }
    }
    return;                                               // the epilogue below belongs to the 256 compute threads
  }
  auto cbar = [&]() { asm volatile("bar.sync 3, 256;" ::: "memory"); };      // barrier of the compute threads
  fence_before();
  cbar();
  fence_after();
  {  // the two group images -> GradLayout image of this CTA (fixed order: group 0 + group 1)
    const float* I0 = imgs; const float* I1 = imgs + IMT;
    for (int e = tid; e < H * H; e += P2CTHREADS) {
      const int r = e / H, c = e % H;
      acc[G.gW2() + e] = I0[IM_W2 + r * IM_S + c] + I1[IM_W2 + r * IM_S + c];
      acc[G.gW3() + e] = I0[IM_W3 + r * IM_S + c] + I1[IM_W3 + r * IM_S + c];
      acc[G.Gam() + e] = I0[IM_G + r * IM_S + c] + I1[IM_G + r * IM_S + c];
    }
    for (int e = tid; e < H * d_pad; e += P2CTHREADS) {
      const int r = e / d_pad, c = e % d_pad;
      acc[G.gW1() + e] = (c < d) ? I0[IM_W1 + r * S3 + c] + I1[IM_W1 + r * S3 + c] : 0.f;
    }
    for (int r = tid; r < H; r += P2CTHREADS) {
      acc[G.gb1() + r] = I0[IM_W1 + r * S3 + d] + I1[IM_W1 + r * S3 + d];
      acc[G.gb2() + r] = I0[IM_W2 + r * IM_S + 20] + I1[IM_W2 + r * IM_S + 20];
      acc[G.gb3() + r] = I0[IM_W3 + r * IM_S + 20] + I1[IM_W3 + r * IM_S + 20];
      acc[G.gw4() + r] = I0[IM_W4 + r] + I1[IM_W4 + r];
    }
    if (tid == 0) acc[G.gb4()] = I0[IM_W4 + 20] + I1[IM_W4 + 20];
  }
  cbar();
  float* dst = pg + ((size_t)blockIdx.x * gridDim.y + net) * gn;
  for (int i = tid; i < gn; i += P2CTHREADS) dst[i] = acc[i];
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tm));
}

}  // namespace tc
}  // namespace epde
