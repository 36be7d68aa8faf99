// Tensor-core version of the fused step for the reference architecture (d -> 20 -> 20 -> 20 -> 1, Tanh),
// sm_100a only: tcgen05.mma kind::tf32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi, fp32-level accuracy).
//
//   thread = TMEM lane = ONE sample; a "group" = 4 warps = 128 samples = the M dimension of every MMA.
//   per-sample mat-vecs  D[128 x 32] = A[128 x K] * B[K x 32]:
//       A (activations, hi/lo) is written to TENSOR MEMORY by the owning threads (tcgen05.st) and read by the
//       MMA from there (no shared-memory traffic), B (weights, hi/lo, biases folded in as an extra K row
//       against a ones column of A) is a static K-major shared-memory image built once per step by k_pack_tc.
//   D comes back with tcgen05.ld for the element-wise work (tanh, Jacobian chain, seeds).
// Two groups per CTA work on different tiles so that one group's element-wise work overlaps the other's MMAs.
//
// Measured building blocks (tools/tc_test/umma_probe.cu, B200): TS-form M128 N32 K8 = 17.9 cycles/MMA (floor 16),
// SS-form M128 N32 = 47 cycles (operand A read from shared memory: 4 KB / MMA), fp32 accumulator truncates (RZ),
// MN-major TF32 operands without swizzle read as zero (CUTLASS: only SW128_32B exists for MN-major tf32).
#pragma once
#include "epde_common.cuh"
#include "epde_coef.cuh"
#include "epde_fused.cuh"

namespace epde {
namespace tc {

#ifdef EPDE_PHASE_TIMING
__device__ unsigned long long g_tc_cycles[16];
#define TC_DECL() long long tt_ = 0
#define TC_T0() tt_ = clock64()
#define TC_ADD(slot) do { long long n_ = clock64(); if (threadIdx.x == 0) atomicAdd(&g_tc_cycles[slot], (unsigned long long)(n_ - tt_)); tt_ = n_; } while (0)
#else
#define TC_DECL() do {} while (0)
#define TC_T0() do {} while (0)
#define TC_ADD(slot) do {} while (0)
#endif

constexpr int TH = 20;              // hidden width handled by this family
constexpr int TKH = 24;             // hidden contraction length padded to k-steps of 8 (+ ones column at index 20)
constexpr int TN = 32;              // MMA N (M = 128 needs N % 16 == 0)
constexpr int TG = 3;               // groups per CTA (TMEM: 3 x (2 x 64 + 32) = 480 of 512 columns)
constexpr int TXR = 64;             // rows of the per-group x staging buffer [j][128]
constexpr int TTHREADS = 128 * TG;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (version 1): LBO = stride between the two 16-byte
// k-chunks of one MMA, SBO = stride between 8-row groups
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ constexpr uint32_t idesc_tf32(int M, int N) {   // D fp32, A/B tf32, both K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
// bounded wait: a protocol error traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
#pragma unroll 1
  for (int it = 0; it < (1 << 19); it++) {          // x 2 us suspend hint = ~1 s
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(a), "r"(parity), "r"(2000u) : "memory");    // suspend-time hint (ns)
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" :: "r"(g + 1) : "memory"); }

#define EPDE_U(x) __float_as_uint(x)
__device__ __forceinline__ void tm_st4(uint32_t t, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               :: "r"(t), "r"(EPDE_U(r[0])), "r"(EPDE_U(r[1])), "r"(EPDE_U(r[2])), "r"(EPDE_U(r[3])) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t t, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(t), "r"(EPDE_U(r[0])), "r"(EPDE_U(r[1])), "r"(EPDE_U(r[2])), "r"(EPDE_U(r[3])),
                  "r"(EPDE_U(r[4])), "r"(EPDE_U(r[5])), "r"(EPDE_U(r[6])), "r"(EPDE_U(r[7])) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t t, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(t), "r"(EPDE_U(r[0])), "r"(EPDE_U(r[1])), "r"(EPDE_U(r[2])), "r"(EPDE_U(r[3])),
                  "r"(EPDE_U(r[4])), "r"(EPDE_U(r[5])), "r"(EPDE_U(r[6])), "r"(EPDE_U(r[7])),
                  "r"(EPDE_U(r[8])), "r"(EPDE_U(r[9])), "r"(EPDE_U(r[10])), "r"(EPDE_U(r[11])),
                  "r"(EPDE_U(r[12])), "r"(EPDE_U(r[13])), "r"(EPDE_U(r[14])), "r"(EPDE_U(r[15])) : "memory");
}
#undef EPDE_U
// 24 columns (a padded hidden vector)
__device__ __forceinline__ void tm_st24(uint32_t t, const float (&r)[TKH]) { tm_st16(t, r); tm_st8(t + 16, r + 16); }

__device__ __forceinline__ void tm_ld4(uint32_t t, float* r) {
  uint32_t v[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(t));
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tm_ld16(uint32_t t, float* r) {
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(t));
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = __uint_as_float(v[i]);
}
// the 20 valid columns of a mat-vec result; the values may only be used after wait_ld()
__device__ __forceinline__ void tm_ld20(uint32_t t, float (&r)[TH]) { tm_ld16(t, r); tm_ld4(t + 16, r + 16); wait_ld(); }

// round-to-nearest TF32 split: hi has 10 explicit mantissa bits, lo = x - hi is exact in fp32
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}

// The x columns of one sample, xv[j] = x_j (j < d), 1 (j == d: the ones column of the folded bias), 0 beyond, for
// j < 8 * ks1.  Whole 8-column chunks below d are loaded without per-column tests (the chunk tests are uniform over the
// CTA); only the chunk that contains d pays for predicates.  `col` points at this thread's x_0, stride in floats.
__device__ __forceinline__ void load_x_cols(float (&xv)[64], const float* __restrict__ col, int stride, int d, int ks1) {
#pragma unroll
  for (int c = 0; c < 8; c++) {
    if (c < ks1) {
      if (8 * c + 8 <= d) {
#pragma unroll
        for (int i = 0; i < 8; i++) xv[8 * c + i] = col[(size_t)(8 * c + i) * stride];
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int j = 8 * c + i;
          xv[j] = (j < d) ? col[(size_t)j * stride] : (j == d ? 1.f : 0.f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) xv[8 * c + i] = 0.f;
    }
  }
}

// ---- packed (f32x2) element-wise helpers over feature pairs: half the issue slots of the scalar forms ----------
// tanh of two values + t = 1 - a^2 (FMUL2, 2 x EX2, FADD2, 2 x RCP, FFMA2, FFMA2)
__device__ __forceinline__ void tanh_pair(float z0, float z1, float& a0, float& a1, float& t0, float& t1) {
#ifdef EPDE_FAST_TANH
  const float2 s = fmul2(make_float2(z0, z1), splat(2.885390081777927f));
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s.y));
  const float2 ep = fadd2(make_float2(e0, e1), splat(1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(ep.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(ep.y));
  const float2 a = ffma2(splat(-2.0f), make_float2(r0, r1), splat(1.0f));
  const float2 t = ffma2(make_float2(-a.x, -a.y), a, splat(1.0f));
  a0 = a.x; a1 = a.y; t0 = t.x; t1 = t.y;
#else
  a0 = tanhf(z0); a1 = tanhf(z1); t0 = fmaf(-a0, a0, 1.f); t1 = fmaf(-a1, a1, 1.f);
#endif
}
// round-to-nearest TF32 split of two values (the integer part cannot be packed, the subtraction can)
__device__ __forceinline__ void split_pair(float x0, float x1, float& h0, float& h1, float& l0, float& l1) {
  h0 = __uint_as_float((__float_as_uint(x0) + 0x1000u) & 0xFFFFE000u);
  h1 = __uint_as_float((__float_as_uint(x1) + 0x1000u) & 0xFFFFE000u);
  const float2 l = fadd2(make_float2(x0, x1), make_float2(-h0, -h1));
  l0 = l.x; l1 = l.y;
}

// ---- packed weight image of one net (floats) -------------------------------------------------
// Every matrix is the K-major, no-swizzle image of B[n][k] (n < 32 output columns of the MMA, k contraction):
//   element (n, k) at (n / 8) * (K / 4) * 32 + (k / 4) * 32 + (n % 8) * 4 + (k % 4)      (floats)
// i.e. LBO = 128 B, SBO = K * 32 B, and k-step s of an MMA chain starts 256 * s bytes into the image.
__host__ __device__ inline int kmaj_off(int n, int k, int K) {
  return (n >> 3) * (K >> 2) * 32 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
}
struct TcPack {
  int ks1;                                   // k-steps of layer 1: 8 * ks1 >= d + 1 (ones column at index d)
  __host__ __device__ explicit TcPack(int d) : ks1((d + 1 + 7) / 8) {}
  __host__ __device__ TcPack(int, int ks) : ks1(ks) {}               // from the k-step count (compile-time in the kernels)
  __host__ __device__ int K1() const { return 8 * ks1; }
  __host__ __device__ int W1f(int lo) const { return lo * TN * K1(); }                    // B[k=j][n=o] = W1[o][j], row d = b1
  __host__ __device__ int mat(int m, int lo) const { return 2 * TN * K1() + (2 * m + lo) * TN * TKH; }
  // m: 0 W2f (B[k=i][n=o] = W2[o][i], row 20 = b2)  1 W3f  2 W2b (B[k=o][n=i] = W2[o][i])  3 W3b  4 Km
  __host__ __device__ int w4() const { return mat(5, 0); }                                // 20 floats (+4 pad)
  __host__ __device__ int b4() const { return w4() + TKH; }                               // 4 floats, [0] used
  __host__ __device__ int total() const { return b4() + 8; }                              // multiple of 8 floats
};

// one CTA per net: flat parameters -> TcPack image in global memory (hi / lo split done here, once per step)
__device__ __forceinline__ void pack_tc_body(int net, const float* __restrict__ params, const float* __restrict__ diag_coeff, int d,
                                             float* __restrict__ pwt, float* __restrict__ nrm) {
  constexpr int H = TH;
  const TcPack L(d);
  const int K1 = L.K1();
  const float* p = params + (size_t)net * flat_net_size<H>(d);
  float* o = pwt + (size_t)net * L.total();
  const float* W1 = p;
  const float* b1 = W1 + H * d;
  const float* W2 = b1 + H;
  const float* b2 = W2 + H * H;
  const float* W3 = b2 + H;
  const float* b3 = W3 + H * H;
  const float* W4 = b3 + H;
  const float* b4 = W4 + H;
  // W1 and diag(c) through shared memory: K = W1 diag(c) W1^T reads every element of W1 twenty times (d <= 63 in this family)
  __shared__ float sW1[TH * 64];
  __shared__ float sdg[64];
  for (int e = threadIdx.x; e < H * d; e += blockDim.x) sW1[e] = W1[e];
  for (int e = threadIdx.x; e < d; e += blockDim.x) sdg[e] = diag_coeff[e];
  __syncthreads();
  // gridDim.y CTAs share the elements of a net's image (one CTA per net took 12 us, all of it latency)
  const int t0 = threadIdx.x + blockIdx.y * blockDim.x, ts = blockDim.x * gridDim.y;
  for (int e = t0; e < TN * K1; e += ts) {
    const int n = e / K1, k = e % K1;
    float v = 0.f;
    if (n < H) v = (k < d) ? sW1[n * d + k] : (k == d ? b1[n] : 0.f);
    float hi, lo; split_tf32(v, hi, lo);
    o[L.W1f(0) + kmaj_off(n, k, K1)] = hi; o[L.W1f(1) + kmaj_off(n, k, K1)] = lo;
  }
  for (int e = t0; e < 5 * TN * TKH; e += ts) {
    const int m = e / (TN * TKH), r = e % (TN * TKH), n = r / TKH, k = r % TKH;
    float v = 0.f;
    if (n < H) {
      if (m == 0) v = (k < H) ? W2[n * H + k] : (k == H ? b2[n] : 0.f);
      else if (m == 1) v = (k < H) ? W3[n * H + k] : (k == H ? b3[n] : 0.f);
      else if (m == 2) v = (k < H) ? W2[k * H + n] : 0.f;
      else if (m == 3) v = (k < H) ? W3[k * H + n] : 0.f;
      else if (k < H) {
        double s = 0.0;
        for (int j = 0; j < d; j++) s += (double)sdg[j] * (double)sW1[n * d + j] * (double)sW1[k * d + j];
        v = (float)s;
      }
    }
    float hi, lo; split_tf32(v, hi, lo);
    o[L.mat(m, 0) + kmaj_off(n, k, TKH)] = hi; o[L.mat(m, 1) + kmaj_off(n, k, TKH)] = lo;
  }
  if (blockIdx.y != 0) return;
  for (int e = threadIdx.x; e < TKH + 8; e += blockDim.x)
    o[L.w4() + e] = (e < H) ? W4[e] : (e == TKH ? b4[0] : 0.f);
  // induced infinity norms for the bounds of phase 2's fp16 staging (k_coef): max row / column abs sums, max |w4|
  __shared__ float sn[5][H];
  if (threadIdx.x < H) {
    const int r = threadIdx.x;
    float a = 0.f, b = 0.f, c = 0.f, e = 0.f;
    for (int i = 0; i < H; i++) { a += fabsf(W2[r * H + i]); b += fabsf(W2[i * H + r]); c += fabsf(W3[r * H + i]); e += fabsf(W3[i * H + r]); }
    sn[0][r] = a; sn[1][r] = b; sn[2][r] = c; sn[3][r] = e; sn[4][r] = fabsf(W4[r]);
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float m = 0.f;
    for (int i = 0; i < H; i++) m = fmaxf(m, sn[threadIdx.x][i]);
    nrm[kNrmPerNet * net + threadIdx.x] = m;
  }
}
__global__ void k_pack_tc(const float* __restrict__ params, const float* __restrict__ diag_coeff, int d,
                          float* __restrict__ pwt, float* __restrict__ nrm) {
  pdl_trigger();                     // the gather that follows is independent of the packing: it runs alongside (PDL)
  pack_tc_body(blockIdx.x, params, diag_coeff, d, pwt, nrm);
}
// issue one 3xTF32 mat-vec chain: D[128 x 32] = A[128 x 8 KS] * B, small terms first (the accumulator truncates).
// Straight-line code with compile-time offsets: every operand of UTCHMMA lives in a uniform register, and
// anything computed per MMA on the uniform datapath costs far more than the MMA itself (18 cycles).
template <int KS>
__device__ __forceinline__ void issue_matvec(uint32_t d_t, uint32_t ahi_t, uint32_t alo_t, uint64_t bhi, uint64_t blo,
                                             uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    mma_ts(d_t, alo_t + 8 * ks, bhi + 16 * ks, idesc, ks ? 1u : 0u);      // + 256 bytes = + 16 in the address field
    mma_ts(d_t, ahi_t + 8 * ks, blo + 16 * ks, idesc, 1u);
  }
#pragma unroll
  for (int ks = 0; ks < KS; ks++) mma_ts(d_t, ahi_t + 8 * ks, bhi + 16 * ks, idesc, 1u);
}
__device__ __forceinline__ void issue_matvec_rt(int ksteps, uint32_t d_t, uint32_t ahi_t, uint32_t alo_t, uint64_t bhi,
                                                uint64_t blo, uint32_t idesc) {
  switch (ksteps) {
    case 1: issue_matvec<1>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
    case 2: issue_matvec<2>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
    case 3: issue_matvec<3>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
    case 4: issue_matvec<4>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
    case 5: issue_matvec<5>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
    case 6: issue_matvec<6>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
    case 7: issue_matvec<7>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
    default: issue_matvec<8>(d_t, ahi_t, alo_t, bhi, blo, idesc); break;
  }
}

// split a padded hidden vector and store it as the next A operand (24 hi + 24 lo columns)
__device__ __forceinline__ void store_vec24(uint32_t ahi_t, uint32_t alo_t, const float (&v)[TKH]) {
  float hi[TKH], lo[TKH];
#pragma unroll
  for (int i = 0; i < TKH; i += 2) split_pair(v[i], v[i + 1], hi[i], hi[i + 1], lo[i], lo[i + 1]);
  tm_st24(ahi_t, hi); tm_st24(alo_t, lo);
  wait_st();
}

// ---------------------------------------------------------------------------------------------
// k_pass1_tc: forward + Jacobian chain + energy for ONE net per CTA (blockIdx.y), 128-sample tiles
//   part[(net * gridDim.x + blockIdx.x) * 2 + {0,1}] = sum w y, sum w e  (fp64);   ybuf[n][k]
// ---------------------------------------------------------------------------------------------
// Per-sample intermediates phase 1 hands to phase 2 (SURVEY H2 option C: stash instead of recompute):
//   stash[tile128][net][vector][feature][128 samples], vectors 0 a1, 1 a2, 2 a3, 3 g2, 4 g1   (fp32)
// 400 B per sample and net, written and read exactly once with coalesced 128-byte rows (streaming: far larger than L2).
// (t = K g1 was a sixth vector: phase 1 is bound by the stash write, and phase 2 recomputes t with one mat-vec that runs
//  behind the flush of the previous tile's accumulators.)
constexpr int ST_VECS = 5;
constexpr size_t ST_TILE = (size_t)ST_VECS * TH * 128;          // floats per tile and net
__device__ __forceinline__ void stash_st(float* p, int vec, const float* v) {
#pragma unroll
  for (int i = 0; i < TH; i++) __stcs(p + (size_t)(vec * TH + i) * 128, v[i]);
}
__device__ __forceinline__ void stash_ld(const float* p, int vec, float (&v)[TH]) {
#pragma unroll
  for (int i = 0; i < TH; i++) v[i] = __ldcs(p + (size_t)(vec * TH + i) * 128);
}

// KSC = number of layer-1 k-steps (see k_pass2_tc)
template <int KSC, int DC = 0>
__global__ void __launch_bounds__(TTHREADS, 1)
k_pass1_tc(const float* __restrict__ xt, const float* __restrict__ wt, int64_t B, int d_rt, int d_pad_rt, int K,
           const float* __restrict__ pwt, float* __restrict__ ybuf, double* __restrict__ part, float* __restrict__ mx,
           float* __restrict__ stash) {
  constexpr int H = TH;
  const int d = DC > 0 ? DC : d_rt;                  // DC > 0: the exact input dimension is a compile-time constant as well
  const int d_pad = DC > 0 ? (DC + 3) / 4 * 4 : d_pad_rt;
  extern __shared__ __align__(1024) unsigned char tsm[];
  const TcPack L(d, KSC);
  float* sw = reinterpret_cast<float*>(tsm);
  float* xs0 = sw + L.total() + (size_t)(threadIdx.x >> 7) * TXR * 128;    // this group's x block [j][128]
  const float* xs = xs0 + (threadIdx.x & 127);                             // this thread's column
  uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + ((size_t)L.total() + (size_t)TG * TXR * 128) * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 * TG);
  __shared__ double red[TTHREADS / 32][2];
  const int tid = threadIdx.x, warp = tid >> 5, g = tid >> 7, wq = warp & 3, net = blockIdx.y;
  {
    const float* src = pwt + (size_t)net * L.total();
    for (int i = tid; i < L.total(); i += TTHREADS) sw[i] = src[i];
  }
  if (tid == 0) { for (int i = 0; i < 2 * TG; i++) mbar_init(bars + i, 1); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tm = *slot;
  const int K1 = L.K1(), ks1 = L.ks1;
  const int KA = K1 > TKH ? K1 : TKH;                                // A region: x (K1 columns) or a hidden vector (24)
  const uint32_t gc = (uint32_t)g * (2 * KA + TN);
  const uint32_t tl = tm + ((uint32_t)(wq * 32) << 16);             // this warp's lane quarter
  const uint32_t ahi = gc, alo = gc + KA, dcol = gc + 2 * KA;
  const uint32_t idesc = idesc_tf32(128, TN);
  const uint32_t sbase = smem_u32(sw);
  const uint64_t dW1h = kdesc(sbase + L.W1f(0) * 4, 128, K1 * 32), dW1l = kdesc(sbase + L.W1f(1) * 4, 128, K1 * 32);
  uint64_t dMh[5], dMl[5];
#pragma unroll
  for (int m = 0; m < 5; m++) {
    dMh[m] = kdesc(sbase + L.mat(m, 0) * 4, 128, TKH * 32); dMl[m] = kdesc(sbase + L.mat(m, 1) * 4, 128, TKH * 32);
  }
  float w4r[H];
#pragma unroll
  for (int i = 0; i < H; i++) w4r[i] = sw[L.w4() + i];
  const float b4r = sw[L.b4()];
  uint64_t* bar = bars + g;
  uint32_t ph = 0;
  double accS = 0.0, accE = 0.0;
  float mg2 = 0.f, mg1 = 0.f, mt = 0.f, my = 0.f;      // running maxima of |g2|, |g1|, |K g1|, |y| (bounds of phase 2's fp16 staging)
  TC_DECL();

  const int64_t nt128 = (B + 127) / 128;
  // the x block of a tile ([j][128], contiguous in the 128-tile-major gather copy) travels global -> shared as ONE
  // bulk copy issued by the group's elected thread; the copy of tile t+1 is issued right after the first group
  // barrier of tile t (every thread has read its column by then)
  const uint32_t xbytes = (uint32_t)d_pad * 512u;
  uint64_t* xbar = bars + TG + g;
  uint32_t xph = 0;
  const int64_t tstep = (int64_t)gridDim.x * TG;
  if ((tid & 127) == 0 && (int64_t)blockIdx.x * TG + g < nt128) {
    mbar_expect_tx(xbar, xbytes);
    bulk_g2s(xs0, xt + (size_t)((int64_t)blockIdx.x * TG + g) * d_pad * 128, xbytes, xbar);
  }
#pragma unroll 1
  for (int64_t t = (int64_t)blockIdx.x * TG + g; t < nt128; t += tstep) {
    const int64_t n = t * 128 + (tid & 127);
    const float wn = __ldg(wt + n);                                  // wt is padded to whole 256-sample tiles
    const bool more = t + tstep < nt128;
    float* sp = stash + ((size_t)t * K + net) * ST_TILE + (tid & 127);
    TC_T0();
    // ---- x -> A (hi / lo), ones column at index d ----------------------------------------------
    {
      float xv[64];
      mbar_wait(xbar, xph); xph ^= 1;
      load_x_cols(xv, xs, 128, d, ks1);
      TC_ADD(6);
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (c < ks1) {
          float hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; i++) split_tf32(xv[8 * c + i], hi[i], lo[i]);
          tm_st8(tl + ahi + 8 * c, hi); tm_st8(tl + alo + 8 * c, lo);
        }
      }
    }
    TC_ADD(8);
    wait_st();
    TC_ADD(4);
#define EPDE_TC_STAGE(ISSUE)                                                                         \
    TC_ADD(0); fence_before(); group_bar(g); TC_ADD(1);                                              \
    if (wq == 0) { if (elect_one()) { fence_after(); ISSUE; commit(bar); } __syncwarp(); }           \
    TC_ADD(2); mbar_wait(bar, ph); ph ^= 1; fence_after(); TC_ADD(3);
#define EPDE_TC_MV(m) issue_matvec<3>(tm + dcol, tm + ahi, tm + alo, dMh[m], dMl[m], idesc)
    EPDE_TC_STAGE(issue_matvec_rt(ks1, tm + dcol, tm + ahi, tm + alo, dW1h, dW1l, idesc);
                  if (more) { mbar_expect_tx(xbar, xbytes); bulk_g2s(xs0, xt + (size_t)(t + tstep) * d_pad * 128, xbytes, xbar); })
    float z[H], v[TKH], t1[H], t2[H];
    // ---- a1 ---------------------------------------------------------------------------------
    tm_ld20(tl + dcol, z); TC_ADD(5);
#pragma unroll
    for (int i = 0; i < H; i += 2) tanh_pair(z[i], z[i + 1], v[i], v[i + 1], t1[i], t1[i + 1]);
    v[20] = 1.f; v[21] = 0.f; v[22] = 0.f; v[23] = 0.f;
    stash_st(sp, 0, v);
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_TC_STAGE(EPDE_TC_MV(0))
    // ---- a2 ---------------------------------------------------------------------------------
    tm_ld20(tl + dcol, z); TC_ADD(5);
#pragma unroll
    for (int i = 0; i < H; i += 2) tanh_pair(z[i], z[i + 1], v[i], v[i + 1], t2[i], t2[i + 1]);
    stash_st(sp, 1, v);
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_TC_STAGE(EPDE_TC_MV(1))
    // ---- a3, y, g3 --------------------------------------------------------------------------
    tm_ld20(tl + dcol, z); TC_ADD(5);
    float y;
    {
      float2 yy = make_float2(b4r, 0.f);
#pragma unroll
      for (int i = 0; i < H; i += 2) {
        float a30, a31, t30, t31;
        tanh_pair(z[i], z[i + 1], a30, a31, t30, t31);
        __stcs(sp + (size_t)(2 * TH + i) * 128, a30); __stcs(sp + (size_t)(2 * TH + i + 1) * 128, a31);
        const float2 w2 = make_float2(w4r[i], w4r[i + 1]);
        yy = ffma2(w2, make_float2(a30, a31), yy);
        const float2 g3 = fmul2(make_float2(t30, t31), w2);           // g3 = (1 - a3^2) w4
        v[i] = g3.x; v[i + 1] = g3.y;
      }
      y = yy.x + yy.y;
    }
    v[20] = 0.f;
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_TC_STAGE(EPDE_TC_MV(3))                                  // u2 = W3^T g3
    tm_ld20(tl + dcol, z); TC_ADD(5);
#pragma unroll
    for (int i = 0; i < H; i += 2) { const float2 q = fmul2(make_float2(t2[i], t2[i + 1]), make_float2(z[i], z[i + 1])); v[i] = q.x; v[i + 1] = q.y; }   // g2
#pragma unroll
    for (int i = 0; i < H; i += 2) mg2 = fmaxf(mg2, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
    stash_st(sp, 3, v);
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_TC_STAGE(EPDE_TC_MV(2))                                  // u1 = W2^T g2
    tm_ld20(tl + dcol, z); TC_ADD(5);
#pragma unroll
    for (int i = 0; i < H; i += 2) { const float2 q = fmul2(make_float2(t1[i], t1[i + 1]), make_float2(z[i], z[i + 1])); v[i] = q.x; v[i + 1] = q.y; }   // g1
#pragma unroll
    for (int i = 0; i < H; i += 2) mg1 = fmaxf(mg1, fmaxf(fabsf(v[i]), fabsf(v[i + 1])));
    stash_st(sp, 4, v);
    store_vec24(tl + ahi, tl + alo, v);
    EPDE_TC_STAGE(EPDE_TC_MV(4))                                  // K g1
    tm_ld20(tl + dcol, z); TC_ADD(5);
    float e0 = 0.f, e1 = 0.f;
#pragma unroll
    for (int i = 0; i < H; i += 2) { e0 = fmaf(v[i], z[i], e0); e1 = fmaf(v[i + 1], z[i + 1], e1); mt = fmaxf(mt, fmaxf(fabsf(z[i]), fabsf(z[i + 1]))); }
    const float e = e0 + e1;
    my = fmaxf(my, fabsf(y));
#undef EPDE_TC_STAGE
#undef EPDE_TC_MV
    if (n < B) ybuf[n * K + net] = y;
    accS += (double)(wn * y);
    accE += (double)(wn * e);
  }
  // block reduction in a fixed order
  const double s0 = warp_sum(accS), s1 = warp_sum(accE);
  if ((tid & 31) == 0) { red[warp][0] = s0; red[warp][1] = s1; }
  fence_before();
  __syncthreads();
  if (tid < 2) {
    double v = 0.0;
    for (int q = 0; q < TTHREADS / 32; q++) v += red[q][tid];
    part[((size_t)net * gridDim.x + blockIdx.x) * 2 + tid] = v;
  }
  {  // maxima: non-negative floats order like their bit patterns
    const uint32_t r0 = __reduce_max_sync(0xffffffffu, __float_as_uint(mg2)), r1 = __reduce_max_sync(0xffffffffu, __float_as_uint(mg1));
    const uint32_t r2 = __reduce_max_sync(0xffffffffu, __float_as_uint(mt)), r3 = __reduce_max_sync(0xffffffffu, __float_as_uint(my));
    if ((tid & 31) == 0) {
      uint32_t* m = reinterpret_cast<uint32_t*>(mx) + 2 + kMxPerNet * net;
      volatile uint32_t* mv = m;
      if (r0 > mv[0]) atomicMax(m + 0, r0);
      if (r1 > mv[1]) atomicMax(m + 1, r1);
      if (r2 > mv[2]) atomicMax(m + 2, r2);
      if (r3 > mv[3]) atomicMax(m + 3, r3);
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tm));
}

// W = sum w and S2[i][j] = sum w y_i y_j from wt / ybuf: one fp64 partial row per CTA, part2[cta][1 + K(K+1)/2]
// (latency-bound: 1024 threads per CTA and two samples per iteration keep 8x the loads of the first version in flight)
constexpr int SY_THREADS = 1024;
template <int K>
__global__ void __launch_bounds__(SY_THREADS)
k_sums_y(const float* __restrict__ wt, const float* __restrict__ ybuf, int64_t B, double* __restrict__ part2) {
  constexpr int NS2 = 1 + K * (K + 1) / 2;
  pdl_trigger();                     // the reduction kernel may take its place (it waits for this grid before reading)
  double acc[NS2];
#pragma unroll
  for (int s = 0; s < NS2; s++) acc[s] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * SY_THREADS;
  for (int64_t n0 = (int64_t)blockIdx.x * SY_THREADS + threadIdx.x; n0 < B; n0 += 2 * stride) {
    const int64_t n1 = n0 + stride;
    const bool v1 = n1 < B;
    float w[2], y[2][K];
    w[0] = __ldg(wt + n0); w[1] = v1 ? __ldg(wt + n1) : 0.f;
#pragma unroll
    for (int i = 0; i < K; i++) { y[0][i] = __ldg(ybuf + n0 * K + i); y[1][i] = v1 ? __ldg(ybuf + n1 * K + i) : 0.f; }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      acc[0] += (double)w[u];
      int s = 1;
#pragma unroll
      for (int i = 0; i < K; i++) {
        const float wy = w[u] * y[u][i];
#pragma unroll
        for (int j = 0; j <= i; j++) { acc[s] += (double)(wy * y[u][j]); s++; }
      }
    }
  }
  __shared__ double red[SY_THREADS / 32][NS2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < NS2; s++) {
    const double v = warp_sum(acc[s]);
    if (lane == 0) red[wid][s] = v;
  }
  __syncthreads();
  if (threadIdx.x < NS2) {
    double v = 0.0;
    for (int q = 0; q < SY_THREADS / 32; q++) v += red[q][threadIdx.x];
    part2[(size_t)blockIdx.x * NS2 + threadIdx.x] = v;
  }
}

// sums = [W, S1(k), E(k), S2(k(k+1)/2)] from the two partial arrays; one warp per sum, fixed (lane-strided, then
// shuffle-tree) order
__device__ __forceinline__ void reduce_tc_body(const double* __restrict__ part, int gx, const double* __restrict__ part2, int g2,
                                               int K, double* sums, double* ssums = nullptr) {
  const int lane = threadIdx.x & 31;
  const int ns2 = 1 + K * (K + 1) / 2, ns = 1 + 2 * K + K * (K + 1) / 2;       // 34 sums for k = 6
  for (int s = threadIdx.x >> 5; s < ns; s += blockDim.x >> 5) {
    double v = 0.0;
    if (s == 0 || s > 2 * K) {
      const int c = (s == 0) ? 0 : s - 2 * K;
      for (int b = lane; b < g2; b += 32) v += part2[(size_t)b * ns2 + c];
    } else {
      const int net = (s - 1) % K, which = (s - 1) / K;
      for (int b = lane; b < gx; b += 32) v += part[((size_t)net * gx + b) * 2 + which];
    }
    v = warp_sum(v);
    if (lane == 0) { sums[s] = v; if (ssums) ssums[s] = v; }
  }
}
__global__ void __launch_bounds__(1024)
k_reduce_tc(const double* __restrict__ part, int gx, const double* __restrict__ part2, int g2, int K,
            double* __restrict__ sums) {
  reduce_tc_body(part, gx, part2, g2, K, sums);
}
// single-rank step: the reduction of the partials and the coefficient kernel in one launch (no exchange in between)
__global__ void __launch_bounds__(1024)
k_reduce_coef_tc(const double* __restrict__ part, int gx, const double* __restrict__ part2, int g2, int K, double* sums,
                 double beta, double alpha, EigW eigw, int sort, double* __restrict__ out, Coef* __restrict__ coef,
                 const float* __restrict__ mx, const float* __restrict__ nrm, Peer peer) {
  pdl_trigger();                     // phase 2's CTAs may come in and set up shared / tensor memory while this CTA works
  pdl_wait();                        // k_sums_y (and with it phase 1) complete
  __shared__ double ssums[1 + 2 * EPDE_MAX_K + EPDE_MAX_K * (EPDE_MAX_K + 1) / 2];      // the coefficient part reads the sums from here
  reduce_tc_body(part, gx, part2, g2, K, sums, ssums);
  __syncthreads();
  if (peer.bufs) {      // sharded step: the sums of all ranks, exchanged over NVLink peer memory by this very CTA
    const uint32_t epoch = *peer.epoch + 1u;
    peer_allreduce_block<double>(peer, epoch, ssums, ssums, sums, 1 + 2 * K + K * (K + 1) / 2);
    __syncthreads();
    if (threadIdx.x == 0) *peer.epoch = epoch;
    // the maxima behind phase 2's staging scales stay per rank: every rank scales its own tiles
  }
  if (threadIdx.x < 32) coef_body<0>(ssums, K, beta, alpha, eigw, sort, out, coef, mx, nrm);
}

}  // namespace tc
}  // namespace epde
