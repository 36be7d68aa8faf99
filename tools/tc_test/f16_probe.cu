// Probe for the fp16 two-term split of the gradient rounds (sm_100a, tcgen05 kind::f16, SS form):
//  A  G[p][q] = sum_n P[n][p] Q[n][q] over 128 samples with P = Ph + Pl, Q = Qh + Ql (fp16 each, scaled by a power of
//     two), three products Pl*Qh + Ph*Ql + Ph*Qh, K-major [feature][sample] images (core matrix = 8 rows x 8 halves,
//     LBO 144 B between 8-sample chunks, SBO = 16 LBO between 8-row groups); error against fp64 and against 3xTF32.
//  B  cycles per MMA: kind::f16 M64 / M128, several N (K = 16 samples per MMA), issue-side blocking.
//  C  M = 64 accumulator at lane offset 16 (can two M = 64 accumulators share one column range?)  -- run last.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// D fp32 (1 << 4), A / B format 0 = F16, both K-major
__device__ __forceinline__ uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
// both MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t idesc_f16_mn(int M, int N) { return idesc_f16(M, N) | (1u << 15) | (1u << 16); }
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); it++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ld8(uint32_t t, float* r) {
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(t));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; i++) r[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void st8(uint32_t t, const float* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(t), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])),
                  "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])),
                  "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr uint32_t LBO = 144, SBO16 = 16 * LBO, SBO32 = 32 * LBO;   // fp16: 16 chunks of 8 samples; tf32: 32 chunks of 4 samples
constexpr uint32_t IMG16 = 8 * SBO16;    // 64 rows (the timing tests read across image boundaries: harmless)
constexpr uint32_t IMG32 = 8 * SBO32;    // 64 rows (tf32 reference)

struct Out {
  float G16[64 * 64];
  float G32[64 * 64];
  float G16off[64 * 64];
  long long cyc[24];
  int fail;
};

__global__ void __launch_bounds__(128, 1)
k_probe(const float* __restrict__ P, const float* __restrict__ Q, float sP, float sQ, Out* out, int do_c) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Ph = smem;                 // fp16 images (128 rows each)
  unsigned char* Pl = Ph + IMG16;
  unsigned char* Qh = Pl + IMG16;
  unsigned char* Ql = Qh + IMG16;
  unsigned char* T = Ql + IMG16;            // tf32 images, 64 rows: Ph Pl Qh Ql
  uint64_t* bar = reinterpret_cast<uint64_t*>(T + 4 * IMG32);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); out->fail = 0; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (uint32_t i = tid; i < (4 * IMG16 + 4 * IMG32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tm = *slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  // ---- staging exactly as the kernel would do it: thread = sample n, loops over feature rows -----------------
  const int n = tid;
  for (int r = 0; r < 64; r++) {
    {
      const float v = P[n * 64 + r] * sP;
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const uint32_t o = (uint32_t)(r >> 3) * SBO16 + (uint32_t)(n >> 3) * LBO + (uint32_t)(r & 7) * 16u + (uint32_t)(n & 7) * 2u;
      *reinterpret_cast<__half*>(Ph + o) = h; *reinterpret_cast<__half*>(Pl + o) = l;
      const float x = P[n * 64 + r];
      const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u), lo = x - hi;
      const uint32_t o2 = (uint32_t)(r >> 3) * SBO32 + (uint32_t)(n >> 2) * LBO + (uint32_t)(r & 7) * 16u + (uint32_t)(n & 3) * 4u;
      *reinterpret_cast<float*>(T + o2) = hi; *reinterpret_cast<float*>(T + IMG32 + o2) = lo;
    }
    {
      const float v = Q[n * 64 + r] * sQ;
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      const uint32_t o = (uint32_t)(r >> 3) * SBO16 + (uint32_t)(n >> 3) * LBO + (uint32_t)(r & 7) * 16u + (uint32_t)(n & 7) * 2u;
      *reinterpret_cast<__half*>(Qh + o) = h; *reinterpret_cast<__half*>(Ql + o) = l;
      const float x = Q[n * 64 + r];
      const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u), lo = x - hi;
      const uint32_t o2 = (uint32_t)(r >> 3) * SBO32 + (uint32_t)(n >> 2) * LBO + (uint32_t)(r & 7) * 16u + (uint32_t)(n & 3) * 4u;
      *reinterpret_cast<float*>(T + 2 * IMG32 + o2) = hi; *reinterpret_cast<float*>(T + 3 * IMG32 + o2) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads();

  // ---- test A: fp16 round (D at column 0) and tf32 round (D at column 64) ---------------------------------------
  if (warp == 0) {
    if (elect_one()) {
      fence_after();
      const uint32_t id16 = idesc_f16(64, 64), id32 = idesc_tf32(64, 64);
      const uint64_t ph = make_desc(smem_u32(Ph), LBO, SBO16), pl = make_desc(smem_u32(Pl), LBO, SBO16);
      const uint64_t qh = make_desc(smem_u32(Qh), LBO, SBO16), ql = make_desc(smem_u32(Ql), LBO, SBO16);
#pragma unroll
      for (int ks = 0; ks < 8; ks++) {          // 16 samples per MMA = two chunks = 288 B = 18 address units
        mma_f16(tm, pl + 18 * ks, qh + 18 * ks, id16, ks ? 1u : 0u);
        mma_f16(tm, ph + 18 * ks, ql + 18 * ks, id16, 1u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ks++) mma_f16(tm, ph + 18 * ks, qh + 18 * ks, id16, 1u);
      const uint64_t th = make_desc(smem_u32(T), LBO, SBO32), tl = make_desc(smem_u32(T + IMG32), LBO, SBO32);
      const uint64_t uh = make_desc(smem_u32(T + 2 * IMG32), LBO, SBO32), ul = make_desc(smem_u32(T + 3 * IMG32), LBO, SBO32);
#pragma unroll
      for (int ks = 0; ks < 16; ks++) {
        mma_tf32(tm + 64, tl + 18 * ks, uh + 18 * ks, id32, ks ? 1u : 0u);
        mma_tf32(tm + 64, th + 18 * ks, ul + 18 * ks, id32, 1u);
      }
#pragma unroll
      for (int ks = 0; ks < 16; ks++) mma_tf32(tm + 64, th + 18 * ks, uh + 18 * ks, id32, 1u);
      commit(bar);
    }
    __syncwarp();
  }
  if (!mbar_wait(bar, phase)) { out->fail = 1; return; }
  phase ^= 1; fence_after();
  // M = 64 layout: row r in lane (r / 16) * 32 + r % 16
  {
    const int lane = tid & 31;
    const int r = 16 * warp + (lane & 15);
    for (int c = 0; c < 8; c++) {       // every lane executes the .sync.aligned loads; lanes < 16 hold the rows
      float v[8];
      ld8(tm + lane_base + 8 * c, v);
      if (lane < 16) for (int i = 0; i < 8; i++) out->G16[r * 64 + 8 * c + i] = v[i];
      ld8(tm + lane_base + 64 + 8 * c, v);
      if (lane < 16) for (int i = 0; i < 8; i++) out->G32[r * 64 + 8 * c + i] = v[i];
    }
  }
  fence_before(); __syncthreads(); fence_after();

  // ---- test B: timing ------------------------------------------------------------------------------------------
  // 0..4: f16 M64 N = 32, 48, 56, 64, 128     5..9: f16 M128 N = 64, 96, 128, 176, 256     10: tf32 M64 N64
  for (int t = 0; t < 11; t++) {
    fence_before(); __syncthreads();
    long long t0 = 0, t1 = 0;
    if (warp == 0 && elect_one()) {
      fence_after();
      const int Ns[11] = {32, 48, 56, 64, 128, 64, 96, 128, 176, 256, 64};
      const int M = (t >= 5 && t < 10) ? 128 : 64;
      const uint32_t id = t == 10 ? idesc_tf32(M, Ns[t]) : idesc_f16(M, Ns[t]);
      uint64_t da[4], db[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { da[i] = make_desc(smem_u32(Ph) + i * 288, LBO, SBO16); db[i] = make_desc(smem_u32(Pl) + i * 288, LBO, SBO16); }
      t0 = clock64();
      if (t == 10) {
#pragma unroll 1
        for (int r = 0; r < 32; r++) {
          mma_tf32(tm + 128, da[0], db[0], id, 1u); mma_tf32(tm + 128, da[1], db[1], id, 1u);
          mma_tf32(tm + 128, da[2], db[2], id, 1u); mma_tf32(tm + 128, da[3], db[3], id, 1u);
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < 32; r++) {
          mma_f16(tm + 128, da[0], db[0], id, 1u); mma_f16(tm + 128, da[1], db[1], id, 1u);
          mma_f16(tm + 128, da[2], db[2], id, 1u); mma_f16(tm + 128, da[3], db[3], id, 1u);
        }
      }
      t1 = clock64();
      commit(bar);
    }
    if (!mbar_wait(bar, phase)) { out->fail = 10 + t; return; }
    phase ^= 1; fence_after();
    if (t0) { out->cyc[t] = clock64() - t0; out->cyc[12 + t] = t1 - t0; }
  }

  // ---- test C: M = 64 accumulator with a lane offset of 16 in the D address ------------------------------------
  if (do_c) {
    fence_before(); __syncthreads();
    { float sent[8] = {-7.f, -7.f, -7.f, -7.f, -7.f, -7.f, -7.f, -7.f};
      for (int c = 0; c < 8; c++) st8(tm + lane_base + 256 + 8 * c, sent);
      wait_st(); }
    fence_before(); __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        fence_after();
        const uint32_t id16 = idesc_f16(64, 64);
        const uint64_t ph = make_desc(smem_u32(Ph), LBO, SBO16), pl = make_desc(smem_u32(Pl), LBO, SBO16);
        const uint64_t qh = make_desc(smem_u32(Qh), LBO, SBO16), ql = make_desc(smem_u32(Ql), LBO, SBO16);
        const uint32_t dd = tm + 256 + (16u << 16);
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
          mma_f16(dd, pl + 18 * ks, qh + 18 * ks, id16, ks ? 1u : 0u);
          mma_f16(dd, ph + 18 * ks, ql + 18 * ks, id16, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ks++) mma_f16(dd, ph + 18 * ks, qh + 18 * ks, id16, 1u);
        commit(bar);
      }
      __syncwarp();
    }
    if (!mbar_wait(bar, phase)) { out->fail = 40; return; }
    phase ^= 1; fence_after();
    // dump lanes: thread tid reads its own lane, columns 256..319; G16off[row-as-lane-sees-it]
    for (int c = 0; c < 8; c++) {
      float v[8];
      ld8(tm + lane_base + 256 + 8 * c, v);
      const int lane = tid & 31;
      if (lane >= 16) { const int r = 16 * warp + (lane - 16); for (int i = 0; i < 8; i++) out->G16off[r * 64 + 8 * c + i] = v[i]; }
      else if (c == 0 && v[0] != -7.f) out->fail = 41;     // the lower half must stay untouched
    }
  }
  fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tm));
}

int main(int argc, char** argv) {
  const int do_c = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> P(128 * 64), Q(128 * 64);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  // mixed magnitudes: most samples O(1), some tiny, some near the top of the range
  for (int n = 0; n < 128; n++) {
    const float mp = (n % 7 == 0) ? 1e-4f : (n % 11 == 0 ? 30.f : 1.f), mq = (n % 5 == 0) ? 1e-3f : 1.f;
    for (int r = 0; r < 64; r++) { P[n * 64 + r] = rnd() * mp; Q[n * 64 + r] = rnd() * mq; }
  }
  float *dP, *dQ; Out* dO;
  CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dQ, Q.size() * 4)); CK(cudaMalloc(&dO, sizeof(Out)));
  CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 4 * IMG16 + 4 * IMG32 + 64;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static Out o;
  const float scalesP[3] = {1024.f, 1.f, 1.f / 1024.f}, scalesQ[3] = {16384.f, 1.f, 1.f / 64.f};
  for (int s = 0; s < 3; s++) {
    CK(cudaMemset(dO, 0, sizeof(Out)));
    k_probe<<<1, 128, smem>>>(dP, dQ, scalesP[s], scalesQ[s], dO, s == 2 ? do_c : 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&o, dO, sizeof(Out), cudaMemcpyDeviceToHost));
    printf("scale P %g Q %g: fail=%d\n", scalesP[s], scalesQ[s], o.fail);
    double e16 = 0, e32 = 0, eoff = 0, sref = 0, rms = 0;
    const double inv = 1.0 / ((double)scalesP[s] * scalesQ[s]);
    for (int p = 0; p < 64; p++) for (int q = 0; q < 64; q++) {
      double r = 0; for (int n = 0; n < 128; n++) r += (double)P[n * 64 + p] * Q[n * 64 + q];
      e16 = fmax(e16, fabs(o.G16[p * 64 + q] * inv - r)); e32 = fmax(e32, fabs(o.G32[p * 64 + q] - r));
      eoff = fmax(eoff, fabs(o.G16off[p * 64 + q] * inv - r));
      sref = fmax(sref, fabs(r)); rms += r * r;
    }
    printf("  A fp16 2-split round: max abs err %.3e   3xTF32: %.3e   (max |ref| %.3f, rms %.3f)\n", e16, e32, sref, sqrt(rms / 4096));
    if (s == 2 && do_c) printf("  C lane-offset-16 accumulator: max abs err %.3e\n", eoff);
    if (s == 0) {
      const char* names[11] = {"f16 M64 N32", "f16 M64 N48", "f16 M64 N56", "f16 M64 N64", "f16 M64 N128", "f16 M128 N64", "f16 M128 N96",
                               "f16 M128 N128", "f16 M128 N176", "f16 M128 N256", "tf32 M64 N64"};
      for (int t = 0; t < 11; t++) printf("  B %-16s %.1f cycles/MMA (issue side: %.1f)\n", names[t], o.cyc[t] / 128.0, o.cyc[12 + t] / 128.0);
    }
  }
  return 0;
}
