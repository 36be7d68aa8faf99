// Probe: kind::f16 SS-form MMAs with MN-MAJOR, no-swizzle operands (sm_100a).
// Images are [sample][feature]: element (f, n) at (f / 8) * S_MN + (n / 8) * 128 + (n % 8) * 16 + (f % 8) * 2 bytes, i.e. a
// core matrix = 8 samples (K) x 8 features (MN, contiguous 16 bytes), core matrices contiguous along K.  A thread (= sample)
// writes 8 features with ONE 16-byte store.  Both LBO / SBO assignments are tried; G = P^T Q against fp64; cycles per MMA.
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
__device__ __forceinline__ uint32_t idesc_f16_mn(int M, int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int it = 0; it < (1 << 22); it++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void ld8(uint32_t t, float* r) {
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(t));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; i++) r[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr uint32_t S_K = 128, S_MN = 16 * 128;         // core matrices contiguous along K (16 per 128 samples); 2048 B per 8 features
constexpr uint32_t IMG = 8 * S_MN;                      // 64 features

struct Out { float G[4][64 * 64]; long long cyc[8]; int fail; };

__global__ void __launch_bounds__(128, 1)
k_probe(const float* __restrict__ P, const float* __restrict__ Q, Out* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Ph = smem; unsigned char* Pl = Ph + IMG; unsigned char* Qh = Pl + IMG; unsigned char* Ql = Qh + IMG;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Ql + IMG);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); out->fail = 0; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_before(); __syncthreads(); fence_after();
  const uint32_t tm = *slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;
  // staging: thread = sample n; 8 features per 16-byte store
  const int n = tid;
  long long ts0 = clock64();
  for (int fg = 0; fg < 8; fg++) {
    uint32_t ph[4], pl[4], qh[4], ql[4];
    for (int j = 0; j < 4; j++) {
      const float a0 = P[n * 64 + 8 * fg + 2 * j], a1 = P[n * 64 + 8 * fg + 2 * j + 1];
      __half2 h = __floats2half2_rn(a0, a1); float2 hf = __half22float2(h);
      __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
      ph[j] = *reinterpret_cast<uint32_t*>(&h); pl[j] = *reinterpret_cast<uint32_t*>(&l);
      const float b0 = Q[n * 64 + 8 * fg + 2 * j], b1 = Q[n * 64 + 8 * fg + 2 * j + 1];
      h = __floats2half2_rn(b0, b1); hf = __half22float2(h);
      l = __floats2half2_rn(b0 - hf.x, b1 - hf.y);
      qh[j] = *reinterpret_cast<uint32_t*>(&h); ql[j] = *reinterpret_cast<uint32_t*>(&l);
    }
    const uint32_t o = (uint32_t)fg * S_MN + (uint32_t)(n >> 3) * S_K + (uint32_t)(n & 7) * 16u;
    *reinterpret_cast<uint4*>(Ph + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(Pl + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    *reinterpret_cast<uint4*>(Qh + o) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
    *reinterpret_cast<uint4*>(Ql + o) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
  }
  long long ts1 = clock64();
  if (tid == 0) out->cyc[7] = ts1 - ts0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before(); __syncthreads();
  // variant 0: LBO = S_K (K direction), SBO = S_MN     variant 1: LBO = S_MN, SBO = S_K
  for (int variant = 0; variant < 2; variant++) {
    fence_before(); __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        fence_after();
        const uint32_t id = idesc_f16_mn(64, 64);
        const uint32_t lbo = variant == 0 ? S_K : S_MN, sbo = variant == 0 ? S_MN : S_K;
        for (int pass = 0; pass < 3; pass++)
          for (int ks = 0; ks < 8; ks++) {               // 16 samples per MMA = two core matrices along K = 256 B
            const unsigned char* pa = pass == 0 ? Pl : Ph;
            const unsigned char* pb = pass == 1 ? Ql : Qh;
            mma_f16(tm + 64 * variant, make_desc(smem_u32(pa) + ks * 256, lbo, sbo), make_desc(smem_u32(pb) + ks * 256, lbo, sbo), id, (pass | ks) ? 1u : 0u);
          }
        commit(bar);
      }
      __syncwarp();
    }
    if (!mbar_wait(bar, phase)) { out->fail = 1 + variant; return; }
    phase ^= 1; fence_after();
    const int lane = tid & 31, r = 16 * warp + (lane & 15);
    for (int c = 0; c < 8; c++) {
      float v[8];
      ld8(tm + lane_base + 64 * variant + 8 * c, v);
      if (lane < 16) for (int i = 0; i < 8; i++) out->G[variant][r * 64 + 8 * c + i] = v[i];
    }
  }
  // timing: M64 N64 and M128 N64 MN-major, issue loop of 128 MMAs
  for (int t = 0; t < 3; t++) {
    fence_before(); __syncthreads();
    long long t0 = 0;
    if (warp == 0 && elect_one()) {
      fence_after();
      const uint32_t id = t == 0 ? idesc_f16_mn(64, 64) : (t == 1 ? idesc_f16_mn(64, 48) : idesc_f16_mn(128, 64));
      uint64_t da[4], db[4];
      for (int i = 0; i < 4; i++) { da[i] = make_desc(smem_u32(Ph) + i * 256, S_K, S_MN); db[i] = make_desc(smem_u32(Qh) + i * 256, S_K, S_MN); }
      t0 = clock64();
#pragma unroll 1
      for (int r = 0; r < 32; r++) {
        mma_f16(tm + 256, da[0], db[0], id, 1u); mma_f16(tm + 256, da[1], db[1], id, 1u);
        mma_f16(tm + 256, da[2], db[2], id, 1u); mma_f16(tm + 256, da[3], db[3], id, 1u);
      }
      commit(bar);
    }
    if (!mbar_wait(bar, phase)) { out->fail = 10 + t; return; }
    phase ^= 1; fence_after();
    if (t0) out->cyc[t] = clock64() - t0;
  }
  fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tm));
}

int main() {
  std::vector<float> P(128 * 64), Q(128 * 64);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& v : P) v = rnd() * 100.f;
  for (auto& v : Q) v = rnd() * 100.f;
  float *dP, *dQ; Out* dO;
  CK(cudaMalloc(&dP, P.size() * 4)); CK(cudaMalloc(&dQ, Q.size() * 4)); CK(cudaMalloc(&dO, sizeof(Out)));
  CK(cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 4 * IMG + 64;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static Out o;
  CK(cudaMemset(dO, 0, sizeof(Out)));
  k_probe<<<1, 128, smem>>>(dP, dQ, dO);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(&o, dO, sizeof(Out), cudaMemcpyDeviceToHost));
  printf("fail=%d\n", o.fail);
  for (int v = 0; v < 2; v++) {
    double e = 0, et = 0, s = 0;
    for (int p = 0; p < 64; p++) for (int q = 0; q < 64; q++) {
      double r = 0; for (int n = 0; n < 128; n++) r += (double)P[n * 64 + p] * Q[n * 64 + q];
      e = fmax(e, fabs(o.G[v][p * 64 + q] - r)); et = fmax(et, fabs(o.G[v][q * 64 + p] - r)); s = fmax(s, fabs(r));
    }
    printf("  MN-major variant %d (%s): max abs err %.3e, transposed readout %.3e (max |ref| %.1f)  G[1][2]=%.3f\n", v,
           v == 0 ? "LBO = K stride, SBO = MN stride" : "LBO = MN stride, SBO = K stride", e, et, s, o.G[v][64 + 2]);
  }
  printf("  timing MN-major f16: M64 N64 %.1f  M64 N48 %.1f  M128 N64 %.1f cycles/MMA;  staging of 4 x 64 rows: %lld cycles (thread 0)\n",
         o.cyc[0] / 128.0, o.cyc[1] / 128.0, o.cyc[2] / 128.0, o.cyc[7]);
  return 0;
}
