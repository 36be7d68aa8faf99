// Stage-A probe for the tcgen05 path: 3xTF32 split GEMMs on sm_100a with hand-built descriptors.
//  test 1: D[128x32] = A[128x24] * B^T, A/B K-major, no swizzle, M=128 N=32 (mat-vec role)
//  test 2: G[64x32]  = P^T Q over K=128 samples, operands MN-major views of the same tiles
// Prints max errors against an fp64 host result and, for test 2, the TMEM lane map.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// shared memory matrix descriptor, no swizzle, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // version
  return d;                          // layout_type (bits 61-63) = 0: SWIZZLE_NONE
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

constexpr int KP = 24;        // padded contraction length of the mat-vec (20 -> 24)
constexpr int TILE_BYTES = 128 * KP * 4;   // one [128 x 24] fp32 operand tile = 12288 B
// canonical K-major / no-swizzle offset (floats) of element (row m, k)
__device__ __host__ inline int koff(int m, int k) { return (m >> 3) * (6 * 32) + (k >> 2) * 32 + (m & 7) * 4 + (k & 3); }

__global__ void __launch_bounds__(128, 1)
k_test(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ Q,
       float* __restrict__ D, float* __restrict__ Gout, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* Ahi = reinterpret_cast<float*>(smem);
  float* Alo = Ahi + 128 * KP;
  float* Qhi = Alo + 128 * KP;
  float* Qlo = Qhi + 128 * KP;
  float* Bhi = Qlo + 128 * KP;              // [32 x 24]
  float* Blo = Bhi + 32 * KP;
  float* QThi = Blo + 32 * KP;              // K-major B operand of test 2: [32 rows i][128 samples]
  float* QTlo = QThi + 32 * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- stage operands (hi/lo split), rows = samples --------------------------------------
  for (int k = 0; k < KP; k++) {
    const float a = (k < 20) ? A[tid * 20 + k] : 0.f;
    const float ah = tf32_rna(a);
    Ahi[koff(tid, k)] = ah; Alo[koff(tid, k)] = tf32_rna(a - ah);
    const float q = (k < 20) ? Q[tid * 20 + k] : 0.f;
    const float qh = tf32_rna(q);
    Qhi[koff(tid, k)] = qh; Qlo[koff(tid, k)] = tf32_rna(q - qh);
  }
  for (int e = tid; e < 32 * KP; e += 128) {          // B[n][k] = W[k][n]
    const int n = e / KP, k = e % KP;
    const float b = (n < 20 && k < 20) ? W[k * 20 + n] : 0.f;
    const float bh = tf32_rna(b);
    Bhi[koff(n, k)] = bh; Blo[koff(n, k)] = tf32_rna(b - bh);
  }
  for (int i = 0; i < 32; i++) {             // QT[i][s=tid]: (i%8)*4 + (i/8)*1024 + (s/4)*32 + s%4 floats
    const float q = (i < 20) ? Q[tid * 20 + i] : 0.f;
    const float qh = tf32_rna(q);
    const int off = (i & 7) * 4 + (i >> 3) * 1024 + (tid >> 2) * 32 + (tid & 3);
    QThi[off] = qh; QTlo[off] = tf32_rna(q - qh);
  }
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;

  // ---- test 1: D[128 x 32] (columns 0..31) ------------------------------------------------
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, 32, 0, 0);
    const uint32_t ah = smem_u32(Ahi), al = smem_u32(Alo), bh = smem_u32(Bhi), bl = smem_u32(Blo);
    uint32_t acc = 0;
    for (int ks = 0; ks < 3; ks++) {                  // K = 8 per MMA = two 16-byte k-chunks (LBO = 128 B)
      const uint32_t o = ks * 256;
      mma_tf32(tm, make_desc(ah + o, 128, 768), make_desc(bh + o, 128, 768), idesc, acc); acc = 1;
      mma_tf32(tm, make_desc(ah + o, 128, 768), make_desc(bl + o, 128, 768), idesc, 1);
      mma_tf32(tm, make_desc(al + o, 128, 768), make_desc(bh + o, 128, 768), idesc, 1);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 32; n++) D[tid * 32 + n] = __uint_as_float(r[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- test 2: G[64 x 32] = A^T Q over the 128 samples (columns 32..63), MN-major operands -
  if (tid == 0) {
    const int M2 = (variant & 4) ? 128 : 64;
    const bool b_k = !(variant & 8);                  // B operand K-major (transposed copy) unless bit 3
    const bool swap = variant & 1;                    // swap LBO / SBO of the MN-major operands
    const uint32_t idesc = make_idesc(M2, 32, 1, b_k ? 0 : 1);
    const uint32_t ah = smem_u32(Ahi), al = smem_u32(Alo), qh = smem_u32(Qhi), ql = smem_u32(Qlo);
    const uint32_t th = smem_u32(QThi), tl = smem_u32(QTlo);
    const uint32_t l = swap ? 128 : 768, sb = swap ? 768 : 128;
    uint32_t acc = 0;
    if (variant & 16) {                               // control: repeat test 1 into columns 32..63
      const uint32_t id1 = make_idesc(128, 32, 0, 0), bh = smem_u32(Bhi), bl = smem_u32(Blo);
      for (int ks = 0; ks < 3; ks++) {
        const uint32_t o = ks * 256;
        mma_tf32(tm + 32, make_desc(ah + o, 128, 768), make_desc(bh + o, 128, 768), id1, acc); acc = 1;
        mma_tf32(tm + 32, make_desc(ah + o, 128, 768), make_desc(bl + o, 128, 768), id1, 1);
        mma_tf32(tm + 32, make_desc(al + o, 128, 768), make_desc(bh + o, 128, 768), id1, 1);
      }
    } else
    for (int ks = 0; ks < 16; ks++) {                 // 8 samples per MMA = one 8-row group (768 B apart)
      const uint32_t o = ks * 768;
      const uint64_t dah = make_desc(ah + o, l, sb), dal = make_desc(al + o, l, sb);
      const uint64_t dbh = b_k ? make_desc(th + ks * 256, 128, 4096) : make_desc(qh + o, l, sb);
      const uint64_t dbl = b_k ? make_desc(tl + ks * 256, 128, 4096) : make_desc(ql + o, l, sb);
      mma_tf32(tm + 32, dah, dbh, idesc, acc); acc = 1;
      mma_tf32(tm + 32, dah, dbl, idesc, 1);
      mma_tf32(tm + 32, dal, dbh, idesc, 1);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 32; n++) Gout[tid * 32 + n] = __uint_as_float(r[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tm));
}

int main() {
  std::vector<float> A(128 * 20), W(20 * 20), Q(128 * 20), D(128 * 32), G(128 * 32);
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& v : A) v = rnd();
  for (auto& v : W) v = rnd();
  for (auto& v : Q) v = rnd();
  float *dA, *dW, *dQ, *dD, *dG;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dQ, Q.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dG, G.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dG, 0, G.size() * 4));
  const int smem = 4 * TILE_BYTES + 2 * 32 * KP * 4 + 2 * 32 * 128 * 4 + 1024;
  CK(cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<double> Gr(20 * 20);
  for (int o = 0; o < 20; o++) for (int i = 0; i < 20; i++) { double r = 0; for (int n = 0; n < 128; n++) r += (double)A[n * 20 + o] * Q[n * 20 + i]; Gr[o * 20 + i] = r; }
  for (int variant : {16, 0, 4}) {
    CK(cudaMemset(dG, 0, G.size() * 4));
    k_test<<<1, 128, smem>>>(dA, dW, dQ, dD, dG, variant);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0;
    for (int m = 0; m < 128; m++) for (int n = 0; n < 20; n++) {
      double ref = 0; for (int k = 0; k < 20; k++) ref += (double)A[m * 20 + k] * W[k * 20 + n];
      e1 = fmax(e1, fabs(ref - D[m * 32 + n]));
    }
    int nz = 0; for (auto v : G) nz += (v != 0.f);
    printf("variant %2d: test1 err %.2e | test2 nonzero %d |", variant, e1, nz);
    for (int o : {0, 1, 5, 16, 19}) {
      int best = -1; double be = 1e9;
      for (int lane = 0; lane < 128; lane++) { double e = 0; for (int i = 0; i < 20; i++) e = fmax(e, fabs(Gr[o * 20 + i] - G[lane * 32 + i])); if (e < be) { be = e; best = lane; } }
      printf(" o%d->lane%d(%.1e)", o, best, be);
    }
    printf("\n");
    if (variant & 16) { double e = 0; for (int m = 0; m < 128; m++) for (int n = 0; n < 20; n++) e = fmax(e, fabs(G[m*32+n] - D[m*32+n])); printf("   control: max |G - D| = %.3e\n", e); }
  }
  return 0;
}
