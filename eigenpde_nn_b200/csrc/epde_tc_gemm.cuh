// tcgen05 (5th-gen tensor core) FP32-accurate GEMM for the generic layer-wise path:
//     C[M,N] = sum_k A(m,k) B(k,n)  (+ bias[n])      fp32 in, fp32 out
// with every fp32 operand split on the fly into two TF32 terms (x = hi + lo, both produced by
// cvt.rna.tf32) and three MMAs per product (hi*hi + hi*lo + lo*hi, "3xTF32"): the dropped lo*lo term
// is 2^-22 relative, i.e. fp32-level accuracy (measured against fp64: 4e-7 relative,
// tools/tc_test/tc_gemm_test.cu).  Accumulators live in TMEM (fp32), one 128 x 128 tile per CTA.
//
// Operand staging: tiles are written to shared memory in the canonical K-major / no-swizzle UMMA
// layout (8-row x 16-byte core matrices; include/cute mma_sm100_desc.hpp "INTERLEAVE"), with a
// leading-dimension byte offset of 144 and a stride byte offset of 1168 so that both the plain and
// the transposing loader store (almost) bank-conflict free.  The loader transposes on the fly, so the
// three operand orientations of the layer-wise path (A or B given as [K][MN]) need no extra copies.
// Double-buffered: the MMAs of k-block i run while block i+1 is loaded, split and stored.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace epde {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 16, TC_THREADS = 256;   // BK = 16: 75 KB/CTA -> 3 CTAs per SM
constexpr uint32_t TC_KCH = TC_BK / 4;                            // 16-byte k-chunks per row
constexpr uint32_t TC_LBO = 160, TC_SBO = TC_KCH * TC_LBO + 16;  // bytes (LBO: 2 bank-quads per chunk; SBO = 656 = 4 words mod 32 banks)
constexpr int TC_V = 128 * (TC_BK / 4) / TC_THREADS;               // float4 per thread per tile
constexpr uint32_t TC_TILE = 16 * TC_SBO;                       // one 128-row operand tile (hi or lo)
constexpr uint32_t TC_STAGE = 4 * TC_TILE;                      // A hi, A lo, B hi, B lo
constexpr uint32_t TC_SMEM = 2 * TC_STAGE + 64;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tc_tf32(float x) {
  uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r);
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {    // K-major, SWIZZLE_NONE, version 1
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tTC_WAIT:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
               "@P1 bra TC_DONE;\n\tbra TC_WAIT;\n\tTC_DONE:\n\t}\n" :: "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One 128 x TC_BK operand tile (rows = MN index), TC_V float4 per thread.
//   TR = false: source element (mn, k) at src[mn * ld + k]   (K contiguous)
//   TR = true : source element (mn, k) at src[k * ld + mn]   (MN contiguous) -- transposed while storing
// tc_tile_load issues the global loads (kept in registers across the MMA issue of the previous k-block),
// tc_tile_store splits every value into hi / lo TF32 parts and writes the canonical K-major tiles.
template <bool TR>
__device__ __forceinline__ void tc_tile_load(const float* __restrict__ src, int ld, int mn0, int mn_end, int k0,
                                             int k_end, float4 (&v)[TC_V]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < TC_V; i++) {
    const int e = t + TC_THREADS * i;                            // 128 * TC_KCH float4 per tile
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!TR) {
      const int gm = mn0 + e / (int)TC_KCH, gk = k0 + 4 * (e % (int)TC_KCH);
      if (gm < mn_end && gk < k_end) v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gm * ld + gk));
    } else {
      const int gk = k0 + 4 * (e >> 7) + (e & 3), gm = mn0 + 4 * ((e & 127) >> 2);     // see tc_tile_store
      if (gk < k_end && gm < mn_end) v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gk * ld + gm));
    }
  }
}
template <bool TR>
__device__ __forceinline__ void tc_tile_store(const float4 (&v)[TC_V], unsigned char* hi, unsigned char* lo) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < TC_V; i++) {
    const int e = t + TC_THREADS * i;
    if (!TR) {
      const int row = e / (int)TC_KCH, chunk = e % (int)TC_KCH;
      const float4 h = make_float4(tc_tf32(v[i].x), tc_tf32(v[i].y), tc_tf32(v[i].z), tc_tf32(v[i].w));
      const float4 l = make_float4(tc_tf32(v[i].x - h.x), tc_tf32(v[i].y - h.y), tc_tf32(v[i].z - h.z),
                                   tc_tf32(v[i].w - h.w));
      const uint32_t off = (row >> 3) * TC_SBO + chunk * TC_LBO + (row & 7) * 16;
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    } else {
      // a warp covers 4 k x 8 float4 of mn (four 128-byte row segments in global memory): its 32 scalar stores of
      // one j then hit 32 different banks (k & 3 -> word, (mc & 1) -> +16 words, mc / 2 -> +4 words per 8-row group)
      const int k = 4 * (e >> 7) + (e & 3), mc = (e & 127) >> 2;
      const float vv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int row = 4 * mc + j;
        const float h = tc_tf32(vv[j]);
        const uint32_t off = (row >> 3) * TC_SBO + (k >> 2) * TC_LBO + (row & 7) * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(hi + off) = h;
        *reinterpret_cast<float*>(lo + off) = tc_tf32(vv[j] - h);
      }
    }
  }
}

// TA: A given as [K][M] (else [M][K]);  TBK: B given as [N][K] (K-major, else [K][N]).
// blockIdx = (n tile, m tile, k split); C (or the split's partial C) is plain fp32 row-major.
template <bool TA, bool TBK>
__global__ void __launch_bounds__(TC_THREADS, 2)
g_gemm_tc(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
          float* __restrict__ C, int ldc, const float* __restrict__ bias, int ksplit, size_t split_stride) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + 2 * TC_STAGE);     // empty[0], empty[1], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
  const int kb = blockIdx.z * ksplit, ke = min(K, kb + ksplit);
  const int bn = min(TC_BN, (N - n0 + 15) & ~15);               // MMA N: multiple of 16
  if (tid == 0) {
    for (int i = 0; i < 3; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem_u32(bars + i)));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(tc_smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  const int nk = (ke - kb + TC_BK - 1) / TC_BK;
  float4 ra[TC_V], rb[TC_V];                                     // raw operands of the NEXT k-block
  tc_tile_load<TA>(A, lda, m0, M, kb, ke, ra);
  tc_tile_load<!TBK>(B, ldb, n0, N, kb, ke, rb);
  for (int it = 0; it < nk; it++) {
    const int s = it & 1;
    unsigned char* st = tc_smem + s * TC_STAGE;
    if (it >= 2) tc_mbar_wait(bars + s, ((it >> 1) - 1) & 1);    // MMAs that read this stage are done
    tc_tile_store<TA>(ra, st, st + TC_TILE);
    tc_tile_store<!TBK>(rb, st + 2 * TC_TILE, st + 3 * TC_TILE);
    if (it + 1 < nk) {                                           // global loads of block it+1 fly during the MMAs
      const int k1 = kb + (it + 1) * TC_BK;
      tc_tile_load<TA>(A, lda, m0, M, k1, ke, ra);
      tc_tile_load<!TBK>(B, ldb, n0, N, k1, ke, rb);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah = tc_smem_u32(st), al = ah + TC_TILE, bh = ah + 2 * TC_TILE, bl = ah + 3 * TC_TILE;
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ks++) {                   // K = 8 per MMA = two 16-byte k-chunks
        const uint32_t o = ks * 2 * TC_LBO;
        tc_mma(tm, tc_desc(ah + o), tc_desc(bh + o), idesc, (it | ks) ? 1u : 0u);
        tc_mma(tm, tc_desc(ah + o), tc_desc(bl + o), idesc, 1u);
        tc_mma(tm, tc_desc(al + o), tc_desc(bh + o), idesc, 1u);
      }
      tc_commit(bars + s);
      if (it == nk - 1) tc_commit(bars + 2);
    }
  }
  tc_mbar_wait(bars + 2, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: thread = row (TMEM lane); warps 0-3 take columns 0..63, warps 4-7 columns 64..127.
  // Each warp transposes its 32 x 32 block through (now idle) stage memory so that every global
  // store instruction writes 128 contiguous bytes of one C row.
  {
    const int lane = tid & 31;
    float* tbuf = reinterpret_cast<float*>(tc_smem) + warp * (32 * 33);
    float* Cz = C + (size_t)blockIdx.z * split_stride;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 64; c0 < (warp >> 2) * 64 + 64; c0 += 32) {
      if (c0 >= bn) break;                                       // warp-uniform
      uint32_t r[32];
      tc_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int n = 0; n < 32; n++) tbuf[lane * 33 + n] = __uint_as_float(r[n]);
      __syncwarp();
      const int gn = n0 + c0 + lane;
      const float bv = (bias && gn < N) ? __ldg(bias + gn) : 0.f;
#pragma unroll 4
      for (int rr = 0; rr < 32; rr++) {
        const int gm = m0 + (warp & 3) * 32 + rr;
        if (gm < M && gn < N) Cz[(size_t)gm * ldc + gn] = tbuf[rr * 33 + lane] + bv;
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tm));
}

// dacc[m * ld_dst + n] += sum_z part[z][m * N + n],  n < n_valid   (fp64 accumulation of the split-K partial tiles)
__global__ void g_reduce_splits(const float* __restrict__ part, int nz, int M, int N, int n_valid, size_t stride,
                                double* __restrict__ dacc, int ld_dst) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * n_valid) return;
  const int m = (int)(e / n_valid), n = (int)(e % n_valid);
  double s = 0.0;
  for (int z = 0; z < nz; z++) s += (double)part[(size_t)z * stride + (size_t)m * N + n];
  dacc[(size_t)m * ld_dst + n] += s;
}

}  // namespace epde
