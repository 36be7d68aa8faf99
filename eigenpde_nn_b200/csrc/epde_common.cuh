// Shared device helpers for the EigenPDE-NN step kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace epde {

// ---- packed FP32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2) ---------------------------
// fma.rn.f32x2 issues ONE instruction for two FP32 FMAs; ptxas folds a {w,w} splat operand
// into the scalar-broadcast form (FFMA2 Rd, Rw.F32, Rx.F32x2, Rc.F32x2), which is how the
// broadcast weights enter the per-sample mat-vecs below.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 splat(float w) { return make_float2(w, w); }

// ---- tanh ----------------------------------------------------------------------------------
// EPDE_FAST_TANH: 1 - 2/(exp2(2x log2 e) + 1) with MUFU.EX2 + MUFU.RCP (abs. error ~2e-7);
// otherwise CUDA's tanhf (<= 2 ulp).  Both are MUFU-bound (2 MUFU per value).
__device__ __forceinline__ float tanh1(float x) {
#ifdef EPDE_FAST_TANH
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
#else
  return tanhf(x);
#endif
}
__device__ __forceinline__ float2 tanh2(float2 z) { return make_float2(tanh1(z.x), tanh1(z.y)); }

// y[o] += sum_i W[i][o] * x[i] for a sample pair; W in shared memory, row i contiguous over o
// (16-byte aligned rows).  One broadcast LDS.128 feeds four FFMA2.
template <int NIN, int NOUT>
__device__ __forceinline__ void mv_acc(const float* __restrict__ sW, const float2 (&x)[NIN], float2 (&y)[NOUT]) {
  static_assert(NOUT % 4 == 0, "row length must be a multiple of 4");
#pragma unroll
  for (int i = 0; i < NIN; i++) {
#pragma unroll
    for (int q = 0; q < NOUT / 4; q++) {
      const float4 w = *reinterpret_cast<const float4*>(sW + i * NOUT + 4 * q);
      y[4 * q + 0] = ffma2(splat(w.x), x[i], y[4 * q + 0]);
      y[4 * q + 1] = ffma2(splat(w.y), x[i], y[4 * q + 1]);
      y[4 * q + 2] = ffma2(splat(w.z), x[i], y[4 * q + 2]);
      y[4 * q + 3] = ffma2(splat(w.w), x[i], y[4 * q + 3]);
    }
  }
}

template <int N>
__device__ __forceinline__ void zero2(float2 (&y)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++) y[i] = make_float2(0.f, 0.f);
}

// y[o] = {b[o], b[o]} from a shared-memory bias vector
template <int N>
__device__ __forceinline__ void bias2(const float* __restrict__ sb, float2 (&y)[N]) {
#pragma unroll
  for (int q = 0; q < N / 4; q++) {
    const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);
    y[4 * q + 0] = splat(b.x); y[4 * q + 1] = splat(b.y);
    y[4 * q + 2] = splat(b.z); y[4 * q + 3] = splat(b.w);
  }
}

// Global loads through volatile asm: issued exactly where written (the compiler otherwise sinks
// them next to their first use, which exposes the full memory latency with one warp per scheduler).
// (64-bit register outputs so that the value stays an aligned register pair for FFMA2.)
__device__ __forceinline__ float2 ldg_early2(const float2* p) {
  unsigned long long v;
  asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return *reinterpret_cast<float2*>(&v);
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// A kernel launched with launch_pdl() may START while its predecessor in the stream is still running - as soon as every
// CTA of the predecessor has executed pdl_trigger() (or exited) - and must execute pdl_wait() before it touches anything
// the predecessor writes: pdl_wait() returns when the predecessor grid has completed and its writes are visible.
// Used where a kernel has a prologue that does not depend on its predecessor (shared-memory / tensor-memory set-up) or
// is independent of it altogether; the chain stays inside one stream and is capturable into a CUDA graph.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- one-shot all-reduce over NVLink peer memory, executed by ONE CTA (all of its threads) ----------------------------
// Every rank owns one symmetric buffer [256 B of flags = 2 parities x 32 ranks x uint32 | 2 parities x world slots x
// slot_bytes]; world <= 32.  A call (epoch e = *epoch_ctr + 1, parity e & 1) stores this rank's payload into slot[rank] of
// EVERY rank's buffer with plain peer stores, fences, raises flag[parity][rank] = e in every rank's buffer, waits until all
// world flags of its OWN buffer show e, and sums the world slots in rank order (the same order on every rank: bit-identical
// results everywhere).  A rank cannot run two epochs ahead of a peer (it needs that peer's flag for e + 1, raised only
// after the peer finished e), so two parities suffice.  The epoch counter lives in device memory and is advanced by the
// call: launch arguments never change, the exchange can sit inside a replayed CUDA graph - or inside another kernel.
struct Peer {
  char* const* bufs;        // device array of `world` pointers (NULL: single rank, no exchange)
  int rank, world;
  uint32_t* epoch;          // device counter
  long long slot_bytes;
};
template <typename T>
__device__ __forceinline__ void peer_allreduce_block(const Peer& P, uint32_t epoch, const T* src, T* dst, T* dst2, int n) {
  const int par = epoch & 1;
  const long long data_off = 256 + ((long long)par * P.world + P.rank) * P.slot_bytes;
  for (int p = 0; p < P.world; p++) {
    T* slot = reinterpret_cast<T*>(P.bufs[p] + data_off);
    for (int i = threadIdx.x; i < n; i += blockDim.x) slot[i] = *reinterpret_cast<const volatile T*>(src + i);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < P.world) {
    volatile uint32_t* theirs = reinterpret_cast<volatile uint32_t*>(P.bufs[threadIdx.x]) + par * P.world + P.rank;
    *theirs = epoch;
    __threadfence_system();
    volatile uint32_t* mine = reinterpret_cast<volatile uint32_t*>(P.bufs[P.rank]) + par * P.world + threadIdx.x;
    int spin = 0;
    while (*mine != epoch) { __nanosleep(128); if (++spin > (1 << 28)) __trap(); }   // bounded (~1 min): a lost peer traps
  }
  __syncthreads();
  __threadfence_system();
  const char* base = P.bufs[P.rank] + 256 + (long long)par * P.world * P.slot_bytes;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    T v = 0;
    for (int p = 0; p < P.world; p++) v += reinterpret_cast<const volatile T*>(base + (long long)p * P.slot_bytes)[i];
    dst[i] = v;
    if (dst2) dst2[i] = v;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace epde
