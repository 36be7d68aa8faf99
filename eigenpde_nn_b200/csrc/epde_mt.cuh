// Device-side, bit-exact minibatch index stream of the reference (src/data_set.py:31,67,77): B draws of
// min(floor(u * K), K - 1) from CPython's MT19937 (`random.choices(range(K), cum_weights = 1..K, k = B)` consumes two
// 32-bit words per draw: u = ((a >> 5) * 2^26 + (b >> 6)) / 2^53).
//
// MT19937 is the linear recurrence  x[n + 624] = x[n + 397] ^ f(x[n], x[n + 1]),  f(a, b) = ((a & U | b & L) >> 1) ^ (b & 1 ?
// MAG : 0): a word depends on the word 227 places back, so the textbook form yields 227 independent words per synchronisation.
// Substituting the recurrence into itself twice,
//     x[m] = x[m - 681] ^ f(x[m - 1078], x[m - 1077]) ^ f(x[m - 851], x[m - 850]) ^ f(x[m - 624], x[m - 623]),
// pushes the nearest dependence to 623 places, so ONE CTA produces 623 words per __syncthreads (2^21 words for B = 2^20 in
// about 0.2 ms on one SM; the step's kernels leave one SM idle, and the engine draws one step ahead on a side stream).
// k_mt_generate writes the raw words the draws consume and leaves the advanced generator state behind (the 624-word block that
// holds the last consumed word + the position in it, exactly what CPython's random.getstate() would show);
// k_mt_to_idx tempers the words and forms the indices (embarrassingly parallel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace epde {

constexpr int MT_N = 624;
constexpr int MT_STEP = 623;          // words per synchronisation
constexpr int MT_CHUNK = 8;           // steps between compactions of the shared-memory window
constexpr int MT_THREADS = 640;

__device__ __forceinline__ uint32_t mt_f(uint32_t a, uint32_t b) {
  return (((a & 0x80000000u) | (b & 0x7fffffffu)) >> 1) ^ ((b & 1u) ? 0x9908b0dfu : 0u);
}
__host__ __device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
  return y;
}

// state[0..623] = mt, state[624] = pos (0..624).  Consumes W = 2 * B words (W < 2^31) starting at position pos of the
// current block and writes them (untempered) to raw[0..W).  One CTA of MT_THREADS threads; 32-bit index arithmetic and
// warp-uniform branches only (the CTA is issue-bound: 20 warps on one SM).
// gridDim.x > 1: CTA q continues the generator state + 625 q (a state jumped ahead to word q Wsub of the slice,
// epde_mt_jump_device) for the words [q Wsub, min((q + 1) Wsub, Wtot)) - the sequential recurrence runs on several SMs at once.
static __global__ void __launch_bounds__(MT_THREADS, 1)
k_mt_generate(uint32_t* __restrict__ state, int Wsub, uint32_t* __restrict__ raw, int Wtot) {
  __shared__ uint32_t buf[MT_N + MT_STEP * MT_CHUNK];
  const int tid = threadIdx.x;
  state += (size_t)blockIdx.x * (MT_N + 1);
  raw += (size_t)blockIdx.x * Wsub;
  const int W = min(Wsub, Wtot - (int)blockIdx.x * Wsub);
  const int p0 = (int)state[MT_N];
  for (int i = tid; i < MT_N; i += MT_THREADS) buf[i] = state[i];
  __syncthreads();
  if (W <= 0) return;
  const long long last = (long long)p0 + W - 1;       // index (relative to the current block) of the last consumed word
  const long long blk_f = last / MT_N;                // block that becomes the new state
  // words of the current block that the draws consume
  for (int x = p0 + tid; x < MT_N && x <= last; x += MT_THREADS) raw[x - p0] = buf[x];
  if (blk_f > 0) {
    const long long nsteps = (blk_f * MT_N + MT_N - MT_N + MT_STEP - 1) / MT_STEP;   // steps until the frontier passes the end of block blk_f
    // per-thread running indices (all fit 32 bits: W < 2^31): r = index into raw of the word this thread generates next,
    // q = its index inside the final block (valid when 0 <= q < 624)
    int r = MT_N + tid - p0;
    long long qq = (long long)MT_N + tid - blk_f * MT_N;
    const bool act = tid < MT_STEP;
    for (long long s0 = 0; s0 < nsteps; s0 += MT_CHUNK) {
      const int ns = (int)((nsteps - s0) < MT_CHUNK ? (nsteps - s0) : MT_CHUNK);
      const bool tail = qq + (long long)MT_STEP * ns + MT_STEP > 0;      // warp-uniform enough: this chunk may touch the final block
      for (int s = 0; s < ns; s++) {
        if (act) {
          const uint32_t* q = buf + MT_STEP * s + tid;     // q[i] = x[m - 624 + i], m = the word this thread generates
          uint32_t v;
          if (tid < 227)       v = q[397] ^ mt_f(q[0], q[1]);
          else if (tid < 454)  v = q[170] ^ mt_f(q[-227], q[-226]) ^ mt_f(q[0], q[1]);
          else                 v = q[-57] ^ mt_f(q[-454], q[-453]) ^ mt_f(q[-227], q[-226]) ^ mt_f(q[0], q[1]);
          buf[MT_N + MT_STEP * s + tid] = v;
          if ((unsigned)r < (unsigned)W) raw[r] = v;
          if (tail && qq >= 0 && qq < MT_N) state[qq] = v;
        }
        r += MT_STEP; qq += MT_STEP;
        __syncthreads();
      }
      // compaction: keep the 624 words below the frontier, buf[623 ns .. 623 ns + 624)
      uint32_t keep = 0;
      if (tid < MT_N) keep = buf[MT_STEP * ns + tid];
      __syncthreads();
      if (tid < MT_N) buf[tid] = keep;
      __syncthreads();
    }
  }
  if (tid == 0) state[MT_N] = (uint32_t)(last + 1 - blk_f * MT_N);
}

// idx[i] = min(floor(u * K), K - 1) for draws out_lo + i, i < out_n, from the raw words of k_mt_generate
static __global__ void __launch_bounds__(256)
k_mt_to_idx(const uint32_t* __restrict__ raw, int64_t K, int64_t out_lo, int64_t out_n, int64_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_n) return;
  const uint2 w = *reinterpret_cast<const uint2*>(raw + 2 * (out_lo + i));
  const uint32_t a = mt_temper(w.x) >> 5, b = mt_temper(w.y) >> 6;
  const double u = ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
  int64_t j = (int64_t)floor(u * (double)K);          // bisect over cum_weights = 1..K (src/data_set.py:31,67)
  if (j > K - 1) j = K - 1;
  idx[i] = j;
}

}  // namespace epde
