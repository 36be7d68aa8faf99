// Jump-ahead for the device-side minibatch stream (src/data_set.py:31,67: CPython's MT19937).
//
// A sharded step draws ONE global stream: rank r needs the words [2 lo_r, 2 hi_r) of every step's 2 B words.  Walking the
// whole stream on every rank costs B_global / 1.25e9 s per step (csrc/epde_mt.cuh: the recurrence is sequential), which at
// 8 x 2^20 samples is five times the step itself.  MT19937 is linear over GF(2): with phi its characteristic polynomial
// (degree 19937) and g_J(x) = x^J mod phi = sum_i g_i x^i, the generator state J words ahead is
//     x[t + J + m] = XOR_{i : g_i = 1} x[t + i + m],   m = 0..623,
// i.e. a fixed XOR-combination of the next 20 560 words.  g_J depends only on J, so it is computed once on the host
// (epde_mt_jump_poly: Berlekamp-Massey for phi, square-and-multiply for x^J) and per step every rank
//   1. generates the 20 560-word window of the step's base state (k_mt_window, one CTA, 33 synchronisations),
//   2. applies g_{2 lo_r} -> the state at its slice and g_{2 B} -> the next step's base state (k_mt_apply, 78 CTAs each),
//   3. generates only its own 2 (hi_r - lo_r) words (k_mt_generate).
// The jumped state is the 624-word WINDOW at the target position with pos = 0: a different representation of the same stream
// position than CPython's (block, pos) pair -- random.setstate() continues the stream identically from it.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/eigenpde_b200.h"
#include "epde_mt.cuh"

namespace {

constexpr int DEG = 19937;
constexpr int PW = 312;                    // 64-bit words of a polynomial of degree < 19968
constexpr int MT_WIN_MAX = 624 + 20560 + 623;          // words of the window buffer (position t <= 624, rounded up to whole steps)

typedef std::vector<uint64_t> Poly;

inline int getb(const Poly& p, int i) { return (int)((p[i >> 6] >> (i & 63)) & 1u); }
inline void flipb(Poly& p, int i) { p[i >> 6] ^= (uint64_t)1 << (i & 63); }

// a ^= b << s   (bit shift, both little-endian bit arrays of the same word count)
void xor_shift(Poly& a, const Poly& b, int s) {
  const int ws = s >> 6, bs = s & 63, n = (int)a.size();
  for (int i = n - 1; i >= ws; i--) {
    uint64_t v = b[i - ws] << bs;
    if (bs && i - ws - 1 >= 0) v |= b[i - ws - 1] >> (64 - bs);
    a[i] ^= v;
  }
}

// characteristic polynomial of MT19937 by Berlekamp-Massey on the least significant bit of 2 * 19937 consecutive words
Poly mt_char_poly() {
  const int NB = 2 * DEG;
  std::vector<uint32_t> mt(624);
  mt[0] = 5489u;
  for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  std::vector<uint8_t> s(NB);
  int pos = 624;
  for (int n = 0; n < NB; n++) {
    if (pos >= 624) {
      auto f = [](uint32_t a, uint32_t b) { const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu); return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); };
      int kk = 0;
      for (; kk < 624 - 397; kk++) mt[kk] = mt[kk + 397] ^ f(mt[kk], mt[kk + 1]);
      for (; kk < 623; kk++) mt[kk] = mt[kk + (397 - 624)] ^ f(mt[kk], mt[kk + 1]);
      mt[623] = mt[396] ^ f(mt[623], mt[0]);
      pos = 0;
    }
    s[n] = (uint8_t)(mt[pos++] & 1u);
  }
  const int W = (NB + 64) / 64 + 1;
  Poly C(W, 0), Bp(W, 0), R(W, 0), T;
  C[0] = 1; Bp[0] = 1;
  int L = 0, m = 1;
  for (int n = 0; n < NB; n++) {
    // R: bit j = s[n - j]  (shift in the new bit)
    for (int i = W - 1; i > 0; i--) R[i] = (R[i] << 1) | (R[i - 1] >> 63);
    R[0] = (R[0] << 1) | s[n];
    uint64_t acc = 0;
    for (int i = 0; i <= (L >> 6); i++) acc ^= C[i] & R[i];
    const int d = __builtin_parityll(acc);
    if (!d) { m++; continue; }
    if (2 * L <= n) { T = C; xor_shift(C, Bp, m); L = n + 1 - L; Bp = T; m = 1; }
    else { xor_shift(C, Bp, m); m++; }
  }
  // connection polynomial C (s[n] = sum_{i>=1} c_i s[n-i]) -> characteristic polynomial phi(x) = sum_i c_i x^(L-i)
  Poly phi(PW + 1, 0);
  if (L != DEG) return Poly();
  for (int i = 0; i <= L; i++) if (getb(C, i)) flipb(phi, L - i);
  return phi;
}

// r <- r mod phi for a polynomial with bits up to `top`
void reduce(Poly& r, const Poly& phi, int top) {
  for (int dg = top; dg >= DEG; dg--)
    if (getb(r, dg)) xor_shift(r, phi, dg - DEG);
}

}  // namespace

extern "C" {

// g_J(x) = x^J mod phi as 624 uint32 words (bit i of the polynomial = bit (i % 32) of word i / 32); host only.
int epde_mt_jump_poly(int64_t J, uint32_t* h_poly624) {
  if (!h_poly624) return EPDE_ERR_NULL;
  if (J < 0) return EPDE_ERR_ARG;
  static const Poly phi = mt_char_poly();            // constant after first use (thread-safe initialisation)
  if (phi.empty()) return EPDE_ERR_UNSUPPORTED;
  const int W2 = 2 * PW + 2;
  Poly r(W2, 0);
  r[0] = 1;
  int top = 63;
  while (top > 0 && !((J >> top) & 1)) top--;
  Poly phi2(W2, 0);
  for (size_t i = 0; i < phi.size(); i++) phi2[i] = phi[i];
  for (int b = J ? top : -1; b >= 0; b--) {
    // square: spread the bits
    Poly sq(W2, 0);
    for (int i = 0; i < DEG; i++) if (getb(r, i)) flipb(sq, 2 * i);
    reduce(sq, phi2, 2 * (DEG - 1));
    r = sq;
    if ((J >> b) & 1) {                               // times x
      for (int i = W2 - 1; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
      r[0] <<= 1;
      if (getb(r, DEG)) xor_shift(r, phi2, 0);
    }
  }
  for (int i = 0; i < 624; i++) h_poly624[i] = (uint32_t)(r[i >> 1] >> ((i & 1) * 32));
  h_poly624[623] &= (1u << (DEG - 623 * 32)) - 1u;    // bits >= 19937 are zero
  return EPDE_OK;
}

}  // extern "C"

namespace epde {

// window[i] = x[i] for i < t + 20560 (t = state[624]): the block itself and the following words, 623 per synchronisation
__global__ void __launch_bounds__(MT_THREADS, 1)
k_mt_window(const uint32_t* __restrict__ state, uint32_t* __restrict__ win) {
  extern __shared__ uint32_t wbuf[];                  // MT_WIN_MAX words
  const int tid = threadIdx.x;
  const int t = (int)state[MT_N];
  for (int i = tid; i < MT_N; i += MT_THREADS) wbuf[i] = state[i];
  __syncthreads();
  const int need = t + 20560;
  for (int n = MT_N; n < need; n += MT_STEP) {
    if (tid < MT_STEP) {
      const uint32_t* q = wbuf + (n - MT_N) + tid;    // q[i] = x[m - 624 + i], m = n + tid
      uint32_t v;
      if (tid < 227)       v = q[397] ^ mt_f(q[0], q[1]);
      else if (tid < 454)  v = q[170] ^ mt_f(q[-227], q[-226]) ^ mt_f(q[0], q[1]);
      else                 v = q[-57] ^ mt_f(q[-454], q[-453]) ^ mt_f(q[-227], q[-226]) ^ mt_f(q[0], q[1]);
      wbuf[n + tid] = v;
    }
    __syncthreads();
  }
  // the part the polynomial is applied to starts at t
  for (int i = tid; i < 20560; i += MT_THREADS) win[i] = wbuf[t + i];
}

// out[m] = XOR_{i : g_i = 1} win[i + m], m = 0..623; out[624] = 0 (position).  One warp per output word; blockIdx.y = which
// polynomial / output state: all jumps of a step run in ONE launch (they are latency-bound chains of dependent loads).
struct MtOuts { uint32_t* p[32]; };
__global__ void __launch_bounds__(256)
k_mt_apply(const uint32_t* __restrict__ win, const uint32_t* __restrict__ polys, MtOuts outs) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int m = blockIdx.x * 8 + wid;                 // 78 CTAs x 8 warps = 624 output words
  const uint32_t* poly = polys + (size_t)blockIdx.y * MT_N;
  uint32_t* out = outs.p[blockIdx.y];
  uint32_t acc = 0;
  for (int w = lane; w < MT_N; w += 32) {             // lane-strided words of the polynomial
    uint32_t g = __ldg(poly + w);
    const uint32_t* base = win + 32 * w + m;
    while (g) {
      const int b = __ffs(g) - 1;
      acc ^= __ldg(base + b);
      g &= g - 1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[m] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[MT_N] = 0u;
}

}  // namespace epde

extern "C" {

// d_state_out[k] = the generator d_state_in advanced by the jump of d_polys[k] (k < n_polys), all from one window.
// d_window: scratch of 20560 uint32.  d_polys: n_polys x 624 words (epde_mt_jump_poly).
int epde_mt_jump_device(const uint32_t* d_state_in, const uint32_t* d_polys, int32_t n_polys, uint32_t* const* h_state_out,
                        uint32_t* d_window, void* stream) {
  if (!d_state_in || !d_polys || !h_state_out || !d_window) return EPDE_ERR_NULL;
  if (n_polys < 1 || n_polys > 32) return EPDE_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)MT_WIN_MAX * 4;
  cudaError_t e = cudaFuncSetAttribute(epde::k_mt_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  epde::MtOuts outs;
  for (int k = 0; k < 32; k++) outs.p[k] = nullptr;
  for (int k = 0; k < n_polys; k++) {
    if (!h_state_out[k]) return EPDE_ERR_NULL;
    outs.p[k] = h_state_out[k];
  }
  epde::k_mt_window<<<1, epde::MT_THREADS, smem, st>>>(d_state_in, d_window);
  epde::k_mt_apply<<<dim3(78, n_polys), 256, 0, st>>>(d_window, d_polys, outs);
  return (int)cudaGetLastError();
}

}  // extern "C"
