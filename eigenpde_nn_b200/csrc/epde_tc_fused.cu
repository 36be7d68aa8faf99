// Launchers of the tensor-core fused family (separate translation unit: these kernels are large).
#include <stdlib.h>
#include "epde_tc_api.h"
#include "epde_tc_fused.cuh"
#include "epde_tc_pass2.cuh"

#ifdef EPDE_PHASE_TIMING
extern "C" int epde_debug_tc_cycles(unsigned long long* h16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(h16, epde::tc::g_tc_cycles, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset) { unsigned long long z[16] = {0}; e = cudaMemcpyToSymbol(epde::tc::g_tc_cycles, z, sizeof(z)); }
  return (int)e;
}
#endif

namespace epde {
namespace tc {

size_t pack_floats(int d) { return (size_t)TcPack(d).total(); }
size_t stash_floats(int64_t B, int k) { return (size_t)((B + 127) / 128) * (size_t)k * ST_TILE; }

int pack(int d, int k, const float* params, const float* diag, float* pwt, float* nrm, cudaStream_t st) {
  k_pack_tc<<<dim3(k, 6), 256, 0, st>>>(params, diag, d, pwt, nrm);
  return (int)cudaGetLastError();
}

int partial(int d, int d_pad, int k, int nsm, int64_t B, const float* params, const float* diag, const float* xt,
            const float* wt, float* pwt, float* ybuf, double* part, double* part2, double* sums, float* mx, float* nrm,
            float* stash, cudaStream_t st, int defer_reduce, int* gx_out, int* g2_out) {
  const TcPack TL(d);
  // (the weight images were packed by pack(), launched BEFORE the gather so that the two overlap; packing inside the gather's
  //  launch - extra CTAs scheduled first - was measured: the merged kernel takes 141 us against 115 + 15 us for two launches)
  const int64_t nt128 = (B + 127) / 128;
  int gx = nsm / k; if (gx < 1) gx = 1;
  if ((int64_t)gx * TG > nt128) gx = (int)((nt128 + TG - 1) / TG);
  const size_t smem = ((size_t)TL.total() + (size_t)TG * TXR * 128) * 4 + 64;
  cudaError_t e = cudaSuccess;
#define EPDE_P1_LAUNCH(...)                                                                                 \
  do {                                                                                                     \
    e = cudaFuncSetAttribute(k_pass1_tc<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return (int)e;                                                                   \
    k_pass1_tc<__VA_ARGS__><<<dim3(gx, k), TTHREADS, smem, st>>>(xt, wt, B, d, d_pad, k, pwt, ybuf, part, mx, stash); \
  } while (0)
  const bool std_pad = d_pad == (d + 3) / 4 * 4;
  if (std_pad && d == 50) EPDE_P1_LAUNCH(7, 50);          // the BASELINE shapes: exact d at compile time
  else if (std_pad && d == 30) EPDE_P1_LAUNCH(4, 30);
  else if (std_pad && d == 2) EPDE_P1_LAUNCH(1, 2);
  else switch (TL.ks1) {   // any other d <= 63: one instantiation per layer-1 k-step count
    case 1: EPDE_P1_LAUNCH(1); break;  case 2: EPDE_P1_LAUNCH(2); break;  case 3: EPDE_P1_LAUNCH(3); break;
    case 4: EPDE_P1_LAUNCH(4); break;  case 5: EPDE_P1_LAUNCH(5); break;  case 6: EPDE_P1_LAUNCH(6); break;
    case 7: EPDE_P1_LAUNCH(7); break;  default: EPDE_P1_LAUNCH(8); break;
  }
#undef EPDE_P1_LAUNCH
  e = cudaGetLastError(); if (e != cudaSuccess) return (int)e;
  int g2 = (int)((B + SY_THREADS - 1) / SY_THREADS); if (g2 > nsm) g2 = nsm;
  switch (k) {
    case 1: k_sums_y<1><<<g2, SY_THREADS, 0, st>>>(wt, ybuf, B, part2); break;
    case 2: k_sums_y<2><<<g2, SY_THREADS, 0, st>>>(wt, ybuf, B, part2); break;
    case 3: k_sums_y<3><<<g2, SY_THREADS, 0, st>>>(wt, ybuf, B, part2); break;
    case 4: k_sums_y<4><<<g2, SY_THREADS, 0, st>>>(wt, ybuf, B, part2); break;
    case 5: k_sums_y<5><<<g2, SY_THREADS, 0, st>>>(wt, ybuf, B, part2); break;
    default: k_sums_y<6><<<g2, SY_THREADS, 0, st>>>(wt, ybuf, B, part2); break;
  }
  if (gx_out) *gx_out = gx;
  if (g2_out) *g2_out = g2;
  if (!defer_reduce) k_reduce_tc<<<1, 1024, 0, st>>>(part, gx, part2, g2, k, sums);      // one warp per sum (34 sums for k = 6)
  return (int)cudaGetLastError();
}

int reduce_coef(const double* part, int gx, const double* part2, int g2, int k, double* sums, double beta, double alpha,
                const EigW& ew, int sort, double* out, Coef* coef, const float* mx, const float* nrm, cudaStream_t st,
                const Peer* peer) {
  const Peer none = {nullptr, 0, 1, nullptr, 0};
  cudaError_t e = launch_pdl(k_reduce_coef_tc, dim3(1), dim3(1024), 0, st, part, gx, part2, g2, k, sums, beta, alpha, ew, sort,
                             out, coef, mx, nrm, peer ? *peer : none);
  return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

int finish(int d, int d_pad, int k, int nsm, int64_t B, const float* xt, const float* wt, const float* pwt,
           const float* ybuf, const Coef* coef, const float* stash, float* pg, int* gx_out, cudaStream_t st) {
  const TcPack TL(d);
  const int64_t nt128 = (B + 127) / 128;
  int gx = nsm / k; if (gx < 1) gx = 1;
  if ((int64_t)gx * P2G > nt128) gx = (int)((nt128 + P2G - 1) / P2G);
  const int gn = GradLayout<TH>(d_pad).total();
  const size_t smem = (size_t)P2G * SG_BYTES + (size_t)TL.total() * 4 + (size_t)P2G * im_total(TL.K1()) * 4 + 768;
  cudaError_t e = cudaSuccess;
#define EPDE_P2_LAUNCH(...)                                                                                          \
  do {                                                                                                              \
    e = cudaFuncSetAttribute(k_pass2_tc<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
    if (e != cudaSuccess) return (int)e;                                                                            \
    e = launch_pdl(k_pass2_tc<__VA_ARGS__>, dim3(gx, k), dim3(P2THREADS), smem, st, xt, wt, B, d, d_pad, k, pwt, ybuf, coef, stash, pg); \
    if (e != cudaSuccess) return (int)e;                                                                            \
  } while (0)
  const bool std_pad = d_pad == (d + 3) / 4 * 4;
  if (std_pad && d == 50) EPDE_P2_LAUNCH(7, 50);
  else if (std_pad && d == 30) EPDE_P2_LAUNCH(4, 30);
  else if (std_pad && d == 2) EPDE_P2_LAUNCH(1, 2);
  else switch (TL.ks1) {
    case 1: EPDE_P2_LAUNCH(1); break;  case 2: EPDE_P2_LAUNCH(2); break;  case 3: EPDE_P2_LAUNCH(3); break;
    case 4: EPDE_P2_LAUNCH(4); break;  case 5: EPDE_P2_LAUNCH(5); break;  case 6: EPDE_P2_LAUNCH(6); break;
    case 7: EPDE_P2_LAUNCH(7); break;  default: EPDE_P2_LAUNCH(8); break;
  }
#undef EPDE_P2_LAUNCH
  *gx_out = gx;
  return (int)cudaGetLastError();
}

}  // namespace tc
}  // namespace epde
