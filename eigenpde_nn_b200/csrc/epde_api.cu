// C ABI of libeigenpde_b200.so (see include/eigenpde_b200.h) and the small glue kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/eigenpde_b200.h"
#include "epde_coef.cuh"
#include "epde_fused.cuh"
#include "epde_generic.cuh"
#include "epde_tc_api.h"
#include "epde_md_align.cuh"
#include "epde_mt.cuh"

using namespace epde;

namespace {

constexpr int kMaxCtas = 160;          // upper bound on persistent CTAs (B200: 148 SMs)
constexpr int kMxWords = 2 + kMxPerNet * EPDE_MAX_K;      // maxima (floats), followed by one ticket word
constexpr int kMetricCtas = 16 * kMaxCtas;  // upper bound on k_metric's grid
constexpr int kFusedH = 20;
// kernel family of the fused path: EPDE_FLAG_FFMA2 in the descriptor selects the FFMA2 kernels (epde_fused.cuh), the
// default are the tcgen05 kernels (epde_tc_fused.cuh).  The choice travels with the descriptor: no process-wide state.
// A per-row metric (MD alignment layer) is implemented by the FFMA2 kernels only.
// ... and so are the activations other than Tanh (Act<true> in epde_fused.cuh).
inline int fused_impl(const epde_desc* D) {
  return ((D->flags & EPDE_FLAG_FFMA2) || D->metric || D->act_id != EPDE_ACT_TANH) ? 0 : 1;
}
inline bool general_act(const epde_desc* D) { return D->act_id != EPDE_ACT_TANH; }
constexpr size_t kSumsTail = 2048;
constexpr int64_t kFwdSlab = 1 << 18;   // samples per slab of epde_forward (bounds its workspace)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int check_desc(const epde_desc* D) {
  if (!D) return EPDE_ERR_NULL;
  if (D->k < 1 || D->k > EPDE_MAX_K) return EPDE_ERR_DESC;
  if (D->n_layers < 1 || D->n_layers > EPDE_MAX_LAYERS) return EPDE_ERR_DESC;
  if (D->d < 1 || D->widths[0] != D->d) return EPDE_ERR_DESC;
  if (D->d_pad < D->d || (D->d_pad & 3)) return EPDE_ERR_ALIGN;
  for (int l = 1; l <= D->n_layers; l++)
    if (D->widths[l] < 1 || D->widths[l] > 4096) return EPDE_ERR_DESC;
  if (D->widths[D->n_layers] != 1) return EPDE_ERR_DESC;
  if (D->act_id < EPDE_ACT_TANH || D->act_id > EPDE_ACT_TANHSHRINK) return EPDE_ERR_DESC;
  return EPDE_OK;
}

bool fused_ok(const epde_desc* D) {
  if (D->metric && (D->d > 2 * kFusedH || (D->flags & EPDE_FLAG_GENERIC))) return false;   // per-row metric: u0 / p live in 2 H shared-memory rows
  if (D->flags & EPDE_FLAG_GENERIC) return false;
  if (D->n_layers != 4) return false;
  if (D->act_id != EPDE_ACT_TANH) {      // the FFMA2 kernels take every activation whose derivatives are functions of its value
    // (not Tanhshrink: its derivatives are not functions of its value; not Sigmoid: all-positive units make the gradient a
    //  small difference of large sums, and the layer-wise family - fp64 accumulation of every weight gradient - is the
    //  one that holds the reference's golden Sigmoid case to 2e-4)
    if (D->act_id == EPDE_ACT_TANHSHRINK || D->act_id == EPDE_ACT_SIGMOID || D->metric) return false;
  }
  if (D->widths[1] != kFusedH || D->widths[2] != kFusedH || D->widths[3] != kFusedH) return false;
  if (D->k > 6) return false;
  if (D->d + 1 > 8 * 8) return false;           // tensor-core family: layer-1 k-steps 8 ks1 >= d + 1 with ks1 <= 8 (x staging rows TXR = 64)
  return pass2_smem_bytes<kFusedH>(D->d, D->d_pad) <= 227 * 1024 &&
         (size_t)D->k * PackLayout<kFusedH>(D->d_pad).total() * 4 + 4096 <= 200 * 1024;
}

int64_t param_count(const epde_desc* D) {
  int64_t n = 0;
  for (int l = 1; l <= D->n_layers; l++) n += (int64_t)D->widths[l] * D->widths[l - 1] + D->widths[l];
  return n * D->k;
}

int num_sums(int k) { return 1 + 2 * k + k * (k + 1) / 2; }

// ---- workspace ------------------------------------------------------------------------------
struct FusedWs {
  float* pw; float* ybuf; double* part; Coef* coef; float* pg; float* xt; float* wt; float* pwt; double* part2;
  float* mx; float* nrm;      // bounds for the fp16 gradient rounds (epde_coef.cuh); mx is followed by k_fin2's ticket word
  float* stash;               // phase-1 intermediates of the tensor-core family (400 B per sample and net)
  float* U; double* epart;    // per-row metric: u0 -> p of every net [k][tiles][d][256], k_metric's partial energy sums
  size_t total;
};
// slot: the tensor-core family keeps TWO sets of gathered rows (xt, wt) and maxima (mx), so that the gather of the next
// step's minibatch (epde_gather_ahead, side stream) can fill one while the current step reads the other.
FusedWs fused_ws(const epde_desc* D, int64_t B, void* base, int slot = 0) {
  char* p = (char*)base; size_t off = 0; FusedWs W;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off = align_up(off + bytes, 256); return r; };
  W.pw = (float*)take((size_t)D->k * PackLayout<kFusedH>(D->d_pad).total() * 4);
  W.ybuf = (float*)take((size_t)B * D->k * 4);
  W.part = (double*)take((size_t)kMaxCtas * num_sums(D->k) * 8);
  W.coef = (Coef*)take(sizeof(Coef));
  W.pg = (float*)take((size_t)kMaxCtas * GradLayout<kFusedH>(D->d_pad).total() * 4);
  const size_t ntl = (size_t)((B + GT - 1) / GT);
  W.xt = (float*)take(ntl * D->d_pad * GT * 4);
  W.wt = (float*)take(ntl * GT * 4);
  W.pwt = (float*)take((size_t)D->k * tc::pack_floats(D->d) * 4);
  W.part2 = (double*)take((size_t)kMaxCtas * 32 * 8);
  W.mx = (float*)take((size_t)(kMxWords + 1) * 4);
  W.nrm = (float*)take((size_t)kNrmPerNet * EPDE_MAX_K * 4);
  W.stash = (float*)take(D->metric ? 0 : tc::stash_floats(B, D->k) * 4);      // (a metric runs the FFMA2 kernels: no stash)
  {
    const bool two = !D->metric;      // (both kernel families: the workspace layout does not depend on EPDE_FLAG_FFMA2)
    float* xt1 = (float*)take(two ? ntl * D->d_pad * GT * 4 : 0);
    float* wt1 = (float*)take(two ? ntl * GT * 4 : 0);
    float* mx1 = (float*)take(two ? (size_t)(kMxWords + 1) * 4 : 0);
    if (slot == 1 && two) { W.xt = xt1; W.wt = wt1; W.mx = mx1; }
  }
  W.U = (float*)take(D->metric ? (size_t)D->k * ntl * D->d * GT * 4 : 0);
  W.epart = (double*)take(D->metric ? (size_t)kMetricCtas * EPDE_MAX_K * 8 : 0);
  W.total = off;
  return W;
}

int sm_count(int* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return (int)e;
  if (*out > kMaxCtas) *out = kMaxCtas;
  return 0;
}

// sum the per-CTA partials in CTA order -> sums[NS]
// (+ per-row metric: the energy sums E_i = sums[1 + k + i] come from k_metric's partials epart[ne][k])
// One warp per sum: the lanes stride over the partial rows, a fixed shuffle tree adds the lane sums.
__global__ void __launch_bounds__(1024)
k_reduce_part(const double* __restrict__ part, int nb, int ns, double* __restrict__ sums,
              const double* __restrict__ epart = nullptr, int ne = 0, int k = 0) {
  const int lane = threadIdx.x & 31;
  for (int s = threadIdx.x >> 5; s < ns; s += blockDim.x >> 5) {
    double v = 0.0;
    for (int b = lane; b < nb; b += 32) v += part[(size_t)b * ns + s];
    if (epart && s > k && s <= 2 * k)
      for (int b = lane; b < ne; b += 32) v += epart[(size_t)b * k + (s - 1 - k)];
    v = warp_sum(v);
    if (lane == 0) sums[s] = v;
  }
}

template <int K>
cudaError_t launch_pass1(int grid, size_t smem, cudaStream_t st, const float* xt, const float* wt, int64_t B, int d,
                         int d_pad, const float* pw, float* ybuf, double* part, float* U, int act_id) {
  if (act_id != EPDE_ACT_TANH) {       // general activation (never together with a per-row metric: fused_ok)
    cudaError_t e = cudaFuncSetAttribute(k_pass1<kFusedH, K, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_pass1<kFusedH, K, false, true><<<grid, P1_THREADS, smem, st>>>(xt, wt, B, d, d_pad, pw, ybuf, part, nullptr, act_id);
    return cudaGetLastError();
  }
  if (U) {
    cudaError_t e = cudaFuncSetAttribute(k_pass1<kFusedH, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_pass1<kFusedH, K, true><<<grid, P1_THREADS, smem, st>>>(xt, wt, B, d, d_pad, pw, ybuf, part, U);
    return cudaGetLastError();
  }
  cudaError_t e = cudaFuncSetAttribute(k_pass1<kFusedH, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_pass1<kFusedH, K><<<grid, P1_THREADS, smem, st>>>(xt, wt, B, d, d_pad, pw, ybuf, part);
  return cudaGetLastError();
}

// gather + transpose the minibatch rows into the tile-major copy (workspace xt / wt)
cudaError_t launch_gather(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
                          float* xt, float* wt, cudaStream_t st, int tile128 = 0, float* mx = nullptr) {
  if (tile128) {     // tensor-core family: 128-sample tiles, wt padded to whole 256-sample tiles by the workspace layout
    const size_t sm1 = (size_t)128 * (D->d_pad + 1) * 4;
    cudaError_t e1 = cudaFuncSetAttribute(k_gather128<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    if (e1 != cudaSuccess) return e1;
    // behind k_pack_tc, which it does not depend on: PDL lets the two run side by side (the gather waits for the packing
    // only at its very end, so that "gather complete" implies "packing complete" for phase 1)
    e1 = launch_pdl(k_gather128<0>, dim3((unsigned)((B + 127) / 128)), dim3(256), sm1, st, X, w, idx, B, D->d_pad, xt, wt, mx);
    return e1 != cudaSuccess ? e1 : cudaGetLastError();
  }
  const size_t smem = (size_t)GT * (D->d_pad + 1) * 4;
  cudaError_t e = cudaFuncSetAttribute(k_gather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_gather<0><<<(unsigned)((B + GT - 1) / GT), 256, smem, st>>>(X, w, idx, B, D->d_pad, xt, wt, tile128);
  return cudaGetLastError();
}

struct Deferred { int on, gx, g2; };      // epde_step: phase 1's final reduction is merged into the coefficient launch
// Gather look-ahead (tensor-core family): slot = which set of gathered rows the step uses; pregathered = that set was filled
// by epde_gather_ahead (the step skips its own gather); reserve = SMs the two persistent pass kernels leave free for the
// gather of the NEXT step, which runs beside them on a second stream.
struct Ahead { int slot, pregathered, reserve; };

int gather_ahead(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B, void* ws, int slot,
                 cudaStream_t st) {
  FusedWs W = fused_ws(D, B, ws, slot);
  cudaError_t e = cudaMemsetAsync(W.mx, 0, (size_t)(kMxWords + 1) * 4, st);
  if (e != cudaSuccess) return (int)e;
  return (int)launch_gather(D, X, w, idx, B, W.xt, W.wt, st, 1, W.mx);
}

int fused_partial(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
                  const float* params, const float* diag, void* ws, double* sums, cudaStream_t st, Deferred* df = nullptr,
                  const Ahead* ah = nullptr) {
  FusedWs W = fused_ws(D, B, ws, ah ? ah->slot : 0);
  int nsm; int rc = sm_count(&nsm); if (rc) return rc;
  const int impl = fused_impl(D);
  const bool pre = impl == 1 && ah && ah->pregathered;
  cudaError_t e;
  if (impl != 1) k_pack<kFusedH><<<D->k, 256, 0, st>>>(params, diag, D->d, D->d_pad, W.pw);
  else {
    if (!pre) {
      e = cudaMemsetAsync(W.mx, 0, (size_t)(kMxWords + 1) * 4, st);     // maxima are gathered with atomicMax; + k_fin2's ticket
      if (e != cudaSuccess) return (int)e;
    }
    rc = tc::pack(D->d, D->k, params, diag, W.pwt, W.nrm, st);
    if (rc) return rc;
  }
  if (!pre) {
    e = launch_gather(D, X, w, idx, B, W.xt, W.wt, st, impl == 1, impl == 1 ? W.mx : nullptr);
    if (e != cudaSuccess) return (int)e;
  }
  if (impl == 1) {
    int nsm1 = nsm - (ah ? ah->reserve : 0); if (nsm1 < D->k) nsm1 = D->k;
    return tc::partial(D->d, D->d_pad, D->k, nsm1, B, params, diag, W.xt, W.wt, W.pwt, W.ybuf, W.part, W.part2, sums,
                       W.mx, W.nrm, W.stash, st, df ? 1 : 0, df ? &df->gx : nullptr, df ? &df->g2 : nullptr);
  }
  const int64_t npairs = (B + 1) / 2;
  int grid = (int)((npairs + P1_THREADS - 1) / P1_THREADS);
  if (grid > nsm) grid = nsm;
  const size_t smem = (size_t)D->k * PackLayout<kFusedH>(D->d_pad).total() * 4;
  float* U = D->metric ? W.U : nullptr;
  switch (D->k) {
    case 1: e = launch_pass1<1>(grid, smem, st, W.xt, W.wt, B, D->d, D->d_pad, W.pw, W.ybuf, W.part, U, D->act_id); break;
    case 2: e = launch_pass1<2>(grid, smem, st, W.xt, W.wt, B, D->d, D->d_pad, W.pw, W.ybuf, W.part, U, D->act_id); break;
    case 3: e = launch_pass1<3>(grid, smem, st, W.xt, W.wt, B, D->d, D->d_pad, W.pw, W.ybuf, W.part, U, D->act_id); break;
    case 4: e = launch_pass1<4>(grid, smem, st, W.xt, W.wt, B, D->d, D->d_pad, W.pw, W.ybuf, W.part, U, D->act_id); break;
    case 5: e = launch_pass1<5>(grid, smem, st, W.xt, W.wt, B, D->d, D->d_pad, W.pw, W.ybuf, W.part, U, D->act_id); break;
    default: e = launch_pass1<6>(grid, smem, st, W.xt, W.wt, B, D->d, D->d_pad, W.pw, W.ybuf, W.part, U, D->act_id); break;
  }
  if (e != cudaSuccess) return (int)e;
  int ne = 0;
  if (D->metric) {      // p = M_r u0 in place in W.U, partial energy sums: one warp per sample over the whole GPU
    ne = (int)((B + MET_THREADS / 32 - 1) / (MET_THREADS / 32));
    if (ne > 16 * nsm) ne = 16 * nsm;
    k_metric<0><<<ne, MET_THREADS, 0, st>>>(D->metric, idx, W.wt, B, D->d, D->k, W.U, (int64_t)((B + GT - 1) / GT), W.epart);
    e = cudaGetLastError(); if (e != cudaSuccess) return (int)e;
  }
  k_reduce_part<<<1, 1024, 0, st>>>(W.part, grid, num_sums(D->k), sums, D->metric ? W.epart : nullptr, ne, D->k);
  return (int)cudaGetLastError();
}

int fused_finish(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
                 const float* params, const float* diag, double beta, double alpha, const EigW& ew,
                 const double* sums, void* ws, double* out, float* grad, cudaStream_t st, const Deferred* df = nullptr,
                 const Peer* peer = nullptr, const Ahead* ah = nullptr) {
  const Peer none = {nullptr, 0, 1, nullptr, 0};
  FusedWs W = fused_ws(D, B, ws, ah ? ah->slot : 0);
  int nsm; int rc = sm_count(&nsm); if (rc) return rc;
  const int impl = fused_impl(D);
  if (impl == 1 && df && df->on) {
    rc = tc::reduce_coef(W.part, df->gx, W.part2, df->g2, D->k, const_cast<double*>(sums), beta, alpha, ew,
                         (D->flags & EPDE_FLAG_SORT) ? 1 : 0, out, W.coef, W.mx, W.nrm, st, peer);
    if (rc) return rc;
  } else
  k_coef<0><<<1, 32, 0, st>>>(sums, D->k, beta, alpha, ew, (D->flags & EPDE_FLAG_SORT) ? 1 : 0, out, W.coef,
                              impl == 1 ? W.mx : nullptr, impl == 1 ? W.nrm : nullptr);
  if (impl == 1) {
    int gx = 0;
    int nsm2 = nsm - (ah ? ah->reserve : 0); if (nsm2 < D->k) nsm2 = D->k;
    rc = tc::finish(D->d, D->d_pad, D->k, nsm2, B, W.xt, W.wt, W.pwt, W.ybuf, W.coef, W.stash, W.pg, &gx, st);
    if (rc) return rc;
    cudaError_t e2 = launch_pdl(k_fin2<kFusedH>, dim3(D->k, fin2_items<kFusedH>(D->d)), dim3(1024), 0, st, (const float*)W.pg, gx,
                                D->k, D->d, D->d_pad, params, diag, grad, peer ? *peer : none,
                                reinterpret_cast<uint32_t*>(W.mx) + kMxWords);
    return (int)(e2 != cudaSuccess ? e2 : cudaGetLastError());
  }
  const int64_t ntiles = (B + P2_T - 1) / P2_T;
  int gx = nsm / D->k; if (gx < 1) gx = 1;
  if (gx > ntiles) gx = (int)ntiles;
  const size_t smem = pass2_smem_bytes<kFusedH>(D->d, D->d_pad);
  cudaError_t e;
  if (general_act(D)) {
    e = cudaFuncSetAttribute(k_pass2<kFusedH, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_pass2<kFusedH, false, true><<<dim3(gx, D->k), P2_THREADS, smem, st>>>(W.xt, W.wt, B, D->d, D->d_pad, D->k, W.pw, W.ybuf,
                                                                             W.coef, W.pg, nullptr, D->act_id);
  } else if (D->metric) {
    e = cudaFuncSetAttribute(k_pass2<kFusedH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_pass2<kFusedH, true><<<dim3(gx, D->k), P2_THREADS, smem, st>>>(W.xt, W.wt, B, D->d, D->d_pad, D->k, W.pw, W.ybuf,
                                                                      W.coef, W.pg, W.U);
  } else {
    e = cudaFuncSetAttribute(k_pass2<kFusedH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_pass2<kFusedH><<<dim3(gx, D->k), P2_THREADS, smem, st>>>(W.xt, W.wt, B, D->d, D->d_pad, D->k, W.pw, W.ybuf,
                                                                W.coef, W.pg);
  }
  e = cudaGetLastError(); if (e != cudaSuccess) return (int)e;
  k_fin2<kFusedH><<<dim3(D->k, fin2_items<kFusedH>(D->d)), 1024, 0, st>>>(W.pg, gx, D->k, D->d, D->d_pad, params, diag, grad, none,
                                                                           nullptr);
  return (int)cudaGetLastError();
}

int check_common(const epde_desc* D, const float* X, const float* w, int64_t B, const float* params,
                 const float* diag, void* ws, size_t ws_bytes) {
  int rc = check_desc(D); if (rc) return rc;
  if (!X || !w || !params || !diag || !ws) return EPDE_ERR_NULL;
  if (B < 1) return EPDE_ERR_ARG;
  if (((uintptr_t)X & 15) || ((uintptr_t)ws & 255)) return EPDE_ERR_ALIGN;
  size_t need; rc = epde_workspace_bytes(D, B, &need); if (rc) return rc;
  if (ws_bytes < need) return EPDE_ERR_WORKSPACE;
  return EPDE_OK;
}

// ---- one-shot all-reduce over NVLink peer memory (the two tiny exchanges of the sharded step) --------------------------
// (protocol: peer_allreduce_block in epde_common.cuh)
template <typename T>
__global__ void __launch_bounds__(1024)
k_peer_allreduce(char* const* __restrict__ bufs, int rank, int world, uint32_t epoch_arg, uint32_t* __restrict__ epoch_ctr,
                 const T* __restrict__ src, T* __restrict__ dst, int n, int64_t slot_bytes) {
  // epoch: an argument, or (epoch_ctr != NULL) a counter in device memory that every call advances by one -- the launch
  // arguments are then invariant and the exchange can sit inside a replayed CUDA graph
  const uint32_t epoch = epoch_ctr ? *epoch_ctr + 1u : epoch_arg;
  const Peer P = {bufs, rank, world, epoch_ctr, (long long)slot_bytes};
  peer_allreduce_block<T>(P, epoch, src, dst, (T*)nullptr, n);
  if (epoch_ctr && threadIdx.x == 0) *epoch_ctr = epoch;      // every thread read it before the first __syncthreads
}

}  // namespace

extern "C" {

#ifdef EPDE_PHASE_TIMING
// profiling aid (only in -DEPDE_PHASE_TIMING builds): cycles thread 0 of each CTA spent per phase
int epde_debug_phase_cycles(unsigned long long* h8, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(h8, g_phase_cycles, sizeof(unsigned long long) * 8);
  if (e == cudaSuccess && reset) { unsigned long long z[8] = {0}; e = cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
  return (int)e;
}
#endif

int epde_abi_version(void) { return EPDE_ABI_VERSION; }

const char* epde_error_string(int code) {
  switch (code) {
    case EPDE_OK: return "ok";
    case EPDE_ERR_NULL: return "required pointer is NULL";
    case EPDE_ERR_DESC: return "malformed descriptor";
    case EPDE_ERR_UNSUPPORTED: return "configuration not supported by this build";
    case EPDE_ERR_WORKSPACE: return "workspace too small";
    case EPDE_ERR_ALIGN: return "pointer or stride alignment";
    case EPDE_ERR_ARG: return "invalid argument";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int epde_param_count(const epde_desc* D, int64_t* out) {
  int rc = check_desc(D); if (rc) return rc;
  if (!out) return EPDE_ERR_NULL;
  *out = param_count(D);
  return EPDE_OK;
}

int epde_num_sums(const epde_desc* D, int32_t* out) {
  int rc = check_desc(D); if (rc) return rc;
  if (!out) return EPDE_ERR_NULL;
  *out = num_sums(D->k);
  return EPDE_OK;
}

int epde_has_fused_path(const epde_desc* D) {
  int rc = check_desc(D); if (rc) return rc;
  return fused_ok(D) ? 1 : 0;
}

int epde_workspace_bytes(const epde_desc* D, int64_t B, size_t* out) {
  int rc = check_desc(D); if (rc) return rc;
  if (!out) return EPDE_ERR_NULL;
  if (B < 1) return EPDE_ERR_ARG;
  // + kSumsTail bytes at the end: scratch for the sums of the single-call epde_step
  if (fused_ok(D)) { *out = fused_ws(D, B, nullptr).total + kSumsTail; return EPDE_OK; }
  rc = generic_workspace_bytes(D, B, out);
  if (rc == EPDE_OK) *out = align_up(*out, 256) + kSumsTail;
  return rc;
}

int epde_step_partial(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
                      const float* params, const float* diag, void* ws, size_t ws_bytes, double* sums,
                      void* stream) {
  int rc = check_common(D, X, w, B, params, diag, ws, ws_bytes); if (rc) return rc;
  if (!sums) return EPDE_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (fused_ok(D)) return fused_partial(D, X, w, idx, B, params, diag, ws, sums, st);
  return generic_partial(D, X, w, idx, B, params, diag, ws, sums, st);
}

int epde_step_finish(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
                     const float* params, const float* diag, double beta, double alpha, const double* h_eig_w,
                     const double* sums, void* ws, size_t ws_bytes, double* out, float* grad, void* stream) {
  int rc = check_common(D, X, w, B, params, diag, ws, ws_bytes); if (rc) return rc;
  if (!h_eig_w || !sums || !out || !grad) return EPDE_ERR_NULL;
  EigW ew; memset(&ew, 0, sizeof(ew));
  for (int i = 0; i < D->k; i++) ew.v[i] = h_eig_w[i];
  cudaStream_t st = (cudaStream_t)stream;
  if (fused_ok(D)) return fused_finish(D, X, w, idx, B, params, diag, beta, alpha, ew, sums, ws, out, grad, st);
  return generic_finish(D, X, w, idx, B, params, diag, beta, alpha, ew, sums, ws, out, grad, st);
}

int epde_step(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
              const float* params, const float* diag, double beta, double alpha, const double* h_eig_w,
              void* ws, size_t ws_bytes, double* out, float* grad, void* stream) {
  int rc = check_common(D, X, w, B, params, diag, ws, ws_bytes); if (rc) return rc;
  if (!out) return EPDE_ERR_NULL;
  size_t need; epde_workspace_bytes(D, B, &need);
  double* sums = (double*)((char*)ws + need - kSumsTail);
  if (fused_ok(D) && fused_impl(D) == 1) {       // single rank, tensor-core family: one launch less between the phases
    if (!h_eig_w || !grad) return EPDE_ERR_NULL;
    EigW ew; memset(&ew, 0, sizeof(ew));
    for (int i = 0; i < D->k; i++) ew.v[i] = h_eig_w[i];
    Deferred df = {1, 0, 0};
    rc = fused_partial(D, X, w, idx, B, params, diag, ws, sums, (cudaStream_t)stream, &df); if (rc) return rc;
    return fused_finish(D, X, w, idx, B, params, diag, beta, alpha, ew, sums, ws, out, grad, (cudaStream_t)stream, &df);
  }
  rc = epde_step_partial(D, X, w, idx, B, params, diag, ws, ws_bytes, sums, stream); if (rc) return rc;
  return epde_step_finish(D, X, w, idx, B, params, diag, beta, alpha, h_eig_w, sums, ws, ws_bytes, out, grad, stream);
}

int epde_step_sharded(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
                      const float* params, const float* diag, double beta, double alpha, const double* h_eig_w,
                      void* ws, size_t ws_bytes, double* out, float* grad, const epde_peer* peer, void* stream) {
  int rc = check_common(D, X, w, B, params, diag, ws, ws_bytes); if (rc) return rc;
  if (!out || !h_eig_w || !grad || !peer || !peer->d_peer_bufs || !peer->d_epoch) return EPDE_ERR_NULL;
  if (peer->world < 2 || peer->world > 32 || peer->rank < 0 || peer->rank >= peer->world || (peer->slot_bytes & 15))
    return EPDE_ERR_ARG;
  if (!fused_ok(D) || fused_impl(D) != 1) return EPDE_ERR_UNSUPPORTED;      // tensor-core family only
  if (param_count(D) * 4 > peer->slot_bytes || num_sums(D->k) * 8 > peer->slot_bytes) return EPDE_ERR_ARG;
  size_t need; epde_workspace_bytes(D, B, &need);
  double* sums = (double*)((char*)ws + need - kSumsTail);
  EigW ew; memset(&ew, 0, sizeof(ew));
  for (int i = 0; i < D->k; i++) ew.v[i] = h_eig_w[i];
  const Peer P = {reinterpret_cast<char* const*>(peer->d_peer_bufs), peer->rank, peer->world, peer->d_epoch,
                  (long long)peer->slot_bytes};
  Deferred df = {1, 0, 0};
  rc = fused_partial(D, X, w, idx, B, params, diag, ws, sums, (cudaStream_t)stream, &df); if (rc) return rc;
  return fused_finish(D, X, w, idx, B, params, diag, beta, alpha, ew, sums, ws, out, grad, (cudaStream_t)stream, &df, &P);
}

int epde_gather_ahead(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B, void* ws,
                      size_t ws_bytes, int32_t slot, void* stream) {
  int rc = check_desc(D); if (rc) return rc;
  if (!X || !w || !ws) return EPDE_ERR_NULL;
  if (B < 1 || slot < 0 || slot > 1) return EPDE_ERR_ARG;
  if (((uintptr_t)X & 15) || ((uintptr_t)ws & 255)) return EPDE_ERR_ALIGN;
  if (!fused_ok(D) || fused_impl(D) != 1) return EPDE_ERR_UNSUPPORTED;
  size_t need; rc = epde_workspace_bytes(D, B, &need); if (rc) return rc;
  if (ws_bytes < need) return EPDE_ERR_WORKSPACE;
  return gather_ahead(D, X, w, idx, B, ws, slot, (cudaStream_t)stream);
}

int epde_step_ahead(const epde_desc* D, const float* X, const float* w, const int64_t* idx, int64_t B,
                    const float* params, const float* diag, double beta, double alpha, const double* h_eig_w,
                    void* ws, size_t ws_bytes, double* out, float* grad, int32_t slot, int32_t pregathered,
                    int32_t reserve_sms, const epde_peer* peer, void* event_phase1, void* stream) {
  int rc = check_common(D, X, w, B, params, diag, ws, ws_bytes); if (rc) return rc;
  if (!out || !h_eig_w || !grad) return EPDE_ERR_NULL;
  if (slot < 0 || slot > 1 || reserve_sms < 0 || reserve_sms > 64) return EPDE_ERR_ARG;
  if (!fused_ok(D) || fused_impl(D) != 1) return EPDE_ERR_UNSUPPORTED;      // tensor-core family only
  Peer P = {nullptr, 0, 1, nullptr, 0};
  if (peer) {
    if (!peer->d_peer_bufs || !peer->d_epoch) return EPDE_ERR_NULL;
    if (peer->world < 2 || peer->world > 32 || peer->rank < 0 || peer->rank >= peer->world || (peer->slot_bytes & 15) ||
        param_count(D) * 4 > peer->slot_bytes || num_sums(D->k) * 8 > peer->slot_bytes)
      return EPDE_ERR_ARG;
    P = Peer{reinterpret_cast<char* const*>(peer->d_peer_bufs), peer->rank, peer->world, peer->d_epoch, (long long)peer->slot_bytes};
  }
  size_t need; epde_workspace_bytes(D, B, &need);
  double* sums = (double*)((char*)ws + need - kSumsTail);
  EigW ew; memset(&ew, 0, sizeof(ew));
  for (int i = 0; i < D->k; i++) ew.v[i] = h_eig_w[i];
  const Ahead ah = {slot, pregathered ? 1 : 0, reserve_sms};
  Deferred df = {1, 0, 0};
  const Ahead ah1 = {slot, pregathered ? 1 : 0, 0};      // phase 1 is bound by HBM itself: it keeps all SMs, and the caller
  rc = fused_partial(D, X, w, idx, B, params, diag, ws, sums, (cudaStream_t)stream, &df, &ah1); if (rc) return rc;      // starts
  if (event_phase1) {                                    // the gather only behind it (event_phase1)
    cudaError_t e = cudaEventRecord((cudaEvent_t)event_phase1, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
  }
  return fused_finish(D, X, w, idx, B, params, diag, beta, alpha, ew, sums, ws, out, grad, (cudaStream_t)stream, &df,
                      peer ? &P : nullptr, &ah);
}

int epde_forward(const epde_desc* D, const float* X, const int64_t* idx, int64_t B, const float* params,
                 const float* diag, void* ws, size_t ws_bytes, float* y, void* stream) {
  int rc = check_desc(D); if (rc) return rc;
  if (!X || !params || !y || !ws || !diag) return EPDE_ERR_NULL;
  if (B < 1) return EPDE_ERR_ARG;
  if (((uintptr_t)X & 15) || ((uintptr_t)ws & 255)) return EPDE_ERR_ALIGN;
  const int64_t slab = B < kFwdSlab ? B : kFwdSlab;
  size_t need; rc = epde_workspace_bytes(D, slab, &need); if (rc) return rc;
  if (ws_bytes < need) return EPDE_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (!fused_ok(D)) return generic_forward(D, X, idx, B, params, ws, ws_bytes, y, st);
  // forward in slabs of at most kFwdSlab samples so the workspace stays small (sized for kFwdSlab)
  FusedWs W = fused_ws(D, slab, ws);
  int nsm; rc = sm_count(&nsm); if (rc) return rc;
  k_pack<kFusedH><<<D->k, 256, 0, st>>>(params, diag, D->d, D->d_pad, W.pw);
  const size_t smem = (size_t)D->k * PackLayout<kFusedH>(D->d_pad).total() * 4;
  cudaError_t e = general_act(D)
      ? cudaFuncSetAttribute(k_forward<kFusedH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
      : cudaFuncSetAttribute(k_forward<kFusedH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  for (int64_t b0 = 0; b0 < B; b0 += kFwdSlab) {
    const int64_t bc = (B - b0 < kFwdSlab) ? B - b0 : kFwdSlab;
    // identity indexing: shift the base pointer; explicit indices: shift the index pointer
    e = launch_gather(D, idx ? X : X + b0 * D->d_pad, nullptr, idx ? idx + b0 : nullptr, bc, W.xt, W.wt, st);
    if (e != cudaSuccess) return (int)e;
    const int64_t npairs = (bc + 1) / 2;
    int grid = (int)((npairs + P1_THREADS - 1) / P1_THREADS);
    if (grid > 4 * nsm) grid = 4 * nsm;
    if (general_act(D))
      k_forward<kFusedH, true><<<grid, P1_THREADS, smem, st>>>(W.xt, bc, D->d, D->d_pad, D->k, W.pw, y + b0 * D->k, D->act_id);
    else
      k_forward<kFusedH><<<grid, P1_THREADS, smem, st>>>(W.xt, bc, D->d, D->d_pad, D->k, W.pw, y + b0 * D->k);
  }
  return (int)cudaGetLastError();
}

// ---- host-side MT19937 (CPython `random` compatible) ----------------------------------------
static void mt_twist(uint32_t* mt) {        // three index ranges instead of two modulo operations per word
  const int N = 624, M = 397;
  auto f = [](uint32_t a, uint32_t b) { const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu); return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); };
  int kk = 0;
  for (; kk < N - M; kk++) mt[kk] = mt[kk + M] ^ f(mt[kk], mt[kk + 1]);
  for (; kk < N - 1; kk++) mt[kk] = mt[kk + (M - N)] ^ f(mt[kk], mt[kk + 1]);
  mt[N - 1] = mt[M - 1] ^ f(mt[N - 1], mt[0]);
}
static inline uint32_t mt_next(uint32_t* mt, int32_t* pos) {
  if (*pos >= 624) { mt_twist(mt); *pos = 0; }
  return mt_temper(mt[(*pos)++]);
}

int epde_minibatch_indices(uint32_t* mt, int32_t* pos, int64_t K, int64_t B, int64_t* out) {
  if (!mt || !pos || !out) return EPDE_ERR_NULL;
  if (K < 1 || B < 0 || *pos < 0 || *pos > 624) return EPDE_ERR_ARG;
  const double total = (double)K;
  for (int64_t i = 0; i < B; i++) {
    const uint32_t a = mt_next(mt, pos) >> 5, b = mt_next(mt, pos) >> 6;
    const double u = ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
    int64_t j = (int64_t)floor(u * total);     // bisect over cum_weights = 1..K (src/data_set.py:31,67)
    if (j > K - 1) j = K - 1;
    out[i] = j;
  }
  return EPDE_OK;
}

int epde_minibatch_indices_device(uint32_t* d_state, int64_t K, int64_t B_total, int64_t out_lo, int64_t out_n,
                                  uint32_t* d_raw, int64_t* d_idx, void* stream) {
  if (!d_state || !d_raw || (!d_idx && out_n > 0)) return EPDE_ERR_NULL;
  if (K < 1 || B_total < 0 || B_total >= (1ll << 30) || out_lo < 0 || out_n < 0 || out_lo + out_n > B_total) return EPDE_ERR_ARG;
  if (((uintptr_t)d_raw & 7) || ((uintptr_t)d_state & 3)) return EPDE_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  k_mt_generate<<<1, MT_THREADS, 0, st>>>(d_state, (int)(2 * B_total), d_raw, (int)(2 * B_total));
  if (out_n > 0) k_mt_to_idx<<<(unsigned)((out_n + 255) / 256), 256, 0, st>>>(d_raw, K, out_lo, out_n, d_idx);
  return (int)cudaGetLastError();
}

int epde_minibatch_indices_device_split(uint32_t* d_states, int32_t parts, int64_t K, int64_t n_sub, int64_t n_total,
                                        uint32_t* d_raw, int64_t* d_idx, void* stream) {
  if (!d_states || !d_raw || !d_idx) return EPDE_ERR_NULL;
  if (K < 1 || parts < 1 || parts > 64 || n_sub < 1 || n_total < 1 || n_total >= (1ll << 30) ||
      (int64_t)(parts - 1) * n_sub >= n_total || (int64_t)parts * n_sub < n_total)
    return EPDE_ERR_ARG;
  if (((uintptr_t)d_raw & 7) || ((uintptr_t)d_states & 3)) return EPDE_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  k_mt_generate<<<parts, MT_THREADS, 0, st>>>(d_states, (int)(2 * n_sub), d_raw, (int)(2 * n_total));
  k_mt_to_idx<<<(unsigned)((n_total + 255) / 256), 256, 0, st>>>(d_raw, K, 0, n_total, d_idx);
  return (int)cudaGetLastError();
}

int epde_mt_seed(const uint32_t* key, int32_t key_len, uint32_t* mt, int32_t* pos) {
  if (!key || !mt || !pos) return EPDE_ERR_NULL;
  if (key_len < 1) return EPDE_ERR_ARG;
  const int N = 624;
  mt[0] = 19650218u;
  for (int i = 1; i < N; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  int i = 1, j = 0;
  for (int kk = (N > key_len ? N : key_len); kk; kk--) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    i++; j++;
    if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
    if (j >= key_len) j = 0;
  }
  for (int kk = N - 1; kk; kk--) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    i++;
    if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
  }
  mt[0] = 0x80000000u;
  *pos = N;
  return EPDE_OK;
}

// ---- fused Adam -----------------------------------------------------------------------------
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, float lr, float bc1, float bc2_sqrt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = 0.9f * m[i] + 0.1f * gi;
  const float vi = 0.999f * v[i] + 0.001f * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + 1e-8f;
  p[i] -= (lr / bc1) * (mi / denom);
}

// The whole per-step epilogue of `train` (:277-284) in ONE launch whose arguments never change, so that a step can be
// replayed as a CUDA graph: Adam with the hyper-parameters read from device memory (state[0] = lr, state[1] = number
// of optimiser steps taken so far, incremented here), then averaged_model.nets[i] += model.nets[cvec[i]] with the
// UPDATED parameters, then the running sums of the sorted eigenvalue estimates.  One CTA (the parameter vector is a
// few thousand floats): the averaging reads other nets' parameters, so it has to follow a CTA-wide barrier.
__global__ void __launch_bounds__(1024)
k_epilogue(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
           double* __restrict__ state, double* __restrict__ avg, double* __restrict__ stats,
           const double* __restrict__ out, int k) {
  const double lr = state[0], step = state[1] + 1.0;
  const float bc1 = (float)(1.0 - pow(0.9, step)), bc2s = (float)sqrt(1.0 - pow(0.999, step));
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i];
    const float mi = 0.9f * m[i] + 0.1f * gi;
    const float vi = 0.999f * v[i] + 0.001f * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= ((float)lr / bc1) * (mi / (sqrtf(vi) / bc2s + 1e-8f));
  }
  __syncthreads();
  const int64_t per = n / k;
  for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = (int)(e / per);
    const int src = (int)out[EPDE_OUT_CVEC(k) + i];
    avg[e] += (double)p[(int64_t)src * per + (e - (int64_t)i * per)];
  }
  if (threadIdx.x < k) { const double ev = out[EPDE_OUT_EIG(k) + threadIdx.x]; stats[threadIdx.x] += ev; stats[k + threadIdx.x] += ev * ev; }
  if (threadIdx.x == 0) state[1] = step;
}

// The same epilogue with fp64 master parameters and fp64 Adam moments (what the reference's `model.double()` +
// torch.optim.Adam hold, src/pytorch_eigen_solver.py:374-379): the fp32 copy p32 is what the next step's kernels read.
// torch's update order: m.lerp_(g, 1 - b1); v = b2 v + (1 - b2) g g; denom = sqrt(v) / sqrt(1 - b2^t) + eps;
// p -= (lr / (1 - b1^t)) m / denom.
__global__ void __launch_bounds__(1024)
k_epilogue64(double* __restrict__ p, float* __restrict__ p32, const float* __restrict__ g, double* __restrict__ m,
             double* __restrict__ v, int64_t n, double* __restrict__ state, double* __restrict__ avg,
             double* __restrict__ stats, const double* __restrict__ out, int k) {
  const double lr = state[0], step = state[1] + 1.0;
  const double bc1 = 1.0 - pow(0.9, step), bc2s = sqrt(1.0 - pow(0.999, step));
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double gi = (double)g[i];
    const double mi = m[i] + (gi - m[i]) * (1.0 - 0.9);
    const double vi = v[i] * 0.999 + (1.0 - 0.999) * gi * gi;
    m[i] = mi; v[i] = vi;
    const double pi = p[i] - (lr / bc1) * (mi / (sqrt(vi) / bc2s + 1e-8));
    p[i] = pi; p32[i] = (float)pi;
  }
  __syncthreads();
  const int64_t per = n / k;
  for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = (int)(e / per);
    const int src = (int)out[EPDE_OUT_CVEC(k) + i];
    avg[e] += p[(int64_t)src * per + (e - (int64_t)i * per)];
  }
  if (threadIdx.x < k) { const double ev = out[EPDE_OUT_EIG(k) + threadIdx.x]; stats[threadIdx.x] += ev; stats[k + threadIdx.x] += ev * ev; }
  if (threadIdx.x == 0) state[1] = step;
}

// averaged_model.nets[i] += model.nets[cvec[i]]  (record_model_parameters, src/pytorch_eigen_solver.py:114-118)
__global__ void k_avg_acc(double* __restrict__ avg, const float* __restrict__ p, const double* __restrict__ out,
                          int k, int64_t per) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per * k) return;
  const int i = (int)(e / per);
  const int src = (int)out[EPDE_OUT_CVEC(k) + i];
  avg[e] += (double)p[(int64_t)src * per + (e - (int64_t)i * per)];
}
// running sums of the sorted eigenvalue estimates and their squares (:280-282)
__global__ void k_eig_stats(double* __restrict__ stats, const double* __restrict__ out, int k) {
  const int i = threadIdx.x;
  if (i < k) { const double e = out[EPDE_OUT_EIG(k) + i]; stats[i] += e; stats[k + i] += e * e; }
}

static int peer_allreduce(void* const* d_peer_bufs, int rank, int world, uint32_t epoch, uint32_t* d_epoch, const void* src,
                          void* dst, int64_t n, int is_f64, int64_t slot_bytes, void* stream) {
  if (!d_peer_bufs || !src || !dst) return EPDE_ERR_NULL;
  if (world < 1 || world > 32 || rank < 0 || rank >= world || n < 1 || (!d_epoch && epoch == 0) ||
      n * (is_f64 ? 8 : 4) > slot_bytes || (slot_bytes & 15)) return EPDE_ERR_ARG;
  char* const* bufs = reinterpret_cast<char* const*>(d_peer_bufs);
  if (is_f64) k_peer_allreduce<double><<<1, 1024, 0, (cudaStream_t)stream>>>(bufs, rank, world, epoch, d_epoch, (const double*)src, (double*)dst, (int)n, slot_bytes);
  else k_peer_allreduce<float><<<1, 1024, 0, (cudaStream_t)stream>>>(bufs, rank, world, epoch, d_epoch, (const float*)src, (float*)dst, (int)n, slot_bytes);
  return (int)cudaGetLastError();
}

int epde_peer_allreduce(void* const* d_peer_bufs, int rank, int world, uint32_t epoch, const void* src, void* dst,
                        int64_t n, int is_f64, int64_t slot_bytes, void* stream) {
  return peer_allreduce(d_peer_bufs, rank, world, epoch, nullptr, src, dst, n, is_f64, slot_bytes, stream);
}

int epde_peer_allreduce_graph(void* const* d_peer_bufs, int rank, int world, uint32_t* d_epoch, const void* src, void* dst,
                              int64_t n, int is_f64, int64_t slot_bytes, void* stream) {
  if (!d_epoch) return EPDE_ERR_NULL;
  return peer_allreduce(d_peer_bufs, rank, world, 0, d_epoch, src, dst, n, is_f64, slot_bytes, stream);
}

int epde_md_align(int64_t K, int32_t n_atoms, int32_t d_pad, const float* X, const double* ref_x, const int32_t* ref_index,
                  int32_t m, const float* diag, float* Xa, float* metric, void* stream) {
  if (!X || !ref_x || !ref_index || !diag || !Xa) return EPDE_ERR_NULL;
  if (K <= 0 || n_atoms < 1 || n_atoms > MD_MAX_ATOMS || m < 1 || m > MD_MAX_REF || d_pad < 3 * n_atoms) return EPDE_ERR_ARG;
  MdRef ref;
  ref.m = m;
  for (int s = 0; s < m; s++) {
    if (ref_index[s] < 0 || ref_index[s] >= n_atoms) return EPDE_ERR_ARG;
    ref.index[s] = ref_index[s];
    for (int i = 0; i < 3; i++) ref.q[s][i] = ref_x[3 * s + i];
  }
  k_md_align<<<(unsigned)K, 64, 0, (cudaStream_t)stream>>>(X, K, n_atoms, d_pad, ref, diag, Xa, metric);
  return (int)cudaGetLastError();
}

int epde_train_epilogue(const epde_desc* D, float* params, const float* grad, float* m, float* v, double* state,
                        double* avg, double* stats, const double* out, void* stream) {
  int rc = check_desc(D); if (rc) return rc;
  if (!params || !grad || !m || !v || !state || !avg || !stats || !out) return EPDE_ERR_NULL;
  k_epilogue<<<1, 1024, 0, (cudaStream_t)stream>>>(params, grad, m, v, param_count(D), state, avg, stats, out, D->k);
  return (int)cudaGetLastError();
}

int epde_train_epilogue64(const epde_desc* D, double* params64, float* params32, const float* grad, double* m, double* v,
                          double* state, double* avg, double* stats, const double* out, void* stream) {
  int rc = check_desc(D); if (rc) return rc;
  if (!params64 || !params32 || !grad || !m || !v || !state || !avg || !stats || !out) return EPDE_ERR_NULL;
  k_epilogue64<<<1, 1024, 0, (cudaStream_t)stream>>>(params64, params32, grad, m, v, param_count(D), state, avg, stats, out, D->k);
  return (int)cudaGetLastError();
}

int epde_average_accumulate(const epde_desc* D, double* avg, const float* params, const double* out, void* stream) {
  int rc = check_desc(D); if (rc) return rc;
  if (!avg || !params || !out) return EPDE_ERR_NULL;
  const int64_t n = param_count(D), per = n / D->k;
  k_avg_acc<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(avg, params, out, D->k, per);
  return (int)cudaGetLastError();
}

int epde_eig_stats_accumulate(const epde_desc* D, double* stats, const double* out, void* stream) {
  int rc = check_desc(D); if (rc) return rc;
  if (!stats || !out) return EPDE_ERR_NULL;
  k_eig_stats<<<1, 32, 0, (cudaStream_t)stream>>>(stats, out, D->k);
  return (int)cudaGetLastError();
}

int epde_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, int64_t step, void* stream) {
  if (!p || !g || !m || !v) return EPDE_ERR_NULL;
  if (n < 1 || step < 1) return EPDE_ERR_ARG;
  const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = 1.0 - pow(0.999, (double)step);
  k_adam<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)lr, (float)bc1,
                                                                         (float)sqrt(bc2));
  return (int)cudaGetLastError();
}

}  // extern "C"
