// Host-side entry points of the tensor-core fused family (kernels in epde_tc_fused.cuh, compiled in
// epde_tc_fused.cu as a separate translation unit).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace epde {
struct Coef;
struct EigW;
struct Peer;
namespace tc {

size_t pack_floats(int d);          // floats of one net's packed weight image
size_t stash_floats(int64_t B, int k);   // floats of the phase-1 -> phase-2 stash (epde_tc_fused.cuh)

// weight images of all k nets (k_pack_tc); to be launched before the gather, which then runs alongside it (PDL)
int pack(int d, int k, const float* params, const float* diag, float* pwt, float* nrm, cudaStream_t st);

// phase 1 after pack() and the gather: k_pass1_tc + k_sums_y + k_reduce_tc -> sums[1 + 2k + k(k+1)/2], ybuf[B][k]
int partial(int d, int d_pad, int k, int nsm, int64_t B, const float* params, const float* diag, const float* xt,
            const float* wt, float* pwt, float* ybuf, double* part, double* part2, double* sums, float* mx, float* nrm,
            float* stash, cudaStream_t st, int defer_reduce = 0, int* gx_out = nullptr, int* g2_out = nullptr);

// single-rank step only: reduction of phase 1's partials + the coefficient kernel in one launch (epde_step passes
// defer_reduce = 1 to partial() and calls this instead of k_reduce_tc / k_coef)
int reduce_coef(const double* part, int gx, const double* part2, int g2, int k, double* sums, double beta, double alpha,
                const EigW& ew, int sort, double* out, Coef* coef, const float* mx, const float* nrm, cudaStream_t st,
                const Peer* peer = nullptr);      // peer: sharded step - the sums are all-reduced inside the kernel

// phase 2: k_pass2_tc -> per-CTA gradient images pg[(cta * k + net) * GradLayout.total()], *gx_out = CTAs per net
int finish(int d, int d_pad, int k, int nsm, int64_t B, const float* xt, const float* wt, const float* pwt,
           const float* ybuf, const Coef* coef, const float* stash, float* pg, int* gx_out, cudaStream_t st);

}  // namespace tc
}  // namespace epde
