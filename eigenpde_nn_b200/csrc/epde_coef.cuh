// Batch-global coefficient kernel shared by the fused and the generic kernel families.
#pragma once
#include "epde_common.cuh"
#include <cuda_runtime.h>
#include "../../include/eigenpde_b200.h"

namespace epde {

struct Coef {                      // written by k_coef, read by k_pass2
  float ecoef[8];                  // ebar_{n,i} = w_n * ecoef[i]
  float cm[8];                     // ybar_{n,i} = w_n * (sum_j Cw[i][j] y_{n,j} - cm[i])
  float Cw[8][8];
  // tensor-core family, fp16 gradient rounds: power-of-two staging scales per net and vector type
  //   [0] g2  [1] g3  [2] ebar g1  [3] ubar1  [4] ubar2  [5] g1  [6] zbar2  [7] zbar3  [8] s4  [9] ybar  [10] a / 1
  //   [11] zbar1  [12] x / 1
  // and the factors that undo them when an accumulator block is added to the gradient image
  //   [0..2] round 1 blocks  [3] zbar2 x a1  [4] zbar3 x a2  [5] s4 x 1  [6] ybar x 1  [7] zbar1 x x
  float sc[8][16];
  float fl[8][8];
};

// Bounds for the tensor-core family's fp16 staging (all device pointers, fp32):
//   mx[0] = max w over the minibatch, mx[1] = max |x|, mx[2 + 4 net + {0: |g2|, 1: |g1|, 2: |K g1|, 3: |y|}]  (as float bits,
//   written with atomicMax by k_gather128 / k_pass1_tc), nrm[8 net + {0: |W2|_inf, 1: |W2^T|_inf, 2: |W3|_inf,
//   3: |W3^T|_inf, 4: max |w4|}] (k_pack_tc)
constexpr int kMxPerNet = 4, kNrmPerNet = 8;


struct EigW { double v[EPDE_MAX_K]; };

// Everything that depends only on the batch-global sums: eig_vals, cvec, losses
// (src/pytorch_eigen_solver.py:215-229, 177-189) and the per-sample seed coefficients of the
// reverse pass (SURVEY.md 8(a) closed form).  One thread (k <= 8); the staging bounds of the tensor-core family: one thread per net.
template <int DUMMY = 0>
__device__ __forceinline__ void coef_body(const double* sums, int k, double beta, double alpha, const EigW& eigw, int sort,
                                          double* __restrict__ out, Coef* __restrict__ coef, const float* __restrict__ mx,
                                          const float* __restrict__ nrm) {
  if (threadIdx.x == 0) {
  const double W = sums[0];
  const double* S1 = sums + 1;
  const double* E = sums + 1 + k;
  const double* S2 = sums + 1 + 2 * k;
  double m[EPDE_MAX_K], var[EPDE_MAX_K], Rq[EPDE_MAX_K], lam[EPDE_MAX_K], wt[EPDE_MAX_K];
  double cov[EPDE_MAX_K][EPDE_MAX_K];
  int cvec[EPDE_MAX_K];
  const double iW = 1.0 / W, iWb = 1.0 / (W * beta);      // (fp64 divisions are ~100 cycles each on this single thread)
  for (int i = 0; i < k; i++) m[i] = S1[i] * iW;
  int t = 0;
  for (int i = 0; i < k; i++)
    for (int j = 0; j <= i; j++) { cov[i][j] = cov[j][i] = S2[t] * iW - m[i] * m[j]; t++; }
  double ivar[EPDE_MAX_K];
  for (int i = 0; i < k; i++) {
    var[i] = cov[i][i];
    ivar[i] = 1.0 / var[i];
    Rq[i] = E[i] * iWb;
    lam[i] = Rq[i] * ivar[i];
    cvec[i] = i;
  }
  if (sort) {   // ascending, stable insertion sort (np.argsort order for distinct values)
    for (int i = 1; i < k; i++) {
      const int c = cvec[i]; int j = i - 1;
      while (j >= 0 && lam[cvec[j]] > lam[c]) { cvec[j + 1] = cvec[j]; j--; }
      cvec[j + 1] = c;
    }
  }
  for (int p = 0; p < k; p++) wt[cvec[p]] = eigw.v[p];
  double nonpen = 0.0, pen = 0.0;
  for (int i = 0; i < k; i++) { nonpen += wt[i] * lam[i]; pen += (var[i] - 1.0) * (var[i] - 1.0); }
  for (int i = 0; i < k; i++)
    for (int j = i + 1; j < k; j++) pen += cov[i][j] * cov[i][j];
  out[EPDE_OUT_LOSS] = nonpen + alpha * pen;
  out[EPDE_OUT_NONPEN] = nonpen;
  out[EPDE_OUT_PENALTY] = pen;
  out[EPDE_OUT_TOTW] = W;
  for (int p = 0; p < k; p++) {
    out[EPDE_OUT_EIG(k) + p] = lam[cvec[p]];
    out[EPDE_OUT_CVEC(k) + p] = (double)cvec[p];
    out[EPDE_OUT_MEAN(k) + p] = m[p];
    out[EPDE_OUT_VAR(k) + p] = var[p];
  }
  if (coef) {
    for (int i = 0; i < k; i++) {
      coef->ecoef[i] = (float)(wt[i] * iWb * ivar[i]);
      double cmi = 0.0;
      for (int j = 0; j < k; j++) {
        const double c = (i == j) ? 2.0 * (-wt[i] * Rq[i] * ivar[i] * ivar[i] + 2.0 * alpha * (var[i] - 1.0))
                                  : 2.0 * alpha * cov[i][j];
        coef->Cw[i][j] = (float)(c * iW);
        cmi += c * iW * m[j];
      }
      coef->cm[i] = (float)cmi;
    }
  }
  __threadfence_block();        // the per-net bound computation below reads coef->ecoef / Cw / cm written here
  }   // thread 0
  __syncwarp();
  if (coef && mx && nrm) {
      // a-priori bounds of every vector the gradient rounds stage (per net), from the measured maxima of the forward /
      // Jacobian quantities and the induced infinity norms of the weight matrices; 5 % slack for fp32 rounding
      auto pow2 = [](double bound) -> double {        // s = 2^q with bound * 1.05 * s < 2^14
        const float b = (float)(bound * 1.05);
        if (!(b > 0.f) || !isfinite(b)) return 1.0;
        int q = 14 - ((int)(__float_as_uint(b) >> 23) - 126);      // b in [2^(E-127), 2^(E-126))
        q = q > 60 ? 60 : (q < -60 ? -60 : q);
        return (double)__uint_as_float((uint32_t)(q + 127) << 23);
      };
      const double wmax = mx[0], xmax = fmax((double)mx[1], 1.0);
      {
        const int i = threadIdx.x;                    // one thread per net
        if (i < k) {
        const float* M = mx + 2 + kMxPerNet * i;
        const float* N = nrm + kNrmPerNet * i;
        const double g2m = M[0], g1m = M[1], tm = M[2], g3m = N[4];
        const double nW2 = N[0], nW2T = N[1], nW3 = N[2], nW3T = N[3];
        const double emax = wmax * fabs((double)coef->ecoef[i]);
        double Y = fabs((double)coef->cm[i]);
        for (int j = 0; j < k; j++) Y += fabs((double)coef->Cw[i][j]) * (double)mx[2 + kMxPerNet * j + 3];
        Y *= wmax;
        const double U1 = 2.0 * emax * tm, GB2 = nW2 * U1, GB3 = nW3 * GB2;
        const double S4 = Y + GB3, Z3 = g3m * (2.0 * GB3 + Y);
        const double Z2 = nW3T * Z3 + 2.0 * g2m * GB2, Z1 = nW2T * Z2 + 2.0 * g1m * U1;
        double s[13];
        s[0] = pow2(g2m); s[1] = pow2(g3m); s[2] = pow2(emax * g1m); s[3] = pow2(U1); s[4] = pow2(GB2); s[5] = pow2(g1m);
        s[6] = pow2(Z2); s[7] = pow2(Z3); s[8] = pow2(S4); s[9] = pow2(Y); s[10] = 16384.0; s[11] = pow2(Z1);
        s[12] = pow2(xmax);
        for (int q = 0; q < 16; q++) coef->sc[i][q] = q < 13 ? (float)s[q] : 1.f;
        coef->fl[i][0] = (float)(1.0 / (s[0] * s[3])); coef->fl[i][1] = (float)(1.0 / (s[1] * s[4]));
        coef->fl[i][2] = (float)(1.0 / (s[2] * s[5])); coef->fl[i][3] = (float)(1.0 / (s[6] * s[10]));
        coef->fl[i][4] = (float)(1.0 / (s[7] * s[10])); coef->fl[i][5] = (float)(1.0 / (s[8] * s[10]));
        coef->fl[i][6] = (float)(1.0 / (s[9] * s[10])); coef->fl[i][7] = (float)(1.0 / (s[11] * s[12]));
        }
      }
  }
}
template <int DUMMY = 0>
__global__ void k_coef(const double* __restrict__ sums, int k, double beta, double alpha, EigW eigw, int sort,
                       double* __restrict__ out, Coef* __restrict__ coef, const float* __restrict__ mx = nullptr,
                       const float* __restrict__ nrm = nullptr) {
  pdl_trigger();                     // phase 2's CTAs may set up shared / tensor memory meanwhile (they wait before reading coef)
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  coef_body<DUMMY>(sums, k, beta, alpha, eigw, sort, out, coef, mx, nrm);
}


}  // namespace epde
