// Batch-global coefficient kernel shared by the fused and the generic kernel families.
#pragma once
#include <cuda_runtime.h>
#include "../../include/eigenpde_b200.h"

namespace epde {

struct Coef {                      // written by k_coef, read by k_pass2
  float ecoef[8];                  // ebar_{n,i} = w_n * ecoef[i]
  float cm[8];                     // ybar_{n,i} = w_n * (sum_j Cw[i][j] y_{n,j} - cm[i])
  float Cw[8][8];
};


struct EigW { double v[EPDE_MAX_K]; };

// Everything that depends only on the batch-global sums: eig_vals, cvec, losses
// (src/pytorch_eigen_solver.py:215-229, 177-189) and the per-sample seed coefficients of the
// reverse pass (SURVEY.md 8(a) closed form).  One thread; k <= 8.
template <int DUMMY = 0>
__global__ void k_coef(const double* __restrict__ sums, int k, double beta, double alpha, EigW eigw, int sort,
                       double* __restrict__ out, Coef* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double W = sums[0];
  const double* S1 = sums + 1;
  const double* E = sums + 1 + k;
  const double* S2 = sums + 1 + 2 * k;
  double m[EPDE_MAX_K], var[EPDE_MAX_K], Rq[EPDE_MAX_K], lam[EPDE_MAX_K], wt[EPDE_MAX_K];
  double cov[EPDE_MAX_K][EPDE_MAX_K];
  int cvec[EPDE_MAX_K];
  for (int i = 0; i < k; i++) m[i] = S1[i] / W;
  int t = 0;
  for (int i = 0; i < k; i++)
    for (int j = 0; j <= i; j++) { cov[i][j] = cov[j][i] = S2[t] / W - m[i] * m[j]; t++; }
  for (int i = 0; i < k; i++) {
    var[i] = cov[i][i];
    Rq[i] = E[i] / (W * beta);
    lam[i] = Rq[i] / var[i];
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
      coef->ecoef[i] = (float)(wt[i] / (W * beta * var[i]));
      double cmi = 0.0;
      for (int j = 0; j < k; j++) {
        const double c = (i == j) ? 2.0 * (-wt[i] * Rq[i] / (var[i] * var[i]) + 2.0 * alpha * (var[i] - 1.0))
                                  : 2.0 * alpha * cov[i][j];
        coef->Cw[i][j] = (float)(c / W);
        cmi += c / W * m[j];
      }
      coef->cm[i] = (float)cmi;
    }
  }
}


}  // namespace epde
