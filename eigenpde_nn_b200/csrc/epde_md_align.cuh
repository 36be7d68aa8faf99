// MD pre-processing layer: Kabsch alignment of every configuration to a reference configuration
// (reference: MD_data_set.align, src/data_set.py:126-153) and the per-sample metric M = J diag(c) J^T of its Jacobian
// J = d aligned / d raw.  The layer has no parameters, so both are functions of the data alone and are computed ONCE
// per data set; the step kernels then work on the aligned rows and use e = u0^T M u0 (see epde_generic.cuh).
//
// One CTA per configuration, fp64 throughout:
//   c = mean of the selected atoms, H = sum_s (p_s - c)^T q_s, V / singular directions from the Jacobi
//   eigen-decomposition of H^T H, u_i = H v_i / |H v_i| (i = 1, 2), u_3 = u_1 x u_2, v_3 = v_1 x v_2,
//   R = sum_i u_i v_i^T  ( = U diag(1, 1, sign det(U V^T)) V^T of the reference's SVD expression ),
//   aligned = (p - c) R;
//   Jacobian column for raw coordinate (b, i): e_i R on atom b; if b is a selected atom additionally -e_i R / m on
//   every atom and (p - c) R [omega]_x with (tr(P) I - P) omega = axial(R^T dH - dH^T R), P = R^T H,
//   dH = e_i^T (q_b - mean q)   (derivative of the orthogonal polar factor; the sign is locally constant).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace epde {

constexpr int MD_MAX_ATOMS = 21;      // d <= 63
constexpr int MD_MAX_REF = 21;

struct MdRef {
  int m;                              // selected atoms
  int index[MD_MAX_REF];
  double q[MD_MAX_REF][3];            // reference configuration (rows as in ./data/ref_state.txt)
};

__device__ inline void md_jacobi3(double A[3][3], double V[3][3]) {   // A symmetric -> A = V diag V^T (A is overwritten)
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; sweep++) {
    for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
      if (fabs(A[p][q]) < 1e-300) continue;
      const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
      for (int k = 0; k < 3; k++) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq; }
      for (int k = 0; k < 3; k++) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk; }
      for (int k = 0; k < 3; k++) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
    }
  }
}

// X [K][d_pad] raw fp32 -> Xa [K][d_pad] aligned fp32, metric [K][d][d] fp32 (may be NULL)
__global__ void __launch_bounds__(64)
k_md_align(const float* __restrict__ X, int64_t K, int n_atoms, int d_pad, MdRef ref, const float* __restrict__ diag,
           float* __restrict__ Xa, float* __restrict__ metric) {
  const int64_t n = blockIdx.x;
  if (n >= K) return;
  const int d = 3 * n_atoms, tid = threadIdx.x;
  __shared__ double sp[MD_MAX_ATOMS][3];          // p - c
  __shared__ double sR[3][3], sTi[3][3], sQc[MD_MAX_REF][3];
  __shared__ double sJ[3 * MD_MAX_ATOMS][3 * MD_MAX_ATOMS + 1];
  __shared__ int ssel[MD_MAX_ATOMS];
  if (tid < n_atoms) ssel[tid] = -1;
  __syncthreads();
  if (tid == 0) {
    const float* x = X + n * d_pad;
    double c[3] = {0, 0, 0}, qm[3] = {0, 0, 0};
    for (int s = 0; s < ref.m; s++) {
      ssel[ref.index[s]] = s;
      for (int i = 0; i < 3; i++) { c[i] += (double)x[3 * ref.index[s] + i]; qm[i] += ref.q[s][i]; }
    }
    for (int i = 0; i < 3; i++) { c[i] /= ref.m; qm[i] /= ref.m; }
    for (int a = 0; a < n_atoms; a++) for (int i = 0; i < 3; i++) sp[a][i] = (double)x[3 * a + i] - c[i];
    for (int s = 0; s < ref.m; s++) for (int i = 0; i < 3; i++) sQc[s][i] = ref.q[s][i] - qm[i];
    double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int s = 0; s < ref.m; s++)
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) H[i][j] += sp[ref.index[s]][i] * ref.q[s][j];
    double G[3][3], V[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += H[k][i] * H[k][j]; G[i][j] = s; }
    md_jacobi3(G, V);
    int o0 = 0, o1 = 1, o2 = 2;                   // eigenvalues descending
    if (G[o0][o0] < G[o1][o1]) { int t = o0; o0 = o1; o1 = t; }
    if (G[o1][o1] < G[o2][o2]) { int t = o1; o1 = o2; o2 = t; }
    if (G[o0][o0] < G[o1][o1]) { int t = o0; o0 = o1; o1 = t; }
    double v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
    for (int i = 0; i < 3; i++) { v1[i] = V[i][o0]; v2[i] = V[i][o1]; }
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    double nn = 0;
    for (int i = 0; i < 3; i++) { u1[i] = H[i][0] * v1[0] + H[i][1] * v1[1] + H[i][2] * v1[2]; nn += u1[i] * u1[i]; }
    nn = 1.0 / sqrt(nn); for (int i = 0; i < 3; i++) u1[i] *= nn;
    double dot = 0; nn = 0;
    for (int i = 0; i < 3; i++) { u2[i] = H[i][0] * v2[0] + H[i][1] * v2[1] + H[i][2] * v2[2]; dot += u2[i] * u1[i]; }
    for (int i = 0; i < 3; i++) { u2[i] -= dot * u1[i]; nn += u2[i] * u2[i]; }
    nn = 1.0 / sqrt(nn); for (int i = 0; i < 3; i++) u2[i] *= nn;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    double R[3][3], P[3][3], T[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R[i][j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j]; sR[i][j] = R[i][j]; }
    double tr = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += R[k][i] * H[k][j]; P[i][j] = s; }
    for (int i = 0; i < 3; i++) tr += P[i][i];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i][j] = (i == j ? tr : 0.0) - 0.5 * (P[i][j] + P[j][i]);
    // inverse of the 3x3 matrix T (adjugate)
    const double det = T[0][0] * (T[1][1] * T[2][2] - T[1][2] * T[2][1]) - T[0][1] * (T[1][0] * T[2][2] - T[1][2] * T[2][0]) +
                       T[0][2] * (T[1][0] * T[2][1] - T[1][1] * T[2][0]);
    const double id = 1.0 / det;
    sTi[0][0] = (T[1][1] * T[2][2] - T[1][2] * T[2][1]) * id; sTi[0][1] = (T[0][2] * T[2][1] - T[0][1] * T[2][2]) * id; sTi[0][2] = (T[0][1] * T[1][2] - T[0][2] * T[1][1]) * id;
    sTi[1][0] = (T[1][2] * T[2][0] - T[1][0] * T[2][2]) * id; sTi[1][1] = (T[0][0] * T[2][2] - T[0][2] * T[2][0]) * id; sTi[1][2] = (T[0][2] * T[1][0] - T[0][0] * T[1][2]) * id;
    sTi[2][0] = (T[1][0] * T[2][1] - T[1][1] * T[2][0]) * id; sTi[2][1] = (T[0][1] * T[2][0] - T[0][0] * T[2][1]) * id; sTi[2][2] = (T[0][0] * T[1][1] - T[0][1] * T[1][0]) * id;
  }
  __syncthreads();
  // aligned coordinates
  for (int o = tid; o < d_pad; o += 64) {
    float v = 0.f;
    if (o < d) { const int a = o / 3, j = o % 3; v = (float)(sp[a][0] * sR[0][j] + sp[a][1] * sR[1][j] + sp[a][2] * sR[2][j]); }
    Xa[n * d_pad + o] = v;
  }
  if (!metric) return;
  // Jacobian columns: thread = raw coordinate (b, i)
  for (int in = tid; in < d; in += 64) {
    const int b = in / 3, i = in % 3, s = ssel[b];
    double W[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};       // R [omega]_x
    if (s >= 0) {
      // A = R^T dH with dH = e_i^T qc_s  ->  A[r][c] = R[i][r] * qc[c];  k = axial(A - A^T)
      double A[3][3];
      for (int r = 0; r < 3; r++) for (int c2 = 0; c2 < 3; c2++) A[r][c2] = sR[i][r] * sQc[s][c2];
      const double k0 = A[2][1] - A[1][2], k1 = A[0][2] - A[2][0], k2 = A[1][0] - A[0][1];
      const double w0 = sTi[0][0] * k0 + sTi[0][1] * k1 + sTi[0][2] * k2, w1 = sTi[1][0] * k0 + sTi[1][1] * k1 + sTi[1][2] * k2,
                   w2 = sTi[2][0] * k0 + sTi[2][1] * k1 + sTi[2][2] * k2;
      const double S[3][3] = {{0, -w2, w1}, {w2, 0, -w0}, {-w1, w0, 0}};
      for (int r = 0; r < 3; r++) for (int c2 = 0; c2 < 3; c2++) W[r][c2] = sR[r][0] * S[0][c2] + sR[r][1] * S[1][c2] + sR[r][2] * S[2][c2];
    }
    for (int a = 0; a < n_atoms; a++)
      for (int j = 0; j < 3; j++) {
        double v = (a == b) ? sR[i][j] : 0.0;
        if (s >= 0) v += -sR[i][j] / ref.m + sp[a][0] * W[0][j] + sp[a][1] * W[1][j] + sp[a][2] * W[2][j];
        sJ[3 * a + j][in] = v;
      }
  }
  __syncthreads();
  for (int e = tid; e < d * d; e += 64) {
    const int o = e / d, o2 = e % d;
    double s = 0;
    for (int in = 0; in < d; in++) s += (double)diag[in] * sJ[o][in] * sJ[o2][in];
    metric[n * (int64_t)d * d + e] = (float)s;
  }
}

}  // namespace epde
