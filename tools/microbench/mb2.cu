// Microbenchmark 2: candidate inner loops for the step kernel at the occupancies
// the on-chip state budget allows (see DESIGN.md "kernel design notes").
//   A  : sample-pair thread, weights by broadcast LDS.128 (1 LDS : 4 FFMA2), CTA sizes 64..256
//   B  : single-sample thread, FFMA2 packed over output neurons (1 LDS : 2 FFMA2)
//   D  : outer-product role, r_o x r_i register tile, pairs over samples, operands from smem
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
constexpr int NL = 6;

template <int NT>
__global__ void __launch_bounds__(NT)
k_A(const float* __restrict__ gw, float* __restrict__ out, int reps, long long* cyc) {
  extern __shared__ float4 s_w4[];
  float* s_w = reinterpret_cast<float*>(s_w4);
  for (int i = threadIdx.x; i < NL * 400; i += NT) s_w[i] = gw[i];
  __syncthreads();
  float2 in[20];
#pragma unroll
  for (int i = 0; i < 20; i++) in[i] = make_float2(1.0f + 0.001f * threadIdx.x, 1.0f - 0.001f * i);
  long long t0 = clock64();
  for (int r = 0; r < reps; r += NL)
  for (int l = 0; l < NL; l++) {
    float2 acc[20];
#pragma unroll
    for (int o = 0; o < 20; o++) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 20; i++) {
#pragma unroll
      for (int q = 0; q < 5; q++) {
        float4 w = s_w4[l * 100 + i * 5 + q];
        acc[4 * q + 0] = ffma2(make_float2(w.x, w.x), in[i], acc[4 * q + 0]);
        acc[4 * q + 1] = ffma2(make_float2(w.y, w.y), in[i], acc[4 * q + 1]);
        acc[4 * q + 2] = ffma2(make_float2(w.z, w.z), in[i], acc[4 * q + 2]);
        acc[4 * q + 3] = ffma2(make_float2(w.w, w.w), in[i], acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int o = 0; o < 20; o++) in[o] = acc[o];
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int o = 0; o < 20; o++) s += in[o].x + in[o].y;
  out[blockIdx.x * NT + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// B: one sample per thread; acc pair = two neighbouring output neurons
template <int NT>
__global__ void __launch_bounds__(NT)
k_B(const float* __restrict__ gw, float* __restrict__ out, int reps, long long* cyc) {
  extern __shared__ float4 s_w4[];
  float* s_w = reinterpret_cast<float*>(s_w4);
  for (int i = threadIdx.x; i < NL * 400; i += NT) s_w[i] = gw[i];
  __syncthreads();
  float in[20];
#pragma unroll
  for (int i = 0; i < 20; i++) in[i] = 1.0f + 0.001f * threadIdx.x - 0.001f * i;
  long long t0 = clock64();
  for (int r = 0; r < reps; r += NL)
  for (int l = 0; l < NL; l++) {
    float2 acc[10];
#pragma unroll
    for (int o = 0; o < 10; o++) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 20; i++) {
#pragma unroll
      for (int q = 0; q < 5; q++) {
        float4 w = s_w4[l * 100 + i * 5 + q];
        acc[2 * q + 0] = ffma2(make_float2(in[i], in[i]), make_float2(w.x, w.y), acc[2 * q + 0]);
        acc[2 * q + 1] = ffma2(make_float2(in[i], in[i]), make_float2(w.z, w.w), acc[2 * q + 1]);
      }
    }
#pragma unroll
    for (int o = 0; o < 10; o++) { in[2 * o] = acc[o].x; in[2 * o + 1] = acc[o].y; }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int o = 0; o < 20; o++) s += in[o];
  out[blockIdx.x * NT + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// D: outer-product role.  P[20][TS], Q[20][TS] in smem (TS = T + 4 pad);
// unit u -> (ot = u % (20/RO), it = u / (20/RO)); thread = (unit, sample group).
template <int NT, int RO, int RI, int T>
__global__ void __launch_bounds__(NT)
k_D(float* __restrict__ out, int reps, long long* cyc) {
  constexpr int TS = T + 4;
  constexpr int NUO = 20 / RO, NUI = 20 / RI, NU = NUO * NUI;
  constexpr int NG = NT / NU;           // sample groups
  constexpr int SG = (T / NG) & ~3;     // samples per group (multiple of 4)
  extern __shared__ float4 s_w4[];
  float* P = reinterpret_cast<float*>(s_w4);
  float* Q = P + 20 * TS;
  for (int i = threadIdx.x; i < 40 * TS; i += NT) P[i] = 0.001f * (i % 97);
  __syncthreads();
  int u = threadIdx.x % NU, grp = threadIdx.x / NU;
  int ot = u % NUO, it = u / NUO;
  float2 acc[RO][RI];
#pragma unroll
  for (int a = 0; a < RO; a++)
#pragma unroll
    for (int b = 0; b < RI; b++) acc[a][b] = make_float2(0.f, 0.f);
  long long t0 = clock64();
  if (grp < NG) {
    for (int r = 0; r < reps; r++) {
      const float* p0 = P + grp * SG;
      const float* q0 = Q + grp * SG;
#pragma unroll 2
      for (int n = 0; n < SG; n += 4) {
        float4 pv[RO], qv[RI];
#pragma unroll
        for (int a = 0; a < RO; a++) pv[a] = *reinterpret_cast<const float4*>(p0 + (ot + a * NUO) * TS + n);
#pragma unroll
        for (int b = 0; b < RI; b++) qv[b] = *reinterpret_cast<const float4*>(q0 + (it + b * NUI) * TS + n);
#pragma unroll
        for (int a = 0; a < RO; a++)
#pragma unroll
          for (int b = 0; b < RI; b++) {
            acc[a][b] = ffma2(make_float2(pv[a].x, pv[a].y), make_float2(qv[b].x, qv[b].y), acc[a][b]);
            acc[a][b] = ffma2(make_float2(pv[a].z, pv[a].w), make_float2(qv[b].z, qv[b].w), acc[a][b]);
          }
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < RO; a++)
#pragma unroll
    for (int b = 0; b < RI; b++) s += acc[a][b].x + acc[a][b].y;
  out[blockIdx.x * NT + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ms; double cyc; };
template <class F>
Res timeit(F launch, long long* d_cyc, int nblk) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 5; i++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  std::vector<long long> h(nblk);
  CK(cudaMemcpy(h.data(), d_cyc, nblk * sizeof(long long), cudaMemcpyDeviceToHost));
  double s = 0; for (auto c : h) s += (double)c;
  return {best, s / nblk};
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  std::vector<float> hw(NL * 400);
  for (size_t i = 0; i < hw.size(); i++) hw[i] = 0.05f + 1e-4f * (float)((i * 7919) % 13 - 6);
  float *d_w, *d_out; long long* d_cyc;
  CK(cudaMalloc(&d_w, hw.size() * 4)); CK(cudaMemcpy(d_w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, (size_t)nsm * 8 * 1024 * 4)); CK(cudaMalloc(&d_cyc, nsm * 8 * sizeof(long long)));
  auto report = [&](const char* name, Res r, double fma_total) {
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"frac_of_73\": %.3f}\n", name, r.ms,
           2.0 * fma_total / (r.ms * 1e-3) / 1e12, 2.0 * fma_total / (r.ms * 1e-3) / 1e12 / 73.0);
    fflush(stdout);
  };
  const int reps = 3000;
  size_t smw = NL * 400 * 4;
#define RUN_A(NT, OCC) { int nb = nsm * OCC; size_t sm = (OCC == 1) ? 120 * 1024 : (200 * 1024 / OCC); \
    CK(cudaFuncSetAttribute(k_A<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    char nm[64]; snprintf(nm, 64, "A_pair_nt%d_occ%d", NT, OCC); \
    report(nm, timeit([&] { k_A<NT><<<nb, NT, sm>>>(d_w, d_out, reps, d_cyc); }, d_cyc, nb), 800.0 * reps * NT * nb); }
#define RUN_B(NT, OCC) { int nb = nsm * OCC; size_t sm = (OCC == 1) ? 120 * 1024 : (200 * 1024 / OCC); \
    CK(cudaFuncSetAttribute(k_B<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    char nm[64]; snprintf(nm, 64, "B_single_nt%d_occ%d", NT, OCC); \
    report(nm, timeit([&] { k_B<NT><<<nb, NT, sm>>>(d_w, d_out, reps, d_cyc); }, d_cyc, nb), 400.0 * reps * NT * nb); }
  (void)smw;
  RUN_A(64, 1) RUN_A(128, 1) RUN_A(160, 1) RUN_A(192, 1) RUN_A(256, 1)
  RUN_A(64, 2) RUN_A(64, 3) RUN_A(96, 2) RUN_A(128, 2) RUN_A(32, 4) RUN_A(32, 6)
  RUN_B(128, 1) RUN_B(256, 1) RUN_B(384, 1) RUN_B(512, 1) RUN_B(256, 2) RUN_B(128, 3)
#define RUN_D(NT, RO, RI, T) { int nb = nsm; size_t sm = 40 * (T + 4) * 4; \
    CK(cudaFuncSetAttribute(k_D<NT, RO, RI, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    constexpr int NU = (20 / RO) * (20 / RI); constexpr int NG = NT / NU; constexpr int SG = (T / NG) & ~3; \
    char nm[64]; snprintf(nm, 64, "D_outer_nt%d_%dx%d_T%d", NT, RO, RI, T); \
    report(nm, timeit([&] { k_D<NT, RO, RI, T><<<nb, NT, sm>>>(d_out, 2000, d_cyc); }, d_cyc, nb), \
           (double)NU * NG * SG * RO * RI * 2000.0 * nb); }
  RUN_D(128, 4, 5, 256) RUN_D(256, 4, 5, 512) RUN_D(128, 5, 5, 256) RUN_D(128, 4, 4, 256) RUN_D(128, 2, 5, 256)
  RUN_D(256, 4, 4, 256) RUN_D(256, 2, 5, 256) RUN_D(128, 5, 10, 256) RUN_D(256, 5, 10, 512)
  return 0;
}
