// FP32 pipe microbenchmark for B200 (sm_100a): measures the denominators the
// EigenPDE step kernel is judged against.  MEASURED_PEAKS.json carries no FP32
// figure, so this program records
//   * scalar FFMA and packed FFMA2 (fma.rn.f32x2) register-only peak,
//   * the same dense-layer inner loop with weights delivered by broadcast
//     LDS.128 from shared memory, and by the uniform datapath (LDCU from
//     __constant__), for 1 and 2 sample pairs per thread,
//   * the accurate-tanh (ex2 + rcp) cost.
// Output: one JSON line per variant on stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ffma_peak ffma_peak.cu
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

constexpr int NL = 23;              // distinct 20x20 matrices cycled through (23*1600 B = 36.8 KB)
__constant__ __align__(16) float c_w[NL * 400];

enum { SRC_REG = 0, SRC_LDS = 1, SRC_LDC = 2 };

// One "layer": out[o] = sum_i W[i][o] * in[i], o,i < 20, for NP sample pairs.
template <int SRC, int NP>
__global__ void __launch_bounds__(256, 1)
k_layer(const float* __restrict__ gw, float* __restrict__ out, int reps, int nl, long long* cyc) {
  extern __shared__ float4 s_w4[];
  float* s_w = reinterpret_cast<float*>(s_w4);
  if (SRC == SRC_LDS) {
    for (int i = threadIdx.x; i < nl * 400; i += blockDim.x) s_w[i] = gw[i];
  }
  __syncthreads();
  float2 in[NP][20];
#pragma unroll
  for (int p = 0; p < NP; p++)
#pragma unroll
    for (int i = 0; i < 20; i++) in[p][i] = make_float2(1.0f + 0.001f * threadIdx.x, 1.0f - 0.001f * i);
  float4 wr[5];
#pragma unroll
  for (int q = 0; q < 5; q++) wr[q] = reinterpret_cast<const float4*>(gw)[q];
  long long t0 = clock64();
  for (int r = 0; r < reps; r += nl)
  for (int l = 0; l < nl; l++) {
    float2 acc[NP][20];
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int o = 0; o < 20; o++) acc[p][o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 20; i++) {
#pragma unroll
      for (int q = 0; q < 5; q++) {
        float4 w;
        if (SRC == SRC_REG) w = wr[q];
        else if (SRC == SRC_LDS) w = s_w4[l * 100 + i * 5 + q];
        else w = reinterpret_cast<const float4*>(c_w)[l * 100 + i * 5 + q];
#pragma unroll
        for (int p = 0; p < NP; p++) {
          acc[p][4 * q + 0] = ffma2(make_float2(w.x, w.x), in[p][i], acc[p][4 * q + 0]);
          acc[p][4 * q + 1] = ffma2(make_float2(w.y, w.y), in[p][i], acc[p][4 * q + 1]);
          acc[p][4 * q + 2] = ffma2(make_float2(w.z, w.z), in[p][i], acc[p][4 * q + 2]);
          acc[p][4 * q + 3] = ffma2(make_float2(w.w, w.w), in[p][i], acc[p][4 * q + 3]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
      for (int o = 0; o < 20; o++) in[p][o] = acc[p][o];
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < NP; p++)
#pragma unroll
    for (int o = 0; o < 20; o++) s += in[p][o].x + in[p][o].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// scalar FFMA, register only: 40 independent accumulators
__global__ void __launch_bounds__(256, 1)
k_ffma_scalar(float* __restrict__ out, int reps, long long* cyc, float a, float b) {
  float acc[40];
#pragma unroll
  for (int i = 0; i < 40; i++) acc[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int j = 0; j < 10; j++)
#pragma unroll
      for (int i = 0; i < 40; i++) acc[i] = fmaf(acc[i], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 40; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// packed FFMA2, register only: 40 independent pair accumulators
__global__ void __launch_bounds__(256, 1)
k_ffma2_reg(float* __restrict__ out, int reps, long long* cyc, float a, float b) {
  float2 acc[40];
#pragma unroll
  for (int i = 0; i < 40; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, 1.0f);
  float2 aa = make_float2(a, a * 0.999f), bb = make_float2(b, b * 1.001f);
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int j = 0; j < 10; j++)
#pragma unroll
      for (int i = 0; i < 40; i++) acc[i] = ffma2(acc[i], aa, bb);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 40; i++) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// tanh via ex2.approx + rcp.approx: t = 1 - 2/(exp(2x)+1)
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
template <int MODE>
__global__ void __launch_bounds__(256, 1)
k_tanh(float* __restrict__ out, int reps, long long* cyc) {
  float v[20];
#pragma unroll
  for (int i = 0; i < 20; i++) v[i] = 0.01f * threadIdx.x + 0.1f * i - 1.0f;
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int i = 0; i < 20; i++) v[i] = (MODE == 0) ? tanh_fast(v[i]) : tanhf(v[i]);
#pragma unroll
    for (int i = 0; i < 20; i++) v[i] = v[i] * 1.5f + 0.1f;
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 20; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ms; double cyc_med; };

template <class F>
Res timeit(F launch, long long* d_cyc, int nblk) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
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
  double s = 0; for (auto c : h) s += (double)c; s /= nblk;
  return {best, s};
}

int main() {
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  int nsm = prop.multiProcessorCount;
  int nblk = nsm;  // one 256-thread CTA per SM (8 warps/SM, 2 per SMSP) like the step kernel
  std::vector<float> hw(NL * 400);
  for (size_t i = 0; i < hw.size(); i++) hw[i] = 0.05f + 1e-4f * (float)((i * 7919) % 13 - 6);
  float *d_w, *d_out; long long* d_cyc;
  CK(cudaMalloc(&d_w, hw.size() * 4)); CK(cudaMemcpy(d_w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpyToSymbol(c_w, hw.data(), hw.size() * 4));
  CK(cudaMalloc(&d_out, (size_t)nblk * 4 * 1024 * 4)); CK(cudaMalloc(&d_cyc, nblk * 4 * sizeof(long long)));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d}\n", prop.name, nsm, prop.clockRate);

  auto report = [&](const char* name, Res r, double fma_per_thread, int threads_per_sm, int nb) {
    double fma_total = fma_per_thread * threads_per_sm * (double)nb;
    double tflops = 2.0 * fma_total / (r.ms * 1e-3) / 1e12;
    double fma_per_clk_sm = fma_per_thread * threads_per_sm / r.cyc_med * ((double)nb / nsm);
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"fma_per_clk_per_sm\": %.2f, \"cycles\": %.0f, \"mhz_eff\": %.0f}\n",
           name, r.ms, tflops, fma_per_clk_sm, r.cyc_med, r.cyc_med / (r.ms * 1e-3) / 1e6);
    fflush(stdout);
  };

  const int reps = 4000;
  for (int occ = 1; occ <= 2; occ++) {
    int nb = nblk * occ; char nm[64];
    snprintf(nm, 64, "ffma_scalar_reg_occ%d", occ);
    report(nm, timeit([&] { k_ffma_scalar<<<nb, 256>>>(d_out, reps, d_cyc, 0.999f, 0.001f); }, d_cyc, nb), 400.0 * reps, 256, nb);
    snprintf(nm, 64, "ffma2_reg_occ%d", occ);
    report(nm, timeit([&] { k_ffma2_reg<<<nb, 256>>>(d_out, reps, d_cyc, 0.999f, 0.001f); }, d_cyc, nb), 800.0 * reps, 256, nb);
  }
  const int lreps = 2000;
  size_t smem = NL * 400 * 4;
  CK(cudaFuncSetAttribute(k_layer<SRC_LDS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_layer<SRC_LDS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int nl : {1, 6, 23}) {
    char nm[64];
    snprintf(nm, 64, "layer_reg_np1_nl%d", nl);
    report(nm, timeit([&] { k_layer<SRC_REG, 1><<<nblk, 256, 0>>>(d_w, d_out, lreps, nl, d_cyc); }, d_cyc, nblk), 800.0 * lreps, 256, nblk);
    snprintf(nm, 64, "layer_reg_np2_nl%d", nl);
    report(nm, timeit([&] { k_layer<SRC_REG, 2><<<nblk, 256, 0>>>(d_w, d_out, lreps, nl, d_cyc); }, d_cyc, nblk), 1600.0 * lreps, 256, nblk);
    snprintf(nm, 64, "layer_lds_np1_nl%d", nl);
    report(nm, timeit([&] { k_layer<SRC_LDS, 1><<<nblk, 256, smem>>>(d_w, d_out, lreps, nl, d_cyc); }, d_cyc, nblk), 800.0 * lreps, 256, nblk);
    snprintf(nm, 64, "layer_lds_np2_nl%d", nl);
    report(nm, timeit([&] { k_layer<SRC_LDS, 2><<<nblk, 256, smem>>>(d_w, d_out, lreps, nl, d_cyc); }, d_cyc, nblk), 1600.0 * lreps, 256, nblk);
    snprintf(nm, 64, "layer_ldc_np1_nl%d", nl);
    report(nm, timeit([&] { k_layer<SRC_LDC, 1><<<nblk, 256, 0>>>(d_w, d_out, lreps, nl, d_cyc); }, d_cyc, nblk), 800.0 * lreps, 256, nblk);
    snprintf(nm, 64, "layer_ldc_np2_nl%d", nl);
    report(nm, timeit([&] { k_layer<SRC_LDC, 2><<<nblk, 256, 0>>>(d_w, d_out, lreps, nl, d_cyc); }, d_cyc, nblk), 1600.0 * lreps, 256, nblk);
  }
  // 2 CTAs/SM for the uniform path (more warps to hide constant-cache latency)
  {
    int nb = nblk * 2;
    report("layer_ldc_np1_nl23_occ2", timeit([&] { k_layer<SRC_LDC, 1><<<nb, 256, 0>>>(d_w, d_out, lreps, 23, d_cyc); }, d_cyc, nb), 800.0 * lreps, 256, nb);
    report("layer_lds_np1_nl23_occ2", timeit([&] { k_layer<SRC_LDS, 1><<<nb, 256, smem>>>(d_w, d_out, lreps, 23, d_cyc); }, d_cyc, nb), 800.0 * lreps, 256, nb);
  }
  // tanh: count "ops" as tanh evaluations (reported in the tflops column as 2*evals/s/1e12)
  report("tanh_fast_20", timeit([&] { k_tanh<0><<<nblk, 256>>>(d_out, 20000, d_cyc); }, d_cyc, nblk), 20.0 * 20000, 256, nblk);
  report("tanhf_20", timeit([&] { k_tanh<1><<<nblk, 256>>>(d_out, 20000, d_cyc); }, d_cyc, nblk), 20.0 * 20000, 256, nblk);
  return 0;
}
