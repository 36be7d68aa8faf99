// Legacy warp-level mma.sync.m16n8k8 TF32 rate on B200 (for the 3xTF32 weight-gradient products).
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NACC>
__global__ void k(float* out, int reps) {
  float c[NACC][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; i++) a[i] = __float_as_uint(1.0f + threadIdx.x * 0.001f + i);
  for (int i = 0; i < 2; i++) b[i] = __float_as_uint(0.5f + i);
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) c[j][i] = 0.f;
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) mma_tf32(c[j], a, b);
  }
  float s = 0; for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  float* d; CK(cudaMalloc(&d, 148 * 1024 * 4 * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int nt : {128, 256, 512}) {
    int reps = 20000;
    k<12><<<p.multiProcessorCount, nt>>>(d, 10); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k<12><<<p.multiProcessorCount, nt>>>(d, reps); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)reps * 12 * (nt / 32) * p.multiProcessorCount;
    printf("{\"variant\": \"mma_sync_m16n8k8_tf32_nt%d\", \"ms\": %.3f, \"mma_per_us_per_sm\": %.1f, \"tf32_tflops\": %.1f}\n", nt, ms,
           mmas / (ms * 1e3) / p.multiProcessorCount, mmas * 2048.0 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
