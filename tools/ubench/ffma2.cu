// micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 issue throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, float a, float b, int iters) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void k_packed(float* out, float a, float b, int iters) {
  unsigned long long A, B, x0, x1, x2, x3;
  asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
  float t = threadIdx.x;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "f"(t), "f"(t + 1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x1) : "f"(t + 2), "f"(t + 3));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(t + 4), "f"(t + 5));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x3) : "f"(t + 6), "f"(t + 7));
  for (int i = 0; i < iters; ++i) {
    fma2(x0, A, B); fma2(x1, A, B); fma2(x2, A, B); fma2(x3, A, B);
  }
  float lo, hi, s = 0;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x0)); s += lo + hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x1)); s += lo + hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x2)); s += lo + hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x3)); s += lo + hi;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 100000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int v = 0; v < 2; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (v == 0) k_scalar<<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters); else k_packed<<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("%s: %.3f ms, %.2f TFLOP/s fp32 (8 fma per thread per iteration)\n", v ? "fma.rn.f32x2" : "scalar fmaf", ms,
                      2.0 * 8 * iters * 148 * 8 * 256 / ms / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
