// Microbenchmark: issue rate of scalar FFMA (3 register operands) vs packed fma.rn.f32x2 on sm_100a, per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fma_rate.bin tools/micro/fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float s) {
  float a[16], w[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = s + i;
  unsigned long long pk[8], pw[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(pk[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
#pragma unroll
  for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(pw[i]) : "f"(w[2 * i]), "f"(w[2 * i + 1]));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w[(i + r) & 7], w[(i + r + 3) & 7]);
      } else if (MODE == 2) {
        // operands stay packed in 64-bit registers across iterations: the pure FFMA2 issue rate
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pk[i]) : "l"(pw[(i + r) & 3]), "l"(pw[(i + r + 1) & 3]));
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          unsigned long long d, x, y, z;
          asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(w[(i + r) & 7]), "f"(w[(i + r + 1) & 7]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(w[(i + r + 3) & 7]), "f"(w[(i + r + 4) & 7]));
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
        }
      }
    }
  }
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += a[i];
  if (MODE == 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pk[i])); t += lo + hi; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 2, 512>>>(out, iters, 1.0001f); else if (mode == 1) k<1><<<148 * 2, 512>>>(out, iters, 1.0001f);
      else k<2><<<148 * 2, 512>>>(out, iters, 1.0001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 2 * 512 * iters * 8.0 * 16.0;
      if (rep) printf("%s: %.3f ms, %.2f TFLOP/s fp32 (%.1f FMA/clk/SM at 1.9 GHz)\n", mode == 0 ? "FFMA scalar" : (mode == 1 ? "fma.rn.f32x2 (repacked every time)" : "fma.rn.f32x2 (packed operands)"), ms, 2 * fma / ms / 1e9,
                      fma / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
