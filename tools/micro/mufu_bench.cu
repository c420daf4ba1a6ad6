// Microbenchmark: throughput of ex2 variants per SM (results/clk/SM).  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + threadIdx.x * 1e-3f + i * 0.01f;
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = 0x38003800u + i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 2) {  // polynomial 2^x on the FMA pipe: floor via magic add, degree-3 poly, exponent add
        float xf = x[i];
        float fl = floorf(xf);
        float fr = xf - fl;
        float p = fmaf(fr, 0.0555054f, 0.2402265f);
        p = fmaf(p, fr, 0.6931472f);
        p = fmaf(p, fr, 1.0f);
        int e = (int)fl;
        x[i] = __int_as_float(__float_as_int(p) + (e << 23)) * 1e-3f - 1.0f;
      } else if (MODE == 3) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 4) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_op) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int iters = 4096;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148 * 4, 512>>>(out, 16, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<148 * 4, 512>>>(out, iters, 0.5f);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = 148.0 * 4 * 512 * 16.0 * iters * per_op;
  double per_clk_sm = ops / (ms * 1e-3) / 148.0 / (clk * 1e3);
  printf("%-28s %.3f ms  %.1f Gop/s  %.2f results/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e6, per_clk_sm, clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 2);
  run<3>("ex2.approx.ftz.bf16x2", 2);
  run<2>("poly3 2^x (FMA pipe)", 1);
  run<4>("tanh.approx.f32", 1);
  return 0;
}
