// Micro-benchmark: exp2 throughput per SM — MUFU.EX2 (f32), packed bf16x2 / f16x2 ex2, and an FMA-pipe polynomial emulation.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

constexpr int ITERS = 4096, UNROLL = 8;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// 2^x for x <= 0 on the FMA/ALU pipes: round-to-nearest split + degree-3 minimax on [-0.5, 0.5] + exponent add
__device__ __forceinline__ float ex2poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;            // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void k(float* out, float seed) {
  float v[UNROLL];
  uint32_t u[UNROLL];
  for (int i = 0; i < UNROLL; ++i) { v[i] = -seed * (threadIdx.x % 7 + i) * 0.01f; u[i] = 0xbc00bc00u + i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (MODE == 0) v[i] = ex2f(v[i]) - 1.0f;
      else if (MODE == 1) u[i] = ex2bf2(u[i]) ^ 0x80008000u;
      else if (MODE == 3) u[i] = ex2h2(u[i]) ^ 0x80008000u;
      else v[i] = ex2poly(v[i]) - 1.0f;
    }
  }
  float s = 0.f;
  for (int i = 0; i < UNROLL; ++i) s += v[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char* name, int per_op) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 256>>>(out, 1.f);
  cudaEventRecord(e0);
  k<MODE><<<sms * 8, 256>>>(out, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * 8 * 256 * ITERS * UNROLL * per_op;
  printf("%-28s %8.3f ms  %7.2f Gexp/s  %6.2f exp/clk/SM (at max clock %d MHz)\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("poly3 on FMA/ALU pipes", 1);
  run<3>("ex2.approx.f16x2", 2);
  return 0;
}
