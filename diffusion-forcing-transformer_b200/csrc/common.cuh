// Shared helpers for the dfot_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dfot_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL) && !defined(__CUDA_ARCH_FEAT_SM103_ALL)
#error "dfot_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace dfot {

// ---- host-side error plumbing (thread-local message, never throws) ----
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DFOT_REQUIRE(cond, code, ...)          \
  do {                                         \
    if (!(cond)) {                             \
      ::dfot::set_error(__VA_ARGS__);          \
      return (code);                           \
    }                                          \
  } while (0)

#define DFOT_CHECK_LAUNCH(name)                                                            \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      ::dfot::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));           \
      return DFOT_ERR_CUDA;                                                                \
    }                                                                                      \
    ::dfot::count_launch();                                                                \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// GroupNorm statistics workspace helpers shared by uvit.cu (stand-alone pass) and gemm_tcgen05.cu (epilogue side output):
// sums = [n_img*groups][2] f64 (sum, sum of squares) followed by [n_img*groups] float2 (mean, rstd)
int gn_zero_sums(double* sums, int64_t n_img, int64_t groups, cudaStream_t s);
int gn_finalize(double* sums, int64_t n_img, int64_t groups, int64_t count_per_group, float eps, cudaStream_t s);

// ---- device helpers ----
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// silu(x) = 0.5 x (1 + tanh(x/2)): one MUFU.TANH + 3 FP32 ops (the IEEE division above costs ~15 instructions);
// tanh.approx has ~2^-11 relative error, far below the bf16 rounding of every consumer
__device__ __forceinline__ float silu_fast_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f * x, t, 0.5f * x);
}
__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  ==  x * sigmoid(2 u)
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return x / (1.0f + __expf(-2.0f * u));
}

// 128-bit streaming loads/stores (read-once / write-once data: bypass L1 allocation)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u2(void* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dfot
