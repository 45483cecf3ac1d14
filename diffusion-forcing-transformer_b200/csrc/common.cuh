// Shared helpers for the dfot_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dfot_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL) && !defined(__CUDA_ARCH_FEAT_SM103_ALL)
#error "dfot_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace dfot {

// ---- host-side error plumbing (thread-local message, never throws) ----
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool latency_mode();   // dfot_set_latency_mode / DFOT_LATENCY_MODE (abi.cu)

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

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// With DFOT_PDL=1 every kernel of the library is launched with the programmatic-stream-serialization attribute; all
// kernels start with
//   pdl_trigger()  — the next kernel of the stream may be scheduled as soon as every CTA of this grid has started, so
//                    its CTAs become resident (and run their prologue: barrier init, TMEM allocation, tensor-map
//                    prefetch) while this grid drains, instead of after a full drain + launch latency;
//   pdl_wait()     — blocks until the PREVIOUS grid of the stream has completed and its memory is visible; no global
//                    memory is read or written before it, so ordering is exactly that of a plain stream (the chain of
//                    waits is transitive).  Without the attribute both instructions are no-ops.
// Measured on B200 inside CUDA graphs (bench.py, power-capped at ~1.65 GHz): RE10K 147.0 NFE/s with PDL vs 149.9
// without, K600 561 vs 559 — the launch gaps it removes are not what bounds a power-limited step, so it is OFF by
// default and kept as an opt-in (tests pass either way).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFOT_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// One element of AdaLayerNorm[Zero]'s output, with the rounding of every step pinned: the GATE_LNRESID GEMM epilogue rebuilds
// the value dfot_adaln_layernorm produced (its bf16 copy fed the GEMM) and must match it bit for bit.
__device__ __forceinline__ float adaln_value(float x, float mean, float rstd, float scale, float shift) {
  return __fmaf_rn(__fmul_rn(__fsub_rn(x, mean), rstd), __fadd_rn(1.f, scale), shift);
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);   // a launch error is picked up by DFOT_CHECK_LAUNCH
}

// ---------------------------------------------------------------- warp-uniform issue of tcgen05 / TMA instructions
// UTCHMMA / UTMALDG / UTCBAR take their operands from UNIFORM registers.  Issued under `if (lane == 0)` they sit in
// divergent control flow, and ptxas wraps every one of them in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop whose R2URs wait
// for the previous instruction's operand read: ~16 instructions and a serialising scoreboard per MMA (measured in the
// attention kernel: 700 cycles to issue four MMAs).  The issuing warps therefore run their loops with all 32 lanes
// (warp-uniform control flow, warp index made uniform with a shuffle as CUTLASS does) and only the instruction itself
// sits under elect.sync, which ptxas recognises.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dfot
