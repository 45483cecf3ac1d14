// K4 — fused per-frame sampler step + history-guidance combine + next-step prepare.
// One pass over HBM per sampling step; see include/dfot_b200.h for the contract.
// HBM-bound: algorithmic bytes per latent element = 4 (x_t) + 4 (x_{t+1}) + nfe*s_out + nfe*s_in.
#include "common.cuh"

namespace dfot {

constexpr int kSamplerThreads = 256;
constexpr int kMaxNfe = 16;

template <typename T> struct Vec4;  // 4 elements of T
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) {
    uint4 u = ld_stream_u4(p);
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
  }
  static __device__ __forceinline__ void store(float* p, float4 v) {
    st_stream_u4(p, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    uint2 u = ld_stream_u2(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    st_stream_u2(p, make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)));
  }
};

__device__ __forceinline__ float4 clamp4(float4 v, float c) {
  return make_float4(fminf(fmaxf(v.x, -c), c), fminf(fmaxf(v.y, -c), c), fminf(fmaxf(v.z, -c), c),
                     fminf(fmaxf(v.w, -c), c));
}

// grid = (chunks, T, B); each thread owns 4 consecutive elements per iteration.
// NFE > 0: the branch count is a compile-time constant, so every stream of an element (x_t, the NFE model outputs, the
// DDIM / history / excluded-token noise) is loaded before the first use — one DRAM round trip per element with
// (2 + 2*NFE) x 16 bytes in flight per thread instead of one dependent load after another.  NFE == 0: generic loop.
template <typename TOut, typename TIn, int NFE>
__global__ void __launch_bounds__(kSamplerThreads)
sampler_step_hg_kernel(float* __restrict__ x, const TOut* __restrict__ model_out, TIn* __restrict__ model_in_next,
                       const dfot_frame_update* __restrict__ upd, const dfot_frame_prepare* __restrict__ prep,
                       const float* __restrict__ noise_ddim, const float* __restrict__ noise_hist,
                       const float* __restrict__ noise_excl, int nfe_rt, int T, int64_t F) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int nfe = NFE > 0 ? NFE : nfe_rt;
  const int t = blockIdx.y, b = blockIdx.z;
  __shared__ dfot_frame_update s_upd[kMaxNfe];
  __shared__ dfot_frame_prepare s_prep[kMaxNfe];
  if (threadIdx.x < nfe) {
    const int64_t r = (int64_t)b * nfe + threadIdx.x;
    if (upd) s_upd[threadIdx.x] = upd[r * T + t];
    if (prep) s_prep[threadIdx.x] = prep[r * T + t];
  }
  __syncthreads();
  const bool do_update = (model_out != nullptr) && (upd != nullptr) && s_upd[0].generate != 0;
  const int64_t frame_x = ((int64_t)b * T + t) * F;
  const int64_t row_stride = (int64_t)T * F;                    // between the branches of a sample
  const int64_t frame_r = ((int64_t)b * nfe * T + t) * F;       // branch 0 of this (sample, frame)
  for (int64_t e = ((int64_t)blockIdx.x * kSamplerThreads + threadIdx.x) * 4; e < F;
       e += (int64_t)gridDim.x * kSamplerThreads * 4) {
    // x is updated in place: plain (coherent) load, not the .nc streaming path
    float4 xv = *reinterpret_cast<const float4*>(x + frame_x + e);
    if constexpr (NFE > 0) {
      float4 o[NFE], nd[NFE], np[NFE];
      if (do_update) {
#pragma unroll
        for (int j = 0; j < NFE; ++j) {
          if (s_upd[j].w == 0.f) continue;
          if (s_upd[j].b != 0.f) o[j] = Vec4<TOut>::load(model_out + frame_r + j * row_stride + e);
          if (s_upd[j].sigma != 0.f && noise_ddim != nullptr) nd[j] = Vec4<float>::load(noise_ddim + frame_r + j * row_stride + e);
        }
      }
      if (model_in_next != nullptr) {
#pragma unroll
        for (int j = 0; j < NFE; ++j) {
          const int mode = s_prep[j].mode;
          if (mode == 1) np[j] = Vec4<float>::load(noise_hist + ((int64_t)s_prep[j].noise_row * T + t) * F + e);
          else if (mode == 2) np[j] = Vec4<float>::load(noise_excl + frame_r + j * row_stride + e);
        }
      }
      if (do_update) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NFE; ++j) {
          const dfot_frame_update u = s_upd[j];
          if (u.w == 0.f) continue;
          float4 ov = o[j];
          if (u.clip > 0.f) ov = clamp4(ov, u.clip);
          // b == 0 marks a frame whose level does not change (a = 1): the reference keeps x through torch.where, so a
          // NaN / Inf in the model output of such a frame must not reach x (0 * NaN = NaN)
          float4 v = u.b == 0.f ? make_float4(u.a * xv.x, u.a * xv.y, u.a * xv.z, u.a * xv.w)
                                : make_float4(u.a * xv.x + u.b * ov.x, u.a * xv.y + u.b * ov.y, u.a * xv.z + u.b * ov.z,
                                              u.a * xv.w + u.b * ov.w);
          if (u.sigma != 0.f && noise_ddim != nullptr) {
            v.x += u.sigma * nd[j].x; v.y += u.sigma * nd[j].y; v.z += u.sigma * nd[j].z; v.w += u.sigma * nd[j].w;
          }
          acc.x += u.w * v.x; acc.y += u.w * v.y; acc.z += u.w * v.z; acc.w += u.w * v.w;
        }
        xv = acc;
        Vec4<float>::store(x + frame_x + e, xv);
      }
      if (model_in_next != nullptr) {
#pragma unroll
        for (int j = 0; j < NFE; ++j) {
          const dfot_frame_prepare p = s_prep[j];
          float4 v = xv;
          if (p.mode == 1)
            v = make_float4(p.qa * xv.x + p.qb * np[j].x, p.qa * xv.y + p.qb * np[j].y, p.qa * xv.z + p.qb * np[j].z,
                            p.qa * xv.w + p.qb * np[j].w);
          else if (p.mode == 2)
            v = np[j];
          Vec4<TIn>::store(model_in_next + frame_r + j * row_stride + e, v);
        }
      }
    } else {
      if (do_update) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < nfe; ++j) {
          const dfot_frame_update u = s_upd[j];
          if (u.w == 0.f) continue;
          const int64_t off = frame_r + j * row_stride + e;
          float4 v = make_float4(u.a * xv.x, u.a * xv.y, u.a * xv.z, u.a * xv.w);
          if (u.b != 0.f) {   // (b == 0: kept frame, the model output is not read — see above)
            float4 o = Vec4<TOut>::load(model_out + off);
            if (u.clip > 0.f) o = clamp4(o, u.clip);
            v.x += u.b * o.x; v.y += u.b * o.y; v.z += u.b * o.z; v.w += u.b * o.w;
          }
          if (u.sigma != 0.f && noise_ddim != nullptr) {
            float4 n = Vec4<float>::load(noise_ddim + off);
            v.x += u.sigma * n.x; v.y += u.sigma * n.y; v.z += u.sigma * n.z; v.w += u.sigma * n.w;
          }
          acc.x += u.w * v.x; acc.y += u.w * v.y; acc.z += u.w * v.z; acc.w += u.w * v.w;
        }
        xv = acc;
        Vec4<float>::store(x + frame_x + e, xv);
      }
      if (model_in_next != nullptr) {
        for (int j = 0; j < nfe; ++j) {
          const dfot_frame_prepare p = s_prep[j];
          const int64_t off = frame_r + j * row_stride + e;
          float4 v = xv;
          if (p.mode == 1) {
            float4 n = Vec4<float>::load(noise_hist + ((int64_t)p.noise_row * T + t) * F + e);
            v = make_float4(p.qa * xv.x + p.qb * n.x, p.qa * xv.y + p.qb * n.y, p.qa * xv.z + p.qb * n.z,
                            p.qa * xv.w + p.qb * n.w);
          } else if (p.mode == 2) {
            v = Vec4<float>::load(noise_excl + off);
          }
          Vec4<TIn>::store(model_in_next + off, v);
        }
      }
    }
  }
}

}  // namespace dfot

extern "C" int dfot_sampler_step_hg(float* x, const void* model_out, int model_out_dtype, void* model_in_next,
                                    int model_in_dtype, const dfot_frame_update* upd,
                                    const dfot_frame_prepare* prep, const float* noise_ddim,
                                    const float* noise_hist, const float* noise_excl, int64_t B, int64_t nfe,
                                    int64_t T, int64_t F, void* stream) {
  using namespace dfot;
  DFOT_REQUIRE(x != nullptr && B > 0 && T > 0 && F > 0, DFOT_ERR_INVALID_ARG, "sampler_step_hg: bad x/B/T/F");
  DFOT_REQUIRE(nfe >= 1 && nfe <= kMaxNfe, DFOT_ERR_UNSUPPORTED, "sampler_step_hg: nfe=%lld not in [1,%d]",
               (long long)nfe, kMaxNfe);
  DFOT_REQUIRE(F % 4 == 0, DFOT_ERR_UNSUPPORTED, "sampler_step_hg: frame size %lld must be a multiple of 4",
               (long long)F);
  DFOT_REQUIRE(T <= 65535 && B <= 65535, DFOT_ERR_UNSUPPORTED, "sampler_step_hg: B/T exceed grid limits");
  DFOT_REQUIRE(model_out == nullptr || upd != nullptr, DFOT_ERR_INVALID_ARG, "sampler_step_hg: upd is required");
  DFOT_REQUIRE(model_in_next == nullptr || prep != nullptr, DFOT_ERR_INVALID_ARG,
               "sampler_step_hg: prep is required");
  int64_t chunks = ceil_div(F, (int64_t)kSamplerThreads * 4);
  // keep >= 2 waves of 148 SMs when the problem is large enough, otherwise one block per 1024 elements
  const int64_t frames = B * T;
  const int64_t cap = ceil_div(148 * 8, frames);
  if (chunks > cap) chunks = cap < 1 ? 1 : cap;
  dim3 grid((unsigned)chunks, (unsigned)T, (unsigned)B), block(kSamplerThreads);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_N(TO, TI, N)                                                                                       \
  launch_pdl(sampler_step_hg_kernel<TO, TI, N>, dim3(grid), dim3(block), 0, s, x, (const TO*)model_out, (TI*)model_in_next, upd, prep,  \
                                                           noise_ddim, noise_hist, noise_excl, (int)nfe, (int)T, F)
#define LAUNCH(TO, TI)                                                                                            \
  do {                                                                                                            \
    switch (nfe) {                                                                                                \
      case 1: LAUNCH_N(TO, TI, 1); break;                                                                         \
      case 2: LAUNCH_N(TO, TI, 2); break;                                                                         \
      case 3: LAUNCH_N(TO, TI, 3); break;                                                                         \
      case 4: LAUNCH_N(TO, TI, 4); break;                                                                         \
      default: LAUNCH_N(TO, TI, 0);                                                                               \
    }                                                                                                             \
  } while (0)
  const bool ob = model_out_dtype == DFOT_BF16, ib = model_in_dtype == DFOT_BF16;
  DFOT_REQUIRE((model_out_dtype == DFOT_F32 || ob) && (model_in_dtype == DFOT_F32 || ib), DFOT_ERR_INVALID_ARG,
               "sampler_step_hg: dtype tags must be DFOT_F32 or DFOT_BF16");
  if (ob && ib) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (ob) LAUNCH(__nv_bfloat16, float);
  else if (ib) LAUNCH(float, __nv_bfloat16);
  else LAUNCH(float, float);
#undef LAUNCH_N
#undef LAUNCH
  DFOT_CHECK_LAUNCH("sampler_step_hg");
  return DFOT_OK;
}
