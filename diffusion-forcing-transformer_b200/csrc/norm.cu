// K1 — fused AdaLN modulate + LayerNorm with per-frame shift/scale gather.
// One warp per token row; the row stays in registers between the statistics and the
// modulate pass, so HBM traffic is exactly: read x (4 B/elem) + write y (4 and/or 2 B/elem).
#include "common.cuh"

namespace dfot {

constexpr int kNormWarps = 4;

template <int NV>  // NV = float4 vectors held per lane (row length D <= NV*128)
__global__ void __launch_bounds__(kNormWarps * 32)
adaln_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ mod, int64_t mod_ld,
                       int64_t shift_col, int64_t scale_col, float* __restrict__ y_f32,
                       __nv_bfloat16* __restrict__ y_bf16, int64_t M, int D, int64_t tokens_per_frame, float eps) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * kNormWarps + warp;
  if (m >= M) return;
  const int nvec = D >> 2;  // float4 per row
  const float4* xr = reinterpret_cast<const float4*>(x + m * D);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      uint4 u = ld_stream_u4(xr + c);
      v[i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
  const int64_t f = m / tokens_per_frame;
  const float4* sh = reinterpret_cast<const float4*>(mod + f * mod_ld + shift_col);
  const float4* sc = reinterpret_cast<const float4*>(mod + f * mod_ld + scale_col);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 s = __ldg(sh + c), g = __ldg(sc + c);
      float4 y;
      y.x = (v[i].x - mean) * rstd * (1.f + g.x) + s.x;
      y.y = (v[i].y - mean) * rstd * (1.f + g.y) + s.y;
      y.z = (v[i].z - mean) * rstd * (1.f + g.z) + s.z;
      y.w = (v[i].w - mean) * rstd * (1.f + g.w) + s.w;
      if (y_f32)
        st_stream_u4(y_f32 + m * D + 4 * c, make_uint4(__float_as_uint(y.x), __float_as_uint(y.y),
                                                       __float_as_uint(y.z), __float_as_uint(y.w)));
      if (y_bf16) st_stream_u2(y_bf16 + m * D + 4 * c, make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w)));
    }
  }
}

}  // namespace dfot

extern "C" int dfot_adaln_layernorm(const float* x, const float* mod, int64_t mod_ld, int64_t shift_col,
                                    int64_t scale_col, float* y_f32, void* y_bf16, int64_t M, int64_t D,
                                    int64_t tokens_per_frame, float eps, void* stream) {
  using namespace dfot;
  DFOT_REQUIRE(x && mod && (y_f32 || y_bf16), DFOT_ERR_INVALID_ARG, "adaln_layernorm: null pointer");
  DFOT_REQUIRE(M > 0 && D > 0 && tokens_per_frame > 0, DFOT_ERR_INVALID_ARG, "adaln_layernorm: bad sizes");
  DFOT_REQUIRE(D % 4 == 0 && mod_ld % 4 == 0 && shift_col % 4 == 0 && scale_col % 4 == 0, DFOT_ERR_UNSUPPORTED,
               "adaln_layernorm: D, mod_ld and column offsets must be multiples of 4 (128-bit access)");
  DFOT_REQUIRE(D <= 4096, DFOT_ERR_UNSUPPORTED, "adaln_layernorm: D=%lld > 4096 unsupported", (long long)D);
  const unsigned grid = (unsigned)ceil_div(M, kNormWarps);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(NV)                                                                                             \
  launch_pdl(adaln_layernorm_kernel<NV>, dim3(grid), dim3(kNormWarps * 32), 0, s, x, mod, mod_ld, shift_col, scale_col, y_f32,     \
                                                              (__nv_bfloat16*)y_bf16, M, (int)D,               \
                                                              tokens_per_frame, eps)
  if (D <= 256) LAUNCH(2);
  else if (D <= 512) LAUNCH(4);
  else if (D <= 768) LAUNCH(6);
  else if (D <= 1024) LAUNCH(8);
  else if (D <= 1152) LAUNCH(9);      // (DiT-XL: 36 instead of 48 data registers per lane — more rows resident per SM)
  else if (D <= 1536) LAUNCH(12);
  else if (D <= 2048) LAUNCH(16);
  else LAUNCH(32);
#undef LAUNCH
  DFOT_CHECK_LAUNCH("adaln_layernorm");
  return DFOT_OK;
}
