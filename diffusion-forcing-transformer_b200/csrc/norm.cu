// K1 — fused AdaLN modulate + LayerNorm with per-frame shift/scale gather.
// One warp per token row; the row stays in registers between the statistics and the
// modulate pass, so HBM traffic is exactly: read x (4 B/elem) + write y (4 and/or 2 B/elem).
#include "common.cuh"

namespace dfot {

constexpr int kNormWarps = 4;

// (register caps measured on B200, M = 10240, D = 1152: 9 blocks/SM (56 registers) 23.2 us vs 23.5 uncapped, 10 blocks (48,
// spills) 27.6 — the 10 k-row launch is bound by its ramp, not by occupancy: 5.0 TB/s against 6.25 at 8x the rows)
template <int NV>  // NV = float4 vectors held per lane (row length D <= NV*128)
__global__ void __launch_bounds__(kNormWarps * 32)
adaln_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ mod, int64_t mod_ld,
                       int64_t shift_col, int64_t scale_col, float* __restrict__ y_f32,
                       __nv_bfloat16* __restrict__ y_bf16, float2* __restrict__ stats, int64_t M, int D,
                       int64_t tokens_per_frame, float eps) {
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
  if (stats != nullptr && lane == 0) stats[m] = make_float2(mean, rstd);   // for the GATE_LNRESID GEMM epilogue
  const int64_t f = m / tokens_per_frame;
  const float4* sh = reinterpret_cast<const float4*>(mod + f * mod_ld + shift_col);
  const float4* sc = reinterpret_cast<const float4*>(mod + f * mod_ld + scale_col);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const float4 s = __ldg(sh + c), g = __ldg(sc + c);
      float4 y;
      y.x = adaln_value(v[i].x, mean, rstd, g.x, s.x);
      y.y = adaln_value(v[i].y, mean, rstd, g.y, s.y);
      y.z = adaln_value(v[i].z, mean, rstd, g.z, s.z);
      y.w = adaln_value(v[i].w, mean, rstd, g.w, s.w);
      if (y_f32)
        st_stream_u4(y_f32 + m * D + 4 * c, make_uint4(__float_as_uint(y.x), __float_as_uint(y.y),
                                                       __float_as_uint(y.z), __float_as_uint(y.w)));
      if (y_bf16) st_stream_u2(y_bf16 + m * D + 4 * c, make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w)));
    }
  }
}


// Latency-regime variant (a few hundred rows — small-batch sampling): one BLOCK per row, one float4 column group per thread
// (two above D = 1024), so every load a row needs is in flight at once and a row costs one memory round trip plus two block
// reductions (warp shuffles + a fixed-order sum over the warps: deterministic), instead of one warp walking NV vectors.
// It is also the consumer of a split-K GEMM (dfot_gemm_bf16_splitk) at the end of a DiT block half, fused with the AdaLN
// that follows it (parts != NULL):
//   x[m, :] = resid[m, :] + gate[f(m), :] * (sum_s parts[m, s*D + :] + bias)          (dit_blocks.py:504-509, gated residual)
//   y[m, :] = LN_eps(x[m, :]) * (1 + scale[f(m), :]) + shift[f(m), :]                  (the next AdaLayerNorm[Zero])
// with the partial sums added in split order.  x itself is only written on request: in the reference's blocks the residual
// base of the next half is the MODULATED tensor y (quirk Q1), so x has no other reader.  parts == NULL: x = resid (plain K1).
constexpr int kRowThreads = 256;
template <int NV>
__global__ void __launch_bounds__(kRowThreads)
row_adaln_kernel(const float* __restrict__ parts, int splits, const float* __restrict__ bias,
                 const float* __restrict__ resid, const float* __restrict__ mod, int64_t mod_ld, int64_t gate_col,
                 int64_t shift_col, int64_t scale_col, float* __restrict__ x_out, float* __restrict__ y_f32,
                 __nv_bfloat16* __restrict__ y_bf16, int D, int64_t tokens_per_frame, float eps) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float red[2][kRowThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = blockIdx.x;
  const int nvec = D >> 2;
  const int64_t f = m / tokens_per_frame;
  const bool norm = shift_col >= 0;
  const float4* pr = reinterpret_cast<const float4*>(parts + m * (int64_t)splits * D);
  const float4* rr = reinterpret_cast<const float4*>(resid + m * D);
  const float4* md = reinterpret_cast<const float4*>(mod + f * mod_ld);
  float4 v[NV], s4[NV], g4[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nvec) {
      const float4 r = rr[c];
      if (norm) { s4[i] = __ldg(md + (shift_col >> 2) + c); g4[i] = __ldg(md + (scale_col >> 2) + c); }
      if (parts != nullptr) {
        float4 acc = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 g = __ldg(md + (gate_col >> 2) + c);
#pragma unroll 8
        for (int sp = 0; sp < splits; ++sp) {
          const float4 t = pr[sp * nvec + c];
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        v[i] = make_float4(fmaf(g.x, acc.x, r.x), fmaf(g.y, acc.y, r.y), fmaf(g.z, acc.z, r.z), fmaf(g.w, acc.w, r.w));
        if (x_out) *reinterpret_cast<float4*>(x_out + m * D + 4 * c) = v[i];
      } else {
        v[i] = r;
      }
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  if (!norm) return;
  auto block_sum = [&](float t, int slot) {
    t = warp_sum(t);
    if (lane == 0) red[slot][warp] = t;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kRowThreads / 32; ++w) tot += red[slot][w];
    return tot;
  };
  const float mean = block_sum(sum, 0) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (threadIdx.x + i * kRowThreads < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(block_sum(sq, 1) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nvec) {
      float4 y;
      y.x = (v[i].x - mean) * rstd * (1.f + g4[i].x) + s4[i].x;
      y.y = (v[i].y - mean) * rstd * (1.f + g4[i].y) + s4[i].y;
      y.z = (v[i].z - mean) * rstd * (1.f + g4[i].z) + s4[i].z;
      y.w = (v[i].w - mean) * rstd * (1.f + g4[i].w) + s4[i].w;
      if (y_f32) *reinterpret_cast<float4*>(y_f32 + m * D + 4 * c) = y;
      if (y_bf16) *reinterpret_cast<uint2*>(y_bf16 + m * D + 4 * c) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
    }
  }
}

constexpr int64_t kRowKernelMaxRows = 2048;   // up to here K1 runs one block per row in latency mode (see row_adaln_kernel)

}  // namespace dfot

extern "C" int dfot_adaln_layernorm(const float* x, const float* mod, int64_t mod_ld, int64_t shift_col,
                                    int64_t scale_col, float* y_f32, void* y_bf16, int64_t M, int64_t D,
                                    int64_t tokens_per_frame, float eps, void* stream) {
  return dfot_adaln_layernorm_stats(x, mod, mod_ld, shift_col, scale_col, y_f32, y_bf16, nullptr, M, D, tokens_per_frame,
                                    eps, stream);
}

extern "C" int dfot_adaln_layernorm_stats(const float* x, const float* mod, int64_t mod_ld, int64_t shift_col,
                                          int64_t scale_col, float* y_f32, void* y_bf16, float* stats, int64_t M, int64_t D,
                                          int64_t tokens_per_frame, float eps, void* stream) {
  using namespace dfot;
  DFOT_REQUIRE(x && mod && (y_f32 || y_bf16), DFOT_ERR_INVALID_ARG, "adaln_layernorm: null pointer");
  DFOT_REQUIRE((uintptr_t)stats % 8 == 0, DFOT_ERR_UNSUPPORTED, "adaln_layernorm: stats must be 8-byte aligned");
  DFOT_REQUIRE(M > 0 && D > 0 && tokens_per_frame > 0, DFOT_ERR_INVALID_ARG, "adaln_layernorm: bad sizes");
  DFOT_REQUIRE(D % 4 == 0 && mod_ld % 4 == 0 && shift_col % 4 == 0 && scale_col % 4 == 0, DFOT_ERR_UNSUPPORTED,
               "adaln_layernorm: D, mod_ld and column offsets must be multiples of 4 (128-bit access)");
  DFOT_REQUIRE(D <= 4096, DFOT_ERR_UNSUPPORTED, "adaln_layernorm: D=%lld > 4096 unsupported", (long long)D);
  cudaStream_t s = (cudaStream_t)stream;
  if (stats == nullptr && latency_mode() && M <= kRowKernelMaxRows && D <= 2048 && (uintptr_t)x % 16 == 0 && (uintptr_t)mod % 16 == 0 && (uintptr_t)y_f32 % 16 == 0 &&
      (uintptr_t)y_bf16 % 8 == 0) {
    if (D <= 1024)
      launch_pdl(row_adaln_kernel<1>, dim3((unsigned)M), dim3(kRowThreads), 0, s, (const float*)nullptr, 0,
                 (const float*)nullptr, x, mod, mod_ld, (int64_t)0, shift_col, scale_col, (float*)nullptr, y_f32,
                 (__nv_bfloat16*)y_bf16, (int)D, tokens_per_frame, eps);
    else
      launch_pdl(row_adaln_kernel<2>, dim3((unsigned)M), dim3(kRowThreads), 0, s, (const float*)nullptr, 0,
                 (const float*)nullptr, x, mod, mod_ld, (int64_t)0, shift_col, scale_col, (float*)nullptr, y_f32,
                 (__nv_bfloat16*)y_bf16, (int)D, tokens_per_frame, eps);
    DFOT_CHECK_LAUNCH("adaln_layernorm");
    return DFOT_OK;
  }
  const unsigned grid = (unsigned)ceil_div(M, kNormWarps);
#define LAUNCH(NV)                                                                                             \
  launch_pdl(adaln_layernorm_kernel<NV>, dim3(grid), dim3(kNormWarps * 32), 0, s, x, mod, mod_ld, shift_col, scale_col, y_f32,     \
                                                              (__nv_bfloat16*)y_bf16, (float2*)stats, M, (int)D, \
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

extern "C" int dfot_splitk_gate_resid_adaln(const float* parts, int64_t splits, const float* bias, const float* resid,
                                            const float* mod, int64_t mod_ld, int64_t gate_col, int64_t shift_col,
                                            int64_t scale_col, float* x_out, float* y_f32, void* y_bf16, int64_t M,
                                            int64_t D, int64_t tokens_per_frame, float eps, void* stream) {
  using namespace dfot;
  DFOT_REQUIRE(parts && resid && mod && splits >= 1 && splits <= 64, DFOT_ERR_INVALID_ARG,
               "splitk_gate_resid_adaln: null pointer or bad split count");
  DFOT_REQUIRE(M > 0 && D > 0 && tokens_per_frame > 0 && gate_col >= 0, DFOT_ERR_INVALID_ARG,
               "splitk_gate_resid_adaln: bad sizes");
  const bool norm = shift_col >= 0;
  DFOT_REQUIRE(norm ? (scale_col >= 0 && (y_f32 || y_bf16)) : (x_out != nullptr), DFOT_ERR_INVALID_ARG,
               "splitk_gate_resid_adaln: the norm needs scale_col and an output; without a norm x_out is the output");
  DFOT_REQUIRE(D % 4 == 0 && mod_ld % 4 == 0 && gate_col % 4 == 0 && (!norm || (shift_col % 4 == 0 && scale_col % 4 == 0)),
               DFOT_ERR_UNSUPPORTED, "splitk_gate_resid_adaln: D, mod_ld and column offsets must be multiples of 4");
  DFOT_REQUIRE(((uintptr_t)parts | (uintptr_t)resid | (uintptr_t)mod | (uintptr_t)bias | (uintptr_t)x_out |
                (uintptr_t)y_f32) % 16 == 0 && (uintptr_t)y_bf16 % 8 == 0,
               DFOT_ERR_UNSUPPORTED, "splitk_gate_resid_adaln: buffers must be 16-byte aligned");
  DFOT_REQUIRE(D <= 2048 && M < (1ll << 31), DFOT_ERR_UNSUPPORTED, "splitk_gate_resid_adaln: D=%lld > 2048 unsupported",
               (long long)D);
  cudaStream_t s = (cudaStream_t)stream;
  if (D <= 1024)
    launch_pdl(row_adaln_kernel<1>, dim3((unsigned)M), dim3(kRowThreads), 0, s, parts, (int)splits, bias, resid, mod, mod_ld,
               gate_col, shift_col, scale_col, x_out, y_f32, (__nv_bfloat16*)y_bf16, (int)D, tokens_per_frame, eps);
  else
    launch_pdl(row_adaln_kernel<2>, dim3((unsigned)M), dim3(kRowThreads), 0, s, parts, (int)splits, bias, resid, mod, mod_ld,
               gate_col, shift_col, scale_col, x_out, y_f32, (__nv_bfloat16*)y_bf16, (int)D, tokens_per_frame, eps);
  DFOT_CHECK_LAUNCH("splitk_gate_resid_adaln");
  return DFOT_OK;
}
