// Glue kernels of the DiT3D backbone: noise-level features, embedding combine,
// patchify (im2col for the patch-embed GEMM), unpatchify, fp32->bf16 cast.
#include "common.cuh"

namespace dfot {

// embeddings.py:112-153 (flip_sin_to_cos → [cos | sin], max_period 1e4) and :94-109 (Fourier).
__global__ void noise_features_kernel(const void* __restrict__ levels, int levels_dtype,
                                      const float* __restrict__ freqs, const float* __restrict__ phases,
                                      __nv_bfloat16* __restrict__ out, int64_t n, int dim) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * dim) return;
  const int64_t r = idx / dim;
  const int c = (int)(idx % dim);
  const float k = levels_dtype == DFOT_I64 ? (float)reinterpret_cast<const int64_t*>(levels)[r]
                                           : reinterpret_cast<const float*>(levels)[r];
  float v;
  if (freqs != nullptr) {
    v = cosf(k * freqs[c] + phases[c]) * 1.4142135623730951f;
  } else {
    const int half = dim >> 1;
    const int i = c < half ? c : c - half;
    const float f = expf(-9.210340371976184f * (float)i / (float)half);
    v = c < half ? cosf(k * f) : sinf(k * f);
  }
  out[idx] = __float2bfloat16_rn(v);
}

__global__ void silu_sum_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                     const uint8_t* __restrict__ row_mask, int64_t rows_per_mask,
                                     __nv_bfloat16* __restrict__ out, int64_t n_rows, int64_t D) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rows * D) return;
  float v = a[idx];
  if (b != nullptr) {
    const int64_t r = idx / D;
    const bool masked = row_mask != nullptr && row_mask[r / rows_per_mask] != 0;
    if (!masked) v += b[idx];
  }
  out[idx] = __float2bfloat16_rn(silu_f(v));
}

// out[token, c*p*p + py*p + px] = x[frame, c, gy*p + py, gx*p + px]; token = (frame, gy, gx)
template <typename TX>
__global__ void patchify_kernel(const TX* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t ld,
                                int64_t frames, int C, int H, int W, int p) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int gw = W / p, gh = H / p, kk = C * p * p;
  const int64_t total = frames * gh * gw * kk;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % kk);
  const int64_t tok = idx / kk;
  const int gx = (int)(tok % gw), gy = (int)((tok / gw) % gh);
  const int64_t fr = tok / ((int64_t)gw * gh);
  const int px = k % p, py = (k / p) % p, c = k / (p * p);
  const float v = (float)x[((fr * C + c) * H + gy * p + py) * W + gx * p + px];
  out[tok * ld + k] = __float2bfloat16_rn(v);
}

// x[frame, c, gy*p + py, gx*p + px] = tok[token, (py*p + px)*C + c]
template <typename TX>
__global__ void unpatchify_kernel(const float* __restrict__ tok, int64_t ld, TX* __restrict__ x, int64_t frames,
                                  int C, int H, int W, int p) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int64_t total = frames * C * H * W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = (int)(idx % W), h = (int)((idx / W) % H), c = (int)((idx / ((int64_t)W * H)) % C);
  const int64_t fr = idx / ((int64_t)W * H * C);
  const int gw = W / p, gh = H / p;
  const int64_t token = (fr * gh + h / p) * gw + w / p;
  const int col = ((h % p) * p + (w % p)) * C + c;
  x[idx] = (TX)tok[token * ld + col];
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
// 8 elements per thread: 2 x 16-byte streaming loads, one 16-byte store (pointers 16-byte aligned, n % 8 == 0)
__global__ void cast_bf16_vec8_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n8) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 a = ld_stream_u4(in + 8 * i), b = ld_stream_u4(in + 8 * i + 4);
  st_stream_u4(out + 8 * i, make_uint4(pack_bf16x2(__uint_as_float(a.x), __uint_as_float(a.y)),
                                       pack_bf16x2(__uint_as_float(a.z), __uint_as_float(a.w)),
                                       pack_bf16x2(__uint_as_float(b.x), __uint_as_float(b.y)),
                                       pack_bf16x2(__uint_as_float(b.z), __uint_as_float(b.w))));
}

// ---------------------------------------------------------------- matrix attention (dit_blocks.py:211-350, MatrixAttention)
// Frames are the attention tokens of the matrix variants: the P patch rows of a frame are contracted with the columns of
// `qkv_u` BEFORE the row projection (einsum 'nm,blnd,dk->blmk' with the u factor first: P times fewer GEMM rows).
//   out[((r*Mc + c)*L + l), d] = sum_n u[n*Mc + c] * y[((r*L + l)*P + n), d]        (bf16, the A operand of the QKV GEMM)
// A work item = (frame, 128-column slab): 8 warps stride over the patches (four rows = 2 KB per warp in flight), float4 per
// lane, fixed-order reduction in shared memory (deterministic).  Column heads c are handled four at a time (y re-read from
// L2 when Mc > 4).
__global__ void __launch_bounds__(256)
patch_mix_bf16_kernel(const float* __restrict__ y, const float* __restrict__ u, __nv_bfloat16* __restrict__ out,
                      int n_frames, int L, int P, int Mc, int D) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ float4 part[8][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slabs = (D + 127) >> 7;
  // persistent over (frame, slab) items, slab fastest: the blocks running at any moment read whole rows between them
  for (int item = blockIdx.x; item < n_frames * n_slabs; item += gridDim.x) {
    const int frame = item / n_slabs, r = frame / L, l = frame % L;
    const int d = (item % n_slabs) * 128 + lane * 4;
    const bool live = d < D;
    const float* yf = y + (int64_t)frame * P * D + d;
    for (int c0 = 0; c0 < Mc; c0 += 4) {
      float4 acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        for (int n0 = warp; n0 < P; n0 += 32) {           // rows n0, n0 + 8, n0 + 16, n0 + 24 of this warp
          uint4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            v[k] = n0 + 8 * k < P ? ld_stream_u4(yf + (int64_t)(n0 + 8 * k) * D) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int n = n0 + 8 * k;
            if (n >= P) break;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float w = (c0 + j < Mc) ? __ldg(u + n * Mc + c0 + j) : 0.f;
              acc[j].x = fmaf(w, __uint_as_float(v[k].x), acc[j].x);
              acc[j].y = fmaf(w, __uint_as_float(v[k].y), acc[j].y);
              acc[j].z = fmaf(w, __uint_as_float(v[k].z), acc[j].z);
              acc[j].w = fmaf(w, __uint_as_float(v[k].w), acc[j].w);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) part[warp][j][lane] = acc[j];
      __syncthreads();
      if (warp < 4 && c0 + warp < Mc && live) {          // warp j finishes column head c0 + j
        float4 s = part[0][warp][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
          const float4 t = part[w][warp][lane];
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        const int64_t row = ((int64_t)r * Mc + c0 + warp) * L + l;
        *reinterpret_cast<uint2*>(out + row * D + d) = make_uint2(pack_bf16x2(s.x, s.y), pack_bf16x2(s.z, s.w));
      }
      __syncthreads();
    }
  }
}

// The way back (the `proj_u` factor of the output projection, the AdaLN-Zero gate and the block's residual in one pass):
//   x[((r*L + l)*P + n), d] = y[same] + gate[(r*L + l)*ld_gate + d] * (sum_c pu[c*P + n] * z[((r*Mc + c)*L + l), d] + pb[n*D + d])
// (y == NULL: no residual, gate == NULL: gate 1 — the bare attention output a MatrixCrossDiTBlock attends to.)
// One block per (frame, 32-patch chunk); a thread owns one float4 column group (two when D > 1024): the frame's gate and —
// with one column head, every shipped configuration — its z row stay in registers, the patch rows stream through four at a
// time.  No index arithmetic per element.
template <bool ONE_COL>
__global__ void __launch_bounds__(256)
patch_expand_gate_resid_kernel(float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                               const float* __restrict__ pu, const float* __restrict__ pb,
                               const float* __restrict__ gate, int64_t ld_gate, int L, int P, int Mc, int D) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int frame = blockIdx.x, r = frame / L, l = frame % L;
  const int n_begin = blockIdx.y * 32, n_end = min(P, n_begin + 32);
  for (int d = threadIdx.x * 4; d < D; d += 1024) {
    const float4 g = gate != nullptr ? __ldg(reinterpret_cast<const float4*>(gate + (int64_t)frame * ld_gate + d))
                                     : make_float4(1.f, 1.f, 1.f, 1.f);
    const float* zf = z + ((int64_t)r * Mc * L + l) * D + d;          // column head c: + c * L * D
    float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ONE_COL) z0 = __ldg(reinterpret_cast<const float4*>(zf));
    const int64_t row0 = (int64_t)frame * P;
    for (int n0 = n_begin; n0 < n_end; n0 += 4) {
      uint4 yv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        yv[k] = (y != nullptr && n0 + k < n_end) ? ld_stream_u4(y + (row0 + n0 + k) * D + d) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int n = n0 + k;
        if (n >= n_end) break;
        float4 s = pb != nullptr ? __ldg(reinterpret_cast<const float4*>(pb + (int64_t)n * D + d))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        if (ONE_COL) {
          const float w = __ldg(pu + n);
          s.x = fmaf(w, z0.x, s.x); s.y = fmaf(w, z0.y, s.y); s.z = fmaf(w, z0.z, s.z); s.w = fmaf(w, z0.w, s.w);
        } else {
          for (int c = 0; c < Mc; ++c) {
            const float w = __ldg(pu + c * P + n);
            const float4 t = __ldg(reinterpret_cast<const float4*>(zf + (int64_t)c * L * D));
            s.x = fmaf(w, t.x, s.x); s.y = fmaf(w, t.y, s.y); s.z = fmaf(w, t.z, s.z); s.w = fmaf(w, t.w, s.w);
          }
        }
        st_stream_u4(x + (row0 + n) * D + d,
                     make_uint4(__float_as_uint(fmaf(g.x, s.x, __uint_as_float(yv[k].x))),
                                __float_as_uint(fmaf(g.y, s.y, __uint_as_float(yv[k].y))),
                                __float_as_uint(fmaf(g.z, s.z, __uint_as_float(yv[k].z))),
                                __float_as_uint(fmaf(g.w, s.w, __uint_as_float(yv[k].w)))));
      }
    }
  }
}

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)ceil_div(n, threads); }

}  // namespace dfot

using namespace dfot;

extern "C" int dfot_noise_features(const void* levels, int levels_dtype, const float* fourier_freqs,
                                   const float* fourier_phases, void* out_bf16, int64_t n, int64_t dim,
                                   void* stream) {
  DFOT_REQUIRE(levels && out_bf16 && n > 0 && dim > 0 && dim % 2 == 0, DFOT_ERR_INVALID_ARG,
               "noise_features: bad arguments");
  DFOT_REQUIRE(levels_dtype == DFOT_I64 || levels_dtype == DFOT_F32, DFOT_ERR_INVALID_ARG,
               "noise_features: levels must be int64 or f32");
  DFOT_REQUIRE((fourier_freqs == nullptr) == (fourier_phases == nullptr), DFOT_ERR_INVALID_ARG,
               "noise_features: freqs and phases go together");
  launch_pdl(noise_features_kernel, dim3(blocks_for(n * dim, 256)), dim3(256), 0, (cudaStream_t)stream, 
      levels, levels_dtype, fourier_freqs, fourier_phases, (__nv_bfloat16*)out_bf16, n, (int)dim);
  DFOT_CHECK_LAUNCH("noise_features");
  return DFOT_OK;
}

extern "C" int dfot_silu_sum_bf16(const float* a, const float* b, const uint8_t* row_mask, int64_t rows_per_mask,
                                  void* out_bf16, int64_t n_rows, int64_t D, void* stream) {
  DFOT_REQUIRE(a && out_bf16 && n_rows > 0 && D > 0, DFOT_ERR_INVALID_ARG, "silu_sum_bf16: bad arguments");
  DFOT_REQUIRE(row_mask == nullptr || rows_per_mask > 0, DFOT_ERR_INVALID_ARG, "silu_sum_bf16: rows_per_mask");
  launch_pdl(silu_sum_bf16_kernel, dim3(blocks_for(n_rows * D, 256)), dim3(256), 0, (cudaStream_t)stream, 
      a, b, row_mask, rows_per_mask, (__nv_bfloat16*)out_bf16, n_rows, D);
  DFOT_CHECK_LAUNCH("silu_sum_bf16");
  return DFOT_OK;
}

extern "C" int dfot_patchify_bf16(const void* x, int x_dtype, void* out_bf16, int64_t ld, int64_t frames, int64_t C,
                                  int64_t H, int64_t W, int64_t p, void* stream) {
  DFOT_REQUIRE(x && out_bf16 && frames > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0 && ld >= C * p * p,
               DFOT_ERR_INVALID_ARG, "patchify: bad arguments");
  const int64_t total = frames * C * H * W;
  if (x_dtype == DFOT_F32)
    launch_pdl(patchify_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const float*)x, (__nv_bfloat16*)out_bf16, ld, frames, (int)C, (int)H, (int)W, (int)p);
  else if (x_dtype == DFOT_BF16)
    launch_pdl(patchify_kernel<__nv_bfloat16>, dim3(blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)x, (__nv_bfloat16*)out_bf16, ld, frames, (int)C, (int)H, (int)W, (int)p);
  else
    DFOT_REQUIRE(false, DFOT_ERR_INVALID_ARG, "patchify: x dtype must be f32 or bf16");
  DFOT_CHECK_LAUNCH("patchify");
  return DFOT_OK;
}

extern "C" int dfot_unpatchify(const float* tok, int64_t ld, void* x, int x_dtype, int64_t frames, int64_t C,
                               int64_t H, int64_t W, int64_t p, void* stream) {
  DFOT_REQUIRE(tok && x && frames > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0 && ld >= C * p * p,
               DFOT_ERR_INVALID_ARG, "unpatchify: bad arguments");
  const int64_t total = frames * C * H * W;
  if (x_dtype == DFOT_F32)
    launch_pdl(unpatchify_kernel<float>, dim3(blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
        tok, ld, (float*)x, frames, (int)C, (int)H, (int)W, (int)p);
  else if (x_dtype == DFOT_BF16)
    launch_pdl(unpatchify_kernel<__nv_bfloat16>, dim3(blocks_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
        tok, ld, (__nv_bfloat16*)x, frames, (int)C, (int)H, (int)W, (int)p);
  else
    DFOT_REQUIRE(false, DFOT_ERR_INVALID_ARG, "unpatchify: x dtype must be f32 or bf16");
  DFOT_CHECK_LAUNCH("unpatchify");
  return DFOT_OK;
}

extern "C" int dfot_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream) {
  DFOT_REQUIRE(in && out_bf16 && n > 0, DFOT_ERR_INVALID_ARG, "cast_bf16: bad arguments");
  if (n % 8 == 0 && (uintptr_t)in % 16 == 0 && (uintptr_t)out_bf16 % 16 == 0)
    launch_pdl(cast_bf16_vec8_kernel, dim3(blocks_for(n / 8, 256)), dim3(256), 0, (cudaStream_t)stream, in, (__nv_bfloat16*)out_bf16, n / 8);
  else
    launch_pdl(cast_bf16_kernel, dim3(blocks_for(n, 256)), dim3(256), 0, (cudaStream_t)stream, in, (__nv_bfloat16*)out_bf16, n);
  DFOT_CHECK_LAUNCH("cast_bf16");
  return DFOT_OK;
}

extern "C" int dfot_patch_mix_bf16(const float* y, const float* u, void* out_bf16, int64_t R, int64_t L, int64_t P,
                                   int64_t Mc, int64_t D, void* stream) {
  DFOT_REQUIRE(y && u && out_bf16 && R > 0 && L > 0 && P > 0 && Mc > 0 && D > 0, DFOT_ERR_INVALID_ARG,
               "patch_mix_bf16: bad arguments");
  DFOT_REQUIRE(D % 4 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)out_bf16 % 8 == 0, DFOT_ERR_INVALID_ARG,
               "patch_mix_bf16: D must be a multiple of 4 and the buffers 16-byte aligned");
  DFOT_REQUIRE(R * L * ceil_div(D, 128) < (1ll << 31), DFOT_ERR_INVALID_ARG, "patch_mix_bf16: too many frames");
  const int64_t items = R * L * ceil_div(D, 128);
  launch_pdl(patch_mix_bf16_kernel, dim3((unsigned)(items < 148 * 8 ? items : 148 * 8)), dim3(256), 0, (cudaStream_t)stream,
             y, u, (__nv_bfloat16*)out_bf16, (int)(R * L), (int)L, (int)P, (int)Mc, (int)D);
  DFOT_CHECK_LAUNCH("patch_mix_bf16");
  return DFOT_OK;
}

extern "C" int dfot_patch_expand_gate_resid(float* x, const float* y, const float* z, const float* pu, const float* pb,
                                            const float* gate, int64_t ld_gate, int64_t R, int64_t L, int64_t P,
                                            int64_t Mc, int64_t D, void* stream) {
  DFOT_REQUIRE(x && z && pu && x != y && R > 0 && L > 0 && P > 0 && Mc > 0 && D > 0, DFOT_ERR_INVALID_ARG,
               "patch_expand_gate_resid: bad arguments (x must not alias y)");
  DFOT_REQUIRE(D % 4 == 0 && (gate == nullptr || (ld_gate % 4 == 0 && ld_gate >= D)), DFOT_ERR_INVALID_ARG,
               "patch_expand_gate_resid: D and ld_gate must be multiples of 4, ld_gate >= D");
  DFOT_REQUIRE(((uintptr_t)x | (uintptr_t)y | (uintptr_t)z | (uintptr_t)gate | (uintptr_t)pb) % 16 == 0, DFOT_ERR_INVALID_ARG,
               "patch_expand_gate_resid: buffers must be 16-byte aligned");
  DFOT_REQUIRE(R * L < (1ll << 31) && P < (1ll << 20), DFOT_ERR_INVALID_ARG, "patch_expand_gate_resid: too many frames");
  const dim3 grid((unsigned)(R * L), (unsigned)ceil_div(P, 32));
  if (Mc == 1)
    launch_pdl(patch_expand_gate_resid_kernel<true>, grid, dim3(256), 0, (cudaStream_t)stream, x, y, z, pu, pb, gate,
               ld_gate, (int)L, (int)P, (int)Mc, (int)D);
  else
    launch_pdl(patch_expand_gate_resid_kernel<false>, grid, dim3(256), 0, (cudaStream_t)stream, x, y, z, pu, pb, gate,
               ld_gate, (int)L, (int)P, (int)Mc, (int)D);
  DFOT_CHECK_LAUNCH("patch_expand_gate_resid");
  return DFOT_OK;
}
