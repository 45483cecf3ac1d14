// Glue kernels of the VAE-decode row (reference: algorithms/vae/video_vae/model.py Decoder,
// algorithms/vae/common/modules/{updownsample,attention}.py).  Activations are channel-last clips with a padded frame
// axis:  [B, kPad + T, H, W, C],  kPad = 2 leading slots per clip that hold copies of the clip's first frame in the bf16
// conv inputs (the causal window of dfot_conv3d_causal_bf16) and are unused in the fp32 residual stream.  HBM-bound.
#include "common.cuh"

namespace dfot {
namespace vae {

constexpr int kThreads = 256;
constexpr int kPad = 2;

__device__ __forceinline__ void lin_coord(int dst, int size_in, int& i0, int& i1, float& w1) {
  // torch upsample, align_corners=False, scale 2: src = max(0, (dst + 0.5) / 2 - 0.5)
  const float src = fmaxf(0.f, ((float)dst + 0.5f) * 0.5f - 0.5f);
  i0 = (int)src;
  i1 = i0 + (i0 < size_in - 1 ? 1 : 0);
  w1 = src - (float)i0;
}

__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// x2 spatial (and, with TIME, x2 temporal on all frames but the first) upsampling, fp32 clip -> bf16 clip with pads.
//   TIME = false: nearest x2 in (H, W)                               (SpatialUpsample2x, updownsample.py:73-80)
//   TIME = true : frame 0 bilinear x2; frames 1.. trilinear x(2,2,2) (Spatial2xTime2x3DUpsample, updownsample.py:131-147)
// One thread = 8 channels of one output pixel; output slots 0..pad-1 of a clip replicate its first frame (pad = kPad for
// video clips, 0 for plain image batches).
template <bool TIME>
__global__ void __launch_bounds__(kThreads)
upsample2x_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Tin, int H, int W, int C,
                  int pad) {
  pdl_trigger();
  pdl_wait();
  const int vecs = C >> 3, Ho = 2 * H, Wo = 2 * W;
  const int Tout = TIME ? 2 * Tin - 1 : Tin;
  const int64_t total = (int64_t)B * (pad + Tout) * Ho * Wo * vecs;
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % vecs);
  int64_t r = idx / vecs;
  const int xo = (int)(r % Wo); r /= Wo;
  const int yo = (int)(r % Ho); r /= Ho;
  const int slot = (int)(r % (pad + Tout));
  const int b = (int)(r / (pad + Tout));
  const int to = slot < pad ? 0 : slot - pad;                      // pads replicate output frame 0
  const float* clip = in + ((int64_t)b * (pad + Tin) + pad) * H * W * C + 8 * v;
  auto px = [&](int t, int y, int x) { return clip + (((int64_t)t * H + y) * W + x) * C; };
  float acc[8];
  if constexpr (!TIME) {
    load8f(px(to, yo >> 1, xo >> 1), acc);
  } else {
    int y0, y1, x0, x1, t0 = 0, t1 = 0;
    float wy, wx, wt = 0.f;
    lin_coord(yo, H, y0, y1, wy);
    lin_coord(xo, W, x0, x1, wx);
    if (to > 0) {                                                    // frames 1.. of the input, interpolated among themselves
      lin_coord(to - 1, Tin - 1, t0, t1, wt);
      t0 += 1; t1 += 1;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float wk = k == 0 ? 1.f - wt : wt;
      if (wk == 0.f) continue;
      const int t = k == 0 ? t0 : t1;
      float a[8], bq[8], c[8], d[8];
      load8f(px(t, y0, x0), a); load8f(px(t, y0, x1), bq); load8f(px(t, y1, x0), c); load8f(px(t, y1, x1), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float top = a[j] + wx * (bq[j] - a[j]), bot = c[j] + wx * (d[j] - c[j]);
        acc[j] += wk * (top + wy * (bot - top));
      }
    }
  }
  __nv_bfloat16* dst = out + ((((int64_t)b * (pad + Tout) + slot) * Ho + yo) * Wo + xo) * C + 8 * v;
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                              pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
}

// pad slots of a bf16 clip <- its first frame (after a kernel that wrote the valid frames only)
__global__ void __launch_bounds__(kThreads)
fill_pad_frames_kernel(__nv_bfloat16* __restrict__ x, int B, int T, int64_t frame_vec8) {
  pdl_trigger();
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= (int64_t)B * frame_vec8) return;
  const int b = (int)(idx / frame_vec8);
  const int64_t o = idx - (int64_t)b * frame_vec8;
  uint4* clip = reinterpret_cast<uint4*>(x) + (int64_t)b * (kPad + T) * frame_vec8;
  const uint4 v = clip[kPad * frame_vec8 + o];
#pragma unroll
  for (int p = 0; p < kPad; ++p) clip[p * frame_vec8 + o] = v;
}

// row softmax of fp32 logits (pre-scaled by `scale`) -> bf16 probabilities; one warp per row, n <= 1024, n % 4 == 0
__global__ void __launch_bounds__(kThreads)
softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int64_t rows, int n, int64_t ld_s,
                    int64_t ld_p, float scale) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(s + row * ld_s);
  float4 v[8];
  float mx = -INFINITY;
  const int nv = n >> 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      v[i] = __ldg(src + c);
      v[i].x *= scale; v[i].y *= scale; v[i].z *= scale; v[i].w *= scale;
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx); v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  uint2* dst = reinterpret_cast<uint2*>(p + row * ld_p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) dst[c] = make_uint2(pack_bf16x2(v[i].x * inv, v[i].y * inv), pack_bf16x2(v[i].z * inv, v[i].w * inv));
  }
}

}  // namespace vae
}  // namespace dfot

using namespace dfot;

extern "C" int dfot_vae_upsample2x_bf16(const float* in, void* out_bf16, int64_t B, int64_t T_in, int64_t H, int64_t W,
                                        int64_t C, int temporal, void* stream) {
  DFOT_REQUIRE(in && out_bf16 && B > 0 && T_in > 0 && H > 0 && W > 0 && C > 0, DFOT_ERR_INVALID_ARG,
               "vae_upsample2x: bad arguments");
  DFOT_REQUIRE(C % 8 == 0 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out_bf16 % 16 == 0), DFOT_ERR_UNSUPPORTED,
               "vae_upsample2x: C %% 8 == 0 and 16-byte aligned pointers required");
  const int64_t T_out = temporal ? 2 * T_in - 1 : T_in;
  const int64_t total = B * (vae::kPad + T_out) * 4 * H * W * (C / 8);
  DFOT_REQUIRE(total < (1ll << 40), DFOT_ERR_UNSUPPORTED, "vae_upsample2x: problem too large");
  const dim3 grid((unsigned)ceil_div(total, vae::kThreads));
  if (temporal)
    launch_pdl(vae::upsample2x_kernel<true>, grid, dim3(vae::kThreads), 0, (cudaStream_t)stream, in,
               (__nv_bfloat16*)out_bf16, (int)B, (int)T_in, (int)H, (int)W, (int)C, vae::kPad);
  else
    launch_pdl(vae::upsample2x_kernel<false>, grid, dim3(vae::kThreads), 0, (cudaStream_t)stream, in,
               (__nv_bfloat16*)out_bf16, (int)B, (int)T_in, (int)H, (int)W, (int)C, vae::kPad);
  DFOT_CHECK_LAUNCH("vae_upsample2x");
  return DFOT_OK;
}

extern "C" int dfot_upsample2x_nearest_bf16(const float* in, void* out_bf16, int64_t n_img, int64_t H, int64_t W, int64_t C,
                                            void* stream) {
  DFOT_REQUIRE(in && out_bf16 && n_img > 0 && H > 0 && W > 0 && C > 0, DFOT_ERR_INVALID_ARG,
               "upsample2x_nearest: bad arguments");
  DFOT_REQUIRE(C % 8 == 0 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out_bf16 % 16 == 0), DFOT_ERR_UNSUPPORTED,
               "upsample2x_nearest: C %% 8 == 0 and 16-byte aligned pointers required");
  const int64_t total = n_img * 4 * H * W * (C / 8);
  DFOT_REQUIRE(total < (1ll << 40) && n_img < (1ll << 31), DFOT_ERR_UNSUPPORTED, "upsample2x_nearest: problem too large");
  launch_pdl(vae::upsample2x_kernel<false>, dim3((unsigned)ceil_div(total, vae::kThreads)), dim3(vae::kThreads), 0,
             (cudaStream_t)stream, in, (__nv_bfloat16*)out_bf16, (int)n_img, 1, (int)H, (int)W, (int)C, 0);
  DFOT_CHECK_LAUNCH("upsample2x_nearest");
  return DFOT_OK;
}

extern "C" int dfot_vae_fill_pad_frames(void* x_bf16, int64_t B, int64_t T, int64_t frame_elems, void* stream) {
  DFOT_REQUIRE(x_bf16 && B > 0 && T > 0 && frame_elems > 0, DFOT_ERR_INVALID_ARG, "vae_fill_pad_frames: bad arguments");
  DFOT_REQUIRE(frame_elems % 8 == 0 && ((uintptr_t)x_bf16 % 16 == 0), DFOT_ERR_UNSUPPORTED,
               "vae_fill_pad_frames: frame size must be a multiple of 8 elements, pointer 16-byte aligned");
  const int64_t fv = frame_elems / 8;
  launch_pdl(vae::fill_pad_frames_kernel, dim3((unsigned)ceil_div(B * fv, vae::kThreads)), dim3(vae::kThreads), 0,
             (cudaStream_t)stream, (__nv_bfloat16*)x_bf16, (int)B, (int)T, fv);
  DFOT_CHECK_LAUNCH("vae_fill_pad_frames");
  return DFOT_OK;
}

extern "C" int dfot_softmax_rows_bf16(const float* s, int64_t ld_s, void* p_bf16, int64_t ld_p, int64_t rows, int64_t n,
                                      float scale, void* stream) {
  DFOT_REQUIRE(s && p_bf16 && rows > 0 && n > 0, DFOT_ERR_INVALID_ARG, "softmax_rows: bad arguments");
  DFOT_REQUIRE(n % 4 == 0 && n <= 1024 && ld_s % 4 == 0 && ld_p % 4 == 0 && ld_s >= n && ld_p >= n &&
                   ((uintptr_t)s % 16 == 0) && ((uintptr_t)p_bf16 % 8 == 0),
               DFOT_ERR_UNSUPPORTED, "softmax_rows: n %% 4 == 0, n <= 1024, leading dimensions multiples of 4");
  launch_pdl(vae::softmax_rows_kernel, dim3((unsigned)ceil_div(rows, vae::kThreads / 32)), dim3(vae::kThreads), 0,
             (cudaStream_t)stream, s, (__nv_bfloat16*)p_bf16, rows, (int)n, ld_s, ld_p, scale);
  DFOT_CHECK_LAUNCH("softmax_rows");
  return DFOT_OK;
}
