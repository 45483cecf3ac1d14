// Glue kernels of the VAE-decode row (reference: algorithms/vae/video_vae/model.py Decoder,
// algorithms/vae/common/modules/{updownsample,attention}.py).  Activations are channel-last clips with a padded frame
// axis:  [B, kPad + T, H, W, C],  kPad = 2 leading slots per clip that hold copies of the clip's first frame in the bf16
// conv inputs (the causal window of dfot_conv3d_causal_bf16) and are unused in the fp32 residual stream.  HBM-bound.
#include "common.cuh"

namespace dfot {
namespace vae {

constexpr int kThreads = 256;
constexpr int kPad = 2;

// x2 spatial (and, with TIME, x2 temporal on all frames but the first) upsampling, fp32 clip -> bf16 clip with pads.
//   TIME = false: nearest x2 in (H, W)                               (SpatialUpsample2x, updownsample.py:73-80)
//   TIME = true : frame 0 bilinear x2; frames 1.. trilinear x(2,2,2) (Spatial2xTime2x3DUpsample, updownsample.py:131-147)
// align_corners = False with scale 2 has fixed weights: output 2i blends inputs (i-1, i) with (1/4, 3/4), output 2i+1
// blends (i, i+1) with (3/4, 1/4), indices clamped at the borders.  One thread = 4 channels of one INPUT pixel of one
// "unit" and writes the whole output block that pixel owns, so every input tap is loaded once per block instead of once
// per output pixel (the first version, one thread per output pixel, re-read 16 B per byte stored and ran at 1.1 TB/s):
//   nearest: unit = input frame t                -> 2 x 2 outputs (+ the pad slots when t = 0);
//   TIME   : unit 0 = input frame 0              -> 2 x 2 outputs of output frame 0 (+ the pad slots);
//            unit 2 + k, k = -1 .. Tn - 1        -> frames (k, k + 1) of the Tn = Tin - 1 later frames (clamped) ->
//                                                   outputs 2k + 1 with (3/4, 1/4) and 2k + 2 with (1/4, 3/4).
// Output slots 0 .. pad-1 of a clip replicate its first frame (pad = kPad for video clips, 0 for plain image batches).
constexpr int kUpVec = 4;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 mix(float wa, float4 a, float wb, float4 b) {
  return make_float4(wa * a.x + wb * b.x, wa * a.y + wb * b.y, wa * a.z + wb * b.z, wa * a.w + wb * b.w);
}
__device__ __forceinline__ void st4_bf16(__nv_bfloat16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// bilinear x2 block of input pixel (y, x) of one frame: o[dy][dx] = output pixel (2y + dy, 2x + dx)
__device__ __forceinline__ void bilinear_block(const float* frame, int y, int x, int H, int W, int C, float4 (&o)[2][2]) {
  const int xl = max(x - 1, 0), xr = min(x + 1, W - 1);
  const int rows[3] = {max(y - 1, 0), y, min(y + 1, H - 1)};
  float4 h[3][2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* row = frame + (int64_t)rows[r] * W * C;
    const float4 l = ld4(row + (int64_t)xl * C), c = ld4(row + (int64_t)x * C), rr = ld4(row + (int64_t)xr * C);
    h[r][0] = mix(0.25f, l, 0.75f, c);
    h[r][1] = mix(0.75f, c, 0.25f, rr);
  }
#pragma unroll
  for (int dx = 0; dx < 2; ++dx) {
    o[0][dx] = mix(0.25f, h[0][dx], 0.75f, h[1][dx]);
    o[1][dx] = mix(0.75f, h[1][dx], 0.25f, h[2][dx]);
  }
}

template <bool TIME>
__global__ void __launch_bounds__(kThreads)
upsample2x_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Tin, int H, int W, int C,
                  int pad) {
  pdl_trigger();
  pdl_wait();
  const int vecs = C / kUpVec, Ho = 2 * H, Wo = 2 * W;
  const int Tn = Tin - 1;                                            // frames that are interpolated in time
  const int Tout = TIME ? 2 * Tin - 1 : Tin;
  const int units = TIME ? (Tn > 0 ? Tn + 2 : 1) : Tin;
  const int64_t total = (int64_t)B * units * H * W * vecs;
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % vecs);
  int64_t r = idx / vecs;
  const int x = (int)(r % W); r /= W;
  const int y = (int)(r % H); r /= H;
  const int u = (int)(r % units);
  const int b = (int)(r / units);
  const int64_t fin = (int64_t)H * W * C, fout = (int64_t)Ho * Wo * C;
  const float* clip = in + ((int64_t)b * (pad + Tin) + pad) * fin + kUpVec * v;
  __nv_bfloat16* oclip = out + (int64_t)b * (pad + Tout) * fout + ((int64_t)(2 * y) * Wo + 2 * x) * C + kUpVec * v;
  auto store_block = [&](int slot, const float4 (&o)[2][2]) {
    __nv_bfloat16* d = oclip + (int64_t)slot * fout;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) st4_bf16(d + ((int64_t)dy * Wo + dx) * C, o[dy][dx]);
  };
  if constexpr (!TIME) {
    const float4 val = ld4(clip + (int64_t)u * fin + ((int64_t)y * W + x) * C);
    const float4 o[2][2] = {{val, val}, {val, val}};
    store_block(pad + u, o);
    if (u == 0)
      for (int p = 0; p < pad; ++p) store_block(p, o);
  } else {
    if (u == 0) {                                                    // first frame: spatial only; pads replicate it
      float4 o[2][2];
      bilinear_block(clip, y, x, H, W, C, o);
      for (int p = 0; p <= pad; ++p) store_block(p, o);
      return;
    }
    const int k = u - 2;                                             // -1 .. Tn - 1
    const int fa = max(k, 0), fb = min(k + 1, Tn - 1);
    float4 a[2][2], bb[2][2];
    bilinear_block(clip + (int64_t)(1 + fa) * fin, y, x, H, W, C, a);
    if (fb != fa) {
      bilinear_block(clip + (int64_t)(1 + fb) * fin, y, x, H, W, C, bb);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) bb[q >> 1][q & 1] = a[q >> 1][q & 1];
    }
    float4 o[2][2];
    if (k >= 0) {                                                    // output 2k + 1 of the later frames = frame 2k + 2
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q >> 1][q & 1] = mix(0.75f, a[q >> 1][q & 1], 0.25f, bb[q >> 1][q & 1]);
      store_block(pad + 2 * k + 2, o);
    }
    if (k + 1 <= Tn - 1) {                                           // output 2k + 2 = frame 2k + 3
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q >> 1][q & 1] = mix(0.25f, a[q >> 1][q & 1], 0.75f, bb[q >> 1][q & 1]);
      store_block(pad + 2 * k + 3, o);
    }
  }
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
  const int64_t units = temporal ? (T_in > 1 ? T_in + 1 : 1) : T_in;      // see upsample2x_kernel
  const int64_t total = B * units * H * W * (C / vae::kUpVec);
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
  const int64_t total = n_img * H * W * (C / vae::kUpVec);
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
