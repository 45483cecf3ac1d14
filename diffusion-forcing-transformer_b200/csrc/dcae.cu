// Glue kernels of the DC-AE image decoder (the VAE of the DMLab / Minecraft latent configurations;
// reference: algorithms/vae/dc_ae/autoencoder_dc_model.py — Decoder :383-470, DCUpBlock2d :222-260, ResBlock :109-136,
// EfficientViTBlock :139-172, SanaMultiscaleLinearAttention :46-106 — plus the diffusers==0.32.2 pieces it imports:
// GLUMBConv, RMSNorm, SanaMultiscaleAttnProcessor2_0).  All 3x3 / 1x1 convolutions and linears run on the tcgen05
// GEMM / implicit-GEMM kernels; what is left is data movement and small per-token / per-image arithmetic on channel-last
// tensors [n, H, W, C].  HBM-bound and small (a DMLab frame is 8x8 latents -> 64x64 pixels): clarity over tuning.
#include "common.cuh"

namespace dfot {
namespace dcae {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ ReLU in place (ResBlock nonlinearity, conv_act)
__global__ void __launch_bounds__(kThreads) relu_bf16_kernel(__nv_bfloat16* __restrict__ x, int64_t n8) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n8) return;
  uint4 v = *reinterpret_cast<uint4*>(x + 8 * i);
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = unpack_bf16x2(w[k]);
    w[k] = pack_bf16x2(fmaxf(f.x, 0.f), fmaxf(f.y, 0.f));
  }
  *reinterpret_cast<uint4*>(x + 8 * i) = v;
}

// ------------------------------------------------------------------ pixel shuffle x2 (+ channel-repeat shortcut)
// DCUpBlock2d.forward: out[n, c, 2h+i, 2w+j] = conv[n, 4c + 2i + j, h, w] + x[n, (4c + 2i + j) / repeats, h, w]
// (F.pixel_shuffle of the conv output plus F.pixel_shuffle of x.repeat_interleave(repeats, dim=1)); channel-last here.
// One thread per OUTPUT element; conv has ld_conv >= 4C channels per pixel (its columns beyond 4C are padding).
__global__ void __launch_bounds__(kThreads)
pixel_shuffle2x_kernel(const float* __restrict__ conv, int64_t ld_conv, const float* __restrict__ x, int Cx, int repeats,
                       float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int64_t total, int H, int W,
                       int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= total) return;
  const int c = (int)(e % C);
  int64_t p = e / C;
  const int ox = (int)(p % (2 * W));
  p /= 2 * W;
  const int oy = (int)(p % (2 * H));
  const int64_t n = p / (2 * H);
  const int k = 4 * c + 2 * (oy & 1) + (ox & 1);
  const int64_t src = (n * H + (oy >> 1)) * W + (ox >> 1);
  float v = conv[src * ld_conv + k];
  if (x != nullptr) v += x[src * Cx + k / repeats];
  if (out_f32 != nullptr) out_f32[e] = v;
  if (out_bf16 != nullptr) out_bf16[e] = __float2bfloat16(v);
}

// ------------------------------------------------------------------ ReLU linear attention (EfficientViT / Sana)
// SanaMultiscaleAttnProcessor2_0 + apply_linear_attention (:87-96): the [q | k | v] projection of a token is read as
// `heads` consecutive groups of 3*d channels (query, key, value of the group — the processor reshapes the concatenated
// tensor, so a group does NOT line up with one projection; the weights are trained with exactly this layout), q and k
// pass a ReLU, and per (image, group)
//     KV[a][b] = sum_t k[t][a] * [v[t] | 1][b]            (d x (d + 1), fp32)
//     out[t][b] = (sum_a q[t][a] KV[a][b]) / (sum_a q[t][a] KV[a][d] + eps)
// grid = (groups, images); the tokens of an image (H*W, 64 for DMLab) are walked by the block.  d <= 32.
constexpr int kLaMaxD = 32;
__global__ void __launch_bounds__(128)
linear_attention_kernel(const float* __restrict__ qkv, int64_t ld, float* __restrict__ out, int64_t ld_out, int HW, int d,
                        float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_kv[kLaMaxD][kLaMaxD + 1];
  const int g = blockIdx.x;
  const int64_t row0 = (int64_t)blockIdx.y * HW;
  const float* base = qkv + row0 * ld + (int64_t)g * 3 * d;
  // KV: thread (a, b) for a < d, b <= d
  for (int idx = threadIdx.x; idx < d * (d + 1); idx += blockDim.x) {
    const int a = idx / (d + 1), b = idx % (d + 1);
    float acc = 0.f;
    for (int t = 0; t < HW; ++t) {
      const float* tok = base + (int64_t)t * ld;
      const float kk = fmaxf(tok[d + a], 0.f);
      acc = fmaf(kk, b < d ? tok[2 * d + b] : 1.f, acc);
    }
    s_kv[a][b] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < HW * d; idx += blockDim.x) {
    const int t = idx / d, b = idx % d;
    const float* tok = base + (int64_t)t * ld;
    float num = 0.f, den = 0.f;
    for (int a = 0; a < d; ++a) {
      const float q = fmaxf(tok[a], 0.f);
      num = fmaf(q, s_kv[a][b], num);
      den = fmaf(q, s_kv[a][d], den);
    }
    out[(row0 + t) * ld_out + (int64_t)g * d + b] = num / (den + eps);
  }
}

// ------------------------------------------------------------------ GLUMBConv middle: depthwise 3x3 + bias, then GLU
// x [n, H, W, 2*Ch] (already SiLU'd by the 1x1 expansion's epilogue), w [2*Ch, 3, 3], b [2*Ch]:
//   y = dwconv3x3(x) + b;   out[.., c] = y[.., c] * silu(y[.., Ch + c])   -> bf16 [n, H, W, Ch]  (conv_point's operand)
__global__ void __launch_bounds__(kThreads)
dwconv3x3_glu_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                     __nv_bfloat16* __restrict__ out, int64_t total, int H, int W, int Ch) {
  pdl_trigger();
  pdl_wait();
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= total) return;
  const int c = (int)(e % Ch);
  int64_t p = e / Ch;
  const int px = (int)(p % W);
  p /= W;
  const int py = (int)(p % H);
  const int64_t n = p / H;
  float y[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int ch = c + half * Ch;
    float acc = b[ch];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = py + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = px + dx;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(__bfloat162float(x[((n * H + yy) * W + xx) * (2 * Ch) + ch]), w[ch * 9 + (dy + 1) * 3 + (dx + 1)], acc);
      }
    }
    y[half] = acc;
  }
  out[e] = __float2bfloat16(y[0] * silu_f(y[1]));
}

// ------------------------------------------------------------------ RMSNorm over channels with affine (+ residual, ReLU)
// diffusers RMSNorm(dim, eps, elementwise_affine=True, bias=True): y = x * rsqrt(mean(x^2) + eps) * w + b; then the
// caller's residual add (EfficientViT attention / GLUMBConv) or ReLU (Decoder.conv_act).  One warp per token.
__global__ void __launch_bounds__(kThreads)
rmsnorm_affine_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ w, const float* __restrict__ b,
                      float eps, const float* __restrict__ resid, int relu, float* __restrict__ out_f32,
                      __nv_bfloat16* __restrict__ out_bf16, int64_t M, int C) {
  pdl_trigger();
  pdl_wait();
  const int64_t row = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * ld_x;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) ss = fmaf(xr[c], xr[c], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / (float)C + eps);
  for (int c = lane; c < C; c += 32) {
    float y = xr[c] * r * w[c] + b[c];
    if (resid != nullptr) y += resid[row * C + c];
    if (relu) y = fmaxf(y, 0.f);
    if (out_f32 != nullptr) out_f32[row * C + c] = y;
    if (out_bf16 != nullptr) out_bf16[row * C + c] = __float2bfloat16(y);
  }
}

static inline unsigned blocks_for(int64_t n) { return (unsigned)ceil_div(n, kThreads); }

}  // namespace dcae
}  // namespace dfot

using namespace dfot;
using namespace dfot::dcae;

extern "C" int dfot_relu_bf16(void* x, int64_t n, void* stream) {
  DFOT_REQUIRE(x && n > 0 && n % 8 == 0 && (uintptr_t)x % 16 == 0, DFOT_ERR_INVALID_ARG,
               "relu_bf16: need a 16-byte aligned buffer of a multiple of 8 elements");
  launch_pdl(relu_bf16_kernel, dim3(blocks_for(n / 8)), dim3(kThreads), 0, (cudaStream_t)stream, (__nv_bfloat16*)x, n / 8);
  DFOT_CHECK_LAUNCH("relu_bf16");
  return DFOT_OK;
}

extern "C" int dfot_pixel_shuffle2x(const float* conv, int64_t ld_conv, const float* x, int64_t Cx, int64_t repeats,
                                    float* out_f32, void* out_bf16, int64_t n_img, int64_t H, int64_t W, int64_t C,
                                    void* stream) {
  DFOT_REQUIRE(conv && (out_f32 || out_bf16) && n_img > 0 && H > 0 && W > 0 && C > 0 && ld_conv >= 4 * C,
               DFOT_ERR_INVALID_ARG, "pixel_shuffle2x: bad arguments");
  DFOT_REQUIRE(x == nullptr || (repeats >= 1 && Cx * repeats == 4 * C), DFOT_ERR_INVALID_ARG,
               "pixel_shuffle2x: shortcut needs Cx * repeats == 4 * C (got %lld * %lld vs %lld)", (long long)Cx,
               (long long)repeats, (long long)(4 * C));
  const int64_t total = n_img * 4 * H * W * C;
  launch_pdl(pixel_shuffle2x_kernel, dim3(blocks_for(total)), dim3(kThreads), 0, (cudaStream_t)stream, conv, ld_conv, x,
             (int)Cx, (int)(repeats < 1 ? 1 : repeats), out_f32, (__nv_bfloat16*)out_bf16, total, (int)H, (int)W, (int)C);
  DFOT_CHECK_LAUNCH("pixel_shuffle2x");
  return DFOT_OK;
}

extern "C" int dfot_linear_attention_relu(const float* qkv, int64_t ld_qkv, float* out, int64_t ld_out, int64_t n_img,
                                          int64_t HW, int64_t heads, int64_t head_dim, float eps, void* stream) {
  DFOT_REQUIRE(qkv && out && n_img > 0 && HW > 0 && heads > 0, DFOT_ERR_INVALID_ARG, "linear_attention: bad arguments");
  DFOT_REQUIRE(head_dim >= 1 && head_dim <= kLaMaxD && ld_qkv >= 3 * heads * head_dim && ld_out >= heads * head_dim &&
                   n_img < 65536,
               DFOT_ERR_UNSUPPORTED, "linear_attention: head_dim must be <= %d and the row strides must cover the heads",
               kLaMaxD);
  launch_pdl(linear_attention_kernel, dim3((unsigned)heads, (unsigned)n_img), dim3(128), 0, (cudaStream_t)stream, qkv,
             ld_qkv, out, ld_out, (int)HW, (int)head_dim, eps);
  DFOT_CHECK_LAUNCH("linear_attention_relu");
  return DFOT_OK;
}

extern "C" int dfot_dwconv3x3_glu_bf16(const void* x, const float* w, const float* b, void* out, int64_t n_img, int64_t H,
                                       int64_t W, int64_t Ch, void* stream) {
  DFOT_REQUIRE(x && w && b && out && n_img > 0 && H > 0 && W > 0 && Ch > 0, DFOT_ERR_INVALID_ARG,
               "dwconv3x3_glu: bad arguments");
  const int64_t total = n_img * H * W * Ch;
  launch_pdl(dwconv3x3_glu_kernel, dim3(blocks_for(total)), dim3(kThreads), 0, (cudaStream_t)stream,
             (const __nv_bfloat16*)x, w, b, (__nv_bfloat16*)out, total, (int)H, (int)W, (int)Ch);
  DFOT_CHECK_LAUNCH("dwconv3x3_glu");
  return DFOT_OK;
}

extern "C" int dfot_rmsnorm_affine(const float* x, int64_t ld_x, const float* w, const float* b, float eps,
                                   const float* resid, int relu, float* out_f32, void* out_bf16, int64_t M, int64_t C,
                                   void* stream) {
  DFOT_REQUIRE(x && w && b && (out_f32 || out_bf16) && M > 0 && C > 0 && ld_x >= C, DFOT_ERR_INVALID_ARG,
               "rmsnorm_affine: bad arguments");
  launch_pdl(rmsnorm_affine_kernel, dim3((unsigned)ceil_div(M, kThreads / 32)), dim3(kThreads), 0, (cudaStream_t)stream, x,
             ld_x, w, b, eps, resid, relu, out_f32, (__nv_bfloat16*)out_bf16, M, (int)C);
  DFOT_CHECK_LAUNCH("rmsnorm_affine");
  return DFOT_OK;
}
