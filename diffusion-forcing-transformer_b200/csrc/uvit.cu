// HBM-bound kernels of the U-ViT3DPose backbone (channel-last activations everywhere):
//   GroupNorm statistics / GroupNorm+FiLM+SiLU apply, RMSNorm+FiLM, q/k RMSNorm + RoPE-3D, 2x2 average pooling,
//   nearest-2x upsample + skip add, subtraction, camera-ray encoding straight into PatchEmbed rows.
// Every kernel reads each input element once and writes each output element once with 128-bit accesses.
#include "common.cuh"

namespace dfot {
namespace uvit {

constexpr int kThreads = 256;

// ---- 8-channel vector load/store helpers (f32: 2 x 16 B, bf16: 1 x 16 B) ----
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const uint4 a = ld_stream_u4(p), b = ld_stream_u4(p + 4);
  v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
  v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = ld_stream_u4(p);
  float2 t;
  t = unpack_bf16x2(a.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16x2(a.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16x2(a.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16x2(a.w); v[6] = t.x; v[7] = t.y;
}
// cached (re-read) variants for small tables
__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  st_stream_u4(p, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                             pack_bf16x2(v[6], v[7])));
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  st_stream_u4(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
  st_stream_u4(p + 4, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
}

// ------------------------------------------------------------------ GroupNorm statistics
// DETERMINISTIC and BATCH-INVARIANT: (sum, sum of squares) are accumulated as 64-bit FIXED-POINT integers (2^-32 units).
// Integer addition is associative, so neither the order in which warps / blocks reach their atomics nor the way the
// pixels are split over blocks (which depends on the number of images in the launch) can change a single bit; the only
// floating-point sums are fp32 "micro-partials" over one thread's 8 channels x kGnMicro pixels of a globally aligned
// pixel chunk, whose composition depends on (HW, C) alone.  Two forwards of the same image — alone, inside a batch, on
// another rank — therefore see identical statistics (the r01 version used fp32 shared-memory atomics and agreed to ~1e-6).
#ifndef DFOT_GN_MICRO_BF16
#define DFOT_GN_MICRO_BF16 8
#endif
// pixels per fp32 micro-partial (per input type: a bf16 pixel vector is one 16-byte load, an fp32 one two)
template <typename TX> struct GnMicro { static constexpr int value = 8; };
template <> struct GnMicro<__nv_bfloat16> { static constexpr int value = DFOT_GN_MICRO_BF16; };
constexpr float kGnFix = 4294967296.f;                    // 2^32: |partial| < 2^31 is ample for activations
__device__ __forceinline__ unsigned long long gn_fix(float v) { return (unsigned long long)__float2ll_rn(v * kGnFix); }

// grid (slabs, n_img); a thread owns one 8-channel vector (fixed) and walks the slab's pixels; GPV = groups per vector.
// (register cap measured on B200, 64 x 16384 x 128: with a bare __launch_bounds__(256) ptxas keeps 48 (f32) / 92 (bf16)
// registers and serialises the micro-partial's loads — 4.1 / 3.2 TB/s; min-blocks 1 lets it keep all of them in flight with
// 96 / 64 registers — 6.15 TB/s = 94 % of the copy peak / 4.7 TB/s; caps of 3, 4, 6 blocks: 6.2 / 3.9, 6.0 / 3.2, 4.3 / 1.7.
// bf16 micro-partials of 16 pixels: 2.4 TB/s, of 4: 4.8 — the bf16 kernel is bound by its unpack + accumulate instructions.)
#ifndef DFOT_GN_STATS_MIN_BLOCKS
#define DFOT_GN_STATS_MIN_BLOCKS 1
#endif
template <typename TX, int GPV>
__global__ void __launch_bounds__(kThreads, DFOT_GN_STATS_MIN_BLOCKS)
gn_stats_kernel(const TX* __restrict__ x, double* __restrict__ sums, int64_t HW, int64_t img_stride, int C, int G,
                int pix_per_block) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  __shared__ unsigned long long s_sum[64], s_sq[64];
  const int vecs = C >> 3, cpg = C / G;
  const int img = blockIdx.y;
  if (threadIdx.x < 64) { s_sum[threadIdx.x] = 0ull; s_sq[threadIdx.x] = 0ull; }
  __syncthreads();
  const int v = threadIdx.x % vecs, lane_pix = threadIdx.x / vecs, pix_step = kThreads / vecs;
  constexpr int CPV = 8 / GPV;                            // channels of one group inside the vector
  constexpr int kGnMicro = GnMicro<TX>::value;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;  // multiple of kGnMicro * pix_step (host): chunks are aligned
  const int64_t p1 = min(HW, p0 + pix_per_block);
  unsigned long long as[GPV], aq[GPV];
#pragma unroll
  for (int g = 0; g < GPV; ++g) { as[g] = 0ull; aq[g] = 0ull; }
  if (lane_pix < pix_step) {
    for (int64_t base = p0 + lane_pix; base < p1; base += (int64_t)kGnMicro * pix_step) {
      float a[kGnMicro][8];
#pragma unroll
      for (int k = 0; k < kGnMicro; ++k) {                // all loads of the micro-partial in flight
        const int64_t pix = base + (int64_t)k * pix_step;
        if (pix < p1) load8(x + (int64_t)img * img_stride + pix * C + 8 * v, a[k]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[k][j] = 0.f;
        }
      }
#pragma unroll
      for (int g = 0; g < GPV; ++g) {
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int k = 0; k < kGnMicro; ++k)
#pragma unroll
          for (int j = 0; j < CPV; ++j) { const float e = a[k][g * CPV + j]; ts += e; tq = fmaf(e, e, tq); }
        as[g] += gn_fix(ts);
        aq[g] += gn_fix(tq);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < GPV; ++g) {
    const int grp = (8 * v + g * CPV) / cpg;
    atomicAdd(&s_sum[grp], as[g]);
    atomicAdd(&s_sq[grp], aq[g]);
  }
  __syncthreads();
  if (threadIdx.x < G) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(sums) + ((int64_t)img * G + threadIdx.x) * 2;
    atomicAdd(dst, s_sum[threadIdx.x]);
    atomicAdd(dst + 1, s_sq[threadIdx.x]);
  }
}

// fixed-point (sum, sum of squares) → (mean, rstd) in f32, once per (image, group): keeps every f64 operation out of the
// per-element kernel
__global__ void gn_finalize_kernel(const double* __restrict__ sums, float2* __restrict__ stats, int64_t n, double inv_n,
                                   float eps) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long* fx = reinterpret_cast<const long long*>(sums);
  constexpr double kUnit = 1.0 / 4294967296.0;
  const double m = (double)fx[2 * i] * kUnit * inv_n;
  const double var = fmax((double)fx[2 * i + 1] * kUnit * inv_n - m * m, 0.0);
  stats[i] = make_float2((float)m, rsqrtf((float)var + eps));
}

// ------------------------------------------------------------------ GroupNorm (+FiLM) + SiLU -> bf16
// grid (pixel slabs, images); a thread owns ONE 8-channel vector (so gamma / beta / statistics / per-image FiLM are
// loaded once and folded into a per-channel affine) and walks kGnPix pixels of its slab with all loads issued up
// front — no 64-bit index arithmetic, 4 independent 16/32-byte loads in flight per thread.
// Pixels in flight per thread and resident blocks per SM.  Measured on B200 (bf16 input + per-pixel FiLM stream, 64 x 128² x
// 128): 2 pixels x 3 blocks 169.5 us (4.75 TB/s), 4 x 2 181, 2 x 4 186, 1 x 4 141.9 us (5.68 TB/s = 87 % of the copy peak),
// 1 x 5 154, 1 x 6 163: many light threads beat few heavy ones; the f32 path is unchanged (123.5-124.4 us, 6.5 TB/s).
#ifndef DFOT_GN_BLOCKS
#define DFOT_GN_BLOCKS 4
#endif
#ifndef DFOT_GN_PIX
#define DFOT_GN_PIX 1
#endif
constexpr int kGnPix = DFOT_GN_PIX, kGnIter = 8 / DFOT_GN_PIX;
template <typename TX, bool SILU = true>
__global__ void __launch_bounds__(kThreads, DFOT_GN_BLOCKS)
gn_silu_kernel(const TX* __restrict__ x, const float2* __restrict__ stats, const float* __restrict__ gamma,
               const float* __restrict__ beta, const float* __restrict__ mod_img, int64_t ld_img,
               int64_t scale_col, int64_t shift_col, const __nv_bfloat16* __restrict__ mod_pix,
               const int32_t* __restrict__ img_map, __nv_bfloat16* __restrict__ y, int HW, int C, int G,
               int64_t img_stride) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int vecs = C >> 3, cpg = C / G;
  const int img = blockIdx.y;
  const int v = threadIdx.x % vecs, prow = threadIdx.x / vecs, pix_step = kThreads / vecs;
  const int c0 = 8 * v;
  // per-channel affine of this image: t = x * A + B  (GroupNorm), then y = t * (1 + scale) + shift (FiLM)
  float A[8], B[8], sc[8], sh[8];
  {
    float ga[8], be[8];
    ldg8(gamma + c0, ga);
    ldg8(beta + c0, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 st = __ldg(stats + (int64_t)img * G + (c0 + j) / cpg);
      A[j] = st.y * ga[j];
      B[j] = be[j] - st.x * A[j];
    }
  }
  const bool film = mod_img != nullptr;
  if (film) {
    ldg8(mod_img + (int64_t)img * ld_img + scale_col + c0, sc);
    ldg8(mod_img + (int64_t)img * ld_img + shift_col + c0, sh);
  }
  const int32_t src = (film && mod_pix != nullptr && img_map != nullptr) ? __ldg(img_map + img) : -1;
  if (film && src < 0) {   // no per-pixel part: fold FiLM into the affine
#pragma unroll
    for (int j = 0; j < 8; ++j) { A[j] *= 1.f + sc[j]; B[j] = fmaf(B[j], 1.f + sc[j], sh[j]); }
  }
  const TX* xb = x + (int64_t)img * img_stride + c0;
  __nv_bfloat16* yb = y + (int64_t)img * img_stride + c0;
  const __nv_bfloat16* pb = src >= 0 ? mod_pix + (int64_t)src * HW * (2 * C) + c0 : nullptr;
  // kGnIter groups of kGnPix pixels: the prologue above (≈20 cached loads) is amortised over kGnIter*kGnPix pixels.
  // Images without a per-pixel part have a third of the loads per pixel, so they walk the same pixels in half as many
  // groups of 2*kGnPix to keep as many bytes in flight.
  const int first = blockIdx.x * (kGnIter * kGnPix) * pix_step + prow;
  if (pb != nullptr) {
#pragma unroll 1
    for (int it = 0; it < kGnIter; ++it) {
      const int pix0 = first + it * (kGnPix * pix_step);
      if (pix0 >= HW) break;
      float a[kGnPix][8], ps[kGnPix][8], ph[kGnPix][8];
#pragma unroll
      for (int k = 0; k < kGnPix; ++k) {
        const int pix = pix0 + k * pix_step;
        if (pix < HW) {
          load8(xb + (int64_t)pix * C, a[k]);
          load8(pb + (int64_t)pix * (2 * C), ps[k]);
          load8(pb + (int64_t)pix * (2 * C) + C, ph[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < kGnPix; ++k) {
        const int pix = pix0 + k * pix_step;
        if (pix < HW) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            a[k][j] = silu_fast_f(fmaf(fmaf(a[k][j], A[j], B[j]), 1.f + (sc[j] + ps[k][j]), sh[j] + ph[k][j]));
          store8(yb + (int64_t)pix * C, a[k]);
        }
      }
    }
  } else {
    constexpr int kPix = 2 * kGnPix;
#pragma unroll 1
    for (int it = 0; it < kGnIter / 2; ++it) {
      const int pix0 = first + it * (kPix * pix_step);
      if (pix0 >= HW) break;
      float a[kPix][8];
#pragma unroll
      for (int k = 0; k < kPix; ++k) {
        const int pix = pix0 + k * pix_step;
        if (pix < HW) load8(xb + (int64_t)pix * C, a[k]);
      }
#pragma unroll
      for (int k = 0; k < kPix; ++k) {
        const int pix = pix0 + k * pix_step;
        if (pix < HW) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = fmaf(a[k][j], A[j], B[j]);
            a[k][j] = SILU ? silu_fast_f(t) : t;
          }
          store8(yb + (int64_t)pix * C, a[k]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ RMSNorm + FiLM -> bf16 (warp per token)
constexpr int kNormWarps = 4;
// Resident blocks per SM (register cap): measured on B200, D = 1152 (NV = 9): 5 blocks 30.1 us, 6 31.4, 8 35.8 (spills),
// uncapped 31.3-33.5; D = 576 (NV = 5): 8 blocks 55.0 us, uncapped 56.9.
template <int NV>
__global__ void __launch_bounds__(kNormWarps * 32, NV >= 8 ? 5 : 8)
rmsnorm_film_kernel(const float* __restrict__ x, const float* __restrict__ weight, float eps,
                    const float* __restrict__ mod_img, int64_t ld_img, int64_t scale_col, int64_t shift_col,
                    const __nv_bfloat16* __restrict__ mod_pix, const int32_t* __restrict__ img_map,
                    __nv_bfloat16* __restrict__ y, int64_t M, int D, int64_t tokens_per_img) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * kNormWarps + warp;
  if (m >= M) return;
  const int nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + m * D);
  float4 v[NV];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      const uint4 u = ld_stream_u4(xr + c);
      v[i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // the per-pixel FiLM part is a second DRAM stream: issue its loads before the reduction so that both streams of the
  // token are in flight together (one round trip per token instead of two)
  const int64_t img = m / tokens_per_img, pix = m - img * tokens_per_img;
  const int32_t src = (mod_pix != nullptr && img_map != nullptr) ? __ldg(img_map + img) : -1;
  const __nv_bfloat16* prow = src >= 0 ? mod_pix + ((int64_t)src * tokens_per_img + pix) * (2 * D) : nullptr;
  uint2 ps[NV], ph[NV];
  if (prow != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) { ps[i] = ld_stream_u2(prow + 4 * c); ph[i] = ld_stream_u2(prow + D + 4 * c); }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
  const float4* sc = reinterpret_cast<const float4*>(mod_img + img * ld_img + scale_col);
  const float4* sh = reinterpret_cast<const float4*>(mod_img + img * ld_img + shift_col);
  const float4* wr = reinterpret_cast<const float4*>(weight);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 g = __ldg(sc + c), s = __ldg(sh + c);
      const float4 w = __ldg(wr + c);
      if (prow != nullptr) {
        const float2 a = unpack_bf16x2(ps[i].x), b = unpack_bf16x2(ps[i].y), e = unpack_bf16x2(ph[i].x), f = unpack_bf16x2(ph[i].y);
        g.x += a.x; g.y += a.y; g.z += b.x; g.w += b.y;
        s.x += e.x; s.y += e.y; s.z += f.x; s.w += f.y;
      }
      float4 o;
      o.x = fmaf(v[i].x * rstd * w.x, 1.f + g.x, s.x);
      o.y = fmaf(v[i].y * rstd * w.y, 1.f + g.y, s.y);
      o.z = fmaf(v[i].z * rstd * w.z, 1.f + g.z, s.z);
      o.w = fmaf(v[i].w * rstd * w.w, 1.f + g.w, s.w);
      st_stream_u2(y + m * D + 4 * c, make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w)));
    }
  }
}

// ------------------------------------------------------------------ q/k RMSNorm(head_dim) + RoPE-3D, in place
// One warp per token.  The q and k columns of a token are one contiguous run of 2*heads*DH bf16, streamed as 16-byte
// pieces: a lane owns 8 adjacent elements (4 rotation pairs) of one head, LPH = DH/8 lanes share a head and one load
// instruction covers 32/LPH heads (512 contiguous bytes).  ALL pieces of the token are loaded before the first reduction
// (NI loads in flight per lane, one DRAM round trip per token); head sums are xor-shuffles inside the LPH-lane group.
// (register cap measured on B200, d = 64: uncapped = 3 blocks/SM 63.9 us, 1 block 73.3, 4 blocks 99.0 (spills); d = 128 needs
// the uncapped allocation: 28.4 us vs 60.0 with a 3-block cap)
template <int DH, int NI>
__global__ void __launch_bounds__(kThreads)
qk_norm_rope_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ld, const float* __restrict__ qw,
                    const float* __restrict__ kw, float eps, const float* __restrict__ rope_cs,
                    int64_t tokens_per_sample, int64_t M, int heads, float q_scale) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  constexpr int LPH = DH / 8, HPI = 32 / LPH;       // lanes per head, heads per load instruction
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int sub = lane % LPH, hsel = lane / LPH;    // position inside the head, head inside the instruction
  uint4* row = reinterpret_cast<uint4*>(qkv + m * ld) + lane;
  const int n_seg = 2 * heads;                      // segment s: columns [s*DH, (s+1)*DH); s < heads → q, else k
  uint4 raw[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i)
    if (i * HPI + hsel < n_seg) raw[i] = *(row + i * 32);
  float wq[8], wk[8];
  float2 cs[4];
  const int64_t tok = m % tokens_per_sample;
  ldg8(qw + sub * 8, wq);
  ldg8(kw + sub * 8, wk);
#pragma unroll
  for (int j = 0; j < 8; ++j) wq[j] *= q_scale;
  {
    const float4* c4 = reinterpret_cast<const float4*>(rope_cs + (tok * (DH / 2) + sub * 4) * 2);
    const float4 a = __ldg(c4), b = __ldg(c4 + 1);
    cs[0] = make_float2(a.x, a.y); cs[1] = make_float2(a.z, a.w);
    cs[2] = make_float2(b.x, b.y); cs[3] = make_float2(b.z, b.w);
  }
  float sq[NI];             // the raw bf16 pieces stay in registers; they are unpacked twice rather than kept as f32
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    sq[i] = 0.f;
    if (i * HPI + hsel < n_seg) {
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2(u[j]);
        sq[i] = fmaf(t.x, t.x, fmaf(t.y, t.y, sq[i]));
      }
    }
  }
#pragma unroll
  for (int o = LPH / 2; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < NI; ++i) sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int seg = i * HPI + hsel;
    if (seg >= n_seg) continue;
    const float rstd = rsqrtf(sq[i] / (float)DH + eps);
    const bool is_q = seg < heads;
    const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_bf16x2(u[j]);
      const float x0 = t.x * rstd * (is_q ? wq[2 * j] : wk[2 * j]);
      const float x1 = t.y * rstd * (is_q ? wq[2 * j + 1] : wk[2 * j + 1]);
      o[j] = pack_bf16x2(x0 * cs[j].x - x1 * cs[j].y, x1 * cs[j].x + x0 * cs[j].y);
    }
    *(row + i * 32) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------ pooling / upsampling / subtraction
template <typename TI, typename TO>
__global__ void __launch_bounds__(kThreads)
avgpool2x2_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n_img, int H, int W, int C) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int vecs = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= n_img * Ho * Wo * vecs) return;
  const int v = (int)(idx % vecs);
  int64_t t = idx / vecs;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho);
  const int64_t img = t / Ho;
  const TI* base = in + (((img * H + 2 * yo) * W + 2 * xo) * (int64_t)C) + 8 * v;
  float a[8], b[8], c[8], d[8], o[8];
  load8(base, a);
  load8(base + C, b);
  load8(base + (int64_t)W * C, c);
  load8(base + (int64_t)W * C + C, d);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (((a[j] + b[j]) + c[j]) + d[j]) * 0.25f;
  store8(out + (((img * Ho + yo) * Wo + xo) * (int64_t)C) + 8 * v, o);
}

__global__ void __launch_bounds__(kThreads)
sub_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b, __nv_bfloat16* __restrict__ out, int64_t n8) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= n8) return;
  float x[8], y[8];
  load8(a + idx * 8, x);
  load8(b + idx * 8, y);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] -= y[j];
  store8(out + idx * 8, x);
}

// out (H x W) = nearest-2x(low (H/2 x W/2)) + skip (H x W)
__global__ void __launch_bounds__(kThreads)
upsample2x_add_kernel(const float* __restrict__ low, const float* __restrict__ skip, float* __restrict__ out,
                      int64_t n_img, int H, int W, int C) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int vecs = C >> 3;
  const int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (idx >= n_img * H * W * vecs) return;
  const int v = (int)(idx % vecs);
  int64_t t = idx / vecs;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H);
  const int64_t img = t / H;
  float a[8], s[8];
  ldg8(low + (((img * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * (int64_t)C) + 8 * v, a);   // re-read 4x: keep cached
  load8(skip + idx * 8, s);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += s[j];
  store8(out + idx * 8, a);
}

// ------------------------------------------------------------------ camera rays -> ray encoding -> patch rows
// thread per pixel: 6 ray components x n_freq frequencies x {sin(e), sin(e + pi/2)} = 12*n_freq contiguous bf16.
// fp32 operation order follows the reference (geometry_utils.py:50-81): e = v * (2^s * pi); e2 = e + pi/2.
// Pixels are taken in the order of the output — token-major, (row, column) of the patch inside — so a block's kPoseThreads
// pixels own one contiguous run of a dense output (ld == p*p*12*n_freq): every thread stages its 12*n_freq values in shared
// memory (2-byte stores, a pixel's run 6*n_freq words apart: at most two lanes of a warp per bank for n_freq = 15) and the block
// writes the run with 16-byte stores.  The first version stored the values straight from the
// pixel's thread: 2-byte stores 24*n_freq B apart across a warp = one 32-byte sector per lane and store instruction, 2.85 ms
// for the 377 MB of an RE10K window (0.13 TB/s).
constexpr int kPoseThreads = 128;
__global__ void __launch_bounds__(kPoseThreads)
pose_ray_patches_kernel(const float* __restrict__ cams, const float* __restrict__ freq_scale, int n_freq,
                        __nv_bfloat16* __restrict__ out, int64_t ld, int64_t frames, int res, int p) {
  extern __shared__ __align__(16) unsigned char pose_smem[];
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(pose_smem);
  pdl_trigger();   // programmatic dependent launch: see common.cuh
  pdl_wait();
  const int pp = p * p, g = res / p, per_pix = 12 * n_freq;
  const int64_t n_pix = frames * res * res;
  const int64_t q0 = (int64_t)blockIdx.x * kPoseThreads, q = q0 + threadIdx.x;   // pixel slot in output order
  if (q < n_pix) {
    const int64_t tok = q / pp;
    const int sub = (int)(q - tok * pp);
    const int64_t fr = tok / ((int64_t)g * g);
    const int ty = (int)((tok / g) % g), tx = (int)(tok % g);
    const int X = tx * p + sub % p, Y = ty * p + sub / p;
    const float* c = cams + fr * 16;
    const float fx = __ldg(c), fy = __ldg(c + 1), px = __ldg(c + 2), py = __ldg(c + 3);
    const float dx = (((float)X + 0.5f) - px) / fx, dy = (((float)Y + 0.5f) - py) / fy;
    float v[6];
    v[0] = __ldg(c + 13); v[1] = __ldg(c + 14); v[2] = __ldg(c + 15);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      v[3 + i] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(c + 4 + 3 * i), dx), __fmul_rn(__ldg(c + 5 + 3 * i), dy)),
                           __ldg(c + 6 + 3 * i));
    const int per_part = 6 * n_freq;   // 3 components x n_freq x 2
    __nv_bfloat16* o = stage + threadIdx.x * per_pix;
    const float half_pi = 1.5707963267948966f;
    for (int part = 0; part < 2; ++part)
      for (int i = 0; i < 3; ++i)
        for (int s = 0; s < n_freq; ++s) {
          const float e = __fmul_rn(v[3 * part + i], __ldg(freq_scale + s));
          const int ch = part * per_part + i * n_freq + s;
          o[ch] = __float2bfloat16_rn(sinf(e));
          o[ch + 3 * n_freq] = __float2bfloat16_rn(sinf(__fadd_rn(e, half_pi)));
        }
  }
  __syncthreads();
  const int n_here = (int)min((int64_t)kPoseThreads, n_pix - q0);       // pixel slots of this block
  if (ld == (int64_t)pp * per_pix && (((uintptr_t)out | (uintptr_t)(q0 * per_pix * 2)) & 15) == 0) {
    // dense rows: one contiguous run of n_here * per_pix elements (per_pix * 2 B = 24 * n_freq B: a multiple of 8)
    const int n_bytes = n_here * per_pix * 2;
    unsigned char* dst = reinterpret_cast<unsigned char*>(out + q0 * per_pix);
    for (int b = threadIdx.x * 16; b + 16 <= n_bytes; b += kPoseThreads * 16)
      *reinterpret_cast<uint4*>(dst + b) = *reinterpret_cast<const uint4*>(pose_smem + b);
    if ((n_bytes & 8) && threadIdx.x == 0)
      *reinterpret_cast<uint2*>(dst + (n_bytes & ~15)) = *reinterpret_cast<const uint2*>(pose_smem + (n_bytes & ~15));
  } else {                                                              // padded rows: element by element
    for (int e = threadIdx.x; e < n_here * per_pix; e += kPoseThreads) {
      const int64_t qq = q0 + e / per_pix, tok = qq / pp;
      out[tok * ld + (qq - tok * pp) * per_pix + e % per_pix] = stage[e];
    }
  }
}

static inline unsigned blocks_for(int64_t n) { return (unsigned)ceil_div(n, kThreads); }

}  // namespace uvit

int gn_zero_sums(double* sums, int64_t n_img, int64_t groups, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * n_img * groups, s);
  DFOT_REQUIRE(e == cudaSuccess, DFOT_ERR_CUDA, "groupnorm: memset failed: %s", cudaGetErrorString(e));
  return DFOT_OK;
}
int gn_finalize(double* sums, int64_t n_img, int64_t groups, int64_t count_per_group, float eps, cudaStream_t s) {
  float2* stats = reinterpret_cast<float2*>(sums + 2 * n_img * groups);
  const int64_t n = n_img * groups;
  launch_pdl(uvit::gn_finalize_kernel, dim3((unsigned)ceil_div(n, 128)), dim3(128), 0, s, sums, stats, n, 1.0 / (double)count_per_group, eps);
  DFOT_CHECK_LAUNCH("groupnorm_finalize");
  return DFOT_OK;
}
}  // namespace dfot

using namespace dfot;
using namespace dfot::uvit;

extern "C" int dfot_groupnorm_stats(const void* x, int x_dtype, double* sums, int64_t n_img, int64_t HW, int64_t C,
                                    int64_t groups, float eps, void* stream) {
  return dfot_groupnorm_stats_strided(x, x_dtype, sums, n_img, HW, HW * C, C, groups, eps, stream);
}

extern "C" int dfot_groupnorm_stats_strided(const void* x, int x_dtype, double* sums, int64_t n_img, int64_t HW,
                                            int64_t img_stride, int64_t C, int64_t groups, float eps, void* stream) {
  DFOT_REQUIRE(x && sums && n_img > 0 && HW > 0 && C > 0 && groups > 0, DFOT_ERR_INVALID_ARG, "groupnorm_stats: bad arguments");
  DFOT_REQUIRE(img_stride >= HW * C && img_stride % 8 == 0, DFOT_ERR_INVALID_ARG,
               "groupnorm_stats: img_stride must be a multiple of 8 and >= HW*C");
  DFOT_REQUIRE(C % groups == 0 && C % 8 == 0 && groups <= 64 && n_img < 65536, DFOT_ERR_UNSUPPORTED,
               "groupnorm_stats: need C %% groups == 0, C %% 8 == 0, groups <= 64");
  const int vecs = (int)(C / 8);
  DFOT_REQUIRE(vecs <= kThreads && kThreads % vecs == 0, DFOT_ERR_UNSUPPORTED,
               "groupnorm_stats: C/8 = %d must divide %d", vecs, kThreads);
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = gn_zero_sums(sums, n_img, groups, s)) return rc;
  // ~8 resident blocks per SM over the whole batch; a block's pixel range is a whole number of aligned micro-partial
  // chunks (kGnMicro pixels per thread-row), so the fp32 micro-partials do not depend on how many blocks share an image
  const int pix_rows = kThreads / vecs;
  const int64_t chunk = (int64_t)(x_dtype == DFOT_BF16 ? GnMicro<__nv_bfloat16>::value : GnMicro<float>::value) * pix_rows;
  int64_t slabs = ceil_div(148 * 8, n_img);
  int64_t ppb = ceil_div(ceil_div(HW, slabs), chunk) * chunk;
  slabs = ceil_div(HW, ppb);
  dim3 grid((unsigned)slabs, (unsigned)n_img);
  const int64_t cpg = C / groups;
  const int gpv = cpg >= 8 ? 1 : (int)(8 / cpg);
  DFOT_REQUIRE(cpg >= 8 ? cpg % 8 == 0 : 8 % cpg == 0, DFOT_ERR_UNSUPPORTED,
               "groupnorm_stats: channels per group (%lld) must divide 8 or be a multiple of 8", (long long)cpg);
#define DFOT_GN_STATS(T, GPV)                                                                                     \
  launch_pdl(gn_stats_kernel<T, GPV>, dim3(grid), dim3(kThreads), 0, s, (const T*)x, sums, HW, img_stride, (int)C, \
             (int)groups, (int)ppb)
#define DFOT_GN_STATS_T(T)                                                                                        \
  do {                                                                                                            \
    if (gpv == 1) DFOT_GN_STATS(T, 1); else if (gpv == 2) DFOT_GN_STATS(T, 2);                                    \
    else if (gpv == 4) DFOT_GN_STATS(T, 4); else DFOT_GN_STATS(T, 8);                                             \
  } while (0)
  if (x_dtype == DFOT_F32) DFOT_GN_STATS_T(float);
  else if (x_dtype == DFOT_BF16) DFOT_GN_STATS_T(__nv_bfloat16);
  else
    DFOT_REQUIRE(false, DFOT_ERR_INVALID_ARG, "groupnorm_stats: x dtype must be f32 or bf16");
#undef DFOT_GN_STATS_T
#undef DFOT_GN_STATS
  DFOT_CHECK_LAUNCH("groupnorm_stats");
  // (mean, rstd) as f32 pairs, stored behind the f64 sums in the same buffer
  return gn_finalize(sums, n_img, groups, HW * (C / groups), eps, s);
}

extern "C" int dfot_groupnorm_silu_bf16(const void* x, int x_dtype, const double* sums, const float* gamma,
                                        const float* beta, const float* mod_img, int64_t ld_img,
                                        int64_t scale_col, int64_t shift_col, const void* mod_pix,
                                        const int32_t* img_map, void* y_bf16, int64_t n_img, int64_t HW, int64_t C,
                                        int64_t groups, void* stream) {
  DFOT_REQUIRE(x && sums && gamma && beta && y_bf16 && n_img > 0 && HW > 0 && C > 0 && groups > 0, DFOT_ERR_INVALID_ARG,
               "groupnorm_silu: bad arguments");
  DFOT_REQUIRE(C % groups == 0 && C % 8 == 0, DFOT_ERR_UNSUPPORTED, "groupnorm_silu: need C %% groups == 0 and C %% 8 == 0");
  DFOT_REQUIRE(mod_img == nullptr || (ld_img % 4 == 0 && scale_col % 4 == 0 && shift_col % 4 == 0), DFOT_ERR_UNSUPPORTED,
               "groupnorm_silu: modulation offsets must be multiples of 4");
  DFOT_REQUIRE((mod_pix == nullptr) == (img_map == nullptr) && (mod_pix == nullptr || mod_img != nullptr),
               DFOT_ERR_INVALID_ARG, "groupnorm_silu: mod_pix needs img_map and mod_img");
  cudaStream_t s = (cudaStream_t)stream;
  const int vecs = (int)(C / 8);
  DFOT_REQUIRE(vecs <= kThreads && kThreads % vecs == 0 && n_img < 65536 && HW < (1ll << 30), DFOT_ERR_UNSUPPORTED,
               "groupnorm_silu: C/8 = %d must divide %d", vecs, kThreads);
  const float2* stats = reinterpret_cast<const float2*>(sums + 2 * n_img * groups);
  const dim3 grid((unsigned)ceil_div(HW, (kThreads / vecs) * kGnPix * kGnIter), (unsigned)n_img);
  if (x_dtype == DFOT_F32)
    launch_pdl(gn_silu_kernel<float>, dim3(grid), dim3(kThreads), 0, s, 
        (const float*)x, stats, gamma, beta, mod_img, ld_img, scale_col, shift_col, (const __nv_bfloat16*)mod_pix,
        img_map, (__nv_bfloat16*)y_bf16, (int)HW, (int)C, (int)groups, HW * C);
  else if (x_dtype == DFOT_BF16)
    launch_pdl(gn_silu_kernel<__nv_bfloat16>, dim3(grid), dim3(kThreads), 0, s, 
        (const __nv_bfloat16*)x, stats, gamma, beta, mod_img, ld_img, scale_col, shift_col,
        (const __nv_bfloat16*)mod_pix, img_map, (__nv_bfloat16*)y_bf16, (int)HW, (int)C, (int)groups, HW * C);
  else
    DFOT_REQUIRE(false, DFOT_ERR_INVALID_ARG, "groupnorm_silu: x dtype must be f32 or bf16");
  DFOT_CHECK_LAUNCH("groupnorm_silu");
  return DFOT_OK;
}

// GroupNorm (+ optional SiLU) -> bf16 over strided "images" (a clip's valid frames inside a padded frame axis), no FiLM
extern "C" int dfot_groupnorm_apply_bf16(const float* x, const double* sums, const float* gamma, const float* beta,
                                         void* y_bf16, int64_t n_img, int64_t HW, int64_t img_stride, int64_t C,
                                         int64_t groups, int silu, void* stream) {
  DFOT_REQUIRE(x && sums && gamma && beta && y_bf16 && n_img > 0 && HW > 0 && C > 0 && groups > 0, DFOT_ERR_INVALID_ARG,
               "groupnorm_apply: bad arguments");
  DFOT_REQUIRE(C % groups == 0 && C % 8 == 0 && img_stride >= HW * C && img_stride % 8 == 0, DFOT_ERR_UNSUPPORTED,
               "groupnorm_apply: need C %% groups == 0, C %% 8 == 0, img_stride %% 8 == 0 and >= HW*C");
  const int vecs = (int)(C / 8);
  DFOT_REQUIRE(vecs <= kThreads && kThreads % vecs == 0 && n_img < 65536 && HW < (1ll << 30), DFOT_ERR_UNSUPPORTED,
               "groupnorm_apply: C/8 = %d must divide %d", vecs, kThreads);
  const float2* stats = reinterpret_cast<const float2*>(sums + 2 * n_img * groups);
  const dim3 grid((unsigned)ceil_div(HW, (kThreads / vecs) * kGnPix * kGnIter), (unsigned)n_img);
  cudaStream_t s = (cudaStream_t)stream;
  if (silu)
    launch_pdl(gn_silu_kernel<float, true>, grid, dim3(kThreads), 0, s, x, stats, gamma, beta, (const float*)nullptr,
               (int64_t)0, (int64_t)0, (int64_t)0, (const __nv_bfloat16*)nullptr, (const int32_t*)nullptr,
               (__nv_bfloat16*)y_bf16, (int)HW, (int)C, (int)groups, img_stride);
  else
    launch_pdl(gn_silu_kernel<float, false>, grid, dim3(kThreads), 0, s, x, stats, gamma, beta, (const float*)nullptr,
               (int64_t)0, (int64_t)0, (int64_t)0, (const __nv_bfloat16*)nullptr, (const int32_t*)nullptr,
               (__nv_bfloat16*)y_bf16, (int)HW, (int)C, (int)groups, img_stride);
  DFOT_CHECK_LAUNCH("groupnorm_apply");
  return DFOT_OK;
}

extern "C" int dfot_rmsnorm_film_bf16(const float* x, const float* weight, float eps, const float* mod_img,
                                      int64_t ld_img, int64_t scale_col, int64_t shift_col, const void* mod_pix,
                                      const int32_t* img_map, void* y_bf16, int64_t M, int64_t D,
                                      int64_t tokens_per_img, void* stream) {
  DFOT_REQUIRE(x && weight && mod_img && y_bf16 && M > 0 && D > 0 && tokens_per_img > 0, DFOT_ERR_INVALID_ARG,
               "rmsnorm_film: bad arguments");
  DFOT_REQUIRE(D % 4 == 0 && D <= 4096 && ld_img % 4 == 0 && scale_col % 4 == 0 && shift_col % 4 == 0,
               DFOT_ERR_UNSUPPORTED, "rmsnorm_film: D (<= 4096) and modulation offsets must be multiples of 4");
  DFOT_REQUIRE((mod_pix == nullptr) == (img_map == nullptr), DFOT_ERR_INVALID_ARG, "rmsnorm_film: mod_pix needs img_map");
  const unsigned grid = (unsigned)ceil_div(M, kNormWarps);
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(NV)                                                                                                   \
  launch_pdl(rmsnorm_film_kernel<NV>, dim3(grid), dim3(kNormWarps * 32), 0, s, x, weight, eps, mod_img, ld_img, scale_col, shift_col,    \
                                                           (const __nv_bfloat16*)mod_pix, img_map,                   \
                                                           (__nv_bfloat16*)y_bf16, M, (int)D, tokens_per_img)
  if (D <= 128) LAUNCH(1);
  else if (D <= 256) LAUNCH(2);
  else if (D <= 640) LAUNCH(5);
  else if (D <= 1152) LAUNCH(9);
  else if (D <= 2048) LAUNCH(16);
  else LAUNCH(32);
#undef LAUNCH
  DFOT_CHECK_LAUNCH("rmsnorm_film");
  return DFOT_OK;
}

extern "C" int dfot_qk_norm_rope(void* qkv, int64_t ld, const float* q_weight, const float* k_weight, float eps,
                                 const float* rope_cs, int64_t tokens_per_sample, int64_t M, int64_t heads,
                                 int64_t head_dim, float q_scale, void* stream) {
  DFOT_REQUIRE(qkv && q_weight && k_weight && rope_cs && M > 0 && heads > 0 && tokens_per_sample > 0,
               DFOT_ERR_INVALID_ARG, "qk_norm_rope: bad arguments");
  DFOT_REQUIRE(head_dim == 64 || head_dim == 128, DFOT_ERR_UNSUPPORTED, "qk_norm_rope: head_dim must be 64 or 128");
  DFOT_REQUIRE(ld % 8 == 0 && ld >= 3 * heads * head_dim && (uintptr_t)qkv % 16 == 0, DFOT_ERR_UNSUPPORTED,
               "qk_norm_rope: ld must be a multiple of 8 and >= 3*heads*head_dim, qkv 16-byte aligned");
  DFOT_REQUIRE(2 * heads * (head_dim / 8) <= 16 * 32, DFOT_ERR_UNSUPPORTED,
               "qk_norm_rope: 2*heads*head_dim = %lld > 4096 unsupported", (long long)(2 * heads * head_dim));
  DFOT_REQUIRE((uintptr_t)rope_cs % 16 == 0 && (uintptr_t)q_weight % 16 == 0 && (uintptr_t)k_weight % 16 == 0,
               DFOT_ERR_UNSUPPORTED, "qk_norm_rope: weights and rope table must be 16-byte aligned");
  const unsigned grid = (unsigned)ceil_div(M, kThreads / 32);
  cudaStream_t s = (cudaStream_t)stream;
  // NI = load instructions per token: 2*heads segments, 32/(head_dim/8) of them per instruction
  const int64_t ni = ceil_div(2 * heads * (head_dim / 8), 32);
#define LAUNCH(DH, NI)                                                                                                 \
  launch_pdl(qk_norm_rope_kernel<DH, NI>, dim3(grid), dim3(kThreads), 0, s, (__nv_bfloat16*)qkv, ld, q_weight, k_weight, eps, rope_cs,     \
                                                        tokens_per_sample, M, (int)heads, q_scale)
  if (head_dim == 64) {
    if (ni <= 3) LAUNCH(64, 3);
    else if (ni <= 5) LAUNCH(64, 5);
    else if (ni <= 8) LAUNCH(64, 8);
    else LAUNCH(64, 16);
  } else {
    if (ni <= 5) LAUNCH(128, 5);
    else if (ni <= 9) LAUNCH(128, 9);
    else LAUNCH(128, 16);
  }
#undef LAUNCH
  DFOT_CHECK_LAUNCH("qk_norm_rope");
  return DFOT_OK;
}

extern "C" int dfot_avgpool2x2(const void* in, int in_dtype, void* out, int out_dtype, int64_t n_img, int64_t H,
                               int64_t W, int64_t C, void* stream) {
  DFOT_REQUIRE(in && out && n_img > 0 && H > 0 && W > 0 && C > 0, DFOT_ERR_INVALID_ARG, "avgpool2x2: bad arguments");
  DFOT_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, DFOT_ERR_UNSUPPORTED, "avgpool2x2: H, W even and C %% 8 == 0");
  const int64_t total = n_img * (H / 2) * (W / 2) * (C / 8);
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = blocks_for(total);
  if (in_dtype == DFOT_F32 && out_dtype == DFOT_F32)
    launch_pdl(avgpool2x2_kernel<float, float>, dim3(grid), dim3(kThreads), 0, s, (const float*)in, (float*)out, n_img, (int)H, (int)W, (int)C);
  else if (in_dtype == DFOT_F32 && out_dtype == DFOT_BF16)
    launch_pdl(avgpool2x2_kernel<float, __nv_bfloat16>, dim3(grid), dim3(kThreads), 0, s, (const float*)in, (__nv_bfloat16*)out, n_img, (int)H, (int)W, (int)C);
  else if (in_dtype == DFOT_BF16 && out_dtype == DFOT_BF16)
    launch_pdl(avgpool2x2_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(kThreads), 0, s, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, n_img, (int)H, (int)W, (int)C);
  else
    DFOT_REQUIRE(false, DFOT_ERR_INVALID_ARG, "avgpool2x2: unsupported dtype combination");
  DFOT_CHECK_LAUNCH("avgpool2x2");
  return DFOT_OK;
}

extern "C" int dfot_sub_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream) {
  DFOT_REQUIRE(a && b && out_bf16 && n > 0, DFOT_ERR_INVALID_ARG, "sub_bf16: bad arguments");
  DFOT_REQUIRE(n % 8 == 0, DFOT_ERR_UNSUPPORTED, "sub_bf16: n must be a multiple of 8");
  launch_pdl(sub_bf16_kernel, dim3(blocks_for(n / 8)), dim3(kThreads), 0, (cudaStream_t)stream, a, b, (__nv_bfloat16*)out_bf16, n / 8);
  DFOT_CHECK_LAUNCH("sub_bf16");
  return DFOT_OK;
}

extern "C" int dfot_upsample2x_add(const float* low, const float* skip, float* out, int64_t n_img, int64_t H, int64_t W,
                                   int64_t C, void* stream) {
  DFOT_REQUIRE(low && skip && out && n_img > 0 && H > 0 && W > 0 && C > 0, DFOT_ERR_INVALID_ARG, "upsample2x_add: bad arguments");
  DFOT_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, DFOT_ERR_UNSUPPORTED, "upsample2x_add: H, W even and C %% 8 == 0");
  launch_pdl(upsample2x_add_kernel, dim3(blocks_for(n_img * H * W * (C / 8))), dim3(kThreads), 0, (cudaStream_t)stream, 
      low, skip, out, n_img, (int)H, (int)W, (int)C);
  DFOT_CHECK_LAUNCH("upsample2x_add");
  return DFOT_OK;
}

extern "C" int dfot_pose_ray_patches(const float* cams, const float* freq_scale, int64_t n_freq, void* out_bf16,
                                     int64_t ld, int64_t frames, int64_t res, int64_t p, void* stream) {
  DFOT_REQUIRE(cams && freq_scale && out_bf16 && n_freq > 0 && frames > 0 && res > 0 && p > 0, DFOT_ERR_INVALID_ARG,
               "pose_ray_patches: bad arguments");
  DFOT_REQUIRE(res % p == 0 && ld >= p * p * 12 * n_freq, DFOT_ERR_INVALID_ARG, "pose_ray_patches: res %% p, ld");
  const size_t smem = (size_t)kPoseThreads * 12 * n_freq * sizeof(__nv_bfloat16);
  DFOT_REQUIRE(smem <= 200 * 1024, DFOT_ERR_UNSUPPORTED, "pose_ray_patches: n_freq %lld too large", (long long)n_freq);
  static size_t reserved = 48 * 1024;
  if (smem > reserved) {
    cudaError_t e = cudaFuncSetAttribute(pose_ray_patches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DFOT_REQUIRE(e == cudaSuccess, DFOT_ERR_CUDA, "pose_ray_patches: cannot reserve %zu B of shared memory", smem);
    reserved = smem;
  }
  const int64_t blocks = (frames * res * res + kPoseThreads - 1) / kPoseThreads;
  launch_pdl(pose_ray_patches_kernel, dim3((unsigned)blocks), dim3(kPoseThreads), smem, (cudaStream_t)stream,
      cams, freq_scale, (int)n_freq, (__nv_bfloat16*)out_bf16, ld, frames, (int)res, (int)p);
  DFOT_CHECK_LAUNCH("pose_ray_patches");
  return DFOT_OK;
}
