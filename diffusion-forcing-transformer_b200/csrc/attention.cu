// K3 — attention over space-time latent tokens (non-causal, unmasked), flash-style:
// Q tile of 128 rows per CTA (8 warps x 16 rows), K/V streamed in 64-key tiles through a
// cp.async double buffer, S = QK^T and O += PV on bf16 tensor-core MMAs with fp32
// accumulation, online softmax in fp32 with exp2 (q is pre-scaled by scale*log2e in the
// QKV-GEMM epilogue, where RoPE-3D is also applied), row max/sum via quad shuffles.
// Reads q/k/v in place from the [tokens, 3*D] QKV matrix; head_dim 72 is zero-padded to 80 in
// shared memory only.  No N x N score matrix ever reaches HBM (the reference materialises it).
#include "common.cuh"

namespace dfot {
namespace attn {

constexpr int BQ = 128, BKV = 64, kThreads = 256;

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 → zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// DH: true head dim (multiple of 8); DP: DH rounded up to a multiple of 16.
template <int DH, int DP>
__global__ void __launch_bounds__(kThreads)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int Ntok, int heads) {
  constexpr int PITCH = DP + 8;                 // elements; +16 B keeps ldmatrix rows on distinct banks
  constexpr int CH = DH / 8;                    // 16-byte chunks per row actually loaded
  constexpr int KS = DP / 16;                   // k-steps for QK^T
  constexpr int ND = DP / 8;                    // 8-wide output column blocks
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sK = sQ + BQ * PITCH;          // 2 stages
  __nv_bfloat16* sV = sK + 2 * BKV * PITCH;     // 2 stages

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, r = blockIdx.z;
  const int D = heads * DH;
  const int64_t ld = 3 * (int64_t)D;
  const __nv_bfloat16* base = qkv + (int64_t)r * Ntok * ld + (int64_t)h * DH;

  // zero the padding columns once (cp.async never touches them)
  if constexpr (DP > DH) {
    constexpr int PADC = DP - DH;
    for (int i = tid; i < (BQ + 4 * BKV) * PADC; i += kThreads) {
      const int row = i / PADC, c = DH + i % PADC;
      sQ[row * PITCH + c] = __float2bfloat16(0.f);  // sQ, sK, sV are contiguous with the same pitch
    }
  }
  auto load_q = [&]() {
    for (int i = tid; i < BQ * CH; i += kThreads) {
      const int row = i / CH, c = i % CH;
      const bool ok = q0 + row < Ntok;
      cp_async_16((uint32_t)__cvta_generic_to_shared(sQ + row * PITCH + c * 8),
                  base + (int64_t)(ok ? q0 + row : 0) * ld + c * 8, ok);
    }
  };
  auto load_kv = [&](int tile, int stage) {
    const int k0 = tile * BKV;
    for (int i = tid; i < BKV * CH; i += kThreads) {
      const int row = i / CH, c = i % CH;
      const bool ok = k0 + row < Ntok;
      const __nv_bfloat16* src = base + (int64_t)(ok ? k0 + row : 0) * ld + c * 8;
      cp_async_16((uint32_t)__cvta_generic_to_shared(sK + (stage * BKV + row) * PITCH + c * 8), src + D, ok);
      cp_async_16((uint32_t)__cvta_generic_to_shared(sV + (stage * BKV + row) * PITCH + c * 8), src + 2 * D, ok);
    }
  };

  const int n_tiles = (Ntok + BKV - 1) / BKV;
  load_q();
  load_kv(0, 0);
  cp_async_commit();
  if (n_tiles > 1) load_kv(1, 1);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();

  // Q fragments stay in registers for the whole kernel
  uint32_t qf[KS][4];
  {
    const int row = warp * 16 + (lane & 15), col = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldsm_x4((uint32_t)__cvta_generic_to_shared(sQ + row * PITCH + ks * 16 + col), qf[ks][0], qf[ks][1], qf[ks][2],
              qf[ks][3]);
  }
  float o[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float row_max[2] = {-INFINITY, -INFINITY}, row_sum[2] = {0.f, 0.f};

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1;
    const __nv_bfloat16* tK = sK + stage * BKV * PITCH;
    const __nv_bfloat16* tV = sV + stage * BKV * PITCH;
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[BKV / 8][4];
#pragma unroll
    for (int nb = 0; nb < BKV / 8; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nb = 0; nb < BKV / 8; nb += 2) {
        uint32_t b0, b1, b2, b3;
        const int krow = nb * 8 + (lane & 7) + ((lane >> 4) << 3), kcol = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4((uint32_t)__cvta_generic_to_shared(tK + krow * PITCH + kcol), b0, b1, b2, b3);
        mma_bf16(s[nb], qf[ks], b0, b1);
        mma_bf16(s[nb + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask keys beyond the sequence (last tile only)
    const int k0 = tile * BKV;
    if (k0 + BKV > Ntok) {
#pragma unroll
      for (int nb = 0; nb < BKV / 8; ++nb) {
        const int key = k0 + nb * 8 + (lane & 3) * 2;
        if (key >= Ntok) s[nb][0] = s[nb][2] = -INFINITY;
        if (key + 1 >= Ntok) s[nb][1] = s[nb][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows g = lane/4 and g + 8)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < BKV / 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float alpha[2], m_use[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float m_new = fmaxf(row_max[i], mx[i]);
      m_use[i] = (m_new == -INFINITY) ? 0.f : m_new;      // fully masked row (cannot happen for valid queries)
      alpha[i] = exp2f(row_max[i] - m_use[i]);
      row_max[i] = m_new;
      row_sum[i] *= alpha[i];
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      o[nd][0] *= alpha[0]; o[nd][1] *= alpha[0];
      o[nd][2] *= alpha[1]; o[nd][3] *= alpha[1];
    }
    uint32_t pf[BKV / 16][4];
#pragma unroll
    for (int nb = 0; nb < BKV / 8; ++nb) {
      const float p0 = exp2f(s[nb][0] - m_use[0]), p1 = exp2f(s[nb][1] - m_use[0]);
      const float p2 = exp2f(s[nb][2] - m_use[1]), p3 = exp2f(s[nb][3] - m_use[1]);
      row_sum[0] += p0 + p1;
      row_sum[1] += p2 + p3;
      pf[nb >> 1][(nb & 1) * 2] = pack_bf16x2(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    // ---- O += P V
#pragma unroll
    for (int j = 0; j < BKV / 16; ++j) {
#pragma unroll
      for (int nd = 0; nd < ND; nd += 2) {
        uint32_t b0, b1, b2, b3;
        const int vrow = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, vcol = nd * 8 + ((lane >> 4) << 3);
        ldsm_x4_t((uint32_t)__cvta_generic_to_shared(tV + vrow * PITCH + vcol), b0, b1, b2, b3);
        mma_bf16(o[nd], pf[j], b0, b1);
        mma_bf16(o[nd + 1], pf[j], b2, b3);
      }
    }
    // ---- advance the K/V ring
    __syncthreads();                      // everyone is done reading this stage
    if (tile + 2 < n_tiles) load_kv(tile + 2, stage);
    cp_async_commit();
    cp_async_wait<1>();                   // tile + 1 has landed
    __syncthreads();
  }

  // ---- finalize: O /= l, write bf16
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    row_sum[i] += __shfl_xor_sync(0xffffffffu, row_sum[i], 1);
    row_sum[i] += __shfl_xor_sync(0xffffffffu, row_sum[i], 2);
  }
  const float inv0 = 1.f / row_sum[0], inv1 = 1.f / row_sum[1];
  const int g = lane >> 2, cpair = (lane & 3) * 2;
  const int qa = q0 + warp * 16 + g, qb = qa + 8;
  __nv_bfloat16* obase = out + (int64_t)r * Ntok * D + (int64_t)h * DH;
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    const int c = nd * 8 + cpair;
    if (c < DH) {
      if (qa < Ntok)
        *reinterpret_cast<uint32_t*>(obase + (int64_t)qa * D + c) = pack_bf16x2(o[nd][0] * inv0, o[nd][1] * inv0);
      if (qb < Ntok)
        *reinterpret_cast<uint32_t*>(obase + (int64_t)qb * D + c) = pack_bf16x2(o[nd][2] * inv1, o[nd][3] * inv1);
    }
  }
}

template <int DH, int DP>
static int launch(const void* qkv, void* out, int64_t R, int64_t Ntok, int64_t heads, cudaStream_t s) {
  constexpr int smem_bytes = (BQ + 4 * BKV) * (DP + 8) * 2;
  auto kern = attention_kernel<DH, DP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    DFOT_REQUIRE(e == cudaSuccess, DFOT_ERR_CUDA, "attention: cannot reserve %d B shared memory: %s", smem_bytes,
                 cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(Ntok, BQ), (unsigned)heads, (unsigned)R);
  kern<<<grid, kThreads, smem_bytes, s>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, (int)Ntok, (int)heads);
  DFOT_CHECK_LAUNCH("attention");
  return DFOT_OK;
}

}  // namespace attn
}  // namespace dfot

extern "C" int dfot_attention(const void* qkv, void* out, int64_t R, int64_t Ntok, int64_t heads, int64_t head_dim,
                              void* stream) {
  using namespace dfot;
  DFOT_REQUIRE(qkv && out && R > 0 && Ntok > 0 && heads > 0, DFOT_ERR_INVALID_ARG, "attention: bad arguments");
  DFOT_REQUIRE(R <= 65535 && heads <= 65535 && Ntok < (1 << 30), DFOT_ERR_UNSUPPORTED,
               "attention: grid limits exceeded");
  DFOT_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 4 == 0), DFOT_ERR_UNSUPPORTED,
               "attention: qkv must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (head_dim) {
    case 64: return attn::launch<64, 64>(qkv, out, R, Ntok, heads, s);
    case 72: return attn::launch<72, 80>(qkv, out, R, Ntok, heads, s);
    case 128: return attn::launch<128, 128>(qkv, out, R, Ntok, heads, s);
  }
  set_error("attention: head_dim %lld unsupported (64, 72, 128)", (long long)head_dim);
  return DFOT_ERR_UNSUPPORTED;
}
