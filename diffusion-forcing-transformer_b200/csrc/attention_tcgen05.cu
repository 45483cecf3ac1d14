// K3 — attention over space-time latent tokens on the 5th-gen tensor cores (non-causal, unmasked).
//
// Persistent, warp-specialised flash attention; one work item = (sample r, head h, 128-query tile):
//   warp 0    : TMA producer — Q tile once per item, K / V tiles of 128 keys through two 2-stage rings.
//               q/k/v are read in place from the [tokens, 3, heads, d] QKV matrix via a 3-D tensor map whose
//               innermost extent is the true head_dim, so head_dim 72 is zero-padded to 80 by TMA OOB fill.
//   warp 1    : MMA issuer  — S_j = Q·K_jᵀ (128x128xd, both operands K-major) into a double-buffered TMEM
//               score tile, then PV_j = P_j·V_j (128 x d x 128; P K-major from shared memory, V MN-major
//               straight from its row-major tile — no transpose) into a double-buffered TMEM tile.
//   warps 2-9 : softmax + accumulate.  Two warpgroups split every 128x128 score tile by key halves; a thread
//               owns (query row = TMEM lane, 64 keys), so the row max needs ONE exchange with its partner thread
//               (via shared memory + a 256-thread named barrier) and no shuffles.  p = exp2(s - m) in fp32 (q is
//               pre-scaled by scale·log2e and RoPE-rotated in the QKV-GEMM epilogue), P written as bf16 in the
//               UMMA 128B-swizzled layout; the running output O lives in registers (each warpgroup keeps half
//               of the head_dim columns) and is updated as O = O·alpha + PV_j one tile behind, so the PV MMA of
//               tile j overlaps the softmax of tile j+1.  Two warps per scheduler hide TMEM/MUFU latency.
// The N x N score matrix never leaves the SM (the reference materialises it in HBM: dit_blocks.py:21-44).
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace dfot {
namespace fattn {

constexpr int BQ = 128, BKV = 128;
constexpr int kSoftmaxWarps = 8;
constexpr int kThreads = 64 + 32 * kSoftmaxWarps;   // TMA warp, MMA warp, 2 softmax warpgroups
constexpr int kAtomBytes = 128 * 128;   // one 128-row x 128-byte swizzle plane (64 bf16 wide)
constexpr uint32_t kSuspendHintNs = 2000;

// Debug build only (-DDFOT_ATTN_TRACE, scripts/attn_trace.py): per-warp clock64 stamps of the phase boundaries of CTA 0
#ifdef DFOT_ATTN_TRACE
__device__ unsigned long long g_trace[16 * 2048];
#define DFOT_TRACE_DECL uint32_t tr_n = 0
#define DFOT_TRACE(tag)                                                                                    \
  do {                                                                                                     \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < 2048)                                         \
      g_trace[(threadIdx.x >> 5) * 2048 + tr_n++] = ((unsigned long long)clock64() << 8) | (unsigned)(tag); \
  } while (0)
#else
#define DFOT_TRACE_DECL
#define DFOT_TRACE(tag)
#endif

// ------------------------------------------------------------------ PTX wrappers (same idioms as the GEMM)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(kSuspendHintNs)   // sleep in hardware instead of spinning on issue slots
        : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u) {   // bounded: a protocol bug must trap, never hang the GPU
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("dfot_attention: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
               threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
// Latency-critical hand-offs (softmax <-> MMA issuer): plain try_wait polling, no suspend-time hint.
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
#ifdef DFOT_ATTN_SUSPEND_WAITS
  mbar_wait(bar, parity);
  return;
#endif
  uint32_t done = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 28)) {   // bounded: a protocol bug must trap, never hang the GPU
      printf("dfot_attention: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (P, bf16 pairs packed in 32-bit columns, lane = row) is read
// straight from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32x2 arithmetic (sm_100: one issue slot for two lanes of work)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 2^x for a pair of fp32 values WITHOUT the MUFU: round-to-nearest split x = j + f (|f| <= 0.5) by the 1.5*2^23 trick,
// degree-3 minimax polynomial for 2^f (max relative error 7.5e-5, far below the bf16 rounding of P), and the exponent
// added with integer arithmetic.  All fp32 work is packed f32x2 (FMA pipe); used for a fraction of the scores so that
// the 16/clk/SM MUFU and the FMA pipe work side by side (the exp2 rate is what bounds head_dim-64 attention).
__device__ __forceinline__ void ex2_poly_x2(float xa, float xb, float& pa, float& pb) {
  const uint64_t magic = pack_f32x2(12582912.f, 12582912.f), neg_magic = pack_f32x2(-12582912.f, -12582912.f);
  const uint64_t x = pack_f32x2(fmaxf(xa, -125.f), fmaxf(xb, -125.f));
  const uint64_t t = add_f32x2(x, magic);                         // integer part in the low mantissa bits
  const uint64_t f = fma_f32x2(add_f32x2(t, neg_magic), pack_f32x2(-1.f, -1.f), x);
  uint64_t p = fma_f32x2(f, pack_f32x2(0.05517168f, 0.05517168f), pack_f32x2(0.24261113f, 0.24261113f));
  p = fma_f32x2(p, f, pack_f32x2(0.69326097f, 0.69326097f));
  p = fma_f32x2(p, f, pack_f32x2(0.99992806f, 0.99992806f));
  float ta, tb, qa, qb;
  unpack_f32x2(t, ta, tb);
  unpack_f32x2(p, qa, qb);
  pa = __int_as_float(__float_as_int(qa) + (__float_as_int(ta) << 23));
  pb = __int_as_float(__float_as_int(qb) + (__float_as_int(tb) << 23));
}

// same for bounded arguments (|x| <= ~100: no clamp needed)
__device__ __forceinline__ void ex2_poly_x2_bounded(float xa, float xb, float& pa, float& pb) {
  const uint64_t magic = pack_f32x2(12582912.f, 12582912.f), neg_magic = pack_f32x2(-12582912.f, -12582912.f);
  const uint64_t x = pack_f32x2(xa, xb);
  const uint64_t t = add_f32x2(x, magic);
  const uint64_t f = fma_f32x2(add_f32x2(t, neg_magic), pack_f32x2(-1.f, -1.f), x);
  uint64_t p = fma_f32x2(f, pack_f32x2(0.05517168f, 0.05517168f), pack_f32x2(0.24261113f, 0.24261113f));
  p = fma_f32x2(p, f, pack_f32x2(0.69326097f, 0.69326097f));
  p = fma_f32x2(p, f, pack_f32x2(0.99992806f, 0.99992806f));
  float ta, tb, qa, qb;
  unpack_f32x2(t, ta, tb);
  unpack_f32x2(p, qa, qb);
  pa = __int_as_float(__float_as_int(qa) + (__float_as_int(ta) << 23));
  pb = __int_as_float(__float_as_int(qb) + (__float_as_int(tb) << 23));
}

// UMMA shared-memory descriptors (cute::UMMA::SmemDescriptor), 128-byte swizzle, version 1.
//   K-major  operand: 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major operand: K rows are 128-byte lines, 8-row groups 1024 B apart (SBO); 64-element MN atoms LBO apart.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, M=128; b_mn selects MN-major B
__device__ __forceinline__ constexpr uint32_t make_idesc(int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

struct Params {
  __nv_bfloat16* out;
  int64_t ld_out;   // row stride of `out` in elements (>= heads*head_dim): lets the output land inside a wider buffer
  int no_max;       // scores are bounded (|s| <= ~96 in log2 units, e.g. QK-normalised attention): p = 2^s needs no running
                    // maximum, no subtraction and no rescaling — mathematically identical softmax, a third fewer instructions
  int R, Ntok, heads, q_tiles, kv_tiles, num_items;
  int pair_items;   // kernel 2: items [0, pair_items) are query-tile PAIRS; the rest are SINGLE tiles, two per pair of the
                    // last (partial) wave — the tail of the persistent loop then costs half a wave instead of a whole one
};

// DH: true head dim; DP: head dim padded to a multiple of 16 (MMA K / N granularity)
template <int DH, int DP>
__global__ void __launch_bounds__(kThreads, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh (pdl_wait() follows the prologue)
  constexpr int ATOMS = (DP + 63) / 64;              // 64-wide swizzle planes per Q/K/V tile
  constexpr int TILE_BYTES = ATOMS * kAtomBytes;     // Q, K or V tile
  constexpr int P_BYTES = 2 * kAtomBytes;            // 128 rows x 128 keys bf16
  constexpr int KS_QK = DP / 16, KS_PV = BKV / 16;
  constexpr uint32_t IDESC_S = make_idesc(BKV, false);
  constexpr uint32_t IDESC_PV = make_idesc(DP, true);
  constexpr int TMEM_S0 = 0, TMEM_PV0 = 256;         // S: 2 x 128 cols, PV: 2 x 128 cols

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);          // 128B-swizzle atoms need 1024-byte alignment
  if ((base & 1023u) != 0) {
    if (threadIdx.x == 0) printf("dfot_attention: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t sQ = base;
  const uint32_t sK = sQ + TILE_BYTES;               // 2 stages
  const uint32_t sV = sK + 2 * TILE_BYTES;           // 2 stages
  const uint32_t sP = sV + 2 * TILE_BYTES;           // 2 buffers
  const uint32_t bars = sP + 2 * P_BYTES;
  enum { Q_FULL = 0, Q_EMPTY = 1, K_FULL = 2, K_EMPTY = 4, V_FULL = 6, V_EMPTY = 8, S_FULL = 10, P_FULL = 12,
         PV_DONE = 14, O_EMPTY = 16, N_BARS = 18 };
  auto bar = [&](int id) { return bars + 8u * id; };
  const uint32_t tmem_slot = bars + 8u * N_BARS;
  const uint32_t s_xchg = bars + 160u;               // float [2 (parity)][2 (warpgroup)][128 rows]: max / sum exchange

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;   // (uniform for the compiler: see elect_one_sync)
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int i = 0; i < N_BARS; ++i) {
      const bool from_softmax = (i >= P_FULL && i < P_FULL + 2) || (i >= O_EMPTY && i < O_EMPTY + 2);
      mbar_init(bar(i), from_softmax ? kSoftmaxWarps : 1);   // one arrival per softmax warp, else one producer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();   // barriers and tensor memory are set up; global memory is only touched from here on

  const int n_kv = p.kv_tiles;
  auto item_coord = [&](int item, int& r, int& h, int& qt) {
    qt = item % p.q_tiles;                           // consecutive items share (r, h): K/V stay hot in L2
    const int rh = item / p.q_tiles;
    h = rh % p.heads;
    r = rh / p.heads;
  };

  if (warp == 0) {
    // ===================== TMA producer (all lanes walk the loop; the copies are issued under elect.sync) ===========
    {
      uint32_t g = 0, it = 0;                        // global KV-tile counter, item counter
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        int r, h, qt;
        item_coord(item, r, h, qt);
        const int row0 = r * p.Ntok;
        mbar_wait(bar(Q_EMPTY), (it & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(bar(Q_FULL), TILE_BYTES);
#pragma unroll
          for (int a = 0; a < ATOMS; ++a) tma_load_3d(sQ + a * kAtomBytes, &tmap, bar(Q_FULL), a * 64, h, row0 + qt * BQ);
        }
        __syncwarp();
        for (int j = 0; j < n_kv; ++j, ++g) {
          const uint32_t st = g & 1u, ph = (g >> 1) & 1u;
          mbar_wait(bar(K_EMPTY + st), ph ^ 1u);
          if (elect_one_sync()) {
            mbar_expect_tx(bar(K_FULL + st), TILE_BYTES);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a)
              tma_load_3d(sK + st * TILE_BYTES + a * kAtomBytes, &tmap, bar(K_FULL + st), a * 64, p.heads + h,
                          row0 + j * BKV);
          }
          __syncwarp();
          mbar_wait(bar(V_EMPTY + st), ph ^ 1u);
          if (elect_one_sync()) {
            mbar_expect_tx(bar(V_FULL + st), TILE_BYTES);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a)
              tma_load_3d(sV + st * TILE_BYTES + a * kAtomBytes, &tmap, bar(V_FULL + st), a * 64, 2 * p.heads + h,
                          row0 + j * BKV);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, tcgen05 instructions under elect.sync) =====================
    {
      uint32_t g = 0, it = 0;
      auto issue_s = [&](uint32_t gt, bool last_of_item) {       // S = Q · K^T for global tile gt
        const uint32_t st = gt & 1u, ph = (gt >> 1) & 1u;
        mbar_wait(bar(K_FULL + st), ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d = tmem_base + TMEM_S0 + st * 128u;
#pragma unroll
          for (int s = 0; s < KS_QK; ++s) {
            const uint32_t off = (uint32_t)(s >> 2) * kAtomBytes + (uint32_t)(s & 3) * 32u;
            umma_bf16(d, desc_kmajor(sQ + off), desc_kmajor(sK + st * TILE_BYTES + off), IDESC_S, s > 0 ? 1u : 0u);
          }
          umma_commit(bar(K_EMPTY + st));
          umma_commit(bar(S_FULL + st));
          if (last_of_item) umma_commit(bar(Q_EMPTY));
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        mbar_wait(bar(Q_FULL), it & 1u);
        tc_fence_after();
        // S buffer st was last read by the softmax of tile g-2, which preceded the PV MMA issued for it: free.
        issue_s(g, n_kv == 1);
        if (n_kv > 1) issue_s(g + 1, n_kv == 2);
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t gt = g + j, st = gt & 1u, ph = (gt >> 1) & 1u;
          mbar_wait(bar(V_FULL + st), ph);
          mbar_wait(bar(O_EMPTY + st), ph ^ 1u);     // softmax consumed PV of tile gt-2
          mbar_wait(bar(P_FULL + st), ph);           // P_j is in shared memory (and S_j has been read)
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t d = tmem_base + TMEM_PV0 + st * 128u;
#pragma unroll
            for (int s = 0; s < KS_PV; ++s) {
              const uint32_t a_off = (uint32_t)(s >> 2) * kAtomBytes + (uint32_t)(s & 3) * 32u;   // 16 keys along K
              umma_bf16(d, desc_kmajor(sP + st * P_BYTES + a_off),
                        desc_mnmajor(sV + st * TILE_BYTES + (uint32_t)s * 2048u, kAtomBytes), IDESC_PV, s > 0 ? 1u : 0u);
            }
            umma_commit(bar(V_EMPTY + st));
            umma_commit(bar(PV_DONE + st));
          }
          __syncwarp();
          if (j + 2 < n_kv) issue_s(gt + 2, j + 3 == n_kv);
        }
        g += n_kv;
      }
    }
  } else {
    // ===================== softmax + accumulate (thread = query row x 64 keys) =====================
    constexpr int HC = DP / 2;                       // output columns owned by this thread
    const int q = warp & 3;                          // TMEM lane quarter (hardware rule: warp_id % 4)
    const int wg = (warp - 2) >> 2;                  // warpgroup: which half of the keys / output columns
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    auto xchg = [&](uint32_t par, int g_, int r_) { return s_xchg + 4u * (uint32_t)((par * 2 + g_) * 128 + r_); };
    auto pair_barrier = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    uint32_t g = 0, n_xchg = 0;                      // exchange slots alternate so one barrier per exchange suffices
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int r, h, qt;
      item_coord(item, r, h, qt);
      float o[HC];
#pragma unroll
      for (int c = 0; c < HC; ++c) o[c] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

      auto accumulate = [&](uint32_t gt, float alpha) {          // O = O*alpha + PV(gt), own column half
        const uint32_t st = gt & 1u, ph = (gt >> 1) & 1u;
        mbar_wait(bar(PV_DONE + st), ph);
        tc_fence_after();
        const uint32_t t_pv = t_lane + TMEM_PV0 + st * 128u + (uint32_t)(wg * HC);
        uint32_t pv[HC / 8][8];
#pragma unroll
        for (int c8 = 0; c8 < HC / 8; ++c8) tmem_ld_x8(t_pv + 8 * c8, pv[c8]);   // all loads in flight, one wait
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(O_EMPTY + st));             // TMEM tile is free again
#pragma unroll
        for (int c = 0; c < HC; ++c) o[c] = o[c] * alpha + __uint_as_float(pv[c >> 3][c & 7]);
      };

      for (int j = 0; j < n_kv; ++j) {
        const uint32_t gt = g + j, st = gt & 1u, ph = (gt >> 1) & 1u;
        mbar_wait(bar(S_FULL + st), ph);
        tc_fence_after();
        const uint32_t t_s = t_lane + TMEM_S0 + st * 128u + (uint32_t)(wg * 64);
        const int valid = p.Ntok - j * BKV - wg * 64;   // keys of this thread's half that exist (may be <= 0)
        uint32_t v[2][32];                              // the 64 scores of this thread stay in registers
        tmem_ld_x32(t_s, v[0]);
        tmem_ld_x32(t_s + 32, v[1]);
        tmem_ld_wait();
        if (valid < 64) {                               // partial tile: mask once, in registers
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= valid) v[c >> 5][c & 31] = 0xff800000u;   // -inf
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; ++c) mx = fmaxf(mx, __uint_as_float(v[c >> 5][c & 31]));
        // one exchange with the partner thread (same row, other key half)
        const uint32_t par = n_xchg++ & 1u;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(xchg(par, wg, row)), "f"(mx) : "memory");
        pair_barrier();
        float mx_other;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mx_other) : "r"(xchg(par, wg ^ 1, row)));
        const float m_new = fmaxf(m_run, fmaxf(mx, mx_other));
        const float alpha = ex2(m_run - m_new);      // 0 on the first tile (m_run = -inf)
        // p = exp2(s - m), partial row sum, bf16 P into plane `wg` of the 128B-swizzled K-major tile
        float sum0 = 0.f, sum1 = 0.f;
        const uint32_t p_row = sP + st * P_BYTES + (uint32_t)wg * kAtomBytes + (uint32_t)row * 128u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {                // 16-byte unit = 8 keys
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 8 * u + 2 * e;
            const float p0 = ex2(__uint_as_float(v[c >> 5][c & 31]) - m_new);
            const float p1 = ex2(__uint_as_float(v[(c + 1) >> 5][(c + 1) & 31]) - m_new);
            sum0 += p0;
            sum1 += p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
          const uint32_t addr = p_row + (uint32_t)((u ^ (row & 7)) << 4);   // swizzle: unit ^= row % 8
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                       "r"(pk[3])
                       : "memory");
        }
        l_run = l_run * alpha + (sum0 + sum1);
        m_run = m_new;
        // make the generic-proxy writes of P visible to the tensor core (async proxy), then signal
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(P_FULL + st));
        if (j > 0) accumulate(gt - 1, alpha_prev);   // one tile behind: overlaps the PV MMA with this softmax
        alpha_prev = alpha;
      }
      accumulate(g + n_kv - 1, alpha_prev);
      // total row sum = own half + partner's half (both are relative to the same running maximum)
      const uint32_t par = n_xchg++ & 1u;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xchg(par, wg, row)), "f"(l_run) : "memory");
      pair_barrier();
      float l_other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l_other) : "r"(xchg(par, wg ^ 1, row)));
      g += n_kv;

      const int qrow = qt * BQ + row;
      if (qrow < p.Ntok) {
        const float inv = 1.f / (l_run + l_other);
        __nv_bfloat16* dst = p.out + ((int64_t)r * p.Ntok + qrow) * p.ld_out + (int64_t)h * DH + wg * HC;
#pragma unroll
        for (int c = 0; c < HC; c += 8) {
          if (wg * HC + c < DH) {
            uint4 w;
            w.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv);
            w.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
            w.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv);
            w.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
            *reinterpret_cast<uint4*>(dst + c) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// =====================================================================================================================
// Kernel 2 — two query tiles per CTA, FA4-style role split (used whenever a sample has more than one query tile):
//   warp 0    : TMA producer — both 128-row Q tiles once per work item, K / V tiles through 2-stage rings
//   warp 1    : MMA issuer  — S_t = Q_t·K_jᵀ into TMEM; O_t += P_t·V_j with the A operand P_t read FROM TENSOR MEMORY
//               (P aliases the first 64 columns of its S tile as packed bf16 pairs) and O_t accumulated in TMEM.
//               Issue order  S0(0) S1(0) | PV0(j) S0(j+1) PV1(j) S1(j+1) ...  — tensor-core ops execute in issue order,
//               which is what makes the S/P aliasing safe without extra barriers.
//   warps 2-3 : idle (they pad warpgroup 0 so that the register reallocation below is warpgroup-aligned)
//   warps 4-7 : softmax warpgroup of Q tile 0;  warps 8-11 : softmax warpgroup of Q tile 1.  A thread owns a whole
//               query row (128 scores in registers): row max / row sum need no cross-thread exchange at all, and the
//               two warpgroups run half a tile out of phase, so the MUFU-bound exp2 phase of one overlaps the
//               TMEM-load / max / store phases of the other and both overlap the MMAs.
//   Lazy rescaling: the running maximum used for exp2 is only raised when a tile's maximum exceeds it by more than 8
//               (log2 units; p <= 256 is harmless in fp32/bf16), so the O_t tile in TMEM is rescaled
//               (tcgen05.ld → mul → tcgen05.st, by the row's own thread) a handful of times per row instead of once
//               per KV tile; l and O always share the same reference maximum, so the result is exact.
//   head_dim <= 64 (SEP_P): tensor memory has room for P_t outside S_t (S 2x128 | P 2x64 | O 2x64 columns), so
//               S_t(j+1) no longer has to wait for PV_t(j): it is issued as soon as the softmax warpgroup has pulled
//               S_t(j) into registers (S_FREE) and runs under the exp2 phase — the softmax never waits for the MMAs.
//   Registers: setmaxnreg moves the register budget from the TMA/MMA warpgroup (96) to the softmax warpgroups (200),
//               so the 128 scores + packing temporaries of a row never spill.
// With head_dim 64 the kernel is bound by the 16/clk/SM MUFU (exp2) rate, not the tensor pipe: 128x128 exps = 1024
// cycles vs 512 cycles of MMA per tile — the roofline is ~50 % of the bf16 tensor peak.
constexpr int kThreads2 = 384;
// Keys per KV tile of kernel 2.  The separate-P pipeline needs 2 x (KV + KV/2 + DP) <= 512 tensor-memory columns: 128
// keys fit for head_dim 64; head_dim 72 (padded to 80) takes 112-key tiles (496 columns) — 12 instead of 10 tiles at
// N = 1280, but S_t(j+1) no longer waits for PV_t(j) (the aliased layout spent 44 % of its softmax time waiting for S);
// head_dim 128 has no room for it at any useful width and keeps the aliased layout.
template <int DP> struct KvTile { static constexpr int value = DP == 80 ? 112 : 128; };
// NOMAX: bounded scores, p = 2^s without a running maximum (Params::no_max); KV: keys per tile (KvTile<DP>)
template <int DH, int DP, bool NOMAX, int KV>
__global__ void __launch_bounds__(kThreads2, 1)
attention2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_kv,
                          const Params p) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh (pdl_wait() follows the prologue)
  constexpr int BKV = KV;                            // (shadows the 128 of kernel 1)
  static_assert(KV % 16 == 0 && KV <= 128 && KV > 96, "KV tile: three full 32-score chunks plus a 16- or 32-wide one");
  constexpr int NCH = (KV + 31) / 32;                // 32-score chunks of a row; the last one may hold 16
  constexpr int ATOMS = (DP + 63) / 64;
  constexpr int TILE_BYTES = ATOMS * kAtomBytes;     // smem slot of a Q / K / V tile (K / V fill KV of its 128 rows)
  constexpr int KV_BYTES = ATOMS * KV * 128;         // bytes one K or V tile brings in
  constexpr int KS_QK = DP / 16, KS_PV = BKV / 16;
  constexpr uint32_t IDESC_S = make_idesc(BKV, false);
  constexpr uint32_t IDESC_PV = make_idesc(DP, true);
  constexpr bool SEP_P = 2 * (KV + KV / 2 + DP) <= 512;   // P_t has its own TMEM columns
  constexpr bool DUAL = SEP_P;                       // one MMA issuer warp per query tile (measured: helps only SEP_P)
  // aliased: S_t at 128*t (P_t = its first 64 columns), O_t at 256 + 128*t;
  // separate: S_t at KV*t, P_t at 2*KV + (KV/2)*t, O_t at 3*KV + DP*t
  constexpr uint32_t TMEM_S = 0, S_STRIDE = SEP_P ? KV : 128;
  constexpr uint32_t TMEM_P = SEP_P ? 2 * KV : 0, P_STRIDE = SEP_P ? KV / 2 : 128;
  constexpr uint32_t TMEM_O = SEP_P ? 3 * KV : 256, O_STRIDE = SEP_P ? DP : 128;
  static_assert(SEP_P || KV == 128, "the aliased layout is written for 128-key tiles");
  constexpr float kRescaleThreshold = 8.0f;
  // which of the 16 score pairs of every 32-score group go through the polynomial exp2 (bit i: iteration i of 8).
  // Measured on B200 (d = 64, N = 8192): 0 % poly 786 TFLOP/s, 25 % 753, 31 % 774, 37.5 % 746, 50 % 693 — the kernel is
  // issue/latency-bound before it is MUFU-bound (XU pipe 70 % busy), so extra FMA-pipe instructions do not pay: off.
#ifndef DFOT_ATTN_POLY_FIRST
#define DFOT_ATTN_POLY_FIRST 0x00
#define DFOT_ATTN_POLY_SECOND 0x55
#endif
#ifndef DFOT_ATTN_POLY_ALIASED
#define DFOT_ATTN_POLY_ALIASED 1
#endif
  constexpr bool kPolyOn = SEP_P || DFOT_ATTN_POLY_ALIASED;
  constexpr uint32_t kPolyFirst = kPolyOn ? DFOT_ATTN_POLY_FIRST : 0u, kPolySecond = kPolyOn ? DFOT_ATTN_POLY_SECOND : 0u;
  // bounded-score path (2 instead of 3 base instructions per score): polynomial share of the score pairs.  Measured
  // on B200 (d = 64, N = 8192, R = 8; run-to-run spread of one box ~3 %): 0 % 824, 12.5 % 853, 19 % 874, 25 % 884-918,
  // 31 % 857, 37.5 % 886, 44 % 874, 50 % 864-877, 62.5 % 867, 75 % 813 TFLOP/s — flat between 25 % and 50 %, so the
  // smaller share is used (fewer FMA-pipe instructions = less power under the cap).  Also measured and dropped: two
  // softmax warpgroups per query tile, each owning 64 of the 128 score columns (4 warps per scheduler; needs
  // setmaxnreg 64/104 because a CTA can only re-use registers its own warps released): 845 vs 870 TFLOP/s.
#ifndef DFOT_ATTN_POLYB_FIRST
#define DFOT_ATTN_POLYB_FIRST 0x00
#define DFOT_ATTN_POLYB_SECOND 0x55
#endif
  constexpr uint32_t kPolyFirstB = kPolyOn ? DFOT_ATTN_POLYB_FIRST : 0u, kPolySecondB = kPolyOn ? DFOT_ATTN_POLYB_SECOND : 0u;
  // Measured and dropped (aliased layout, d = 72 / 128): starting tile 1 one softmax phase behind tile 0, so that one
  // warpgroup exponentiates while the other tile's PV + S MMAs run — 505 vs 522 TFLOP/s at d = 72 (N = 1280), 958 vs 1003
  // at d = 128 (N = 2048): the hand-off latencies, not the phase of the two tiles, bound these shapes
  // (profiles/r02_ncu_attn72.txt: 44 % of the softmax warps' samples wait for S_FULL with the tensor pipe 27 % busy).
#ifndef DFOT_ATTN_PV_WAIT_CHUNK
#define DFOT_ATTN_PV_WAIT_CHUNK -1
#endif
  // 32-score chunk (0..3) after whose exps PV_DONE is awaited.  Measured on B200 (d = 64, N = 8192, R = 8), chunk 0 / 1 /
  // 2 / 3: bounded-score path 903 / 901 / 900 / 824 TFLOP/s (its exp2 phase is short enough that the stores at the end
  // lengthen the tile), running-max path 788 / 775 / 761 / 844 (there the wait at the very end never stalls).
  constexpr int kPvWaitChunk = DFOT_ATTN_PV_WAIT_CHUNK >= 0 ? DFOT_ATTN_PV_WAIT_CHUNK : (NOMAX ? 0 : 3);
  // Phase stagger of the two query tiles (SEP_P only).  The warps of tile 0 and tile 1 that share a scheduler also share
  // its MUFU; started together they stay in lockstep — both in the exp2 phase (each at half the MUFU rate), then both in
  // the MUFU-free phase (TMEM load, barrier hand-offs, P stores) — and an offset between them, once there, persists
  // (profiles/r02_ncu_attn64_stalls.txt: exp2 phase = 2 x its solo length, XU pipe idle a third of the time).  So tile 1
  // starts every work item one exp2 phase late: its first tile waits until tile 0 has published its first P.
#ifndef DFOT_ATTN_STAGGER
#define DFOT_ATTN_STAGGER 0
#endif
  constexpr int kStagger = SEP_P ? DFOT_ATTN_STAGGER : 0;   // 0 off, 1 after tile 0's first P, 2 after half of it

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) {
    if (threadIdx.x == 0) printf("dfot_attention: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t sQ = base;                          // 2 tiles
  const uint32_t sK = sQ + 2 * TILE_BYTES;           // 2 stages
  const uint32_t sV = sK + 2 * TILE_BYTES;           // 2 stages
  const uint32_t bars = sV + 2 * TILE_BYTES;
  enum { Q_FULL = 0, Q_EMPTY = 1, K_FULL = 2, K_EMPTY = 4, V_FULL = 6, V_EMPTY = 8, S_FULL = 10, P_FULL = 12,
         O_DONE = 14, O_FREE = 16, S_FREE = 18, PV_DONE = 20, STAGGER = 22, N_BARS = 23 };
  auto bar = [&](int id) { return bars + 8u * id; };
  const uint32_t tmem_slot = bars + 8u * N_BARS;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;   // (uniform for the compiler: see elect_one_sync)
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_kv) : "memory");
    for (int i = 0; i < N_BARS; ++i) {
      const bool from_softmax = (i >= P_FULL && i < P_FULL + 2) || (i >= O_FREE && i < O_FREE + 2) ||
                                (i >= S_FREE && i < S_FREE + 2) || i == STAGGER;
      const bool from_both_issuers = i == Q_EMPTY || (i >= K_EMPTY && i < K_EMPTY + 2) || (i >= V_EMPTY && i < V_EMPTY + 2);
      // softmax → one arrival per warp of the tile's warpgroup; ring slots → one per MMA issuer; else one producer
      mbar_init(bar(i), from_softmax ? 4 : ((DUAL && from_both_issuers) ? 2 : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();   // barriers and tensor memory are set up; global memory is only touched from here on

  const int n_kv = p.kv_tiles;
  const int q_pairs = (p.q_tiles + 1) >> 1;
  // item -> (sample r, head h, first query tile q0, number of query tiles nt in {1, 2})
  auto item_coord = [&](int item, int& r, int& h, int& q0, int& nt) {
    int pair = item, sel = 0;
    nt = 2;
    if (item >= p.pair_items) {                      // split tail: two single-tile items per pair
      const int k = item - p.pair_items;
      pair = p.pair_items + (k >> 1);
      sel = k & 1;
      nt = 1;
    }
    q0 = 2 * (pair % q_pairs) + sel;                 // consecutive items share (r, h): K/V stay hot in L2
    const int rh = pair / q_pairs;
    h = rh % p.heads;
    r = rh / p.heads;
  };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
  if (warp == 0) {
    // ===================== TMA producer (all lanes walk the loop; the copies are issued under elect.sync) ===========
    {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        int r, h, q0, nt;
        item_coord(item, r, h, q0, nt);
        const int row0 = r * p.Ntok;
        mbar_wait(bar(Q_EMPTY), (it & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(bar(Q_FULL), 2 * TILE_BYTES);
#pragma unroll
          for (int t = 0; t < 2; ++t)   // (the second slot of a single-tile item is loaded too and simply not used)
#pragma unroll
            for (int a = 0; a < ATOMS; ++a)
              tma_load_3d(sQ + t * TILE_BYTES + a * kAtomBytes, &tmap, bar(Q_FULL), a * 64, h, row0 + (q0 + t) * BQ);
        }
        __syncwarp();
        for (int j = 0; j < n_kv; ++j, ++g) {
          const uint32_t st = g & 1u, ph = (g >> 1) & 1u;
          mbar_wait(bar(K_EMPTY + st), ph ^ 1u);
          if (elect_one_sync()) {
            mbar_expect_tx(bar(K_FULL + st), KV_BYTES);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a)
              tma_load_3d(sK + st * TILE_BYTES + a * kAtomBytes, &tmap_kv, bar(K_FULL + st), a * 64, p.heads + h,
                          row0 + j * BKV);
          }
          __syncwarp();
          mbar_wait(bar(V_EMPTY + st), ph ^ 1u);
          if (elect_one_sync()) {
            mbar_expect_tx(bar(V_FULL + st), KV_BYTES);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a)
              tma_load_3d(sV + st * TILE_BYTES + a * kAtomBytes, &tmap_kv, bar(V_FULL + st), a * 64, 2 * p.heads + h,
                          row0 + j * BKV);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp <= 2) {
    if constexpr (DUAL) {
    // ===================== MMA issuers: warp 1 drives query tile 0, warp 2 drives query tile 1 =====================
    // Two independent in-order streams with hardware-suspended waits (a polling single issuer steals issue slots from
    // the softmax warps of its scheduler).  Per tile t:  S_t(0) | [S_t(j+1)] PV_t(j) ...   where
    //   S_t(j+1) needs K(j+1) and the S_t buffer — SEP_P: softmax pulled S_t(j) into registers (S_FREE);
    //            aliased: it is issued right after PV_t(j) (tensor-core ops of one thread execute in issue order);
    //   PV_t(j)  needs V(j), P_t(j) (P_FULL) and, for j == 0, the previous item's epilogue to have drained O_t (O_FREE).
    // K / V / Q ring slots are released by BOTH issuers (barrier count 2).
    {
      const int t = warp - 1;
      uint32_t g = 0, it = 0, n_p = 0, n_f = 0, n_o = 0;
      DFOT_TRACE_DECL;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        int r, h, q0, nt;
        item_coord(item, r, h, q0, nt);
        const bool active = t < nt && (q0 + t) * BQ < p.Ntok;   // (tile 1 of the last pair may lie outside the sample)
        mbar_wait(bar(Q_FULL), it & 1u);
        if (!active) {                                 // keep the shared rings moving
          for (int j = 0; j < n_kv; ++j) {
            const uint32_t gt = g + j, st = gt & 1u, ph = (gt >> 1) & 1u;
            mbar_wait(bar(K_FULL + st), ph);
            if (lane == 0) mbar_arrive(bar(K_EMPTY + st));
            mbar_wait(bar(V_FULL + st), ph);
            if (lane == 0) mbar_arrive(bar(V_EMPTY + st));
          }
          if (lane == 0) mbar_arrive(bar(Q_EMPTY));
          g += n_kv;
          continue;
        }
        // S_t = Q_t · K^T (K stage st).  The item's last S releases the Q tiles: the next item's Q (and its first K / V
        // tiles behind it in the producer's queue) then load under the last KV step instead of after it.
        auto issue_s = [&](uint32_t st, bool last_s) {
          if (elect_one_sync()) {
            const uint32_t d = tmem_base + TMEM_S + (uint32_t)t * S_STRIDE;
#pragma unroll
            for (int s = 0; s < KS_QK; ++s) {
              const uint32_t off = (uint32_t)(s >> 2) * kAtomBytes + (uint32_t)(s & 3) * 32u;
              umma_bf16(d, desc_kmajor(sQ + t * TILE_BYTES + off), desc_kmajor(sK + st * TILE_BYTES + off), IDESC_S,
                        s > 0 ? 1u : 0u);
            }
            umma_commit(bar(S_FULL + t));
            umma_commit(bar(K_EMPTY + st));
            if (last_s) umma_commit(bar(Q_EMPTY));
          }
          __syncwarp();
        };
        // O_t (+)= P_t · V (V stage st), A operand from TMEM; then the step's commits
        auto issue_pv = [&](uint32_t st, bool first, bool more) {
          if (elect_one_sync()) {
            const uint32_t d = tmem_base + TMEM_O + (uint32_t)t * O_STRIDE;
            const uint32_t a = tmem_base + TMEM_P + (uint32_t)t * P_STRIDE;
#pragma unroll
            for (int s = 0; s < KS_PV; ++s)
              umma_bf16_ts(d, a + (uint32_t)(8 * s), desc_mnmajor(sV + st * TILE_BYTES + (uint32_t)s * 2048u, kAtomBytes),
                           IDESC_PV, (first && s == 0) ? 0u : 1u);
            umma_commit(bar(V_EMPTY + st));
            if (!more) umma_commit(bar(O_DONE + t));
            else if constexpr (SEP_P) umma_commit(bar(PV_DONE + t));
          }
          __syncwarp();
        };
        {
          const uint32_t st = g & 1u, ph = (g >> 1) & 1u;
          mbar_wait(bar(K_FULL + st), ph);
          tc_fence_after();
          issue_s(st, n_kv == 1);
        }
        if (n_o > 0) mbar_wait(bar(O_FREE + t), (n_o - 1) & 1u);
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t gt = g + j, st = gt & 1u, ph = (gt >> 1) & 1u;
          const uint32_t gn = gt + 1, stn = gn & 1u, phn = (gn >> 1) & 1u;
          const bool more = j + 1 < n_kv;
          if (SEP_P && more) {
            mbar_wait(bar(K_FULL + stn), phn);
            DFOT_TRACE(10);
            mbar_wait_fast(bar(S_FREE + t), n_f++ & 1u);
            tc_fence_after();
            DFOT_TRACE(11);
            issue_s(stn, j + 2 == n_kv);
            DFOT_TRACE(12);
          }
          mbar_wait(bar(V_FULL + st), ph);
          DFOT_TRACE(13);
          mbar_wait_fast(bar(P_FULL + t), n_p++ & 1u);
          tc_fence_after();
          DFOT_TRACE(14);
          issue_pv(st, j == 0, more);
          DFOT_TRACE(15);
          if (!SEP_P && more) {
            mbar_wait(bar(K_FULL + stn), phn);
            tc_fence_after();
            issue_s(stn, j + 2 == n_kv);
          }
        }
        ++n_o;
        g += n_kv;
      }
    }
    } else if (warp == 1) {
    // ===================== MMA issuer of the aliased layout (head_dim > 64): one warp drives both query tiles ==========
    // Issue order  S0(0) S1(0) | PV0(j) S0(j+1) PV1(j) S1(j+1) ...: tensor-core ops of one thread execute in issue order,
    // which is what makes the S/P aliasing safe without extra barriers.
    {
      uint32_t g = 0, it = 0;
      uint32_t n_p[2] = {0, 0};                      // P_FULL phases consumed per tile
      uint32_t n_o[2] = {0, 0};                      // items in which tile t was active (O_FREE phases)
      auto one = [&](auto&& f) {                     // f() by one elected lane, warp-uniform control flow around it
        if (elect_one_sync()) f();
        __syncwarp();
      };
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        int r, h, q0, nt;
        item_coord(item, r, h, q0, nt);
        const bool has1 = nt == 2 && (q0 + 1) * BQ < p.Ntok;     // second query tile holds rows of this sample
        auto issue_s = [&](int t, uint32_t st) {          // S_t = Q_t · K^T (K stage st)
          const uint32_t d = tmem_base + TMEM_S + (uint32_t)t * S_STRIDE;
#pragma unroll
          for (int s = 0; s < KS_QK; ++s) {
            const uint32_t off = (uint32_t)(s >> 2) * kAtomBytes + (uint32_t)(s & 3) * 32u;
            umma_bf16(d, desc_kmajor(sQ + t * TILE_BYTES + off), desc_kmajor(sK + st * TILE_BYTES + off), IDESC_S,
                      s > 0 ? 1u : 0u);
          }
          umma_commit(bar(S_FULL + t));
        };
        auto issue_pv = [&](int t, uint32_t st, bool first) {   // O_t (+)= P_t · V (V stage st), A from TMEM
          const uint32_t d = tmem_base + TMEM_O + (uint32_t)t * O_STRIDE;
          const uint32_t a = tmem_base + TMEM_P + (uint32_t)t * P_STRIDE;
#pragma unroll
          for (int s = 0; s < KS_PV; ++s)
            umma_bf16_ts(d, a + (uint32_t)(8 * s), desc_mnmajor(sV + st * TILE_BYTES + (uint32_t)s * 2048u, kAtomBytes),
                         IDESC_PV, (first && s == 0) ? 0u : 1u);
        };
        mbar_wait(bar(Q_FULL), it & 1u);
        {   // prologue: S_t(0)
          const uint32_t st = g & 1u, ph = (g >> 1) & 1u;
          mbar_wait(bar(K_FULL + st), ph);
          tc_fence_after();
          one([&] {
            issue_s(0, st);
            if (has1) issue_s(1, st);
            umma_commit(bar(K_EMPTY + st));
            if (n_kv == 1) umma_commit(bar(Q_EMPTY));   // the item's last S releases the Q tiles (see the dual path)
          });
        }
        // the epilogue of the previous item must have drained O_t before the first (overwriting) PV
        if (n_o[0] > 0) mbar_wait(bar(O_FREE + 0), (n_o[0] - 1) & 1u);
        if (has1 && n_o[1] > 0) mbar_wait(bar(O_FREE + 1), (n_o[1] - 1) & 1u);
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t gt = g + j, st = gt & 1u, ph = (gt >> 1) & 1u;
          const uint32_t gn = gt + 1, stn = gn & 1u, phn = (gn >> 1) & 1u;
          const bool more = j + 1 < n_kv;
          mbar_wait(bar(V_FULL + st), ph);
          if (more) mbar_wait(bar(K_FULL + stn), phn);         // (loaded long ago: K runs a whole step ahead)
          mbar_wait_fast(bar(P_FULL + 0), n_p[0]++ & 1u);      // P_0(j) is in TMEM (O_0 rescaled if it had to be)
          tc_fence_after();
          one([&] {
            issue_pv(0, st, j == 0);
            if (more) issue_s(0, stn); else umma_commit(bar(O_DONE + 0));
            if (!has1) {
              umma_commit(bar(V_EMPTY + st));
              if (more) umma_commit(bar(K_EMPTY + stn));
              if (j + 2 == n_kv) umma_commit(bar(Q_EMPTY));
            }
          });
          if (has1) {
            mbar_wait_fast(bar(P_FULL + 1), n_p[1]++ & 1u);
            tc_fence_after();
            one([&] {
              issue_pv(1, st, j == 0);
              if (more) issue_s(1, stn); else umma_commit(bar(O_DONE + 1));
              umma_commit(bar(V_EMPTY + st));
              if (more) umma_commit(bar(K_EMPTY + stn));
              if (j + 2 == n_kv) umma_commit(bar(Q_EMPTY));
            });
          }
        }
        ++n_o[0];
        if (has1) ++n_o[1];
        g += n_kv;
      }
    }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ===================== softmax warpgroups (thread = one query row of tile t) =====================
    const int t = (warp - 4) >> 2;                   // which query tile of the pair
    const int q = warp & 3;                          // TMEM lane quarter (hardware rule: warp_id % 4)
    const int row = q * 32 + lane;
    const uint32_t t_s = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_S + (uint32_t)t * S_STRIDE;
    const uint32_t t_p = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_P + (uint32_t)t * P_STRIDE;
    const uint32_t t_o = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_O + (uint32_t)t * O_STRIDE;
    uint32_t n_s = 0, n_items = 0, n_pv = 0, it_all = 0;
    DFOT_TRACE_DECL;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int r, h, q0, nt;
      item_coord(item, r, h, q0, nt);
      const int qt = q0 + t;
      const uint32_t it_cur = it_all++;              // tile 0 is active in every item: one STAGGER phase per item
      if (t >= nt || qt * BQ >= p.Ntok) continue;    // single-tile item, or tile 1 of the last pair outside the sample
      if (kStagger != 0 && t == 1) mbar_wait_fast(bar(STAGGER), it_cur & 1u);
      float m_used = 0.f, l_run = 0.f;
      uint32_t v[NCH][32];                           // the KV scores of this row stay in registers
      // pull S_t(jj) into registers (all four 32-column loads in flight), hand the buffer back, mask, row maximum
      auto load_begin = [&]() {
        mbar_wait_fast(bar(S_FULL + t), n_s++ & 1u);
        tc_fence_after();
        DFOT_TRACE(1);
      };
      auto load_chunk = [&](int c) {                 // (the last chunk of a 112-key tile holds 16 scores)
        if (32 * c + 32 <= BKV) tmem_ld_x32(t_s + 32 * c, v[c]);
        else tmem_ld_x16(t_s + 32 * c, *reinterpret_cast<uint32_t(*)[16]>(&v[c][0]));
      };
      auto store_chunk = [&](int c, const uint32_t (&pk)[16]) {   // 32 (16) scores = 16 (8) packed bf16 pairs
        if (32 * c + 32 <= BKV) tmem_st_x16(t_p + 16 * c, pk);
        else tmem_st_x8(t_p + 16 * c, *reinterpret_cast<const uint32_t(*)[8]>(&pk[0]));
      };
      auto load_end = [&](int jj) -> float {
        tmem_ld_wait();
        DFOT_TRACE(2);
        if constexpr (SEP_P) {                       // hand the S_t buffer back: S_t(jj+1) may be computed now
          if (jj + 1 < n_kv) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(S_FREE + t));
          }
        }
        const int valid = p.Ntok - jj * BKV;         // keys of this tile that exist
        if (valid < BKV) {
#pragma unroll
          for (int c = 0; c < BKV; ++c)
            if (c >= valid) v[c >> 5][c & 31] = 0xff800000u;   // -inf
        }
        if constexpr (NOMAX) return 0.f;
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          mx0 = fmaxf(mx0, __uint_as_float(v[0][c]));
          mx1 = fmaxf(mx1, __uint_as_float(v[1][c]));
          mx2 = fmaxf(mx2, __uint_as_float(v[2][c]));
          if (96 + c < BKV) mx3 = fmaxf(mx3, __uint_as_float(v[3][c]));
        }
        return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      };
      // (Measured and dropped, r02: prefetching S_t(j+1) chunk by chunk into the registers the exp2 loop has finished with —
      // 559-623 against 762-786 TFLOP/s at d = 64: a tcgen05.ld in flight slows the exp2 loop of the same warp down.)
      for (int j = 0; j < n_kv; ++j) {
        load_begin();
#pragma unroll
        for (int c = 0; c < NCH; ++c) load_chunk(c);
        const float mx = load_end(j);
        if constexpr (NOMAX) {
          // Bounded scores: p = 2^s directly (reference maximum 0 for every row and tile) — no row maximum, no
          // subtraction, no O rescaling; everything else (P in TMEM, PV, row sums) is unchanged.
          uint64_t sum_a = 0ull, sum_b = 0ull;
          // a quarter of the score pairs take the FMA-pipe polynomial exp2, the rest the MUFU (measured: 824 -> ~900
          // TFLOP/s at d = 64, N = 8192).  A partial last tile carries -inf masks the polynomial cannot take: MUFU only.
          auto exp_tile = [&](auto poly_tag) {
            constexpr bool POLY = decltype(poly_tag)::value;
            uint32_t pka[NCH][16];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              uint32_t (&pk)[16] = pka[c];
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                if (32 * c + 2 * e >= BKV) continue;   // (the last chunk of a 112-key tile holds 16 scores)
                float p0, p1, p2, p3;
                if (POLY && ((kPolyFirstB >> (e >> 1)) & 1))
                  ex2_poly_x2_bounded(__uint_as_float(v[c][2 * e]), __uint_as_float(v[c][2 * e + 1]), p0, p1);
                else { p0 = ex2(__uint_as_float(v[c][2 * e])); p1 = ex2(__uint_as_float(v[c][2 * e + 1])); }
                if (POLY && ((kPolySecondB >> (e >> 1)) & 1))
                  ex2_poly_x2_bounded(__uint_as_float(v[c][2 * e + 2]), __uint_as_float(v[c][2 * e + 3]), p2, p3);
                else { p2 = ex2(__uint_as_float(v[c][2 * e + 2])); p3 = ex2(__uint_as_float(v[c][2 * e + 3])); }
                sum_a = add_f32x2(sum_a, pack_f32x2(p0, p1));
                sum_b = add_f32x2(sum_b, pack_f32x2(p2, p3));
                pk[e] = pack_bf16x2(p0, p1);
                pk[e + 1] = pack_bf16x2(p2, p3);
              }
              // PV_t(j-1) still reads P_t(j-1) from these very columns: it must have retired before the first store of
              // P_t(j).  The wait sits after chunk kPvWaitChunk's exps (the packed P of the earlier chunks stays in
              // registers until then), so that the MMA's latency hides under exp2 work instead of stalling the warp.
              if constexpr (SEP_P) {
                if (c == kPvWaitChunk) {
                  if (j > 0) {
                    DFOT_TRACE(3);
                    mbar_wait_fast(bar(PV_DONE + t), n_pv++ & 1u);
                    tc_fence_after();
                    DFOT_TRACE(4);
                  }
#pragma unroll
                  for (int cc = 0; cc < kPvWaitChunk; ++cc) store_chunk(cc, pka[cc]);
                }
                if (c >= kPvWaitChunk) store_chunk(c, pk);
              } else {
                store_chunk(c, pk);
              }
              if (kStagger == 2 && t == 0 && j == 0 && c == 1 && lane == 0) mbar_arrive(bar(STAGGER));
            }
          };
          if (p.Ntok - j * BKV >= BKV) exp_tile(std::true_type{});
          else exp_tile(std::false_type{});
          float s0, s1, s2, s3;
          unpack_f32x2(sum_a, s0, s1);
          unpack_f32x2(sum_b, s2, s3);
          l_run += (s0 + s1) + (s2 + s3);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(P_FULL + t));
            if (kStagger == 1 && t == 0 && j == 0) mbar_arrive(bar(STAGGER));
          }
          DFOT_TRACE(5);
          continue;
        }
        // Lazy rescaling decision now (the exps below already use the new maximum); the O_t tile itself is rescaled
        // after the exp2 phase, when PV_t(j-1) has long retired.
        float alpha = 1.f;
        bool rescale = false;
        if (j == 0) {
          m_used = mx;                               // the first PV overwrites O: nothing to rescale
        } else {
          const bool raise = mx > m_used + kRescaleThreshold;
          rescale = __any_sync(0xffffffffu, raise);  // warp-uniform: tcgen05.ld/st are warp-collective
          if (raise) {
            alpha = ex2(m_used - mx);
            l_run *= alpha;
            m_used = mx;
          }
        }
        // p = exp2(s - m) → packed bf16 pairs into P_t (the PV MMA's A operand); fp32x2 packed sub / row sum
        const uint64_t neg_m2 = pack_f32x2(-m_used, -m_used);
        uint64_t sum_a = 0ull, sum_b = 0ull;         // (+0.f, +0.f)
        uint32_t pka[NCH][16];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t (&pk)[16] = pka[c];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            if (32 * c + 2 * e >= BKV) continue;     // (the last chunk of a 112-key tile holds 16 scores)
            float x0, x1, x2, x3;
#ifndef DFOT_ATTN_NO_PACKED
            unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[c][2 * e]), __uint_as_float(v[c][2 * e + 1])), neg_m2), x0, x1);
            unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[c][2 * e + 2]), __uint_as_float(v[c][2 * e + 3])), neg_m2), x2, x3);
#else
            x0 = __uint_as_float(v[c][2 * e]) - m_used; x1 = __uint_as_float(v[c][2 * e + 1]) - m_used;
            x2 = __uint_as_float(v[c][2 * e + 2]) - m_used; x3 = __uint_as_float(v[c][2 * e + 3]) - m_used;
#endif
            // pairs selected by kPolyFirst / kPolySecond take the FMA-pipe exp2, the others the MUFU
            float p0, p1, p2, p3;
            if ((kPolyFirst >> (e >> 1)) & 1) ex2_poly_x2(x0, x1, p0, p1); else { p0 = ex2(x0); p1 = ex2(x1); }
            if ((kPolySecond >> (e >> 1)) & 1) ex2_poly_x2(x2, x3, p2, p3); else { p2 = ex2(x2); p3 = ex2(x3); }
            sum_a = add_f32x2(sum_a, pack_f32x2(p0, p1));
            sum_b = add_f32x2(sum_b, pack_f32x2(p2, p3));
            pk[e] = pack_bf16x2(p0, p1);
            pk[e + 1] = pack_bf16x2(p2, p3);
          }
          if constexpr (SEP_P) {
            // PV_t(j-1) reads P_t(j-1) from these very columns (and writes O_t): it must have retired before the first
            // store of P_t(j) — see the bounded-score path above for the placement of the wait.
            if (c == kPvWaitChunk) {
              if (j > 0) {
                DFOT_TRACE(3);
                mbar_wait_fast(bar(PV_DONE + t), n_pv++ & 1u);
                tc_fence_after();
                DFOT_TRACE(4);
              }
#pragma unroll
              for (int cc = 0; cc < kPvWaitChunk; ++cc) store_chunk(cc, pka[cc]);
            }
            if (c >= kPvWaitChunk) store_chunk(c, pk);
          } else {
            store_chunk(c, pk);
          }
          if (kStagger == 2 && t == 0 && j == 0 && c == 1 && lane == 0) mbar_arrive(bar(STAGGER));
        }
        float sum0, sum1, sum2, sum3;
        unpack_f32x2(sum_a, sum0, sum1);
        unpack_f32x2(sum_b, sum2, sum3);
        if (j > 0) {
          if (rescale) {
            // aliased layout: S_FULL(j) was committed after PV(j-1); separate layout: PV_DONE(j-1) was awaited above.
            // Either way O_t is complete and idle until this warpgroup arrives on P_FULL
#pragma unroll
            for (int c0 = 0; c0 < DP; c0 += 32) {
              if constexpr (DP % 32 != 0) {
                if (c0 + 32 > DP) {
                  uint32_t o[16];
                  tmem_ld_x16(t_o + c0, o);
                  tmem_ld_wait();
#pragma unroll
                  for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
                  tmem_st_x16(t_o + c0, o);
                  continue;
                }
              }
              uint32_t o[32];
              tmem_ld_x32(t_o + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
              tmem_st_x32(t_o + c0, o);
            }
          }
        }
        l_run += (sum0 + sum1) + (sum2 + sum3);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(P_FULL + t));
          if (kStagger == 1 && t == 0 && j == 0) mbar_arrive(bar(STAGGER));
        }
      }
      // ---- epilogue: O_t / l → bf16 rows
      mbar_wait_fast(bar(O_DONE + t), n_items++ & 1u);
      tc_fence_after();
      const int qrow = qt * BQ + row;
      const float inv = 1.f / l_run;
      __nv_bfloat16* dst = p.out + ((int64_t)r * p.Ntok + qrow) * p.ld_out + (int64_t)h * DH;
#pragma unroll
      for (int c0 = 0; c0 < DP; c0 += 32) {
        uint32_t o[32];
        if (DP % 32 != 0 && c0 + 32 > DP) {
          uint32_t o16[16];
          tmem_ld_x16(t_o + c0, o16);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = o16[c];
        } else {
          tmem_ld_x32(t_o + c0, o);
          tmem_ld_wait();
        }
        if (qrow < p.Ntok) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            if (c0 + c < DH) {   // DH % 8 == 0
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv);
              w.y = pack_bf16x2(__uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
              w.z = pack_bf16x2(__uint_as_float(o[c + 4]) * inv, __uint_as_float(o[c + 5]) * inv);
              w.w = pack_bf16x2(__uint_as_float(o[c + 6]) * inv, __uint_as_float(o[c + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + c0 + c) = w;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(O_FREE + t));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

static int attention_impl_override() {   // DFOT_ATTENTION_IMPL=1|2 pins the kernel (benchmarking); default: by shape
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFOT_ATTENTION_IMPL");
    v = (e != nullptr && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }
  return v;
}
static bool attention_split_tail() {     // DFOT_ATTENTION_SPLIT_TAIL=0 keeps whole pairs in the last wave (benchmarking)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFOT_ATTENTION_SPLIT_TAIL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int DH, int DP>
static int launch(const void* qkv, void* out, int64_t ld_out, int64_t R, int64_t Ntok, int64_t heads, float score_bound,
                  cudaStream_t s) {
  constexpr int ATOMS = (DP + 63) / 64;
  constexpr int smem1 = 5 * ATOMS * kAtomBytes + 2 * 2 * kAtomBytes + 160 /*barriers*/ + 2048 /*exchange*/;
  constexpr int smem2 = 6 * ATOMS * kAtomBytes + 512 /*barriers*/;
  static_assert(smem1 <= 232448 && smem2 <= 232448, "attention: shared memory budget exceeded");
  const int ov = attention_impl_override();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // Kernel 2 (two query tiles per CTA) whenever a sample has more than one query tile — except, in latency mode (the choice
  // depends on the batch, see dfot_set_latency_mode), in the latency regime of small-batch DiT sampling: short sequences whose single query tiles all fit one wave run as twice as many CTAs of half
  // the serial work each (measured in the DMLab bench, profiles/r02_dmlab_attn_impl_ab.txt: T = 36 batch 1 735 -> 771
  // frames/s, T = 16 batch 4 989 -> 1032, batch 1 332 -> 338).
  const bool latency_regime = latency_mode() && score_bound == 0.f && Ntok <= 1024 && R * heads * ceil_div(Ntok, BQ) <= sms;
  const bool paired = ov == 2 || (ov == 0 && Ntok > BQ && !latency_regime);
  const int smem_bytes = paired ? smem2 : smem1;
  constexpr int KV2 = KvTile<DP>::value;                      // keys per KV tile of kernel 2
  EncodeTiledFn enc = get_encode_fn();
  DFOT_REQUIRE(enc != nullptr, DFOT_ERR_DRIVER, "attention: cuTensorMapEncodeTiled unavailable from the driver");
  // qkv viewed as [tokens][3*heads][DH]: box = 64 (d) x 1 x rows (tokens); d beyond DH is zero-filled by TMA.  Query
  // tiles are 128 rows; the K / V tiles of kernel 2 may be narrower (KvTile), hence a second map.
  CUtensorMap tmap, tmap_kv;
  cuuint64_t gdim[3] = {(cuuint64_t)DH, (cuuint64_t)(3 * heads), (cuuint64_t)(R * Ntok)};
  cuuint64_t gstr[2] = {(cuuint64_t)DH * 2, (cuuint64_t)(3 * heads * DH) * 2};
  cuuint32_t estr[3] = {1, 1, 1};
  for (int m = 0; m < 2; ++m) {
    cuuint32_t box[3] = {64, 1, (cuuint32_t)(m == 0 ? 128 : KV2)};
    CUresult cr = enc(m == 0 ? &tmap : &tmap_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr,
                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DFOT_REQUIRE(cr == CUDA_SUCCESS, DFOT_ERR_DRIVER, "attention: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
  }
  const bool nomax = paired && score_bound > 0.f && score_bound <= 96.f;   // kernel 2 only
  void (*kern1)(const CUtensorMap, const Params) = attention_tcgen05_kernel<DH, DP>;
  void (*kern2)(const CUtensorMap, const CUtensorMap, const Params) =
      nomax ? attention2_tcgen05_kernel<DH, DP, true, KV2> : attention2_tcgen05_kernel<DH, DP, false, KV2>;
  const int which = paired ? (nomax ? 2 : 1) : 0;
  static bool configured[3] = {false, false, false};
  if (!configured[which]) {
    cudaError_t e = paired ? cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
                           : cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    DFOT_REQUIRE(e == cudaSuccess, DFOT_ERR_CUDA, "attention: cannot reserve %d B shared memory: %s", smem_bytes,
                 cudaGetErrorString(e));
    configured[which] = true;
  }
  Params p;
  p.out = (__nv_bfloat16*)out;
  p.ld_out = ld_out;
  p.no_max = nomax ? 1 : 0;
  p.R = (int)R; p.Ntok = (int)Ntok; p.heads = (int)heads;
  p.q_tiles = (int)ceil_div(Ntok, BQ);
  p.kv_tiles = (int)ceil_div(Ntok, paired ? KV2 : BKV);
  p.num_items = (int)(R * heads * (paired ? (p.q_tiles + 1) / 2 : p.q_tiles));
  p.pair_items = p.num_items;
  if (paired && attention_split_tail() && p.q_tiles % 2 == 0 && p.num_items > sms) {
    // Persistent CTAs take items round-robin, so the last wave holds num_items % sms pairs on as many SMs while the
    // others idle for the length of a whole item.  Those pairs are handed out as single query tiles instead (twice the
    // items, half the work each) when they still fit one wave: the tail then costs about half a wave.  (With more than
    // sms / 2 pairs left the singles would need two rounds, and a single costs more than half a pair — K / V tiles are
    // shared inside a pair — so those tails stay whole: measured 866 vs 1000 TFLOP/s at d = 128, R = 8, N = 2048.)
    const int rem = p.num_items % sms;
    if (rem > 0 && 2 * rem <= sms) {
      p.pair_items = p.num_items - rem;
      p.num_items = p.pair_items + 2 * rem;
    }
  }
  const int grid = p.num_items < sms ? p.num_items : sms;
  if (paired) launch_pdl(kern2, dim3(grid), dim3(kThreads2), smem_bytes, s, tmap, tmap_kv, p);
  else launch_pdl(kern1, dim3(grid), dim3(kThreads), smem_bytes, s, tmap, p);
  DFOT_CHECK_LAUNCH("attention_tcgen05");
  return DFOT_OK;
}

}  // namespace fattn
}  // namespace dfot

#ifdef DFOT_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int dfot_debug_attn_trace(void* dst, int64_t bytes) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(dst, dfot::fattn::g_trace, (size_t)bytes);
}
#endif

extern "C" int dfot_attention(const void* qkv, void* out, int64_t R, int64_t Ntok, int64_t heads, int64_t head_dim,
                              void* stream) {
  return dfot_attention_strided(qkv, out, heads * head_dim, R, Ntok, heads, head_dim, stream);
}

extern "C" int dfot_attention_strided(const void* qkv, void* out, int64_t ld_out, int64_t R, int64_t Ntok, int64_t heads,
                                      int64_t head_dim, void* stream) {
  return dfot_attention_bounded(qkv, out, ld_out, R, Ntok, heads, head_dim, 0.f, stream);
}

extern "C" int dfot_attention_bounded(const void* qkv, void* out, int64_t ld_out, int64_t R, int64_t Ntok, int64_t heads,
                                      int64_t head_dim, float score_bound, void* stream) {
  using namespace dfot;
  DFOT_REQUIRE(score_bound >= 0.f, DFOT_ERR_INVALID_ARG, "attention: score_bound must be >= 0 (0 = unknown)");
  DFOT_REQUIRE(qkv && out && R > 0 && Ntok > 0 && heads > 0, DFOT_ERR_INVALID_ARG, "attention: bad arguments");
  DFOT_REQUIRE(ld_out >= heads * head_dim && ld_out % 8 == 0, DFOT_ERR_INVALID_ARG,
               "attention: ld_out must be a multiple of 8 and >= heads*head_dim");
  DFOT_REQUIRE(R * Ntok < (1ll << 31) && R * heads * ceil_div(Ntok, 128) < (1ll << 31), DFOT_ERR_UNSUPPORTED,
               "attention: problem too large");
  DFOT_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), DFOT_ERR_UNSUPPORTED,
               "attention: qkv and out must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (head_dim) {
    case 64: return fattn::launch<64, 64>(qkv, out, ld_out, R, Ntok, heads, score_bound, s);
    case 72: return fattn::launch<72, 80>(qkv, out, ld_out, R, Ntok, heads, score_bound, s);
    case 128: return fattn::launch<128, 128>(qkv, out, ld_out, R, Ntok, heads, score_bound, s);
  }
  set_error("attention: head_dim %lld unsupported (64, 72, 128)", (long long)head_dim);
  return DFOT_ERR_UNSUPPORTED;
}
