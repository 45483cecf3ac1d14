// K2 — bf16 GEMM on the 5th-gen tensor cores: C[M,N] = epi(A[M,K] · W[N,K]^T + bias).
//
// Persistent, warp-specialised kernel (one CTA per SM):
//   warp 0   : TMA producer  — cp.async.bulk.tensor 2D loads of A (128x64) and W (BNx64) tiles into a
//              128B-swizzled shared-memory ring, completion on mbarriers
//   warp 1   : MMA issuer    — one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16),
//              accumulating in tensor memory; tcgen05.commit releases ring slots / publishes accumulators
//   warps 2-9: epilogue      — tcgen05.ld (32 lanes x 32 columns per warp) → XOR-swizzled shared-memory
//              transpose → row-wise pass where a warp touches one contiguous row segment per instruction, so
//              bias / gate / residual / RoPE-table reads and the output stores are all fully coalesced.
//              Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// Both operands are K-major (A row-major activations, W = nn.Linear.weight), so no transposes exist anywhere.
#include <cuda.h>

#include "common.cuh"

namespace dfot {
namespace gemm {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // one 128-byte swizzle atom of bf16 along K
constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;  // TMA warp + MMA warp + epilogue warps
constexpr int kRasterBand = 16;  // m-blocks per rasterisation band (L2 reuse of A and W among concurrent CTAs)

template <int BN> struct Cfg {
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = BN * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int kTmemCols = BN == 192 ? 512 : 2 * BN;  // two accumulator stages at columns 0 and BN (power of two >= 32)
  static constexpr int kEpiStageBytes = 32 * 32 * 4;  // per epilogue warp: 32 rows x 32 fp32, swizzled
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct Params {
  int M, N, K;
  void* C;
  int64_t ldc;
  dfot_gemm_epilogue e;
  // implicit-GEMM 3x3 convolution over NHWC activations (conv_cblks == 0: plain GEMM).  K = 9 taps x Cin; the A tile
  // of tap (dy, dx) is the pixel tile shifted by (dy, dx), fetched by a 4-D TMA whose out-of-bounds zero fill IS
  // the convolution's zero padding.
  int conv_cblks;   // ceil(Cin / 64)
  int conv_W, conv_H;
  int conv_taps;    // 9 (3x3) or 9*kt (causal kt x 3 x 3 over a frame axis: tap / 9 = frame offset of the TMA box)
  // split-K (single-CTA kernel, plain GEMM, F32 epilogue without bias): N = split_k * (weight rows) output columns;
  // n-block nb covers weight rows (nb % num_n_w) * BN and the k-blocks of split nb / num_n_w — the partial sums of split s
  // land in columns [s * N_w, (s + 1) * N_w) of C (N_w % BN == 0, host-checked), so the epilogue is unchanged.
  int split_k;      // 1 = off
};

// ------------------------------------------------------------------ PTX wrappers
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
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(2000u)   // suspend-time hint (ns): sleep in hardware, do not burn issue slots
        : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("dfot_gemm: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
               threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the TMA of either CTA signals the LEADER CTA's mbarrier (peer bit cleared) ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_x32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait_all() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;             // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset between 8-row core-matrix groups
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=BN
template <int BN> __device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------ epilogue
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float u = x * (0.7978845608028654f + 0.035677408136300125f * x * x);
  return 0.5f * x * (1.0f + tanh_approx(u));
}
__device__ __forceinline__ float silu_fast(float x) { return 0.5f * x * (1.0f + tanh_approx(0.5f * x)); }

// Phase 1: this warp's 32x32 fp32 accumulator chunk (thread = row) → swizzled smem (16-byte unit j of row r
// lives at unit j ^ (r & 7)): conflict-free for the row-owner writes AND for the row-wise reads below.
__device__ __forceinline__ void stage_chunk(uint32_t stage, int lane, const uint32_t (&r)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t addr = stage + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                 "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                 : "memory");
  }
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// silu(x) = h + h*tanh(h), h = x/2: one MUFU + 2 FP32 ops
__device__ __forceinline__ float silu_fast2(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}

// ---------------------------------------------------------------------------------------------------------------------
// Phase 2 of the epilogue walks the staged 32x32 fp32 chunk ROW-WISE WITH 128-BIT ACCESSES: a lane owns one 16-byte
// unit of the staged row per step (4 fp32 outputs = 16 B, or 8 bf16 outputs = 16 B), so a chunk costs 8 (f32) or
// 4 x 2 (bf16) shared loads and 8 / 4 global stores per lane instead of 32 / 16 scalar ones, and row addresses advance
// by one 64-bit add per step.  The XOR swizzle (unit j of row r at j ^ (r & 7)) keeps every access at the 4-wavefront
// minimum.  (At K = 576 the scalar version made the kernel epilogue-bound: 483 instructions per chunk and warp.)
// Side inputs (residual rows) are fetched BEFORE the accumulator is touched so that 8 independent 16-byte loads per
// lane are in flight during the TMEM read + transpose.  FULL = chunk completely inside the matrix: no predicates.

// GroupNorm side output (F32 / RESID_F32 / BF16 epilogues).  A lane holds per-column (sum, sum of squares) of its NC
// adjacent columns over the rows it visited; LPR lanes share a row, lanes with equal (lane % LPR) share columns.  All
// 32 rows of a chunk belong to one image (rows_per_img % 32 == 0); channels per group is a power of two.
// GPL = groups inside a lane's NC columns (1 when a group is at least NC wide).
template <int NC, int LPR, int GPL>
__device__ __forceinline__ void gn_flush_impl(const Params& p, int lane, int m0, int col0, int cpg, const float (&s)[NC],
                                              const float (&q)[NC], bool col_ok) {
  constexpr int CPL = NC / GPL;                        // columns of one group inside the lane
  const int64_t img = m0 / p.e.gn_rows_per_img;
  double* base = p.e.gn_sums + img * p.e.gn_groups * 2;
#pragma unroll
  for (int g = 0; g < GPL; ++g) {
    float ts = 0.f, tq = 0.f;
    if (col_ok) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) { ts += s[g * CPL + j]; tq += q[g * CPL + j]; }
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {               // fold the rows (lanes with equal lane % LPR share columns)
      ts += __shfl_xor_sync(0xffffffffu, ts, o);
      tq += __shfl_xor_sync(0xffffffffu, tq, o);
    }
    int lpg = 1;                                       // lanes of the row that share the group (GPL == 1 only)
    if (GPL == 1) {
      lpg = cpg / NC;
      if (lpg > LPR) lpg = LPR;
      for (int o = 1; o < lpg; o <<= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
        tq += __shfl_xor_sync(0xffffffffu, tq, o);
      }
    }
    if (lane < LPR && (lane % lpg) == 0 && col_ok) {
      // 64-bit fixed point (2^-32 units, the format of uvit.cu's statistics kernel): integer atomics are associative, so
      // the order in which the chunks of an image arrive cannot change the statistics; `ts` / `tq` themselves are fp32
      // sums over a fixed pattern (32 rows x the group's columns of one chunk) that no tile shape or batch size alters
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(base) + ((col0 + g * CPL) / cpg) * 2;
      atomicAdd(dst, (unsigned long long)__float2ll_rn(ts * 4294967296.f));
      atomicAdd(dst + 1, (unsigned long long)__float2ll_rn(tq * 4294967296.f));
    }
  }
}
template <int NC, int LPR>
__device__ __forceinline__ void gn_flush(const Params& p, int lane, int m0, int col0, const float (&s)[NC],
                                         const float (&q)[NC], bool col_ok) {
  const int cpg = p.N / (int)p.e.gn_groups;            // power of two (host-checked), warp-uniform
  if (cpg >= NC) gn_flush_impl<NC, LPR, 1>(p, lane, m0, col0, cpg, s, q, col_ok);
  else if (2 * cpg == NC) gn_flush_impl<NC, LPR, 2>(p, lane, m0, col0, cpg, s, q, col_ok);
  else if (4 * cpg == NC) gn_flush_impl<NC, LPR, 4>(p, lane, m0, col0, cpg, s, q, col_ok);
  else gn_flush_impl<NC, LPR, NC>(p, lane, m0, col0, cpg, s, q, col_ok);       // one channel per group
}

// epilogue families: a residual operand (fetched ahead of the accumulator), a per-frame gate
__host__ __device__ constexpr bool epi_has_resid(int epi) {
  return epi == DFOT_EPI_RESID_F32 || epi == DFOT_EPI_GATE_RESID_F32 || epi == DFOT_EPI_GATE_LNRESID_F32;
}
__host__ __device__ constexpr bool epi_has_gate(int epi) { return epi == DFOT_EPI_GATE_RESID_F32 || epi == DFOT_EPI_GATE_LNRESID_F32; }

template <int EPI> struct ChunkSide {
  float4 r[8];   // *_RESID_F32: residual of (row 4*it + lane/8, columns 4*(lane%8)..+3), it = 0..7
  float v[32];   // QKV_ROPE: (cos, sin) of row 2k + (lane>>4) for the lane's rotation pair
};

template <int EPI, bool FULL>
__device__ __forceinline__ void prefetch_side(const Params& p, ChunkSide<EPI>& sd, int lane, int m0, int n0) {
  if constexpr (epi_has_resid(EPI)) {
    const int col = n0 + ((lane & 7) << 2), rsub = lane >> 3;
    const float* res = p.e.resid + (int64_t)(m0 + rsub) * p.e.ld_resid + col;
    const int64_t step = 4 * p.e.ld_resid;
    if constexpr (FULL) {
#pragma unroll
      for (int it = 0; it < 8; ++it, res += step) sd.r[it] = __ldg(reinterpret_cast<const float4*>(res));
    } else {
      const int rows = p.M - m0;
      const bool col_ok = col < p.N;   // N % 4 == 0: a 16-byte unit is entirely inside or outside
#pragma unroll
      for (int it = 0; it < 8; ++it, res += step)
        sd.r[it] = (col_ok && 4 * it + rsub < rows) ? __ldg(reinterpret_cast<const float4*>(res))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if constexpr (EPI == DFOT_EPI_QKV_ROPE_BF16) {
    const int col = n0 + ((lane & 15) << 1);
    const int dh = (int)p.e.head_dim, tps = (int)p.e.tokens_per_sample;
    const bool rotate = col < 2 * (int)p.e.model_dim && col < p.N;
    const float2* cs = reinterpret_cast<const float2*>(p.e.rope_cs) + ((col % dh) >> 1);
    const int tok0 = (m0 + (lane >> 4)) % tps;
    const int hd2 = dh >> 1;
    if (!rotate) {   // v columns: identity rotation
#pragma unroll
      for (int k = 0; k < 16; ++k) { sd.v[2 * k] = 1.f; sd.v[2 * k + 1] = 0.f; }
    } else if (FULL && tok0 + 31 < tps) {   // no wrap into the next sample inside this chunk
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float2 c = __ldg(cs + (tok0 + 2 * k) * hd2);
        sd.v[2 * k] = c.x;
        sd.v[2 * k + 1] = c.y;
      }
    } else {
      const int rows = p.M - m0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        int tok = tok0 + 2 * k;
        if (tok >= tps) tok %= tps;
        float2 c = make_float2(1.f, 0.f);
        if (2 * k + (lane >> 4) < rows) c = __ldg(cs + tok * hd2);
        sd.v[2 * k] = c.x;
        sd.v[2 * k + 1] = c.y;
      }
    }
  }
}

// fp32 outputs (F32, RESID_F32, GATE_RESID_F32, GATE_LNRESID_F32): lane = (row 4*it + lane/8, columns 4*(lane%8)..+3), 8 steps.
// GATE_LNRESID: the residual base is not the stored tensor r but its modulated LayerNorm, rebuilt here from the row's
// (mean, rstd) and the frame's shift / scale vectors with the expression dfot_adaln_layernorm uses — so that kernel need not
// store the fp32 copy of its output; the output may overwrite r (every element is read and written by the same lane).
template <int EPI, bool FULL>
__device__ __forceinline__ void epilogue_rows_f32(const Params& p, const ChunkSide<EPI>& sd, uint32_t stage, int lane,
                                                  int m0, int n0) {
  const int cq = lane & 7, rsub = lane >> 3;
  const int col = n0 + (cq << 2);
  const bool col_ok = FULL || col < p.N;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.e.bias != nullptr && col_ok) bias = __ldg(reinterpret_cast<const float4*>(p.e.bias + col));
  const int rows = FULL ? 32 : p.M - m0;
  float* out = reinterpret_cast<float*>(p.C) + (int64_t)(m0 + rsub) * p.ldc + col;
  const int64_t step = 4 * p.ldc;
  float4 gate0 = make_float4(0.f, 0.f, 0.f, 0.f), gate1 = gate0;
  float4 sc0 = gate0, sc1 = gate0, sh0 = gate0, sh1 = gate0;   // GATE_LNRESID: scale / shift of the chunk's two frames
  // tokens_per_frame >= 32: rows [0, split) of the chunk belong to frame f0, the rest to f0 + 1 — two gate vectors per
  // lane.  Smaller frames (8x8 latents with patch 2: 16 tokens, the DMLab / Minecraft DiT configurations): a chunk spans
  // several frames and every row looks its (L1-resident) gate vector up.
  int split = 32, P = 1, f0 = 0, rem0 = 0;
  bool small_frames = false;     // warp-uniform
  if constexpr (epi_has_gate(EPI)) {
    P = (int)p.e.tokens_per_frame;
    f0 = m0 / P;
    rem0 = m0 - f0 * P;
    small_frames = P < 32;
    split = min(32, (f0 + 1) * P - m0);
    if (col_ok && !small_frames) {
      gate0 = __ldg(reinterpret_cast<const float4*>(p.e.gate + (int64_t)f0 * p.e.ld_gate + col));
      if (split < 32) gate1 = __ldg(reinterpret_cast<const float4*>(p.e.gate + (int64_t)(f0 + 1) * p.e.ld_gate + col));
      if constexpr (EPI == DFOT_EPI_GATE_LNRESID_F32) {
        sc0 = __ldg(reinterpret_cast<const float4*>(p.e.ln_scale + (int64_t)f0 * p.e.ld_gate + col));
        sh0 = __ldg(reinterpret_cast<const float4*>(p.e.ln_shift + (int64_t)f0 * p.e.ld_gate + col));
        if (split < 32) {
          sc1 = __ldg(reinterpret_cast<const float4*>(p.e.ln_scale + (int64_t)(f0 + 1) * p.e.ld_gate + col));
          sh1 = __ldg(reinterpret_cast<const float4*>(p.e.ln_shift + (int64_t)(f0 + 1) * p.e.ld_gate + col));
        }
      }
    }
  }
  const bool gn = p.e.gn_sums != nullptr;   // warp-uniform
  float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < 8; ++it, out += step) {
    const int i = 4 * it + rsub;
    float4 y = lds_f32x4(stage + (uint32_t)i * 128u + (uint32_t)((cq ^ (i & 7)) << 4));
    y.x += bias.x; y.y += bias.y; y.z += bias.z; y.w += bias.w;
    if constexpr (EPI == DFOT_EPI_GATE_RESID_F32) {
      float4 g = i < split ? gate0 : gate1;
      if (small_frames && col_ok && (FULL || i < rows))
        g = __ldg(reinterpret_cast<const float4*>(p.e.gate + (int64_t)(f0 + (rem0 + i) / P) * p.e.ld_gate + col));
      y.x = fmaf(g.x, y.x, sd.r[it].x); y.y = fmaf(g.y, y.y, sd.r[it].y);
      y.z = fmaf(g.z, y.z, sd.r[it].z); y.w = fmaf(g.w, y.w, sd.r[it].w);
    }
    if constexpr (EPI == DFOT_EPI_GATE_LNRESID_F32) {
      float4 g = i < split ? gate0 : gate1, sc = i < split ? sc0 : sc1, sh = i < split ? sh0 : sh1;
      float2 st = make_float2(0.f, 0.f);                       // (mean, rstd) of row m0 + i
      if (col_ok && (FULL || i < rows)) {
        st = __ldg(reinterpret_cast<const float2*>(p.e.ln_stats) + m0 + i);
        if (small_frames) {
          const int64_t fo = (int64_t)(f0 + (rem0 + i) / P) * p.e.ld_gate + col;
          g = __ldg(reinterpret_cast<const float4*>(p.e.gate + fo));
          sc = __ldg(reinterpret_cast<const float4*>(p.e.ln_scale + fo));
          sh = __ldg(reinterpret_cast<const float4*>(p.e.ln_shift + fo));
        }
      }
      const float4 r = sd.r[it];
      // the residual base, bit for bit what adaln_layernorm_kernel computes: ((x - mean) * rstd) * (1 + scale) + shift
      const float bx = adaln_value(r.x, st.x, st.y, sc.x, sh.x), by = adaln_value(r.y, st.x, st.y, sc.y, sh.y);
      const float bz = adaln_value(r.z, st.x, st.y, sc.z, sh.z), bw = adaln_value(r.w, st.x, st.y, sc.w, sh.w);
      y.x = fmaf(g.x, y.x, bx); y.y = fmaf(g.y, y.y, by); y.z = fmaf(g.z, y.z, bz); y.w = fmaf(g.w, y.w, bw);
    }
    if constexpr (EPI == DFOT_EPI_RESID_F32) {
      y.x += sd.r[it].x; y.y += sd.r[it].y; y.z += sd.r[it].z; y.w += sd.r[it].w;
    }
    if (FULL || (col_ok && i < rows)) {
      *reinterpret_cast<float4*>(out) = y;
      if (gn) {
        gs[0] += y.x; gq[0] = fmaf(y.x, y.x, gq[0]); gs[1] += y.y; gq[1] = fmaf(y.y, y.y, gq[1]);
        gs[2] += y.z; gq[2] = fmaf(y.z, y.z, gq[2]); gs[3] += y.w; gq[3] = fmaf(y.w, y.w, gq[3]);
      }
    }
  }
  if (gn) gn_flush<4, 8>(p, lane, m0, col, gs, gq, col_ok);
}

// bf16 outputs without rotation (BF16, GELU_BF16, SILU_BF16): lane = (row 8*it + lane/4, columns 8*(lane%4)..+7), 4 steps.
template <int EPI, bool FULL>
__device__ __forceinline__ void epilogue_rows_bf16(const Params& p, uint32_t stage, int lane, int m0, int n0) {
  const int cq = lane & 3, rsub = lane >> 2;
  const int col = n0 + (cq << 3);
  const bool col_ok = FULL || col < p.N;    // N % 8 == 0
  float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (p.e.bias != nullptr && col_ok) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.e.bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.e.bias + col + 4));
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
  const int rows = FULL ? 32 : p.M - m0;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)(m0 + rsub) * p.ldc + col;
  const int64_t step = 8 * p.ldc;
  const bool gn = p.e.gn_sums != nullptr;   // warp-uniform
  float gs[8], gq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { gs[j] = 0.f; gq[j] = 0.f; }
#pragma unroll
  for (int it = 0; it < 4; ++it, out += step) {
    const int i = 8 * it + rsub;
    const uint32_t row = stage + (uint32_t)i * 128u;
    const float4 lo = lds_f32x4(row + (uint32_t)(((2 * cq) ^ (i & 7)) << 4));
    const float4 hi = lds_f32x4(row + (uint32_t)(((2 * cq + 1) ^ (i & 7)) << 4));
    float v[8] = {lo.x + b[0], lo.y + b[1], lo.z + b[2], lo.w + b[3], hi.x + b[4], hi.y + b[5], hi.z + b[6], hi.w + b[7]};
    if constexpr (EPI == DFOT_EPI_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = gelu_tanh_fast(v[j]);
    } else if constexpr (EPI == DFOT_EPI_SILU_BF16) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = silu_fast2(v[j]);
    }
    if (FULL || (col_ok && i < rows)) {
      const uint4 w = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                 pack_bf16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(out) = w;
      if (gn) {   // statistics of the values as stored (bf16-rounded), like a separate pass over the output would see
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 r = unpack_bf16x2(ww[j]);
          gs[2 * j] += r.x; gq[2 * j] = fmaf(r.x, r.x, gq[2 * j]);
          gs[2 * j + 1] += r.y; gq[2 * j + 1] = fmaf(r.y, r.y, gq[2 * j + 1]);
        }
      }
    }
  }
  if (gn) gn_flush<8, 4>(p, lane, m0, col, gs, gq, col_ok);
}

// bf16 outputs with RoPE-3D (QKV_ROPE): a lane owns two adjacent columns (a rotation pair); lanes 0-15 take row 2k,
// lanes 16-31 row 2k+1.
template <bool FULL>
__device__ __forceinline__ void epilogue_rows_rope(const Params& p, const ChunkSide<DFOT_EPI_QKV_ROPE_BF16>& sd,
                                                   uint32_t stage, int lane, int m0, int n0) {
  const int cl = (lane & 15) << 1;
  const int col = n0 + cl;
  const bool col_ok = FULL || col < p.N;  // N is even
  float2 bias = make_float2(0.f, 0.f);
  if (p.e.bias != nullptr && col_ok) bias = __ldg(reinterpret_cast<const float2*>(p.e.bias + col));
  const int rows = FULL ? 32 : p.M - m0;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)m0 * p.ldc + col;
  const float qs = col < (int)p.e.model_dim ? p.e.q_scale : 1.f;
  const int half = lane >> 4;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int i = 2 * k + half;
    float2 v = lds_f32x2(stage + (uint32_t)i * 128u + (uint32_t)((((cl >> 2) ^ (i & 7))) << 4) +
                         (uint32_t)((cl & 3) << 2));
    v.x += bias.x;
    v.y += bias.y;
    const float c = sd.v[2 * k], sn = sd.v[2 * k + 1];   // (1, 0) on v columns
    const float x0 = v.x, x1 = v.y;
    v.x = (x0 * c - x1 * sn) * qs;
    v.y = (x1 * c + x0 * sn) * qs;
    if (FULL || (col_ok && i < rows)) *reinterpret_cast<uint32_t*>(out + (int64_t)i * p.ldc) = pack_bf16x2(v.x, v.y);
  }
}

// QKNORM_ROPE: one HEAD (DH = 64 or 128 columns) of the lane-quarter's 32 rows.  Thread = row while the accumulators are in
// registers, so bias, the head's sum of squares and 1/rms are thread-local (no shuffles, fp32 accumulators); the
// normalised values then go through the usual swizzled transpose, and the row-wise pass applies the per-column norm
// weight and the RoPE-3D rotation (a lane's 8 columns = 4 rotation pairs) before the 16-byte bf16 stores.
template <int DH, bool FULL>
__device__ __forceinline__ void epilogue_head_qknorm(const Params& p, uint32_t t_addr, uint32_t stage, int lane, int m0,
                                                     int n0) {
  constexpr int NCH = DH / 32;
  const int which = n0 / (int)p.e.model_dim;            // 0 = q, 1 = k, 2 = v (heads never straddle: model_dim % DH == 0)
  // pass 1: the head's sum of squares (accumulators + bias), 32 columns at a time — TMEM reads are cheap, registers
  // are not (a 128-wide head would need 128 live accumulators)
  float ss = 0.f;
  if (which < 2) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t r[32];
      tmem_ld_x32(t_addr + 32 * c, r);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.e.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.e.bias + n0 + 32 * c + j));   // warp-uniform
        const float v0 = __uint_as_float(r[j]) + b.x, v1 = __uint_as_float(r[j + 1]) + b.y;
        const float v2 = __uint_as_float(r[j + 2]) + b.z, v3 = __uint_as_float(r[j + 3]) + b.w;
        ss = fmaf(v0, v0, ss); ss = fmaf(v1, v1, ss); ss = fmaf(v2, v2, ss); ss = fmaf(v3, v3, ss);
      }
    }
  }
  const float rstd = which < 2 ? rsqrtf(ss / (float)DH + p.e.qk_eps) * (which == 0 ? p.e.q_scale : 1.f) : 1.f;
  const float* wn = which == 0 ? p.e.qn_w : p.e.kn_w;
  const int cq = lane & 3, rsub = lane >> 2;
  const int tps = (int)p.e.tokens_per_sample;
  const int rows = FULL ? 32 : p.M - m0;
  // pass 2: (accumulator + bias) * rstd → transpose → weight, RoPE, bf16
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[32];
    tmem_ld_x32(t_addr + 32 * c, r);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.e.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.e.bias + n0 + 32 * c + j));
      r[j] = __float_as_uint((__uint_as_float(r[j]) + b.x) * rstd);
      r[j + 1] = __float_as_uint((__uint_as_float(r[j + 1]) + b.y) * rstd);
      r[j + 2] = __float_as_uint((__uint_as_float(r[j + 2]) + b.z) * rstd);
      r[j + 3] = __float_as_uint((__uint_as_float(r[j + 3]) + b.w) * rstd);
    }
    stage_chunk(stage, lane, r);
    __syncwarp();
    const int hc = 32 * c + 8 * cq;                     // first of the lane's 8 columns inside the head
    float w[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (which < 2) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wn + hc)), w1 = __ldg(reinterpret_cast<const float4*>(wn + hc + 4));
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
    }
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)(m0 + rsub) * p.ldc + n0 + hc;
    const int64_t step = 8 * p.ldc;
#pragma unroll
    for (int it = 0; it < 4; ++it, out += step) {
      const int i = 8 * it + rsub;
      const uint32_t row = stage + (uint32_t)i * 128u;
      const float4 lo = lds_f32x4(row + (uint32_t)(((2 * cq) ^ (i & 7)) << 4));
      const float4 hi = lds_f32x4(row + (uint32_t)(((2 * cq + 1) ^ (i & 7)) << 4));
      float v[8] = {lo.x * w[0], lo.y * w[1], lo.z * w[2], lo.w * w[3], hi.x * w[4], hi.y * w[5], hi.z * w[6], hi.w * w[7]};
      if (which < 2 && (FULL || i < rows)) {
        const int tok = (m0 + i) % tps;
        const float4* cs = reinterpret_cast<const float4*>(p.e.rope_cs + ((int64_t)tok * (DH / 2) + (hc >> 1)) * 2);
        const float4 c0 = __ldg(cs), c1 = __ldg(cs + 1);   // (cos, sin) of the lane's 4 pairs
        const float cs8[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = v[2 * e], x1 = v[2 * e + 1];
          v[2 * e] = x0 * cs8[2 * e] - x1 * cs8[2 * e + 1];
          v[2 * e + 1] = x1 * cs8[2 * e] + x0 * cs8[2 * e + 1];
        }
      }
      if (FULL || i < rows)
        *reinterpret_cast<uint4*>(out) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                    pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
    __syncwarp();
  }
}

// chunk with its side inputs already in flight (see epilogue_tile_resid)
template <int EPI, bool FULL>
__device__ __forceinline__ void epilogue_chunk_f32_side(const Params& p, const ChunkSide<EPI>& side, uint32_t t_addr,
                                                        uint32_t stage_buf, int lane, int m0, int n0) {
  uint32_t r[32];
  tmem_ld_x32(t_addr, r);
  stage_chunk(stage_buf, lane, r);
  __syncwarp();
  epilogue_rows_f32<EPI, FULL>(p, side, stage_buf, lane, m0, n0);
  __syncwarp();
}

// Residual epilogues (RESID_F32, GATE_RESID_F32) are bound by HBM, not by the tensor pipe, whenever K is small (a
// 128 x 256 fp32 tile reads and writes 256 KB for 1-5 k cycles of MMA work), and the eight epilogue warps are all the
// memory-level parallelism a CTA has: with the residual of ONE chunk in flight per warp (8 x 16 B per lane, 32 KB per SM
// at best) these launches ran at ~60 % of the copy bandwidth.  The residual rows do not depend on the accumulator, so
// the warp's chunks are software-pipelined: chunk c+1's residual is requested before chunk c is transposed and stored,
// and the first chunk's before the wait for the accumulator — i.e. under the main loop of the tile.
#ifndef DFOT_GEMM_L2_PREFETCH_MAX_K
#define DFOT_GEMM_L2_PREFETCH_MAX_K 1536
#endif
template <int EPI, int BN_>
__device__ __forceinline__ void epilogue_tile_resid(const Params& p, uint32_t t_row, uint32_t stage_buf, int lane, int m0,
                                                    int n_base, int half, uint32_t full_bar, uint32_t parity) {
  auto valid = [&](int c) { return c < BN_ / 32 && n_base + c * 32 < p.N && m0 < p.M; };   // warp-uniform
  auto full = [&](int c) { return m0 + 32 <= p.M && n_base + c * 32 + 32 <= p.N; };
  auto fetch = [&](ChunkSide<EPI>& sd, int c) {
    if (full(c)) prefetch_side<EPI, true>(p, sd, lane, m0, n_base + c * 32);
    else prefetch_side<EPI, false>(p, sd, lane, m0, n_base + c * 32);
  };
  // ... and all of the warp's residual rows of this tile are pulled into L2 first (prefetch.global.L2 needs no
  // scoreboard, unlike a register prefetch, whose completion the loads of the chunk in hand wait for as well): the
  // register loads below then see L2 latency, and up to 16 KB per warp (128 KB per SM) of DRAM requests are in flight.
  // (scripts/micro/rowpiece_bench.cu: 8 warps x 4 KB in flight top out at 2.7-3.7 TB/s of the 6.5 TB/s copy peak.)
  // Measured on one box, M = 65536, N = K = 576: 93.8 -> 75-83 us; N = K = 1152: 58.4 -> 55-58 us; no gain for K >= 2304
  // (tensor-bound) and -2 % on the 128-channel convolutions, which are bound by the SM's TMA fill rate (9 taps re-read
  // the input tile: 94 B/clk/SM at the full MMA rate against the 69 B/clk/SM scripts/micro/tma_fill_bench.cu measures)
  // — so small-K plain GEMMs only.
#ifndef DFOT_GEMM_NO_L2_PREFETCH
  if (p.K <= DFOT_GEMM_L2_PREFETCH_MAX_K && p.conv_cblks == 0) {   // (measured: see the comment above)
    const int col = n_base + ((lane & 7) << 2), rsub = lane >> 3;
#pragma unroll 1
    for (int cc = half; valid(cc); cc += 2) {
      const int cn0 = col + cc * 32;
      if (cn0 < p.N) {
        const float* res = p.e.resid + (int64_t)(m0 + rsub) * p.e.ld_resid + cn0;
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if (m0 + 4 * it + rsub < p.M)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(res + (int64_t)(4 * it) * p.e.ld_resid));
      }
    }
  }
#endif
  ChunkSide<EPI> cur, nxt;
  int c = half;
  bool v = valid(c);
  if (v) fetch(cur, c);
  mbar_wait(full_bar, parity);
  tc_fence_after();
#pragma unroll 1
  while (v) {
    const int cn = c + 2;
    const bool vn = valid(cn);
    if (vn) fetch(nxt, cn);
    if (full(c)) epilogue_chunk_f32_side<EPI, true>(p, cur, t_row + (uint32_t)(c * 32), stage_buf, lane, m0, n_base + c * 32);
    else epilogue_chunk_f32_side<EPI, false>(p, cur, t_row + (uint32_t)(c * 32), stage_buf, lane, m0, n_base + c * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) cur.r[i] = nxt.r[i];
    c = cn;
    v = vn;
  }
}

template <int EPI, bool FULL>
__device__ __forceinline__ void epilogue_chunk(const Params& p, uint32_t t_addr, uint32_t stage_buf, int lane, int m0,
                                               int n0) {
  ChunkSide<EPI> side;
  prefetch_side<EPI, FULL>(p, side, lane, m0, n0);   // loads overlap the TMEM read + transpose below
  uint32_t r[32];
  tmem_ld_x32(t_addr, r);
  stage_chunk(stage_buf, lane, r);
  __syncwarp();
  if constexpr (EPI == DFOT_EPI_F32 || epi_has_resid(EPI))
    epilogue_rows_f32<EPI, FULL>(p, side, stage_buf, lane, m0, n0);
  else if constexpr (EPI == DFOT_EPI_QKV_ROPE_BF16)
    epilogue_rows_rope<FULL>(p, side, stage_buf, lane, m0, n0);
  else
    epilogue_rows_bf16<EPI, FULL>(p, stage_buf, lane, m0, n0);
  __syncwarp();
}

// ------------------------------------------------------------------ the kernel
template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const Params p) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh (pdl_wait() follows the prologue)
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms must start on 1024-byte boundaries
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::kStages + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * C::kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * C::kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::kStages + 4);
  const uint32_t epi_stage0 = bars + 256u;  // 8 x 4 KB transpose buffers (16-byte aligned)

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;   // (uniform for the compiler: common.cuh)
  const int num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.conv_cblks > 0 ? p.conv_taps * p.conv_cblks : (p.K + BK - 1) / BK;
  // banded rasterisation: consecutive tiles walk the n-blocks of a band of kRasterBand m-blocks, so the CTAs
  // resident at any time share a small set of A and W tiles in L2
  auto tile_coord = [&](int tile, int& m_blk, int& n_blk) {
    const int band_tiles = kRasterBand * num_n;
    const int band = tile / band_tiles, in_band = tile - band * band_tiles;
    const int band_m0 = band * kRasterBand;
    const int band_h = min(kRasterBand, num_m - band_m0);
    m_blk = band_m0 + in_band % band_h;
    n_blk = in_band / band_h;
  };
  // split-K (Params::split_k): weight n-block and k-block range [kb0, kb1) of output n-block n_blk
  const int num_n_w = num_n / p.split_k;
  auto k_range = [&](int n_blk, int& n_blk_w, int& kb0, int& kb1) {
    const int sp = n_blk / num_n_w;
    n_blk_w = n_blk - sp * num_n_w;
    kb0 = (int)((int64_t)sp * num_kb / p.split_k);
    kb1 = (int)((int64_t)(sp + 1) * num_kb / p.split_k);
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), kEpiWarps);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();   // barriers, tensor memory and tensor maps are set up; global memory is only touched from here on

  if (warp == 0) {
    // ===================== TMA producer (all lanes walk the loop, the copies sit under elect.sync: common.cuh) ========
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coord(tile, m_blk, n_blk);
        if (p.conv_cblks == 0) {
          int n_blk_w, kb0, kb1;
          k_range(n_blk, n_blk_w, kb0, kb1);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (elect_one_sync()) {
              mbar_expect_tx(full_bar(stage), C::kStageBytes);
              const uint32_t sa = smem_base + stage * C::kStageBytes;
              tma_load_2d(sa, &tma_a, full_bar(stage), kb * BK, m_blk * BM);
              tma_load_2d(sa + C::kStageBytesA, &tma_b, full_bar(stage), kb * BK, n_blk_w * BN);
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
        } else {
          // pixel tile origin (the 128 rows of a tile are bw x bh x bn pixels; host guarantees the divisibility)
          const int m0 = m_blk * BM;
          const int x0 = m0 % p.conv_W, y0 = (m0 / p.conv_W) % p.conv_H, img0 = m0 / (p.conv_W * p.conv_H);
          int tap = 0, cb = 0;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (elect_one_sync()) {
              mbar_expect_tx(full_bar(stage), C::kStageBytes);
              const uint32_t sa = smem_base + stage * C::kStageBytes;
              const int dt = tap / 9, sp = tap - dt * 9;
              const int dy = sp / 3 - 1, dx = sp - (sp / 3) * 3 - 1;
              tma_load_4d(sa, &tma_a, full_bar(stage), cb * BK, x0 + dx, y0 + dy, img0 + dt);
              tma_load_3d(sa + C::kStageBytesA, &tma_b, full_bar(stage), cb * BK, tap, n_blk * BN);
            }
            __syncwarp();
            if (++cb == p.conv_cblks) { cb = 0; ++tap; }
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, tcgen05 instructions under elect.sync) =====================
    {
      constexpr uint32_t idesc = make_idesc<BN>();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        int kb0 = 0, kb1 = num_kb;
        if (p.split_k > 1) {                               // (conv launches never split)
          int m_blk, n_blk, n_blk_w;
          tile_coord(tile, m_blk, n_blk);
          k_range(n_blk, n_blk_w, kb0, kb1);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);               // TMA bytes have landed
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + stage * C::kStageBytes;
            const uint64_t a_desc = make_smem_desc_sw128(sa);
            const uint64_t b_desc = make_smem_desc_sw128(sa + C::kStageBytesA);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
              umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));                 // ring slot reusable once these MMAs retire
            if (kb + 1 == kb1) umma_commit(tmem_full_bar(acc));   // accumulator complete → epilogue
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;          // TMEM lane quarter this warp may access (hardware rule: warp_id % 4)
    const int half = (warp - 2) >> 2;  // the two warps of a quarter alternate over the 32-column chunks
    const uint32_t stage_buf = epi_stage0 + (uint32_t)(warp - 2) * C::kEpiStageBytes;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coord(tile, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m0 = m_blk * BM + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (epi_has_resid(EPI)) {
        epilogue_tile_resid<EPI, BN>(p, t_row, stage_buf, lane, m0, n_blk * BN, half, tmem_full_bar(acc), acc_phase);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        continue;
      }
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      if constexpr (EPI == DFOT_EPI_QKNORM_ROPE_BF16) {
        // head-wise: the two warps of a lane quarter alternate over the heads of the 256-column tile
        const int dh = (int)p.e.head_dim;               // 64 or 128 (host-checked), N % dh == 0
#pragma unroll 1
        for (int hh = half; hh * dh < BN; hh += 2) {
          const int n0 = n_blk * BN + hh * dh;
          if (n0 >= p.N || m0 >= p.M) break;  // warp-uniform
          const uint32_t ta = t_row + (uint32_t)(hh * dh);
          if (dh == 64) {
            if (m0 + 32 <= p.M) epilogue_head_qknorm<64, true>(p, ta, stage_buf, lane, m0, n0);
            else epilogue_head_qknorm<64, false>(p, ta, stage_buf, lane, m0, n0);
          } else {
            if (m0 + 32 <= p.M) epilogue_head_qknorm<128, true>(p, ta, stage_buf, lane, m0, n0);
            else epilogue_head_qknorm<128, false>(p, ta, stage_buf, lane, m0, n0);
          }
        }
      } else {
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= p.N || m0 >= p.M) break;  // warp-uniform
          if (m0 + 32 <= p.M && n0 + 32 <= p.N)
            epilogue_chunk<EPI, true>(p, t_row + (uint32_t)(c * 32), stage_buf, lane, m0, n0);
          else
            epilogue_chunk<EPI, false>(p, t_row + (uint32_t)(c * 32), stage_buf, lane, m0, n0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ the CTA-pair kernel (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) compute one 256 x 256 tile: CTA r holds rows [128r, 128r+128) of A and of
// the accumulator, and HALF of the B tile (rows [128r, 128r+128) of the 256 output columns); the MMA — issued by the
// leader CTA only, UMMA M = 256 — reads the other half from the peer's shared memory.  Per k-block every SM now moves
// 32 KB (16 A + 16 B) through L2 / TMA / its shared-memory port instead of 48 KB for the same 128 x 256 x 64 of MMA work:
// the operand traffic that caps the single-CTA kernel near 2/3 of the tensor rate drops by a third, and the ring gets
// six stages.  Barriers: every TMA (both CTAs) completes on the LEADER's full barrier; the MMA's commits are multicast
// to both CTAs' empty / tmem_full barriers; both CTAs' epilogue warps arrive on the LEADER's tmem_empty barrier.
template <int BN_> struct Cfg2 {                            // BN_ in {128, 192, 256}: columns of the pair's tile
  static constexpr int BN = BN_;
  static constexpr int kStageBytesA = BM * BK * 2;          // this CTA's 128 rows of A
  static constexpr int kStageBytesB = (BN / 2) * BK * 2;    // this CTA's half of the B tile
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = BN == 256 ? 6 : (BN == 192 ? 6 : 8);
  static constexpr int kTmemCols = BN == 128 ? 256 : 512;   // two accumulator stages at columns 0 and BN
  static constexpr int kEpiStageBytes = 32 * 32 * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024 + 256;
};
template <int BN> __device__ __forceinline__ constexpr uint32_t make_idesc_pair() {   // M = 256 (pair), N = BN
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                          const Params p) {
  pdl_trigger();   // programmatic dependent launch: see common.cuh (pdl_wait() follows the prologue)
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::kStages + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * C::kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * C::kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::kStages + 4);
  const uint32_t epi_stage0 = bars + 256u;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;   // (uniform for the compiler: common.cuh)
  const int rank = (int)cluster_ctarank();              // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (p.M + 2 * BM - 1) / (2 * BM), num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.conv_cblks > 0 ? p.conv_taps * p.conv_cblks : (p.K + BK - 1) / BK;
  auto tile_coord = [&](int tile, int& m_blk, int& n_blk) {   // banded rasterisation (bands of 8 pair-rows = 2048 rows)
    constexpr int kBand = kRasterBand / 2;
    const int band_tiles = kBand * num_n;
    const int band = tile / band_tiles, in_band = tile - band * band_tiles;
    const int band_m0 = band * kBand;
    const int band_h = min(kBand, num_m - band_m0);
    m_blk = band_m0 + in_band % band_h;
    n_blk = in_band / band_h;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);                // leader's copy: one expect_tx arrival, bytes of both CTAs
      mbar_init(empty_bar(s), 1);               // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);           // multicast commit
      mbar_init(tmem_empty_bar(a), 2 * kEpiWarps);   // leader's copy: epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem2_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  cluster_sync_all();                            // barriers of both CTAs are initialised before anyone signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();   // barriers, tensor memory and tensor maps are set up; global memory is only touched from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own A rows, own half of B) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        tile_coord(tile, m_blk, n_blk);
        const int m0 = m_blk * 2 * BM + rank * BM;
        const int nb0 = n_blk * BN + rank * (BN / 2);
        const int x0 = m0 % p.conv_W, y0 = (m0 / p.conv_W) % p.conv_H, img0 = m0 / (p.conv_W * p.conv_H);
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::kStageBytes);
            const uint32_t sa = smem_base + stage * C::kStageBytes;
            if (p.conv_cblks == 0) {
              tma2_load_2d(sa, &tma_a, full_bar(stage), kb * BK, m0);
              tma2_load_2d(sa + C::kStageBytesA, &tma_b, full_bar(stage), kb * BK, nb0);
            } else {
              const int dt = tap / 9, sp = tap - dt * 9;
              const int dy = sp / 3 - 1, dx = sp - (sp / 3) * 3 - 1;
              tma2_load_4d(sa, &tma_a, full_bar(stage), cb * BK, x0 + dx, y0 + dy, img0 + dt);
              tma2_load_3d(sa + C::kStageBytesA, &tma_b, full_bar(stage), cb * BK, tap, nb0);
            }
          }
          __syncwarp();
          if (p.conv_cblks != 0 && ++cb == p.conv_cblks) { cb = 0; ++tap; }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {                                     // warp-uniform loop, tcgen05 instructions under elect.sync
      constexpr uint32_t idesc = make_idesc_pair<BN>();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);    // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);                 // both CTAs' TMA bytes have landed
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + stage * C::kStageBytes;
            const uint64_t a_desc = make_smem_desc_sw128(sa);
            const uint64_t b_desc = make_smem_desc_sw128(sa + C::kStageBytesA);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma2_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
            umma2_commit_both(empty_bar(stage));             // ring slot reusable in BOTH CTAs
            if (kb + 1 == num_kb) umma2_commit_both(tmem_full_bar(acc));   // accumulator complete → both epilogues
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs: their own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t stage_buf = epi_stage0 + (uint32_t)(warp - 2) * C::kEpiStageBytes;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      int m_blk, n_blk;
      tile_coord(tile, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m0 = m_blk * 2 * BM + rank * BM + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (epi_has_resid(EPI)) {
        epilogue_tile_resid<EPI, BN>(p, t_row, stage_buf, lane, m0, n_blk * BN, half, tmem_full_bar(acc), acc_phase);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tmem_empty_bar(acc));
        continue;
      }
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      if constexpr (EPI == DFOT_EPI_QKNORM_ROPE_BF16) {
        const int dh = (int)p.e.head_dim;               // 64 or 128 (host-checked), N % dh == 0, BN % dh == 0
#pragma unroll 1
        for (int hh = half; hh * dh < BN; hh += 2) {
          const int n0 = n_blk * BN + hh * dh;
          if (n0 >= p.N || m0 >= p.M) break;  // warp-uniform
          const uint32_t ta = t_row + (uint32_t)(hh * dh);
          if (dh == 64) {
            if (m0 + 32 <= p.M) epilogue_head_qknorm<64, true>(p, ta, stage_buf, lane, m0, n0);
            else epilogue_head_qknorm<64, false>(p, ta, stage_buf, lane, m0, n0);
          } else {
            if (m0 + 32 <= p.M) epilogue_head_qknorm<128, true>(p, ta, stage_buf, lane, m0, n0);
            else epilogue_head_qknorm<128, false>(p, ta, stage_buf, lane, m0, n0);
          }
        }
      } else {
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= p.N || m0 >= p.M) break;  // warp-uniform
          if (m0 + 32 <= p.M && n0 + 32 <= p.N)
            epilogue_chunk<EPI, true>(p, t_row + (uint32_t)(c * 32), stage_buf, lane, m0, n0);
          else
            epilogue_chunk<EPI, false>(p, t_row + (uint32_t)(c * 32), stage_buf, lane, m0, n0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tmem_empty_bar(acc));
    }
  }
  tc_fence_before();
  cluster_sync_all();                            // the peer may still be reading this CTA's shared memory / signalling it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem2_dealloc(tmem_base, C::kTmemCols);
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

// 2D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64], 128B swizzle
static int make_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  DFOT_REQUIRE(enc != nullptr, DFOT_ERR_DRIVER, "gemm: cuTensorMapEncodeTiled unavailable from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFOT_REQUIRE(r == CUDA_SUCCESS, DFOT_ERR_DRIVER, "gemm: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DFOT_OK;
}

// bf16 tensor of `rank` dims (dim 0 contiguous), box[0] = 64 elements = one 128-byte swizzle atom
static int make_tmap_nd(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* gdim, const cuuint64_t* gstr_bytes,
                        const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode_fn();
  DFOT_REQUIRE(enc != nullptr, DFOT_ERR_DRIVER, "conv: cuTensorMapEncodeTiled unavailable from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr_bytes,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFOT_REQUIRE(r == CUDA_SUCCESS, DFOT_ERR_DRIVER, "conv: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DFOT_OK;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int EPI>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t s) {
  using C = Cfg<BN>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    DFOT_REQUIRE(e == cudaSuccess, DFOT_ERR_CUDA, "gemm: cannot reserve %d B of shared memory: %s", C::kSmemBytes,
                 cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = (int)(ceil_div(p.M, BM) * ceil_div(p.N, BN));
  const int grid = tiles < num_sms() ? tiles : num_sms();
  launch_pdl(kern, dim3(grid), dim3(kThreads), C::kSmemBytes, s, ta, tb, p);
  DFOT_CHECK_LAUNCH("gemm_bf16_tcgen05");
  return DFOT_OK;
}

template <int BN, int EPI>
static int launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t s) {
  using C = Cfg2<BN>;
  auto kern = gemm2_bf16_tcgen05_kernel<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    DFOT_REQUIRE(e == cudaSuccess, DFOT_ERR_CUDA, "gemm: cannot reserve %d B of shared memory: %s", C::kSmemBytes,
                 cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = (int)(ceil_div(p.M, 2 * BM) * ceil_div(p.N, BN));
  int pairs = num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), C::kSmemBytes, s, ta, tb, p);   // static __cluster_dims__(2, 1, 1)
  DFOT_CHECK_LAUNCH("gemm2_bf16_tcgen05");
  return DFOT_OK;
}
template <int BN>
static int dispatch_pair(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t s) {
  switch (epi) {
    case DFOT_EPI_F32: return launch_pair<BN, DFOT_EPI_F32>(ta, tb, p, s);
    case DFOT_EPI_BF16: return launch_pair<BN, DFOT_EPI_BF16>(ta, tb, p, s);
    case DFOT_EPI_GELU_BF16: return launch_pair<BN, DFOT_EPI_GELU_BF16>(ta, tb, p, s);
    case DFOT_EPI_SILU_BF16: return launch_pair<BN, DFOT_EPI_SILU_BF16>(ta, tb, p, s);
    case DFOT_EPI_GATE_RESID_F32: return launch_pair<BN, DFOT_EPI_GATE_RESID_F32>(ta, tb, p, s);
    case DFOT_EPI_GATE_LNRESID_F32: return launch_pair<BN, DFOT_EPI_GATE_LNRESID_F32>(ta, tb, p, s);
    case DFOT_EPI_QKV_ROPE_BF16: return launch_pair<BN, DFOT_EPI_QKV_ROPE_BF16>(ta, tb, p, s);
    case DFOT_EPI_RESID_F32: return launch_pair<BN, DFOT_EPI_RESID_F32>(ta, tb, p, s);
    case DFOT_EPI_QKNORM_ROPE_BF16:
      if constexpr (BN == 256) return launch_pair<256, DFOT_EPI_QKNORM_ROPE_BF16>(ta, tb, p, s);
      break;
  }
  set_error("gemm: epilogue %d has no CTA-pair kernel at this tile width", epi);
  return DFOT_ERR_INVALID_ARG;
}
static int dispatch_pair_bn(int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t s) {
  if (bn == 256) return dispatch_pair<256>(epi, ta, tb, p, s);
  if (bn == 192) return dispatch_pair<192>(epi, ta, tb, p, s);
  return dispatch_pair<128>(epi, ta, tb, p, s);
}
// Tile width of the CTA-pair kernel (0 = use the single-CTA kernel): the width with the fewest padded columns (wider on
// ties), provided every SM pair gets at least one 256-row tile.  DFOT_GEMM_PAIR=0|1 pins the choice (benchmarking).
static int pick_pair_bn(int64_t M, int64_t N, int epilogue) {
  static int ov = -2;
  if (ov == -2) {
    const char* e = getenv("DFOT_GEMM_PAIR");
    ov = e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }
  if (ov == 0 || N < 128) return 0;
  int bn = 256;
  static int force_bn = -1;                          // DFOT_GEMM_BN=128|192|256 pins the pair tile width (benchmarking)
  if (force_bn < 0) {
    const char* e = getenv("DFOT_GEMM_BN");
    force_bn = e == nullptr ? 0 : atoi(e);
    if (force_bn != 128 && force_bn != 192 && force_bn != 256) force_bn = 0;
  }
  if (force_bn != 0 && epilogue != DFOT_EPI_QKNORM_ROPE_BF16) {
    bn = force_bn;
  } else if (epilogue != DFOT_EPI_QKNORM_ROPE_BF16) {
    const int64_t pad256 = ceil_div(N, 256) * 256, pad192 = ceil_div(N, 192) * 192, pad128 = ceil_div(N, 128) * 128;
    // (measured r02, M = 16384: N = 1152 -> 192 columns 139 us vs 256 columns 146 us at K = 4608; N = 3456 is a tie)
    if (pad192 * 100 <= pad256 * 92) bn = 192;
    if (pad128 * 100 <= (bn == 256 ? pad256 : pad192) * 85) bn = 128;
  }
  const int64_t tiles = ceil_div(M, 2 * BM) * ceil_div(N, bn);
  if (ov != 1 && tiles < num_sms() / 2) return 0;
  return bn;
}

// GroupNorm side output: validate, zero the workspace before the launch (gn_begin) and finalise after it (gn_end)
static int gn_begin(const Params& p, int epilogue, cudaStream_t s) {
  if (p.e.gn_sums == nullptr) return DFOT_OK;
  DFOT_REQUIRE(epilogue == DFOT_EPI_F32 || epilogue == DFOT_EPI_RESID_F32 || epilogue == DFOT_EPI_BF16,
               DFOT_ERR_UNSUPPORTED, "gemm: GroupNorm side output needs the F32, RESID_F32 or BF16 epilogue");
  const int64_t G = p.e.gn_groups, rpi = p.e.gn_rows_per_img;
  DFOT_REQUIRE(G > 0 && p.N % G == 0 && rpi > 0 && rpi % 32 == 0 && p.M % rpi == 0, DFOT_ERR_UNSUPPORTED,
               "gemm: GroupNorm side output needs N %% groups == 0, rows_per_img %% 32 == 0 and M %% rows_per_img == 0");
  const int64_t cpg = p.N / G;
  DFOT_REQUIRE((cpg & (cpg - 1)) == 0, DFOT_ERR_UNSUPPORTED, "gemm: channels per group (%lld) must be a power of two",
               (long long)cpg);
  return gn_zero_sums(p.e.gn_sums, p.M / rpi, G, s);
}
static int gn_end(const Params& p, cudaStream_t s) {
  if (p.e.gn_sums == nullptr) return DFOT_OK;
  const int64_t G = p.e.gn_groups, rpi = p.e.gn_rows_per_img;
  return gn_finalize(p.e.gn_sums, p.M / rpi, G, rpi * (p.N / G), p.e.gn_eps, s);
}

// Tile width: the widest tile unless 192 columns (3 x 64, a valid UMMA N) leave at least 15 % fewer padded columns —
// e.g. N = 576 = 3 x 192 instead of 3 x 256 (measured: 225 -> 205 us at K = 2880).  A 10 % saving (N = 1152) does not
// pay for the narrower tile's worse operand-bandwidth ratio at large K (measured: 103 -> 113 us at K = 4608).
static int pick_bn(int64_t M, int64_t N, int epilogue) {
  if (N <= 64) return 64;
  if (epilogue == DFOT_EPI_QKNORM_ROPE_BF16) return 256;
  int bn = 128;
  if (N > 128) {
    const int64_t pad256 = ceil_div(N, 256) * 256, pad192 = ceil_div(N, 192) * 192;
    bn = (pad192 * 100 <= pad256 * 85) ? 192 : 256;
  }
  // latency regime (small batches: fewer tiles than SMs): narrower tiles spread the problem over more SMs and shorten
  // every CTA's serial k-loop.  Time model per wave: MMA time proportional to the tile width plus a fixed per-tile cost
  // (pipeline fill, epilogue tail) worth about 64 columns; the width with the fewest wave-units wins.
  const int64_t m_tiles = ceil_div(M, BM), sms = num_sms();
  if (m_tiles * ceil_div(N, bn) < sms) {
    int64_t best = ceil_div(m_tiles * ceil_div(N, bn), sms) * (bn + 64);
    for (int cand : {192, 128, 64}) {
      if (cand >= bn) continue;
      const int64_t t = ceil_div(m_tiles * ceil_div(N, cand), sms) * (cand + 64);
      if (t < best) { best = t; bn = cand; }
    }
  }
  return bn;
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t s) {
  switch (epi) {
    case DFOT_EPI_F32: return launch<BN, DFOT_EPI_F32>(ta, tb, p, s);
    case DFOT_EPI_BF16: return launch<BN, DFOT_EPI_BF16>(ta, tb, p, s);
    case DFOT_EPI_GELU_BF16: return launch<BN, DFOT_EPI_GELU_BF16>(ta, tb, p, s);
    case DFOT_EPI_SILU_BF16: return launch<BN, DFOT_EPI_SILU_BF16>(ta, tb, p, s);
    case DFOT_EPI_GATE_RESID_F32: return launch<BN, DFOT_EPI_GATE_RESID_F32>(ta, tb, p, s);
    case DFOT_EPI_GATE_LNRESID_F32: return launch<BN, DFOT_EPI_GATE_LNRESID_F32>(ta, tb, p, s);
    case DFOT_EPI_QKV_ROPE_BF16: return launch<BN, DFOT_EPI_QKV_ROPE_BF16>(ta, tb, p, s);
    case DFOT_EPI_RESID_F32: return launch<BN, DFOT_EPI_RESID_F32>(ta, tb, p, s);
    case DFOT_EPI_QKNORM_ROPE_BF16:
      if constexpr (BN == 256) return launch<256, DFOT_EPI_QKNORM_ROPE_BF16>(ta, tb, p, s);
      break;
  }
  set_error("gemm: unknown epilogue %d", epi);
  return DFOT_ERR_INVALID_ARG;
}

}  // namespace gemm
}  // namespace dfot

extern "C" int dfot_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* Cout, int64_t ldc,
                              int64_t M, int64_t N, int64_t K, int epilogue, const dfot_gemm_epilogue* epi,
                              void* stream) {
  using namespace dfot;
  using namespace dfot::gemm;
  DFOT_REQUIRE(A && W && Cout && epi, DFOT_ERR_INVALID_ARG, "gemm: null pointer");
  DFOT_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31),
               DFOT_ERR_INVALID_ARG, "gemm: bad sizes M=%lld N=%lld K=%lld", (long long)M, (long long)N,
               (long long)K);
  DFOT_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, DFOT_ERR_UNSUPPORTED,
               "gemm: K, lda, ldw must be multiples of 8 (16-byte TMA strides); got K=%lld lda=%lld ldw=%lld",
               (long long)K, (long long)lda, (long long)ldw);
  DFOT_REQUIRE(N % 8 == 0 && ldc % 8 == 0 && ldc >= N, DFOT_ERR_UNSUPPORTED,
               "gemm: N and ldc must be multiples of 8; got N=%lld ldc=%lld", (long long)N, (long long)ldc);
  DFOT_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)Cout % 16 == 0),
               DFOT_ERR_UNSUPPORTED, "gemm: A, W, C must be 16-byte aligned");
  // the epilogue reads bias / residual / gate and writes the output with 16-byte accesses
  DFOT_REQUIRE(epi->bias == nullptr || (uintptr_t)epi->bias % 16 == 0, DFOT_ERR_UNSUPPORTED,
               "gemm: bias must be 16-byte aligned");
  const bool gated = epilogue == DFOT_EPI_GATE_RESID_F32 || epilogue == DFOT_EPI_GATE_LNRESID_F32;
  if (gated)
    DFOT_REQUIRE(epi->resid && epi->gate && epi->tokens_per_frame >= 1 && epi->tokens_per_frame < (1ll << 30),
                 DFOT_ERR_INVALID_ARG, "gemm: GATE_RESID needs resid, gate and tokens_per_frame");
  if (gated)
    DFOT_REQUIRE((uintptr_t)epi->gate % 16 == 0 && epi->ld_gate % 4 == 0, DFOT_ERR_UNSUPPORTED,
                 "gemm: gate must be 16-byte aligned with ld_gate %% 4 == 0");
  if (epilogue == DFOT_EPI_GATE_LNRESID_F32)
    DFOT_REQUIRE(epi->ln_stats && epi->ln_shift && epi->ln_scale && (uintptr_t)epi->ln_stats % 8 == 0 &&
                     (uintptr_t)epi->ln_shift % 16 == 0 && (uintptr_t)epi->ln_scale % 16 == 0,
                 DFOT_ERR_INVALID_ARG, "gemm: GATE_LNRESID needs ln_stats (8-byte aligned), ln_shift and ln_scale (16-byte aligned)");
  if (gated || epilogue == DFOT_EPI_RESID_F32)
    DFOT_REQUIRE(epi->resid != nullptr && (uintptr_t)epi->resid % 16 == 0 && epi->ld_resid % 4 == 0, DFOT_ERR_UNSUPPORTED,
                 "gemm: resid must be 16-byte aligned with ld_resid %% 4 == 0");
  if (epilogue == DFOT_EPI_RESID_F32)
    DFOT_REQUIRE(epi->resid != nullptr && epi->ld_resid >= N, DFOT_ERR_INVALID_ARG, "gemm: RESID needs resid");
  if (epilogue == DFOT_EPI_QKNORM_ROPE_BF16)
    DFOT_REQUIRE(epi->rope_cs && epi->qn_w && epi->kn_w && epi->tokens_per_sample > 0 &&
                     (epi->head_dim == 64 || epi->head_dim == 128) && epi->model_dim % epi->head_dim == 0 &&
                     N == 3 * epi->model_dim && N > 128 && (uintptr_t)epi->rope_cs % 16 == 0 &&
                     (uintptr_t)epi->qn_w % 16 == 0 && (uintptr_t)epi->kn_w % 16 == 0,
                 DFOT_ERR_INVALID_ARG,
                 "gemm: QKNORM_ROPE needs rope table, q/k norm weights (16-byte aligned), head_dim 64|128, N = 3D > 128");
  if (epilogue == DFOT_EPI_QKV_ROPE_BF16)
    DFOT_REQUIRE(epi->rope_cs && epi->tokens_per_sample > 0 && epi->head_dim > 0 && epi->head_dim % 2 == 0 &&
                     epi->model_dim > 0 && epi->model_dim % epi->head_dim == 0 && N == 3 * epi->model_dim,
                 DFOT_ERR_INVALID_ARG, "gemm: QKV_ROPE needs rope table, tokens_per_sample, head_dim | model_dim, N=3D");
  Params p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.C = Cout; p.ldc = ldc; p.e = *epi;
  p.conv_cblks = 0; p.conv_W = p.conv_H = 1; p.conv_taps = 9; p.split_k = 1;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, M, K, lda, BM);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = gn_begin(p, epilogue, s))) return rc;
  if (const int pbn = pick_pair_bn(M, N, epilogue)) {
    rc = make_tmap(&tb, W, N, K, ldw, pbn / 2);          // each CTA loads half of the pair's B tile
    if (!rc) rc = dispatch_pair_bn(pbn, epilogue, ta, tb, p, s);
    return rc ? rc : gn_end(p, s);
  }
  const int bn = pick_bn(M, N, epilogue);
  if (bn == 256) {
    rc = make_tmap(&tb, W, N, K, ldw, 256);
    if (!rc) rc = dispatch_epi<256>(epilogue, ta, tb, p, s);
  } else if (bn == 192) {
    rc = make_tmap(&tb, W, N, K, ldw, 192);
    if (!rc) rc = dispatch_epi<192>(epilogue, ta, tb, p, s);
  } else if (bn == 128) {
    rc = make_tmap(&tb, W, N, K, ldw, 128);
    if (!rc) rc = dispatch_epi<128>(epilogue, ta, tb, p, s);
  } else {
    rc = make_tmap(&tb, W, N, K, ldw, 64);
    if (!rc) rc = dispatch_epi<64>(epilogue, ta, tb, p, s);
  }
  return rc ? rc : gn_end(p, s);
}

// Split-K for the latency regime (a few hundred rows: fewer output tiles than SMs and a long serial k-loop per CTA, e.g.
// DiT-B fc2 at batch 1: 256 x 768 x 3072 = 24 tiles of 48 k-blocks, bound by one SM's TMA fill rate).  The k-blocks are
// dealt to `splits` CTAs per output tile; split s writes its fp32 partial sums to columns [s*N, (s+1)*N) of
// parts[M, splits*N] — no atomics, the consumer (dfot_splitk_gate_resid_adaln) adds them in a fixed order.
extern "C" int dfot_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* parts, int64_t M,
                                     int64_t N, int64_t K, int64_t splits, void* stream) {
  using namespace dfot;
  using namespace dfot::gemm;
  DFOT_REQUIRE(A && W && parts, DFOT_ERR_INVALID_ARG, "gemm_splitk: null pointer");
  DFOT_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1 && M < (1ll << 31) && N * splits < (1ll << 31) && K < (1ll << 31),
               DFOT_ERR_INVALID_ARG, "gemm_splitk: bad sizes M=%lld N=%lld K=%lld splits=%lld", (long long)M, (long long)N,
               (long long)K, (long long)splits);
  DFOT_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, DFOT_ERR_UNSUPPORTED,
               "gemm_splitk: K, lda, ldw must be multiples of 8 (16-byte TMA strides)");
  DFOT_REQUIRE(N % 64 == 0, DFOT_ERR_UNSUPPORTED, "gemm_splitk: N must be a multiple of 64 (whole n-blocks per split)");
  DFOT_REQUIRE(splits <= ceil_div(K, BK), DFOT_ERR_INVALID_ARG, "gemm_splitk: more splits (%lld) than k-blocks (%lld)",
               (long long)splits, (long long)ceil_div(K, BK));
  DFOT_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)parts % 16 == 0), DFOT_ERR_UNSUPPORTED,
               "gemm_splitk: A, W, parts must be 16-byte aligned");
  Params p;
  memset(&p.e, 0, sizeof(p.e));
  p.M = (int)M; p.N = (int)(N * splits); p.K = (int)K; p.C = parts; p.ldc = N * splits;
  p.conv_cblks = 0; p.conv_W = p.conv_H = 1; p.conv_taps = 9; p.split_k = (int)splits;
  p.e.tokens_per_frame = 1; p.e.tokens_per_sample = 1;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, M, K, lda, BM);
  if (rc) return rc;
  // the widest n-block that divides N and still leaves the tiles x splits grid at or under one wave
  const int64_t m_tiles = ceil_div(M, BM);
  int bn = 64;
  if (N % 128 == 0 && m_tiles * (N / 64) * splits > num_sms()) bn = 128;
  rc = make_tmap(&tb, W, N, K, ldw, bn);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  return bn == 128 ? launch<128, DFOT_EPI_F32>(ta, tb, p, s) : launch<64, DFOT_EPI_F32>(ta, tb, p, s);
}

// 3x3 convolution, stride 1, zero padding 1, over NHWC bf16 activations as an implicit GEMM on the same kernel:
// M = n_img*H*W pixels, N = Cout, K = 9*Cin.  No im2col buffer exists: tap (dy, dx) of a pixel tile is the tile's
// TMA box moved by (dy, dx); rows/columns that fall outside the image are zero-filled by the TMA unit.
// kt = 1: 3x3 convolution of n_img images.  kt > 1: "causal" kt x 3 x 3 convolution along a frame axis — x holds
// n_img + kt - 1 frames and output frame j = sum over dt < kt of conv3x3(x[j + dt], w[:, dt]); the caller lays the frames
// out so that x[j .. j + kt - 1] are frame j's causal window (first frame repeated in front, see dfot_b200.h).
static int conv_impl(const void* x, const void* w, void* out, int64_t ldc, int64_t n_img, int64_t H, int64_t W,
                     int64_t Cin, int64_t Cout, int64_t kt, int epilogue, const dfot_gemm_epilogue* epi, void* stream) {
  using namespace dfot;
  using namespace dfot::gemm;
  DFOT_REQUIRE(x && w && out && epi, DFOT_ERR_INVALID_ARG, "conv3x3: null pointer");
  DFOT_REQUIRE(kt >= 1 && kt <= 7, DFOT_ERR_INVALID_ARG, "conv3d: temporal kernel size %lld not in [1, 7]", (long long)kt);
  DFOT_REQUIRE(n_img > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && n_img * H * W < (1ll << 31),
               DFOT_ERR_INVALID_ARG, "conv3x3: bad sizes");
  DFOT_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && ldc % 8 == 0 && ldc >= Cout, DFOT_ERR_UNSUPPORTED,
               "conv3x3: Cin, Cout, ldc must be multiples of 8; got %lld %lld %lld", (long long)Cin, (long long)Cout,
               (long long)ldc);
  DFOT_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0),
               DFOT_ERR_UNSUPPORTED, "conv3x3: x, w, out must be 16-byte aligned");
  DFOT_REQUIRE(epilogue == DFOT_EPI_F32 || epilogue == DFOT_EPI_BF16 || epilogue == DFOT_EPI_RESID_F32 ||
                   epilogue == DFOT_EPI_SILU_BF16,
               DFOT_ERR_UNSUPPORTED, "conv3x3: epilogue %d unsupported", epilogue);
  if (epilogue == DFOT_EPI_RESID_F32)
    DFOT_REQUIRE(epi->resid != nullptr && epi->ld_resid >= Cout && (uintptr_t)epi->resid % 16 == 0 &&
                     epi->ld_resid % 4 == 0,
                 DFOT_ERR_INVALID_ARG, "conv3x3: RESID needs a 16-byte aligned resid with ld_resid %% 4 == 0");
  DFOT_REQUIRE(epi->bias == nullptr || (uintptr_t)epi->bias % 16 == 0, DFOT_ERR_UNSUPPORTED,
               "conv3x3: bias must be 16-byte aligned");
  // tile = bw x bh x bn pixels = 128 GEMM rows in (img, y, x) order
  const bool pow2w = (W & (W - 1)) == 0, pow2h = (H & (H - 1)) == 0;
  int64_t bw, bh, bn;
  if (W >= BM) {
    DFOT_REQUIRE(W % BM == 0, DFOT_ERR_UNSUPPORTED, "conv3x3: W=%lld must be a multiple of 128 or a power of two",
                 (long long)W);
    bw = BM; bh = 1; bn = 1;
  } else {
    DFOT_REQUIRE(pow2w, DFOT_ERR_UNSUPPORTED, "conv3x3: W=%lld must be a power of two below 128", (long long)W);
    bw = W;
    if (H * W >= BM) {
      DFOT_REQUIRE(H % (BM / W) == 0, DFOT_ERR_UNSUPPORTED, "conv3x3: H=%lld must be a multiple of %lld",
                   (long long)H, (long long)(BM / W));
      bh = BM / W; bn = 1;
    } else {
      DFOT_REQUIRE(pow2h, DFOT_ERR_UNSUPPORTED, "conv3x3: H=%lld must be a power of two for small images", (long long)H);
      bh = H; bn = BM / (W * H);
    }
  }
  Params p;
  p.M = (int)(n_img * H * W); p.N = (int)Cout; p.K = (int)(9 * kt * Cin); p.C = out; p.ldc = ldc; p.e = *epi;
  p.conv_cblks = (int)ceil_div(Cin, BK); p.conv_W = (int)W; p.conv_H = (int)H; p.conv_taps = (int)(9 * kt); p.split_k = 1;
  CUtensorMap ta, tb;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(n_img + kt - 1)};
    cuuint64_t gstr[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    int rc = make_tmap_nd(&ta, x, 4, gdim, gstr, box);
    if (rc) return rc;
  }
  const int pbn = pick_pair_bn(n_img * H * W, Cout, epilogue);
  const bool pair = pbn != 0;
  const int bnt = pair ? pbn / 2 : pick_bn(n_img * H * W, Cout, epilogue);
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Cin, (cuuint64_t)(9 * kt), (cuuint64_t)Cout};
    cuuint64_t gstr[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)(9 * kt) * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, 1, (cuuint32_t)bnt};
    int rc = make_tmap_nd(&tb, w, 3, gdim, gstr, box);
    if (rc) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int rc = gn_begin(p, epilogue, s);
  if (rc) return rc;
  if (pair) rc = dispatch_pair_bn(pbn, epilogue, ta, tb, p, s);
  else if (bnt == 256) rc = dispatch_epi<256>(epilogue, ta, tb, p, s);
  else if (bnt == 192) rc = dispatch_epi<192>(epilogue, ta, tb, p, s);
  else if (bnt == 128) rc = dispatch_epi<128>(epilogue, ta, tb, p, s);
  else rc = dispatch_epi<64>(epilogue, ta, tb, p, s);
  return rc ? rc : gn_end(p, s);
}

extern "C" int dfot_conv3x3_bf16(const void* x, const void* w, void* out, int64_t ldc, int64_t n_img, int64_t H,
                                 int64_t W, int64_t Cin, int64_t Cout, int epilogue, const dfot_gemm_epilogue* epi,
                                 void* stream) {
  return conv_impl(x, w, out, ldc, n_img, H, W, Cin, Cout, 1, epilogue, epi, stream);
}

extern "C" int dfot_conv3d_causal_bf16(const void* x, const void* w, void* out, int64_t ldc, int64_t n_frames_out,
                                       int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t kt, int epilogue,
                                       const dfot_gemm_epilogue* epi, void* stream) {
  return conv_impl(x, w, out, ldc, n_frames_out, H, W, Cin, Cout, kt, epilogue, epi, stream);
}
