// Micro-benchmark: how many bytes per clock can TMA land in one SM's shared memory, and does cluster multicast raise it?
//   mode 0: every CTA loads its own 16 KB tiles (64 x 128 bf16, 128B swizzle) through a 4-stage ring
//   mode 1: CTA pairs (cluster 2): each CTA issues HALF of every tile with a multicast mask to both CTAs — each SM still
//           receives 16 KB per tile, L2 serves half the bytes
//   each with tiles streamed from HBM (512 MB footprint) and with an L2-resident footprint (38 MB, every unit its own lines)
// If mode 1 > mode 0 the limit is L2 -> SM, if equal it is the shared-memory fill port.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/micro/tma_fill_bench.cu -o /tmp/tma_fill -lcuda && /tmp/tma_fill
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

constexpr int kStages = 6, kTile = 16384, kTilesPerStage = 2;   // 32 KB per stage, like the CTA-pair GEMM

__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar, uint32_t cta) {   // arrive on `bar` of CTA `cta` of the cluster
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {   // a protocol bug must not hang the GPU
  uint32_t done = 0;
  for (uint32_t spin = 0; !done && spin < (1u << 24); ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}

template <int MODE>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_half,
                                         int iters, int rows_total, long long* cycles_out, int resident) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t full = base + kStages * kTilesPerStage * kTile, empty = full + 8 * kStages;
  uint32_t rank = 0;
  if (MODE == 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, MODE == 1 ? 2 : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE == 1) { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
  long long t0 = clock64();
  bool ok = true;
  const int pair = MODE == 1 ? blockIdx.x >> 1 : blockIdx.x;
  const int n_units = MODE == 1 ? gridDim.x >> 1 : gridDim.x;
  if (threadIdx.x == 0) {            // producer (as in the GEMM: one thread, one stage = kTilesPerStage tiles)
    for (int i = 0; i < iters && ok; ++i) {
      const int s = i % kStages, round = i / kStages;
      if (round > 0) ok = mbar_wait_bounded(empty + 8 * s, (round - 1) & 1);
      mbar_expect_tx(full + 8 * s, kTilesPerStage * kTile);
      for (int t = 0; t < kTilesPerStage; ++t) {
        const long long j = (long long)i * kTilesPerStage + t;
        // resident: 16 tiles per unit, 38 MB in all — L2 hits after the first pass, every unit its own lines
        long long row = resident ? ((j % 16) * n_units + pair) * 128 : ((j * n_units + pair) * 128) % rows_total;
        const uint32_t dst = base + (s * kTilesPerStage + t) * kTile;
        if (MODE == 1) tma_load_2d_mc(dst + rank * (kTile / 2), &map_half, full + 8 * s, 0, (int)row + rank * 64, 3);
        else tma_load_2d(dst, &map_full, full + 8 * s, 0, (int)row);
      }
    }
  } else if (threadIdx.x == 32) {    // consumer: releases a stage in every CTA its loads write to
    for (int c = 0; c < iters && ok; ++c) {
      const int s = c % kStages, round = c / kStages;
      ok = mbar_wait_bounded(full + 8 * s, round & 1);
      if (MODE == 1) { mbar_arrive_cta(empty + 8 * s, 0); mbar_arrive_cta(empty + 8 * s, 1); }
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * s) : "memory");
    }
    if (!ok) cycles_out[blockIdx.x] = -1;
  }
  __syncthreads();
  if (MODE == 1) { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = ok ? clock64() - t0 : -1;   // (a consumer timeout shows as a producer timeout too)
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const long long rows = 1 << 22;                        // 4 Mi rows x 64 bf16 = 512 MB (larger than L2)
  void* buf;
  cudaMalloc(&buf, rows * 128);
  cudaMemset(buf, 0, rows * 128);
  CUtensorMap mf, mh;
  cuuint64_t gdim[2] = {64, (cuuint64_t)rows}, gstr[1] = {128};
  cuuint32_t estr[2] = {1, 1};
  cuuint32_t boxf[2] = {64, 128}, boxh[2] = {64, 64};
  enc(&mf, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, boxf, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&mh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, boxh, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  long long* cyc;
  cudaMallocManaged(&cyc, 148 * 8);
  const int smem = kStages * kTilesPerStage * kTile + 2048, iters = 4000;
  auto report = [&](const char* name, float ms) {
    long long mx = 0;
    for (int i = 0; i < 148; ++i) { if (cyc[i] < 0) printf("  (CTA %d timed out)\n", i); mx = cyc[i] > mx ? cyc[i] : mx; }
    printf("%-46s %8.1f us  %6.1f B/clk/SM  %6.2f TB/s landed chip-wide\n", name, ms * 1e3, (double)iters * kTilesPerStage * kTile / mx,
           148.0 * iters * kTilesPerStage * kTile / (ms * 1e-3) / 1e12);
  };
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    for (int resident = 0; resident < 2; ++resident) {
      const char* where = resident ? "L2-resident" : "HBM";
      char name[96];
      cudaEventRecord(a); k<0><<<148, 128, smem>>>(mf, mh, iters, (int)rows, cyc, resident); cudaEventRecord(b); cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b); snprintf(name, 96, "own tiles (%s)", where); if (rep) report(name, ms);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaEventRecord(a); cudaLaunchKernelEx(&cfg, k<1>, mf, mh, iters, (int)rows, cyc, resident); cudaEventRecord(b); cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b); snprintf(name, 96, "pairs, halves multicast to both (%s)", where); if (rep) report(name, ms);
    }
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
