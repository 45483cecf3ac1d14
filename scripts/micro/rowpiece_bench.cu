// Micro-benchmark: HBM bandwidth of "out = in + 1" over an fp32 [M, N] matrix when a CTA owns a 128-row x BN-column tile
// and its 8 warps walk it the way the GEMM epilogue does:
//   mode 0: warp = 32 rows x 32 columns at a time, one instruction = 4 rows x 128 B   (today's epilogue row pass)
//   mode 1: warp = 32 rows x 128 columns at a time, one instruction = 1 row x 512 B
//   mode 2: warp = 32 rows x 64 columns at a time, one instruction = 2 rows x 256 B
// Build + run on the GPU box:  nvcc -O3 -arch=sm_100a scripts/micro/rowpiece_bench.cu -o /tmp/rowpiece && /tmp/rowpiece
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* __restrict__ out, int M, int N, int BN) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2;
  const int num_m = M / 128, num_n = N / BN;
  for (int tile = blockIdx.x; tile < num_m * num_n; tile += gridDim.x) {
    const int m_blk = tile % num_m, n_blk = tile / num_m;
    const int m0 = m_blk * 128 + q * 32, nb = n_blk * BN;
    if (MODE == 0) {
      for (int c = half; c < BN / 32; c += 2) {
        const int col = nb + c * 32 + (lane & 7) * 4, rsub = lane >> 3;
        float4 v[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) v[it] = __ldg(reinterpret_cast<const float4*>(in + (size_t)(m0 + 4 * it + rsub) * N + col));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          v[it].x += 1.f; v[it].y += 1.f; v[it].z += 1.f; v[it].w += 1.f;
          *reinterpret_cast<float4*>(out + (size_t)(m0 + 4 * it + rsub) * N + col) = v[it];
        }
      }
    } else if (MODE == 1) {
      for (int c = half; c < BN / 128; c += 2) {
        const int col = nb + c * 128 + lane * 4;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
          float4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) v[it] = __ldg(reinterpret_cast<const float4*>(in + (size_t)(m0 + r0 + it) * N + col));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            v[it].x += 1.f; v[it].y += 1.f; v[it].z += 1.f; v[it].w += 1.f;
            *reinterpret_cast<float4*>(out + (size_t)(m0 + r0 + it) * N + col) = v[it];
          }
        }
      }
    } else {
      for (int c = half; c < BN / 64; c += 2) {
        const int col = nb + c * 64 + (lane & 15) * 4, rsub = lane >> 4;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 16) {
          float4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) v[it] = __ldg(reinterpret_cast<const float4*>(in + (size_t)(m0 + r0 + 2 * it + rsub) * N + col));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            v[it].x += 1.f; v[it].y += 1.f; v[it].z += 1.f; v[it].w += 1.f;
            *reinterpret_cast<float4*>(out + (size_t)(m0 + r0 + 2 * it + rsub) * N + col) = v[it];
          }
        }
      }
    }
  }
}

template <int MODE> void run(const float* in, float* out, int M, int N, int BN, const char* name) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) k<MODE><<<148, 256>>>(in, out, M, N, BN);
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) k<MODE><<<148, 256>>>(in, out, M, N, BN);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("M=%d N=%d BN=%d %-28s %7.1f us  %6.0f GB/s\n", M, N, BN, name, ms * 50.f, 2.0 * M * N * 4 / (ms / 20 * 1e-3) / 1e9);
}

int main() {
  const int shapes[][3] = {{65536, 576, 192}, {16384, 1152, 256}, {65536, 2304, 256}, {1048576, 128, 128}};
  float *in, *out;
  cudaMalloc(&in, (size_t)1 << 30); cudaMalloc(&out, (size_t)1 << 30);
  cudaMemset(in, 0, (size_t)1 << 30);
  for (auto& s : shapes) {
    run<0>(in, out, s[0], s[1], s[2], "4 rows x 128 B / instr");
    if (s[2] % 64 == 0) run<2>(in, out, s[0], s[1], s[2], "2 rows x 256 B / instr");
    if (s[2] % 128 == 0) run<1>(in, out, s[0], s[1], s[2], "1 row x 512 B / instr");
  }
  return 0;
}
