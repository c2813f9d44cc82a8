// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1, M=128, K=16) for the operand sources and N
// shapes the InfoNCE backward kernel uses.  One CTA per SM, one issuing warp, back-to-back MMAs, clock64 around
// issue + commit + wait.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../clip-for-dl_b200/csrc/common.cuh"
using namespace b200;

// mode 0: SS (A smem K-major, B smem K-major); 1: TS (A tmem); 2: SS with B MN-major (dX shape); 3: SS, A MN-major
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, MODE == 3, MODE == 2);
    // A: 8 chunks of [128 x 64] (16 KB each) at smem+0 ; B: chunks of [N x 64] at smem + 128 KB (K-major) ;
    // MN-major B ([16 K-rows] x N): 64-element groups 2048 B apart like the backward kernel
    const uint32_t a_lo = desc_lo(smem_u32(smem), 16);
    const uint32_t b_lo = (MODE == 2) ? desc_lo(smem_u32(smem + 64 * 1024), 4096) : desc_lo(smem_u32(smem + 64 * 1024), 16);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (MODE == 1) mma_ts_lo(tmem + 256, tmem + c * 32 + j * 8, b_lo + c * ((N * 128) >> 4) + 2 * j, idesc, true);
              else if (MODE == 2) {                     // the dX MMA of nce_bwd4: G tile (A, K-major) x Y in-half tile (B, MN-major)
                const int t = (c * 4 + j) >> 1, jj = j & 1;
                mma_ss_lo(tmem, a_lo + (t & 1) * 1024 + ((t >> 1) & 1) * 4 + jj * 2, b_lo + (t & 7) * 1024 + jj * 128, idesc, true);
              }
              else mma_ss_lo(tmem + 256, a_lo + c * (16384 >> 4) + 2 * j, b_lo + c * ((N * 128) >> 4) + 2 * j, idesc, true);
            }
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// Mixed issue as in nce_bwd4: warp 1 issues the S MMAs of an own tile (28 TS + 4 SS, N = 32), warp 2 the dX MMAs of two
// tiles (4 x N = 256, B MN-major), concurrently, into different accumulators.  Reports cycles per tile PAIR.
__global__ void __launch_bounds__(128, 1) mixed_kernel(long long* out, int iters, int which) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_lo = desc_lo(smem_u32(smem), 16);
  long long t0 = clock64();
  if (warp == 1 && (which & 1)) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 32, false, false);
    const uint32_t b_lo = desc_lo(smem_u32(smem + 64 * 1024), 16);
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 7; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_ts_lo(tmem + 256, tmem + 288 + c * 32 + j * 8, b_lo + c * (4096 >> 4) + 2 * j, idesc, true);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_ss_lo(tmem + 256, a_lo + 2 * j, b_lo + 7 * (4096 >> 4) + 2 * j, idesc, true);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], 0);
  }
  if (warp == 2 && (which & 2)) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, false, true);
    const uint32_t g_lo = desc_lo(smem_u32(smem + 16 * 1024), 16);
    const uint32_t b_lo = desc_lo(smem_u32(smem + 64 * 1024), 4096);
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
            mma_ss_lo_ab(tmem, g_lo + ((it * 2 + t) & 3) * 512 + jj * 2, DESC_HI_SW64, b_lo + ((it * 2 + t) & 7) * 1024 + jj * 128, DESC_HI_SW128, idesc, true);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar[1]);
    __syncwarp();
    mbar_wait(&bar[1], 0);
  }
  long long t1 = clock64();
  __shared__ long long tt[4];
  if (lane == 0) tt[warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = max(tt[1], tt[2]);
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void run_mixed(const char* name, long long* d_out, int iters, int which) {
  cudaFuncSetAttribute(mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mixed_kernel<<<148, 128, 200 * 1024>>>(d_out, iters, which);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %8.1f cyc per tile pair   %s\n", name, double(cyc) / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int MODE, int N>
static void run(const char* name, long long* d_out, int iters) {
  auto k = rate_kernel<MODE, N>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<148, 128, 200 * 1024>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  const double per = double(cyc) / (double(iters) * 16);
  printf("%-28s N=%3d: %8.1f cyc/MMA  (floor N/2 = %d)  -> %.0f%% of tensor peak   %s\n", name, N, per, N / 2, 100.0 * (N / 2) / per,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  const int iters = 2000;
  run<0, 32>("SS  A,B K-major", d_out, iters);
  run<1, 32>("TS  A tmem", d_out, iters);
  run<0, 64>("SS  A,B K-major", d_out, iters);
  run<1, 64>("TS  A tmem", d_out, iters);
  run<0, 128>("SS  A,B K-major", d_out, iters);
  run<1, 128>("TS  A tmem", d_out, iters);
  run<0, 256>("SS  A,B K-major", d_out, iters);
  run<1, 256>("TS  A tmem", d_out, iters);
  run<2, 256>("SS  B MN-major (dX)", d_out, iters);
  run<3, 128>("SS  A MN-major", d_out, iters);
  run<3, 256>("SS  A MN-major", d_out, iters);
  run_mixed("S MMAs only (28 TS + 4 SS, N=32)", d_out, iters, 1);
  run_mixed("dX MMAs only (4 x N=256)", d_out, iters, 2);
  run_mixed("both warps concurrently", d_out, iters, 3);
  return 0;
}
