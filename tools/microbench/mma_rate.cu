// Microbenchmark: issue rate / execution rate of tcgen05.mma (kind::f16, bf16) for the operand placements the
// InfoNCE kernels use.  All operands sit in shared memory (zero-filled); one thread issues `iters` batches of
// MMAs, commits, and waits.  Prints cycles per MMA and the implied fraction of the 8192 flop/clk/SM peak.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../rangeclip_b200/csrc mma_rate.cu -o mma_rate
#include "umma.cuh"
#include <cstdio>
#include <cstdlib>
using namespace rc::umma;

struct __align__(8) Bars { uint64_t done; uint32_t tmem_base, pad; };

// variant: 0 = pair S-type (M256 N256, A MN-major, B K-major), 1 = pair dX-type (M256 N128 K-major both),
//          2 = pair dX-type N256, 3 = single-CTA S-type (M128 N256), 4 = single-CTA dX-type (M128 N128),
//          5 = pair S-type with K-major A
template <int kPair>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int variant, int iters, int per_iter, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + 192 * 1024);
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (kPair) rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bars->done, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 1) { if (kPair) tmem_alloc_2sm<512>(&bars->tmem_base); else tmem_alloc<512>(&bars->tmem_base); }
  tc_fence_before();
  if (kPair) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  if (threadIdx.x == 0 && rank == 0) {
    const int M = kPair ? 256 : 128;
    int N = 256, a_mn = 0;
    if (variant == 0 || variant == 3) a_mn = 1;
    if (variant == 1 || variant == 4) N = 128;
    const uint32_t idesc = make_idesc_bf16(M, N, a_mn, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int j = 0; j < per_iter; ++j) {
        const int ch = (j >> 2) & 3, ks = j & 3;     // cycle over 4 chunks of 16 KB (A) / 16-32 KB (B)
        const uint64_t a = a_mn ? desc_mnmajor_sw128(a0 + ch * 16384 + ks * 2048, 8192) : desc_kmajor_sw128(a0 + ch * 16384 + ks * 32);
        const uint64_t b = desc_kmajor_sw128(b0 + ch * 16384 + ks * 32);
        const uint32_t d = tmem + ((variant == 1 || variant == 4) ? 256 + (j & 16 ? 128 : 0) : 0);
        if (kPair) mma_bf16_ss_2sm(d, a, b, idesc, 1); else mma_bf16_ss(d, a, b, idesc, 1);
      }
    }
    const long long t1 = clock64();
    if (kPair) mma_commit_2sm(&bars->done); else mma_commit(&bars->done);
    mbar_wait(&bars->done, 0, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (kPair && threadIdx.x == 0 && rank == 1) mbar_wait_cluster(&bars->done, 0, 2);
  tc_fence_before();
  if (kPair) cluster_sync(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); if (kPair) tmem_dealloc_2sm<512>(tmem); else tmem_dealloc<512>(tmem); }
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200;
  long long* out; cudaMalloc(&out, 16);
  const int smem = 192 * 1024 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"pair S (M256 N256, A MN-major)", "pair dX (M256 N128)", "pair dX (M256 N256)", "1cta S (M128 N256, A MN-major)",
                         "1cta dX (M128 N128)", "pair S (M256 N256, A K-major)"};
  for (int grid_mode = 0; grid_mode < 2; ++grid_mode)
    for (int v = 0; v < 6; ++v) {
      const bool pair = !(v == 3 || v == 4);
      const int grid = grid_mode == 0 ? (pair ? 2 : 1) : 148;
      cudaMemset(out, 0, 16);
      if (pair) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, mma_rate_kernel<1>, v, iters, 32, out);
      } else {
        mma_rate_kernel<0><<<grid, 128, smem>>>(v, iters, 32, out);
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const int N = (v == 1 || v == 4) ? 128 : 256;
      const double per = (double)h[1] / (iters * 32.0);
      const double ideal = 128.0 * N * 16 * 2 / 8192.0;    // per-SM cycles at 8192 flop/clk/SM
      printf("%-36s grid=%3d  issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f, %.0f%% of peak) %s\n", names[v], grid,
             (double)h[0] / (iters * 32.0), per, ideal, 100.0 * ideal / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
