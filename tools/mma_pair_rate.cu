// Probe: issue rate of tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16, SS) from a CTA pair, operands zero-filled.
//   mma_pair_rate <iters>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_pair_rate tools/mma_pair_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../sdf-representation_b200/csrc/ptx.cuh"
using namespace sdf;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (uint32_t i = threadIdx.x; i < 196608 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc2<512>(&slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, 256, false, false, false);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 65536;
    uint64_t dbv[16], dav[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      dbv[k] = make_smem_desc(sB + k * 2 * (128 * 16), 128 * 16, 128);   // this CTA's 128 rows of B
      dav[k] = make_smem_desc(sA + k * 4096, 2048, 128);
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) umma_f16_ss2(tmem, dav[k], dbv[k], idesc, 1u);
    }
    umma_commit2(&bar, 3);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x >> 1] = t1 - t0;
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2<512>(tmem);
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 1000;
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 74);
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  pair_kernel<<<148, 128, 196608>>>(10, d);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  pair_kernel<<<148, 128, 196608>>>(iters, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mma_pair_rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c0;
  cudaMemcpy(&c0, d, sizeof(c0), cudaMemcpyDeviceToHost);
  const double n = iters * 16.0;
  printf("mma_pair_rate: %.1f cycles per M=256,N=256,K=16 pair MMA (pair 0), %.3f ms, %.1f TFLOP/s\n", c0 / n, ms,
         n * 2.0 * 256 * 256 * 16 * 74 / (ms * 1e-3) / 1e12);
  return 0;
}
