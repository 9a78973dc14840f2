// Stand-alone tcgen05.mma issue-rate probe: how fast does ONE CTA per SM retire M=128 x N x K=16 kind::f16
// instructions (cta_group::1), with A from shared memory (SS) or from tensor memory (TS), and with B re-read from
// the same shared-memory bytes.  Numbers decide the tile shape of the layer-chain kernel.
//   mma_rate <N> <ts> <iters> <grid>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../sdf-representation_b200/csrc/ptx.cuh"
using namespace sdf;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int ts, int iters, long long* cycles, int pre) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (uint32_t i = threadIdx.x; i < 196608 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, false, false, false);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 65536;
    // descriptors precomputed (as the production issue loop does): the timed loop is MMA issue only
    uint64_t dbv[16], dav[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      dbv[k] = make_smem_desc(sB + k * 2 * (N * 16), N * 16, 128);
      dav[k] = make_smem_desc(sA + k * 4096, 2048, 128);
    }
    long long t0 = clock64();
    if (pre) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          // pre == 2: alternate between two accumulators (independent dependency chains)
          const uint32_t d = tmem + ((pre == 2 && (k & 1)) ? static_cast<uint32_t>(N) : 0u);
          if (ts) umma_f16_ts(d, tmem + 2 * N + k * 8, dbv[k], idesc, 1u);
          else umma_f16_ss(d, dav[k], dbv[k], idesc, 1u);
        }
      }
    } else
    for (int it = 0; it < iters; ++it) {
      // K = 256 per "chunk": 16 instructions walking distinct operand bytes like the real kernel does
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t db = make_smem_desc(sB + k * 2 * (N * 16), N * 16, 128);
        if (ts) {
          umma_f16_ts(tmem, tmem + 256 + k * 8, db, idesc, 1u);
        } else {
          const uint64_t da = make_smem_desc(sA + k * 4096, 2048, 128);
          umma_f16_ss(tmem, da, db, idesc, 1u);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 256;
  int ts = argc > 2 ? atoi(argv[2]) : 0;
  int iters = argc > 3 ? atoi(argv[3]) : 2000;
  int grid = argc > 4 ? atoi(argv[4]) : 148;
  int pre = argc > 5 ? atoi(argv[5]) : 1;
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  rate_kernel<<<grid, 128, 196608>>>(N, ts, 10, d, pre);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  rate_kernel<<<grid, 128, 196608>>>(N, ts, iters, d, pre);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mma_rate N=%d ts=%d: CUDA error %s\n", N, ts, cudaGetErrorString(e));
    return 1;
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c0;
  cudaMemcpy(&c0, d, sizeof(c0), cudaMemcpyDeviceToHost);
  const double n_mma = static_cast<double>(iters) * 16;
  const double flop = n_mma * 2.0 * 128 * N * 16 * grid;
  printf("mma_rate pre=%d N=%d %s grid=%d: %.1f cycles/MMA (cta0), %.3f ms, %.1f TFLOP/s\n", pre, N, ts ? "TS" : "SS", grid,
         c0 / n_mma, ms, flop / (ms * 1e-3) / 1e12);
  return 0;
}
