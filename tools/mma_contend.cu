// Probe: does concurrent epilogue-style traffic from other warps slow tcgen05.mma (M=128, N=256, K=16, SS)?
//   mma_contend <bg mask> <iters>
//   bg bits: 1 = tcgen05.ld x16 loop on the other accumulator half, 2 = st.shared.v4 loop into a scratch region,
//            4 = tcgen05.st x8 loop, 8 = FP32/MUFU arithmetic loop, 16 = mbarrier try_wait polling loop,
//            32 = one warp streams 32 KB bulk copies (TMA) from global memory into an unused shared-memory stage,
//            64 = operands filled with pseudo-random fp16 data instead of zeros
// 1 issuer warp + 16 background warps (lane quarter = warp % 4).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_contend tools/mma_contend.cu
#include <cstdio>
#include <cstdlib>
#include "../sdf-representation_b200/csrc/ptx.cuh"
using namespace sdf;

__global__ void __launch_bounds__(544, 1) contend_kernel(int bg, int iters, long long* cycles, float* sink, const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, pollbar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  for (uint32_t i = threadIdx.x; i < 196608 / 16; i += blockDim.x) {
    uint32_t v = 0;
    if (bg & 64) { v = (i * 2654435761u) ^ (i >> 3); v = (v & 0x03ff03ffu) | 0x38003800u; }  // fp16 values in [0.5, 1)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(v, v * 3u & 0x3bff3bffu | 0x38003800u, v ^ 0x00550055u, v);
  }
  __shared__ uint64_t tbar;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&pollbar, 1);
    mbar_init(&tbar, 1);
    mbar_fence_init();
    stop = 0;
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 256, false, false, false);
      const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 65536;
      uint64_t dbv[16], dav[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        dbv[k] = make_smem_desc(sB + k * 2 * (256 * 16), 256 * 16, 128);
        dav[k] = make_smem_desc(sA + k * 4096, 2048, 128);
      }
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) umma_f16_ss(tmem, dav[k], dbv[k], idesc, 1u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t1 = clock64();
      cycles[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (warp == 16 && (bg & 32)) {
    if (lane == 0) {
      uint32_t ph = 0;
      while (!stop) {
        mbar_arrive_expect_tx(&tbar, 32768);
        bulk_g2s(smem + 131072 + 32768, gsrc + (blockIdx.x & 63) * 32768, 32768, &tbar);
        mbar_wait(&tbar, ph & 1);
        ++ph;
      }
    }
  } else {
    const int q = (warp - 1) & 3;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + 256 + ((warp - 1) >> 2) * 32;
    const uint32_t sS = smem_u32(smem) + 131072 + 65536 - 32768 + (warp - 1) * 2048 + lane * 16;  // scratch beyond A, B
    float acc = lane * 0.001f;
    uint32_t r[16];
    long n = 0;
    while (!stop) {
      if (bg & 1) {
        tmem_ld16(taddr, r);
        tmem_ld_wait();
        acc += __uint_as_float(r[3]);
      }
      if (bg & 4) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(acc + j);
        tmem_st8(taddr + 16, w);
        tmem_st_wait();
      }
      if (bg & 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sS + j * 512), "r"(__float_as_uint(acc)) : "memory");
      }
      if (bg & 8) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float e;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc * -0.01f));
          acc = fmaf(e, 0.5f, acc * 0.999f);
        }
      }
      if (bg & 16) {
        for (int j = 0; j < 8; ++j) acc += mbar_try_wait(&pollbar, 0) ? 1.f : 0.f;
      }
      ++n;
    }
    if (acc == 12345.678f) sink[threadIdx.x] = acc + n;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  int bg = argc > 1 ? atoi(argv[1]) : 0;
  int iters = argc > 2 ? atoi(argv[2]) : 1000;
  long long* d;
  float* sink;
  cudaMalloc(&d, sizeof(long long) * 148);
  cudaMalloc(&sink, sizeof(float) * 1024);
  uint8_t* gsrc;
  cudaMalloc(&gsrc, 64 * 32768);
  cudaMemset(gsrc, 0x3c, 64 * 32768);
  cudaFuncSetAttribute(contend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  contend_kernel<<<148, 544, 196608>>>(bg, 10, d, sink, gsrc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  contend_kernel<<<148, 544, 196608>>>(bg, iters, d, sink, gsrc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mma_contend bg=%d: CUDA error %s\n", bg, cudaGetErrorString(e)); return 1; }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c0;
  cudaMemcpy(&c0, d, sizeof(c0), cudaMemcpyDeviceToHost);
  printf("mma_contend bg=%2d: %.1f cycles/MMA (cta0), %.3f ms\n", bg, c0 / (iters * 16.0), ms);
  return 0;
}
