// Probe: the weight ring of the layer-chain kernel in isolation.  One producer thread streams 32 KB (or <chunk>) bulk
// copies from an L2-resident global buffer into an <stages>-deep shared-memory ring, one issuer thread runs <k16> MMAs
// (M = 128, N = 256, K = 16, SS, A from a fixed 128 KB image) per chunk and commits the stage.  Reports cycles per MMA.
//   ring_probe <stages> <chunk_kb> <chunks> [mode]   mode bits: 1 = producer arrives without copying, 2 = issuer does not
//   wait on the full barriers (and no producer), 4 = no per-chunk commit, 8 = descriptors precomputed per stage,
//   16 = commit of chunk c-1 issued after the first MMA of chunk c (hidden behind a queued MMA)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ring_probe tools/ring_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../sdf-representation_b200/csrc/ptx.cuh"
using namespace sdf;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128, 1) ring_kernel(const uint8_t* w, int wbytes, int stages, int chunk_bytes, int chunks, int no_copy,
                                                      long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8], done;
  __shared__ uint32_t slot;
  uint8_t* sA = smem;
  uint8_t* sW = smem + 131072;
  for (uint32_t i = threadIdx.x; i < 131072 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x38003800u, 0x39003900u, 0x3a003a00u, 0x3b003b00u);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0 && !(no_copy & 2)) {
    uint32_t stage = 0, round = 0;
    size_t off = (static_cast<size_t>(blockIdx.x) * 32768) % wbytes;
    for (int c = 0; c < chunks; ++c) {
      if (!(no_copy & 4)) mbar_wait(&empty[stage], (round & 1) ^ 1);
      if (no_copy & 1) {
        mbar_arrive(&full[stage]);
      } else {
        mbar_arrive_expect_tx(&full[stage], chunk_bytes);
        bulk_g2s(sW + stage * chunk_bytes, w + off, chunk_bytes, &full[stage]);
      }
      off += chunk_bytes;
      if (off + chunk_bytes > static_cast<size_t>(wbytes)) off = 0;
      if (++stage == static_cast<uint32_t>(stages)) { stage = 0; ++round; }
    }
  } else if (warp == 0 && (no_copy & 32)) {
    // warp-uniform issue loop: all lanes walk the loop (scalars stay in uniform registers), one elected lane issues
    const uint32_t idesc = make_idesc_f16(128, 256, false, false, false);
    const int k16 = chunk_bytes / 8192;
    const uint32_t b_lbo = 256 * 16;
    const uint64_t db_step = (2 * b_lbo) >> 4;
    const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
    uint32_t stage = 0, round = 0;
    long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      mbar_wait(&full[stage], round & 1);
      tc_fence_after();
      const uint64_t da0 = make_smem_desc(sA_u + (c & 7) * 16384, 2048, 128);
      const uint64_t db0 = make_smem_desc(sW_u + stage * chunk_bytes, b_lbo, 128);
      const uint32_t td = tmem + ((c / 8) & 1) * 256;
      if (elect_one()) {
        if (k16 == 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(td, da0 + static_cast<uint64_t>(k) * 256, db0 + static_cast<uint64_t>(k) * db_step, idesc, 1u);
        } else {
          for (int k = 0; k < k16; ++k) umma_f16_ss(td, da0 + static_cast<uint64_t>(k & 3) * 256, db0 + static_cast<uint64_t>(k) * db_step, idesc, 1u);
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == static_cast<uint32_t>(stages)) { stage = 0; ++round; }
    }
    if (elect_one()) {
      umma_commit(&done);
      mbar_wait(&done, 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (warp == 0 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, 256, false, false, false);
    const int k16 = chunk_bytes / 8192;  // 256 rows x 16 K x 2 B per K=16 slice
    const uint32_t b_lbo = 256 * 16;
    uint32_t stage = 0, round = 0;
    uint64_t dbs[8], das[8];
    for (int i = 0; i < 8; ++i) {
      dbs[i] = make_smem_desc(smem_u32(sW) + i * chunk_bytes, b_lbo, 128);
      das[i] = make_smem_desc(smem_u32(sA) + i * 16384, 2048, 128);
    }
    const uint64_t db_step = (2 * b_lbo) >> 4;
    long long t0 = clock64();
    int prev_stage = -1;
    for (int c = 0; c < chunks; ++c) {
      if (!(no_copy & 2)) {
        mbar_wait(&full[stage], round & 1);
        tc_fence_after();
      }
      uint64_t da0, db0;
      if (no_copy & 8) {
        da0 = das[c & 7];
        db0 = dbs[stage];
      } else {
        da0 = make_smem_desc(smem_u32(sA) + (c & 7) * 16384, 2048, 128);
        db0 = make_smem_desc(smem_u32(sW) + stage * chunk_bytes, b_lbo, 128);
      }
      const uint32_t td = tmem + ((c / 8) & 1) * 256;
      umma_f16_ss(td, da0, db0, idesc, 1u);
      if ((no_copy & 16) && prev_stage >= 0) umma_commit(&empty[prev_stage]);
      for (int k = 1; k < k16; ++k) umma_f16_ss(td, da0 + static_cast<uint64_t>(k & 3) * 256, db0 + static_cast<uint64_t>(k) * db_step, idesc, 1u);
      if (!(no_copy & 4) && !(no_copy & 16)) umma_commit(&empty[stage]);
      prev_stage = stage;
      if (++stage == static_cast<uint32_t>(stages)) { stage = 0; ++round; }
    }
    if ((no_copy & 16) && prev_stage >= 0) umma_commit(&empty[prev_stage]);
    umma_commit(&done);
    mbar_wait(&done, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
  int stages = argc > 1 ? atoi(argv[1]) : 3;
  int chunk_kb = argc > 2 ? atoi(argv[2]) : 32;
  int chunks = argc > 3 ? atoi(argv[3]) : 4000;
  int no_copy = argc > 4 ? atoi(argv[4]) : 0;
  const int wbytes = 3670016;  // ~ the packed weight stream of the 8x512 net (L2 resident)
  uint8_t* w;
  long long* d;
  cudaMalloc(&w, wbytes + 65536);
  cudaMemset(w, 0x3a, wbytes + 65536);
  cudaMalloc(&d, sizeof(long long) * 148);
  const int smem = 131072 + stages * chunk_kb * 1024;
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ring_kernel<<<148, 128, smem>>>(w, wbytes, stages, chunk_kb * 1024, 64, no_copy, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ring_probe: %s\n", cudaGetErrorString(e)); return 1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  ring_kernel<<<148, 128, smem>>>(w, wbytes, stages, chunk_kb * 1024, chunks, no_copy, d);
  cudaEventRecord(e1);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ring_probe: %s\n", cudaGetErrorString(e)); return 1; }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c0;
  cudaMemcpy(&c0, d, sizeof(c0), cudaMemcpyDeviceToHost);
  const double n_mma = static_cast<double>(chunks) * (chunk_kb / 8);
  printf("ring_probe stages=%d chunk=%dKB no_copy=%d: %.1f cycles/MMA, %.1f B/clk/SM weight stream, %.3f ms, L2->SM %.2f TB/s\n", stages, chunk_kb, no_copy,
         c0 / n_mma, static_cast<double>(chunks) * chunk_kb * 1024 / c0, ms, 148.0 * chunks * chunk_kb * 1024 / (ms * 1e-3) / 1e12);
  return 0;
}
