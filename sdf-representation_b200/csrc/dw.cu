// Weight-gradient GEMM:  out[n][m] += sum_points  A[p][m] * B[p][n]   (K = points)
//
// dW_l = sum_p ( Zbar_l (x) h_l + q_l (x) hdot_l )   (the reference gets it from autograd's
// mm(grad^T, input) nodes, executor/executor.py:176).  Both operands are the spilled tile images of chain.cu,
// consumed as MN-major no-swizzle UMMA operands (K = the 128 rows of a tile image), loaded with two 1-D bulk
// copies per tile and term.  A persistent CTA owns one 128 x N accumulator in TMEM per work item (two items in
// flight: the epilogue of one overlaps the MMAs of the next) and reduces it into the fp32 gradient accumulator
// with red.global.add at the end of the item.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane = output feature m), warp 4 producer, warp 5 MMA issuer.
#include <cstdio>
#include "ptx.cuh"
#include "sdf_internal.h"

namespace sdf {

constexpr int kDwStages = 2;
constexpr int kDwThreads = 192;
constexpr uint32_t kDwABytes = 16 * kGranuleBytes;   // 128 features x 128 points
constexpr uint32_t kDwBBytes = 32 * kGranuleBytes;   // 256 features x 128 points
constexpr uint32_t kDwStageBytes = kDwABytes + kDwBBytes;
constexpr uint32_t kDwSmem = kDwStages * kDwStageBytes + 128;

template <int DT>
__global__ void __launch_bounds__(kDwThreads, 1) dw_kernel(const DwItem* __restrict__ items, int n_items, int parity) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kDwStageBytes);
  uint64_t* full = bars;                    // [2]
  uint64_t* empty = bars + 2;               // [2]
  uint64_t* acc_full = bars + 4;            // [2]
  uint64_t* acc_empty = bars + 6;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, round = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const DwItem d = items[it];
        const uint32_t b_bytes = (d.n / 8) * kGranuleBytes;
        const int nterms = d.A2 ? 2 : 1;
        for (uint32_t tile = d.tile0; tile < d.tile1; ++tile) {
          for (int term = 0; term < nterms; ++term) {
            const uint8_t* a = (term ? d.A2 : d.A1) + static_cast<size_t>(tile) * d.a_tstride + d.a_off;
            const uint8_t* b = (term ? d.B2 : d.B1) + static_cast<size_t>(tile) * d.b_tstride + d.b_off;
            mbar_wait(&empty[stage], (round & 1) ^ 1);
            mbar_arrive_expect_tx(&full[stage], kDwABytes + b_bytes);
            uint8_t* dst = smem + stage * kDwStageBytes;
            bulk_g2s(dst, a, kDwABytes, &full[stage]);
            bulk_g2s(dst + kDwABytes, b, b_bytes, &full[stage]);
            if (++stage == kDwStages) { stage = 0; ++round; }
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      uint32_t stage = 0, round = 0, item_cnt = 0;
      const uint32_t lbo = 128u;            // K-group (8 image rows) stride
      const uint32_t sbo = kGranuleBytes;   // MN-block (8 features) stride
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_cnt) {
        const DwItem d = items[it];
        const uint32_t as = item_cnt & 1;
        mbar_wait(&acc_empty[as], ((item_cnt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(128, static_cast<int>(d.n), DT == DT_BF16, true, true);
        const int nterms = d.A2 ? 2 : 1;
        uint32_t first = 1;
        for (uint32_t tile = d.tile0; tile < d.tile1; ++tile) {
          for (int term = 0; term < nterms; ++term) {
            mbar_wait(&full[stage], round & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * kDwStageBytes);
            const uint32_t b_addr = a_addr + kDwABytes;
            if (!parity) {
#pragma unroll
              for (int k = 0; k < kTileRows / 16; ++k) {
                const uint64_t da = make_smem_desc(a_addr + k * 256, lbo, sbo);
                const uint64_t db = make_smem_desc(b_addr + k * 256, lbo, sbo);
                umma_f16_ss(tmem_base + as * 256, da, db, idesc, first ? 0u : 1u);
                first = 0;
              }
            } else {
              // image rows 0-63 = hi parts, 64-127 = lo parts of the same 64 points: hi.hi + lo.hi + hi.lo
              // (lo.lo is below fp32 resolution)
#pragma unroll
              for (int k = 0; k < kParityPoints / 16; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  const uint32_t ka = k + (c == 1 ? 4 : 0), kb = k + (c == 2 ? 4 : 0);
                  const uint64_t da = make_smem_desc(a_addr + ka * 256, lbo, sbo);
                  const uint64_t db = make_smem_desc(b_addr + kb * 256, lbo, sbo);
                  umma_f16_ss(tmem_base + as * 256, da, db, idesc, first ? 0u : 1u);
                  first = 0;
                }
              }
            }
            umma_commit(&empty[stage]);
            if (++stage == kDwStages) { stage = 0; ++round; }
          }
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    uint32_t item_cnt = 0;
    const int m = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_cnt) {
      const DwItem d = items[it];
      const uint32_t as = item_cnt & 1;
      mbar_wait(&acc_full[as], (item_cnt >> 1) & 1);
      tc_fence_after();
      if (d.tile1 > d.tile0) {
        const float im = d.inv_mult ? __ldg(d.inv_mult) : 1.f;
        for (uint32_t c0 = 0; c0 < d.n; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + as * 256 + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(d.out + static_cast<size_t>(c0 + j) * d.out_ld + m, __uint_as_float(r[j]) * im);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem_base);
}

cudaError_t launch_dw(const DwItem* items_dev, int n_items, int dtype, int parity, int num_sms, cudaStream_t stream) {
  static bool attr_set[2] = {false, false};
  auto kfn = dtype == DT_BF16 ? dw_kernel<DT_BF16> : dw_kernel<DT_FP16>;
  if (!attr_set[dtype]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kDwSmem));
    if (e != cudaSuccess) return e;
    attr_set[dtype] = true;
  }
  if (n_items <= 0) return cudaSuccess;
  const int grid = n_items < num_sms ? n_items : num_sms;
  kfn<<<grid, kDwThreads, kDwSmem, stream>>>(items_dev, n_items, parity);
  return cudaGetLastError();
}

}  // namespace sdf
