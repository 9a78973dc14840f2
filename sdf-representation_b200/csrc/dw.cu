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

// Split-K partial sums meet in the fp32 gradient accumulator: a reduction without return value on a pointer that is known
// to be global memory (atomicAdd on the generic pointer of a DwItem compiles to ATOM + a shared-memory CAS-spin fallback
// behind a QSPC branch; this is one RED.E.ADD.F32 per element).
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}

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
          for (int j = 0; j < 16; ++j) {
            float* o = d.out + static_cast<size_t>(c0 + j) * d.out_ld + m;
            if (d.store) *o = __uint_as_float(r[j]) * im;
            else red_add_f32(o, __uint_as_float(r[j]) * im);
          }
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

// deterministic mode, second pass: acc[i] += partial[0][i] + partial[1][i] + ... in split order
__global__ void reduce_partials_kernel(float* __restrict__ acc, const float* __restrict__ partial, long n, int slots) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < slots; ++k) s += partial[static_cast<long>(k) * n + i];
    acc[i] += s;
  }
}
cudaError_t launch_reduce_partials(float* acc, const float* partial, long n, int slots, cudaStream_t stream) {
  reduce_partials_kernel<<<1184, 256, 0, stream>>>(acc, partial, n, slots);
  return cudaGetLastError();
}

cudaError_t launch_dw(const DwItem* items_dev, int n_items, int dtype, int parity, int num_sms, cudaStream_t stream) {
  static bool attr_set[2][kMaxDevices] = {};  // per device: the shared-memory opt-in is a per-context attribute
  auto kfn = dtype == DT_BF16 ? dw_kernel<DT_BF16> : dw_kernel<DT_FP16>;
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dtype][dev]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kDwSmem));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) attr_set[dtype][dev] = true;
  }
  if (n_items <= 0) return cudaSuccess;
  const int grid = n_items < num_sms ? n_items : num_sms;
  kfn<<<grid, kDwThreads, kDwSmem, stream>>>(items_dev, n_items, parity);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cluster of 2, tcgen05 cta_group::2): one pair owns a 256 x 256 block of dW.
// Each CTA loads the 128-feature half of the A block that its TMEM lanes accumulate and one 128-feature half of the
// B block (the tensor cores of the pair read both halves), i.e. 64 KB per tile and term instead of 96 KB for a
// 128 x 256 block: half the L2 -> SM bytes and half the shared-memory traffic per FLOP of dw_kernel.
// Warp roles per CTA (192 threads): warps 0-3 epilogue, warp 4 producer, warp 5 = MMA issuer in the leader CTA
// (rank 0) and "my half has landed" relay in the peer CTA.
// ------------------------------------------------------------------------------------------------
constexpr int kDw2Stages = 3;
constexpr uint32_t kDw2Half = 16 * kGranuleBytes;        // 128 features x 128 points
constexpr uint32_t kDw2XBytes = kGranuleBytes;            // this CTA's 8 columns of a 16-column block image
constexpr uint32_t kDw2StageBytes = 2 * kDw2Half + kDw2XBytes;  // A half + B half + block-image granule
constexpr uint32_t kDw2Smem = kDw2Stages * kDw2StageBytes + 256;
#ifndef SDF_DW_SYNC_TILES
#define SDF_DW_SYNC_TILES 8
#endif
constexpr int kDw2SyncTiles = SDF_DW_SYNC_TILES;  // tiles between two lock-step points of a group of pairs (power of two)

template <int DT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDwThreads, 1)
dw2_kernel(const DwItem* __restrict__ items, int n_items, int parity, int* __restrict__ sync_ctr, int group, int n_pairs_eff) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDw2Stages * kDw2StageBytes);
  uint64_t* full = bars;                         // [3] local: this CTA's halves have landed
  uint64_t* empty = bars + kDw2Stages;           // [3] both CTAs: the pair's MMAs on this stage are done
  uint64_t* peer_full = bars + 2 * kDw2Stages;   // [3] leader only: the peer's halves have landed
  uint64_t* acc_full = bars + 3 * kDw2Stages;    // [2] both CTAs
  uint64_t* acc_empty = acc_full + 2;            // [2] leader only: both epilogues drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // Work items are dealt to the first n_pairs_eff pairs only (a multiple of the group size): the `group` consecutive
  // items that share operand tiles (the four 256 x 256 blocks of one layer and tile range) are then always processed by
  // `group` neighbouring pairs in the same round, and those pairs keep in step (sync_ctr, see the producer) so that the
  // second read of every half tile is an L2 hit instead of a DRAM read.
  const int pair = (static_cast<int>(blockIdx.x >> 1) < n_pairs_eff) ? static_cast<int>(blockIdx.x >> 1) : n_items, n_pairs = n_pairs_eff;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kDw2Stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&peer_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 256);
    }
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc2<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barrier inits and the TMEM allocation of both CTAs are in place
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, round = 0;
      for (int it = pair; it < n_items; it += n_pairs) {
        const DwItem d = items[it];
        const int nterms = d.A2 ? 2 : 1;
        for (uint32_t tile = d.tile0; tile < d.tile1; ++tile) {
          for (int term = 0; term < nterms; ++term) {
            const uint8_t* a = (term ? d.A2 : d.A1) + static_cast<size_t>(tile) * d.a_tstride + d.a_off + rank * kDw2Half;
            const uint8_t* b = (term ? d.B2 : d.B1) + static_cast<size_t>(tile) * d.b_tstride + d.b_off + rank * kDw2Half;
            const bool with_x = d.X != nullptr && term == 0;
            if (group > 1 && rank == 0 && term == 0 && ((tile - d.tile0) & (kDw2SyncTiles - 1)) == 0) {
              // lock-step with the other pairs of the group every kDw2SyncTiles tiles: one arrival per pair, then wait
              // until all `group` pairs have arrived at this point of this item
              int* ctr = sync_ctr + it / group;
              const int target = group * (static_cast<int>(tile - d.tile0) / kDw2SyncTiles + 1);
              atomicAdd(ctr, 1);
              uint32_t spins = 0;
              while (*reinterpret_cast<volatile int*>(ctr) < target) {
                if (++spins > SDF_SPIN_LIMIT) { printf("sdf: dw2 group sync timeout (pair %d item %d)\n", pair, it); __trap(); }
              }
            }
            mbar_wait(&empty[stage], (round & 1) ^ 1);
            mbar_arrive_expect_tx(&full[stage], 2 * kDw2Half + (with_x ? kDw2XBytes : 0u));
            uint8_t* dst = smem + stage * kDw2StageBytes;
            bulk_g2s(dst, a, kDw2Half, &full[stage]);
            bulk_g2s(dst + kDw2Half, b, kDw2Half, &full[stage]);
            if (with_x) bulk_g2s(dst + 2 * kDw2Half, d.X + static_cast<size_t>(tile) * d.x_tstride + rank * kDw2XBytes, kDw2XBytes, &full[stage]);
            if (++stage == kDw2Stages) { stage = 0; ++round; }
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      uint32_t stage = 0, round = 0, item_cnt = 0;
      if (rank != 0) {
        // relay: tell the leader when this CTA's halves of a stage are in shared memory
        for (int it = pair; it < n_items; it += n_pairs) {
          const DwItem d = items[it];
          const uint32_t uses = (d.tile1 - d.tile0) * (d.A2 ? 2u : 1u);
          for (uint32_t u = 0; u < uses; ++u) {
            mbar_wait(&full[stage], round & 1);
            mbar_arrive_cluster(&peer_full[stage], 0);
            if (++stage == kDw2Stages) { stage = 0; ++round; }
          }
        }
      } else {
        const uint32_t lbo = 128u;            // K-group (8 image rows) stride
        const uint32_t sbo = kGranuleBytes;   // MN-block (8 features) stride
        const uint32_t idesc = make_idesc_f16(256, 256, DT == DT_BF16, true, true);
        const uint32_t idesc_x = make_idesc_f16(256, 16, DT == DT_BF16, true, true);
        uint32_t uses[2] = {0, 0};  // acquisitions of each accumulator slot so far (barrier phases)
        for (int it = pair; it < n_items; it += n_pairs, ++item_cnt) {
          const DwItem d = items[it];
          const uint32_t as = item_cnt & 1;
          // an item with a block-image product keeps its 16 extra columns in the other slot: it owns both
          mbar_wait_cluster(&acc_empty[as], (uses[as] & 1) ^ 1);
          ++uses[as];
          if (d.X) {
            mbar_wait_cluster(&acc_empty[as ^ 1], (uses[as ^ 1] & 1) ^ 1);
            ++uses[as ^ 1];
          }
          tc_fence_after();
          const int nterms = d.A2 ? 2 : 1;
          uint32_t first = 1, first_x = 1;
          for (uint32_t tile = d.tile0; tile < d.tile1; ++tile) {
            for (int term = 0; term < nterms; ++term) {
              mbar_wait(&full[stage], round & 1);
              mbar_wait_cluster(&peer_full[stage], round & 1);
              tc_fence_after();
              const uint32_t a_addr = smem_u32(smem + stage * kDw2StageBytes);
              const uint32_t b_addr = a_addr + kDw2Half;
              const uint64_t da0 = make_smem_desc(a_addr, lbo, sbo), db0 = make_smem_desc(b_addr, lbo, sbo);
              if (!parity) {
#pragma unroll
                for (int k = 0; k < kTileRows / 16; ++k) {
                  umma_f16_ss2(tmem_base + as * 256, da0 + static_cast<uint64_t>(k) * 16, db0 + static_cast<uint64_t>(k) * 16, idesc,
                               first ? 0u : 1u);
                  first = 0;
                }
              } else {
                // image rows 0-63 = hi parts, 64-127 = lo parts of the same 64 points: hi.hi + lo.hi + hi.lo
#pragma unroll
                for (int k = 0; k < kParityPoints / 16; ++k) {
#pragma unroll
                  for (int c = 0; c < 3; ++c) {
                    const uint32_t ka = k + (c == 1 ? 4 : 0), kb = k + (c == 2 ? 4 : 0);
                    umma_f16_ss2(tmem_base + as * 256, da0 + static_cast<uint64_t>(ka) * 16, db0 + static_cast<uint64_t>(kb) * 16, idesc,
                                 first ? 0u : 1u);
                    first = 0;
                  }
                }
              }
              if (d.X && term == 0) {
                const uint64_t dx0 = make_smem_desc(b_addr + kDw2Half, lbo, sbo);
                const uint32_t tx = tmem_base + (as ^ 1) * 256;
                if (!parity) {
#pragma unroll
                  for (int k = 0; k < kTileRows / 16; ++k) {
                    umma_f16_ss2(tx, da0 + static_cast<uint64_t>(k) * 16, dx0 + static_cast<uint64_t>(k) * 16, idesc_x, first_x ? 0u : 1u);
                    first_x = 0;
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < kParityPoints / 16; ++k) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                      const uint32_t ka = k + (c == 1 ? 4 : 0), kb = k + (c == 2 ? 4 : 0);
                      umma_f16_ss2(tx, da0 + static_cast<uint64_t>(ka) * 16, dx0 + static_cast<uint64_t>(kb) * 16, idesc_x, first_x ? 0u : 1u);
                      first_x = 0;
                    }
                  }
                }
              }
              umma_commit2(&empty[stage], 3);
              if (++stage == kDw2Stages) { stage = 0; ++round; }
            }
          }
          umma_commit2(&acc_full[as], 3);
          if (d.X) umma_commit2(&acc_full[as ^ 1], 3);
        }
      }
    }
  } else {
    uint32_t item_cnt = 0;
    uint32_t uses[2] = {0, 0};
    const int m = static_cast<int>(rank) * 128 + warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int it = pair; it < n_items; it += n_pairs, ++item_cnt) {
      const DwItem d = items[it];
      const uint32_t as = item_cnt & 1;
      mbar_wait(&acc_full[as], uses[as] & 1);
      ++uses[as];
      if (d.X) {
        mbar_wait(&acc_full[as ^ 1], uses[as ^ 1] & 1);
        ++uses[as ^ 1];
      }
      tc_fence_after();
      if (d.tile1 > d.tile0) {
        const float im = d.inv_mult ? __ldg(d.inv_mult) : 1.f;
        for (uint32_t c0 = 0; c0 < 256; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + as * 256 + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float* o = d.out + static_cast<size_t>(c0 + j) * d.out_ld + m;
            if (d.store) *o = __uint_as_float(r[j]) * im;
            else red_add_f32(o, __uint_as_float(r[j]) * im);
          }
        }
        if (d.out_x) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (as ^ 1) * 256, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float* o = d.out_x + static_cast<size_t>(j) * d.out_ld + m;
            if (d.store) *o = __uint_as_float(r[j]) * im;
            else red_add_f32(o, __uint_as_float(r[j]) * im);
          }
        }
      }
      tc_fence_before();
      if (rank == 0) mbar_arrive(&acc_empty[as]);
      else mbar_arrive_cluster(&acc_empty[as], 0);
      if (d.X) {
        if (rank == 0) mbar_arrive(&acc_empty[as ^ 1]);
        else mbar_arrive_cluster(&acc_empty[as ^ 1], 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's shared memory and TMEM stay valid until both CTAs are done
  if (warp == 5) tmem_dealloc2<512>(tmem_base);
}

cudaError_t launch_dw2(const DwItem* items_dev, int n_items, int dtype, int parity, int num_sms, int* sync_ctr, int group,
                       cudaStream_t stream) {
  static bool attr_set[2][kMaxDevices] = {};
  auto kfn = dtype == DT_BF16 ? dw2_kernel<DT_BF16> : dw2_kernel<DT_FP16>;
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dtype][dev]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kDw2Smem));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) attr_set[dtype][dev] = true;
  }
  if (n_items <= 0) return cudaSuccess;
  int pairs = num_sms / 2;
  if (n_items < pairs) pairs = n_items;
  // lock-step groups need every group inside one round of the item deal: pairs_eff is a multiple of the group size and
  // all of them are co-resident (one CTA per SM), so the group barrier cannot deadlock
  int pairs_eff = pairs;
  if (group > 1 && sync_ctr) {
    pairs_eff = pairs / group * group;
    if (pairs_eff == 0 || n_items % group != 0) { pairs_eff = pairs; group = 1; }
  } else {
    group = 1;
  }
  if (group > 1) {
    cudaError_t e = cudaMemsetAsync(sync_ctr, 0, sizeof(int) * (n_items / group), stream);
    if (e != cudaSuccess) return e;
  }
  if (group > 1) {
    // The group barrier needs every CTA of the grid resident at once.  On an idle GPU that is what happens anyway (one
    // CTA per SM by shared memory); a cooperative launch makes it a guarantee when another stream's kernel holds SMs
    // (gang scheduling) instead of a possible deadlock.  If the launch is refused, run without the lock-step.
    static bool coop_refused[kMaxDevices] = {};
    bool& coop_ok_ref = coop_refused[dev >= 0 && dev < kMaxDevices ? dev : 0];
    const bool coop_ok = !coop_ok_ref;
    if (coop_ok) {
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      cfg.gridDim = dim3(2 * pairs);
      cfg.blockDim = dim3(kDwThreads);
      cfg.dynamicSmemBytes = kDw2Smem;
      cfg.stream = stream;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, items_dev, n_items, parity, sync_ctr, group, pairs_eff);
      if (e == cudaSuccess) return cudaGetLastError();
      cudaGetLastError();  // clear; fall back below
      coop_ok_ref = true;
    }
    group = 1;
    pairs_eff = pairs;
  }
  kfn<<<2 * pairs, kDwThreads, kDw2Smem, stream>>>(items_dev, n_items, parity, sync_ctr, group, pairs_eff);
  return cudaGetLastError();
}

}  // namespace sdf
