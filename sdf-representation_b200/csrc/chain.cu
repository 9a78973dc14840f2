// Layer-chain kernel: one persistent CTA per SM walks point tiles through a *program* of layer GEMMs
// (tcgen05.mma, fp32 accumulators in TMEM) whose epilogues apply the activation / derivative math and write
// the next A operand straight back into shared memory, so point activations never leave the SM inside a sweep.
//
// The same kernel runs four sweeps of the hot path (reference model/networks.py:87-106,
// model/losses.py:283-296):
//   forward            z_l = W_l h_l + b_l, a_l = act(z_l)                   (EPI_FWD_*)
//   reverse            q_{l-1} = act'(z_{l-1}) * (W_l^T q_l)   -> g = grad_x f (EPI_REV_*)
//   tangent            adot_l = act'(z_l) * (W_l hdot_l)                     (EPI_TAN_*)
//   adjoint            Zbar_l = act'(z_l) * (W_{l+1}^T Zbar_{l+1}) + u_l     (EPI_ADJ_*)
// State that must survive between sweeps (a_l, q_l, adot_l, u_l, Zbar_l) is spilled as operand images that the
// weight-gradient kernel (dw.cu) consumes directly.
//
// Two precision modes (template parameter PAR):
//   fast   : tile = 128 points, one 16-bit operand per value, one MMA pass             (error ~ 2^-11 relative)
//   parity : tile = 64 points; image rows 0-63 hold hi = rnd16(v), rows 64-127 hold lo = rnd16(v - hi); the weight
//            stream carries W_hi chunks then W_lo chunks.  One M=128 accumulator then holds v_hi.(W_hi+W_lo) in lanes
//            0-63 and v_lo.(W_hi+W_lo) in lanes 64-127; the epilogue adds the two lane halves (fp32-grade result); each
//            half does the arithmetic of 16 of every 32 columns (load_acc16_split).
//
// Overlap (H = 512): the 128 x 512 accumulator is two 256-column halves with their own completion barriers.  The
// epilogue of half 0 runs while the MMAs of half 1 still read the A image -- its results go straight into the quarters
// those MMAs have left (SDF_DIRECT_A, forward / reverse) or are parked in TMEM until they have finished (tangent /
// adjoint, parity mode) --; the epilogue of half 1 runs while the next layer's MMAs already consume the K < 256 part
// of the new A image (hidden_step_staged, hidden_step_staged_v1, hidden_step_staged_par).
//
// Warp roles (640 threads = 5 warpgroups; setmaxnreg moves registers from the last warpgroup to the epilogue):
//   warps 0-15  epilogue: warp w owns TMEM lanes 32*(w%4).. (tile rows) and column quarter w/4 of every 128-column block
//   warp 16     weight producer (cp.async.bulk into a 3-stage ring)
//   warp 17     MMA issuer: warp-uniform loop, one elected lane issues (pair kernel, peer CTA: "my weight half has
//               landed" relay)
//   warp 18     pair kernel, peer CTA: relays the completed A-image quarters to the leader CTA; otherwise idle
//   warp 19     forward / reverse programs: spill-store thread (bulk copies of completed A-image quarters to the spill
//               images, SDF_BULK_SPILL); otherwise idle
// Four epilogue warps per SM sub-partition give the latency hiding (TMEM reads, spill-image loads, MUFU) that the
// accumulator hand-off needs.
//
// At width 512 in fast mode the kernel runs on CTA pairs by default (PAIR = true: cluster of 2, tcgen05 cta_group::2,
// each CTA streams half of every weight chunk); the single-CTA form serves the other widths, the parity mode and
// SDFNET_PAIR=0.  Each form is instantiated per program family (SW: forward / reverse, tangent / adjoint).
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
#include "sdf_internal.h"

namespace sdf {

// measurement build (-DSDF_CLOCKS): thread 0 accumulates cycles per epilogue kind and phase
#ifdef SDF_CLOCKS
#define CLK_DECL long long clk_t0 = clock64()
#define CLK_MARK(kind, phase)                         \
  if (threadIdx.x == 0) {                             \
    const long long clk_t1 = clock64();               \
    e.clk[((kind) == 8 ? 5 : (kind) == 9 ? 0 : (kind)) * 8 + (phase)] += clk_t1 - clk_t0;   \
    clk_t0 = clk_t1;                                  \
  }
#else
#define CLK_DECL
#define CLK_MARK(kind, phase)
#endif

// register split between the epilogue warps and the producer / issuer warpgroup (640 x 96 = 61440 registers in all)
#ifndef SDF_REGS_EPI
#define SDF_REGS_EPI 104
#define SDF_REGS_AUX 56
#endif
#ifndef SDF_ISSUER_UNROLL
#define SDF_ISSUER_UNROLL 1
#endif

constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kChainThreads = kEpiThreads + 128;
constexpr uint32_t kABytes = kMaxH / 8 * kGranuleBytes;  // 131072
constexpr uint32_t kScratchBytes = 4 * 32 * kParityPoints * 4;  // parity: lane-half exchange, [h][32 cols][64 points] fp32

// PAIR: the kernel runs on CTA pairs (cluster of 2, tcgen05 cta_group::2).  Each CTA keeps its own 128-point tile, A
// image, accumulators and epilogue; the leader CTA issues M = 256 MMAs over both tiles, and each CTA streams only its
// half (N / 2 rows) of every weight chunk: half the L2 -> SM weight bytes and half the shared-memory operand traffic
// per FLOP; the ring stages keep their 32 KB (bulk copies have a large fixed cost) and hold half chunks of twice the K.
template <bool PAR, bool PAIR = false>
struct ChainCfg {
  static constexpr int kStages = PAR ? 2 : 3;
  static constexpr uint32_t kStage = kStageBytes;  // bytes of one ring stage (pair: half of a K = 128 chunk)
  static constexpr int kPoints = PAR ? kParityPoints : kTileRows;
  static constexpr uint32_t kBarOff = kABytes + kStages * kStage;
  static constexpr uint32_t kGskipOff = kBarOff + 256;
  static constexpr uint32_t kScratchOff = kGskipOff + kTileRows * 3 * 4;
  static constexpr uint32_t kSmem = kScratchOff + (PAR ? kScratchBytes : 0);
};

template <int DT>
struct H16;
template <>
struct H16<DT_FP16> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  static __device__ __forceinline__ float rnd(float a) { return __half2float(__float2half_rn(a)); }
};
template <>
struct H16<DT_BF16> {
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
  static __device__ __forceinline__ float rnd(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
};

__device__ __forceinline__ void st_smem_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_smem_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

template <int DT>
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 o;
  o.x = H16<DT>::pack(v[0], v[1]);
  o.y = H16<DT>::pack(v[2], v[3]);
  o.z = H16<DT>::pack(v[4], v[5]);
  o.w = H16<DT>::pack(v[6], v[7]);
  return o;
}
template <int DT>
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  float2 t;
  t = H16<DT>::unpack(u.x); v[0] = t.x; v[1] = t.y;
  t = H16<DT>::unpack(u.y); v[2] = t.x; v[3] = t.y;
  t = H16<DT>::unpack(u.z); v[4] = t.x; v[5] = t.y;
  t = H16<DT>::unpack(u.w); v[6] = t.x; v[7] = t.y;
}
template <int DT>
__device__ __forceinline__ void unpack8_add(const uint4& u, float* v) {
  float2 t;
  t = H16<DT>::unpack(u.x); v[0] += t.x; v[1] += t.y;
  t = H16<DT>::unpack(u.y); v[2] += t.x; v[3] += t.y;
  t = H16<DT>::unpack(u.z); v[4] += t.x; v[5] += t.y;
  t = H16<DT>::unpack(u.w); v[6] += t.x; v[7] += t.y;
}

// Split a float into three 16-bit terms hi + lo + lolo (exact to ~fp32 for bf16, beyond for fp16).
template <int DT>
__device__ __forceinline__ void split3(float x, float& h, float& l, float& ll) {
  h = H16<DT>::rnd(x);
  float r1 = x - h;
  l = H16<DT>::rnd(r1);
  ll = H16<DT>::rnd(r1 - l);
}

// Per-thread epilogue context
struct Epi {
  uint32_t sA;      // shared address of the A image
  uint32_t taddr;   // TMEM base + this warp's lane quarter
  int prow;         // point row inside the tile (image row of the hi / only copy)
  int h;            // column quarter within each 128-column block (this thread's chunks start at h*32 + 128*i)
  bool math;        // fast: always; parity: warps holding the hi lanes do the arithmetic, the others ferry lanes 64-127
  float* scr;       // parity: exchange buffer of this thread's column group, [32][64]
  int64_t grow;     // global point index
  bool valid;
  float x[3];       // input point (forward) or g-bar (tangent)
  float fb;         // f-bar (adjoint start)
  float fval;       // network output of this row (kept for the tanh mode)
#ifdef SDF_CLOCKS
  long long* clk;
#endif
};

// byte offset of (granule of column c, row) inside an operand image; c is a multiple of 8
__device__ __forceinline__ uint32_t img_off(int c, int row) {
  return static_cast<uint32_t>(c >> 3) * kGranuleBytes + row * 16;
}

// 32 accumulator columns of this thread's point.  Parity: every thread of the column group must call (the
// threads of lanes 64-127 hand their half over through shared memory); only e.math threads get a result.
template <bool PAR>
__device__ __forceinline__ void load_acc32(const Epi& e, int c0, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(e.taddr + c0, r);
  tmem_ld_wait();
  if (!PAR) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  } else {
    if (!e.math) {
#pragma unroll
      for (int j = 0; j < 32; ++j) e.scr[j * kParityPoints + e.prow] = __uint_as_float(r[j]);
    }
    named_bar_sync(2 + e.h, 128);
    if (e.math) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + e.scr[j * kParityPoints + e.prow];
    }
    named_bar_sync(2 + e.h, 128);
  }
}

// Parity mode, both lane halves working: of the 32 accumulator columns at c0, the threads of lanes 0-63 (e.math) end up
// with the summed columns c0 .. c0+15 of their point and the threads of lanes 64-127 with c0+16 .. c0+31 -- each half
// hands the other the 16 columns it does not process through shared memory.  (Round-1 form: lanes 64-127 only ferried
// their accumulators and half of the epilogue warps did no arithmetic; the parity epilogue is 3 x its MMA time.)
__device__ __forceinline__ void load_acc16_split(const Epi& e, int c0, float (&v)[16]) {
  uint32_t r[32];
  tmem_ld32(e.taddr + c0, r);
  tmem_ld_wait();
  const int give = e.math ? 16 : 0;   // first column handed over
  const int keep = e.math ? 0 : 16;   // first column processed here
#pragma unroll
  for (int j = 0; j < 16; ++j) e.scr[(give + j) * kParityPoints + e.prow] = __uint_as_float(e.math ? r[16 + j] : r[j]);
  named_bar_sync(2 + e.h, 128);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(e.math ? r[j] : r[16 + j]) + e.scr[(keep + j) * kParityPoints + e.prow];
  named_bar_sync(2 + e.h, 128);
}

__device__ __forceinline__ uint8_t* img_ptr(uint8_t* base, int layer, int n_tiles, size_t tile, uint32_t ib) {
  return base + (static_cast<size_t>(layer) * n_tiles + tile) * ib;
}

// Build the 16-column input block of one row: cols 0-2 hi, 3-5 lo, 6-8 lolo, 9-11 hi, 12-14 lo, 15 = last.
// (The matching layer-0 weight columns are W_hi x3, W_lo x2, 0: the product reproduces fp32 x . W to ~2^-22.)
template <int DT>
__device__ __forceinline__ uint4 xblock_granule(const float (&x)[3], int h, float last) {
  float hi[3], lo[3], ll[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) split3<DT>(x[j], hi[j], lo[j], ll[j]);
  float v[8];
  if (h == 0) {
    v[0] = hi[0]; v[1] = hi[1]; v[2] = hi[2]; v[3] = lo[0]; v[4] = lo[1]; v[5] = lo[2]; v[6] = ll[0]; v[7] = ll[1];
  } else {
    v[0] = ll[2]; v[1] = hi[0]; v[2] = hi[1]; v[3] = hi[2]; v[4] = lo[0]; v[5] = lo[1]; v[6] = lo[2]; v[7] = last;
  }
  return pack8<DT>(v);
}

// write granule g (0 / 1) of a 16-column block for this thread's point (parity: the lo row of a block is zero, the
// split lives in the columns) into shared memory and / or global memory
template <bool PAR>
__device__ __forceinline__ void put_block(const Epi& e, int g, const uint4& v, bool to_smem, uint8_t* gimg) {
  const uint32_t off = g * kGranuleBytes + e.prow * 16;
  const uint4 z = make_uint4(0, 0, 0, 0);
  if (to_smem) {
    st_smem_v4(e.sA + off, v);
    if (PAR) st_smem_v4(e.sA + off + kParityPoints * 16, z);
  }
  if (gimg) {
    *reinterpret_cast<uint4*>(gimg + off) = v;
    if (PAR) *reinterpret_cast<uint4*>(gimg + off + kParityPoints * 16) = z;
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogue building blocks
// ------------------------------------------------------------------------------------------------
struct Bars {
  uint64_t* acc_full;  // [2]  MMA issuer -> epilogue: accumulator half g is complete
  uint64_t* a_q;       // [5]  epilogue -> MMA issuer: A-image columns [128 j, 128 j + 128) are in place (j < 4); [4]: the
                       //      EVEN 16-column slices of quarter 3 are in place (SDF_SPLIT_Q3)
  bool remote;         //      CTA pair, rank 1: a relay warp forwards the quarter barriers to the leader CTA
  uint64_t* a_drain;   //      spill-store thread -> epilogue: the bulk stores of the last spilled A image have read it
  uint64_t* a_free;    // [2]  MMA issuer -> epilogue: the MMAs of accumulator half 1 have read A-image quarter j (j = 0, 1)
};
// SDF_DIRECT_A = 1: the epilogue of accumulator half 0 writes its results straight into quarters 0 / 1 of the A image as
// soon as the MMAs of half 1 -- which walk the K quarters in order -- have read them (a_free[j], committed by the issuer
// after the quarter's last weight chunk).  Round-1 form (0): half 0 parks its packed results in the TMEM columns it has
// consumed and moves them to shared memory after half 1's MMAs have completed ("dump": 0.75-1.3 k cycles per layer step on
// the critical path of a kernel whose critical path is the epilogue, plus 64 KB of tcgen05.st + tcgen05.ld per step).
// Measured (same-session A/B, 1M-point step): forward + reverse 9.40 -> 8.84 ms, 1024^3-class grid inference 321 -> 329 M
// queries/s; tangent + adjoint 13.2 -> 13.7 ms (its half-0 epilogue waits for HBM operands anyway and then collides with the
// a_free waits), so that family keeps the parked form (SDF_DIRECT_A_TA = 0).
#ifndef SDF_DIRECT_A
#define SDF_DIRECT_A 1
#endif
#ifndef SDF_DIRECT_A_TA
#define SDF_DIRECT_A_TA 0
#endif
// SDF_SPLIT_Q3 = 1 (CTA-pair issuer, hot shape): what bounds a layer step is epilogue(half 1) -> the MMAs on the LAST K
// quarter of the next layer -> epilogue(half 0).  Every epilogue thread writes quarter 3 as two 16-column pieces, and the
// first pieces of all threads together are exactly the even K = 16 slices of the quarter: they are announced on a_q[4]
// one piece early, the issuer runs those four MMAs of the last weight chunk at once and only the four odd slices wait for
// a_q[3] -- the MMA tail after the last epilogue arrival is K = 64 instead of K = 128 (~0.5 k of ~1.3 k cycles per step).
// Measured (same-session A/B, 1M-point step): tangent + adjoint 13.1 -> 12.6 ms; forward + reverse 8.86 -> 9.03 ms (there the
// earlier start of the half-0 epilogue runs into its a_free waits), grid inference equal -- so it is on for the tangent /
// adjoint family only (SDF_SPLIT_Q3 = 2: every family).  The K order of the last chunk changes (even slices, then odd), so
// results differ from the unsplit order by fp32 summation order, i.e. by fp16 rounding flips downstream.
#ifndef SDF_SPLIT_Q3
#define SDF_SPLIT_Q3 1
#endif
// Spill images by TMA (SDF_BULK_SPILL = 1).  What the forward / reverse / tangent / adjoint epilogues spill for the
// weight-gradient GEMM (a_l, q_l, adot_l, Zbar_l) is byte for byte the A operand image they have just written to shared
// memory for the next layer's MMAs.  Instead of every epilogue thread also storing its 16-byte granules to global memory,
// one thread of an otherwise idle warp issues cp.async.bulk shared -> global per completed 32 KB quarter of the image
// and tells the epilogue (a_drain barrier) when the image may be overwritten.
#ifndef SDF_BULK_SPILL
#define SDF_BULK_SPILL 1
#endif
// tangent / adjoint programs: off -- measured 14.2-14.5 vs 13.0-13.3 ms with the bulk stores (the kernel is HBM-bound;
// 32 KB store bursts serve it worse than the per-thread stores that are spread over the pieces)
#ifndef SDF_BULK_SPILL_TA
#define SDF_BULK_SPILL_TA 0
#endif
// before the first write to the A image after a bulk-spilled production: its copies have read shared memory
// (drain: -1 = nothing pending, else the parity to wait for; warp-uniform)
__device__ __forceinline__ void drain_wait(const Bars& B, int drain) {
  if (SDF_BULK_SPILL && drain >= 0) mbar_wait(B.a_drain, static_cast<uint32_t>(drain));
}
#ifdef SDF_CLOCKS
#define DRAIN_WAIT(B, d)                                                    \
  {                                                                         \
    const long long dw_t0 = clock64();                                      \
    drain_wait(B, d);                                                       \
    if (threadIdx.x == 0 && (d) >= 0) { e.clk[48] += clock64() - dw_t0; e.clk[54] += 1; } \
  }
#else
#define DRAIN_WAIT(B, d) drain_wait(B, d)
#endif

// One arrival per epilogue warp on quarter barrier j (every lane has fenced its own shared-memory / TMEM accesses).
__device__ __forceinline__ void arrive_quarter(const Bars& B, int j) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    mbar_arrive(&B.a_q[j]);
  }
}
__device__ __forceinline__ void arrive_all_quarters(const Bars& B) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      mbar_arrive(&B.a_q[j]);
    }
  }
}

struct EpiConst {
  int H;
  float beta, nbl2, ln2_over_beta;  // nbl2 = -beta*log2(e)
  bool relu;
#ifdef SDF_CLOCKS
  bool no_math, no_mufu, no_aux;  // FL_DBG_* measurement switches
#endif
  // fast-mode softplus correction log1p(e)/beta = e (sp2[0] + e (sp2[1] + e (sp2[2] + e sp2[3]))), e = exp(-beta |z|),
  // coefficients as packed 16-bit pairs (the arithmetic runs two elements per instruction)
  uint32_t sp2[4];
};

// inputs / outputs of one step for this thread's tile
struct StepIO {
  const uint8_t* in_a;  // activation image whose act' multiplies the result (nullptr: forward)
  const uint8_t* in_b;  // q image (tangent) or u image (adjoint); nullptr if absent
  uint8_t* out;         // spill image of the result (nullptr: not spilled)
  uint8_t* out_u;       // tangent: u image
  const float* vec;     // forward: bias row
  bool special;         // the last three columns of this layer carry the raw input (skip injection)
  // tangent / adjoint sweeps: the images the NEXT step of the program will read (just-in-time L2 prefetch, see
  // hidden_step_staged_v1); cold: the previous step did not request this step's first pieces
  const uint8_t* next_a;
  const uint8_t* next_b;
  bool cold;
  const float* wl;      // EPI_FWD_FUSED: fp32 row of the output layer lin_L
  int drain;            // bulk spill stores: parity of a_drain to wait for before the A image is overwritten (-1: none)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// softplus_beta(z) = max(z,0) + log1p(exp(-|beta z|))/beta   (torch.nn.Softplus(beta, threshold=20): for beta*z > 20
// this differs from the reference's linear branch by < 2.1e-9/beta, far under fp32 resolution of z)
__device__ __forceinline__ float softplus_k(float z, const EpiConst& k) {
#ifdef SDF_CLOCKS
  if (k.no_mufu) return fmaf(fabsf(z * k.nbl2) + 1.f, k.ln2_over_beta, fmaxf(z, 0.f));
#endif
  const float e = ex2_approx(-fabsf(z * k.nbl2));
  return fmaf(lg2_approx(1.f + e), k.ln2_over_beta, fmaxf(z, 0.f));
}
// 1 - act'(z) recovered from the stored activation a = act(z):  exp(-beta a) = 1 - sigmoid(beta z);  relu: [a <= 0]
__device__ __forceinline__ float one_minus_dact(float a, const EpiConst& k) {
#ifdef SDF_CLOCKS
  if (k.no_mufu) return k.relu ? (a > 0.f ? 0.f : 1.f) : a * k.nbl2;
#endif
  return k.relu ? (a > 0.f ? 0.f : 1.f) : ex2_approx(a * k.nbl2);
}

// memory operands of one 16-column piece (two granules): hi rows, and lo rows in parity mode
template <bool PAR>
struct Piece {
  uint4 h[2];
  uint4 l[PAR ? 2 : 1];
};
// L2 eviction-priority hints (createpolicy encodings, as CUTLASS's TMA::CacheHintSm90): the spill images stream through
// L2 exactly once per use -- 148 CTAs move ~76 MB per layer step, of the order of the L2 capacity -- so a line that has
// been consumed (or written for a later kernel) should be the first victim, and a prefetched line should survive until
// its one use.
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
__device__ __forceinline__ uint4 ldg_v4_ef(const void* p) {
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(kPolicyEvictFirst) : "memory");
  return v;
}
__device__ __forceinline__ void stg_v4_ef(void* p, const uint4& v) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(kPolicyEvictFirst) : "memory");
}
#ifndef SDF_AUX_EF
#define SDF_AUX_EF 1
#endif
#ifndef SDF_SPILL_EF
#define SDF_SPILL_EF 1
#endif
#ifndef SDF_FR_SPILL_EF
#define SDF_FR_SPILL_EF 0
#endif
#ifndef SDF_REV_PF_DIST
#define SDF_REV_PF_DIST 0
#endif
// spill-image store of the forward / reverse sweeps (a, q images: read again only many layer steps later)
__device__ __forceinline__ void stg_spill_fr(uint8_t* p, const uint4& v) {
  if (SDF_FR_SPILL_EF) stg_v4_ef(p, v);
  else *reinterpret_cast<uint4*>(p) = v;
}
// spill-image store of the tangent / adjoint sweeps
__device__ __forceinline__ void stg_spill(uint8_t* p, const uint4& v) {
  if (SDF_SPILL_EF) stg_v4_ef(p, v);
  else *reinterpret_cast<uint4*>(p) = v;
}
template <bool PAR>
__device__ __forceinline__ void ldg_piece(const Epi& e, int c, const uint8_t* img, Piece<PAR>& pk) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (SDF_AUX_EF && !PAR) { pk.h[j] = ldg_v4_ef(img + img_off(c + 8 * j, e.prow)); continue; }
    pk.h[j] = *reinterpret_cast<const uint4*>(img + img_off(c + 8 * j, e.prow));
    if (PAR) pk.l[j] = *reinterpret_cast<const uint4*>(img + img_off(c + 8 * j, e.prow + kParityPoints));
  }
}
template <int DT, bool PAR>
__device__ __forceinline__ void granule8(const Piece<PAR>& pk, int j, float (&o)[8]) {
  unpack8<DT>(pk.h[j], o);
  if (PAR) unpack8_add<DT>(pk.l[j], o);
}
// 8 values of this thread's point -> the granule of column c of a global image (parity: hi row and lo row)
template <int DT, bool PAR>
__device__ __forceinline__ void stg_granule8(const Epi& e, uint8_t* img, int c, const float (&v)[8]) {
  if (!PAR) {
    *reinterpret_cast<uint4*>(img + img_off(c, e.prow)) = pack8<DT>(v);
  } else {
    float hi[8], lo[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      hi[t] = H16<DT>::rnd(v[t]);
      lo[t] = v[t] - hi[t];
    }
    *reinterpret_cast<uint4*>(img + img_off(c, e.prow)) = pack8<DT>(hi);
    *reinterpret_cast<uint4*>(img + img_off(c, e.prow + kParityPoints)) = pack8<DT>(lo);
  }
}
// 8 values of this thread's point -> the granule of column c of the A image in shared memory and (optionally) of a
// global image
template <int DT, bool PAR>
__device__ __forceinline__ void put_granule8(const Epi& e, uint8_t* img, int c, const float (&v)[8]) {
  if (!PAR) {
    const uint4 pk = pack8<DT>(v);
    st_smem_v4(e.sA + img_off(c, e.prow), pk);
    if (img) *reinterpret_cast<uint4*>(img + img_off(c, e.prow)) = pk;
  } else {
    float hi[8], lo[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      hi[t] = H16<DT>::rnd(v[t]);
      lo[t] = v[t] - hi[t];
    }
    const uint4 ph = pack8<DT>(hi), pl = pack8<DT>(lo);
    st_smem_v4(e.sA + img_off(c, e.prow), ph);
    st_smem_v4(e.sA + img_off(c, e.prow + kParityPoints), pl);
    if (img) {
      *reinterpret_cast<uint4*>(img + img_off(c, e.prow)) = ph;
      *reinterpret_cast<uint4*>(img + img_off(c, e.prow + kParityPoints)) = pl;
    }
  }
}

// Operands of one 16-column piece that come from memory: forward = the bias row (fp32: granule j = xa.h[j] | xb.h[j]);
// other sweeps = the activation image (xa) and the q / u image (xb).
template <bool PAR, int EPI>
__device__ __forceinline__ void load_aux(const Epi& e, const StepIO& io, int c, Piece<PAR>& xa, Piece<PAR>& xb) {
  if (EPI == EPI_FWD_HIDDEN || EPI == EPI_FWD_FUSED) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      xa.h[j] = __ldg(reinterpret_cast<const uint4*>(io.vec + c + 8 * j));
      xb.h[j] = __ldg(reinterpret_cast<const uint4*>(io.vec + c + 8 * j + 4));
    }
  } else {
    ldg_piece<PAR>(e, c, io.in_a, xa);
    if (io.in_b) ldg_piece<PAR>(e, c, io.in_b, xb);
  }
}

// The arithmetic of one granule (8 columns starting at column c) of a hidden-layer epilogue.
// v: accumulator columns in, result out.  j: granule index (0 / 1) inside the piece the operands were loaded for.
template <int DT, bool PAR, int EPI>
__device__ __forceinline__ void granule_math(const Epi& e, const StepIO& io, const EpiConst& k, int c, int j, float (&v)[8],
                                             const Piece<PAR>& xa, const Piece<PAR>& xb, float* gskip) {
  const bool last = io.special && (c + 8 == k.H);  // columns H-3..H-1 carry the raw input (skip-in)
  if (EPI == EPI_FWD_HIDDEN) {
    // a_l = act(z_l + b_l)                                          reference model/networks.py:100-102
    const float b[8] = {__uint_as_float(xa.h[j].x), __uint_as_float(xa.h[j].y), __uint_as_float(xa.h[j].z), __uint_as_float(xa.h[j].w),
                        __uint_as_float(xb.h[j].x), __uint_as_float(xb.h[j].y), __uint_as_float(xb.h[j].z), __uint_as_float(xb.h[j].w)};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float z = v[t] + b[t];
      v[t] = k.relu ? fmaxf(z, 0.f) : softplus_k(z, k);
    }
    if (last) { v[5] = e.x[0]; v[6] = e.x[1]; v[7] = e.x[2]; }  // next layer sees [a ; x] (networks.py:94-95)
  } else if (EPI == EPI_REV_HIDDEN) {
    // q_{l-1} = (q_l W_l) * act'(z_{l-1})                            autograd of networks.py:100-102 (losses.py:289-295)
    if (last) {
      // rows H-3..H-1 of W_skip multiply the raw input: their product is d f / d x through the skip
      gskip[e.prow * 3 + 0] = v[5]; gskip[e.prow * 3 + 1] = v[6]; gskip[e.prow * 3 + 2] = v[7];
    }
    float a[8];
    granule8<DT, PAR>(xa, j, a);
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] *= 1.f - one_minus_dact(a[t], k);
    if (last) v[5] = v[6] = v[7] = 0.f;
  } else if (EPI == EPI_TAN_HIDDEN) {
    // zdot_l in v;  adot_l = s' zdot_l -> next A;  u_l = act''(z_l) s_l zdot_l = beta (1 - s') q_l zdot_l
    float a[8];
    granule8<DT, PAR>(xa, j, a);
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = one_minus_dact(a[t], k);
    if (last) a[5] = a[6] = a[7] = 1.f;  // input columns: no activation state
    if (!k.relu) {
      float q[8];
      granule8<DT, PAR>(xb, j, q);
#pragma unroll
      for (int t = 0; t < 8; ++t) q[t] = k.beta * a[t] * q[t] * v[t];
      stg_granule8<DT, PAR>(e, io.out_u, c, q);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] *= 1.f - a[t];
    if (last) { v[5] = e.x[0]; v[6] = e.x[1]; v[7] = e.x[2]; }  // tangent of the injected input = g-bar
  } else if (EPI == EPI_TAN_LAST) {
    // zdot_{L-1} in v.  adot = s' zdot -> AD spill (out_u);  Zbar_{L-1} = s' w_L fbar + beta (1 - s') q zdot -> result
    float a[8];
    granule8<DT, PAR>(xa, j, a);
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = one_minus_dact(a[t], k);
    float ad[8], q[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { ad[t] = v[t] * (1.f - a[t]); q[t] = 0.f; }
    if (!k.relu) granule8<DT, PAR>(xb, j, q);
    stg_granule8<DT, PAR>(e, io.out_u, c, ad);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(io.vec + c));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(io.vec + c + 4));
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = fmaf(k.beta * a[t] * q[t], v[t], (1.f - a[t]) * (w[t] * e.fb));
  } else if (EPI == EPI_ADJ_HIDDEN) {
    // Zbar_l = s' (Zbar_{l+1} W_{l+1}) + u_l
    float a[8];
    granule8<DT, PAR>(xa, j, a);
#pragma unroll
    for (int t = 0; t < 8; ++t) a[t] = 1.f - one_minus_dact(a[t], k);
    if (last) a[5] = a[6] = a[7] = 0.f;  // those rows fed the raw input
    if (io.in_b) {
      float u[8];
      granule8<DT, PAR>(xb, j, u);
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = fmaf(a[t], v[t], u[t]);
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] *= a[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fast-mode arithmetic on packed 16-bit pairs.  The epilogue is instruction-issue bound (about one warp instruction
// per clock and SM sub-partition against ~8000 elements per accumulator half), so everything that can run on two
// elements per instruction does: exp(-beta a) is evaluated by one packed multiply and one packed MUFU.EX2 straight
// on the image words, and second operands (q, u) stay packed.  The 16-bit rounding this adds is of the size of the
// rounding the stored images carry anyway (parity mode keeps the fp32 path).
// ------------------------------------------------------------------------------------------------
template <int DT>
struct P2;
template <>
struct P2<DT_FP16> {
  static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t ex2(uint32_t a) {
    uint32_t d;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a));
    return d;
  }
  static __device__ __forceinline__ uint32_t fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
  }
  static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t max(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t abs(uint32_t a) {
    uint32_t d;
    asm("abs.f16x2 %0, %1;" : "=r"(d) : "r"(a));
    return d;
  }
};
template <>
struct P2<DT_BF16> {
  static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t ex2(uint32_t a) {
    uint32_t d;
    asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(d) : "r"(a));
    return d;
  }
  static __device__ __forceinline__ uint32_t fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
  }
  static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t max(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  }
  static __device__ __forceinline__ uint32_t abs(uint32_t a) {
    uint32_t d;
    asm("abs.bf16x2 %0, %1;" : "=r"(d) : "r"(a));
    return d;
  }
};
__device__ __forceinline__ uint32_t word(const uint4& u, int i) { return i == 0 ? u.x : (i == 1 ? u.y : (i == 2 ? u.z : u.w)); }

// One granule (8 columns) of a hidden-layer epilogue in fast mode, softplus activation, no skip-injection columns.
// v: accumulator columns; xa / xb: the granule's words of the activation image and of the q / u image (forward: the
// bias row as fp32 bits).  Returns the packed result; *ou receives the packed u (tangent sweep).
template <int DT, int EPI>
__device__ __forceinline__ uint4 granule_fast(const float (&v)[8], const uint4& xa, const uint4& xb, const EpiConst& k,
                                              uint32_t nbl2_2, bool has_b, uint4* ou, const float* wl = nullptr, float* dot = nullptr) {
  uint32_t o[4];
  if (EPI == EPI_FWD_FUSED) {
    // a = softplus(z + b) as in EPI_FWD_HIDDEN (-> *ou, the A spill); f accumulates w_L . a in fp32; has_b: a reverse sweep
    // follows and the result is q = w_L * act'(z) = w_L (1 - exp(-beta a)) (model/losses.py:289-295 seeds grad_outputs = 1)
    const float b[8] = {__uint_as_float(xa.x), __uint_as_float(xa.y), __uint_as_float(xa.z), __uint_as_float(xa.w),
                        __uint_as_float(xb.x), __uint_as_float(xb.y), __uint_as_float(xb.z), __uint_as_float(xb.w)};
    uint32_t ao[4];
    float d = *dot;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t zh = H16<DT>::pack(v[2 * i] + b[2 * i], v[2 * i + 1] + b[2 * i + 1]);
      const uint32_t ex = P2<DT>::ex2(P2<DT>::mul(P2<DT>::abs(zh), nbl2_2));
      uint32_t pl = P2<DT>::fma(ex, k.sp2[3], k.sp2[2]);
      pl = P2<DT>::fma(pl, ex, k.sp2[1]);
      pl = P2<DT>::fma(pl, ex, k.sp2[0]);
      ao[i] = P2<DT>::fma(pl, ex, P2<DT>::max(zh, 0u));
      const float2 af = H16<DT>::unpack(ao[i]);
      d = fmaf(af.x, wl[2 * i], d);
      d = fmaf(af.y, wl[2 * i + 1], d);
      if (has_b) {
        const float2 ef = H16<DT>::unpack(P2<DT>::ex2(P2<DT>::mul(ao[i], nbl2_2)));  // exp(-beta a) = 1 - act'(z)
        o[i] = H16<DT>::pack(fmaf(-ef.x, wl[2 * i], wl[2 * i]), fmaf(-ef.y, wl[2 * i + 1], wl[2 * i + 1]));
      } else {
        o[i] = ao[i];
      }
    }
    *dot = d;
    *ou = make_uint4(ao[0], ao[1], ao[2], ao[3]);
  } else if (EPI == EPI_FWD_HIDDEN) {
    // a = max(z, 0) + log1p(exp(-beta |z|)) / beta.  The correction is at most ln2 / beta and the result is rounded to
    // 16 bits anyway, so it is evaluated on packed pairs: one MUFU (ex2) per element instead of ex2 + lg2 - the
    // forward epilogue was MUFU-bound (ncu: mio-throttle stalls on every MUFU) - and a cubic on the FMA pipe.
    const float b[8] = {__uint_as_float(xa.x), __uint_as_float(xa.y), __uint_as_float(xa.z), __uint_as_float(xa.w),
                        __uint_as_float(xb.x), __uint_as_float(xb.y), __uint_as_float(xb.z), __uint_as_float(xb.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t zh = H16<DT>::pack(v[2 * i] + b[2 * i], v[2 * i + 1] + b[2 * i + 1]);
      const uint32_t ex = P2<DT>::ex2(P2<DT>::mul(P2<DT>::abs(zh), nbl2_2));
      uint32_t pl = P2<DT>::fma(ex, k.sp2[3], k.sp2[2]);
      pl = P2<DT>::fma(pl, ex, k.sp2[1]);
      pl = P2<DT>::fma(pl, ex, k.sp2[0]);
      o[i] = P2<DT>::fma(pl, ex, P2<DT>::max(zh, 0u));
    }
  } else {
    uint32_t uo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 ef = H16<DT>::unpack(P2<DT>::ex2(P2<DT>::mul(word(xa, i), nbl2_2)));  // exp(-beta a) = 1 - act'(z)
      const float v0 = v[2 * i], v1 = v[2 * i + 1];
      if (EPI == EPI_REV_HIDDEN) {
        o[i] = H16<DT>::pack(fmaf(-ef.x, v0, v0), fmaf(-ef.y, v1, v1));
      } else if (EPI == EPI_TAN_HIDDEN) {
        const float w0 = v0 * ef.x, w1 = v1 * ef.y;             // (1 - s') zdot
        o[i] = H16<DT>::pack(v0 - w0, v1 - w1);                 // adot = s' zdot
        uo[i] = P2<DT>::mul(H16<DT>::pack(w0 * k.beta, w1 * k.beta), word(xb, i));  // u = beta (1 - s') zdot q
      } else if (EPI == EPI_TAN_LAST) {
        const float w0 = v0 * ef.x, w1 = v1 * ef.y;             // (1 - s') zdot
        uo[i] = H16<DT>::pack(v0 - w0, v1 - w1);                // adot = s' zdot                    -> AD spill
        const float2 qf = H16<DT>::unpack(word(xb, i));
        // Zbar = s' (w_L fbar) + beta (1 - s') zdot q  (u stays in fp32 registers)           -> Z spill + next A operand
        o[i] = H16<DT>::pack(fmaf(w0 * k.beta, qf.x, fmaf(-ef.x, wl[2 * i], wl[2 * i])),
                             fmaf(w1 * k.beta, qf.y, fmaf(-ef.y, wl[2 * i + 1], wl[2 * i + 1])));
      } else {  // EPI_ADJ_HIDDEN
        o[i] = H16<DT>::pack(fmaf(-ef.x, v0, v0), fmaf(-ef.y, v1, v1));
        if (has_b) o[i] = P2<DT>::add(o[i], word(xb, i));       // + u
      }
    }
    if (EPI == EPI_TAN_HIDDEN || EPI == EPI_TAN_LAST) *ou = make_uint4(uo[0], uo[1], uo[2], uo[3]);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

#ifndef SDF_PARITY_SPLIT
#define SDF_PARITY_SPLIT 1
#endif
// Hidden-layer epilogue, general form: wait for the whole accumulator, then walk the columns.  Used in parity mode
// and whenever the GEMM has a single column group (H <= 256).
template <int DT, bool PAR, int EPI>
__device__ __forceinline__ void hidden_step_plain(Epi& e, const StepIO& io, const EpiConst& k, const StepDesc& st, const Bars& B,
                                                  uint32_t par0, uint32_t par1, float* gskip) {
  mbar_wait(&B.acc_full[0], par0);
  if (st.n_groups == 2) mbar_wait(&B.acc_full[1], par1);
  tc_fence_after();
  drain_wait(B, io.drain);
  for (int c0 = e.h * 32; c0 < k.H; c0 += 128) {
    if constexpr (PAR && SDF_PARITY_SPLIT) {
      float v16[16];
      load_acc16_split(e, c0, v16);
      const int cs = c0 + (e.math ? 0 : 16);
      Piece<PAR> xa, xb;
      load_aux<PAR, EPI>(e, io, cs, xa, xb);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float w[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) w[t] = v16[8 * j + t];
        const int c = cs + 8 * j;
        granule_math<DT, PAR, EPI>(e, io, k, c, j, w, xa, xb, gskip);
        put_granule8<DT, PAR>(e, io.out, c, w);
      }
      continue;
    }
    float v[32];
    load_acc32<PAR>(e, c0, v);
    if (e.math) {
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        Piece<PAR> xa, xb;
        load_aux<PAR, EPI>(e, io, c0 + 16 * sub, xa, xb);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float w[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) w[t] = v[16 * sub + 8 * j + t];
          const int c = c0 + 16 * sub + 8 * j;
          granule_math<DT, PAR, EPI>(e, io, k, c, j, w, xa, xb, gskip);
          put_granule8<DT, PAR>(e, io.out, c, w);
        }
      }
    }
  }
  fence_proxy_async_smem();
  if (st.signal) {
    tc_fence_before();
    arrive_all_quarters(B);
  }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// Hidden-layer epilogue, overlapped form (fast mode, H = 512: two accumulator halves of 256 columns).
//   half 0 is processed while the MMAs of half 1 still read the A image, so its packed 16-bit result cannot go
//   to shared memory yet: it is parked in TMEM, in the first 16 of the 32 accumulator columns the thread has just
//   consumed.  Once half 1 is complete the parked columns move into the A image (-> quarters 0, 1 released: the
//   next GEMM may start on K < 256 and overwrite accumulator half 0) and half 1 is processed straight into the A
//   image (-> quarter 2 released half-way, quarter 3 at the end).
// Each thread owns 2 chunks of 32 columns per half (4 pieces of 16).  Every spill-image line the step will read is
// first requested into L2 (prefetch.global.L2, while the MMAs of the step are still running); the four epilogue
// warps of an SM sub-partition cover each other's load latencies.
// Overlapped hidden-layer epilogue, loop form: one piece at a time, both halves share the code.  Used by the tangent /
// adjoint sweeps (the software-pipelined form below spills in their piece loop and measured 16 % slower there).
// Just-in-time L2 prefetch of the tangent / adjoint sweeps (SDF_TA_PF_DIST = D > 0): while piece P is processed, the
// lines of piece P + D are requested -- pieces of the NEXT step once P + D runs past this step's 8 pieces.  A line then
// has to survive in L2 for D pieces (a few microseconds) instead of up to a whole layer step: with 148 CTAs streaming
// ~76 MB per step through a 126 MB L2, the step-start burst (D = 0) loses its late lines again before they are used
// (round 1: DRAM reads 16 % above the operand bytes, long-scoreboard stalls on the image loads).
// (Requesting the same lines with cp.async.bulk.prefetch.L2 from an idle warp, 64 KB per instruction and half a layer step
// ahead, measured equal within noise -- 13.1 vs 13.2 ms -- as did D = 3, 4, 6: the sweeps are not bound by L2 misses.)
#ifndef SDF_TA_PF_DIST
#define SDF_TA_PF_DIST 2
#endif
#ifndef SDF_PF_EL
#define SDF_PF_EL 0
#endif
__device__ __forceinline__ void prefetch_l2_hint(const void* ptr) {
  if (SDF_PF_EL) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(ptr));
  else asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}
// first column of piece P (0..7) of a 512-wide step for this thread: half P / 4, chunk (P / 2) % 2, sub-piece P % 2
__device__ __forceinline__ int piece_col(const Epi& e, int P) { return (P >> 2) * 256 + ((P >> 1) & 1) * 128 + e.h * 32 + (P & 1) * 16; }
__device__ __forceinline__ void prefetch_piece(const Epi& e, const uint8_t* a, const uint8_t* b, int P) {
  const int c = piece_col(e, P);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    prefetch_l2_hint(a + img_off(c + 8 * j, e.prow));
    if (b) prefetch_l2_hint(b + img_off(c + 8 * j, e.prow));
  }
}

template <int DT, int EPI>
__device__ __forceinline__ void hidden_step_staged_v1(Epi& e, const StepIO& io, const EpiConst& k, const StepDesc& st, const Bars& B,
                                                   uint32_t par0, uint32_t par1, float* gskip) {
#if SDF_TA_PF_DIST > 0
  if (io.cold) {
#pragma unroll
    for (int P = 0; P < SDF_TA_PF_DIST && P < 8; ++P) prefetch_piece(e, io.in_a, io.in_b, P);
  }
#elif !defined(SDF_NO_L2PF)
  // Every spill-image line the step will read is first requested into L2 (same-session A/B on B200: +3 % for the
  // tangent / adjoint sweeps, whose operands come from HBM; requesting the *next* step's lines instead: -18 %, L2 thrash)
  if (EPI == EPI_TAN_HIDDEN || EPI == EPI_ADJ_HIDDEN || EPI == EPI_TAN_LAST) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        prefetch_l2(io.in_a + img_off(ci * 128 + e.h * 32 + 8 * j, e.prow));
        if (io.in_b) prefetch_l2(io.in_b + img_off(ci * 128 + e.h * 32 + 8 * j, e.prow));
      }
    }
  }
#endif
  const uint32_t nbl2_2 = H16<DT>::pack(k.nbl2, k.nbl2);
  Piece<false> xa, xb;
#ifdef SDF_CLOCKS
  xa.h[0] = xa.h[1] = xb.h[0] = xb.h[1] = make_uint4(0, 0, 0, 0);
  if (!k.no_aux)
#endif
  load_aux<false, EPI>(e, io, e.h * 32, xa, xb);
  CLK_DECL;
  mbar_wait(&B.acc_full[0], par0);
  tc_fence_after();
  CLK_MARK(EPI, 0);
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    if (half == 1) {
      if (SDF_DIRECT_A_TA) {
        // half 0 went straight into quarters 0 / 1 of the A image: release them (the next GEMM may start on K < 256 and
        // overwrite accumulator half 0, which this thread has finished reading)
        fence_proxy_async_smem();
        if (st.signal) {
          tc_fence_before();
          arrive_quarter(B, 0);
          arrive_quarter(B, 1);
        }
        CLK_MARK(EPI, 1);
        mbar_wait(&B.acc_full[1], par1);
        tc_fence_after();
        CLK_MARK(EPI, 2);
      } else {
        tmem_st_wait();
        CLK_MARK(EPI, 1);
        mbar_wait(&B.acc_full[1], par1);
        tc_fence_after();
        CLK_MARK(EPI, 2);
        {
          uint32_t r0[16], r1[16];
          const int c0 = e.h * 32;
          tmem_ld16(e.taddr + c0, r0);
          tmem_ld16(e.taddr + c0 + 128, r1);
          if (SDF_BULK_SPILL_TA) DRAIN_WAIT(B, io.drain);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_smem_v4(e.sA + img_off(c0 + 8 * j, e.prow), make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]));
            st_smem_v4(e.sA + img_off(c0 + 128 + 8 * j, e.prow), make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]));
          }
        }
        fence_proxy_async_smem();
        if (st.signal) {
          tc_fence_before();
          arrive_quarter(B, 0);
          arrive_quarter(B, 1);
        }
      }
      CLK_MARK(EPI, 3);
    }
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
      // piece pc of this half: chunk pc / 2, 16-column sub-piece pc % 2
      const int c0 = half * 256 + (pc >> 1) * 128 + e.h * 32;  // chunk base column
      const int c = c0 + (pc & 1) * 16;                        // first column of the piece
      uint32_t r[16];
      tmem_ld16(e.taddr + c, r);
#if SDF_TA_PF_DIST > 0
      {
        const int T = half * 4 + pc + SDF_TA_PF_DIST;
        if (T < 8) prefetch_piece(e, io.in_a, io.in_b, T);
        else if (io.next_a) prefetch_piece(e, io.next_a, io.next_b, T - 8);
      }
#endif
#ifdef SDF_CLOCKS
      if ((pc > 0 || half == 1) && !k.no_aux) load_aux<false, EPI>(e, io, c, xa, xb);
#else
      if (pc > 0 || half == 1) load_aux<false, EPI>(e, io, c, xa, xb);
#endif
      tmem_ld_wait();
      uint4 pq[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float w[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) w[t] = __uint_as_float(r[8 * j + t]);
#ifdef SDF_NO_PACKED
        const bool generic = EPI != EPI_FWD_HIDDEN || k.relu || (io.special && c + 8 * j + 8 == k.H);
#else
        const bool generic = k.relu || (io.special && c + 8 * j + 8 == k.H);
#endif
#ifdef SDF_CLOCKS
        if (k.no_math) {
          pq[j] = pack8<DT>(w);
        } else
#endif
        if (generic) {
          granule_math<DT, false, EPI>(e, io, k, c + 8 * j, j, w, xa, xb, gskip);
          pq[j] = pack8<DT>(w);
        } else {
          uint4 ou;
          float wl[8];
          if (EPI == EPI_TAN_LAST) {
            // w_L[c..c+7] * fbar: the adjoint-start seed of this point (fp32 row of lin_L, L1-resident)
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(io.vec + c + 8 * j));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(io.vec + c + 8 * j + 4));
            wl[0] = w0.x * e.fb; wl[1] = w0.y * e.fb; wl[2] = w0.z * e.fb; wl[3] = w0.w * e.fb;
            wl[4] = w1.x * e.fb; wl[5] = w1.y * e.fb; wl[6] = w1.z * e.fb; wl[7] = w1.w * e.fb;
          }
          pq[j] = granule_fast<DT, EPI>(w, xa.h[j], xb.h[j], k, nbl2_2, io.in_b != nullptr, &ou, wl);
          if (EPI == EPI_TAN_HIDDEN || EPI == EPI_TAN_LAST) stg_spill(io.out_u + img_off(c + 8 * j, e.prow), ou);
        }
      }
      const uint32_t pk[8] = {pq[0].x, pq[0].y, pq[0].z, pq[0].w, pq[1].x, pq[1].y, pq[1].z, pq[1].w};
      const uint4 p0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), p1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      if (io.out) {
        stg_spill(io.out + img_off(c, e.prow), p0);
        stg_spill(io.out + img_off(c + 8, e.prow), p1);
      }
      if (half == 0 && !SDF_DIRECT_A_TA) {
        tmem_st8(e.taddr + c0 + (pc & 1) * 8, pk);  // park: columns [c0, c0+16) hold the chunk's 32 packed results
      } else {
        if (half == 0) {
          // quarter pc / 2 of the A image: the MMAs of accumulator half 1 have read it
          if (pc == 0 && SDF_BULK_SPILL_TA) DRAIN_WAIT(B, io.drain);
          if ((pc & 1) == 0) mbar_wait(&B.a_free[pc >> 1], par1);
        }
        st_smem_v4(e.sA + img_off(c, e.prow), p0);
        st_smem_v4(e.sA + img_off(c + 8, e.prow), p1);
        if (half == 1 && pc == 1) {
          // this thread's share of A columns [256,384) is written: the next GEMM may run two more weight chunks
          fence_proxy_async_smem();
          if (st.signal) arrive_quarter(B, 2);
        }
        if (half == 1 && pc == 2) {
          // ... and the even 16-column slices of A columns [384,512)
          fence_proxy_async_smem();
          if (st.signal) arrive_quarter(B, 4);
        }
      }
    }
  }
  fence_proxy_async_smem();
  if (st.signal) {
    tc_fence_before();
    arrive_quarter(B, 3);
  }
  CLK_MARK(EPI, 4);
#ifdef SDF_CLOCKS
  if (threadIdx.x == 0) e.clk[(EPI == 8 ? 5 : EPI == 9 ? 0 : EPI) * 8 + 5] += 1;
#endif
}

// bit masks over the sweeps (1 forward, 2 reverse, 4 tangent, 8 adjoint): read the next accumulator piece / request the
// next piece's spill-image operands one piece ahead
#ifndef SDF_TMEM_AHEAD
#define SDF_TMEM_AHEAD 15
#endif
#ifndef SDF_AUX_AHEAD
#define SDF_AUX_AHEAD 14
#endif
template <int DT, int EPI>
__device__ __forceinline__ void hidden_step_staged(Epi& e, const StepIO& io, const EpiConst& k, const StepDesc& st, const Bars& B,
                                                   uint32_t par0, uint32_t par1, float* gskip) {
  constexpr bool kFwd = EPI == EPI_FWD_HIDDEN || EPI == EPI_FWD_FUSED;
  constexpr bool kFused = EPI == EPI_FWD_FUSED;
  // kFused, forward-only program (st.signal == 0): nothing reads the A image after this step -- the epilogue only
  // accumulates the output dot product (no parking, no shared-memory stores)
  const bool keep = !kFused || st.signal != 0;
  float dot = 0.f;
  // Software pipeline of a warp (ncu: the sweeps that read spill images sat on the scoreboard of those loads, 60 % of
  // the reverse epilogue's samples): the accumulator piece p + 1 is read from TMEM and the spill-image operands of
  // piece p + 1 are requested while piece p is being processed.  The forward sweep's only memory operand is the bias
  // row (L1-resident), loaded per piece.
  constexpr int kSweepBit = kFwd ? 1 : EPI == EPI_REV_HIDDEN ? 2 : EPI == EPI_TAN_HIDDEN ? 4 : 8;
  constexpr bool kTmemAhead = (SDF_TMEM_AHEAD & kSweepBit) != 0;
  constexpr bool kAuxAhead = !kFwd && (SDF_AUX_AHEAD & kSweepBit) != 0;
  constexpr int kUnrollHalves = (EPI == EPI_TAN_HIDDEN || EPI == EPI_ADJ_HIDDEN) ? 1 : 2;
  constexpr int kAuxDist = kAuxAhead ? 1 : 0;  // (two pieces ahead in the reverse sweep: no gain measured)
#ifndef SDF_NO_L2PF
  // (same-session A/B on B200: +3 % for the tangent / adjoint sweeps, whose operands come from HBM; -2 % for the
  // reverse sweep, whose activation images were written moments earlier and are mostly still in L2)
  if (EPI == EPI_TAN_HIDDEN || EPI == EPI_ADJ_HIDDEN) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        prefetch_l2(io.in_a + img_off(ci * 128 + e.h * 32 + 8 * j, e.prow));
        if (io.in_b) prefetch_l2(io.in_b + img_off(ci * 128 + e.h * 32 + 8 * j, e.prow));
      }
    }
  }
#endif
  const uint32_t nbl2_2 = H16<DT>::pack(k.nbl2, k.nbl2);
  const int cb = e.h * 32;  // this thread's column offset inside every 128-column block
  Piece<false> xa[2], xb[2];
  uint32_t r[2][16];
  load_aux<false, EPI>(e, io, cb, xa[0], xb[0]);
  CLK_DECL;
  mbar_wait(&B.acc_full[0], par0);
  tc_fence_after();
  CLK_MARK(EPI, 0);
  if (kTmemAhead) tmem_ld16(e.taddr + cb, r[0]);
  // the tangent / adjoint sweeps keep the two halves as a loop: unrolled, their register pressure spills into the
  // piece loop (measured: 16 % slower)
#pragma unroll kUnrollHalves
  for (int half = 0; half < 2; ++half) {
    if (half == 1) {
      if (SDF_DIRECT_A) {
        // half 0 went straight into quarters 0 / 1 of the A image: release them
        fence_proxy_async_smem();
        if (st.signal) {
          tc_fence_before();
          arrive_quarter(B, 0);
          arrive_quarter(B, 1);
        }
        CLK_MARK(EPI, 1);
        mbar_wait(&B.acc_full[1], par1);
        tc_fence_after();
        CLK_MARK(EPI, 2);
        if (kTmemAhead) tmem_ld16(e.taddr + 256 + cb, r[0]);
      } else {
        tmem_st_wait();
        CLK_MARK(EPI, 1);
        mbar_wait(&B.acc_full[1], par1);
        tc_fence_after();
        CLK_MARK(EPI, 2);
        if (keep) {
          uint32_t r0[16], r1[16];
          tmem_ld16(e.taddr + cb, r0);
          tmem_ld16(e.taddr + cb + 128, r1);
          DRAIN_WAIT(B, io.drain);
          tmem_ld_wait();
          if (kTmemAhead) tmem_ld16(e.taddr + 256 + cb, r[0]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_smem_v4(e.sA + img_off(cb + 8 * j, e.prow), make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]));
            st_smem_v4(e.sA + img_off(cb + 128 + 8 * j, e.prow), make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]));
          }
        } else if (kTmemAhead) {
          tmem_ld16(e.taddr + 256 + cb, r[0]);
        }
        fence_proxy_async_smem();
        if (st.signal) {
          tc_fence_before();
          arrive_quarter(B, 0);
          arrive_quarter(B, 1);
        }
      }
      CLK_MARK(EPI, 3);
    }
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
      // piece pc of this half: chunk pc / 2, 16-column sub-piece pc % 2
      const int cur = pc & 1, nxt = cur ^ 1;
      const int ac = cur;  // buffer holding this piece's spill-image operands
      const int c0 = half * 256 + (pc >> 1) * 128 + cb;  // chunk base column
      const int c = c0 + (pc & 1) * 16;                  // first column of the piece
      const int cn = pc < 3 ? half * 256 + ((pc + 1) >> 1) * 128 + cb + ((pc + 1) & 1) * 16 : 256 + cb;  // next piece
      if (!kTmemAhead) tmem_ld16(e.taddr + c, r[cur]);
#if SDF_REV_PF_DIST > 0
      if (EPI == EPI_REV_HIDDEN) {
        const int T = half * 4 + pc + 1 + SDF_REV_PF_DIST;  // the operands of piece P + 1 are already being loaded
        if (T < 8) prefetch_piece(e, io.in_a, nullptr, T);
        else if (io.next_a) prefetch_piece(e, io.next_a, nullptr, T - 8);
      }
#endif
      if (!kAuxAhead && (pc > 0 || half == 1)) load_aux<false, EPI>(e, io, c, xa[ac], xb[ac]);
      tmem_ld_wait();
      if (kTmemAhead && pc < 3) tmem_ld16(e.taddr + cn, r[nxt]);
      if (kAuxDist == 1 && (pc < 3 || half == 0)) load_aux<false, EPI>(e, io, cn, xa[nxt], xb[nxt]);
      uint4 pq[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float w[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) w[t] = __uint_as_float(r[cur][8 * j + t]);
        uint4 ou;
        if constexpr (kFused) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(io.wl + c + 8 * j));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(io.wl + c + 8 * j + 4));
          const float wl[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          pq[j] = granule_fast<DT, EPI>(w, xa[ac].h[j], xb[ac].h[j], k, nbl2_2, st.signal != 0, &ou, wl, &dot);
          if (io.out_u) stg_spill_fr(io.out_u + img_off(c + 8 * j, e.prow), ou);  // a_{L-1}: operand of the weight-gradient GEMM
        } else {
          pq[j] = granule_fast<DT, EPI>(w, xa[ac].h[j], xb[ac].h[j], k, nbl2_2, io.in_b != nullptr, &ou);
          if (EPI == EPI_TAN_HIDDEN) *reinterpret_cast<uint4*>(io.out_u + img_off(c + 8 * j, e.prow)) = ou;
        }
      }
      const uint32_t pk[8] = {pq[0].x, pq[0].y, pq[0].z, pq[0].w, pq[1].x, pq[1].y, pq[1].z, pq[1].w};
      if (io.out) {
        stg_spill_fr(io.out + img_off(c, e.prow), pq[0]);
        stg_spill_fr(io.out + img_off(c + 8, e.prow), pq[1]);
      }
      if (!keep) {
        // forward-only fused step: the packed results are not needed by anything
      } else if (half == 0 && !SDF_DIRECT_A) {
        tmem_st8(e.taddr + c0 + (pc & 1) * 8, pk);  // park: columns [c0, c0+16) hold the chunk's 32 packed results
      } else {
        if (half == 0) {
          // quarter pc / 2 of the A image: the MMAs of accumulator half 1 have read it
          if (pc == 0) DRAIN_WAIT(B, io.drain);
          if ((pc & 1) == 0) mbar_wait(&B.a_free[pc >> 1], par1);
        }
        st_smem_v4(e.sA + img_off(c, e.prow), pq[0]);
        st_smem_v4(e.sA + img_off(c + 8, e.prow), pq[1]);
        if (half == 1 && pc == 1) {
          // this thread's share of A columns [256,384) is written: the next GEMM may run two more weight chunks
          fence_proxy_async_smem();
          if (st.signal) arrive_quarter(B, 2);
        }
        if (half == 1 && pc == 2) {
          // ... and the even 16-column slices of A columns [384,512)
          fence_proxy_async_smem();
          if (st.signal) arrive_quarter(B, 4);
        }
      }
    }
  }
  if (io.special && cb + 32 == 128) {
    // The last granule of this layer carries the raw input in its last three columns (skip injection, H = 512): redo
    // it in the general form - accumulator half 1 is intact until quarter 3 is released - and overwrite what the
    // loop above stored for it.  One granule of one layer; keeps the special case out of the piece loop.
    const int c = k.H - 8;
    uint32_t r8[8];
    tmem_ld8(e.taddr + c, r8);
    load_aux<false, EPI>(e, io, c - 8, xa[0], xb[0]);
    tmem_ld_wait();
    float w[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) w[t] = __uint_as_float(r8[t]);
    granule_math<DT, false, EPI>(e, io, k, c, 1, w, xa[0], xb[0], gskip);
    const uint4 pk = pack8<DT>(w);
    if (io.out) *reinterpret_cast<uint4*>(io.out + img_off(c, e.prow)) = pk;
    st_smem_v4(e.sA + img_off(c, e.prow), pk);
  }
  if constexpr (kFused) e.fval = dot;  // this thread's share of w_L . a_{L-1} (128 of the 512 columns)
  fence_proxy_async_smem();
  if (st.signal) {
    tc_fence_before();
    arrive_quarter(B, 3);
  }
  CLK_MARK(EPI, 4);
#ifdef SDF_CLOCKS
  if (threadIdx.x == 0) e.clk[(EPI == 8 ? 5 : EPI == 9 ? 0 : EPI) * 8 + 5] += 1;
#endif
}

// Parity mode at width 512, overlapped form (SDF_PARITY_STAGED): as in the fast mode the epilogue of accumulator half 0 runs
// while the MMAs of half 1 still read the A image -- its packed hi / lo results are parked in the TMEM columns it has
// consumed and move to shared memory when half 1 has completed -- and the quarters of the A image are released as they are
// written, so the next layer's hi-pass MMAs start under the half-1 epilogue.  Each thread owns 16 of the 32 columns of a
// chunk (load_acc16_split); arithmetic is the general fp32 form (granule_math).
#ifndef SDF_PARITY_STAGED
#define SDF_PARITY_STAGED 1
#endif
template <int DT, int EPI>
__device__ __forceinline__ void hidden_step_staged_par(Epi& e, const StepIO& io, const EpiConst& k, const StepDesc& st, const Bars& B,
                                                       uint32_t par0, uint32_t par1, float* gskip) {
  const int sub = e.math ? 0 : 16;  // this thread's 16 columns inside every 32-column chunk
  const int lo_row = e.prow + kParityPoints;
  mbar_wait(&B.acc_full[0], par0);
  tc_fence_after();
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    if (half == 1) {
      tmem_st_wait();
      mbar_wait(&B.acc_full[1], par1);
      tc_fence_after();
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c0 = ci * 128 + e.h * 32;
        uint32_t r[16];
        tmem_ld16(e.taddr + c0, r);
        tmem_ld_wait();
        st_smem_v4(e.sA + img_off(c0 + sub, e.prow), make_uint4(r[0], r[1], r[2], r[3]));
        st_smem_v4(e.sA + img_off(c0 + sub + 8, e.prow), make_uint4(r[4], r[5], r[6], r[7]));
        st_smem_v4(e.sA + img_off(c0 + sub, lo_row), make_uint4(r[8], r[9], r[10], r[11]));
        st_smem_v4(e.sA + img_off(c0 + sub + 8, lo_row), make_uint4(r[12], r[13], r[14], r[15]));
      }
      fence_proxy_async_smem();
      if (st.signal) {
        tc_fence_before();
        arrive_quarter(B, 0);
        arrive_quarter(B, 1);
      }
    }
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      const int c0 = half * 256 + ci * 128 + e.h * 32;
      float v16[16];
      load_acc16_split(e, c0, v16);
      const int cs = c0 + sub;
      Piece<true> xa, xb;
      load_aux<true, EPI>(e, io, cs, xa, xb);
      uint4 ph[2], pl[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float w[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) w[t] = v16[8 * j + t];
        const int c = cs + 8 * j;
        granule_math<DT, true, EPI>(e, io, k, c, j, w, xa, xb, gskip);
        float hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          hi[t] = H16<DT>::rnd(w[t]);
          lo[t] = w[t] - hi[t];
        }
        ph[j] = pack8<DT>(hi);
        pl[j] = pack8<DT>(lo);
        if (io.out) {
          *reinterpret_cast<uint4*>(io.out + img_off(c, e.prow)) = ph[j];
          *reinterpret_cast<uint4*>(io.out + img_off(c, lo_row)) = pl[j];
        }
      }
      if (half == 0) {
        const uint32_t pk[16] = {ph[0].x, ph[0].y, ph[0].z, ph[0].w, ph[1].x, ph[1].y, ph[1].z, ph[1].w,
                                 pl[0].x, pl[0].y, pl[0].z, pl[0].w, pl[1].x, pl[1].y, pl[1].z, pl[1].w};
        tmem_st16(e.taddr + c0, pk);  // park in the first 16 of the 32 accumulator columns this thread has consumed
      } else {
        st_smem_v4(e.sA + img_off(cs, e.prow), ph[0]);
        st_smem_v4(e.sA + img_off(cs + 8, e.prow), ph[1]);
        st_smem_v4(e.sA + img_off(cs, lo_row), pl[0]);
        st_smem_v4(e.sA + img_off(cs + 8, lo_row), pl[1]);
        if (ci == 0) {
          fence_proxy_async_smem();
          if (st.signal) arrive_quarter(B, 2);
        }
      }
    }
  }
  fence_proxy_async_smem();
  if (st.signal) {
    tc_fence_before();
    arrive_quarter(B, 4);
    arrive_quarter(B, 3);
  }
}

template <int DT, bool PAR, int EPI>
__device__ __forceinline__ void hidden_step(Epi& e, const StepIO& io, const EpiConst& k, const StepDesc& st, const Bars& B,
                                            uint32_t& n0, uint32_t& n1, float* gskip) {
  if (!PAR && st.n_groups == 2 && !k.relu) {
#ifndef SDF_TA_PIPE
#define SDF_TA_PIPE 0
#endif
    if constexpr (EPI == EPI_TAN_LAST || (!SDF_TA_PIPE && (EPI == EPI_TAN_HIDDEN || EPI == EPI_ADJ_HIDDEN)))
      hidden_step_staged_v1<DT, EPI>(e, io, k, st, B, n0 & 1, n1 & 1, gskip);
    else
      hidden_step_staged<DT, EPI>(e, io, k, st, B, n0 & 1, n1 & 1, gskip);
  }
  else if (PAR && SDF_PARITY_STAGED && SDF_PARITY_SPLIT && st.n_groups == 2 && !k.relu) {
    if constexpr (PAR) hidden_step_staged_par<DT, EPI>(e, io, k, st, B, n0 & 1, n1 & 1, gskip);
  }
  else hidden_step_plain<DT, PAR, EPI>(e, io, k, st, B, n0 & 1, n1 & 1, gskip);
  ++n0;
  if (st.n_groups == 2) ++n1;
}

// The spill image a step's epilogue produces as a copy of the A operand image it leaves in shared memory (nullptr: none).
__device__ __forceinline__ uint8_t* step_spill_image(const ChainParams& p, const StepDesc& st, int tile, uint32_t ib) {
  const int l = st.layer;
  switch (st.epi) {
    case EPI_FWD_HIDDEN: return (p.flags & FL_SPILL_A) ? img_ptr(p.Aimg, l, p.tile_stride, tile, ib) : nullptr;
    case EPI_FWD_FUSED: return (st.signal && (p.flags & FL_SPILL_Q)) ? img_ptr(p.Qimg, l, p.tile_stride, tile, ib) : nullptr;
    case EPI_REV_HIDDEN: return (p.flags & FL_SPILL_Q) ? img_ptr(p.Qimg, l - 1, p.tile_stride, tile, ib) : nullptr;
    case EPI_TAN_HIDDEN: return img_ptr(p.ADimg, l, p.tile_stride, tile, ib);
    case EPI_TAN_LAST: return img_ptr(p.Zimg, l, p.tile_stride, tile, ib);
    case EPI_ADJ_HIDDEN: return img_ptr(p.Zimg, l, p.tile_stride, tile, ib);
    default: return nullptr;
  }
}
// A step spills by bulk store when it runs the overlapped (staged) epilogue -- fast mode, two accumulator halves (H = 512),
// softplus -- and the image it leaves is consumed by a GEMM (its completion is announced on the a_q barriers, which also
// trigger the store).  The epilogue warps and the store warp evaluate the same predicate.
template <bool PAR>
__device__ __forceinline__ bool step_spills_by_bulk(const ChainParams& p, const StepDesc& st) {
  if (PAR || !SDF_BULK_SPILL) return false;
  if (st.n_groups != 2 || st.signal == 0 || (p.flags & (FL_RELU | FL_DBG_NO_STG))) return false;
  switch (st.epi) {
    case EPI_FWD_HIDDEN: return (p.flags & FL_SPILL_A) != 0;
    case EPI_FWD_FUSED:
    case EPI_REV_HIDDEN: return (p.flags & FL_SPILL_Q) != 0;
    case EPI_TAN_HIDDEN:
    case EPI_TAN_LAST:
    case EPI_ADJ_HIDDEN: return SDF_BULK_SPILL_TA != 0;
    default: return false;
  }
}

template <bool PAIR>
__device__ __forceinline__ void umma_any(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (PAIR) umma_f16_ss2(tmem_d, da, db, idesc, accumulate);
  else umma_f16_ss(tmem_d, da, db, idesc, accumulate);
}

// SW: sweeps this instantiation can run (1 = forward / reverse programs, 2 = tangent / adjoint programs, 3 = both).  The
// fast single-CTA kernel is instantiated per program family: half the code per kernel (the full kernel is ~10 k SASS
// instructions, more than the instruction caches hold).
template <int DT, bool PAR, bool PAIR, int SW>
__global__ void __launch_bounds__(kChainThreads, 1) chain_kernel(const __grid_constant__ ChainParams p) {
  using Cfg = ChainCfg<PAR, PAIR>;
  constexpr int NST = Cfg::kStages;
  constexpr int PTS = Cfg::kPoints;
  constexpr uint32_t kStage = Cfg::kStage;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* w_full = bars;                // [NST] this CTA's (half) chunk has landed
  uint64_t* w_empty = bars + NST;         // [NST] the MMAs reading the stage are done (pair: signalled in both CTAs)
  uint64_t* peer_full = bars + 2 * NST;   // [NST] pair, leader only: the peer's half chunk has landed
  Bars B;
  B.acc_full = bars + 3 * NST;            // [2]
  B.a_q = bars + 3 * NST + 2;             // [5]  this CTA's epilogue warps
  uint64_t* a_qr = bars + 3 * NST + 7;    // [5]  pair, leader only: the peer CTA's quarter is in place (one relayed arrival)
  B.a_drain = bars + 3 * NST + 12;        //      bulk spill stores of the A image have read it (one arrival by the store thread)
  B.a_free = bars + 3 * NST + 13;         // [2]  accumulator half 1's MMAs have read A-image quarter 0 / 1 (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NST + 15);
  float* gskip = reinterpret_cast<float*>(smem + Cfg::kGskipOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // this instantiation can contain steps that spill by bulk store (see step_spills_by_bulk)
  constexpr bool kBulkK = SDF_BULK_SPILL && !PAR && ((SW & 1) != 0 || SDF_BULK_SPILL_TA);
  constexpr bool kSplitQ3 = SDF_SPLIT_Q3 == 2 || (SDF_SPLIT_Q3 == 1 && SW == 2);
#ifdef SDF_CLOCKS
  // effective SM clock of this launch: SM cycles and global-timer nanoseconds between the start and the end of CTA 0
  long long kclk0 = 0;
  unsigned long long kns0 = 0;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    kclk0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(kns0));
  }
#endif
  B.remote = PAIR && rank != 0;
  // tiles are walked two at a time by a pair: pair i works on tiles 2i (leader) and 2i + 1 (peer); the peer of a
  // ragged last pair works on the scratch tile n_tiles (all rows invalid, images inside the even-padded arrays)
  const int it0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int it_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // device-side point count (octree level loop: the size of a level is the result of the previous level's compaction and
  // never visits the host): lvl = {count, offset into keys / f_out / flag_out}; the grid was sized for p.n_tiles = capacity
  const int64_t n_pts = p.lvl ? p.lvl[0] : p.n;
  const int64_t lvl_off = p.lvl ? p.lvl[1] : 0;
  const int n_tiles_rt = p.lvl ? static_cast<int>((n_pts + Cfg::kPoints - 1) / Cfg::kPoints) : p.n_tiles;
  const int it_end = PAIR ? (n_tiles_rt + 1) / 2 : n_tiles_rt;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
      mbar_init(&peer_full[i], 1);
    }
    mbar_init(&B.acc_full[0], 1);
    mbar_init(&B.acc_full[1], 1);
    mbar_init(B.a_drain, 1);
    mbar_init(&B.a_free[0], 1);
    mbar_init(&B.a_free[1], 1);
    for (int j = 0; j < 5; ++j) {
      mbar_init(&B.a_q[j], kEpiWarps);
      mbar_init(&a_qr[j], 1);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) {
    if (PAIR) tmem_alloc2<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // barrier inits and the TMEM allocations of both CTAs are in place
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ib = img_bytes(p.H);

  // register file re-split (per SM sub-partition: 16384 registers = the CTA is launched with 640 x 96 registers; 128 x 56 + 512 x 104 fits that pool)
  if (warp >= kEpiWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SDF_REGS_AUX));
    if (warp == kProducerWarp) {
      // ------------------------------------------------------------------ weight producer (warp-uniform loop, one
      // elected lane issues the copies)
      {
        uint32_t stage = 0, round = 0;
        for (int it = it0; it < it_end; it += it_step) {
          const int tile = PAIR ? 2 * it + static_cast<int>(rank) : it;
          (void)tile;
          for (int s = 0; s < p.n_steps; ++s) {
            const StepDesc& st = p.steps[s];
            const int nchunks = st.n_groups * st.kchunks;
            const uint8_t* src = p.wpack + st.w_off;
            for (int c = 0; c < nchunks; ++c) {
              mbar_wait_vote(&w_empty[stage], (round & 1) ^ 1);
#ifdef SDF_CLOCKS
              if (p.flags & FL_DBG_NO_W) {
                if (elect_one()) mbar_arrive(&w_full[stage]);
              } else
#endif
              if (elect_one()) {
                const uint32_t bytes = PAIR ? st.chunk_bytes / 2 : st.chunk_bytes;
                mbar_arrive_expect_tx(&w_full[stage], bytes);
                bulk_g2s(sW + stage * kStage, src + static_cast<size_t>(c) * st.chunk_bytes + rank * bytes, bytes, &w_full[stage]);
              }
              __syncwarp();
              if (++stage == NST) { stage = 0; ++round; }
            }
          }
        }
      }
    } else if (warp == kMmaWarp && PAIR && rank != 0) {
      // ------------------------------------------------------------------ peer CTA: tell the leader when this CTA's
      // half of a weight chunk is in shared memory
      if (lane == 0) {
        uint32_t stage = 0, round = 0;
        for (int it = it0; it < it_end; it += it_step) {
          for (int s = 0; s < p.n_steps; ++s) {
            const int nchunks = p.steps[s].n_groups * p.steps[s].kchunks;
            for (int c = 0; c < nchunks; ++c) {
              mbar_wait(&w_full[stage], round & 1);
              mbar_arrive_cluster(&peer_full[stage], 0);
              if (++stage == NST) { stage = 0; ++round; }
            }
          }
        }
      }
    } else if (warp == kMmaWarp && !PAIR) {
      // ------------------------------------------------------------------ MMA issuer, single CTA.  ncu showed this
      // warp busy ~90 % of the time on loop bookkeeping (~1150 cycles per weight chunk against 512 cycles of MMA
      // work) while the epilogue warps waited for accumulators: the loop is the pacing resource of the kernel.  So:
      // step fields are copied to locals once per step, chunk / quarter indices advance by counters (no division),
      // descriptors advance by adds on their start-address field, barriers are addressed as integers, the loop stays
      // warp-uniform (uniform registers) and one elected lane issues.
      const uint32_t sA_d = smem_u32(sA) >> 4, sW_d = smem_u32(sW) >> 4;  // descriptor start-address units (16 B)
      const uint32_t wfull_u = smem_u32(w_full), wempty_u = smem_u32(w_empty);
      const uint32_t accfull_u = smem_u32(B.acc_full), aq_u = smem_u32(B.a_q), afree_u = smem_u32(B.a_free);
      uint32_t stage = 0, wpar = 0, n0 = 0;
      const bool poller = elect_one();  // (opaque to the compiler: a predicate it keeps instead of re-deriving the lane)
      for (int it = it0; it < it_end; it += it_step) {
        for (int s = 0; s < p.n_steps; ++s) {
          const int n_groups = p.steps[s].n_groups;
          if (n_groups == 0) continue;
          const int ng = p.steps[s].ng, kchunks = p.steps[s].kchunks, a_chunks = p.steps[s].a_chunks;
          const int k16 = p.steps[s].k16, k_lo = p.steps[s].k_lo;
          const uint32_t par = n0 & 1;
          const uint32_t idesc = make_idesc_f16(kTileRows, ng, DT == DT_BF16, false, false);
          const uint32_t b_lbo = static_cast<uint32_t>(ng) * 16;  // rows of B x 16 B
          const uint32_t db_step = (2 * b_lbo) >> 4;
          const uint32_t a_chunk_d = (static_cast<uint32_t>(k16 * 2) * kGranuleBytes) >> 4;
          const uint64_t da_hi = make_smem_desc(0, kGranuleBytes, 128);  // all fields but the start address
          const uint64_t db_hi = make_smem_desc(0, b_lbo, 128);
          if (SDF_ISSUER_UNROLL && k16 == 4 && kchunks == 8 && a_chunks == 8 && n_groups == 2) {
            // The hot shape (512 x 512 layer, fast mode): both column groups and their 8 weight chunks unrolled, so
            // chunk indices, accumulate flags and A-descriptor offsets are immediates; only the ring stage is state.
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint32_t tmem_d = tmem_base + g * ng;
#pragma unroll
              for (int kc = 0; kc < 8; ++kc) {
                if (g == 0) {
                  if (k_lo == 0) {
                    if (kc == 0) {
                      for (int j = 0; j < 4; ++j) mbar_wait_vote_u32(aq_u + 8 * j, par, poller);
                      tc_fence_after();
                    }
                  } else if ((kc & 1) == 0) {  // k_lo == 2: two weight chunks per 128-column quarter
                    mbar_wait_vote_u32(aq_u + 8 * (kc >> 1), par, poller);
                    tc_fence_after();
                  }
                }
                mbar_wait_vote_u32(wfull_u + 8 * stage, wpar, poller);
                tc_fence_after();
                const uint64_t da0 = da_hi | (sA_d + kc * ((4 * 2 * kGranuleBytes) >> 4));
                const uint64_t db0 = db_hi | (sW_d + stage * (kStage >> 4));
                if (elect_one()) {
#ifdef SDF_CLOCKS
                  if (!(p.flags & FL_DBG_NO_MMA))
#endif
                  {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_f16_ss(tmem_d, da0 + static_cast<uint64_t>(k) * (2 * kGranuleBytes >> 4), db0 + static_cast<uint64_t>(k) * db_step, idesc,
                                  (kc | k) != 0 ? 1u : 0u);
                  }
                  umma_commit_u32(wempty_u + 8 * stage);
                  if (g == 1 && (kc == 1 || kc == 3)) umma_commit_u32(afree_u + 8 * (kc >> 1));  // K quarter kc / 2 has been read
                  if (kc == 7) umma_commit_u32(accfull_u + 8 * g);
                }
                __syncwarp();
                if (++stage == NST) { stage = 0; wpar ^= 1; }
              }
            }
            ++n0;
            continue;
          }
          for (int g = 0; g < n_groups; ++g) {
            const uint32_t tmem_d = tmem_base + g * ng;
            uint32_t a_d = sA_d;    // start-address field of the current A chunk (A chunk of weight chunk kc = kc % a_chunks)
            int a_left = a_chunks;
            int q = 0, q_left = 0;  // group 0: next A-image quarter to wait for, weight chunks until then
            for (int kc = 0; kc < kchunks; ++kc) {
              if (g == 0) {
                // the A image becomes available quarter by quarter (hidden_step_staged)
                if (k_lo == 0) {
                  if (kc == 0) {
                    for (int j = 0; j < 4; ++j) mbar_wait_vote_u32(aq_u + 8 * j, par, poller);
                    tc_fence_after();
                  }
                } else if (q < 4 && q_left == 0) {
                  mbar_wait_vote_u32(aq_u + 8 * q, par, poller);
                  tc_fence_after();
                  ++q;
                  q_left = k_lo;
                }
                --q_left;
              }
              mbar_wait_vote_u32(wfull_u + 8 * stage, wpar, poller);
              tc_fence_after();
              // descriptors of the chunk's first K = 16 slice; the following slices only advance the 14-bit start
              // address fields (shared memory addresses stay below 2^18, so the adds never carry out of the field)
              const uint64_t da0 = da_hi | a_d;
              const uint64_t db0 = db_hi | (sW_d + stage * (kStage >> 4));
              if (elect_one()) {
#ifdef SDF_CLOCKS
                if (!(p.flags & FL_DBG_NO_MMA))
#endif
                {
                  if (k16 == 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_f16_ss(tmem_d, da0 + static_cast<uint64_t>(k) * (2 * kGranuleBytes >> 4), db0 + static_cast<uint64_t>(k) * db_step, idesc,
                                  (kc | k) != 0 ? 1u : 0u);
                  } else {
                    for (int k = 0; k < k16; ++k)
                      umma_f16_ss(tmem_d, da0 + static_cast<uint64_t>(k) * (2 * kGranuleBytes >> 4), db0 + static_cast<uint64_t>(k) * db_step, idesc,
                                  (kc | k) != 0 ? 1u : 0u);
                  }
                }
                umma_commit_u32(wempty_u + 8 * stage);
                if (kc + 1 == kchunks) {
                  if (g == 1) { umma_commit_u32(afree_u); umma_commit_u32(afree_u + 8); }  // other shapes: with the group's last chunk
                  umma_commit_u32(accfull_u + 8 * g);
                }
              }
              __syncwarp();
              a_d += a_chunk_d;
              if (--a_left == 0) { a_d = sA_d; a_left = a_chunks; }
              if (++stage == NST) { stage = 0; wpar ^= 1; }
            }
          }
          ++n0;
        }
      }
    } else if (PAIR && warp == kMmaWarp + 1 && rank != 0) {
      // ------------------------------------------------------------------ peer CTA: quarter-barrier relay.  The peer's
      // 16 epilogue warps arrive on their own CTA's barriers (a release.cluster arrive is a MEMBAR + ERRBAR in every
      // warp: 10 % of the pair kernel's samples); this otherwise idle warp forwards each completed quarter to the
      // leader with one remote arrival.
      if (lane == 0) {
        uint32_t n0 = 0;
        for (int it = it0; it < it_end; it += it_step) {
          for (int s = 0; s < p.n_steps; ++s) {
            if (p.steps[s].n_groups == 0) continue;
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
              const int j = jj < 3 ? jj : (jj == 3 ? 4 : 3);  // completion order: quarters 0, 1, 2, even slices of 3, quarter 3
              mbar_wait(&B.a_q[j], n0 & 1);
              tc_fence_after();
              tc_fence_before();
              mbar_arrive_cluster(&a_qr[j], 0);
            }
            ++n0;
          }
        }
      }
    } else if (kBulkK && warp == kMmaWarp + 2) {
      // ------------------------------------------------------------------ spill-store thread (every CTA): one bulk copy
      // shared -> global per completed quarter (32 KB) of the A image of every step that spills it (step_spills_by_bulk).
      // The quarter barriers a_q[j] the epilogue warps arrive on (after their fence.proxy.async) are the trigger; a_drain
      // tells the epilogue when quarter j may be overwritten again (the copy has read it).  Images are read back by LATER
      // launches (kernel boundary) except the forward activations, which the reverse sweep of the same launch loads again:
      // see out_full below.
      if (lane == 0) {
        const uint32_t sA_u = smem_u32(sA);
        const uint32_t qbytes = ib / 4;
        // Every GEMM step consumes one "production" of the A image (the tile prologue's or the previous step's), announced
        // on a_q[0..3] with the parity of the GEMM's index.  This thread observes EVERY production in order -- a parity
        // wait that skipped one would alias with the phase two back -- and copies out those whose step spills by bulk.
        uint32_t n0 = 0;
        bool any = false;  // inference programs spill nothing: this thread has nothing to pace
        for (int s = 0; s < p.n_steps; ++s) any = any || step_spills_by_bulk<PAR>(p, p.steps[s]);
        for (int it = it0; any && it < it_end; it += it_step) {
          const int tile = PAIR ? 2 * it + static_cast<int>(rank) : it;
          uint8_t* out = nullptr;  // spill image of the production the next GEMM step consumes (nullptr: none)
          bool out_full = false;
          for (int s = 0; s < p.n_steps; ++s) {
            const StepDesc& st = p.steps[s];
            if (st.n_groups != 0) {
              const uint32_t par = n0 & 1;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                mbar_wait(&B.a_q[j], par);
                if (out) {
                  if (SW == 2) bulk_s2g_hint(out + j * qbytes, sA_u + j * qbytes, qbytes, kPolicyEvictFirst);
                  else bulk_s2g_u32(out + j * qbytes, sA_u + j * qbytes, qbytes);
                  bulk_commit();
                }
              }
              if (out) {
                if (out_full) {
                  // the last activation image before the output layer: the reverse sweep of this launch starts to read the
                  // forward sweep's images back with generic loads -- wait until every bulk store issued so far is COMPLETE
                  // (written and visible), fence the proxies, then release (the drain wait of the epilogue acquires)
                  bulk_wait<0>();
                  fence_proxy_async_all();
                } else {
                  bulk_wait_read<0>();  // shared memory may be overwritten as soon as the copies have READ it
                }
                mbar_arrive(B.a_drain);
              }
              ++n0;
            }
            out = step_spills_by_bulk<PAR>(p, st) ? step_spill_image(p, st, tile, ib) : nullptr;
            out_full = out && st.epi == EPI_FWD_HIDDEN && s + 1 < p.n_steps &&
                       (p.steps[s + 1].epi == EPI_FWD_FUSED || p.steps[s + 1].epi == EPI_FWD_LAST);
          }
        }
        bulk_wait<0>();
      }
    } else if (warp == kMmaWarp) {
      // ------------------------------------------------------------------ MMA issuer, CTA pair (leader CTA), lean form
      // of the loop above: M = 256 MMAs over both CTAs' tiles, 8 per weight chunk (K = 128), so the issue cost per
      // FLOP and the L2 -> SM weight bytes per SM are half of the single-CTA kernel's and a ring stage holds twice the
      // tensor time.
      const uint32_t sA_d = smem_u32(sA) >> 4, sW_d = smem_u32(sW) >> 4;
      const uint32_t wfull_u = smem_u32(w_full), wempty_u = smem_u32(w_empty), pfull_u = smem_u32(peer_full);
      const uint32_t accfull_u = smem_u32(B.acc_full), aq_u = smem_u32(B.a_q), aqr_u = smem_u32(a_qr), afree_u = smem_u32(B.a_free);
      uint32_t stage = 0, wpar = 0, n0 = 0;
      const bool poller = elect_one();
      for (int it = it0; it < it_end; it += it_step) {
        for (int s = 0; s < p.n_steps; ++s) {
          const int n_groups = p.steps[s].n_groups;
          if (n_groups == 0) continue;
          const int ng = p.steps[s].ng, kchunks = p.steps[s].kchunks, a_chunks = p.steps[s].a_chunks;
          const int k16 = p.steps[s].k16, k_lo = p.steps[s].k_lo;
          const uint32_t par = n0 & 1;
          const uint32_t idesc = make_idesc_f16(2 * kTileRows, ng, DT == DT_BF16, false, false);
          const uint32_t b_lbo = static_cast<uint32_t>(ng / 2) * 16;  // rows of B in this CTA x 16 B
          const uint32_t db_step = (2 * b_lbo) >> 4;
          const uint32_t a_chunk_d = (static_cast<uint32_t>(k16 * 2) * kGranuleBytes) >> 4;
          const uint64_t da_hi = make_smem_desc(0, kGranuleBytes, 128);
          const uint64_t db_hi = make_smem_desc(0, b_lbo, 128);
          if (k16 == 8 && kchunks == 4 && a_chunks == 4 && n_groups == 2 && k_lo <= 1) {
            // hot shape (512 x 512 layer): 2 column groups x 4 weight chunks, unrolled
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const uint32_t tmem_d = tmem_base + g * ng;
#pragma unroll
              for (int kc = 0; kc < 4; ++kc) {
                if (g == 0) {
                  if (k_lo == 0) {
                    if (kc == 0) {
                      for (int j = 0; j < 4; ++j) { mbar_wait_vote_u32(aq_u + 8 * j, par, poller); mbar_wait_cluster_vote_u32(aqr_u + 8 * j, par, poller); }
                      tc_fence_after();
                    }
                  } else {  // one weight chunk per 128-column quarter of the A images
                    const int jq = (kSplitQ3 && kc == 3) ? 4 : kc;  // last quarter: its even 16-column slices first
                    mbar_wait_vote_u32(aq_u + 8 * jq, par, poller);
                    mbar_wait_cluster_vote_u32(aqr_u + 8 * jq, par, poller);
                    tc_fence_after();
                  }
                }
                mbar_wait_vote_u32(wfull_u + 8 * stage, wpar, poller);
                mbar_wait_cluster_vote_u32(pfull_u + 8 * stage, wpar, poller);
                tc_fence_after();
                const uint64_t da0 = da_hi | (sA_d + kc * ((8 * 2 * kGranuleBytes) >> 4));
                const uint64_t db0 = db_hi | (sW_d + stage * (kStage >> 4));
                if (kSplitQ3 && g == 0 && kc == 3 && k_lo != 0) {
                  // even K = 16 slices now, odd slices when the whole quarter has arrived
                  if (elect_one()) {
#ifdef SDF_CLOCKS
                    if (!(p.flags & FL_DBG_NO_MMA))
#endif
#pragma unroll
                    for (int k = 0; k < 8; k += 2)
                      umma_f16_ss2(tmem_d, da0 + static_cast<uint64_t>(k) * (2 * kGranuleBytes >> 4), db0 + static_cast<uint64_t>(k) * db_step, idesc, 1u);
                  }
                  __syncwarp();
                  mbar_wait_vote_u32(aq_u + 8 * 3, par, poller);
                  mbar_wait_cluster_vote_u32(aqr_u + 8 * 3, par, poller);
                  tc_fence_after();
                  if (elect_one()) {
#ifdef SDF_CLOCKS
                    if (!(p.flags & FL_DBG_NO_MMA))
#endif
#pragma unroll
                    for (int k = 1; k < 8; k += 2)
                      umma_f16_ss2(tmem_d, da0 + static_cast<uint64_t>(k) * (2 * kGranuleBytes >> 4), db0 + static_cast<uint64_t>(k) * db_step, idesc, 1u);
                    umma_commit2_u32(wempty_u + 8 * stage, 3);
                    umma_commit2_u32(accfull_u + 8 * g, 3);
                  }
                  __syncwarp();
                  if (++stage == NST) { stage = 0; wpar ^= 1; }
                  continue;
                }
                if (elect_one()) {
#ifdef SDF_CLOCKS
                  if (!(p.flags & FL_DBG_NO_MMA))
#endif
#pragma unroll
                  for (int k = 0; k < 8; ++k)
                    umma_f16_ss2(tmem_d, da0 + static_cast<uint64_t>(k) * (2 * kGranuleBytes >> 4), db0 + static_cast<uint64_t>(k) * db_step, idesc,
                                 (kc | k) != 0 ? 1u : 0u);
                  umma_commit2_u32(wempty_u + 8 * stage, 3);
                  if (g == 1 && kc < 2) umma_commit2_u32(afree_u + 8 * kc, 3);  // K quarter kc of both A images has been read
                  if (kc == 3) umma_commit2_u32(accfull_u + 8 * g, 3);
                }
                __syncwarp();
                if (++stage == NST) { stage = 0; wpar ^= 1; }
              }
            }
            ++n0;
            continue;
          }
          for (int g = 0; g < n_groups; ++g) {
            const uint32_t tmem_d = tmem_base + g * ng;
            uint32_t a_d = sA_d;
            int a_left = a_chunks;
            int q = 0, q_left = 0;
            for (int kc = 0; kc < kchunks; ++kc) {
              if (g == 0) {
                if (k_lo == 0) {
                  if (kc == 0) {
                    for (int j = 0; j < 4; ++j) { mbar_wait_vote_u32(aq_u + 8 * j, par, poller); mbar_wait_cluster_vote_u32(aqr_u + 8 * j, par, poller); }
                    tc_fence_after();
                  }
                } else if (q < 4 && q_left == 0) {
                  mbar_wait_vote_u32(aq_u + 8 * q, par, poller);
                  mbar_wait_cluster_vote_u32(aqr_u + 8 * q, par, poller);
                  tc_fence_after();
                  ++q;
                  q_left = k_lo;
                }
                --q_left;
              }
              mbar_wait_vote_u32(wfull_u + 8 * stage, wpar, poller);
              mbar_wait_cluster_vote_u32(pfull_u + 8 * stage, wpar, poller);
              tc_fence_after();
              const uint64_t da0 = da_hi | a_d;
              const uint64_t db0 = db_hi | (sW_d + stage * (kStage >> 4));
              if (elect_one()) {
                for (int k = 0; k < k16; ++k)
                  umma_f16_ss2(tmem_d, da0 + static_cast<uint64_t>(k) * (2 * kGranuleBytes >> 4), db0 + static_cast<uint64_t>(k) * db_step, idesc,
                               (kc | k) != 0 ? 1u : 0u);
                umma_commit2_u32(wempty_u + 8 * stage, 3);
                if (kc + 1 == kchunks) {
                  if (g == 1) { umma_commit2_u32(afree_u, 3); umma_commit2_u32(afree_u + 8, 3); }
                  umma_commit2_u32(accfull_u + 8 * g, 3);
                }
              }
              __syncwarp();
              a_d += a_chunk_d;
              if (--a_left == 0) { a_d = sA_d; a_left = a_chunks; }
              if (++stage == NST) { stage = 0; wpar ^= 1; }
            }
          }
          ++n0;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SDF_REGS_EPI));
    Epi e;
    e.sA = smem_u32(sA);
    const int q = warp & 3;
    e.h = warp >> 2;
    e.math = !PAR || q < 2;
    e.prow = PAR ? (q & 1) * 32 + lane : q * 32 + lane;
    e.scr = reinterpret_cast<float*>(smem + Cfg::kScratchOff) + e.h * 32 * kParityPoints;
    e.taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    e.fval = 0.f;
    e.fb = 0.f;
#ifdef SDF_CLOCKS
    long long clk_local[64];
    for (int i = 0; i < 64; ++i) clk_local[i] = 0;
    e.clk = clk_local;
#endif
    uint32_t n0 = 0, n1 = 0;  // GEMM steps seen / GEMM steps with two column groups seen (barrier phases)
    EpiConst k;
    k.H = p.H;
    k.beta = p.beta;
    k.nbl2 = -p.beta * 1.4426950408889634f;
    k.ln2_over_beta = 0.6931471805599453f * p.inv_beta;
    k.relu = (p.flags & FL_RELU) != 0;
#ifdef SDF_CLOCKS
    k.no_math = (p.flags & FL_DBG_NO_MATH) != 0;
    k.no_mufu = (p.flags & FL_DBG_NO_MUFU) != 0;
    k.no_aux = (p.flags & FL_DBG_NO_AUX) != 0;
    const bool no_stg = (p.flags & FL_DBG_NO_STG) != 0;
#else
    constexpr bool no_stg = false;
#endif
    {
      // log1p(x) = x P(x) on [0,1], cubic P, |error| < 7.1e-5 (6.4e-4 when evaluated in fp16): after the 1/beta scale
      // an order of magnitude under the 16-bit rounding of the stored activation
      const float c[4] = {0.99744897f, -0.4713013f, 0.22568572f, -0.05875714f};
#pragma unroll
      for (int i = 0; i < 4; ++i) k.sp2[i] = H16<DT>::pack(c[i] * p.inv_beta, c[i] * p.inv_beta);
    }
    const bool relu = k.relu;
    const float mult = p.seed_mult ? __ldg(p.seed_mult) : 1.f;
    // layer whose image carries the raw input in its last three columns (feeds the skip layer): those columns have no
    // activation state, their derivative factor must read as 0
    const int inj_layer = p.skip > 0 ? p.skip - 1 : -1;
    const bool has_u = (p.flags & FL_TANGENT) && !relu;

    // bulk spill stores: a_drain parity still to be waited for before the next write to the A image (-1: none), number
    // of bulk-spilled images so far
    int drain = -1;
    uint32_t n_bulk = 0;
    for (int it = it0; it < it_end; it += it_step) {
          const int tile = PAIR ? 2 * it + static_cast<int>(rank) : it;
          (void)tile;
      e.grow = static_cast<int64_t>(tile) * PTS + e.prow;
      e.valid = e.grow < n_pts;
      // ---------------------------------------------------------------- prologue: input block (column groups 0, 1
      // write its two granules)
      const int first = p.steps[0].epi;
      e.x[0] = e.x[1] = e.x[2] = 0.f;
      drain_wait(B, drain);  // the input block goes into quarter 0 of the A image
      drain = -1;
      const bool fwd_program = SW == 1 || (SW == 3 && first == EPI_FWD_HIDDEN);
      if (fwd_program) {
        {
          if (e.valid) {
            if (p.src == SRC_POINTS) {
              e.x[0] = __ldg(p.x + e.grow * 3 + 0);
              e.x[1] = __ldg(p.x + e.grow * 3 + 1);
              e.x[2] = __ldg(p.x + e.grow * 3 + 2);
            } else if (p.src == SRC_GRID) {
              // reference executor/executor.py:351-356: coords = stack(meshgrid(x,x,x,'ij'),-1).reshape(-1,3)
              const int64_t idx = p.grid_off + e.grow;
              const int64_t nn = p.grid_n;
              const int iz = static_cast<int>(idx % nn);
              const int iy = static_cast<int>((idx / nn) % nn);
              const int ix = static_cast<int>(idx / (nn * nn));
              e.x[0] = __ldg(p.axis + ix);
              e.x[1] = __ldg(p.axis + iy);
              e.x[2] = __ldg(p.axis + iz);
            } else {
              // octree cell centre from a 3-bit-per-level Morton key: c = -1 + (2 i + 1) * 2^-level
              const uint64_t key = __ldg(p.keys + lvl_off + e.grow);
              uint32_t ci[3] = {0, 0, 0};
              for (int b = 0; b < p.level; ++b) {
                ci[0] |= static_cast<uint32_t>((key >> (3 * b + 2)) & 1) << b;
                ci[1] |= static_cast<uint32_t>((key >> (3 * b + 1)) & 1) << b;
                ci[2] |= static_cast<uint32_t>((key >> (3 * b + 0)) & 1) << b;
              }
              const float hs = ldexpf(1.f, -p.level);
#pragma unroll
              for (int j = 0; j < 3; ++j) e.x[j] = -1.f + (2.f * static_cast<float>(ci[j]) + 1.f) * hs;
            }
          }
          if (e.math && e.h < 2) {
            const uint4 gx = xblock_granule<DT>(e.x, e.h, 1.f);
            put_block<PAR>(e, e.h, gx, true,
                           (p.flags & FL_SPILL_X) ? p.Ximg + tile * static_cast<size_t>(2 * kGranuleBytes) : nullptr);
          }
        }
      } else {
        // tangent / adjoint program: seeds
        {
          if (e.valid) {
            e.fb = p.fbar ? __ldg(p.fbar + e.grow) * mult : 0.f;
            if (p.gbar) {
              e.x[0] = __ldg(p.gbar + e.grow * 3 + 0) * mult;
              e.x[1] = __ldg(p.gbar + e.grow * 3 + 1) * mult;
              e.x[2] = __ldg(p.gbar + e.grow * 3 + 2) * mult;
            }
          } else {
            e.fb = 0.f;
          }
          if (e.math && e.h < 2) {
            const uint4 gg = xblock_granule<DT>(e.x, e.h, 0.f);
            put_block<PAR>(e, e.h, gg, true, p.Gimg + tile * static_cast<size_t>(2 * kGranuleBytes));
            // f-bar block: cols 0-2 = fbar hi/lo/lolo, col 15 = 1
            float fh, fl, fll;
            split3<DT>(e.fb, fh, fl, fll);
            float sv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (e.h == 0) { sv[0] = fh; sv[1] = fl; sv[2] = fll; } else { sv[7] = e.valid ? 1.f : 0.f; }
            put_block<PAR>(e, e.h, pack8<DT>(sv), false, p.Simg + tile * static_cast<size_t>(2 * kGranuleBytes));
          }
        }
      }
      fence_proxy_async_smem();
      if (p.steps[0].n_groups != 0) {
        tc_fence_before();
        arrive_all_quarters(B);
      }

      for (int s = 0; s < p.n_steps; ++s) {
        const StepDesc& st = p.steps[s];
        const int l = st.layer;
        StepIO io;
        io.in_a = io.in_b = nullptr;
        io.out = io.out_u = nullptr;
        io.vec = nullptr;
        io.special = false;
        io.next_a = io.next_b = nullptr;
        io.cold = true;
        io.wl = nullptr;
        // this step's spill image leaves through the store thread's bulk copies of the A image (no per-thread stores)
        const bool bulk = kBulkK && step_spills_by_bulk<PAR>(p, st);
        io.drain = kBulkK ? drain : -1;
        // every step but the two N = 16 output steps and the forward-only fused end writes the A image (and waits first)
        if (!(st.epi == EPI_FWD_LAST || st.epi == EPI_REV_LAST || (st.epi == EPI_FWD_FUSED && st.signal == 0))) drain = -1;
#if SDF_TA_PF_DIST > 0
        if ((SW & 2) && (st.epi == EPI_TAN_HIDDEN || st.epi == EPI_ADJ_HIDDEN || st.epi == EPI_TAN_LAST)) {
          // images the next step will read (just-in-time prefetch hands over from step to step)
          if (s + 1 < p.n_steps) {
            const StepDesc& nx = p.steps[s + 1];
            if (nx.epi == EPI_TAN_HIDDEN || nx.epi == EPI_TAN_LAST) {
              io.next_a = img_ptr(p.Aimg, nx.layer, p.tile_stride, tile, ib);
              io.next_b = relu ? nullptr : img_ptr(p.Qimg, nx.layer, p.tile_stride, tile, ib);
            } else if (nx.epi == EPI_ADJ_HIDDEN) {
              io.next_a = img_ptr(p.Aimg, nx.layer, p.tile_stride, tile, ib);
              io.next_b = has_u ? img_ptr(p.Uimg, nx.layer, p.tile_stride, tile, ib) : nullptr;
            }
          }
          io.cold = !(s > 0 && (p.steps[s - 1].epi == EPI_TAN_HIDDEN || p.steps[s - 1].epi == EPI_ADJ_HIDDEN || p.steps[s - 1].epi == EPI_TAN_LAST));
        }
#endif
#if SDF_REV_PF_DIST > 0
        if ((SW & 1) && st.epi == EPI_REV_HIDDEN && s + 1 < p.n_steps && p.steps[s + 1].epi == EPI_REV_HIDDEN)
          io.next_a = img_ptr(p.Aimg, p.steps[s + 1].layer - 1, p.tile_stride, tile, ib);
#endif
        switch (st.epi) {
          case EPI_FWD_HIDDEN: if (SW & 1) {
            io.out = ((p.flags & FL_SPILL_A) && !no_stg && !bulk) ? img_ptr(p.Aimg, l, p.tile_stride, tile, ib) : nullptr;
            io.vec = p.bias + l * p.H;
            io.special = l == inj_layer;
            hidden_step<DT, PAR, EPI_FWD_HIDDEN>(e, io, k, st, B, n0, n1, gskip);
          } break;
          case EPI_FWD_FUSED: if (SW & 1) {
            // last hidden layer + output layer (+ reverse start): see sdf_internal.h
            io.out = (st.signal && (p.flags & FL_SPILL_Q) && !no_stg && !bulk) ? img_ptr(p.Qimg, l, p.tile_stride, tile, ib) : nullptr;
            io.out_u = ((p.flags & FL_SPILL_A) && !no_stg) ? img_ptr(p.Aimg, l, p.tile_stride, tile, ib) : nullptr;
            io.vec = p.bias + l * p.H;
            io.wl = p.w_last;
            hidden_step<DT, PAR, EPI_FWD_FUSED>(e, io, k, st, B, n0, n1, gskip);
            // f = sum of the four column-quarter partial dot products of a row + b_L (fixed order: deterministic)
            if (e.h > 0) gskip[(e.h - 1) * kTileRows + e.prow] = e.fval;
            named_bar_sync(1, kEpiThreads);
            if (e.h == 0) {
              const float f = ((e.fval + gskip[e.prow]) + gskip[kTileRows + e.prow]) + gskip[2 * kTileRows + e.prow] + __ldg(p.bias + (l + 1) * p.H);
              e.fval = f;
              if (e.valid && p.f_out) p.f_out[lvl_off + e.grow] = f;
              if (e.valid && p.flag_out) p.flag_out[lvl_off + e.grow] = (p.flag_all || fabsf(f) <= p.flag_thresh) ? 1 : 0;
            }
            named_bar_sync(1, kEpiThreads);  // the partials live in the reverse sweep's skip buffer
          } break;
          case EPI_FWD_LAST: if (SW & 1) {
            mbar_wait(&B.acc_full[0], n0 & 1);
            tc_fence_after();
            ++n0;
            if (e.h == 0) {
              float v[32];
              load_acc32<PAR>(e, 0, v);
              if (e.math) {
                const float z = v[0] + __ldg(p.bias + l * p.H);
                // reference model/networks.py:104-105: tanh on the output when the activation is ReLU
                e.fval = relu ? tanhf(z) : z;
                if (e.valid && p.f_out) p.f_out[lvl_off + e.grow] = e.fval;
                // octree refinement predicate fused into the last forward epilogue (one fp32 compare, as
                // oracle/sdf_oracle.py:octree_flags): |f| <= band * h_level, or every cell below the uniform level
                if (e.valid && p.flag_out) p.flag_out[lvl_off + e.grow] = (p.flag_all || fabsf(e.fval) <= p.flag_thresh) ? 1 : 0;
              }
            }
          } break;
          case EPI_REV_START: if (SW & 1) {
            // q_{L-1} = w_L * act'(z_{L-1}); the A image still holds a_{L-1}
            uint8_t* qimg = (p.flags & FL_SPILL_Q) ? img_ptr(p.Qimg, l, p.tile_stride, tile, ib) : nullptr;
            drain_wait(B, io.drain);
            if (e.math) {
              for (int c0 = e.h * 32; c0 < k.H; c0 += 128) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int c = c0 + 8 * j;
                  float a[8];
                  unpack8<DT>(ld_smem_v4(e.sA + img_off(c, e.prow)), a);
                  if (PAR) unpack8_add<DT>(ld_smem_v4(e.sA + img_off(c, e.prow + kParityPoints)), a);
                  const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_last + c));
                  const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_last + c + 4));
                  const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                  for (int t = 0; t < 8; ++t) a[t] = w[t] * (1.f - one_minus_dact(a[t], k));
                  put_granule8<DT, PAR>(e, qimg, c, a);
                }
              }
            }
            fence_proxy_async_smem();
            if (st.signal) {
              tc_fence_before();
              arrive_all_quarters(B);
            }
          } break;
          case EPI_REV_HIDDEN: if (SW & 1) {
            // GEMM r = q_l W_l done; q_{l-1} = r * act'(z_{l-1});  l = st.layer
            io.in_a = img_ptr(p.Aimg, l - 1, p.tile_stride, tile, ib);
            io.out = ((p.flags & FL_SPILL_Q) && !no_stg && !bulk) ? img_ptr(p.Qimg, l - 1, p.tile_stride, tile, ib) : nullptr;
            io.special = (l - 1) == inj_layer;
            hidden_step<DT, PAR, EPI_REV_HIDDEN>(e, io, k, st, B, n0, n1, gskip);
          } break;
          case EPI_REV_LAST: if (SW & 1) {
            mbar_wait(&B.acc_full[0], n0 & 1);
            tc_fence_after();
            ++n0;
            named_bar_sync(1, kEpiThreads);  // gskip written by the last column group in an earlier step
            if (e.h == 0) {
              float v[32];
              load_acc32<PAR>(e, 0, v);
              if (e.math) {
                float g[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) g[j] = v[j] + (p.skip > 0 ? gskip[e.prow * 3 + j] : 0.f);
                if (relu) {
                  const float d = 1.f - e.fval * e.fval;  // d tanh
#pragma unroll
                  for (int j = 0; j < 3; ++j) g[j] *= d;
                }
                if (e.valid && p.g_out) {
                  p.g_out[e.grow * 3 + 0] = g[0]; p.g_out[e.grow * 3 + 1] = g[1]; p.g_out[e.grow * 3 + 2] = g[2];
                }
              }
            }
            named_bar_sync(1, kEpiThreads);  // gskip may be rewritten by the next tile
          } break;
          case EPI_TAN_HIDDEN: if (SW & 2) {
            io.in_a = img_ptr(p.Aimg, l, p.tile_stride, tile, ib);
            io.in_b = relu ? nullptr : img_ptr(p.Qimg, l, p.tile_stride, tile, ib);
            io.out = (no_stg || bulk) ? nullptr : img_ptr(p.ADimg, l, p.tile_stride, tile, ib);
            io.out_u = relu ? nullptr : img_ptr(p.Uimg, l, p.tile_stride, tile, ib);
            io.special = l == inj_layer;
            hidden_step<DT, PAR, EPI_TAN_HIDDEN>(e, io, k, st, B, n0, n1, gskip);
          } break;
          case EPI_TAN_LAST: if (SW & 2) {
            // last tangent layer + adjoint start in one epilogue: u_{L-1} is consumed where it is produced
            io.in_a = img_ptr(p.Aimg, l, p.tile_stride, tile, ib);
            io.in_b = relu ? nullptr : img_ptr(p.Qimg, l, p.tile_stride, tile, ib);
            io.out = (no_stg || bulk) ? nullptr : img_ptr(p.Zimg, l, p.tile_stride, tile, ib);
            io.out_u = img_ptr(p.ADimg, l, p.tile_stride, tile, ib);
            io.vec = p.w_last;
            hidden_step<DT, PAR, EPI_TAN_LAST>(e, io, k, st, B, n0, n1, gskip);
          } break;
          case EPI_ADJ_START: if (SW & 2) {
            // Zbar_{L-1} = s' * (w_L fbar) + u_{L-1}
            const uint8_t* aimg = img_ptr(p.Aimg, l, p.tile_stride, tile, ib);
            const uint8_t* uimg = has_u ? img_ptr(p.Uimg, l, p.tile_stride, tile, ib) : nullptr;
            uint8_t* zimg = img_ptr(p.Zimg, l, p.tile_stride, tile, ib);
            drain_wait(B, io.drain);
            if (e.math) {
              for (int c0 = e.h * 32; c0 < k.H; c0 += 128) {
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                  Piece<PAR> xa, xu;
                  ldg_piece<PAR>(e, c0 + 16 * sub, aimg, xa);
                  if (has_u) ldg_piece<PAR>(e, c0 + 16 * sub, uimg, xu);
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const int c = c0 + 16 * sub + 8 * j;
                    float a[8];
                    granule8<DT, PAR>(xa, j, a);
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w_last + c));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w_last + c + 4));
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int t = 0; t < 8; ++t) a[t] = (1.f - one_minus_dact(a[t], k)) * (w[t] * e.fb);
                    if (has_u) {
                      float u[8];
                      granule8<DT, PAR>(xu, j, u);
#pragma unroll
                      for (int t = 0; t < 8; ++t) a[t] += u[t];
                    }
                    put_granule8<DT, PAR>(e, zimg, c, a);
                  }
                }
              }
            }
            fence_proxy_async_smem();
            if (st.signal) {
              tc_fence_before();
              arrive_all_quarters(B);
            }
          } break;
          case EPI_ADJ_HIDDEN: if (SW & 2) {
            io.in_a = img_ptr(p.Aimg, l, p.tile_stride, tile, ib);
            io.in_b = has_u ? img_ptr(p.Uimg, l, p.tile_stride, tile, ib) : nullptr;
            io.out = (no_stg || bulk) ? nullptr : img_ptr(p.Zimg, l, p.tile_stride, tile, ib);
            io.special = l == inj_layer;
            hidden_step<DT, PAR, EPI_ADJ_HIDDEN>(e, io, k, st, B, n0, n1, gskip);
          } break;
          default: break;
        }
        if (bulk) {
          // the A image this step has left is being copied out
          drain = static_cast<int>(n_bulk & 1);
          ++n_bulk;
        }
      }
    }
#ifdef SDF_CLOCKS
    if (threadIdx.x == 0 && p.dbg_clk)
      for (int i = 0; i < 64; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg_clk) + i, static_cast<unsigned long long>(clk_local[i]));
#endif
  }

#ifdef SDF_CLOCKS
  if (threadIdx.x == 0 && blockIdx.x == 0 && p.dbg_clk) {
    unsigned long long kns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(kns1));
    p.dbg_clk[70] = clock64() - kclk0;
    p.dbg_clk[71] = static_cast<long long>(kns1 - kns0);
  }
#endif
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's shared memory and TMEM stay valid until both CTAs are done
  else __syncthreads();
  if (warp == kMmaWarp) {
    if (PAIR) tmem_dealloc2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int DT, bool PAR, bool PAIR, int SW = 3>
static cudaError_t launch_one(const ChainParams& p, int num_sms, cudaStream_t stream) {
  // the opt-in to > 48 KB of dynamic shared memory is a per-device (per-context) function attribute: one flag per device
  // ordinal, so a second device of the same process (nn.DataParallel, net.engine(device=1)) gets it too
  static bool attr_set[kMaxDevices] = {};
  auto kfn = chain_kernel<DT, PAR, PAIR, SW>;
  constexpr uint32_t smem = ChainCfg<PAR, PAIR>::kSmem;
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (PAIR) {
    const int pairs_needed = (p.n_tiles + 1) / 2;
    const int pairs = pairs_needed < num_sms / 2 ? pairs_needed : num_sms / 2;
    cfg.gridDim = dim3(2 * pairs);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(p.n_tiles < num_sms ? p.n_tiles : num_sms);
  }
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
#ifdef SDF_CLOCKS
  {
    // measurement build: synchronous launch, per-kind cycle report of epilogue thread 0 (averaged over CTAs and steps)
    const int grid = static_cast<int>(cfg.gridDim.x);
    ChainParams q = p;
    cudaMalloc(&q.dbg_clk, sizeof(long long) * 72);
    cudaMemsetAsync(q.dbg_clk, 0, sizeof(long long) * 72, stream);
    cudaLaunchKernelEx(&cfg, kfn, q);
    cudaStreamSynchronize(stream);
    long long hb[72];
    cudaMemcpy(hb, q.dbg_clk, sizeof(hb), cudaMemcpyDeviceToHost);
    cudaFree(q.dbg_clk);
    static const char* names[8] = {"FWD", "FWD_LAST", "REV_START", "REV", "REV_LAST", "TAN", "ADJ_START", "ADJ"};
    for (int kind = 0; kind < 8; ++kind) {
      const double n = static_cast<double>(hb[kind * 8 + 5]);
      if (n <= 0) continue;
      fprintf(stderr, "clk %-4s steps/CTA %.0f: wait_acc0 %.0f  half0 %.0f  wait_acc1 %.0f  dump %.0f  half1 %.0f\n", names[kind], n / grid,
              hb[kind * 8 + 0] / n, hb[kind * 8 + 1] / n, hb[kind * 8 + 2] / n, hb[kind * 8 + 3] / n, hb[kind * 8 + 4] / n);
    }
    if (hb[54] > 0) fprintf(stderr, "clk DRAIN wait at the dump (thread 0): %.0f cycles per wait, %.0f waits per CTA\n", hb[48] / (double)hb[54], hb[54] / (double)grid);
    if (hb[71] > 0)
      fprintf(stderr, "clk KERNEL CTA 0: %.3f M cycles in %.3f ms = %.0f MHz effective SM clock\n", hb[70] / 1e6, hb[71] / 1e6, 1e3 * hb[70] / static_cast<double>(hb[71]));
    if (hb[68] > 0) {
      const double n = static_cast<double>(hb[68]);
      fprintf(stderr, "clk ISSUER per GEMM step: wait_a_lo %.0f  wait_a_hi %.0f  wait_weights %.0f  total %.0f\n", hb[64] / n, hb[65] / n, hb[66] / n, hb[67] / n);
    }
    return cudaGetLastError();
  }
#endif
  cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, p);
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_chain(const ChainParams& p, int dtype, int parity, int pair, int num_sms, cudaStream_t stream) {
  if (p.n_tiles <= 0) return cudaSuccess;
  // program family of this launch (api.cu builds forward [+ reverse] programs or tangent / adjoint programs)
  bool fr = false, ta = false;
  for (int s = 0; s < p.n_steps; ++s) ((p.steps[s].epi <= EPI_REV_LAST || p.steps[s].epi == EPI_FWD_FUSED) ? fr : ta) = true;
  if (parity) {
    if (fr && !ta) return dtype == DT_BF16 ? launch_one<DT_BF16, true, false, 1>(p, num_sms, stream) : launch_one<DT_FP16, true, false, 1>(p, num_sms, stream);
    if (ta && !fr) return dtype == DT_BF16 ? launch_one<DT_BF16, true, false, 2>(p, num_sms, stream) : launch_one<DT_FP16, true, false, 2>(p, num_sms, stream);
    return dtype == DT_BF16 ? launch_one<DT_BF16, true, false>(p, num_sms, stream) : launch_one<DT_FP16, true, false>(p, num_sms, stream);
  }
  if (pair) {
    if (fr && !ta) return dtype == DT_BF16 ? launch_one<DT_BF16, false, true, 1>(p, num_sms, stream) : launch_one<DT_FP16, false, true, 1>(p, num_sms, stream);
    if (ta && !fr) return dtype == DT_BF16 ? launch_one<DT_BF16, false, true, 2>(p, num_sms, stream) : launch_one<DT_FP16, false, true, 2>(p, num_sms, stream);
    return dtype == DT_BF16 ? launch_one<DT_BF16, false, true, 3>(p, num_sms, stream) : launch_one<DT_FP16, false, true, 3>(p, num_sms, stream);
  }
  if (fr && !ta) return dtype == DT_BF16 ? launch_one<DT_BF16, false, false, 1>(p, num_sms, stream) : launch_one<DT_FP16, false, false, 1>(p, num_sms, stream);
  if (ta && !fr) return dtype == DT_BF16 ? launch_one<DT_BF16, false, false, 2>(p, num_sms, stream) : launch_one<DT_FP16, false, false, 2>(p, num_sms, stream);
  return dtype == DT_BF16 ? launch_one<DT_BF16, false, false, 3>(p, num_sms, stream) : launch_one<DT_FP16, false, false, 3>(p, num_sms, stream);
}

}  // namespace sdf
