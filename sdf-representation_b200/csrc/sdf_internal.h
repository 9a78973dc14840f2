// Internal definitions shared by the kernels (chain.cu, dw.cu, aux.cu) and the C-ABI (api.cu).
//
// Data layout ("granule-major operand image"), used for every 16-bit matrix this library keeps in shared
// memory or spills to HBM:  a [R rows x C cols] matrix is stored as [C/8 granules][R rows][8 elements],
// i.e. element (r, c) lives at byte ((c/8)*R + r)*16 + (c%8)*2.  With R = 128 this single image is
//   * a legal no-swizzle K-major  UMMA operand (rows = points,  K = features)  -> the layer chain,
//   * a legal no-swizzle MN-major UMMA operand (K = points, MN = features)     -> the weight-gradient GEMM,
//   * perfectly coalesced for epilogue threads (lane = row: 32 lanes x 16 B contiguous),
//   * movable with one 1-D bulk copy per contiguous granule range.
#pragma once
#include <cstdint>

namespace sdf {

constexpr int kTileRows = 128;           // rows of one operand image: 128 points (fast) or 64 hi + 64 lo rows (parity)
constexpr int kParityPoints = 64;        // points per tile in parity mode
constexpr int kGranuleBytes = kTileRows * 16;  // 2048: one 8-column granule of a tile image
constexpr int kMaxH = 512;
constexpr int kMaxLayers = 16;           // hidden layers
constexpr int kMaxSteps = 2 * kMaxLayers + 4;
constexpr int kStageBytes = 32768;       // one weight chunk (<= 256 rows x 64 K x 2 B)
constexpr int kMaxDevices = 64;          // per-device launch state (function attributes) is kept per device ordinal
constexpr int kXCols = 16;               // width of the input block (x / g-bar / f-bar split columns)

// element type of the 16-bit operands
enum : int { DT_BF16 = 0, DT_FP16 = 1 };

// epilogue kinds of the layer chain
enum : int {
  EPI_FWD_HIDDEN = 0,  // a = act(z + b)                         -> next A operand (+ spill A)
  EPI_FWD_LAST = 1,    // f = z_L + b_L (tanh in relu mode)      -> f_out
  EPI_REV_START = 2,   // q_{L-1} = w_L * act'(z_{L-1})          (no GEMM)
  EPI_REV_HIDDEN = 3,  // q_{l-1} = (q_l W_l) * act'(z_{l-1})    (+ skip rows -> g, spill Q)
  EPI_REV_LAST = 4,    // g = q_0 W_0 (+ skip part)              -> g_out
  EPI_TAN_HIDDEN = 5,  // adot_l = act'(z_l) * zdot_l ; u_l = act''(z_l) s_l zdot_l   (spill AD, U)
  EPI_ADJ_START = 6,   // Zbar_{L-1} = act'(z_{L-1}) w_L fbar + u_{L-1}               (no GEMM)
  EPI_ADJ_HIDDEN = 7,  // Zbar_l = act'(z_l) (Zbar_{l+1} W_{l+1}) + u_l               (spill Z)
  EPI_TAN_LAST = 8,    // last tangent layer fused with the adjoint start: adot_{L-1} (spill AD) and
                       // Zbar_{L-1} = act'(z_{L-1}) w_L fbar + u_{L-1} -> next A operand (spill Z); u_{L-1} never leaves the SM
  EPI_FWD_FUSED = 9,   // last hidden forward layer fused with the output layer and the reverse start (fast mode, H = 512):
                       // a_{L-1} (spill A), f = w_L . a_{L-1} + b_L by an in-register dot product (no N = 16 GEMM step),
                       // q_{L-1} = w_L * act'(z_{L-1}) -> next A operand (spill Q) when a reverse sweep follows
};

struct StepDesc {
  int32_t epi;
  int32_t layer;        // meaning depends on epi (index of the layer whose activation state is touched)
  int32_t n_groups;     // TMEM column groups of this GEMM; 0 = element-wise step without GEMM
  int32_t ng;           // N (rows of the weight chunk) per group
  int32_t kchunks;      // weight chunks per group (parity mode: hi chunks then lo chunks)
  int32_t a_chunks;     // distinct A chunks; the A chunk of weight chunk kc is kc % a_chunks
  int32_t k16;          // K=16 instructions per chunk
  uint32_t chunk_bytes;
  uint32_t w_off;       // byte offset of the first chunk of this step in the packed weight stream
  int32_t k_lo;         // weight chunks per 128-column quarter of a 512-wide A image: chunk kc of group 0 may start as
                        // soon as quarter kc / k_lo is in place (0: the step needs the whole A image first)
  int32_t signal;       // 1: the A image this step leaves behind is consumed by a GEMM (arrive on the a_q barriers)
};

// where the points of a tile come from
enum : int { SRC_POINTS = 0, SRC_GRID = 1, SRC_OCTREE = 2 };

// flag bits
enum : uint32_t {
  FL_SPILL_A = 1u << 0,   // forward sweep writes the activation images (needed by reverse / training)
  FL_SPILL_Q = 1u << 1,   // reverse sweep writes q images (training)
  FL_SPILL_X = 1u << 2,   // write the input-block image (training, for dW_0 / biases)
  FL_TANGENT = 1u << 3,   // TA kernel: run the tangent sweep (second-order terms present)
  FL_RELU = 1u << 4,      // ReLU hidden activation + tanh output (beta == 0 in the reference ini)
  // measurement-only switches (SDFNET_DEBUG_FLAGS environment variable; results are wrong when set)
  FL_DBG_NO_MMA = 1u << 16,   // MMA issuer waits / commits but issues no tcgen05.mma
  FL_DBG_NO_MATH = 1u << 17,  // hidden-layer epilogues skip their arithmetic and stores
  FL_DBG_NO_W = 1u << 18,     // weight producer signals the ring without copying
  FL_DBG_NO_MUFU = 1u << 19,  // ex2 / lg2 replaced by plain arithmetic
  FL_DBG_NO_AUX = 1u << 20,   // spill-image / bias operands of the hidden epilogues are not loaded
  FL_DBG_NO_STG = 1u << 21,   // hidden epilogues do not store their spill images
};

struct ChainParams {
  // network
  int32_t H, L, skip, n_steps;
  float beta, inv_beta;
  const float* bias;       // [(L+1)][H] fp32, zero padded
  const float* w_last;     // [H] fp32 (last layer weight row)
  const uint8_t* wpack;    // packed 16-bit weight chunk stream
  StepDesc steps[kMaxSteps];
  // points
  int64_t n;               // valid points
  int32_t n_tiles;
  int32_t tile_stride;     // tiles per layer in the spill arrays (n_tiles rounded up to even)
  int32_t src;
  const float* x;          // SRC_POINTS: [n][3]
  const float* axis;       // SRC_GRID: [grid_n] axis vector (torch.linspace output)
  int32_t grid_n;
  int64_t grid_off;        // flat index of the first point ((ix*n + iy)*n + iz ordering)
  const uint64_t* keys;    // SRC_OCTREE: Morton keys of the cells of one level
  int32_t level;
  const int64_t* lvl;      // optional DEVICE {count, offset}: overrides n (and offsets keys / f_out / flag_out); n_tiles = capacity
  uint8_t* flag_out;       // optional: octree refinement flags, written by the last forward epilogue
  float flag_thresh;       // |f| <= flag_thresh  (band * h_level)
  int32_t flag_all;        // 1: every cell is refined (levels below the uniform level)
  // outputs
  float* f_out;            // [n]
  float* g_out;            // [n][3]
  // seeds (TA kernel); both are multiplied by *seed_mult (a power of two chosen so that the largest seed is O(1):
  // keeps the 16-bit adjoint images out of fp16 subnormals), dw.cu multiplies the result by 1 / *seed_mult
  const float* fbar;       // [n]
  const float* gbar;       // [n][3]
  const float* seed_mult;  // device scalar, may be nullptr (= 1)
  // spill images, each [layer][tile][image]
  uint8_t* Aimg;
  uint8_t* Qimg;
  uint8_t* ADimg;
  uint8_t* Uimg;
  uint8_t* Zimg;
  uint8_t* Ximg;           // [tile][2 granules]  x split block (+ ones column)
  uint8_t* Gimg;           // [tile][2 granules]  g-bar split block
  uint8_t* Simg;           // [tile][2 granules]  f-bar split block (+ ones column)
  uint32_t flags;
#ifdef SDF_CLOCKS
  long long* dbg_clk;      // measurement build only: [8 epilogue kinds][8] cycle counters of thread 0, summed over CTAs
#endif
};

// one unit of work of the weight-gradient kernel:  out[n][m] += sum_tiles  A1^T B1 (+ A2^T B2)
struct DwItem {
  const uint8_t* A1;
  const uint8_t* B1;
  const uint8_t* A2;       // nullptr: single term
  const uint8_t* B2;
  uint32_t a_tstride, b_tstride;   // bytes between consecutive tiles of the A / B images
  uint32_t a_off, b_off;           // byte offset of the first granule used inside a tile image
  uint32_t n;                      // 256 or 16
  uint32_t tile0, tile1;
  uint32_t out_ld;                 // leading dimension (elements) of out (index [n][m])
  float* out;
  const float* inv_mult;           // device scalar multiplied into the result (nullptr = 1)
  // CTA-pair kernel only: the first term's A operand is also multiplied with a 16-column block image (bias / first
  // layer products ride along with the tile that is in shared memory anyway): out_x[16][m] += A1^T X
  const uint8_t* X;                // nullptr: none
  uint32_t x_tstride;
  float* out_x;
  // 0: red.add into out (split-K partial sums meet in the accumulator in completion order); 1: plain store into a slot
  // this item owns alone (deterministic mode: a second pass adds the slots in split order, api.cu)
  uint32_t store;
};

// one launch packs every step image of a weight layout + biases + last-layer vector (aux.cu pack_all_kernel)
struct PackImg {
  const float* W;
  uint16_t* dst;
  int out_dim, in_dim, n_groups, ng, kchunks, k16, kind, split;
  float scale;
  long total;
};
constexpr int kMaxPackImgs = 2 * kMaxLayers + 2;
struct PackAll {
  PackImg img[kMaxPackImgs];
  int n_img;
  const float* b[kMaxLayers + 1];
  int b_out[kMaxLayers + 1];
  float* bias_dst;          // [(L+1)][H]; nullptr: weight images only
  float* w_last_dst;        // [H]
  const float* w_last_src;  // lin_L.weight
  int H, L, dtype, halves;
};

__host__ __device__ inline uint32_t img_bytes(int H) { return static_cast<uint32_t>(H / 8) * kGranuleBytes; }

}  // namespace sdf
