// C ABI (include/sdfnet.h) over the kernels in chain.cu / dw.cu / aux.cu.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>
#include "../../include/sdfnet.h"
#include "sdf_internal.h"

namespace sdf {
cudaError_t launch_chain(const ChainParams& p, int dtype, int parity, int pair, int num_sms, cudaStream_t stream);
cudaError_t launch_dw(const DwItem* items_dev, int n_items, int dtype, int parity, int num_sms, cudaStream_t stream);
cudaError_t launch_dw2(const DwItem* items_dev, int n_items, int dtype, int parity, int num_sms, int* sync_ctr, int group,
                       cudaStream_t stream);
cudaError_t launch_igr_seeds(const float* f, const float* g, const float* y, int64_t n, float delta, float tau,
                             float lambda_g, float thr, float mean_t, const float* mean_t_dev, const double* stats_dev, double inv_n,
                             int data_term, int relu_tanh, unsigned* absmax, float* fbar, float* gbar, double* sums,
                             double* fbar_sum, double* part, cudaStream_t stream);
cudaError_t launch_label_stats(const float* y, int64_t n, float delta, double* out2, cudaStream_t stream);
cudaError_t launch_pcd_seeds(const float* f, const float* g, int64_t n, int which, float coef, int relu_tanh,
                             unsigned* absmax, float* fbar, float* gbar, double* sums, double* fbar_sum, double* part,
                             cudaStream_t stream);
cudaError_t launch_wsl2_label_stats(const float* y, int64_t n, float delta, float wf, double* out4, cudaStream_t stream);
cudaError_t launch_wsl2_seeds(const float* f, const float* y, int64_t n, float delta, float wf, int data_term, const double* stats,
                              int relu_tanh, unsigned* absmax, float* fbar, double* sums, double* fbar_sum, double* part,
                              cudaStream_t stream);
cudaError_t launch_absmax(const float* v, int64_t n, unsigned* absmax, cudaStream_t stream);
cudaError_t launch_seed_scale(unsigned* absmax, float* scale, const double* sums, const double* fbar_sum, float* tail,
                              cudaStream_t stream);
cudaError_t launch_pack_all(const PackAll& pa, cudaStream_t stream);
cudaError_t launch_fused_adam(int H, int L, int skip, const float* main_acc, const float* zx_acc, const float* qg_acc,
                              const float* as_acc, const float* adx_acc, const double* fbar_sum, const float* tail, long num_params,
                              float* const* W, float* const* b, float* m, float* v, float lr, float b1, float b2, float eps, int step,
                              float scale, const float* scale_dev, float* grads_out, cudaStream_t stream);
cudaError_t launch_loss_value(const float* tail, int kind, float tau, float lambda_g, int data_term, double inv_n, double inv_m,
                              float* out4, cudaStream_t stream);
cudaError_t launch_sum(const float* v, int64_t n, double* out, double* part, cudaStream_t stream);
cudaError_t launch_reduce_partials(float* acc, const float* partial, long n, int slots, cudaStream_t stream);
cudaError_t launch_tanh_seeds(const float* f, const float* g, const float* fbar_in, const float* gbar_in, int64_t n,
                              float* fbar, float* gbar, cudaStream_t stream);
cudaError_t launch_unpack_grads(int H, int L, int skip, const float* main_acc, const float* zx_acc, const float* qg_acc,
                                const float* as_acc, const float* adx_acc, const double* fbar_sum, const float* tail, long num_params,
                                float* grads, float scale, cudaStream_t stream);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                        int step, cudaStream_t stream);
cudaError_t launch_octree_flags(const float* sdf, int64_t n, float thresh, uint8_t* flags, cudaStream_t stream);
cudaError_t launch_mc_classify(const float* sdf, int nx, int ny, int nz, uint8_t* cases, cudaStream_t stream);
cudaError_t launch_mc_edge_vertices(const float* sdf, int nx, int ny, int nz, const long* cells, long n_cells, float sx, float sy, float sz,
                                    float* verts, unsigned short* mask, cudaStream_t stream);
long compact_num_blocks(long n);
cudaError_t launch_mc_triangles(const uint8_t* cases, const long* cells, long n, long* tri, long* total, long* scratch, cudaStream_t stream);
cudaError_t launch_octree_refine_dev(const uint8_t* flags, uint64_t* keys, const int64_t* lvl, int64_t* lvl_next, long* scratch,
                                     long level_capacity, long capacity, int* overflow, cudaStream_t stream);
cudaError_t launch_count_flags(const uint8_t* v, long n, int mode, long* total, long* scratch, cudaStream_t stream);
cudaError_t launch_compact(const uint8_t* v, long n, int mode, const uint64_t* keys, long* out_idx, uint64_t* out_keys,
                           long* total, long* scratch, cudaStream_t stream);
}  // namespace sdf

using namespace sdf;

namespace {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(x)                                                                                       \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) return fail(static_cast<int>(e_), "%s: %s", #x, cudaGetErrorString(e_)); \
  } while (0)

enum : int { PK_F0 = 0, PK_FH = 1, PK_FL = 2, PK_RH = 3, PK_R0 = 4 };

struct StepImage {  // shape of one packed step image
  int n_groups, ng, kchunks, k16, a_chunks, split;
  uint32_t chunk_bytes, off;
  uint32_t bytes() const { return static_cast<uint32_t>(n_groups) * kchunks * chunk_bytes; }
};
}  // namespace

namespace sdf {
// where the accumulated weight-gradient tiles live and how they map onto the parameter vector
struct GradLayout {
  int H, L, skip;
  float* main_acc;    // [(L-1)][H (in)][H (out)]   layers 1..L-1
  float* zx_acc;      // [L][16][H]   Zbar_l^T . Xblock     (col 15 -> db_l ; l = 0: cols 0..8 -> dW_0 term 1)
  float* qg_acc;      // [16][H]      q_0^T . Gblock        (cols 0..8 -> dW_0 term 2)
  float* as_acc;      // [16][H]      a_{L-1}^T . Sblock    (cols 0..2 -> dW_L term 1)
  float* adx_acc;     // [16][H]      adot_{L-1}^T . Xblock (col 15 -> dW_L term 2)
  double* fbar_sum;   // [1]          sum fbar -> db_L
  int64_t num_params;
};
}  // namespace sdf

constexpr size_t kGradTail = 32;
constexpr size_t kDetBlocks = 4096;  // >= the grid of the seed kernels

// cached work list of the weight-gradient kernels for one (workspace, tile count, order) combination
struct DwPlan {
  const void* ws = nullptr;
  int tiles = -1, so = -1;
  int n = 0, pair_n = 0, group = 1;
  DwItem* items_dev = nullptr;
  DwItem* items_host = nullptr;   // pinned: the upload is a true async copy, no stream synchronisation
  int* sync = nullptr;
  float* partial = nullptr;       // deterministic mode: [slots][acc_floats] private split-K outputs
  int slots = 0;
  int det = 0;
  uint64_t last_use = 0;
};

struct sdfnet_handle {
  sdfnet_desc d;
  int H, L, skip, num_sms;
  bool relu, parity;
  // fast mode, H = 512: the chain kernel runs on CTA pairs (tcgen05 cta_group::2, packed-weight layout Fp / Rp) --
  // measured against single CTAs: training step -5 %, 1024^3 grid 291 -> 310 M queries/s.  pair_train: training programs
  // (forward + reverse with spills, tangent + adjoint); pair_infer: forward(-gradient) inference programs.
  // SDFNET_PAIR=0 disables both (single-CTA layout F / R), SDFNET_PAIR=train keeps inference on single CTAs.
  bool pair_train, pair_infer;
  int pts;                 // points per tile: 128 (fast) or 64 (parity)
  float* scale_dev = nullptr;      // [0] seed multiplier (power of two), [1] its inverse
  unsigned* absmax_dev = nullptr;  // running max |seed| of the current chunk
  uint8_t* wpack = nullptr;
  size_t wpack_bytes = 0;
  float* bias = nullptr;
  float* w_last = nullptr;
  std::vector<StepImage> F, R;    // F[0..L], R[0..L-1]: single-CTA layout
  std::vector<StepImage> Fp, Rp;  // CTA-pair layout (only when pair_train || pair_infer)
  // Each layout is packed when sdfnet_pack_weights is called only if a program has used it before; a layout that is
  // first needed later is packed on demand from the parameter pointers of the last sdfnet_pack_weights call (the
  // caller packs right before launching, so they are current).  Training then packs the pair layout only, inference
  // the single-CTA layout only.
  std::vector<const float*> Wptr, Bptr;
  bool layout_used[2] = {false, false};
  bool layout_ok[2] = {false, false};
  // gradient accumulators
  float* acc = nullptr;
  size_t acc_floats = 0;
  float* tail = nullptr;           // kGradTail floats behind the accumulators: (hi, lo) pairs of the 8 loss sums and sum(fbar)
  double* fbar_sum = nullptr;
  GradLayout gl;
  // work lists of the weight-gradient kernels, cached per (workspace, tile count, order): a ragged last batch of an
  // epoch alternates with the full batches without any rebuild, synchronisation or pageable copy
  std::vector<DwPlan> plans;
  uint64_t plan_clock = 0;
  bool deterministic = false;      // sdfnet_set_deterministic: fixed-order reductions everywhere (bit-identical reruns)
  double* det_part = nullptr;      // per-block partial loss sums [kDetBlocks][16]
  long* oct_scratch = nullptr;     // block counts of the device-side octree compaction
  long oct_scratch_cap = 0;
  // instrumentation: kernels launched through this handle, and (when enabled) CUDA-event timing per kernel class
  int64_t launches = 0;
  bool prof_on = false;
  struct ProfRec { int cls; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[4] = {0, 0, 0, 0};
  int64_t prof_n[4] = {0, 0, 0, 0};
};

namespace {
enum : int { PC_CHAIN_FG = 0, PC_CHAIN_TA = 1, PC_DW = 2, PC_OTHER = 3 };
// RAII bracket around one kernel launch: counts it and, in profiling mode, records events on the launch stream
struct Launch {
  sdfnet_handle* h; cudaStream_t st; int idx = -1;
  Launch(sdfnet_handle* h_, int cls, cudaStream_t st_, int count = 1) : h(h_), st(st_) {
    h->launches += count;
    if (!h->prof_on) return;
    sdfnet_handle::ProfRec r; r.cls = cls;
    for (cudaEvent_t* e : {&r.e0, &r.e1}) {
      if (!h->ev_pool.empty()) { *e = h->ev_pool.back(); h->ev_pool.pop_back(); }
      else if (cudaEventCreate(e) != cudaSuccess) return;
    }
    cudaEventRecord(r.e0, st);
    h->prof.push_back(r);
    idx = static_cast<int>(h->prof.size()) - 1;
  }
  ~Launch() { if (idx >= 0) cudaEventRecord(h->prof[idx].e1, st); }
};
}  // namespace

extern "C" const char* sdfnet_last_error(void) { return g_err; }

// split: parity-mode image = hi chunks followed by lo chunks of the same weights (A chunk index wraps)
static StepImage make_image(int N, int K16total, int kchunk_k16, bool split, uint32_t& off) {
  StepImage s;
  s.ng = N < 256 ? N : 256;
  s.n_groups = N / s.ng;
  s.k16 = kchunk_k16;
  s.a_chunks = K16total / kchunk_k16;
  s.split = split ? 1 : 0;
  s.kchunks = s.a_chunks * (split ? 2 : 1);
  s.chunk_bytes = static_cast<uint32_t>(s.ng) * 16u * 2u * kchunk_k16;
  s.off = off;
  off += s.bytes();
  return s;
}

static int init_handle(sdfnet_handle* h, const sdfnet_desc* desc, int num_sms);

extern "C" int sdfnet_create(const sdfnet_desc* desc, sdfnet_t* out) {
  if (!desc || !out) return fail(1, "sdfnet_create: null argument");
  if (desc->d_in != 3) return fail(1, "sdfnet_create: d_in must be 3 (got %d)", desc->d_in);
  if (desc->hidden != 64 && desc->hidden != 128 && desc->hidden != 256 && desc->hidden != 512)
    return fail(1, "sdfnet_create: hidden width must be 64, 128, 256 or 512 (got %d)", desc->hidden);
  if (desc->num_hidden < 2 || desc->num_hidden > kMaxLayers)
    return fail(1, "sdfnet_create: num_hidden must be in [2,%d] (got %d)", kMaxLayers, desc->num_hidden);
  if (desc->skip_layer < 0 || desc->skip_layer == 1 || desc->skip_layer >= desc->num_hidden)
    return fail(1, "sdfnet_create: skip layer must be 0 (none) or in [2,%d) (got %d)", desc->num_hidden, desc->skip_layer);
  if (desc->dtype != DT_BF16 && desc->dtype != DT_FP16) return fail(1, "sdfnet_create: dtype must be 0 (bf16) or 1 (fp16)");
  if (desc->precision != 0 && desc->precision != 1) return fail(1, "sdfnet_create: precision must be 0 (fast) or 1 (parity)");
  if (!(desc->beta >= 0.f)) return fail(1, "sdfnet_create: beta must be >= 0");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (desc->device < 0 || desc->device >= ndev) return fail(1, "sdfnet_create: no CUDA device %d (no CPU fallback)", desc->device);
  CU(cudaSetDevice(desc->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, desc->device));
  if (prop.major != 10) return fail(1, "sdfnet_create: device %d is sm_%d%d; this library is sm_100a only", desc->device, prop.major, prop.minor);

  sdfnet_handle* h = new (std::nothrow) sdfnet_handle();
  if (!h) return fail(2, "sdfnet_create: out of host memory");
  const int rc = init_handle(h, desc, prop.multiProcessorCount);
  if (rc) {
    sdfnet_destroy(h);  // releases whatever was allocated before the failure
    return rc;
  }
  *out = h;
  return 0;
}

static int init_handle(sdfnet_handle* h, const sdfnet_desc* desc, int num_sms) {
  {
    // keep freed blocks of the stream-ordered pool cached (default: released at the next synchronisation)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, desc->device) == cudaSuccess && pool) {
      uint64_t keep = 64ull << 20;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  h->d = *desc;
  h->H = desc->hidden;
  h->L = desc->num_hidden;
  h->skip = desc->skip_layer;
  h->relu = desc->beta == 0.f;
  h->parity = desc->precision == 1;
  h->pts = h->parity ? kParityPoints : kTileRows;
  // CTA-pair layer chain (cta_group::2): halves the L2 -> SM weight bytes and doubles the time-depth of the weight ring
  // 
  {
    const char* e = getenv("SDFNET_PAIR");
    const bool can = !h->parity && desc->hidden == 512;
    h->pair_train = can && !(e && e[0] == '0');
    h->pair_infer = can && !(e && (e[0] == '0' || e[0] == 't'));
  }
  h->num_sms = num_sms;
  const int H = h->H, L = h->L;
  uint32_t off = 0;
  const bool sp = h->parity;
  auto build = [&](std::vector<StepImage>& F, std::vector<StepImage>& R, int k16_hidden) {
    F.resize(L + 1);
    R.resize(L);
    F[0] = make_image(H, 1, 1, false, off);                  // K = 16 (input block, already split along K)
    for (int l = 1; l < L; ++l) F[l] = make_image(H, H / 16, k16_hidden, sp, off);
    F[L] = make_image(16, H / 16, H / 16, sp, off);          // N = 16 (row 0 = lin_L)
    for (int l = L - 1; l >= 1; --l) R[l] = make_image(H, H / 16, k16_hidden, sp, off);
    R[0] = make_image(16, H / 16, H / 16, sp, off);          // N = 16 (rows 0..2 = d/dx)
  };
  build(h->F, h->R, 4);
  if (h->pair_train || h->pair_infer) build(h->Fp, h->Rp, 8);
  h->wpack_bytes = off;
  CU(cudaMalloc(&h->wpack, h->wpack_bytes));
  CU(cudaMalloc(&h->bias, sizeof(float) * (L + 1) * H));
  CU(cudaMalloc(&h->w_last, sizeof(float) * H));
  // gradient accumulators
  const size_t n_main = static_cast<size_t>(L - 1) * H * H;
  const size_t n_zx = static_cast<size_t>(L) * 16 * H;
  h->acc_floats = n_main + n_zx + 3 * 16 * static_cast<size_t>(H);
  CU(cudaMalloc(&h->acc, sizeof(float) * (h->acc_floats + kGradTail)));
  h->tail = h->acc + h->acc_floats;
  CU(cudaMalloc(&h->fbar_sum, sizeof(double)));
  CU(cudaMalloc(&h->scale_dev, 2 * sizeof(float)));
  CU(cudaMalloc(&h->absmax_dev, sizeof(unsigned)));
  CU(cudaMemset(h->absmax_dev, 0, sizeof(unsigned)));
  {
    const float one[2] = {1.f, 1.f};
    CU(cudaMemcpy(h->scale_dev, one, sizeof(one), cudaMemcpyHostToDevice));
  }
  CU(cudaMemset(h->acc, 0, sizeof(float) * (h->acc_floats + kGradTail)));
  CU(cudaMemset(h->fbar_sum, 0, sizeof(double)));
  GradLayout& gl = h->gl;
  gl.H = H; gl.L = L; gl.skip = h->skip;
  gl.main_acc = h->acc;
  gl.zx_acc = gl.main_acc + n_main;
  gl.qg_acc = gl.zx_acc + n_zx;
  gl.as_acc = gl.qg_acc + 16 * H;
  gl.adx_acc = gl.as_acc + 16 * H;
  gl.fbar_sum = h->fbar_sum;
  int64_t np = 0;
  for (int l = 0; l <= L; ++l) {
    const int in = l == 0 ? 3 : H;
    const int outd = l == L ? 1 : ((h->skip > 0 && l + 1 == h->skip) ? H - 3 : H);
    np += static_cast<int64_t>(outd) * in + outd;
  }
  gl.num_params = np;
  return 0;
}

extern "C" int sdfnet_destroy(sdfnet_t h) {
  if (!h) return 0;
  cudaFree(h->wpack);
  cudaFree(h->bias);
  cudaFree(h->w_last);
  cudaFree(h->acc);
  cudaFree(h->fbar_sum);
  cudaFree(h->scale_dev);
  cudaFree(h->absmax_dev);
  cudaFree(h->oct_scratch);
  cudaFree(h->det_part);
  for (auto& pl : h->plans) {
    cudaFree(pl.partial);
    cudaFree(pl.items_dev);
    cudaFreeHost(pl.items_host);
    cudaFree(pl.sync);
  }
  for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  delete h;
  return 0;
}

extern "C" int64_t sdfnet_num_params(sdfnet_t h) { return h ? h->gl.num_params : -1; }

static int out_dim_of(const sdfnet_handle* h, int l) {
  if (l == h->L) return 1;
  return (h->skip > 0 && l + 1 == h->skip) ? h->H - 3 : h->H;
}

// pack one weight layout (0 = single CTA, 1 = CTA pair) from the parameter pointers of the last sdfnet_pack_weights call:
// ONE launch for all 2L + 1 step images (+ the bias rows and the last-layer vector when with_bias)
static int pack_layout(sdfnet_t h, int layout, bool with_bias, cudaStream_t stream) {
  const int H = h->H, L = h->L, dt = h->d.dtype;
  const float c = static_cast<float>(1.0 / std::sqrt(2.0));
  const std::vector<StepImage>& F = layout ? h->Fp : h->F;
  const std::vector<StepImage>& R = layout ? h->Rp : h->R;
  const float* const* W = h->Wptr.data();
  PackAll pa;
  memset(&pa, 0, sizeof(pa));
  pa.H = H; pa.L = L; pa.dtype = dt; pa.halves = layout ? 2 : 1;
  auto add = [&](const StepImage& s, int l, int kind) {
    PackImg& im = pa.img[pa.n_img++];
    im.W = W[l];
    im.dst = reinterpret_cast<uint16_t*>(h->wpack + s.off);
    im.out_dim = out_dim_of(h, l); im.in_dim = l == 0 ? 3 : H;
    im.n_groups = s.n_groups; im.ng = s.ng; im.kchunks = s.kchunks; im.k16 = s.k16; im.kind = kind; im.split = s.split;
    im.scale = (h->skip > 0 && l == h->skip) ? c : 1.f;
    im.total = static_cast<long>(s.n_groups) * s.kchunks * (2 * s.k16) * s.ng * 8;
  };
  for (int l = 0; l <= L; ++l) add(F[l], l, l == 0 ? PK_F0 : (l == L ? PK_FL : PK_FH));
  for (int l = L - 1; l >= 0; --l) add(R[l], l, l == 0 ? PK_R0 : PK_RH);
  if (with_bias) {
    for (int l = 0; l <= L; ++l) { pa.b[l] = h->Bptr[l]; pa.b_out[l] = out_dim_of(h, l); }
    pa.bias_dst = h->bias; pa.w_last_dst = h->w_last; pa.w_last_src = W[L];
  }
  Launch lp(h, PC_OTHER, stream, 1);
  CU(launch_pack_all(pa, stream));
  h->layout_ok[layout] = true;
  return 0;
}
// biases + last-layer vector only (no weight layout in use yet)
static int pack_bias_only(sdfnet_t h, cudaStream_t stream) {
  PackAll pa;
  memset(&pa, 0, sizeof(pa));
  pa.H = h->H; pa.L = h->L; pa.dtype = h->d.dtype; pa.halves = 1;
  for (int l = 0; l <= h->L; ++l) { pa.b[l] = h->Bptr[l]; pa.b_out[l] = out_dim_of(h, l); }
  pa.bias_dst = h->bias; pa.w_last_dst = h->w_last; pa.w_last_src = h->Wptr[h->L];
  Launch lp(h, PC_OTHER, stream, 1);
  CU(launch_pack_all(pa, stream));
  return 0;
}
// re-pack everything in use from the recorded parameter pointers (sdfnet_pack_weights, sdfnet_fused_adam_step)
static int repack(sdfnet_t h, cudaStream_t stream) {
  const bool both = !h->Fp.empty();
  bool bias_done = false;
  for (int layout = 0; layout < 2; ++layout) {
    h->layout_ok[layout] = false;
    if (layout == 1 && !both) continue;
    if (h->layout_used[layout]) {
      if (int rc = pack_layout(h, layout, !bias_done, stream)) return rc;
      bias_done = true;
    }
  }
  if (!bias_done) return pack_bias_only(h, stream);
  return 0;
}
// called by every chain launch site: the layout its program was built on must hold the current weights
static int ensure_layout(sdfnet_t h, bool pr, cudaStream_t stream) {
  const int layout = pr ? 1 : 0;
  h->layout_used[layout] = true;
  if (h->layout_ok[layout]) return 0;
  if (h->Wptr.empty()) return fail(1, "sdfnet: sdfnet_pack_weights has not been called");
  return pack_layout(h, layout, false, stream);
}

extern "C" int sdfnet_pack_weights(sdfnet_t h, const float* const* W, const float* const* b, void* stream_) {
  if (!h || !W || !b) return fail(1, "sdfnet_pack_weights: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int L = h->L;
  for (int l = 0; l <= L; ++l)
    if (!W[l] || !b[l]) return fail(1, "sdfnet_pack_weights: null parameter pointer for layer %d", l);
  h->Wptr.assign(W, W + L + 1);
  h->Bptr.assign(b, b + L + 1);
  return repack(h, stream);
}

// ------------------------------------------------------------------------------------------------
// program construction
// ------------------------------------------------------------------------------------------------
static void add_step(ChainParams& p, int epi, int layer, const StepImage* img) {
  StepDesc& s = p.steps[p.n_steps++];
  s.epi = epi;
  s.layer = layer;
  if (img) {
    s.n_groups = img->n_groups; s.ng = img->ng; s.kchunks = img->kchunks; s.k16 = img->k16; s.a_chunks = img->a_chunks;
    s.chunk_bytes = img->chunk_bytes; s.w_off = img->off;
  } else {
    s.n_groups = 0; s.ng = 0; s.kchunks = 0; s.k16 = 0; s.a_chunks = 1; s.chunk_bytes = 0; s.w_off = 0;
  }
  // a 512-wide A image is released in four 128-column quarters (chain.cu hidden_step_staged): k_lo = weight chunks per
  // quarter; 0 = the step waits for the whole image before its first MMA
  s.k_lo = (img && img->a_chunks % 4 == 0 && img->k16 * 16 * img->a_chunks == 512) ? img->a_chunks / 4 : 0;
  s.signal = 0;
}
// a step's A image is consumed iff the next step of the program is a GEMM
static void finish_program(ChainParams& p) {
  for (int s = 0; s < p.n_steps; ++s) p.steps[s].signal = (s + 1 < p.n_steps && p.steps[s + 1].n_groups != 0) ? 1 : 0;
}
static void base_params(const sdfnet_handle* h, ChainParams& p) {
  memset(&p, 0, sizeof(p));
  p.H = h->H; p.L = h->L; p.skip = h->skip;
  p.beta = h->d.beta;
  p.inv_beta = h->relu ? 0.f : 1.f / h->d.beta;
  p.bias = h->bias; p.w_last = h->w_last; p.wpack = h->wpack;
  p.flags = h->relu ? FL_RELU : 0u;
  static const uint32_t dbg = [] { const char* e = getenv("SDFNET_DEBUG_FLAGS"); return e ? static_cast<uint32_t>(strtoul(e, nullptr, 0)) : 0u; }();
  p.flags |= dbg & 0xffff0000u;
}
// pr: build the program on the CTA-pair weight layout (the launch must then use the pair kernel)
// Fast mode at width 512 (the overlapped epilogue): the last hidden layer's epilogue also forms the network output by an
// in-register dot product and, when a reverse sweep follows, its first operand -- no N = 16 output GEMM step and no
// element-wise reverse-start step (EPI_FWD_FUSED).  SDFNET_NO_FWD_FUSED=1 keeps the separate steps (measurement switch).
static bool fuse_fwd_last(const sdfnet_handle* h) {
  static const bool on = [] { const char* e = getenv("SDFNET_NO_FWD_FUSED"); return !(e && e[0] == '1'); }();
  return on && !h->parity && !h->relu && h->H == 512;
}
static void prog_forward(const sdfnet_handle* h, ChainParams& p, bool pr) {
  const std::vector<StepImage>& F = pr ? h->Fp : h->F;
  const bool fused = fuse_fwd_last(h);
  for (int l = 0; l < h->L; ++l) add_step(p, (fused && l + 1 == h->L) ? EPI_FWD_FUSED : EPI_FWD_HIDDEN, l, &F[l]);
  if (!fused) add_step(p, EPI_FWD_LAST, h->L, &F[h->L]);
}
static void prog_reverse(const sdfnet_handle* h, ChainParams& p, bool pr) {
  const std::vector<StepImage>& R = pr ? h->Rp : h->R;
  if (!fuse_fwd_last(h)) add_step(p, EPI_REV_START, h->L - 1, nullptr);
  for (int l = h->L - 1; l >= 1; --l) add_step(p, EPI_REV_HIDDEN, l, &R[l]);
  add_step(p, EPI_REV_LAST, 0, &R[0]);
}
// SDFNET_NO_TAN_LAST=1 (measurement switch): separate last tangent layer and element-wise adjoint start, as in round 1
static bool fuse_tan_last() {
  static const bool on = [] { const char* e = getenv("SDFNET_NO_TAN_LAST"); return !(e && e[0] == '1'); }();
  return on;
}
static void prog_tangent(const sdfnet_handle* h, ChainParams& p, bool pr) {
  const std::vector<StepImage>& F = pr ? h->Fp : h->F;
  // the last tangent layer also forms Zbar_{L-1} (EPI_TAN_LAST): the adjoint sweep that follows starts at layer L-2
  for (int l = 0; l < h->L; ++l) add_step(p, (l + 1 < h->L || !fuse_tan_last()) ? EPI_TAN_HIDDEN : EPI_TAN_LAST, l, &F[l]);
}
static void prog_adjoint(const sdfnet_handle* h, ChainParams& p, bool pr, bool after_tangent) {
  const std::vector<StepImage>& R = pr ? h->Rp : h->R;
  if (!after_tangent || !fuse_tan_last()) add_step(p, EPI_ADJ_START, h->L - 1, nullptr);
  for (int l = h->L - 2; l >= 0; --l) add_step(p, EPI_ADJ_HIDDEN, l, &R[l + 1]);
}

// workspace carving -------------------------------------------------------------------------------
namespace {
struct Spill {
  uint8_t *A, *Q, *AD, *U, *Z, *X, *G, *S;
  float *f, *g, *fbar, *gbar;
};
size_t per_tile_bytes(const sdfnet_handle* h, int mode) {
  const size_t ib = img_bytes(h->H);
  if (mode == 0) return h->L * ib;
  return 5 * h->L * ib + 3 * 2 * kGranuleBytes + kTileRows * 8 * sizeof(float);
}
// carve the workspace for `tiles` tiles
// images are laid out for an even number of tiles: the CTA-pair kernel walks tiles two at a time and its odd
// partner of a ragged last pair works on a scratch tile
static inline size_t even_tiles(size_t tiles) { return (tiles + 1) & ~static_cast<size_t>(1); }
void carve(const sdfnet_handle* h, int mode, uint8_t* ws, size_t tiles_, Spill& s) {
  const size_t tiles = even_tiles(tiles_);
  const size_t ib = img_bytes(h->H);
  const size_t arr = h->L * ib * tiles;
  memset(&s, 0, sizeof(s));
  s.A = ws; ws += arr;
  if (mode == 0) return;
  s.Q = ws; ws += arr;
  s.AD = ws; ws += arr;
  s.U = ws; ws += arr;
  s.Z = ws; ws += arr;
  s.X = ws; ws += tiles * 2 * kGranuleBytes;
  s.G = ws; ws += tiles * 2 * kGranuleBytes;
  s.S = ws; ws += tiles * 2 * kGranuleBytes;
  s.f = reinterpret_cast<float*>(ws); ws += tiles * kTileRows * sizeof(float);
  s.g = reinterpret_cast<float*>(ws); ws += tiles * kTileRows * 3 * sizeof(float);
  s.fbar = reinterpret_cast<float*>(ws); ws += tiles * kTileRows * sizeof(float);
  s.gbar = reinterpret_cast<float*>(ws); ws += tiles * kTileRows * 3 * sizeof(float);
}
}  // namespace

extern "C" size_t sdfnet_workspace_bytes(sdfnet_t h, int64_t n_points, int32_t mode) {
  if (!h || n_points <= 0) return 0;
  const size_t tiles = (static_cast<size_t>((n_points + h->pts - 1) / h->pts) + 1) & ~static_cast<size_t>(1);
  return tiles * per_tile_bytes(h, mode) + 1024;
}

static uint8_t* align_ws(void* ws, size_t& bytes) {
  uintptr_t p = reinterpret_cast<uintptr_t>(ws);
  uintptr_t a = (p + 255) & ~static_cast<uintptr_t>(255);
  const size_t lost = a - p;
  bytes = bytes > lost ? bytes - lost : 0;
  return reinterpret_cast<uint8_t*>(a);
}

// ------------------------------------------------------------------------------------------------
// inference entry points
// ------------------------------------------------------------------------------------------------
static int run_forward(sdfnet_t h, ChainParams& p, int64_t n, cudaStream_t stream) {
  p.n = n;
  p.n_tiles = static_cast<int>((n + h->pts - 1) / h->pts);
  p.tile_stride = (p.n_tiles + 1) & ~1;
  if (int rc = ensure_layout(h, h->pair_infer, stream)) return rc;
  prog_forward(h, p, h->pair_infer);
  finish_program(p);
  Launch lr(h, PC_CHAIN_FG, stream);
  CU(launch_chain(p, h->d.dtype, h->parity ? 1 : 0, h->pair_infer ? 1 : 0, h->num_sms, stream));
  return 0;
}

extern "C" int sdfnet_forward(sdfnet_t h, const float* x, int64_t n, float* sdf_out, void* stream) {
  if (!h || !x || !sdf_out) return fail(1, "sdfnet_forward: null argument");
  if (n <= 0) return 0;
  // tiles are counted in int32: split very large batches
  const int64_t kMax = static_cast<int64_t>(1) << 30;
  for (int64_t o = 0; o < n; o += kMax) {
    ChainParams p;
    base_params(h, p);
    p.src = SRC_POINTS;
    p.x = x + o * 3;
    p.f_out = sdf_out + o;
    const int64_t m = n - o < kMax ? n - o : kMax;
    if (int rc = run_forward(h, p, m, static_cast<cudaStream_t>(stream))) return rc;
  }
  return 0;
}

extern "C" int sdfnet_grid_eval(sdfnet_t h, const float* axis, int32_t n_axis, int64_t first, int64_t count, float* sdf_out,
                                void* stream) {
  if (!h || !axis || !sdf_out) return fail(1, "sdfnet_grid_eval: null argument");
  if (n_axis <= 0) return fail(1, "sdfnet_grid_eval: n_axis must be positive");
  const int64_t total = static_cast<int64_t>(n_axis) * n_axis * n_axis;
  if (first < 0 || count < 0 || first + count > total)
    return fail(1, "sdfnet_grid_eval: range [%lld, %lld) outside the %d^3 grid", (long long)first, (long long)(first + count), n_axis);
  const int64_t kMax = static_cast<int64_t>(1) << 30;
  for (int64_t o = 0; o < count; o += kMax) {
    ChainParams p;
    base_params(h, p);
    p.src = SRC_GRID;
    p.axis = axis;
    p.grid_n = n_axis;
    p.grid_off = first + o;
    p.f_out = sdf_out + o;
    const int64_t m = count - o < kMax ? count - o : kMax;
    if (int rc = run_forward(h, p, m, static_cast<cudaStream_t>(stream))) return rc;
  }
  return 0;
}

extern "C" int sdfnet_octree_level(sdfnet_t h, const uint64_t* keys, int64_t count, int32_t level, float band, float* sdf_out,
                                   uint8_t* refine_flags, void* stream) {
  if (!h || !keys || !sdf_out) return fail(1, "sdfnet_octree_level: null argument");
  if (level < 0 || level > 20) return fail(1, "sdfnet_octree_level: level %d outside [0,20]", level);
  if (count <= 0) return 0;
  ChainParams p;
  base_params(h, p);
  p.src = SRC_OCTREE;
  p.keys = keys;
  p.level = level;
  p.f_out = sdf_out;
  if (int rc = run_forward(h, p, count, static_cast<cudaStream_t>(stream))) return rc;
  if (refine_flags) {
    const float thresh = band * ldexpf(2.f, -level);  // band * h_level, one fp32 rounding (oracle/sdf_oracle.py)
    Launch lf(h, PC_OTHER, static_cast<cudaStream_t>(stream));
    CU(launch_octree_flags(sdf_out, count, thresh, refine_flags, static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

extern "C" int sdfnet_octree_levels(sdfnet_t h, int32_t level_lo, int32_t level_hi, int32_t uni_level, float band, uint64_t* keys,
                                    float* sdf, uint8_t* flags, int64_t capacity, int64_t* level_table, int32_t* overflow,
                                    void* stream_) {
  if (!h || !keys || !sdf || !flags || !level_table || !overflow) return fail(1, "sdfnet_octree_levels: null argument");
  if (level_lo < 0 || level_hi > 20 || level_lo > level_hi) return fail(1, "sdfnet_octree_levels: levels [%d,%d] outside [0,20]", level_lo, level_hi);
  if (capacity <= 0 || capacity > (static_cast<int64_t>(1) << 30)) return fail(1, "sdfnet_octree_levels: capacity must be in (0, 2^30]");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long need = compact_num_blocks(capacity) + 2;
  if (need > h->oct_scratch_cap) {
    CU(cudaStreamSynchronize(stream));
    cudaFree(h->oct_scratch);
    h->oct_scratch = nullptr;
    h->oct_scratch_cap = 0;
    CU(cudaMalloc(&h->oct_scratch, sizeof(long) * need));
    h->oct_scratch_cap = need;
  }
  for (int level = level_lo; level <= level_hi; ++level) {
    // a level holds at most 8^level cells: bounds the grid of the (device-count driven) launches of the coarse levels
    int64_t level_cap = capacity;
    if (3 * level < 40 && (static_cast<int64_t>(1) << (3 * level)) < capacity) level_cap = static_cast<int64_t>(1) << (3 * level);
    ChainParams p;
    base_params(h, p);
    p.src = SRC_OCTREE;
    p.keys = keys;
    p.level = level;
    p.f_out = sdf;
    p.lvl = level_table + 2 * level;
    p.flag_out = flags;
    p.flag_thresh = band * ldexpf(2.f, -level);  // band * h_level, one fp32 rounding (oracle/sdf_oracle.py:octree_flags)
    p.flag_all = level < uni_level ? 1 : 0;
    if (int rc = run_forward(h, p, level_cap, stream)) return rc;
    if (level < level_hi && level >= uni_level) {
      Launch lf(h, PC_OTHER, stream, 3);
      CU(launch_octree_refine_dev(flags, keys, level_table + 2 * level, level_table + 2 * (level + 1), h->oct_scratch, level_cap,
                                  capacity, overflow, stream));
    }
  }
  return 0;
}

extern "C" int sdfnet_forward_grad(sdfnet_t h, const float* x, int64_t n, float* sdf_out, float* grad_out, void* ws_,
                                   size_t ws_bytes, void* stream) {
  if (!h || !x || !grad_out) return fail(1, "sdfnet_forward_grad: null argument");
  if (n <= 0) return 0;
  uint8_t* ws = align_ws(ws_, ws_bytes);
  const size_t ptb = per_tile_bytes(h, 0);
  size_t cap_tiles = (ws_bytes / ptb) & ~static_cast<size_t>(1);
  if (!ws_ || cap_tiles == 0) return fail(1, "sdfnet_forward_grad: workspace too small (%zu bytes, need >= %zu per 128 points)", ws_bytes, ptb);
  if (cap_tiles > (1u << 22)) cap_tiles = 1u << 22;
  const int64_t chunk = static_cast<int64_t>(cap_tiles) * h->pts;
  for (int64_t o = 0; o < n; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    ChainParams p;
    base_params(h, p);
    p.src = SRC_POINTS;
    p.x = x + o * 3;
    p.f_out = sdf_out ? sdf_out + o : nullptr;
    p.g_out = grad_out + o * 3;
    p.n = m;
    p.n_tiles = static_cast<int>((m + h->pts - 1) / h->pts);
  p.tile_stride = (p.n_tiles + 1) & ~1;
    p.tile_stride = (p.n_tiles + 1) & ~1;
    Spill s;
    carve(h, 0, ws, p.n_tiles, s);
    p.Aimg = s.A;
    p.flags |= FL_SPILL_A;
    if (int rc = ensure_layout(h, h->pair_infer, static_cast<cudaStream_t>(stream))) return rc;
    prog_forward(h, p, h->pair_infer);
    prog_reverse(h, p, h->pair_infer);
    finish_program(p);
    Launch lr(h, PC_CHAIN_FG, static_cast<cudaStream_t>(stream));
    CU(launch_chain(p, h->d.dtype, h->parity ? 1 : 0, h->pair_infer ? 1 : 0, h->num_sms, static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// training entry points
// ------------------------------------------------------------------------------------------------
namespace {
// Build and run the weight-gradient work list for one chunk of `tiles` tiles.
int run_dw(sdfnet_t h, const Spill& s, int tiles, bool second_order, cudaStream_t stream) {
  const int H = h->H, L = h->L;
  const uint32_t ib = img_bytes(H);
  const uint32_t xb = 2 * kGranuleBytes;
  const GradLayout& gl = h->gl;
  auto launch_plan = [&](DwPlan& pl) -> int {
    pl.last_use = ++h->plan_clock;
    if (pl.det) CU(cudaMemsetAsync(pl.partial, 0, sizeof(float) * pl.slots * h->acc_floats, stream));
    Launch lr(h, PC_DW, stream, (pl.pair_n > 0 && pl.n > pl.pair_n ? 2 : 1) + (pl.det ? 1 : 0));
    if (pl.pair_n > 0)
      CU(launch_dw2(pl.items_dev, pl.pair_n, h->d.dtype, h->parity ? 1 : 0, h->num_sms, pl.sync, pl.group, stream));
    if (pl.n > pl.pair_n)
      CU(launch_dw(pl.items_dev + pl.pair_n, pl.n - pl.pair_n, h->d.dtype, h->parity ? 1 : 0, h->num_sms, stream));
    if (pl.det) CU(launch_reduce_partials(h->acc, pl.partial, static_cast<long>(h->acc_floats), pl.slots, stream));
    return 0;
  };
  const int det = h->deterministic ? 1 : 0;
  for (auto& pl : h->plans)
    if (pl.ws == s.A && pl.tiles == tiles && pl.so == (second_order ? 1 : 0) && pl.det == det && pl.n > 0) return launch_plan(pl);
  int max_slots = 1;
  int dw_group = 1;
  std::vector<DwItem> items;
  const int nmb = H / 128 > 0 ? H / 128 : 1;
  if (H % 128) return fail(1, "training requires hidden width multiple of 128 (got %d)", H);
  const size_t arr = static_cast<size_t>(ib) * even_tiles(tiles);  // bytes of one layer of one spill array
  auto split_range = [&](int sp, int ksplit, uint32_t& t0, uint32_t& t1) {
    t0 = static_cast<uint32_t>(static_cast<int64_t>(tiles) * sp / ksplit);
    t1 = static_cast<uint32_t>(static_cast<int64_t>(tiles) * (sp + 1) / ksplit);
  };
  int pair_n = 0;
  const bool use_pairs = H % 256 == 0;
  if (use_pairs) {
    // 256 x 256 blocks of dW_l for the CTA-pair kernel (dw2_kernel)
    const int nb2 = H / 256;
    const int types = (L - 1) * nb2 * nb2;
    int ksplit = (8 * (h->num_sms / 2) + types - 1) / types;
    dw_group = getenv("SDFNET_DW_NOSYNC") ? 1 : nb2 * nb2;
    if (dw_group > 1) {
      // lock-step groups (dw.cu): deal the items over a multiple-of-group number of pairs with no ragged last round
      const int pairs_eff = (h->num_sms / 2) / dw_group * dw_group;
      if (pairs_eff > 0) {
        int a = types, b = pairs_eff;
        while (b) { const int t = a % b; a = b; b = t; }
        const int step = pairs_eff / a;  // ksplit must be a multiple of this for (ksplit * types) % pairs_eff == 0
        int k2 = (ksplit + step / 2) / step * step;
        if (k2 < step) k2 = step;
        if (k2 <= tiles) ksplit = k2;
      }
    }
    if (ksplit > tiles) ksplit = tiles;
    if (ksplit < 1) ksplit = 1;
    for (int sp = 0; sp < ksplit; ++sp) {
      uint32_t t0, t1;
      split_range(sp, ksplit, t0, t1);
      if (t1 <= t0) continue;
      for (int l = 1; l < L; ++l)
        for (int mb = 0; mb < nb2; ++mb)
          for (int nh = 0; nh < nb2; ++nh) {
            DwItem d{};
            d.A1 = s.Z + arr * l;        d.B1 = s.A + arr * (l - 1);
            if (second_order) { d.A2 = s.Q + arr * l; d.B2 = s.AD + arr * (l - 1); }
            d.a_tstride = ib; d.b_tstride = ib;
            d.a_off = static_cast<uint32_t>(mb) * 32 * kGranuleBytes;
            d.b_off = static_cast<uint32_t>(nh) * 32 * kGranuleBytes;
            d.n = 256; d.tile0 = t0; d.tile1 = t1; d.out_ld = H; d.inv_mult = h->scale_dev + 1;
            d.out = gl.main_acc + (static_cast<size_t>(l - 1) * H + static_cast<size_t>(nh) * 256) * H + mb * 256;
            // Zbar_l^T . Xblock (bias gradients, column 15) rides along with the Zbar_l tiles.  Every block of the
            // layer does the (cheap, N = 16) product so that the pairs sharing operand tiles stay in step -- which is
            // what lets L2 serve their second reads -- but only the nh == 0 blocks add it to the result.
            d.X = s.X; d.x_tstride = xb;
            d.out_x = nh == 0 ? gl.zx_acc + static_cast<size_t>(l) * 16 * H + mb * 256 : nullptr;
            d.store = static_cast<uint32_t>(sp); items.push_back(d);
          }
    }
    pair_n = static_cast<int>(items.size());
  }
  const int nb = H > 256 ? 256 : H;       // N per main item
  const int nnh = H / nb;
  const int types = (use_pairs ? 0 : (L - 1) * nmb * nnh) + ((use_pairs ? 1 : L) + 3) * nmb;
  int ksplit = (8 * h->num_sms + types - 1) / types;
  if (ksplit > tiles) ksplit = tiles;
  if (ksplit < 1) ksplit = 1;
  for (int sp = 0; sp < ksplit; ++sp) {
    uint32_t t0, t1;
    split_range(sp, ksplit, t0, t1);
    if (t1 <= t0) continue;
    for (int l = 1; l < L && !use_pairs; ++l) {
      for (int mb = 0; mb < nmb; ++mb)
        for (int nh = 0; nh < nnh; ++nh) {
          DwItem d{};
          d.A1 = s.Z + arr * l;        d.B1 = s.A + arr * (l - 1);
          if (second_order) { d.A2 = s.Q + arr * l; d.B2 = s.AD + arr * (l - 1); }
          d.a_tstride = ib; d.b_tstride = ib;
          d.a_off = static_cast<uint32_t>(mb) * 16 * kGranuleBytes;
          d.b_off = static_cast<uint32_t>(nh) * (nb / 8) * kGranuleBytes;
          d.n = nb; d.tile0 = t0; d.tile1 = t1; d.out_ld = H; d.inv_mult = h->scale_dev + 1;
          d.out = gl.main_acc + (static_cast<size_t>(l - 1) * H + static_cast<size_t>(nh) * nb) * H + mb * 128;
          d.store = static_cast<uint32_t>(sp); items.push_back(d);
        }
    }
    for (int mb = 0; mb < nmb; ++mb) {
      const uint32_t aoff = static_cast<uint32_t>(mb) * 16 * kGranuleBytes;
      for (int l = 0; l < (use_pairs ? 1 : L); ++l) {   // Zbar_l^T . Xblock -> biases (col 15) and, for l = 0, dW_0 term 1
        // (with the CTA-pair kernel the layers l >= 1 are folded into its items)
        DwItem d{};
        d.A1 = s.Z + arr * l; d.B1 = s.X; d.a_tstride = ib; d.b_tstride = xb; d.a_off = aoff; d.b_off = 0;
        d.n = 16; d.tile0 = t0; d.tile1 = t1; d.out_ld = H; d.inv_mult = h->scale_dev + 1;
        d.out = gl.zx_acc + static_cast<size_t>(l) * 16 * H + mb * 128;
        d.store = static_cast<uint32_t>(sp); items.push_back(d);
      }
      {  // a_{L-1}^T . Sblock -> dW_L term 1
        DwItem d{};
        d.A1 = s.A + arr * (L - 1); d.B1 = s.S; d.a_tstride = ib; d.b_tstride = xb; d.a_off = aoff; d.b_off = 0;
        d.n = 16; d.tile0 = t0; d.tile1 = t1; d.out_ld = H; d.inv_mult = h->scale_dev + 1; d.out = gl.as_acc + mb * 128;
        d.store = static_cast<uint32_t>(sp); items.push_back(d);
      }
      if (second_order) {
        DwItem d{};  // q_0^T . Gblock -> dW_0 term 2
        d.A1 = s.Q; d.B1 = s.G; d.a_tstride = ib; d.b_tstride = xb; d.a_off = aoff; d.b_off = 0;
        d.n = 16; d.tile0 = t0; d.tile1 = t1; d.out_ld = H; d.inv_mult = h->scale_dev + 1; d.out = gl.qg_acc + mb * 128;
        d.store = static_cast<uint32_t>(sp); items.push_back(d);
        DwItem e{};  // adot_{L-1}^T . Xblock -> dW_L term 2 (col 15)
        e.A1 = s.AD + arr * (L - 1); e.B1 = s.X; e.a_tstride = ib; e.b_tstride = xb; e.a_off = aoff; e.b_off = 0;
        e.n = 16; e.tile0 = t0; e.tile1 = t1; e.out_ld = H; e.inv_mult = h->scale_dev + 1; e.out = gl.adx_acc + mb * 128;
        e.store = static_cast<uint32_t>(sp); items.push_back(e);
      }
    }
  }
  const int n_items = static_cast<int>(items.size());
  if (det) {
    // every item of split sp writes (plain stores) into slot sp: same offsets as in the accumulator
    for (auto& d : items) {
      const int sp = static_cast<int>(d.store);  // the builders below left the split index here
      if (sp + 1 > max_slots) max_slots = sp + 1;
    }
  }
  // new plan: own device list, pinned host copy (true async upload, no synchronisation), own lock-step counters.  The
  // cache holds a few plans (full batch, ragged last batch, first / second order); the least recently used one is
  // recycled only after the stream has drained (a launch may still read it) -- not in steady state.
  DwPlan* pl = nullptr;
  constexpr size_t kMaxPlans = 16;
  if (h->plans.size() < kMaxPlans) {
    h->plans.emplace_back();
    pl = &h->plans.back();
  } else {
    pl = &h->plans[0];
    for (auto& q : h->plans) if (q.last_use < pl->last_use) pl = &q;
    CU(cudaStreamSynchronize(stream));
    cudaFree(pl->items_dev); cudaFreeHost(pl->items_host); cudaFree(pl->sync); cudaFree(pl->partial);
    *pl = DwPlan();
  }
  if (det) {
    CU(cudaMalloc(&pl->partial, sizeof(float) * max_slots * h->acc_floats));
    for (auto& d : items) {
      const size_t sp = d.store;
      d.out = pl->partial + sp * h->acc_floats + (d.out - h->acc);
      if (d.out_x) d.out_x = pl->partial + sp * h->acc_floats + (d.out_x - h->acc);
      d.store = 1;
    }
    pl->slots = max_slots;
  } else {
    for (auto& d : items) d.store = 0;
  }
  pl->det = det;
  CU(cudaMalloc(&pl->items_dev, sizeof(DwItem) * n_items));
  CU(cudaMallocHost(&pl->items_host, sizeof(DwItem) * n_items));
  CU(cudaMalloc(&pl->sync, sizeof(int) * (pair_n + 64)));
  memcpy(pl->items_host, items.data(), sizeof(DwItem) * n_items);
  CU(cudaMemcpyAsync(pl->items_dev, pl->items_host, sizeof(DwItem) * n_items, cudaMemcpyHostToDevice, stream));
  pl->ws = s.A; pl->tiles = tiles; pl->so = second_order ? 1 : 0;
  pl->n = n_items; pl->pair_n = pair_n; pl->group = dw_group;
  return launch_plan(*pl);
}

// forward (+reverse) over one chunk, spilling what training needs
int run_fg(sdfnet_t h, const float* x, int64_t m, const Spill& s, bool need_g, cudaStream_t stream) {
  ChainParams p;
  base_params(h, p);
  p.src = SRC_POINTS;
  p.x = x;
  p.n = m;
  p.n_tiles = static_cast<int>((m + h->pts - 1) / h->pts);
  p.tile_stride = (p.n_tiles + 1) & ~1;
  p.f_out = s.f;
  p.g_out = need_g ? s.g : nullptr;
  p.Aimg = s.A; p.Qimg = s.Q; p.Ximg = s.X;
  p.flags |= FL_SPILL_A | FL_SPILL_X | (need_g ? FL_SPILL_Q : 0u);
  if (int rc = ensure_layout(h, h->pair_train, stream)) return rc;
  prog_forward(h, p, h->pair_train);
  if (need_g) prog_reverse(h, p, h->pair_train);
  finish_program(p);
  Launch lr(h, PC_CHAIN_FG, stream);
  CU(launch_chain(p, h->d.dtype, h->parity ? 1 : 0, h->pair_train ? 1 : 0, h->num_sms, stream));
  return 0;
}
// tangent (+adjoint) sweeps over one chunk
int run_ta(sdfnet_t h, int64_t m, const Spill& s, const float* fbar, const float* gbar, cudaStream_t stream) {
  ChainParams p;
  base_params(h, p);
  p.n = m;
  p.n_tiles = static_cast<int>((m + h->pts - 1) / h->pts);
  p.tile_stride = (p.n_tiles + 1) & ~1;
  p.fbar = fbar; p.gbar = gbar; p.seed_mult = h->scale_dev;
  p.Aimg = s.A; p.Qimg = s.Q; p.ADimg = s.AD; p.Uimg = s.U; p.Zimg = s.Z; p.Ximg = s.X; p.Gimg = s.G; p.Simg = s.S;
  if (int rc = ensure_layout(h, h->pair_train, stream)) return rc;
  if (gbar) {
    p.flags |= FL_TANGENT;
    prog_tangent(h, p, h->pair_train);
  }
  prog_adjoint(h, p, h->pair_train, gbar != nullptr);
  finish_program(p);
  Launch lr(h, PC_CHAIN_TA, stream);
  CU(launch_chain(p, h->d.dtype, h->parity ? 1 : 0, h->pair_train ? 1 : 0, h->num_sms, stream));
  return 0;
}
int64_t chunk_points(sdfnet_t h, size_t ws_bytes) {
  size_t cap_tiles = (ws_bytes / per_tile_bytes(h, 1)) & ~static_cast<size_t>(1);
  if (cap_tiles > (1u << 22)) cap_tiles = 1u << 22;
  return static_cast<int64_t>(cap_tiles) * h->pts;
}
}  // namespace

extern "C" int sdfnet_zero_grads(sdfnet_t h, void* stream) {
  if (!h) return fail(1, "sdfnet_zero_grads: null handle");
  CU(cudaMemsetAsync(h->acc, 0, sizeof(float) * (h->acc_floats + kGradTail), static_cast<cudaStream_t>(stream)));
  CU(cudaMemsetAsync(h->fbar_sum, 0, sizeof(double), static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int sdfnet_read_grads(sdfnet_t h, float* grads, float scale, void* stream) {
  if (!h || !grads) return fail(1, "sdfnet_read_grads: null argument");
  const GradLayout& gl = h->gl;
  Launch lr(h, PC_OTHER, static_cast<cudaStream_t>(stream));
  CU(launch_unpack_grads(gl.H, gl.L, gl.skip, gl.main_acc, gl.zx_acc, gl.qg_acc, gl.as_acc, gl.adx_acc, gl.fbar_sum, h->tail,
                         gl.num_params, grads, scale, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int sdfnet_backward(sdfnet_t h, const float* x, int64_t n, const float* fbar, const float* gbar, void* ws_,
                               size_t ws_bytes, void* stream_) {
  if (!h || !x) return fail(1, "sdfnet_backward: null argument");
  if (!fbar && !gbar) return 0;
  if (n <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = align_ws(ws_, ws_bytes);
  const int64_t chunk = ws_ ? chunk_points(h, ws_bytes) : 0;
  if (chunk <= 0) return fail(1, "sdfnet_backward: workspace too small (%zu bytes, need >= %zu per 128 points)", ws_bytes, per_tile_bytes(h, 1));
  for (int64_t o = 0; o < n; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    const int tiles = static_cast<int>((m + h->pts - 1) / h->pts);
    Spill s;
    carve(h, 1, ws, tiles, s);
    // the ReLU/tanh network needs f (and g) to move the seeds through the output tanh
    const bool need_g = gbar != nullptr;
    if (int rc = run_fg(h, x + o * 3, m, s, need_g, stream)) return rc;
    const float* fb = fbar ? fbar + o : nullptr;
    const float* gb = gbar ? gbar + o * 3 : nullptr;
    if (h->relu) {
      CU(launch_tanh_seeds(s.f, s.g, fb, gb, m, s.fbar, s.gbar, stream));
      fb = s.fbar;
      gb = gbar ? s.gbar : nullptr;
    }
    if (fb) CU(launch_sum(fb, m, h->fbar_sum, h->deterministic ? h->det_part : nullptr, stream));
    {
      Launch ls(h, PC_OTHER, stream, 2 + (fb ? 1 : 0) + (gb ? 1 : 0));
      if (fb) CU(launch_absmax(fb, m, h->absmax_dev, stream));
      if (gb) CU(launch_absmax(gb, 3 * m, h->absmax_dev, stream));
      CU(launch_seed_scale(h->absmax_dev, h->scale_dev, nullptr, h->fbar_sum, h->tail, stream));
    }
    if (int rc = run_ta(h, m, s, fb, gb, stream)) return rc;
    if (int rc = run_dw(h, s, tiles, gbar != nullptr, stream)) return rc;
  }
  return 0;
}

extern "C" int sdfnet_train_step_igr(sdfnet_t h, const float* x, const float* y, int64_t n, const sdfnet_igr_cfg* cfg,
                                     double* loss_sums, void* ws_, size_t ws_bytes, void* stream_) {
  if (!h || !x || !y || !cfg || !loss_sums) return fail(1, "sdfnet_train_step_igr: null argument");
  if (n <= 0) return 0;
  if (cfg->n_global <= 0 && !cfg->label_stats_dev) return fail(1, "sdfnet_train_step_igr: n_global must be positive (or label_stats_dev given)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = align_ws(ws_, ws_bytes);
  const int64_t chunk = ws_ ? chunk_points(h, ws_bytes) : 0;
  if (chunk <= 0) return fail(1, "sdfnet_train_step_igr: workspace too small (%zu bytes, need >= %zu per 128 points)", ws_bytes, per_tile_bytes(h, 1));
  for (int64_t o = 0; o < n; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    const int tiles = static_cast<int>((m + h->pts - 1) / h->pts);
    Spill s;
    carve(h, 1, ws, tiles, s);
    if (int rc = run_fg(h, x + o * 3, m, s, true, stream)) return rc;
    {
      Launch ls(h, PC_OTHER, stream, 2);
      CU(launch_igr_seeds(s.f, s.g, y + o * 4, m, cfg->delta, cfg->tau, cfg->lambda_g, cfg->regularizer_threshold,
                          cfg->mean_t, cfg->mean_t_dev, cfg->label_stats_dev, cfg->n_global > 0 ? 1.0 / static_cast<double>(cfg->n_global) : 0.0,
                          cfg->data_term, h->relu ? 1 : 0,
                          h->absmax_dev, s.fbar, s.gbar, loss_sums, h->fbar_sum, h->deterministic ? h->det_part : nullptr, stream));
      CU(launch_seed_scale(h->absmax_dev, h->scale_dev, loss_sums, h->fbar_sum, h->tail, stream));
    }
    if (int rc = run_ta(h, m, s, s.fbar, s.gbar, stream)) return rc;
    if (int rc = run_dw(h, s, tiles, true, stream)) return rc;
  }
  return 0;
}

extern "C" int sdfnet_train_step_pcd(sdfnet_t h, const float* x, int64_t n, int64_t n_global, const float* xp, int64_t m_,
                                     int64_t m_global, float lambda_g, double* loss_sums, void* ws_, size_t ws_bytes,
                                     void* stream_) {
  if (!h || !loss_sums) return fail(1, "sdfnet_train_step_pcd: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = align_ws(ws_, ws_bytes);
  const int64_t chunk = ws_ ? chunk_points(h, ws_bytes) : 0;
  if (chunk <= 0) return fail(1, "sdfnet_train_step_pcd: workspace too small (%zu bytes, need >= %zu per 128 points)", ws_bytes, per_tile_bytes(h, 1));
  // surface points: L += mean |f|   (first order: forward + adjoint + dW term 1)
  for (int64_t o = 0; x && o < n; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    const int tiles = static_cast<int>((m + h->pts - 1) / h->pts);
    Spill s;
    carve(h, 1, ws, tiles, s);
    if (int rc = run_fg(h, x + o * 3, m, s, false, stream)) return rc;
    {
      Launch ls(h, PC_OTHER, stream, 2);
      CU(launch_pcd_seeds(s.f, nullptr, m, 0, 1.f / static_cast<float>(n_global), h->relu ? 1 : 0, h->absmax_dev, s.fbar,
                          nullptr, loss_sums, h->fbar_sum, h->deterministic ? h->det_part : nullptr, stream));
      CU(launch_seed_scale(h->absmax_dev, h->scale_dev, loss_sums, h->fbar_sum, h->tail, stream));
    }
    if (int rc = run_ta(h, m, s, s.fbar, nullptr, stream)) return rc;
    if (int rc = run_dw(h, s, tiles, false, stream)) return rc;
  }
  // perturbed points: L += lambda mean (|grad f| - 1)^2
  for (int64_t o = 0; xp && o < m_; o += chunk) {
    const int64_t m = m_ - o < chunk ? m_ - o : chunk;
    const int tiles = static_cast<int>((m + h->pts - 1) / h->pts);
    Spill s;
    carve(h, 1, ws, tiles, s);
    if (int rc = run_fg(h, xp + o * 3, m, s, true, stream)) return rc;
    {
      Launch ls(h, PC_OTHER, stream, 2);
      CU(launch_pcd_seeds(s.f, s.g, m, 1, lambda_g / static_cast<float>(m_global), h->relu ? 1 : 0, h->absmax_dev, s.fbar,
                          s.gbar, loss_sums, h->fbar_sum, h->deterministic ? h->det_part : nullptr, stream));
      CU(launch_seed_scale(h->absmax_dev, h->scale_dev, loss_sums, h->fbar_sum, h->tail, stream));
    }
    if (int rc = run_ta(h, m, s, s.fbar, s.gbar, stream)) return rc;
    if (int rc = run_dw(h, s, tiles, true, stream)) return rc;
  }
  return 0;
}

extern "C" int sdfnet_wsl2_label_stats(const float* y, int64_t n, float delta, float weight_factor, double* out4, void* stream_) {
  if (!y || !out4) return fail(1, "sdfnet_wsl2_label_stats: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CU(cudaMemsetAsync(out4, 0, 4 * sizeof(double), stream));
  if (n > 0) CU(launch_wsl2_label_stats(y, n, delta, weight_factor, out4, stream));
  return 0;
}

// model/losses.py:50-69 WeightedSmoothL2Loss as one fused first-order step: forward (spills a_l) -> seeds + loss sums ->
// adjoint sweep -> first term of the weight-gradient GEMMs (3F per point; the generic autograd path runs the forward twice)
extern "C" int sdfnet_train_step_wsl2(sdfnet_t h, const float* x, const float* y, int64_t n, float delta, float weight_factor,
                                      int32_t data_term, const double* label_stats_dev, double* loss_sums, void* ws_, size_t ws_bytes,
                                      void* stream_) {
  if (!h || !x || !y || !label_stats_dev || !loss_sums) return fail(1, "sdfnet_train_step_wsl2: null argument");
  if (data_term != 0 && data_term != 1) return fail(1, "sdfnet_train_step_wsl2: data_term must be 0 (pairwise) or 1 (pointwise)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  uint8_t* ws = align_ws(ws_, ws_bytes);
  const int64_t chunk = ws_ ? chunk_points(h, ws_bytes) : 0;
  if (chunk <= 0) return fail(1, "sdfnet_train_step_wsl2: workspace too small (%zu bytes, need >= %zu per 128 points)", ws_bytes, per_tile_bytes(h, 1));
  for (int64_t o = 0; o < n; o += chunk) {
    const int64_t m = n - o < chunk ? n - o : chunk;
    const int tiles = static_cast<int>((m + h->pts - 1) / h->pts);
    Spill s;
    carve(h, 1, ws, tiles, s);
    if (int rc = run_fg(h, x + o * 3, m, s, false, stream)) return rc;
    {
      Launch ls(h, PC_OTHER, stream, 2);
      CU(launch_wsl2_seeds(s.f, y + o * 4, m, delta, weight_factor, data_term, label_stats_dev, h->relu ? 1 : 0, h->absmax_dev, s.fbar,
                           loss_sums, h->fbar_sum, h->deterministic ? h->det_part : nullptr, stream));
      CU(launch_seed_scale(h->absmax_dev, h->scale_dev, loss_sums, h->fbar_sum, h->tail, stream));
    }
    if (int rc = run_ta(h, m, s, s.fbar, nullptr, stream)) return rc;
    if (int rc = run_dw(h, s, tiles, false, stream)) return rc;
  }
  return 0;
}

extern "C" int sdfnet_set_deterministic(sdfnet_t h, int32_t on) {
  if (!h) return fail(1, "sdfnet_set_deterministic: null handle");
  if (on && !h->det_part) CU(cudaMalloc(&h->det_part, sizeof(double) * kDetBlocks * 16));
  h->deterministic = on != 0;
  return 0;
}

extern "C" int sdfnet_label_stats(const float* y, int64_t n, float delta, double* out2, void* stream_) {
  if (!y || !out2) return fail(1, "sdfnet_label_stats: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CU(cudaMemsetAsync(out2, 0, 2 * sizeof(double), stream));
  if (n > 0) CU(launch_label_stats(y, n, delta, out2, stream));
  return 0;
}

extern "C" int sdfnet_grad_buffer(sdfnet_t h, float** buf, int64_t* n_floats) {
  if (!h || !buf || !n_floats) return fail(1, "sdfnet_grad_buffer: null argument");
  *buf = h->acc;
  *n_floats = static_cast<int64_t>(h->acc_floats + kGradTail);
  return 0;
}

extern "C" int sdfnet_loss_value(sdfnet_t h, int32_t kind, float tau, float lambda_g, int32_t data_term, int64_t n_global,
                                 int64_t m_global, float* out4, void* stream) {
  if (!h || !out4) return fail(1, "sdfnet_loss_value: null argument");
  if (kind < 0 || kind > 2) return fail(1, "sdfnet_loss_value: kind must be 0 (IGRLOSS), 1 (point cloud) or 2 (WeightedSmoothL2Loss)");
  Launch ll(h, PC_OTHER, static_cast<cudaStream_t>(stream));
  CU(launch_loss_value(h->tail, kind, tau, lambda_g, data_term, n_global > 0 ? 1.0 / static_cast<double>(n_global) : 0.0,
                       m_global > 0 ? 1.0 / static_cast<double>(m_global) : 0.0, out4, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int sdfnet_fused_adam_step(sdfnet_t h, float* const* W, float* const* b, float* exp_avg, float* exp_avg_sq, float lr,
                                      float beta1, float beta2, float eps, int32_t step, float grad_scale,
                                      const float* grad_scale_dev, float* grads_out, void* stream_) {
  if (!h || !W || !b || !exp_avg || !exp_avg_sq) return fail(1, "sdfnet_fused_adam_step: null argument");
  if (step < 1) return fail(1, "sdfnet_fused_adam_step: step counts from 1");
  for (int l = 0; l <= h->L; ++l)
    if (!W[l] || !b[l]) return fail(1, "sdfnet_fused_adam_step: null parameter pointer for layer %d", l);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const GradLayout& gl = h->gl;
  {
    Launch la(h, PC_OTHER, stream);
    CU(launch_fused_adam(gl.H, gl.L, gl.skip, gl.main_acc, gl.zx_acc, gl.qg_acc, gl.as_acc, gl.adx_acc, gl.fbar_sum, h->tail,
                         gl.num_params, W, b, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale, grad_scale_dev,
                         grads_out, stream));
  }
  // the updated parameters become the packed operand images of the next step
  h->Wptr.assign(W, W + h->L + 1);
  h->Bptr.assign(b, b + h->L + 1);
  return repack(h, stream);
}

extern "C" int64_t sdfnet_launch_count(sdfnet_t h) { return h ? h->launches : -1; }

extern "C" int sdfnet_profile_enable(sdfnet_t h, int32_t on) {
  if (!h) return fail(1, "sdfnet_profile_enable: null handle");
  h->prof_on = on != 0;
  if (on) {
    for (int i = 0; i < 4; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  }
  return 0;
}

extern "C" int sdfnet_profile_read(sdfnet_t h, double* ms4, int64_t* launches4) {
  if (!h || !ms4 || !launches4) return fail(1, "sdfnet_profile_read: null argument");
  for (auto& r : h->prof) {
    CU(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, r.e0, r.e1));
    h->prof_ms[r.cls] += ms;
    h->prof_n[r.cls] += 1;
    h->ev_pool.push_back(r.e0);
    h->ev_pool.push_back(r.e1);
  }
  h->prof.clear();
  for (int i = 0; i < 4; ++i) { ms4[i] = h->prof_ms[i]; launches4[i] = h->prof_n[i]; }
  return 0;
}

extern "C" int sdfnet_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, int32_t step, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail(1, "sdfnet_adam_step: null argument");
  if (step < 1) return fail(1, "sdfnet_adam_step: step counts from 1");
  CU(launch_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int sdfnet_mc_classify(const float* sdf, int32_t nx, int32_t ny, int32_t nz, uint8_t* case_index,
                                  int64_t* active_cells, int64_t* n_active, void* stream_) {
  if (!sdf || !case_index) return fail(1, "sdfnet_mc_classify: null argument");
  if (nx < 2 || ny < 2 || nz < 2) return fail(1, "sdfnet_mc_classify: block must be at least 2x2x2");
  if (nx > 65536 || ny > 65536) return fail(1, "sdfnet_mc_classify: block larger than 65536 planes / rows");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CU(launch_mc_classify(sdf, nx, ny, nz, case_index, stream));
  if (n_active) {
    const int64_t total = static_cast<int64_t>(nx - 1) * (ny - 1) * (nz - 1);
    return sdfnet_mc_compact(case_index, total, active_cells, n_active, stream_);
  }
  return 0;
}

extern "C" int sdfnet_mc_edge_vertices(const float* sdf, int32_t nx, int32_t ny, int32_t nz, const int64_t* active_cells, int64_t n_active,
                                       const float* spacing3, float* verts, uint16_t* edge_mask, void* stream_) {
  if (!sdf || !spacing3 || (n_active > 0 && (!active_cells || !verts || !edge_mask))) return fail(1, "sdfnet_mc_edge_vertices: null argument");
  if (nx < 2 || ny < 2 || nz < 2 || n_active < 0) return fail(1, "sdfnet_mc_edge_vertices: bad sizes");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CU(launch_mc_edge_vertices(sdf, nx, ny, nz, reinterpret_cast<const long*>(active_cells), n_active, spacing3[0], spacing3[1], spacing3[2],
                             verts, edge_mask, stream));
  return 0;
}

// The three handle-free scan entry points below take their block-sum scratch (8 bytes per 1024 elements) from the device's
// stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the caller's stream): no synchronisation, safe for concurrent
// streams -- a cached buffer would need a handle or a lock -- and after its first use the pool serves the request from
// memory it already holds (sdfnet_create raises the pool's release threshold so it keeps it).
extern "C" int sdfnet_mc_triangles(const uint8_t* case_index, const int64_t* active_cells, int64_t n_active, int64_t* tri_vertex_index,
                                   int64_t* n_tri, void* stream_) {
  if (!case_index || !n_tri || (n_active > 0 && !active_cells)) return fail(1, "sdfnet_mc_triangles: null argument");
  if (n_active < 0) return fail(1, "sdfnet_mc_triangles: negative cell count");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  long* scratch = nullptr;
  CU(cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(long) * (compact_num_blocks(n_active) + 1), stream));
  cudaError_t e = launch_mc_triangles(case_index, reinterpret_cast<const long*>(active_cells), n_active, reinterpret_cast<long*>(tri_vertex_index),
                                      reinterpret_cast<long*>(n_tri), scratch, stream);
  cudaFreeAsync(scratch, stream);
  CU(e);
  return 0;
}

extern "C" int sdfnet_mc_compact(const uint8_t* case_index, int64_t n_cells, int64_t* active_cells, int64_t* n_active,
                                 void* stream_) {
  if (!case_index || !n_active) return fail(1, "sdfnet_mc_compact: null argument");
  if (n_cells < 0) return fail(1, "sdfnet_mc_compact: negative cell count");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static_assert(sizeof(long) == sizeof(int64_t), "LP64 expected");
  long* scratch = nullptr;
  CU(cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(long) * (compact_num_blocks(n_cells) + 1), stream));
  cudaError_t e = active_cells
                      ? launch_compact(case_index, n_cells, 0, nullptr, reinterpret_cast<long*>(active_cells), nullptr,
                                       reinterpret_cast<long*>(n_active), scratch, stream)
                      : launch_count_flags(case_index, n_cells, 0, reinterpret_cast<long*>(n_active), scratch, stream);
  cudaFreeAsync(scratch, stream);
  CU(e);
  return 0;
}

extern "C" int sdfnet_octree_refine(const uint64_t* keys, const uint8_t* flags, int64_t count, uint64_t* child_keys,
                                    int64_t* n_children, void* stream_) {
  if (!keys || !flags || !child_keys || !n_children) return fail(1, "sdfnet_octree_refine: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  long* scratch = nullptr;
  CU(cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(long) * (compact_num_blocks(count) + 1), stream));
  // n_children receives the number of refined CELLS; each contributes 8 consecutive child keys
  cudaError_t e = launch_compact(flags, count, 1, keys, nullptr, child_keys, reinterpret_cast<long*>(n_children), scratch, stream);
  cudaFreeAsync(scratch, stream);
  CU(e);
  return 0;
}
