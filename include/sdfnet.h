/* sdfnet.h -- C ABI of the B200-native neural-SDF hot path.
 *
 * One shared library (libsdfnet.so, built from sdf-representation_b200/csrc by __graft_entry__.build()),
 * extern "C", plain pointers and sizes, no torch types, no exceptions across the boundary.  All data pointers are
 * DEVICE pointers owned by the caller (PyTorch); the library owns only its packed-weight cache and gradient
 * accumulators.  Every entry point returns 0 on success and a non-zero code otherwise; sdfnet_last_error() then
 * describes the failure.  Calls are asynchronous on the CUDA stream passed in (a cudaStream_t cast to void*).
 * There is no CPU fallback.
 *
 * Each entry point names the reference (newton-raphson/sdf-representation) interface it replaces.
 */
#ifndef SDFNET_H_
#define SDFNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Network description == the constructor arguments of model/networks.py:29-38 ImplicitNet(d_in, dims, skip_in,
 * beta, ...) restricted to what the kernels support: d_in = 3, all hidden widths equal (multiple of 64, <= 512),
 * at most one skip layer.  beta > 0 selects Softplus(beta); beta == 0 selects ReLU with tanh on the output
 * (model/networks.py:79-83,104-105). */
typedef struct sdfnet_desc {
  int32_t d_in;
  int32_t hidden;       /* H */
  int32_t num_hidden;   /* number of hidden layers L; the net has L+1 Linear layers lin0..linL */
  int32_t skip_layer;   /* index in skip_in (layer whose input is cat([x, input])/sqrt(2)); 0 = none */
  float beta;
  int32_t dtype;        /* 16-bit operand type of the tensor-core GEMMs: 0 = bf16, 1 = fp16 */
  int32_t precision;    /* 0 = single pass ("fast"), 1 = hi/lo split operands ("parity", fp32-grade) */
  int32_t device;       /* CUDA device ordinal */
} sdfnet_desc;

typedef struct sdfnet_handle* sdfnet_t;

/* IGRLOSS hyper-parameters: model/losses.py:100-105 (+ the data-term variant). */
typedef struct sdfnet_igr_cfg {
  float delta, tau, lambda_g, regularizer_threshold;
  int32_t data_term;    /* 0 = "pairwise": mean over the reference's (N,1)-(N,) broadcast; 1 = "pointwise" */
  float mean_t;         /* pairwise: global mean of clamp(S, +-delta) over the whole (multi-GPU) batch */
  int64_t n_global;     /* loss normaliser: points in the whole batch across ranks */
  const float* mean_t_dev; /* optional DEVICE pointer to mean_t (overrides mean_t; avoids a host sync per step) */
  const double* label_stats_dev; /* optional DEVICE {sum clamp(S), N} over the whole batch (sdfnet_label_stats + cross-rank
                                  * sum): overrides mean_t, mean_t_dev and n_global -- unequal shards need no host value */
} sdfnet_igr_cfg;

const char* sdfnet_last_error(void);

/* model/networks.py:29-83 (construction); the packed copy replaces nn.Linear's own weight storage on the hot path */
int sdfnet_create(const sdfnet_desc* desc, sdfnet_t* out);
int sdfnet_destroy(sdfnet_t h);

/* Re-pack lin{i}.weight / lin{i}.bias (fp32 device pointers, layouts [out,in] and [out], i = 0..L) into UMMA
 * operand images.  Call after every optimizer step (executor/executor.py:177), right before launching.
 * There are two image layouts (single-CTA and CTA-pair kernels): the call packs the biases and the layout(s) a program
 * of this handle has already used; a layout that is first needed by a later launch is packed then, from the parameter
 * pointers recorded here -- so the parameter memory must stay valid and unchanged until the next call. */
int sdfnet_pack_weights(sdfnet_t h, const float* const* weights, const float* const* biases, void* stream);

/* Bytes of scratch one point tile set needs, so callers can size `ws`: mode 0 = forward_grad, 1 = training. */
size_t sdfnet_workspace_bytes(sdfnet_t h, int64_t n_points, int32_t mode);

/* model/networks.py:87-106 ImplicitNet.forward: x [n,3] -> sdf [n] */
int sdfnet_forward(sdfnet_t h, const float* x, int64_t n, float* sdf_out, void* stream);

/* model/losses.py:283-296 compute_normal / executor/executordistributed.py:10-19 gradient /
 * ops/conversion_test/test_loading.cpp:62,99: sdf [n] and grad_x sdf [n,3] */
int sdfnet_forward_grad(sdfnet_t h, const float* x, int64_t n, float* sdf_out, float* grad_out, void* ws,
                        size_t ws_bytes, void* stream);

/* executor/executor.py:351-383 and evaluations/post_process.py:46-53,83-88: SDF on the points
 * [first, first+count) of the flattened n_axis^3 'ij' meshgrid (x slowest), coordinates taken from `axis`
 * (= torch.linspace(-1, 1, n_axis) evaluated by the caller). */
int sdfnet_grid_eval(sdfnet_t h, const float* axis, int32_t n_axis, int64_t first, int64_t count, float* sdf_out,
                     void* stream);

/* evaluations/compare_octree_dl.py:20-36 (per-node query), batched per level: SDF at the centres of the cells
 * `keys` (3 x level-bit Morton codes, x bit highest) of the octree over [-1,1]^3; refine_flags[i] = 1 iff
 * |sdf| <= band * h_level (h_level = 2 / 2^level), the single-rounding fp32 predicate of oracle/sdf_oracle.py:octree_flags. */
int sdfnet_octree_level(sdfnet_t h, const uint64_t* keys, int64_t count, int32_t level, float band, float* sdf_out,
                        uint8_t* refine_flags, void* stream);

/* The whole level loop of the octree build (evaluations/compare_octree_dl.py:20-36 node queries, batched per level; SURVEY.md
 * section 8d cfg 5) WITHOUT a host round trip per level: keys / sdf / flags are arenas of `capacity` entries, level_table
 * (DEVICE, int64 [level_hi + 2][2]) holds {count, offset} per level.  The caller fills the keys and table rows of the
 * levels <= max(level_lo, uni_level) (uniform levels: all 8^level Morton codes, or this rank's share); for every level the
 * forward kernel evaluates the cell centres, its last epilogue writes the refinement flags (1 below uni_level, else
 * |sdf| <= band * h_level), and an ordered device-side compaction appends the children of the flagged cells as the next
 * level and writes its table row.  *overflow (device) is set if the arena was too small (children beyond it are dropped). */
int sdfnet_octree_levels(sdfnet_t h, int32_t level_lo, int32_t level_hi, int32_t uni_level, float band, uint64_t* keys,
                         float* sdf, uint8_t* flags, int64_t capacity, int64_t* level_table, int32_t* overflow, void* stream);

/* Generic backward of the network (autograd path of model/losses.py:283-296 + loss.backward()):
 * given per-point seeds fbar = dL/df [n] (may be NULL) and gbar = dL/d(grad_x f) [n,3] (may be NULL), ACCUMULATE
 * dL/d(parameters) into the library's fp32 accumulators.  Fetch with sdfnet_read_grads. */
int sdfnet_backward(sdfnet_t h, const float* x, int64_t n, const float* fbar, const float* gbar, void* ws,
                    size_t ws_bytes, void* stream);
int sdfnet_zero_grads(sdfnet_t h, void* stream);
/* grads: flat fp32 [num_params] in state_dict order lin0.weight, lin0.bias, lin1.weight, ... (model/networks.py:77) */
int sdfnet_read_grads(sdfnet_t h, float* grads, float scale, void* stream);
int64_t sdfnet_num_params(sdfnet_t h);

/* model/losses.py:109-137 IGRLOSS.forward + executor/executor.py:174-176 backward, fused:
 * x [n,3], y [n,4] = [S, nx, ny, nz].  Accumulates parameter gradients (see sdfnet_read_grads) and adds this
 * shard's loss terms to loss_sums[8] (double, device): sum p, sum p^2, sum t, sum t^2, sum (p-t)^2, sum reg,
 * sum eik, count. */
int sdfnet_train_step_igr(sdfnet_t h, const float* x, const float* y, int64_t n, const sdfnet_igr_cfg* cfg,
                          double* loss_sums, void* ws, size_t ws_bytes, void* stream);

/* executor/executordistributed.py:63-79 point-cloud step: surface points x [n,3] (|f| term) and perturbed points
 * xp [m,3] (eikonal term); loss_sums[0] += sum |f|, loss_sums[1] += sum (|g|-1)^2.  n_global / m_global are the
 * normalisers over all ranks. */
int sdfnet_train_step_pcd(sdfnet_t h, const float* x, int64_t n, int64_t n_global, const float* xp, int64_t m,
                          int64_t m_global, float lambda_g, double* loss_sums, void* ws, size_t ws_bytes,
                          void* stream);

/* executor/executor.py:160,177 torch.optim.Adam(lr).step() fused over the flat parameter vector. */
int sdfnet_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, int32_t step, void* stream);

/* The step after loss.backward() in executor/executor.py:176-177, fused: ONE kernel unpacks the gradient accumulators
 * (after the optional cross-rank sum, see sdfnet_grad_buffer) into state_dict order, scales them by
 * grad_scale * (*grad_scale_dev) (the d loss / d loss seed of loss.backward(); NULL = 1) and applies torch.optim.Adam's
 * update (no amsgrad / weight decay) IN PLACE to the fp32 parameters weights[i] / biases[i] (device pointers, layouts as
 * sdfnet_pack_weights); exp_avg / exp_avg_sq are flat [num_params] in state_dict order; grads_out (optional, flat
 * [num_params]) receives the gradient that was applied.  ONE more kernel re-packs the operand images, bias rows and
 * last-layer vector from the updated parameters (what sdfnet_pack_weights does), so the next step can launch at once. */
int sdfnet_fused_adam_step(sdfnet_t h, float* const* weights, float* const* biases, float* exp_avg, float* exp_avg_sq, float lr,
                           float beta1, float beta2, float eps, int32_t step, float grad_scale, const float* grad_scale_dev,
                           float* grads_out, void* stream);

/* model/losses.py:50-69 WeightedSmoothL2Loss(weight_factor, delta) -- the loss the shipped ini selects -- as one fused
 * first-order step: forward, seeds, adjoint sweep, weight gradients into the accumulators; loss_sums[8] (double, device) +=
 * {sum p, sum p^2, sum w (t-p)^2, sum w, sum w t, sum w t^2, 0, count} of this shard.  label_stats_dev = DEVICE
 * {sum w, sum w t, sum w t^2, N} over the whole (multi-GPU) batch (sdfnet_wsl2_label_stats + cross-rank sum).  data_term
 * 0 = the reference's [N,1] - [N] broadcast mean in closed form, 1 = per-point mean(w (t - p)^2). */
int sdfnet_wsl2_label_stats(const float* y, int64_t n, float delta, float weight_factor, double* out4, void* stream);
int sdfnet_train_step_wsl2(sdfnet_t h, const float* x, const float* y, int64_t n, float delta, float weight_factor,
                           int32_t data_term, const double* label_stats_dev, double* loss_sums, void* ws, size_t ws_bytes,
                           void* stream);

/* Reproducibility switch (the reference's cuBLAS SGEMM + ATen path gives the same bits on every run; the default
 * split-K red.add of the weight-gradient GEMM does not).  on != 0: every split of a dW block stores into a private slot
 * and a second pass adds the slots in split order; the loss sums go through per-block partials added in block order.
 * Training results are then bit-identical from run to run on the same device and batch. */
int sdfnet_set_deterministic(sdfnet_t h, int32_t on);

/* model/losses.py:115 t = clamp(y[:,0], +-delta): out2 (device, double) = {sum t, n} of this shard's labels y [n,4].  Summed
 * over ranks it is the label_stats_dev of sdfnet_igr_cfg. */
int sdfnet_label_stats(const float* y, int64_t n, float delta, double* out2, void* stream);

/* Replaces nn.DataParallel's reduce-add of 18 gradient tensors into GPU 0 (executor/executor.py:53-56): the library's
 * gradient accumulators followed by a 32-float tail holding the running loss sums and sum(fbar) as (hi, lo) float pairs
 * -- ONE fp32 sum-allreduce (NCCL) over [*buf, *buf + *n_floats) carries the gradients and the loss terms of a step.
 * The buffer stays owned by the library; it is valid between sdfnet_train_step_* / sdfnet_backward and the next
 * sdfnet_zero_grads. */
int sdfnet_grad_buffer(sdfnet_t h, float** buf, int64_t* n_floats);

/* Loss value of the last step(s) from the tail of the gradient buffer (after the cross-rank sum): out4 (device) =
 * {total, term0, term1, term2}; n_global <= 0: the point count that travelled with the sums.  kind 0: IGRLOSS model/losses.py:118-136 {data, normal, eikonal};
 * kind 1: point-cloud loss executor/executordistributed.py:73-79 {surface, eikonal, 0}; kind 2: WeightedSmoothL2Loss
 * model/losses.py:50-69 {value, 0, 0}. */
int sdfnet_loss_value(sdfnet_t h, int32_t kind, float tau, float lambda_g, int32_t data_term, int64_t n_global, int64_t m_global,
                      float* out4, void* stream);

/* executor/executor.py:388 marching_cubes(level=0) -- cell classification only: for the (nx-1)(ny-1)(nz-1) cells
 * of an sdf block [nx,ny,nz], case_index (8-bit corner sign mask, bit c set iff corner c < 0) and the compacted
 * list of active cells (mask not 0 or 255). */
int sdfnet_mc_classify(const float* sdf, int32_t nx, int32_t ny, int32_t nz, uint8_t* case_index,
                       int64_t* active_cells, int64_t* n_active, void* stream);
/* Second phase on its own: active_cells == NULL only counts the active cells into *n_active (so the caller can size
 * the index buffer exactly), otherwise writes their flat indices in ascending order. */
int sdfnet_mc_compact(const uint8_t* case_index, int64_t n_cells, int64_t* active_cells, int64_t* n_active, void* stream);

/* executor/executor.py:388 marching_cubes(sdf, level=0.0, spacing=spacing) -- vertex interpolation stage for the active
 * cells: for cell c and edge e (Lorensen-Cline numbering: 0-3 bottom ring, 4-7 top ring, 8-11 verticals) whose corner
 * signs differ, verts[(c*12+e)*3..] = (idx_a + t (idx_b - idx_a)) * spacing with t = f_a / (f_a - f_b) (index space times
 * spacing, as skimage); non-crossing edges get (0,0,0).  edge_mask[c] has bit e set for crossing edges.  spacing3 is a
 * HOST pointer to 3 floats.  The triangle tables of skimage's Lewiner variant are not restated (parity unpinned there). */
int sdfnet_mc_edge_vertices(const float* sdf, int32_t nx, int32_t ny, int32_t nz, const int64_t* active_cells, int64_t n_active,
                            const float* spacing3, float* verts, uint16_t* edge_mask, void* stream);

/* executor/executor.py:388 marching_cubes(...) -- the faces: for the active cells (ascending list of sdfnet_mc_compact) and
 * their case indices, the triangles in cell order, each as three indices r * 12 + e into the [n_active][12] edge-vertex
 * array of sdfnet_mc_edge_vertices (r = rank of the cell in active_cells, e = edge id), normals pointing from negative to
 * positive sdf.  tri_vertex_index == NULL only counts the triangles into *n_tri (device) so the caller can size the
 * buffer.  The triangle table is derived by a watertight-by-construction rule (tools/make_mc_table.py), not
 * transcribed from skimage's Lewiner tables: same surface, same vertices, possibly different diagonals (parity unpinned). */
int sdfnet_mc_triangles(const uint8_t* case_index, const int64_t* active_cells, int64_t n_active, int64_t* tri_vertex_index,
                        int64_t* n_tri, void* stream);

/* Instrumentation (no reference counterpart; used by bench.py): kernels launched through this handle so far, and
 * CUDA-event timing of each kernel class on its launch stream.  Classes: 0 = layer chain forward (+ reverse),
 * 1 = layer chain tangent + adjoint, 2 = weight-gradient GEMM, 3 = everything else.  sdfnet_profile_read waits for the
 * recorded events and returns accumulated milliseconds and launch counts since sdfnet_profile_enable(h, 1). */
int64_t sdfnet_launch_count(sdfnet_t h);
int sdfnet_profile_enable(sdfnet_t h, int32_t on);
int sdfnet_profile_read(sdfnet_t h, double* ms4, int64_t* launches4);

/* Octree refinement (build-defined): children keys of the flagged cells, in order. */
int sdfnet_octree_refine(const uint64_t* keys, const uint8_t* flags, int64_t count, uint64_t* child_keys,
                         int64_t* n_children, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDFNET_H_ */
