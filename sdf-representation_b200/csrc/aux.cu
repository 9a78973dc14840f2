// Auxiliary kernels around the layer chain: weight packing into UMMA operand images, loss seeds and
// reductions (IGRLOSS / point-cloud / WeightedSmoothL2), gradient unpacking, fused Adam, marching-cubes cell
// classification and octree refinement.  All plain CUDA-core kernels: they are HBM- or latency-bound.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "sdf_internal.h"
#include "mc_tables.h"

namespace sdf {

__device__ __forceinline__ float rnd16(float v, int dtype) {
  return dtype == DT_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v));
}
__device__ __forceinline__ uint16_t bits16(float v, int dtype) {
  if (dtype == DT_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// Weight packing.  One launch per step image.  Destination element order (see chain.cu MMA loop):
//   [group g][k-chunk kc][half hf (halves)][granule gr (2*k16 per chunk)][row (ng / halves)][8 elements]
// logical operand B[n][k] with n = g*ng + hf*(ng/halves) + row, k = (kc*2*k16 + gr)*8 + e.
// halves = 2 is the CTA-pair layout: CTA r of a pair streams the contiguous half hf = r of every chunk.
// ------------------------------------------------------------------------------------------------
enum : int { PK_F0 = 0, PK_FH = 1, PK_FL = 2, PK_RH = 3, PK_R0 = 4 };

// split != 0 (parity mode): kchunks counts hi chunks followed by lo chunks, lo = rnd16(v - rnd16(v)).
// (kernel: pack_all_kernel below, one launch per layout)

// ------------------------------------------------------------------------------------------------
// Loss seeds.  Per point: (f, g) -> (fbar, gbar) = dL/df, dL/dg and the partial sums of the loss terms.
// Block-level warp-shuffle trees, one double atomicAdd per block and term.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// part == nullptr: one double atomicAdd per block and term (summation order = block completion order).
// part != nullptr (deterministic mode): block b writes its partial sums to part[b * stride + off + i]; a single-warp kernel
// adds them in block order afterwards (finalize_partials_kernel), so the result is bit-identical from run to run.
template <int NS>
__device__ __forceinline__ void block_accumulate(float (&v)[NS], double* sums, double* part = nullptr, int stride = 0, int off = 0) {
  __shared__ float part_sh[NS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    float w = warp_sum(v[i]);
    if (lane == 0) part_sh[i][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      float w = lane < nw ? part_sh[i][lane] : 0.f;
      w = warp_sum(w);
      if (lane == 0) {
        if (part) part[static_cast<long>(blockIdx.x) * stride + off + i] = static_cast<double>(w);
        else atomicAdd(&sums[i], static_cast<double>(w));
      }
    }
  }
}
// *dst.p[i] += sum over blocks (in block order) of part[b * stride + i], i < stride
struct PartDst { double* p[16]; };
__global__ void finalize_partials_kernel2(const double* __restrict__ part, int nblocks, int stride, PartDst dst) {
  const int i = threadIdx.x;
  if (i >= stride || !dst.p[i]) return;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) acc += part[static_cast<long>(b) * stride + i];
  *dst.p[i] += acc;
}

struct IgrCfgDev {
  float delta, tau, lambda_g, thr, mean_t, inv_n;
  int data_term, relu_tanh;
  const float* mean_t_dev;
  const double* stats_dev;   // {sum clamp(S), N} over the whole (multi-GPU) batch: overrides mean_t and inv_n
  unsigned* absmax;   // running max |seed| (float bits; non-negative floats order like unsigned ints)
};

__device__ __forceinline__ void block_absmax(float m, unsigned* absmax) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m <= 3.0e38f) atomicMax(absmax, __float_as_uint(m));
}

// reference model/losses.py:109-137 (IGRLOSS.forward) differentiated by hand:
//   p = clamp(f, +-delta), t = clamp(S, +-delta), m = |t| < thr
//   cos = n.g / (max(|n|,1e-8) max(|g|,1e-8));  reg = m ? (1-cos)^2 : 1e-8;  eik = m ? (|g|-1)^2 : 1e-8
//   L = D + tau mean(reg) + lambda mean(eik),  D = mean(p^2)+mean(t^2)-2 mean(p) mean(t)  |  mean((p-t)^2)
__global__ void igr_seeds_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ y,
                                 long n, IgrCfgDev c, float* __restrict__ fbar, float* __restrict__ gbar,
                                 double* __restrict__ sums, double* __restrict__ fbar_sum, double* __restrict__ part) {
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float fsum[1] = {0.f};
  float amax = 0.f;
  if (c.mean_t_dev) c.mean_t = __ldg(c.mean_t_dev);
  if (c.stats_dev) {
    const double st = c.stats_dev[0], nn = c.stats_dev[1];
    c.mean_t = static_cast<float>(st / nn);
    c.inv_n = static_cast<float>(1.0 / nn);
  }
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float fi = f[i];
    const float gx = g[3 * i], gy = g[3 * i + 1], gz = g[3 * i + 2];
    const float4 yi = __ldg(reinterpret_cast<const float4*>(y) + i);
    const float p = fminf(fmaxf(fi, -c.delta), c.delta);
    const float t = fminf(fmaxf(yi.x, -c.delta), c.delta);
    const bool inside = (fi >= -c.delta) && (fi <= c.delta);
    const bool m = fabsf(t) < c.thr;
    const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
    const float nn = sqrtf(yi.y * yi.y + yi.z * yi.z + yi.w * yi.w);
    const float gnc = fmaxf(gn, 1e-8f), nnc = fmaxf(nn, 1e-8f);
    const float dot = yi.y * gx + yi.z * gy + yi.w * gz;
    const float cs = dot / (nnc * gnc);
    const float reg = m ? (1.f - cs) * (1.f - cs) : 1e-8f;
    const float eik = m ? (gn - 1.f) * (gn - 1.f) : 1e-8f;
    acc[0] += p; acc[1] += p * p; acc[2] += t; acc[3] += t * t; acc[4] += (p - t) * (p - t);
    acc[5] += reg; acc[6] += eik; acc[7] += 1.f;
    // seeds
    float fb = inside ? 2.f * c.inv_n * (p - (c.data_term == 0 ? c.mean_t : t)) : 0.f;
    float bx = 0.f, by = 0.f, bz = 0.f;
    if (m) {
      // d(1-cos)^2/dg = -2 (1-cos) dcos/dg ; dcos/dg = n/(|n||g|) - cos g/|g|^2   (|g| above the eps clamp)
      const float k1 = -2.f * c.tau * c.inv_n * (1.f - cs);
      const float inv_ng = 1.f / (nnc * gnc);
      const float k2 = gn > 1e-8f ? cs / (gn * gn) : 0.f;
      // d(|g|-1)^2/dg = 2 (|g|-1) g/|g|
      const float k3 = gn > 0.f ? 2.f * c.lambda_g * c.inv_n * (gn - 1.f) / gn : 0.f;
      bx = k1 * (yi.y * inv_ng - k2 * gx) + k3 * gx;
      by = k1 * (yi.z * inv_ng - k2 * gy) + k3 * gy;
      bz = k1 * (yi.w * inv_ng - k2 * gz) + k3 * gz;
    }
    if (c.relu_tanh) {
      // f = tanh(z), g = (1-f^2) G: seeds of the linear-output network are
      //   zbar = fbar (1-f^2) - 2 f (gbar . g),   Gbar = (1-f^2) gbar
      const float d = 1.f - fi * fi;
      fb = fb * d - 2.f * fi * (bx * gx + by * gy + bz * gz);
      bx *= d; by *= d; bz *= d;
    }
    fbar[i] = fb;
    fsum[0] += fb;
    gbar[3 * i] = bx; gbar[3 * i + 1] = by; gbar[3 * i + 2] = bz;
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(fb), fabsf(bx)), fmaxf(fabsf(by), fabsf(bz))));
  }
  block_absmax(amax, c.absmax);
  block_accumulate<8>(acc, sums, part, 9, 0);
  __syncthreads();
  block_accumulate<1>(fsum, fbar_sum, part, 9, 8);  // sum fbar = d L / d b_L
}

// reference executor/executordistributed.py:73-79:  which = 0: surface points, L += mean |f|  (fbar = sign(f)/N);
// which = 1: perturbed points, L += lambda mean (|g|-1)^2  (gbar = (lambda/M) 2 (|g|-1) g/|g|).  coef = 1/N or lambda/M.
__global__ void pcd_seeds_kernel(const float* __restrict__ f, const float* __restrict__ g, long n, int which, float coef,
                                 int relu_tanh, unsigned* __restrict__ absmax, float* __restrict__ fbar, float* __restrict__ gbar,
                                 double* __restrict__ sums, double* __restrict__ fbar_sum, double* __restrict__ part) {
  float acc[1] = {0.f};
  float fsum[1] = {0.f};
  float amax = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float fi = f[i];
    float fb = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (which == 0) {
      acc[0] += fabsf(fi);
      fb = fi > 0.f ? coef : (fi < 0.f ? -coef : 0.f);
    } else {
      gx = g[3 * i]; gy = g[3 * i + 1]; gz = g[3 * i + 2];
      const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
      acc[0] += (gn - 1.f) * (gn - 1.f);
      const float k3 = gn > 0.f ? 2.f * coef * (gn - 1.f) / gn : 0.f;
      bx = k3 * gx; by = k3 * gy; bz = k3 * gz;
    }
    if (relu_tanh) {
      const float d = 1.f - fi * fi;
      fb = fb * d - 2.f * fi * (bx * gx + by * gy + bz * gz);
      bx *= d; by *= d; bz *= d;
    }
    fbar[i] = fb;
    fsum[0] += fb;
    if (gbar) { gbar[3 * i] = bx; gbar[3 * i + 1] = by; gbar[3 * i + 2] = bz; }
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(fb), fabsf(bx)), fmaxf(fabsf(by), fabsf(bz))));
  }
  block_absmax(amax, absmax);
  block_accumulate<1>(acc, sums + which, part, 2, 0);
  __syncthreads();
  block_accumulate<1>(fsum, fbar_sum, part, 2, 1);
}

// reference model/losses.py:50-69 WeightedSmoothL2Loss (the loss the shipped ini selects): t = clamp(S, +-delta),
// p = clamp(f, +-delta), w = 1 + wf exp(-|t|).  The reference's [N,1] - [N] broadcast makes it mean_{i,j} w_j (t_j - p_i)^2
// = mean(w t^2) - 2 mean(p) mean(w t) + mean(p^2) mean(w)  ("pairwise", data_term 0); data_term 1 = mean(w (t - p)^2).
// Label moments {sum w, sum w t, sum w t^2, N} of one shard (summed over ranks before the seeds):
__global__ void wsl2_label_stats_kernel(const float* __restrict__ y, long n, float delta, float wf, double* __restrict__ out4) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float t = fminf(fmaxf(__ldg(y + 4 * i), -delta), delta);
    const float w = 1.f + wf * expf(-fabsf(t));
    acc[0] += w; acc[1] += w * t; acc[2] += w * t * t; acc[3] += 1.f;
  }
  block_accumulate<4>(acc, out4);
}
// Seeds: d L / d f_i = [|f_i| <= delta] (2/N) (p_i mean(w) - mean(w t))  (pairwise)  |  (2/N) w_i (p_i - t_i)  (pointwise);
// no gradient term.  sums += {sum p, sum p^2, sum w (t-p)^2, sum w, sum w t, sum w t^2, 0, count} of this shard.
__global__ void wsl2_seeds_kernel(const float* __restrict__ f, const float* __restrict__ y, long n, float delta, float wf, int data_term,
                                  const double* __restrict__ stats, int relu_tanh, unsigned* __restrict__ absmax,
                                  float* __restrict__ fbar, double* __restrict__ sums, double* __restrict__ fbar_sum,
                                  double* __restrict__ part) {
  const double ng = stats[3];
  const float inv_n = ng > 0.0 ? static_cast<float>(1.0 / ng) : 0.f;
  const float mean_w = static_cast<float>(stats[0] * (ng > 0.0 ? 1.0 / ng : 0.0));
  const float mean_wt = static_cast<float>(stats[1] * (ng > 0.0 ? 1.0 / ng : 0.0));
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float fsum[1] = {0.f};
  float amax = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float fi = f[i];
    const float pi = fminf(fmaxf(fi, -delta), delta);
    const float t = fminf(fmaxf(__ldg(y + 4 * i), -delta), delta);
    const float w = 1.f + wf * expf(-fabsf(t));
    const bool inside = fi >= -delta && fi <= delta;  // torch.clamp passes the gradient on the closed interval
    float fb = data_term == 0 ? 2.f * inv_n * (pi * mean_w - mean_wt) : 2.f * inv_n * w * (pi - t);
    if (!inside) fb = 0.f;
    if (relu_tanh) fb *= 1.f - fi * fi;  // f = tanh(z) (model/networks.py:104-105): seed w.r.t. z
    acc[0] += pi; acc[1] += pi * pi; acc[2] += w * (t - pi) * (t - pi);
    acc[3] += w; acc[4] += w * t; acc[5] += w * t * t; acc[7] += 1.f;
    fbar[i] = fb;
    fsum[0] += fb;
    amax = fmaxf(amax, fabsf(fb));
  }
  block_absmax(amax, absmax);
  block_accumulate<8>(acc, sums, part, 9, 0);
  __syncthreads();
  block_accumulate<1>(fsum, fbar_sum, part, 9, 8);
}

__global__ void sum_kernel(const float* __restrict__ v, long n, double* __restrict__ out, double* __restrict__ part) {
  float acc[1] = {0.f};
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) acc[0] += v[i];
  block_accumulate<1>(acc, out, part, 1, 0);
}

// {sum clamp(y[:,0], +-delta), count} of one shard's labels (y [n,4]; model/losses.py:115): what the pairwise data term and
// the loss normaliser need from the whole batch before any seed can be formed
__global__ void label_stats_kernel(const float* __restrict__ y, long n, float delta, double* __restrict__ out2) {
  float acc[2] = {0.f, 0.f};
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    acc[0] += fminf(fmaxf(__ldg(y + 4 * i), -delta), delta);
    acc[1] += 1.f;
  }
  block_accumulate<2>(acc, out2);
}

static int seed_blocks(long n) {
  long b = (n + 255) / 256;
  if (b > 2368) b = 2368;  // 16 blocks per SM
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

cudaError_t launch_igr_seeds(const float* f, const float* g, const float* y, int64_t n, float delta, float tau,
                             float lambda_g, float thr, float mean_t, const float* mean_t_dev, const double* stats_dev, double inv_n,
                             int data_term, int relu_tanh, unsigned* absmax, float* fbar, float* gbar, double* sums,
                             double* fbar_sum, double* part, cudaStream_t stream) {
  IgrCfgDev c;
  c.mean_t_dev = mean_t_dev;
  c.stats_dev = stats_dev;
  c.absmax = absmax;
  c.delta = delta; c.tau = tau; c.lambda_g = lambda_g; c.thr = thr; c.mean_t = mean_t;
  c.inv_n = static_cast<float>(inv_n); c.data_term = data_term; c.relu_tanh = relu_tanh;
  const int nb = seed_blocks(n);
  igr_seeds_kernel<<<nb, 256, 0, stream>>>(f, g, y, n, c, fbar, gbar, sums, fbar_sum, part);
  if (part) {
    PartDst d{};
    for (int i = 0; i < 8; ++i) d.p[i] = sums + i;
    d.p[8] = fbar_sum;
    finalize_partials_kernel2<<<1, 32, 0, stream>>>(part, nb, 9, d);
  }
  return cudaGetLastError();
}
cudaError_t launch_label_stats(const float* y, int64_t n, float delta, double* out2, cudaStream_t stream) {
  label_stats_kernel<<<seed_blocks(n), 256, 0, stream>>>(y, n, delta, out2);
  return cudaGetLastError();
}
cudaError_t launch_wsl2_label_stats(const float* y, int64_t n, float delta, float wf, double* out4, cudaStream_t stream) {
  wsl2_label_stats_kernel<<<seed_blocks(n), 256, 0, stream>>>(y, n, delta, wf, out4);
  return cudaGetLastError();
}
cudaError_t launch_wsl2_seeds(const float* f, const float* y, int64_t n, float delta, float wf, int data_term, const double* stats,
                              int relu_tanh, unsigned* absmax, float* fbar, double* sums, double* fbar_sum, double* part,
                              cudaStream_t stream) {
  const int nb = seed_blocks(n);
  wsl2_seeds_kernel<<<nb, 256, 0, stream>>>(f, y, n, delta, wf, data_term, stats, relu_tanh, absmax, fbar, sums, fbar_sum, part);
  if (part) {
    PartDst d{};
    for (int i = 0; i < 8; ++i) d.p[i] = sums + i;
    d.p[8] = fbar_sum;
    finalize_partials_kernel2<<<1, 32, 0, stream>>>(part, nb, 9, d);
  }
  return cudaGetLastError();
}
cudaError_t launch_pcd_seeds(const float* f, const float* g, int64_t n, int which, float coef, int relu_tanh,
                             unsigned* absmax, float* fbar, float* gbar, double* sums, double* fbar_sum, double* part,
                             cudaStream_t stream) {
  const int nb = seed_blocks(n);
  pcd_seeds_kernel<<<nb, 256, 0, stream>>>(f, g, n, which, coef, relu_tanh, absmax, fbar, gbar, sums, fbar_sum, part);
  if (part) {
    PartDst d{};
    d.p[0] = sums + which;
    d.p[1] = fbar_sum;
    finalize_partials_kernel2<<<1, 32, 0, stream>>>(part, nb, 2, d);
  }
  return cudaGetLastError();
}
// ReLU/tanh network: turn seeds w.r.t. (f = tanh z, g = (1-f^2) G) into seeds w.r.t. (z, G), see igr_seeds_kernel.
__global__ void tanh_seeds_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ fbar_in,
                                  const float* __restrict__ gbar_in, long n, float* __restrict__ fbar, float* __restrict__ gbar) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float fi = f[i];
    const float d = 1.f - fi * fi;
    float fb = fbar_in ? fbar_in[i] * d : 0.f;
    if (gbar_in) {
      const float bx = gbar_in[3 * i], by = gbar_in[3 * i + 1], bz = gbar_in[3 * i + 2];
      fb -= 2.f * fi * (bx * g[3 * i] + by * g[3 * i + 1] + bz * g[3 * i + 2]);
      gbar[3 * i] = bx * d; gbar[3 * i + 1] = by * d; gbar[3 * i + 2] = bz * d;
    }
    fbar[i] = fb;
  }
}
cudaError_t launch_tanh_seeds(const float* f, const float* g, const float* fbar_in, const float* gbar_in, int64_t n,
                              float* fbar, float* gbar, cudaStream_t stream) {
  tanh_seeds_kernel<<<seed_blocks(n), 256, 0, stream>>>(f, g, fbar_in, gbar_in, n, fbar, gbar);
  return cudaGetLastError();
}
// max |v| over a seed array (generic backward), and the power-of-two scale derived from it
__global__ void absmax_kernel(const float* __restrict__ v, long n, unsigned* __restrict__ absmax) {
  float m = 0.f;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) m = fmaxf(m, fabsf(v[i]));
  block_absmax(m, absmax);
}
// scale[0] = 2^-ceil(log2(absmax)) (largest seed lands in (0.5, 1]), scale[1] = 1 / scale[0]; absmax is reset
// It also mirrors the running double-precision loss sums and sum(fbar) as (hi, lo) float pairs into the tail of the
// gradient buffer, so that ONE fp32 sum-allreduce over [accumulators | tail] carries gradients and loss sums.
__global__ void seed_scale_kernel(unsigned* __restrict__ absmax, float* __restrict__ scale, const double* __restrict__ sums,
                                  const double* __restrict__ fbar_sum, float* __restrict__ tail) {
  if (tail) {
    for (int i = 0; i < 9; ++i) {
      const double d = i < 8 ? (sums ? sums[i] : 0.0) : *fbar_sum;
      const float hi = static_cast<float>(d);
      tail[2 * i] = hi;
      tail[2 * i + 1] = static_cast<float>(d - static_cast<double>(hi));
    }
  }
  const float m = __uint_as_float(*absmax);
  int e = 0;
  float mult = 1.f;
  if (m > 0.f) {
    frexpf(m, &e);              // m = f * 2^e, f in [0.5, 1)
    e = e > 120 ? 120 : (e < -120 ? -120 : e);
    mult = ldexpf(1.f, -e);
  }
  scale[0] = mult;
  scale[1] = 1.f / mult;
  *absmax = 0u;
}
cudaError_t launch_absmax(const float* v, int64_t n, unsigned* absmax, cudaStream_t stream) {
  absmax_kernel<<<seed_blocks(n), 256, 0, stream>>>(v, n, absmax);
  return cudaGetLastError();
}
cudaError_t launch_seed_scale(unsigned* absmax, float* scale, const double* sums, const double* fbar_sum, float* tail,
                              cudaStream_t stream) {
  seed_scale_kernel<<<1, 1, 0, stream>>>(absmax, scale, sums, fbar_sum, tail);
  return cudaGetLastError();
}
cudaError_t launch_sum(const float* v, int64_t n, double* out, double* part, cudaStream_t stream) {
  const int nb = seed_blocks(n);
  sum_kernel<<<nb, 256, 0, stream>>>(v, n, out, part);
  if (part) {
    PartDst d{};
    d.p[0] = out;
    finalize_partials_kernel2<<<1, 32, 0, stream>>>(part, nb, 1, d);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Gradient unpack: accumulator tiles -> flat parameter-gradient vector in state_dict order
// (lin0.weight [out0,3], lin0.bias, lin1.weight [out1,H], ...; reference model/networks.py:62,77).
// ------------------------------------------------------------------------------------------------
struct GradLayoutDev {
  int H, L, skip;
  const float* main_acc; const float* zx_acc; const float* qg_acc; const float* as_acc; const float* adx_acc;
  const double* fbar_sum;
};

__device__ __forceinline__ float unpack_one(const GradLayoutDev& gl, const float* tail, long idx, int& layer, int& is_bias, long& off);

__global__ void unpack_grads_kernel(GradLayoutDev gl, const float* __restrict__ tail, float* __restrict__ grads, long num_params,
                                    float scale) {
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < num_params; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    int layer, is_bias;
    long off;
    grads[idx] = unpack_one(gl, tail, idx, layer, is_bias, off) * scale;
  }
}

cudaError_t launch_unpack_grads(int H, int L, int skip, const float* main_acc, const float* zx_acc, const float* qg_acc,
                                const float* as_acc, const float* adx_acc, const double* fbar_sum, const float* tail, long num_params,
                                float* grads, float scale, cudaStream_t stream) {
  GradLayoutDev gl{H, L, skip, main_acc, zx_acc, qg_acc, as_acc, adx_acc, fbar_sum};
  unpack_grads_kernel<<<1184, 256, 0, stream>>>(gl, tail, grads, num_params, scale);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults: no amsgrad, no weight decay; reference executor/executor.py:160,177)
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                        int step, cudaStream_t stream) {
  const float bc1 = 1.f - powf(b1, static_cast<float>(step));
  const float bc2 = 1.f - powf(b2, static_cast<float>(step));
  adam_kernel<<<1184, 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, sqrtf(bc2));
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// Fused optimizer tail (reference executor/executor.py:160,177: Adam over 18 tensors).
// ONE kernel reads the gradient accumulators (after the optional cross-rank sum), unpacks them into state_dict order,
// applies torch.optim.Adam's update to the fp32 parameters in place; ONE kernel then re-packs all weight images and the
// bias / last-layer vectors the next step's programs read.  Replaces unpack_grads + 18 multiplies + foreach Adam + 26
// pack launches of the generic path.
// ------------------------------------------------------------------------------------------------
struct ParamPtrs {
  float* W[kMaxLayers + 1];
  float* b[kMaxLayers + 1];
};

__device__ __forceinline__ float unpack_one(const GradLayoutDev& gl, const float* tail, long idx, int& layer, int& is_bias, long& off) {
  const int H = gl.H, L = gl.L;
  const float c = 0.70710678118654752440f;
  long rem = idx;
  float v = 0.f;
  for (int l = 0; l <= L; ++l) {
    const int in = l == 0 ? 3 : H;
    const int outd = l == L ? 1 : ((gl.skip > 0 && l + 1 == gl.skip) ? H - 3 : H);
    const long wsz = static_cast<long>(outd) * in;
    if (rem < wsz) {
      const int m = static_cast<int>(rem / in), n = static_cast<int>(rem % in);
      if (l == 0) {
        const float* zx = gl.zx_acc;
        v = zx[(n) * H + m] + zx[(n + 3) * H + m] + zx[(n + 6) * H + m] + gl.qg_acc[(n) * H + m] +
            gl.qg_acc[(n + 3) * H + m] + gl.qg_acc[(n + 6) * H + m];
      } else if (l == L) {
        v = gl.as_acc[0 * H + n] + gl.as_acc[1 * H + n] + gl.as_acc[2 * H + n] + gl.adx_acc[15 * H + n];
      } else {
        v = gl.main_acc[(static_cast<long>(l - 1) * H + n) * H + m];
        if (gl.skip > 0 && l == gl.skip) v *= c;
      }
      layer = l; is_bias = 0; off = rem;
      return v;
    }
    rem -= wsz;
    if (rem < outd) {
      const int m = static_cast<int>(rem);
      // d L / d b_L = sum fbar: from the (hi, lo) tail pair when the buffer went through an allreduce, else the double
      v = l == L ? (tail ? tail[16] + tail[17] : static_cast<float>(*gl.fbar_sum)) : gl.zx_acc[(static_cast<long>(l) * 16 + 15) * H + m];
      layer = l; is_bias = 1; off = rem;
      return v;
    }
    rem -= outd;
  }
  layer = 0; is_bias = 0; off = 0;
  return 0.f;
}

__global__ void fused_adam_kernel(GradLayoutDev gl, const float* __restrict__ tail, ParamPtrs pp, float* __restrict__ m,
                                  float* __restrict__ v, long num_params, float lr_bc1, float b1, float b2, float eps,
                                  float inv_bc2_sqrt, float scale, const float* __restrict__ scale_dev, float* __restrict__ grads_out) {
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < num_params; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    int layer, is_bias;
    long off;
    const float gi = unpack_one(gl, tail, idx, layer, is_bias, off) * sc;
    if (grads_out) grads_out[idx] = gi;
    float* p = (is_bias ? pp.b[layer] : pp.W[layer]) + off;
    const float mi = b1 * m[idx] + (1.f - b1) * gi;
    const float vi = b2 * v[idx] + (1.f - b2) * gi * gi;
    m[idx] = mi;
    v[idx] = vi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    *p -= lr_bc1 * (mi / denom);
  }
}

cudaError_t launch_fused_adam(int H, int L, int skip, const float* main_acc, const float* zx_acc, const float* qg_acc,
                              const float* as_acc, const float* adx_acc, const double* fbar_sum, const float* tail, long num_params,
                              float* const* W, float* const* b, float* m, float* v, float lr, float b1, float b2, float eps, int step,
                              float scale, const float* scale_dev, float* grads_out, cudaStream_t stream) {
  GradLayoutDev gl{H, L, skip, main_acc, zx_acc, qg_acc, as_acc, adx_acc, fbar_sum};
  ParamPtrs pp;
  for (int l = 0; l <= kMaxLayers; ++l) { pp.W[l] = l <= L ? W[l] : nullptr; pp.b[l] = l <= L ? b[l] : nullptr; }
  // bias corrections in double like torch.optim.Adam (python floats)
  const double bc1 = 1.0 - pow(static_cast<double>(b1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(b2), step);
  fused_adam_kernel<<<1184, 256, 0, stream>>>(gl, tail, pp, m, v, num_params, static_cast<float>(lr / bc1), b1, b2, eps,
                                              static_cast<float>(1.0 / sqrt(bc2)), scale, scale_dev, grads_out);
  return cudaGetLastError();
}

// All step images of one weight layout + the bias rows + the last-layer vector in ONE launch (PackAll: sdf_internal.h).
__device__ __forceinline__ void pack_one(const PackImg& im, long idx, int dtype, int halves) {
  const int hrows = im.ng / halves;
  long t = idx;
  const int e = static_cast<int>(t % 8); t /= 8;
  const int row = static_cast<int>(t % hrows); t /= hrows;
  const int gr = static_cast<int>(t % (2 * im.k16)); t /= (2 * im.k16);
  const int hf = static_cast<int>(t % halves); t /= halves;
  int kc = static_cast<int>(t % im.kchunks); t /= im.kchunks;
  const int g = static_cast<int>(t);
  const int n = g * im.ng + hf * hrows + row;
  bool lo_part = false;
  if (im.split) {
    lo_part = kc >= im.kchunks / 2;
    kc %= im.kchunks / 2;
  }
  const int k = (kc * 2 * im.k16 + gr) * 8 + e;
  const float* W = im.W;
  float v = 0.f;
  switch (im.kind) {
    case PK_F0:
      if (n < im.out_dim && k < 15) {
        const float w = W[n * im.in_dim + (k % 3)];
        const float hi = rnd16(w, dtype);
        v = (k < 9) ? hi : rnd16(w - hi, dtype);
      }
      break;
    case PK_FH:
      if (n < im.out_dim && k < im.in_dim) v = im.scale * W[static_cast<long>(n) * im.in_dim + k];
      break;
    case PK_FL:
      if (n == 0 && k < im.in_dim) v = W[k];
      break;
    case PK_RH:
      if (k < im.out_dim && n < im.in_dim) v = im.scale * W[static_cast<long>(k) * im.in_dim + n];
      break;
    case PK_R0:
      if (n < im.in_dim && k < im.out_dim) v = W[k * im.in_dim + n];
      break;
  }
  if (lo_part) v = v - rnd16(v, dtype);
  im.dst[idx] = bits16(v, dtype);
}

__global__ void pack_all_kernel(const __grid_constant__ PackAll pa) {
  const long tid = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long nthr = static_cast<long>(gridDim.x) * blockDim.x;
  for (int i = 0; i < pa.n_img; ++i) {
    const PackImg& im = pa.img[i];
    for (long idx = tid; idx < im.total; idx += nthr) pack_one(im, idx, pa.dtype, pa.halves);
  }
  if (pa.bias_dst) {
    const long nb = static_cast<long>(pa.L + 1) * pa.H;
    for (long idx = tid; idx < nb; idx += nthr) {
      const int l = static_cast<int>(idx / pa.H), j = static_cast<int>(idx % pa.H);
      pa.bias_dst[idx] = j < pa.b_out[l] ? pa.b[l][j] : 0.f;
    }
    for (long idx = tid; idx < pa.H; idx += nthr) pa.w_last_dst[idx] = pa.w_last_src[idx];
  }
}
cudaError_t launch_pack_all(const PackAll& pa, cudaStream_t stream) {
  pack_all_kernel<<<1184, 256, 0, stream>>>(pa);
  return cudaGetLastError();
}


// Loss value from the (hi, lo) tail of the gradient buffer (after the cross-rank sum): out4 = {total, term0, term1, term2}.
// kind 0: IGRLOSS (model/losses.py:118-136), kind 1: point-cloud loss (executor/executordistributed.py:73-79),
// kind 2: WeightedSmoothL2Loss (model/losses.py:50-69).
__global__ void loss_value_kernel(const float* __restrict__ tail, int kind, float tau, float lambda_g, int data_term, double inv_n,
                                  double inv_m, float* __restrict__ out4) {
  double s[8];
  for (int i = 0; i < 8; ++i) s[i] = static_cast<double>(tail[2 * i]) + static_cast<double>(tail[2 * i + 1]);
  if (kind == 0) {
    if (inv_n <= 0.0) inv_n = s[7] > 0.0 ? 1.0 / s[7] : 0.0;  // the count travels with the sums (global after the allreduce)
    const double data = data_term == 0 ? s[1] * inv_n + s[3] * inv_n - 2.0 * (s[0] * inv_n) * (s[2] * inv_n) : s[4] * inv_n;
    const double reg = s[5] * inv_n, eik = s[6] * inv_n;
    out4[0] = static_cast<float>(data + tau * reg + lambda_g * eik);
    out4[1] = static_cast<float>(data); out4[2] = static_cast<float>(reg); out4[3] = static_cast<float>(eik);
  } else if (kind == 2) {
    // WeightedSmoothL2Loss (model/losses.py:50-69): sums = {sum p, sum p^2, sum w (t-p)^2, sum w, sum w t, sum w t^2, -, N}
    if (inv_n <= 0.0) inv_n = s[7] > 0.0 ? 1.0 / s[7] : 0.0;
    const double v = data_term == 0 ? s[5] * inv_n - 2.0 * (s[0] * inv_n) * (s[4] * inv_n) + (s[1] * inv_n) * (s[3] * inv_n) : s[2] * inv_n;
    out4[0] = static_cast<float>(v);
    out4[1] = static_cast<float>(v); out4[2] = 0.f; out4[3] = 0.f;
  } else {
    const double surf = s[0] * inv_n, eik = s[1] * inv_m;
    out4[0] = static_cast<float>(surf + lambda_g * eik);
    out4[1] = static_cast<float>(surf); out4[2] = static_cast<float>(eik); out4[3] = 0.f;
  }
}
cudaError_t launch_loss_value(const float* tail, int kind, float tau, float lambda_g, int data_term, double inv_n, double inv_m,
                              float* out4, cudaStream_t stream) {
  loss_value_kernel<<<1, 1, 0, stream>>>(tail, kind, tau, lambda_g, data_term, inv_n, inv_m, out4);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Octree refinement flags, marching-cubes cell classification, ordered stream compaction.
// Integer / predicate work: bit-exact against oracle/sdf_oracle.py (octree_flags / octree_children / mc_cases) given equal SDF inputs.
// ------------------------------------------------------------------------------------------------
__global__ void octree_flags_kernel(const float* __restrict__ sdf, long n, float thresh, uint8_t* __restrict__ flags) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    flags[i] = fabsf(sdf[i]) <= thresh ? 1 : 0;
}
cudaError_t launch_octree_flags(const float* sdf, int64_t n, float thresh, uint8_t* flags, cudaStream_t stream) {
  octree_flags_kernel<<<seed_blocks(n), 256, 0, stream>>>(sdf, n, thresh, flags);
  return cudaGetLastError();
}

// corner c of cell (i,j,k): Lorensen-Cline order (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),(1,0,1),(1,1,1),(0,1,1)
// One thread classifies 4 consecutive cells along z of one (ix, iy) cell row: blockIdx.z = ix, blockIdx.y = iy (no
// index divisions), 4 x (float4 + 1) loads of the four corner rows, 4 byte stores.  HBM-bound: ~5 B per cell.
__global__ void mc_classify_kernel(const float* __restrict__ sdf, int nx, int ny, int nz, uint8_t* __restrict__ cases) {
  const int cy = ny - 1, cz = nz - 1;
  const int i = blockIdx.z, j = blockIdx.y;
  const int k0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (k0 >= cz) return;
  const long sy = nz, sx = static_cast<long>(ny) * nz;
  const float* b = sdf + (static_cast<long>(i) * ny + j) * nz + k0;
  const long row_off[4] = {0, sx, sx + sy, sy};  // corners 0..3 (z = k), corners 4..7 are the same rows at z = k + 1
  unsigned neg[4];                               // bit t of neg[r]: row r value at z = k0 + t is negative (t = 0..4)
  const bool vec = (nz & 3) == 0 && k0 + 4 < nz;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* p = b + row_off[r];
    float v[5];
    if (vec) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      v[4] = __ldg(p + 4);
    } else {
#pragma unroll
      for (int t = 0; t < 5; ++t) v[t] = (k0 + t < nz) ? __ldg(p + t) : 0.f;
    }
    unsigned m = 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) m |= (v[t] < 0.f ? 1u : 0u) << t;
    neg[r] = m;
  }
  uint8_t* out = cases + (static_cast<long>(i) * cy + j) * cz + k0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (k0 + t < cz) {
      unsigned m = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) m |= ((neg[r] >> t) & 1u) << r | ((neg[r] >> (t + 1)) & 1u) << (r + 4);
      out[t] = static_cast<uint8_t>(m);
    }
  }
}

// Edge vertices of the active cells (the interpolation stage of marching cubes, level 0).  One thread per (cell, edge);
// edge e of a cell joins corners kEdgeA[e], kEdgeB[e] (Lorensen-Cline numbering, as the corner order above).  A crossing
// edge (corner signs differ) gets t = f_a / (f_a - f_b) and the point (idx_a + t (idx_b - idx_a)) * spacing -- index space
// times spacing, the convention of skimage.measure.marching_cubes(volume, 0.0, spacing=...) at the reference call site.
// Every operation is a single correctly rounded fp32 operation (no contraction), so the oracle reproduces the bits.
__constant__ int kEdgeA[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__constant__ int kEdgeB[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
__constant__ int kCornerD[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
__global__ void mc_edge_vertices_kernel(const float* __restrict__ sdf, int nx, int ny, int nz, const long* __restrict__ cells, long n_cells,
                                        float sx, float sy, float sz, float* __restrict__ verts, unsigned short* __restrict__ mask) {
  const long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long ci = t / 12;
  const int e = static_cast<int>(t - ci * 12);
  if (ci >= n_cells) return;
  const long cell = cells[ci];
  const int cy = ny - 1, cz = nz - 1;
  const int k = static_cast<int>(cell % cz);
  const int j = static_cast<int>((cell / cz) % cy);
  const int i = static_cast<int>(cell / (static_cast<long>(cz) * cy));
  const int a = kEdgeA[e], b = kEdgeB[e];
  const int ia[3] = {i + kCornerD[a][0], j + kCornerD[a][1], k + kCornerD[a][2]};
  const int ib[3] = {i + kCornerD[b][0], j + kCornerD[b][1], k + kCornerD[b][2]};
  const float fa = __ldg(sdf + (static_cast<long>(ia[0]) * ny + ia[1]) * nz + ia[2]);
  const float fb = __ldg(sdf + (static_cast<long>(ib[0]) * ny + ib[1]) * nz + ib[2]);
  const bool cross = (fa < 0.f) != (fb < 0.f);
  float v[3] = {0.f, 0.f, 0.f};
  if (cross) {
    const float tt = __fdiv_rn(fa, __fsub_rn(fa, fb));
    const float sp[3] = {sx, sy, sz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float pa = static_cast<float>(ia[d]), pb = static_cast<float>(ib[d]);
      v[d] = __fmul_rn(__fadd_rn(pa, __fmul_rn(tt, __fsub_rn(pb, pa))), sp[d]);
    }
  }
  float* o = verts + t * 3;
  o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
  if (e == 0) {
    // 12-bit mask of the crossing edges of the cell = edges whose corner bits differ in the 8-bit corner sign mask
    unsigned neg = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      neg |= (__ldg(sdf + (static_cast<long>(i + kCornerD[c][0]) * ny + j + kCornerD[c][1]) * nz + k + kCornerD[c][2]) < 0.f ? 1u : 0u) << c;
    unsigned m = 0;
#pragma unroll
    for (int q = 0; q < 12; ++q) m |= (((neg >> kEdgeA[q]) ^ (neg >> kEdgeB[q])) & 1u) << q;
    mask[ci] = static_cast<unsigned short>(m);
  }
}
cudaError_t launch_mc_edge_vertices(const float* sdf, int nx, int ny, int nz, const long* cells, long n_cells, float sx, float sy, float sz,
                                    float* verts, unsigned short* mask, cudaStream_t stream) {
  if (n_cells <= 0) return cudaSuccess;
  const long threads = n_cells * 12;
  const long blocks = (threads + 383) / 384;  // 384 = 32 cells per block
  mc_edge_vertices_kernel<<<static_cast<unsigned>(blocks), 384, 0, stream>>>(sdf, nx, ny, nz, cells, n_cells, sx, sy, sz, verts, mask);
  return cudaGetLastError();
}

// Ordered compaction in three passes: per-block counts, scan of the block counts (one block), scatter.
constexpr int kScanBlock = 1024;
constexpr int kItemsPerThread = 4;
constexpr int kTileItems = kScanBlock * kItemsPerThread;

// mode 0: flag = cases[i] not in {0,255};  mode 1: flag = flags[i] != 0
__device__ __forceinline__ int compact_flag(const uint8_t* v, long i, int mode) {
  const unsigned c = v[i];
  return mode == 0 ? (c != 0u && c != 255u) : (c != 0u);
}
__global__ void compact_count_kernel(const uint8_t* __restrict__ v, long n, int mode, long* __restrict__ block_counts) {
  __shared__ int wsum[32];
  const long base = static_cast<long>(blockIdx.x) * kTileItems;
  int c = 0;
  for (int t = 0; t < kItemsPerThread; ++t) {
    const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
    if (i < n) c += compact_flag(v, i, mode);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = wsum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = w;
  }
}
__global__ void compact_scan_kernel(long* __restrict__ block_counts, long nblocks, long* __restrict__ total) {
  // single block, sequential over chunks of blockDim.x; exclusive scan in place
  __shared__ long sh[kScanBlock];
  __shared__ long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long b0 = 0; b0 < nblocks; b0 += kScanBlock) {
    const long i = b0 + threadIdx.x;
    const long x = i < nblocks ? block_counts[i] : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
      long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) block_counts[i] = carry + sh[threadIdx.x] - x;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[kScanBlock - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
// mode 0: out[rank] = i (active cell index);  mode 1: out[8*rank + c] = (keys[i] << 3) | c
__global__ void compact_scatter_kernel(const uint8_t* __restrict__ v, long n, int mode, const long* __restrict__ block_offsets,
                                       const uint64_t* __restrict__ keys, long* __restrict__ out_idx,
                                       uint64_t* __restrict__ out_keys) {
  __shared__ int wsum[32];
  const long base = static_cast<long>(blockIdx.x) * kTileItems;
  int fl[kItemsPerThread];
  int c = 0;
  for (int t = 0; t < kItemsPerThread; ++t) {
    const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
    fl[t] = i < n ? compact_flag(v, i, mode) : 0;
    c += fl[t];
  }
  // exclusive scan of c over the block
  int incl = c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    int wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    wsum[lane] = wi - w;
  }
  __syncthreads();
  long pos = block_offsets[blockIdx.x] + wsum[warp] + (incl - c);
  for (int t = 0; t < kItemsPerThread; ++t) {
    if (fl[t]) {
      const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
      if (mode == 0) {
        out_idx[pos] = i;
      } else {
        const uint64_t k = keys[i];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) out_keys[pos * 8 + ch] = (k << 3) | static_cast<uint64_t>(ch);
      }
      ++pos;
    }
  }
}

cudaError_t launch_mc_classify(const float* sdf, int nx, int ny, int nz, uint8_t* cases, cudaStream_t stream) {
  const int cz = nz - 1;
  const int threads = 128;
  dim3 grid((cz + 4 * threads - 1) / (4 * threads), ny - 1, nx - 1);
  mc_classify_kernel<<<grid, threads, 0, stream>>>(sdf, nx, ny, nz, cases);
  return cudaGetLastError();
}
// total number of flagged entries only (first two passes of the compaction)
cudaError_t launch_count_flags(const uint8_t* v, long n, int mode, long* total, long* scratch, cudaStream_t stream);
long compact_num_blocks(long n) { return (n + kTileItems - 1) / kTileItems; }
cudaError_t launch_compact(const uint8_t* v, long n, int mode, const uint64_t* keys, long* out_idx, uint64_t* out_keys,
                           long* total, long* scratch, cudaStream_t stream) {
  const long nb = compact_num_blocks(n);
  if (nb == 0) return cudaMemsetAsync(total, 0, sizeof(long), stream);
  compact_count_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, stream>>>(v, n, mode, scratch);
  compact_scan_kernel<<<1, kScanBlock, 0, stream>>>(scratch, nb, total);
  compact_scatter_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, stream>>>(v, n, mode, scratch, keys, out_idx, out_keys);
  return cudaGetLastError();
}
cudaError_t launch_count_flags(const uint8_t* v, long n, int mode, long* total, long* scratch, cudaStream_t stream) {
  const long nb = compact_num_blocks(n);
  if (nb == 0) return cudaMemsetAsync(total, 0, sizeof(long), stream);
  compact_count_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, stream>>>(v, n, mode, scratch);
  compact_scan_kernel<<<1, kScanBlock, 0, stream>>>(scratch, nb, total);
  return cudaGetLastError();
}

// ---- octree level loop without host round trips: the cell count of a level lives in DEVICE memory -------------------
// lvl = {count, offset} of the level inside the key / sdf / flag arenas; the children of its flagged cells become the
// next level, appended behind it: lvl_next = {8 * flagged, offset + count rounded up to even} (clamped to the arena capacity; *overflow = 1
// then).  Same three ordered passes as launch_compact, each block looping over the tiles the device-side count yields.
__global__ void octree_count_dev_kernel(const uint8_t* __restrict__ flags, const int64_t* __restrict__ lvl, long* __restrict__ block_counts) {
  __shared__ int wsum[32];
  const long n = lvl[0];
  const uint8_t* v = flags + lvl[1];
  const long ntiles = (n + kTileItems - 1) / kTileItems;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long base = tile * kTileItems;
    int c = 0;
    for (int t = 0; t < kItemsPerThread; ++t) {
      const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
      if (i < n) c += v[i] != 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
      if (threadIdx.x == 0) block_counts[tile] = w;
    }
    __syncthreads();
  }
}
__global__ void octree_scan_dev_kernel(long* __restrict__ block_counts, const int64_t* __restrict__ lvl, int64_t* __restrict__ lvl_next,
                                       long capacity, int* __restrict__ overflow) {
  __shared__ long sh[kScanBlock];
  __shared__ long carry;
  const long n = lvl[0];
  const long nblocks = (n + kTileItems - 1) / kTileItems;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long b0 = 0; b0 < nblocks; b0 += kScanBlock) {
    const long i = b0 + threadIdx.x;
    const long x = i < nblocks ? block_counts[i] : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int o = 1; o < kScanBlock; o <<= 1) {
      long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) block_counts[i] = carry + sh[threadIdx.x] - x;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[kScanBlock - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long off_next = (lvl[1] + n + 1) & ~1L;  // even: the 16-byte key stores of the scatter stay aligned
    long children = 8 * carry;
    if (off_next + children > capacity) {
      children = off_next < capacity ? (capacity - off_next) / 8 * 8 : 0;
      *overflow = 1;
    }
    lvl_next[0] = children;
    lvl_next[1] = off_next;
  }
}
__global__ void octree_scatter_dev_kernel(const uint8_t* __restrict__ flags, uint64_t* __restrict__ keys, const int64_t* __restrict__ lvl,
                                          const int64_t* __restrict__ lvl_next, const long* __restrict__ block_offsets) {
  __shared__ int wsum[32];
  const long n = lvl[0];
  const uint8_t* v = flags + lvl[1];
  const uint64_t* kin = keys + lvl[1];
  uint64_t* kout = keys + lvl_next[1];
  const long max_cells = lvl_next[0] / 8;
  const long ntiles = (n + kTileItems - 1) / kTileItems;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long base = tile * kTileItems;
    int fl[kItemsPerThread];
    int c = 0;
    for (int t = 0; t < kItemsPerThread; ++t) {
      const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
      fl[t] = i < n ? (v[i] != 0) : 0;
      c += fl[t];
    }
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
      int wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    long pos = block_offsets[tile] + wsum[warp] + (incl - c);
    for (int t = 0; t < kItemsPerThread; ++t) {
      if (fl[t]) {
        if (pos < max_cells) {
          const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
          const uint64_t k = kin[i];
          // 8 consecutive child keys = 64 bytes: four 16-byte stores
          ulonglong2* o2 = reinterpret_cast<ulonglong2*>(kout + pos * 8);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) o2[ch] = make_ulonglong2((k << 3) | static_cast<uint64_t>(2 * ch), (k << 3) | static_cast<uint64_t>(2 * ch + 1));
        }
        ++pos;
      }
    }
    __syncthreads();
  }
}
cudaError_t launch_octree_refine_dev(const uint8_t* flags, uint64_t* keys, const int64_t* lvl, int64_t* lvl_next, long* scratch,
                                     long level_capacity, long capacity, int* overflow, cudaStream_t stream) {
  long nb = compact_num_blocks(level_capacity);
  if (nb < 1) nb = 1;
  if (nb > 1184) nb = 1184;
  octree_count_dev_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, stream>>>(flags, lvl, scratch);
  octree_scan_dev_kernel<<<1, kScanBlock, 0, stream>>>(scratch, lvl, lvl_next, capacity, overflow);
  octree_scatter_dev_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, stream>>>(flags, keys, lvl, lvl_next, scratch);
  return cudaGetLastError();
}

// ---- marching-cubes faces: the triangles of the active cells, ordered (cell order, then table order) ------------------
// Triangle t of active cell r (case index c) has the vertices r * 12 + kMcTriTable[c][3 t + k], k = 0..2 -- indices into
// the [n_active][12] edge-vertex array of mc_edge_vertices_kernel.  Count pass per tile of cells, scan of the tile counts
// (compact_scan_kernel), ordered write pass.  Table: mc_tables.h (generated, tools/make_mc_table.py).
__global__ void mc_tri_count_kernel(const uint8_t* __restrict__ cases, const long* __restrict__ cells, long n, long* __restrict__ block_counts) {
  __shared__ int wsum[32];
  const long base = static_cast<long>(blockIdx.x) * kTileItems;
  int c = 0;
  for (int t = 0; t < kItemsPerThread; ++t) {
    const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
    if (i < n) c += kMcTriCount[cases[cells[i]]];
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = wsum[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = w;
  }
}
__global__ void mc_tri_write_kernel(const uint8_t* __restrict__ cases, const long* __restrict__ cells, long n,
                                    const long* __restrict__ block_offsets, long* __restrict__ tri) {
  __shared__ int wsum[32];
  const long base = static_cast<long>(blockIdx.x) * kTileItems;
  int cs[kItemsPerThread];
  int c = 0;
  for (int t = 0; t < kItemsPerThread; ++t) {
    const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
    cs[t] = i < n ? cases[cells[i]] : 0;
    c += kMcTriCount[cs[t]];
  }
  int incl = c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
    int wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    wsum[lane] = wi - w;
  }
  __syncthreads();
  long pos = block_offsets[blockIdx.x] + wsum[warp] + (incl - c);
  for (int t = 0; t < kItemsPerThread; ++t) {
    const long i = base + static_cast<long>(threadIdx.x) * kItemsPerThread + t;
    const int nt = kMcTriCount[cs[t]];
    for (int q = 0; q < nt; ++q) {
#pragma unroll
      for (int k = 0; k < 3; ++k) tri[(pos + q) * 3 + k] = i * 12 + kMcTriTable[cs[t]][3 * q + k];
    }
    pos += nt;
  }
}
cudaError_t launch_mc_triangles(const uint8_t* cases, const long* cells, long n, long* tri, long* total, long* scratch, cudaStream_t stream) {
  const long nb = compact_num_blocks(n);
  if (nb == 0) return cudaMemsetAsync(total, 0, sizeof(long), stream);
  mc_tri_count_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, stream>>>(cases, cells, n, scratch);
  compact_scan_kernel<<<1, kScanBlock, 0, stream>>>(scratch, nb, total);
  if (tri) mc_tri_write_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, stream>>>(cases, cells, n, scratch, tri);
  return cudaGetLastError();
}

}  // namespace sdf
