// Stand-alone probe for the UMMA shared-memory descriptor conventions this library relies on.
// One CTA, one tcgen05.mma chain over a single 128 x N tile; operands are laid out on the host in the
// no-swizzle "granule-major" image ([granule of 8 elements][row][16 B]) and the result is checked against a
// CPU product.  Descriptor fields come from the command line so one GPU session can test the intended
// encoding and its alternatives, each in its own process:
//
//   umma_probe <mode> <N> <K> <bf16> <swap>
//     mode 0: A, B K-major   (layer chain:  D[128 pts, N] = A[pts, K] * W[N, K]^T)
//     mode 1: A, B MN-major  (weight grad:  D[128 f_out, N f_in] = sum_k A[k pts, f_out] * B[k pts, f_in])
//     swap 1: exchange the LBO/SBO fields (the alternative reading of the descriptor)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "../sdf-representation_b200/csrc/ptx.cuh"

using namespace sdf;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                   \
    }                                                                            \
  } while (0)

struct ProbeArgs {
  int N, K;
  uint32_t a_bytes, b_bytes;
  uint32_t lbo_a, sbo_a, step_a;  // step = start-address advance per K=16 instruction
  uint32_t lbo_b, sbo_b, step_b;
  uint32_t idesc;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint4* __restrict__ a_img, const uint4* __restrict__ b_img,
                                                       float* __restrict__ d_out, ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.a_bytes;
  for (uint32_t i = threadIdx.x; i < p.a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = a_img[i];
  for (uint32_t i = threadIdx.x; i < p.b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = b_img[i];
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) {
    const int ksteps = p.K / 16;
    for (int k = 0; k < ksteps; ++k) {
      uint64_t da = make_smem_desc(smem_u32(sA) + k * p.step_a, p.lbo_a, p.sbo_a);
      uint64_t db = make_smem_desc(smem_u32(sB) + k * p.step_b, p.lbo_b, p.sbo_b);
      umma_f16_ss(tmem_base, da, db, p.idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d_out[row * p.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_base);
}

static uint16_t to16(float v, bool bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
static float from16(uint16_t u, bool bf16) {
  if (bf16) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<__half*>(&u));
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  int N = argc > 2 ? atoi(argv[2]) : 256;
  int K = argc > 3 ? atoi(argv[3]) : 64;
  bool bf16 = argc > 4 ? atoi(argv[4]) != 0 : true;
  int swap = argc > 5 ? atoi(argv[5]) : 0;
  const int M = 128;
  // Logical operands: A[M][K], B[N][K]  (D = A * B^T), small integers/fractions exactly representable.
  std::vector<float> A(M * K), B(N * K);
  srand(1234);
  for (auto& v : A) v = static_cast<float>((rand() % 17) - 8) / 8.0f;
  for (auto& v : B) v = static_cast<float>((rand() % 13) - 6) / 4.0f;
  std::vector<uint16_t> aimg(M * K), bimg(N * K);
  ProbeArgs p{};
  p.N = N;
  p.K = K;
  p.a_bytes = M * K * 2;
  p.b_bytes = N * K * 2;
  if (mode == 0) {
    // K-major granule image: element (r, k) at ((k/8)*R + r)*8 + k%8 (in halves); R = rows of the operand.
    for (int r = 0; r < M; ++r)
      for (int k = 0; k < K; ++k) aimg[((k / 8) * M + r) * 8 + k % 8] = to16(A[r * K + k], bf16);
    for (int r = 0; r < N; ++r)
      for (int k = 0; k < K; ++k) bimg[((k / 8) * N + r) * 8 + k % 8] = to16(B[r * K + k], bf16);
    p.lbo_a = M * 16;  // next K granule
    p.sbo_a = 128;     // next 8-row group
    p.step_a = 2 * M * 16;
    p.lbo_b = N * 16;
    p.sbo_b = 128;
    p.step_b = 2 * N * 16;
    p.idesc = make_idesc_f16(M, N, bf16, false, false);
  } else {
    // MN-major granule image: rows of the image are K (points); element (mn, k) at ((mn/8)*K + k)*8 + mn%8.
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < K; ++k) aimg[((m / 8) * K + k) * 8 + m % 8] = to16(A[m * K + k], bf16);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) bimg[((n / 8) * K + k) * 8 + n % 8] = to16(B[n * K + k], bf16);
    p.lbo_a = 128;     // next 8-row K group
    p.sbo_a = K * 16;  // next 8-element MN block
    p.step_a = 256;    // 16 K rows
    p.lbo_b = 128;
    p.sbo_b = K * 16;
    p.step_b = 256;
    p.idesc = make_idesc_f16(M, N, bf16, true, true);
  }
  if (swap) {
    std::swap(p.lbo_a, p.sbo_a);
    std::swap(p.lbo_b, p.sbo_b);
  }
  uint4 *da, *db;
  float* dd;
  CK(cudaMalloc(&da, p.a_bytes));
  CK(cudaMalloc(&db, p.b_bytes));
  CK(cudaMalloc(&dd, M * N * sizeof(float)));
  CK(cudaMemcpy(da, aimg.data(), p.a_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bimg.data(), p.b_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xFF, M * N * sizeof(float)));
  size_t smem = p.a_bytes + p.b_bytes;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(da, db, dd, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(M * N);
  CK(cudaMemcpy(D.data(), dd, M * N * sizeof(float), cudaMemcpyDeviceToHost));
  double max_err = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k)
        ref += (double)from16(to16(A[m * K + k], bf16), bf16) * (double)from16(to16(B[n * K + k], bf16), bf16);
      double err = fabs(ref - (double)D[m * N + n]);
      if (!(err <= 1e-3)) ++bad;
      if (err > max_err || std::isnan(err)) max_err = err;
    }
  printf("probe mode=%d N=%d K=%d bf16=%d swap=%d : %s (max_err %.3g, bad %d / %d)\n", mode, N, K, (int)bf16, swap,
         bad == 0 ? "PASS" : "FAIL", max_err, bad, M * N);
  return bad == 0 ? 0 : 1;
}
