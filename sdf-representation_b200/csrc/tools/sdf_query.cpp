// sdf_query -- native SDF query tool on top of the C ABI (include/sdfnet.h).
//
// B200-native equivalent of the reference's stand-alone C++ query programs: ops/DeepTrace/src/deeptrace.cpp:16-103
// (ONNX Runtime, one session.Run per point) and ops/conversion_test/test_loading.cpp:48-130 (libtorch forward +
// autograd::grad): read points from a CSV, evaluate the network, write the SDF values and the input gradients as CSV.
// Here every point goes through one batched launch of the fused forward + input-gradient kernel.
//
//   sdf_query <weights.sdfw> <input.csv> <output.csv> [gradient.csv] [--fast] [--bf16]
//
// weights.sdfw (written by sdf_representation_b200.checkpoint.export_flat): little endian
//   char magic[8] = "SDFW0001"; int32 d_in, hidden, num_hidden, skip_layer; float beta;
//   then for i = 0..num_hidden: float weight[out_i * in_i] (row-major [out,in]), float bias[out_i].
// input.csv: one point per line, "x,y,z" (a header line or an index column as written by pandas is skipped, like
// the reference's readers do).  output.csv: one value per line; gradient.csv: "gx,gy,gz".
// Default precision is the hi/lo split ("parity", fp32-grade); --fast selects single-pass 16-bit operands.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "../../../include/sdfnet.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "sdf_query: %s: %s\n", #x, cudaGetErrorString(e_));                       \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)
#define SK(x)                                                                \
  do {                                                                       \
    if ((x) != 0) {                                                          \
      fprintf(stderr, "sdf_query: %s: %s\n", #x, sdfnet_last_error());       \
      return 3;                                                              \
    }                                                                        \
  } while (0)

static bool parse_point(const std::string& line, float (&p)[3]) {
  std::vector<double> v;
  std::stringstream ss(line);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    char* end = nullptr;
    const double d = strtod(tok.c_str(), &end);
    if (end == tok.c_str()) return false;  // header or malformed line
    v.push_back(d);
  }
  if (v.size() < 3) return false;
  const size_t o = v.size() - 3;  // pandas index column(s) come first
  for (int j = 0; j < 3; ++j) p[j] = static_cast<float>(v[o + j]);
  return true;
}

int main(int argc, char** argv) {
  std::vector<const char*> pos;
  int precision = 1, dtype = 1;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--fast")) precision = 0;
    else if (!strcmp(argv[i], "--bf16")) dtype = 0;
    else pos.push_back(argv[i]);
  }
  if (pos.size() < 3) {
    fprintf(stderr, "usage: sdf_query <weights.sdfw> <input.csv> <output.csv> [gradient.csv] [--fast] [--bf16]\n");
    return 1;
  }
  // ---- weights
  std::ifstream wf(pos[0], std::ios::binary);
  if (!wf) { fprintf(stderr, "sdf_query: cannot open %s\n", pos[0]); return 1; }
  char magic[8];
  int32_t hdr[4];
  float beta;
  wf.read(magic, 8);
  wf.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
  wf.read(reinterpret_cast<char*>(&beta), sizeof(beta));
  if (!wf || memcmp(magic, "SDFW0001", 8) != 0) { fprintf(stderr, "sdf_query: %s is not an SDFW0001 file\n", pos[0]); return 1; }
  sdfnet_desc desc{};
  desc.d_in = hdr[0]; desc.hidden = hdr[1]; desc.num_hidden = hdr[2]; desc.skip_layer = hdr[3];
  desc.beta = beta; desc.dtype = dtype; desc.precision = precision; desc.device = 0;
  const int L = desc.num_hidden, H = desc.hidden;
  std::vector<float*> dW(L + 1), dB(L + 1);
  for (int l = 0; l <= L; ++l) {
    const int in = l == 0 ? desc.d_in : H;
    const int out = l == L ? 1 : ((desc.skip_layer > 0 && l + 1 == desc.skip_layer) ? H - desc.d_in : H);
    std::vector<float> w(static_cast<size_t>(out) * in), b(out);
    wf.read(reinterpret_cast<char*>(w.data()), w.size() * sizeof(float));
    wf.read(reinterpret_cast<char*>(b.data()), b.size() * sizeof(float));
    if (!wf) { fprintf(stderr, "sdf_query: %s is truncated (layer %d)\n", pos[0], l); return 1; }
    CK(cudaMalloc(&dW[l], w.size() * sizeof(float)));
    CK(cudaMalloc(&dB[l], b.size() * sizeof(float)));
    CK(cudaMemcpy(dW[l], w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB[l], b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  // ---- points
  std::ifstream pf(pos[1]);
  if (!pf) { fprintf(stderr, "sdf_query: cannot open %s\n", pos[1]); return 1; }
  std::vector<float> pts;
  std::string line;
  while (std::getline(pf, line)) {
    float p[3];
    if (parse_point(line, p)) pts.insert(pts.end(), p, p + 3);
  }
  const int64_t n = static_cast<int64_t>(pts.size() / 3);
  // ---- evaluate
  sdfnet_t net = nullptr;
  SK(sdfnet_create(&desc, &net));
  SK(sdfnet_pack_weights(net, dW.data(), dB.data(), nullptr));
  float *dx = nullptr, *df = nullptr, *dg = nullptr;
  void* ws = nullptr;
  std::vector<float> f(n), g(3 * n);
  if (n > 0) {
    const size_t ws_bytes = sdfnet_workspace_bytes(net, n, 0);
    CK(cudaMalloc(&dx, pts.size() * sizeof(float)));
    CK(cudaMalloc(&df, n * sizeof(float)));
    CK(cudaMalloc(&dg, 3 * n * sizeof(float)));
    CK(cudaMalloc(&ws, ws_bytes));
    CK(cudaMemcpy(dx, pts.data(), pts.size() * sizeof(float), cudaMemcpyHostToDevice));
    SK(sdfnet_forward_grad(net, dx, n, df, dg, ws, ws_bytes, nullptr));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(f.data(), df, n * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g.data(), dg, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  // ---- write
  FILE* of = fopen(pos[2], "w");
  if (!of) { fprintf(stderr, "sdf_query: cannot write %s\n", pos[2]); return 1; }
  for (int64_t i = 0; i < n; ++i) fprintf(of, "%.9g\n", f[i]);
  fclose(of);
  if (pos.size() > 3) {
    FILE* gf = fopen(pos[3], "w");
    if (!gf) { fprintf(stderr, "sdf_query: cannot write %s\n", pos[3]); return 1; }
    for (int64_t i = 0; i < n; ++i) fprintf(gf, "%.9g,%.9g,%.9g\n", g[3 * i], g[3 * i + 1], g[3 * i + 2]);
    fclose(gf);
  }
  printf("sdf_query: %lld points, %s precision\n", static_cast<long long>(n), precision ? "parity" : "fast");
  cudaFree(dx); cudaFree(df); cudaFree(dg); cudaFree(ws);
  for (int l = 0; l <= L; ++l) { cudaFree(dW[l]); cudaFree(dB[l]); }
  sdfnet_destroy(net);
  return 0;
}
