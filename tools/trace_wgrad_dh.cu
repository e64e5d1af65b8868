// Debug tool: in-kernel timeline of the generation-4 weight-gradient kernel (wgrad_dh).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DM2_TRACE -Im2_mixer_b200/csrc
//        tools/trace_wgrad_dh.cu m2_mixer_b200/csrc/wgrad_fused.cu m2_mixer_b200/csrc/chain_ts.cu m2_mixer_b200/csrc/chain.cu m2_mixer_b200/csrc/profile.cu -o tools/trace_wgrad_dh.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
namespace m2 {
int wgrad_dh(const void* xn_b, const void* dy_b, const void* dh_b, int ldh, const void* w1b, const float* b1, float* dw1,
             float* db1, float* dw2, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s);
int wgrad_trace_read(long long* host, int n);
}
int main(int argc, char** argv) {
  const int M = 16384, D = 128, C = 3072;
  const float p = argc > 1 ? atof(argv[1]) : 0.f;
  float *b1, *dw1, *dw2, *db1; void *w1b, *xnb, *dyb, *dhb;
  cudaMalloc(&b1, C * 4); cudaMalloc(&dw1, C * D * 4); cudaMalloc(&dw2, C * D * 4); cudaMalloc(&db1, C * 4);
  cudaMalloc(&w1b, C * D * 2); cudaMalloc(&xnb, M * D * 2); cudaMalloc(&dyb, M * D * 2); cudaMalloc(&dhb, (size_t)M * C * 2);
  cudaMemset(b1, 0, C * 4); cudaMemset(dw1, 0, C * D * 4); cudaMemset(dw2, 0, C * D * 4); cudaMemset(db1, 0, C * 4);
  cudaMemset(dhb, 0, (size_t)M * C * 2);
  std::vector<__nv_bfloat16> w(C * D), x(M * D);
  for (auto& v : w) v = __float2bfloat16((rand() % 2001 - 1000) * 1e-4f);
  for (auto& v : x) v = __float2bfloat16((rand() % 2001 - 1000) * 1e-3f);
  cudaMemcpy(w1b, w.data(), C * D * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(xnb, x.data(), M * D * 2, cudaMemcpyHostToDevice); cudaMemcpy(dyb, x.data(), M * D * 2, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 3; ++rep) {
    int rc = m2::wgrad_dh(xnb, dyb, dhb, C, w1b, b1, dw1, db1, dw2, M, D, C, p, 1234, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("rc=%d err=%s\n", rc, cudaGetErrorString(e)); return 1; }
  }
  std::vector<long long> t(4096);
  m2::wgrad_trace_read(t.data(), 4096);
  const char* names[] = {"", "hg:wait", "hg:go", "hg:issued", "wg:wait", "wg:go1", "wg:issued", "epi:waitH", "epi:H", "epi:hempty",
                         "epi:gfull", "wg:go2", "epi:gempty", "epi:half"};
  long long t0 = 0;
  for (int s = 0; s < 1360; ++s) if (t[3 * s + 2] && (t0 == 0 || t[3 * s + 2] < t0)) t0 = t[3 * s + 2];
  auto pr = [&](int slot) { printf(" %s=%lld", names[t[3 * slot]], t[3 * slot + 2] - t0); };
  for (int i = 0; i < 28; ++i) {
    printf("tile %2d:", i);
    for (int k = 0; k < 3; ++k) pr(4 * i + k);
    printf(" |");
    for (int k = 0; k < 4; ++k) pr(200 + 4 * i + k);
    printf(" |");
    pr(400 + 4 * i); pr(400 + 4 * i + 1); pr(600 + 2 * i + 1); pr(600 + 2 * i); pr(400 + 4 * i + 2); pr(400 + 4 * i + 3);
    printf("\n");
  }
  return 0;
}
