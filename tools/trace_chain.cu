// Debug tool: timeline of the dgrad chain (build: see tools/trace_chain.sh).  Runs chain_bwd_ts once on M=16384,D=128,C=3072
// and prints CTA 0's event clocks relative to the first event.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
namespace m2 {
int chain_bwd_ts(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
                 int ldw2, const float* dy, float* du, float* dln_w, float* dln_b, float* db2, void* xn_b, void* dy_b,
                 void* g_b, void* dh_b, int ldh, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s);
int chain_trace_read(long long* host, int n);
}
int main() {
  const int M = 16384, D = 128, C = 3072;
  float *u, *lnw, *lnb, *b1, *dy, *du, *dlw, *dlb, *db2; void *w1b, *w2b, *xnb, *dyb;
  cudaMalloc(&u, M * D * 4); cudaMalloc(&dy, M * D * 4); cudaMalloc(&du, M * D * 4);
  cudaMalloc(&lnw, D * 4); cudaMalloc(&lnb, D * 4); cudaMalloc(&b1, C * 4); cudaMalloc(&dlw, D * 4); cudaMalloc(&dlb, D * 4); cudaMalloc(&db2, D * 4);
  cudaMalloc(&w1b, C * D * 2); cudaMalloc(&w2b, C * D * 2); cudaMalloc(&xnb, M * D * 2); cudaMalloc(&dyb, M * D * 2);
  std::vector<float> h(M * D);
  for (auto& v : h) v = (rand() % 2001 - 1000) * 1e-3f;
  cudaMemcpy(u, h.data(), M * D * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, h.data(), M * D * 4, cudaMemcpyHostToDevice);
  std::vector<float> ones(C, 1.f);
  cudaMemcpy(lnw, ones.data(), D * 4, cudaMemcpyHostToDevice); cudaMemset(lnb, 0, D * 4); cudaMemset(b1, 0, C * 4);
  std::vector<__nv_bfloat16> w(C * D);
  for (auto& v : w) v = __float2bfloat16((rand() % 2001 - 1000) * 1e-4f);
  cudaMemcpy(w1b, w.data(), C * D * 2, cudaMemcpyHostToDevice); cudaMemcpy(w2b, w.data(), C * D * 2, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 3; ++rep) {
    int rc = m2::chain_bwd_ts(u, lnw, lnb, w1b, b1, w2b, C, dy, du, dlw, dlb, db2, xnb, dyb, nullptr, nullptr, C, M, D, C, 0.f, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("rc=%d err=%s\n", rc, cudaGetErrorString(e)); return 1; }
  }
  std::vector<long long> t(4096);
  m2::chain_trace_read(t.data(), 4096);
  const char* names[] = {"", "dx wait", "dx go", "hg wait", "hg issue", "epi wait", "epi ready", "epi loaded", "epi arrived", "mma", "dx issued"};
  long long t0 = t[3 * 0 + 2];
  for (int s = 0; s < 1360; ++s) if (t[3 * s + 2] && (t0 == 0 || t[3 * s + 2] < t0)) t0 = t[3 * s + 2];
  for (int j = 0; j < 48; ++j) {
    printf("chunk %2d:", j);
    for (int k = 0; k < 4; ++k) printf(" %s=%lld", names[t[3 * (4 * j + k)]], t[3 * (4 * j + k) + 2] - t0);
    printf(" |");
    for (int k = 0; k < 4; ++k) printf(" %s=%lld", names[t[3 * (400 + 4 * j + k)]], t[3 * (400 + 4 * j + k) + 2] - t0);
    printf(" | hg: top / H go / dG go / issued:");
    for (int k = 0; k < 4; ++k) printf(" %lld", t[3 * (800 + 4 * j + k) + 2] - t0);
    printf("\n");
  }
  return 0;
}
