// Debug tool: in-kernel timeline of the forward channel-mix chain.  Build (tools/trace_fwd.sh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DM2_TRACE -Im2_mixer_b200/csrc
//        tools/trace_fwd.cu m2_mixer_b200/csrc/chain_ts.cu m2_mixer_b200/csrc/profile.cu -o tools/trace_fwd.bin
// Runs chain_fwd_ts on M=16384, D=128, C=3072 and prints CTA 0's event clocks (MMA issuer | first warp of the epilogue
// group that owns the chunk) relative to the first event.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
namespace m2 {
int chain_fwd_ts(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
                 int ldw2, const float* b2, float* y, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s);
int chain_trace_read(long long* host, int n);
}
int main(int argc, char** argv) {
  const int M = 16384, D = 128, C = 3072;
  const float p = argc > 1 ? atof(argv[1]) : 0.f;
  float *u, *lnw, *lnb, *b1, *b2, *y; void *w1b, *w2b;
  cudaMalloc(&u, M * D * 4); cudaMalloc(&y, M * D * 4);
  cudaMalloc(&lnw, D * 4); cudaMalloc(&lnb, D * 4); cudaMalloc(&b1, C * 4); cudaMalloc(&b2, D * 4);
  cudaMalloc(&w1b, C * D * 2); cudaMalloc(&w2b, C * D * 2);
  std::vector<float> h(M * D);
  for (auto& v : h) v = (rand() % 2001 - 1000) * 1e-3f;
  cudaMemcpy(u, h.data(), M * D * 4, cudaMemcpyHostToDevice);
  std::vector<float> ones(C, 1.f);
  cudaMemcpy(lnw, ones.data(), D * 4, cudaMemcpyHostToDevice); cudaMemset(lnb, 0, D * 4); cudaMemset(b1, 0, C * 4); cudaMemset(b2, 0, D * 4);
  std::vector<__nv_bfloat16> w(C * D);
  for (auto& v : w) v = __float2bfloat16((rand() % 2001 - 1000) * 1e-4f);
  cudaMemcpy(w1b, w.data(), C * D * 2, cudaMemcpyHostToDevice); cudaMemcpy(w2b, w.data(), C * D * 2, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 3; ++rep) {
    int rc = m2::chain_fwd_ts(u, lnw, lnb, w1b, b1, w2b, C, b2, y, M, D, C, p, 1234, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e != cudaSuccess) { printf("rc=%d err=%s\n", rc, cudaGetErrorString(e)); return 1; }
  }
  std::vector<long long> t(4096);
  m2::chain_trace_read(t.data(), 4096);
  const char* names[] = {"", "iss:wait", "iss:G", "iss:W", "iss:done", "epi:start", "epi:waitH", "epi:H", "epi:arrived"};
  long long t0 = 0;
  for (int s = 0; s < 1360; ++s) if (t[3 * s + 2] && (t0 == 0 || t[3 * s + 2] < t0)) t0 = t[3 * s + 2];
  const char* ph[] = {"entry", "init done", "LN staged", "roles start", "epilogue loop done", "Y complete", "Y staged", "output written", "LN start (warp 0)", "LN done (warp 0)", "bias staged (warp 0)"};
  for (int k = 0; k < 11; ++k) printf("phase %-20s %lld\n", ph[k], t[3 * (1300 + k) + 2] - t0);
  for (int j = 0; j < 48; ++j) {
    printf("chunk %2d:", j);
    for (int k = 0; k < 4; ++k) printf(" %s=%lld", names[t[3 * (4 * j + k)]], t[3 * (4 * j + k) + 2] - t0);
    printf(" |");
    for (int k = 0; k < 4; ++k) printf(" %s=%lld", names[t[3 * (400 + 4 * j + k)]], t[3 * (400 + 4 * j + k) + 2] - t0);
    printf("\n");
  }
  return 0;
}
