// Micro-probe: cost of waiting on an ALREADY COMPLETED mbarrier phase from one warp while other warps are idle / busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I m2_mixer_b200/csrc tools/wait_probe.cu -o tools/wait_probe.bin
#include <cstdio>
#include "common.cuh"
using namespace m2;

__device__ __forceinline__ void wait_asm(uint64_t* bar, uint32_t parity) {   // CUTLASS-style: try_wait loop in PTX
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int MODE>
__global__ void probe(int busy_warps, long long* out, float* sink) {
  __shared__ uint64_t bars[8];
  __shared__ float sb[1024];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sb[i] = i * 1e-3f;
  __syncthreads();
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_arrive(&bars[i]);   // phase 0 complete on every barrier
  __syncthreads();
  if (warp == blockDim.x / 32 - 1) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      if (MODE == 0) mbar_wait(&bars[i & 7], 0);
      else if (MODE == 1) wait_asm(&bars[i & 7], 0);
      else while (!test_wait(&bars[i & 7], 0)) {}
    }
    long long t1 = clock64();
    if (threadIdx.x % 32 == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / 64;
  } else if (warp < busy_warps) {
    float a = threadIdx.x * 1e-3f, b = 0.5f, c = 0.25f;
#pragma unroll 1
    for (int i = 0; i < 4000; ++i) {
      a = fmaf(a, b, c); b = fmaf(b, a, c);
      float t;
      asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(a));
      c = t + sb[(threadIdx.x + i) & 1023];
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c;
  }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 148 * 1024 * 4);
  const char* names[] = {"mbar_wait (C++ spin loop, try_wait)", "PTX loop try_wait", "test_wait loop"};
  for (int busy : {0, 4, 8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      const int threads = 32 * (busy + 1);
      if (mode == 0) probe<0><<<148, threads>>>(busy, out, sink);
      if (mode == 1) probe<1><<<148, threads>>>(busy, out, sink);
      if (mode == 2) probe<2><<<148, threads>>>(busy, out, sink);
      long long h = 0;
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("busy warps %2d  %-40s : %lld clk per completed wait (%s)\n", busy, names[mode], h, cudaGetErrorString(e));
    }
  }
  return 0;
}
