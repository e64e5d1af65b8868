// Micro-probe: issue rates of the GELU building blocks on one SM with every scheduler busy.
//   tanh.approx.f32 | tanh.approx.bf16x2 (two results per instruction?) | ex2.approx.ftz.f32 | fma.rn.f32x2 | fma.rn.bf16x2
// Every thread runs 8 independent dependency chains so latency does not bound the rate; 1024 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mufu_probe.cu -o tools/mufu_probe.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kMode>
__global__ void __launch_bounds__(1024) probe(int iters, long long* clk, float* sink) {
  float f[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = 0.001f * (threadIdx.x + i); h[i] = 0x3c003c00u + threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kMode == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
      if (kMode == 1) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
      if (kMode == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (kMode == 3) {
        unsigned long long v;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(f[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
        asm volatile("{.reg .f32 t; mov.b64 {%0, t}, %1;}" : "=f"(f[i]) : "l"(v));
      }
      if (kMode == 4) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(h[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(h[i]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int kMode>
void run(const char* name, int results_per_instr) {
  long long* clk; float* sink;
  cudaMalloc(&clk, 8 * sizeof(long long)); cudaMalloc(&sink, 1024 * sizeof(float));
  const int iters = 4096;
  probe<kMode><<<1, 1024>>>(64, clk, sink);
  probe<kMode><<<1, 1024>>>(iters, clk, sink);
  long long c; cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
  const double instr = 1024.0 / 32 * 8 * iters;   // warp instructions on the SM
  printf("%-24s %8.2f clk per warp instruction per SM, %7.1f results/clk/SM\n", name, c / instr, 32.0 * results_per_instr * instr / c);
  cudaFree(clk); cudaFree(sink);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("tanh.approx.bf16x2", 2);
  run<2>("ex2.approx.ftz.f32", 1);
  run<3>("fma.rn.f32x2 (+2 movs)", 2);
  run<4>("fma.rn.bf16x2", 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
