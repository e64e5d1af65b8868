// Micro-probe: tcgen05.ld / tcgen05.st throughput per lane quadrant and per SM (warps on the same / different quadrants),
// and MUFU.TANH / FFMA2 issue rates.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I m2_mixer_b200/csrc ...
#include <cstdio>
#include "common.cuh"
using namespace m2;

__device__ __forceinline__ void tmem_st32_(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

// mode 0: x32 loads, wait after each; 1: x32 loads, 4 in flight then wait; 2: x16 loads wait each; 3: x32 stores
__global__ void probe(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      tmem_ld32(base + lane_addr + (i & 7) * 32, r);
      tmem_ld_wait();
      acc += r[0] + r[31];
    }
  } else if (mode == 1) {
    for (int i = 0; i < iters; i += 4) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(base + lane_addr + 0, r0);
      tmem_ld32(base + lane_addr + 32, r1);
      tmem_ld32(base + lane_addr + 64, r2);
      tmem_ld32(base + lane_addr + 96, r3);
      tmem_ld_wait();
      acc += r0[0] + r1[31] + r2[5] + r3[7];
    }
  } else if (mode == 2) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[16];
      tmem_ld16(base + lane_addr + (i & 15) * 16, r);
      tmem_ld_wait();
      acc += r[0] + r[15];
    }
  } else {
    uint32_t r[32];
    for (int k = 0; k < 32; ++k) r[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
      tmem_st32_(base + lane_addr + (i & 7) * 32, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678) sink[0] = 1.f;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

// ALU probes: per-warp issue cost of MUFU.TANH and packed FFMA2 with W warps per SM
__global__ void alu_probe(int mode, int iters, long long* out, float* sink) {
  float x = threadIdx.x * 1e-3f, y = 0.5f, z = 0.25f, w = 0.125f;
  float2 a = make_float2(x, y), b = make_float2(z, w), c = make_float2(w, x), d = make_float2(y, z);
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      x = tanh_ap(x); y = tanh_ap(y); z = tanh_ap(z); w = tanh_ap(w);
    }
  } else {
    const float2 k = make_float2(1.0001f, 0.9999f);
    for (int i = 0; i < iters; ++i) {
      a = __ffma2_rn(a, k, b); b = __ffma2_rn(b, k, c); c = __ffma2_rn(c, k, d); d = __ffma2_rn(d, k, a);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (x + y + z + w + a.x + b.y + c.x + d.y == 123.456f) sink[0] = 1.f;
}

int main() {
  long long* dout; float* sink;
  cudaMalloc(&dout, 16); cudaMalloc(&sink, 16);
  const int iters = 1024;
  const char* names[] = {"ld x32 (4 KB/warp), wait each", "ld x32, 4 in flight", "ld x16 (2 KB/warp), wait each", "st x32, wait each"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {1, 2, 4, 8}) {   // warps w and w+4 share a lane quadrant
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) { probe<<<1, warps * 32>>>(mode, iters, dout, sink); cudaDeviceSynchronize(); }
      cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * (mode == 2 ? 2048 : 4096) * warps;
      printf("%-32s warps=%d : %8lld clk, %6.1f clk/op, %6.1f B/clk/SM\n", names[mode], warps, h, double(h) / iters, bytes / h);
    }
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) { alu_probe<<<1, warps * 32>>>(mode, 4096, dout, sink); cudaDeviceSynchronize(); }
      cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
      printf("%-10s warps=%2d : %8lld clk, %6.2f clk per warp-instr per SMSP\n", mode ? "FFMA2" : "MUFU.TANH", warps, h,
             double(h) / (4096.0 * 4) / ((warps + 3) / 4));
    }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
