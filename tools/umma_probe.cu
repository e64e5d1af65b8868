// Micro-probe: cost of back-to-back tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, of the operand source
// (A from shared memory vs tensor memory) and of the accumulator pattern (one dependent chain vs several rotating
// accumulators).  One CTA per SM, one issuing thread; cycles from first issue to the commit's mbarrier completion.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I m2_mixer_b200/csrc tools/umma_probe.cu -o /tmp/umma_probe
#include <cstdio>
#include <vector>

#include "common.cuh"

using namespace m2;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// MODE: 0 = SS, 1 = TS.  NACC accumulators used round-robin, CHAIN consecutive MMAs accumulate into the same one.
// The pattern is compile-time so that the issue loop contains nothing but descriptor adds and the MMAs.
template <int N, int MODE, int NACC, int CHAIN, int B_MN>
__global__ void probe(int rounds, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // [128][64] bf16 K-major SW128 (16 KB)
  uint8_t* sB = smem + 16384;         // [256][64] bf16 (32 KB)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, B_MN);
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    const uint32_t tA = base + 448;   // 64 columns reserved for a TMEM A operand
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
      for (int acc = 0; acc < NACC; ++acc) {
#pragma unroll
        for (int c = 0; c < CHAIN; ++c) {
          const int kk = c & 3;
          const uint32_t d = base + acc * N;
          const uint64_t bd = B_MN ? umma_desc_sw128(b_addr + kk * 2048, 8192, 1024) : umma_desc_sw128(b_addr + kk * 32, 16, 1024);
          if (MODE == 0) umma_bf16(d, umma_desc_sw128(a_addr + kk * 32, 16, 1024), bd, idesc, c ? 1u : 0u);
          else umma_ts(d, tA + kk * 8, bd, idesc, c ? 1u : 0u);
        }
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    long long tc = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = tc - t1; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(base, 512);
}

template <int N, int MODE, int NACC, int CHAIN, int B_MN>
void run(int grid, long long* dout) {
  auto k = probe<N, MODE, NACC, CHAIN, B_MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int rounds = 512 / (NACC * CHAIN) > 0 ? 512 / (NACC * CHAIN) : 1;
  const int nmma = rounds * NACC * CHAIN;
  long long h[3];
  for (int rep = 0; rep < 2; ++rep) {
    k<<<grid, 128, 64 * 1024>>>(rounds, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d %s B=%s grid=%3d nmma=%4d nacc=%d chain=%3d : issue %7lld clk, commit %5lld clk, total %7lld clk, %6.1f clk/MMA (floor %d)\n", N,
         MODE ? "TS" : "SS", B_MN ? "MN" : "K ", grid, nmma, NACC, CHAIN, h[0], h[2], h[1], double(h[1]) / nmma, N / 2);
}

template <int N, int MODE, int NACC, int CHAIN, int B_MN>
void run_short(int grid, long long* dout) {
  auto k = probe<N, MODE, NACC, CHAIN, B_MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  long long h[3];
  for (int rep = 0; rep < 2; ++rep) { k<<<grid, 128, 64 * 1024>>>(1, dout); cudaDeviceSynchronize(); }
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  printf("SHORT N=%3d %s nmma=%3d : issue %6lld clk, commit %5lld clk, total %6lld clk\n", N, MODE ? "TS" : "SS", NACC * CHAIN, h[0], h[2], h[1]);
}

template <int MODE>
void sweep(int grid, long long* dout) {
  run_short<128, MODE, 1, 4, 1>(grid, dout);
  run_short<64, MODE, 2, 8, 0>(grid, dout);
  run_short<64, MODE, 1, 1, 0>(grid, dout);
  run<64, MODE, 1, 64, 0>(grid, dout);
  run<64, MODE, 1, 8, 0>(grid, dout);
  run<64, MODE, 2, 8, 0>(grid, dout);
  run<64, MODE, 4, 8, 0>(grid, dout);
  run<64, MODE, 4, 1, 0>(grid, dout);
  run<64, MODE, 2, 8, 1>(grid, dout);
  run<128, MODE, 1, 64, 0>(grid, dout);
  run<128, MODE, 1, 4, 0>(grid, dout);
  run<128, MODE, 2, 4, 0>(grid, dout);
  run<128, MODE, 2, 4, 1>(grid, dout);
  run<256, MODE, 1, 64, 0>(grid, dout);
  run<256, MODE, 1, 4, 0>(grid, dout);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 32);
  for (int grid : {148}) {
    sweep<0>(grid, dout);
    sweep<1>(grid, dout);
  }
  return 0;
}
