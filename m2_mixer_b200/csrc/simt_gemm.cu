// fp32 CUDA-core GEMM for the fp32-parity mode (single-pass TF32 tensor-core math misses the 1e-4
// bar by 5x, so parity mode stays on FP32 FMA).  Same contract as the tcgen05 GEMM (kernels.h): either operand
// K-major or MN-major, batched, bias / exact-erf GELU / residual / accumulate epilogue.  It has to be correct,
// not fast: 64x64x16 tiles, 256 threads, 4x4 outputs per thread, bounds-checked loads.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace m2 {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtDev {
  const float* A; const float* B;
  int a_mn, b_mn; long long lda, ldb;
  int M, N, K;
  long long a_batch_rows, b_batch_rows;
  const float* bias; int bias_mode; int act;
  const float* residual; long long ldr, r_batch_stride;
  float* C; long long ldc, c_batch_stride;
  int accumulate;
  Drop drop; long long drop_ld;
  int splitk, batch;     // splitk > 1: blockIdx.z = batch * splitk + slice, partial sums are reduced with atomics into C
};

__global__ void __launch_bounds__(256) simt_gemm_kernel(const SimtDev p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int b = p.splitk > 1 ? blockIdx.z / p.splitk : blockIdx.z;
  const int slice = p.splitk > 1 ? blockIdx.z % p.splitk : 0;
  const int kper = p.splitk > 1 ? ceil_div(ceil_div(p.K, p.splitk), TK) * TK : p.K;
  const int k_lo = slice * kper, k_hi = min(p.K, k_lo + kper);
  const float* A = p.A + b * p.a_batch_rows * p.lda;
  const float* B = p.B + b * p.b_batch_rows * p.ldb;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, each 4 x 4
  float acc[4][4] = {};

  for (int k0 = k_lo; k0 < k_hi; k0 += TK) {
    // A tile -> As[k][m]
    for (int i = tid; i < TM * TK; i += 256) {
      int m, k;
      if (p.a_mn) { m = i % TM; k = i / TM; } else { k = i % TK; m = i / TK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < p.M && gk < k_hi) v = p.a_mn ? A[static_cast<long long>(gk) * p.lda + gm] : A[static_cast<long long>(gm) * p.lda + gk];
      As[k][m] = v;
    }
    for (int i = tid; i < TN * TK; i += 256) {
      int n, k;
      if (p.b_mn) { n = i % TN; k = i / TN; } else { k = i % TK; n = i / TK; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < p.N && gk < k_hi) v = p.b_mn ? B[static_cast<long long>(gk) * p.ldb + gn] : B[static_cast<long long>(gn) * p.ldb + gk];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias_mode == 1) v += p.bias[n];
      else if (p.bias_mode == 2) v += p.bias[m];
      if (p.act == 1) v = gelu_erf(v);
      else if (p.act == 2) v = fmaxf(v, 0.f);
      if (p.drop.thresh) v = drop_apply(p.drop, v, static_cast<unsigned long long>(m) * p.drop_ld + n);
      if (p.residual) v += p.residual[b * p.r_batch_stride + static_cast<long long>(m) * p.ldr + n];
      float* c = p.C + b * p.c_batch_stride + static_cast<long long>(m) * p.ldc + n;
      if (p.splitk > 1) atomicAdd(c, v);           // plain GEMM only (the launcher checks): C was zeroed unless accumulating
      else *c = p.accumulate ? *c + v : v;
    }
  }
}

}  // namespace

int gemm_f32_simt(const GemmArgs& g, cudaStream_t s) {
  if (!g.A || !g.B || !g.C || g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batch <= 0) return M2_ERR_ARG;
  if (g.c_bf16) return M2_ERR_ARG;
  if (g.bias_mode && !g.bias) return M2_ERR_ARG;
  SimtDev d;
  d.A = static_cast<const float*>(g.A); d.B = static_cast<const float*>(g.B);
  d.a_mn = g.a_mn; d.b_mn = g.b_mn; d.lda = g.lda; d.ldb = g.ldb;
  d.M = g.M; d.N = g.N; d.K = g.K;
  d.a_batch_rows = g.a_batch_rows; d.b_batch_rows = g.b_batch_rows;
  d.bias = g.bias; d.bias_mode = g.bias_mode; d.act = g.act;
  d.residual = g.residual; d.ldr = g.ldr; d.r_batch_stride = g.r_batch_stride;
  d.C = static_cast<float*>(g.C); d.ldc = g.ldc; d.c_batch_stride = g.c_batch_stride;
  d.accumulate = g.accumulate;
  d.drop = make_drop(g.drop_p, g.drop_seed, g.drop_site); d.drop_ld = g.drop_ld;
  // Split K when a plain GEMM has a long contraction and too few output tiles to fill the chip: the weight gradients of
  // the small linears (proj / MLP encoder of MIMIC-H: M x N = 64 x 12, K = B * 24 rows) ran as ONE CTA walking K serially
  // (330 us per call at B = 4096).
  const bool plain = !g.bias_mode && !g.act && !g.residual && d.drop.thresh == 0;
  const long long tiles = static_cast<long long>(ceil_div(g.M, TM)) * ceil_div(g.N, TN) * g.batch;
  int split = 1;
  if (plain && g.K >= 1024 && tiles < 148) {
    split = static_cast<int>(std::min<long long>(ceil_div(g.K, 256), (2 * 148 + tiles - 1) / tiles));
    if (g.splitk > 1) split = g.splitk;
  } else if (g.splitk > 1) {
    if (!plain) return M2_ERR_ARG;
    split = g.splitk;
  }
  d.splitk = split; d.batch = g.batch;
  dim3 grid(ceil_div(g.M, TM), ceil_div(g.N, TN), g.batch * split);
  if (grid.y > 65535 || grid.z > 65535) return M2_ERR_ARG;
  LaunchScope scope("simt_gemm", s, (split > 1 && !g.accumulate) ? 2 : 1);
  if (split > 1 && !g.accumulate) {   // the slices add into C
    for (int b = 0; b < g.batch; ++b)
      if (cudaMemset2DAsync(d.C + b * d.c_batch_stride, d.ldc * sizeof(float), 0, static_cast<size_t>(g.N) * sizeof(float), g.M, s) != cudaSuccess)
        return M2_ERR_LAUNCH;
  }
  simt_gemm_kernel<<<grid, 256, 0, s>>>(d);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace m2
