// Token-mixing half of a Mixer block, transpose-free, on the CUDA cores.
//
// Reference arithmetic: MixerBlock.token_mix, modules/mixer.py:30-35,43  (LN -> 'b n d -> b d n' -> FeedForward(N->T->N)
// -> back-permute -> residual).  Restated without the permutes:
//     H[b] = GELU(Wt1[T,N] . LN(x[b])[N,D] + bt1 1^T)      u[b] = x[b] + Wt2[N,T] . H[b] + bt2 1^T
// For every shipped config the contraction lengths are N = 4..25 and T = 8..32 (0.3 % of the model FLOPs):
// a tensor-core tile would be >90 % padding, so this path keeps the [N x D] sample tile in shared memory and runs
// FP32 FMAs, reading x once and writing u once.  One CTA handles a slice of `dt` hidden columns of one sample and
// loops over samples; weights are read through L1 (warp-uniform addresses).
#include "common.cuh"
#include "kernels.h"

namespace m2 {
namespace {

constexpr int kThreads = 256;

// Row statistics of sample b (full D) -> stat[n] = (mean, rstd).  One warp per row.
__device__ __forceinline__ void row_stats(const float* xb, int N, int D, float* stat) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < N; n += kThreads / 32) {
    const float* xr = xb + static_cast<long long>(n) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = xr[c] - mean; ss += d * d; }
    const float rstd = rsqrtf(warp_sum(ss) / D + kLnEps);
    if (lane == 0) { stat[2 * n] = mean; stat[2 * n + 1] = rstd; }
  }
}

template <int DT>
__global__ void __launch_bounds__(kThreads) token_mix_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
    const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
    const float* __restrict__ b2, float* __restrict__ u, int B, int N, int D, int T, int exact_gelu) {
  extern __shared__ float sm[];
  constexpr int LD = DT + 1;
  float* sXn = sm;                       // [N][LD]
  float* sG = sXn + N * LD;              // [T][LD]
  float* stat = sG + T * LD;             // [N][2] (mean, rstd)
  const int dd = threadIdx.x % DT, ty = threadIdx.x / DT;
  constexpr int TY = kThreads / DT;
  const int d = blockIdx.y * DT + dd;
  const bool dok = d < D;
  const float gam = dok ? ln_w[d] : 0.f, bet = dok ? ln_b[d] : 0.f;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* xb = x + static_cast<long long>(b) * N * D;
    __syncthreads();   // previous sample's smem fully consumed
    row_stats(xb, N, D, stat);
    __syncthreads();
    for (int n = ty; n < N; n += TY) {
      sXn[n * LD + dd] = dok ? (xb[static_cast<long long>(n) * D + d] - stat[2 * n]) * stat[2 * n + 1] * gam + bet : 0.f;
    }
    __syncthreads();
    for (int t = ty; t < T; t += TY) {
      float acc = b1[t];
      const float* wr = w1 + static_cast<long long>(t) * N;
      for (int n = 0; n < N; ++n) acc = fmaf(wr[n], sXn[n * LD + dd], acc);
      sG[t * LD + dd] = exact_gelu ? gelu_erf(acc) : gelu_fast(acc);
    }
    __syncthreads();
    if (dok) {
      for (int n = ty; n < N; n += TY) {
        float acc = b2[n] + xb[static_cast<long long>(n) * D + d];
        const float* wr = w2 + static_cast<long long>(n) * T;
        for (int t = 0; t < T; ++t) acc = fmaf(wr[t], sG[t * LD + dd], acc);
        u[(static_cast<long long>(b) * N + n) * D + d] = acc;
      }
    }
  }
}

// Backward: given du = dL/du, recompute LN / H / GELU and produce
//   dxn = dL/dLN(x)  (fp32 [B,N,D]; LayerNorm backward + the residual term are done by ln_bwd with dres = du)
//   dw1[T,N] += dH . Xn^T,  db1[T] += rowsum(dH),  dw2[N,T] += dU . G^T,  db2[N] += rowsum(dU)   (summed over b and d)
template <int DT>
__global__ void __launch_bounds__(kThreads) token_mix_bwd_kernel(
    const float* __restrict__ du, const float* __restrict__ x, const float* __restrict__ ln_w,
    const float* __restrict__ ln_b, const float* __restrict__ w1, const float* __restrict__ b1,
    const float* __restrict__ w2, float* __restrict__ dxn, float* __restrict__ dw1, float* __restrict__ db1,
    float* __restrict__ dw2, float* __restrict__ db2, int B, int N, int D, int T, int exact_gelu, int reg_acc) {
  extern __shared__ float sm[];
  constexpr int LD = DT + 1;
  float* sXn = sm;                    // [N][LD]
  float* sDU = sXn + N * LD;          // [N][LD]
  float* sG = sDU + N * LD;           // [T][LD]
  float* sDH = sG + T * LD;           // [T][LD]
  float* stat = sDH + T * LD;
  const int dd = threadIdx.x % DT, ty = threadIdx.x / DT;
  constexpr int TY = kThreads / DT;
  const int d = blockIdx.y * DT + dd;
  const bool dok = d < D;
  const float gam = dok ? ln_w[d] : 0.f, bet = dok ? ln_b[d] : 0.f;
  const int NT = N * T;
  // register accumulation of the weight gradients across the samples this CTA visits (small N*T only)
  constexpr int kAcc = 4;
  float a1[kAcc] = {0.f, 0.f, 0.f, 0.f}, a2[kAcc] = {0.f, 0.f, 0.f, 0.f}, ab = 0.f;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* xb = x + static_cast<long long>(b) * N * D;
    const float* dub = du + static_cast<long long>(b) * N * D;
    __syncthreads();
    row_stats(xb, N, D, stat);
    __syncthreads();
    for (int n = ty; n < N; n += TY) {
      sXn[n * LD + dd] = dok ? (xb[static_cast<long long>(n) * D + d] - stat[2 * n]) * stat[2 * n + 1] * gam + bet : 0.f;
      sDU[n * LD + dd] = dok ? dub[static_cast<long long>(n) * D + d] : 0.f;
    }
    __syncthreads();
    for (int t = ty; t < T; t += TY) {
      float h = b1[t], dg = 0.f;
      const float* wr = w1 + static_cast<long long>(t) * N;
      for (int n = 0; n < N; ++n) {
        h = fmaf(wr[n], sXn[n * LD + dd], h);
        dg = fmaf(w2[static_cast<long long>(n) * T + t], sDU[n * LD + dd], dg);
      }
      float g, dgelu;
      if (exact_gelu) { g = gelu_erf(h); dgelu = gelu_erf_grad(h); } else { g = gelu_fast_grad(h, dgelu); }
      sG[t * LD + dd] = dok ? g : 0.f;
      sDH[t * LD + dd] = dok ? dg * dgelu : 0.f;
    }
    __syncthreads();
    if (dok) {
      for (int n = ty; n < N; n += TY) {
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc = fmaf(w1[static_cast<long long>(t) * N + n], sDH[t * LD + dd], acc);
        dxn[(static_cast<long long>(b) * N + n) * D + d] = acc;
      }
    }
    // weight gradients: pair p = t*N + n  ->  dw1[t][n] and dw2[n][t]
    if (reg_acc) {
#pragma unroll
      for (int i = 0; i < kAcc; ++i) {
        const int p = threadIdx.x + i * kThreads;
        if (p < NT) {
          const int t = p / N, n = p - t * N;
          float s1 = 0.f, s2 = 0.f;
          for (int c = 0; c < DT; ++c) {
            s1 = fmaf(sDH[t * LD + c], sXn[n * LD + c], s1);
            s2 = fmaf(sDU[n * LD + c], sG[t * LD + c], s2);
          }
          a1[i] += s1; a2[i] += s2;
        }
      }
      // biases: threads [0,T) -> db1[t], threads [T, T+N) -> db2[n]
      if (threadIdx.x < T + N) {
        const float* rowp = threadIdx.x < T ? sDH + threadIdx.x * LD : sDU + (threadIdx.x - T) * LD;
        float s = 0.f;
        for (int c = 0; c < DT; ++c) s += rowp[c];
        ab += s;
      }
    } else {
      for (int p = threadIdx.x; p < NT; p += kThreads) {
        const int t = p / N, n = p - t * N;
        float s1 = 0.f, s2 = 0.f;
        for (int c = 0; c < DT; ++c) {
          s1 = fmaf(sDH[t * LD + c], sXn[n * LD + c], s1);
          s2 = fmaf(sDU[n * LD + c], sG[t * LD + c], s2);
        }
        atomicAdd(&dw1[static_cast<long long>(t) * N + n], s1);
        atomicAdd(&dw2[static_cast<long long>(n) * T + t], s2);
      }
      for (int i = threadIdx.x; i < T + N; i += kThreads) {
        const float* rowp = i < T ? sDH + i * LD : sDU + (i - T) * LD;
        float s = 0.f;
        for (int c = 0; c < DT; ++c) s += rowp[c];
        atomicAdd(i < T ? &db1[i] : &db2[i - T], s);
      }
    }
  }
  if (reg_acc) {
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
      const int p = threadIdx.x + i * kThreads;
      if (p < NT) {
        const int t = p / N, n = p - t * N;
        atomicAdd(&dw1[static_cast<long long>(t) * N + n], a1[i]);
        atomicAdd(&dw2[static_cast<long long>(n) * T + t], a2[i]);
      }
    }
    if (threadIdx.x < T + N) atomicAdd(threadIdx.x < T ? &db1[threadIdx.x] : &db2[threadIdx.x - T], ab);
  }
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) != cudaSuccess)
      return M2_ERR_LAUNCH;
  }
  return M2_OK;
}

constexpr size_t kMaxSmem = 220 * 1024;

}  // namespace

int token_mix_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* u, int B, int N, int D, int T, int exact_gelu, cudaStream_t s) {
  if (B <= 0 || N <= 0 || D <= 0 || T <= 0) return M2_ERR_ARG;
  auto bytes = [&](int dt) { return static_cast<size_t>(N + T) * (dt + 1) * 4 + static_cast<size_t>(N) * 8; };
  int dt = 32;
  while (dt > 8 && bytes(dt) > kMaxSmem) dt >>= 1;
  if (bytes(dt) > kMaxSmem) return M2_ERR_ARG;
  const int ds = ceil_div(D, dt);
  int gx = B < ceil_div(148 * 8, ds) ? B : ceil_div(148 * 8, ds);
  dim3 grid(gx, ds);
  int rc;
#define M2_TM_FWD(DT_)                                                                                         \
  rc = set_smem(token_mix_fwd_kernel<DT_>, bytes(DT_));                                                        \
  if (rc) return rc;                                                                                           \
  token_mix_fwd_kernel<DT_><<<grid, kThreads, bytes(DT_), s>>>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, D, T, exact_gelu);
  if (dt == 32) { M2_TM_FWD(32) } else if (dt == 16) { M2_TM_FWD(16) } else { M2_TM_FWD(8) }
#undef M2_TM_FWD
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int token_mix_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                  const float* w2, float* dxn, float* dw1, float* db1, float* dw2, float* db2, int B, int N, int D, int T,
                  int exact_gelu, cudaStream_t s) {
  if (B <= 0 || N <= 0 || D <= 0 || T <= 0) return M2_ERR_ARG;
  auto bytes = [&](int dt) { return static_cast<size_t>(2 * N + 2 * T) * (dt + 1) * 4 + static_cast<size_t>(N) * 8; };
  int dt = 32;
  while (dt > 8 && bytes(dt) > kMaxSmem) dt >>= 1;
  if (bytes(dt) > kMaxSmem) return M2_ERR_ARG;
  const int ds = ceil_div(D, dt);
  const int reg_acc = (N * T <= 4 * kThreads && N + T <= kThreads) ? 1 : 0;
  const int cap = reg_acc ? ceil_div(148 * 4, ds) : ceil_div(148 * 8, ds);
  int gx = B < cap ? B : cap;
  dim3 grid(gx, ds);
  int rc;
#define M2_TM_BWD(DT_)                                                                                            \
  rc = set_smem(token_mix_bwd_kernel<DT_>, bytes(DT_));                                                           \
  if (rc) return rc;                                                                                              \
  token_mix_bwd_kernel<DT_><<<grid, kThreads, bytes(DT_), s>>>(du, x, ln_w, ln_b, w1, b1, w2, dxn, dw1, db1, dw2, db2, B, N, \
                                                               D, T, exact_gelu, reg_acc);
  if (dt == 32) { M2_TM_BWD(32) } else if (dt == 16) { M2_TM_BWD(16) } else { M2_TM_BWD(8) }
#undef M2_TM_BWD
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace m2
