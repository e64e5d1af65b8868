// Token-mixing half of a Mixer block, transpose-free, on the CUDA cores.
//
// Reference arithmetic: MixerBlock.token_mix, modules/mixer.py:30-35,43  (LN -> 'b n d -> b d n' -> FeedForward(N->T->N)
// -> back-permute -> residual).  Restated without the permutes:
//     H[b] = GELU(Wt1[T,N] . LN(x[b])[N,D] + bt1 1^T)      u[b] = x[b] + Wt2[N,T] . H[b] + bt2 1^T
// For every shipped config the contraction lengths are N = 4..25 and T = 8..32 (0.3 % of the model FLOPs):
// a tensor-core tile would be >90 % padding, so this path keeps the [N x D] sample tile in shared memory and runs
// FP32 FMAs, reading x once and writing u once.  One CTA handles a slice of `dt` hidden columns of one sample and
// loops over samples; weights are read through L1 (warp-uniform addresses).
#include <cstdlib>

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
    const float* __restrict__ b2, float* __restrict__ u, int B, int N, int D, int T, int exact_gelu, const Drop dh,
    const Drop dout) {
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
      float gv = exact_gelu ? gelu_erf(acc) : gelu_fast(acc);
      if (dh.thresh) gv = drop_apply(dh, gv, (static_cast<unsigned long long>(b) * T + t) * D + d);
      sG[t * LD + dd] = gv;
    }
    __syncthreads();
    if (dok) {
      for (int n = ty; n < N; n += TY) {
        float acc = b2[n];
        const float* wr = w2 + static_cast<long long>(n) * T;
        for (int t = 0; t < T; ++t) acc = fmaf(wr[t], sG[t * LD + dd], acc);
        const long long o = (static_cast<long long>(b) * N + n) * D + d;
        if (dout.thresh) acc = drop_apply(dout, acc, o);
        u[o] = xb[static_cast<long long>(n) * D + d] + acc;
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
    float* __restrict__ dw2, float* __restrict__ db2, int B, int N, int D, int T, int exact_gelu, int reg_acc,
    const Drop dh, const Drop dout) {
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
      float dv = dok ? dub[static_cast<long long>(n) * D + d] : 0.f;
      if (dok && dout.thresh) dv = drop_apply(dout, dv, (static_cast<unsigned long long>(b) * N + n) * D + d);
      sDU[n * LD + dd] = dv;   // gradient of the (dropped) branch output; the residual path is added by ln_bwd
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
      if (dok && dh.thresh) {   // G' = m*s*G ; dH = dG' * m*s * gelu'(h)
        const bool keep = drop_keep(dh, (static_cast<unsigned long long>(b) * T + t) * D + d);
        g = keep ? g * dh.scale : 0.f;
        dgelu = keep ? dgelu * dh.scale : 0.f;
      }
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

// ---------------------------------------------------------------------------------------------------------------
// Small-N specialisation (every shipped config: N <= 32, D in {32,64,128,256}): ONE THREAD PER (sample, hidden column).
// The thread keeps its column of the [N x D] sample tile in registers, so LayerNorm -> Wt1 -> GELU -> Wt2 -> residual
// needs no shared-memory tile and only two block reductions (the LN row statistics, reduced across the D threads of
// a sample with warp shuffles).  x is read once and u written once, fully coalesced along d.
template <int NMAX>
__device__ __forceinline__ void small_ln(const float (&x)[NMAX], float (&xn)[NMAX], int N, int D, float gam, float bet,
                                         float* red /*[8][NMAX]*/, int warp, int lane, int first_warp, int nwarps_sample) {
#pragma unroll
  for (int n = 0; n < NMAX; ++n)
    if (n < N) {
      const float v = warp_sum(x[n]);
      if (lane == 0) red[warp * NMAX + n] = v;
    }
  __syncthreads();
  float mean[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n)
    if (n < N) {
      float t = 0.f;
      for (int w = 0; w < nwarps_sample; ++w) t += red[(first_warp + w) * NMAX + n];
      mean[n] = t / D;
    }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NMAX; ++n)
    if (n < N) {
      const float c = x[n] - mean[n];
      const float v = warp_sum(c * c);
      if (lane == 0) red[warp * NMAX + n] = v;
    }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NMAX; ++n)
    if (n < N) {
      float t = 0.f;
      for (int w = 0; w < nwarps_sample; ++w) t += red[(first_warp + w) * NMAX + n];
      xn[n] = (x[n] - mean[n]) * rsqrtf(t / D + kLnEps) * gam + bet;
    }
  __syncthreads();   // red is reused by the next slab
}

// Weights are staged in shared memory padded to NMAX columns per row (zeros beyond N) so that the inner loops are
// fully unrolled without predicates; two hidden units (t, t+1) are processed together with packed fp32x2 math.
template <int NMAX>
__device__ __forceinline__ void stage_weights(const float* w1, const float* b1, const float* w2, int N, int T, int Tp,
                                              float* sW1, float* sW2t, float* sB1) {
  for (int i = threadIdx.x; i < Tp * NMAX; i += kThreads) {
    const int t = i / NMAX, n = i - t * NMAX;
    const bool in = t < T && n < N;
    sW1[i] = in ? w1[t * N + n] : 0.f;
    sW2t[i] = in ? w2[n * T + t] : 0.f;
  }
  for (int i = threadIdx.x; i < Tp; i += kThreads) sB1[i] = i < T ? b1[i] : 0.f;
}

template <int NMAX>
__global__ void __launch_bounds__(kThreads) token_mix_small_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
    const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
    const float* __restrict__ b2, float* __restrict__ u, int B, int N, int D, int T, int exact_gelu, const Drop dh,
    const Drop dout) {
  extern __shared__ float sm[];
  const int Tp = (T + 1) & ~1;
  float* sW1 = sm;                    // [Tp][NMAX]
  float* sW2t = sW1 + Tp * NMAX;      // [Tp][NMAX]  (transposed: both inner loops run over n with t fixed)
  float* sB1 = sW2t + Tp * NMAX;      // [Tp]
  float* sB2 = sB1 + Tp;              // [NMAX]
  float* red = sB2 + NMAX;            // [8][NMAX]
  stage_weights<NMAX>(w1, b1, w2, N, T, Tp, sW1, sW2t, sB1);
  for (int i = threadIdx.x; i < NMAX; i += kThreads) sB2[i] = i < N ? b2[i] : 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int spb = kThreads / D;                       // samples per CTA slab
  const int s_in = threadIdx.x / D, d = threadIdx.x - s_in * D;
  const int wps = D / 32, first_warp = s_in * wps;
  const float gam = ln_w[d], bet = ln_b[d];
  __syncthreads();
  for (int b0 = blockIdx.x * spb; b0 < B; b0 += gridDim.x * spb) {
    const int b = b0 + s_in;
    const bool ok = b < B;
    const float* xb = x + (static_cast<long long>(ok ? b : 0) * N) * D + d;
    float xv[NMAX], xn[NMAX];
    float2 acc[NMAX];                                 // (.x, .y) = partial sums of the even / odd hidden units
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      xv[n] = (n < N && ok) ? xb[static_cast<long long>(n) * D] : 0.f;
      xn[n] = 0.f;
      acc[n] = make_float2(0.f, 0.f);
    }
    small_ln<NMAX>(xv, xn, N, D, gam, bet, red, warp, lane, first_warp, wps);
#pragma unroll 2
    for (int t = 0; t < Tp; t += 2) {
      const float* w1a = sW1 + t * NMAX;
      float2 h = make_float2(sB1[t], sB1[t + 1]);
#pragma unroll
      for (int n = 0; n < NMAX; ++n) h = __ffma2_rn(make_float2(w1a[n], w1a[NMAX + n]), make_float2(xn[n], xn[n]), h);
      float2 g;
      if (exact_gelu) g = make_float2(gelu_erf(h.x), gelu_erf(h.y));
      else g = gelu2(h);
      if (dh.thresh) {
        g.x = drop_apply(dh, g.x, (static_cast<unsigned long long>(b) * T + t) * D + d);
        g.y = drop_apply(dh, g.y, (static_cast<unsigned long long>(b) * T + t + 1) * D + d);
      }
      const float* w2a = sW2t + t * NMAX;
#pragma unroll
      for (int n = 0; n < NMAX; ++n) acc[n] = __ffma2_rn(make_float2(w2a[n], w2a[NMAX + n]), g, acc[n]);
    }
    if (ok) {
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) {
          const long long o = (static_cast<long long>(b) * N + n) * D + d;
          float v = acc[n].x + acc[n].y + sB2[n];
          if (dout.thresh) v = drop_apply(dout, v, o);
          u[o] = xv[n] + v;
        }
    }
  }
}

constexpr int kSmallLd = 260;   // row stride of the phase-2 tiles (float4-aligned, spreads rows over banks)

template <int NMAX>
__global__ void __launch_bounds__(kThreads) token_mix_small_bwd_kernel(
    const float* __restrict__ du, const float* __restrict__ x, const float* __restrict__ ln_w,
    const float* __restrict__ ln_b, const float* __restrict__ w1, const float* __restrict__ b1,
    const float* __restrict__ w2, float* __restrict__ dxn, float* __restrict__ dw1, float* __restrict__ db1,
    float* __restrict__ dw2, float* __restrict__ db2, int B, int N, int D, int T, int exact_gelu, const Drop dh,
    const Drop dout) {
  extern __shared__ float sm[];
  const int Tp = (T + 1) & ~1;
  float* sW1 = sm;                        // [Tp][NMAX]
  float* sW2t = sW1 + Tp * NMAX;          // [Tp][NMAX]
  float* sB1 = sW2t + Tp * NMAX;          // [Tp]
  float* red = sB1 + Tp;                  // [8][NMAX]
  float* tile = red + 8 * NMAX;
  tile += (4 - ((tile - sm) & 3)) & 3;    // 16-byte align the float4 tiles
  float* sDH = tile;                      // [Tp][kSmallLd]
  float* sG = sDH + Tp * kSmallLd;        // [Tp][kSmallLd]
  float* sXn = sG + Tp * kSmallLd;        // [N][kSmallLd]
  float* sDU = sXn + N * kSmallLd;        // [N][kSmallLd]
  stage_weights<NMAX>(w1, b1, w2, N, T, Tp, sW1, sW2t, sB1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int spb = kThreads / D;
  const int s_in = threadIdx.x / D, d = threadIdx.x - s_in * D;
  const int wps = D / 32, first_warp = s_in * wps;
  const float gam = ln_w[d], bet = ln_b[d];
  const int NT = N * T, nred = NT + T + N;       // reduction outputs: dw1/dw2 pairs, db1, db2
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  for (int b0 = blockIdx.x * spb; b0 < B; b0 += gridDim.x * spb) {
    const int b = b0 + s_in;
    const bool ok = b < B;
    const long long base = (static_cast<long long>(ok ? b : 0) * N) * D + d;
    float xv[NMAX], xn[NMAX], dub[NMAX];
    float2 dx[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      const bool in = n < N && ok;
      xv[n] = in ? x[base + static_cast<long long>(n) * D] : 0.f;
      float g = in ? du[base + static_cast<long long>(n) * D] : 0.f;
      if (in && dout.thresh) g = drop_apply(dout, g, static_cast<unsigned long long>(base) + static_cast<unsigned long long>(n) * D);
      dub[n] = g;
      xn[n] = 0.f;
      dx[n] = make_float2(0.f, 0.f);
    }
    small_ln<NMAX>(xv, xn, N, D, gam, bet, red, warp, lane, first_warp, wps);
#pragma unroll 2
    for (int t = 0; t < Tp; t += 2) {
      const float* w1a = sW1 + t * NMAX;
      const float* w2a = sW2t + t * NMAX;
      float2 h = make_float2(sB1[t], sB1[t + 1]), dg = make_float2(0.f, 0.f);
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        h = __ffma2_rn(make_float2(w1a[n], w1a[NMAX + n]), make_float2(xn[n], xn[n]), h);
        dg = __ffma2_rn(make_float2(w2a[n], w2a[NMAX + n]), make_float2(dub[n], dub[n]), dg);
      }
      float2 g, dgelu;
      if (exact_gelu) {
        g = make_float2(gelu_erf(h.x), gelu_erf(h.y));
        dgelu = make_float2(gelu_erf_grad(h.x), gelu_erf_grad(h.y));
      } else {
        g = gelu2_grad(h, dgelu);
      }
      if (dh.thresh) {
        const bool k0 = drop_keep(dh, (static_cast<unsigned long long>(b) * T + t) * D + d);
        const bool k1 = drop_keep(dh, (static_cast<unsigned long long>(b) * T + t + 1) * D + d);
        g.x = k0 ? g.x * dh.scale : 0.f; dgelu.x = k0 ? dgelu.x * dh.scale : 0.f;
        g.y = k1 ? g.y * dh.scale : 0.f; dgelu.y = k1 ? dgelu.y * dh.scale : 0.f;
      }
      float2 dhv = __fmul2_rn(dg, dgelu);
      if (!ok) { dhv = make_float2(0.f, 0.f); g = make_float2(0.f, 0.f); }
#pragma unroll
      for (int n = 0; n < NMAX; ++n) dx[n] = __ffma2_rn(make_float2(w1a[n], w1a[NMAX + n]), dhv, dx[n]);
      sDH[t * kSmallLd + threadIdx.x] = dhv.x;
      sDH[(t + 1) * kSmallLd + threadIdx.x] = dhv.y;
      sG[t * kSmallLd + threadIdx.x] = g.x;
      sG[(t + 1) * kSmallLd + threadIdx.x] = g.y;
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) {
        sXn[n * kSmallLd + threadIdx.x] = ok ? xn[n] : 0.f;
        sDU[n * kSmallLd + threadIdx.x] = dub[n];
        if (ok) dxn[base + static_cast<long long>(n) * D] = dx[n].x + dx[n].y;
      }
    __syncthreads();
    // phase 2: reduce the slab's 256 columns into the weight-gradient accumulators (registers, across slabs)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = threadIdx.x + i * kThreads;
      if (p < NT) {
        const int t = p / N, n = p - t * N;
        const float4* ph = reinterpret_cast<const float4*>(sDH + t * kSmallLd);
        const float4* pg = reinterpret_cast<const float4*>(sG + t * kSmallLd);
        const float4* px = reinterpret_cast<const float4*>(sXn + n * kSmallLd);
        const float4* pu = reinterpret_cast<const float4*>(sDU + n * kSmallLd);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
        for (int c = 0; c < kThreads / 4; ++c) {
          const float4 hh = ph[c], gg = pg[c], xx = px[c], uu = pu[c];
          s1 += hh.x * xx.x + hh.y * xx.y + hh.z * xx.z + hh.w * xx.w;
          s2 += uu.x * gg.x + uu.y * gg.y + uu.z * gg.z + uu.w * gg.w;
        }
        a1[i] += s1; a2[i] += s2;
      } else if (p < nred) {
        const float4* pr = reinterpret_cast<const float4*>(p < NT + T ? sDH + (p - NT) * kSmallLd : sDU + (p - NT - T) * kSmallLd);
        float s1 = 0.f;
#pragma unroll 4
        for (int c = 0; c < kThreads / 4; ++c) { const float4 v = pr[c]; s1 += v.x + v.y + v.z + v.w; }
        a1[i] += s1;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = threadIdx.x + i * kThreads;
    if (p < NT) {
      const int t = p / N, n = p - t * N;
      atomicAdd(&dw1[static_cast<long long>(t) * N + n], a1[i]);
      atomicAdd(&dw2[static_cast<long long>(n) * T + t], a2[i]);
    } else if (p < NT + T) {
      atomicAdd(&db1[p - NT], a1[i]);
    } else if (p < nred) {
      atomicAdd(&db2[p - NT - T], a1[i]);
    }
  }
}

inline bool small_ok(int N, int D, int T) {
  return N <= 32 && (D == 32 || D == 64 || D == 128 || D == 256) && N * T + T + N <= 4 * kThreads && 2 * T + 2 * N <= 160 && T <= 512;
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

int token_generation() {
  static const int gen = []() {
    const char* e = getenv("M2B200_TOKEN_GEN");
    return e ? atoi(e) : 2;
  }();
  return gen;
}

int token_mix_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* u, int B, int N, int D, int T, int exact_gelu, float drop_p,
                  unsigned long long seed, cudaStream_t s) {
  if (B <= 0 || N <= 0 || D <= 0 || T <= 0) return M2_ERR_ARG;
  if (!exact_gelu && token_generation() != 1 && token_mix_mma_supported(N, D, T))
    return token_mix_mma_fwd(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, D, T, drop_p, seed, s);
  if (small_ok(N, D, T)) {
    const Drop dh = make_drop(drop_p, seed, kSiteTokenHidden), dout = make_drop(drop_p, seed, kSiteTokenOut);
    const int spb = kThreads / D, Tp = (T + 1) & ~1;
    int grid = ceil_div(B, spb);
    if (grid > 148 * 8) grid = 148 * 8;
    LaunchScope scope("token_mix_small_fwd", s);
#define M2_TMS_FWD(NM_)                                                                                                  \
  {                                                                                                                      \
    const size_t smem = (static_cast<size_t>(2 * Tp * NM_ + Tp + NM_) + 8 * NM_) * 4;                                    \
    int rc = set_smem(token_mix_small_fwd_kernel<NM_>, smem);                                                            \
    if (rc) return rc;                                                                                                   \
    token_mix_small_fwd_kernel<NM_><<<grid, kThreads, smem, s>>>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, D, T, exact_gelu, \
                                                                 dh, dout);                                              \
  }
    if (N <= 4) M2_TMS_FWD(4) else if (N <= 8) M2_TMS_FWD(8) else if (N <= 16) M2_TMS_FWD(16) else M2_TMS_FWD(32)
#undef M2_TMS_FWD
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
  auto bytes = [&](int dt) { return static_cast<size_t>(N + T) * (dt + 1) * 4 + static_cast<size_t>(N) * 8; };
  int dt = 32;
  while (dt > 8 && bytes(dt) > kMaxSmem) dt >>= 1;
  if (bytes(dt) > kMaxSmem) return M2_ERR_ARG;
  const int ds = ceil_div(D, dt);
  int gx = B < ceil_div(148 * 8, ds) ? B : ceil_div(148 * 8, ds);
  dim3 grid(gx, ds);
  int rc;
  const Drop dh = make_drop(drop_p, seed, kSiteTokenHidden), dout = make_drop(drop_p, seed, kSiteTokenOut);
  LaunchScope scope("token_mix_fwd", s);
#define M2_TM_FWD(DT_)                                                                                         \
  rc = set_smem(token_mix_fwd_kernel<DT_>, bytes(DT_));                                                        \
  if (rc) return rc;                                                                                           \
  token_mix_fwd_kernel<DT_><<<grid, kThreads, bytes(DT_), s>>>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, D, T, exact_gelu, \
                                                               dh, dout);
  if (dt == 32) { M2_TM_FWD(32) } else if (dt == 16) { M2_TM_FWD(16) } else { M2_TM_FWD(8) }
#undef M2_TM_FWD
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int token_mix_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                  const float* w2, float* dxn, float* dw1, float* db1, float* dw2, float* db2, int B, int N, int D, int T,
                  int exact_gelu, float drop_p, unsigned long long seed, cudaStream_t s) {
  if (B <= 0 || N <= 0 || D <= 0 || T <= 0) return M2_ERR_ARG;
  if (small_ok(N, D, T)) {
    const Drop dh = make_drop(drop_p, seed, kSiteTokenHidden), dout = make_drop(drop_p, seed, kSiteTokenOut);
    const int spb = kThreads / D, Tp = (T + 1) & ~1;
    int grid = ceil_div(B, spb);
    if (grid > 148 * 3) grid = 148 * 3;
    LaunchScope scope("token_mix_small_bwd", s);
#define M2_TMS_BWD(NM_)                                                                                                  \
  {                                                                                                                      \
    const size_t smem = (static_cast<size_t>(2 * Tp * NM_ + Tp) + 8 * NM_ + 4 + static_cast<size_t>(2 * Tp + 2 * N) * kSmallLd) * 4; \
    int rc = set_smem(token_mix_small_bwd_kernel<NM_>, smem);                                                            \
    if (rc) return rc;                                                                                                   \
    token_mix_small_bwd_kernel<NM_><<<grid, kThreads, smem, s>>>(du, x, ln_w, ln_b, w1, b1, w2, dxn, dw1, db1, dw2, db2, B, \
                                                                 N, D, T, exact_gelu, dh, dout);                         \
  }
    if (N <= 4) M2_TMS_BWD(4) else if (N <= 8) M2_TMS_BWD(8) else if (N <= 16) M2_TMS_BWD(16) else M2_TMS_BWD(32)
#undef M2_TMS_BWD
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
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
  const Drop dh = make_drop(drop_p, seed, kSiteTokenHidden), dout = make_drop(drop_p, seed, kSiteTokenOut);
  LaunchScope scope("token_mix_bwd", s);
#define M2_TM_BWD(DT_)                                                                                            \
  rc = set_smem(token_mix_bwd_kernel<DT_>, bytes(DT_));                                                           \
  if (rc) return rc;                                                                                              \
  token_mix_bwd_kernel<DT_><<<grid, kThreads, bytes(DT_), s>>>(du, x, ln_w, ln_b, w1, b1, w2, dxn, dw1, db1, dw2, db2, B, N, \
                                                               D, T, exact_gelu, reg_acc, dh, dout);
  if (dt == 32) { M2_TM_BWD(32) } else if (dt == 16) { M2_TM_BWD(16) } else { M2_TM_BWD(8) }
#undef M2_TM_BWD
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace m2
