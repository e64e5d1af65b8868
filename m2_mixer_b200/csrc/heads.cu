// Per-modality + fused classifier heads with their summed loss: one warp-shuffle reduction kernel.
//
// Reference arithmetic (not in modules/losses.py):
//   logits_h = Linear_h(mean over tokens(x_h))                 models/avmnist.py:267-273, classification.py:90
//   CE form : L_h = mean_b CE(logits_h, y);  loss = sum_h head_weight[h] * L_h      models/avmnist.py:276-291,
//             (AV-MNIST: w = (w_f, ow, ow) * 3;  MIMIC: no *3, models/mimic.py:111-121)
//   BCE form: L_h = mean_{b,k} BCEWithLogits(pos_weight)       models/mmimdb.py:47-50,115-125
//   preds   : argmax (softmax is monotone) / logit > 0         models/avmnist.py:296-298, mmimdb.py:128-133
// One warp per sample: lanes stride the hidden axis for the token mean-pool, K dot products are warp-reduced,
// the per-sample loss terms are reduced per CTA and added to the 4 loss scalars with one atomic per CTA.
#include "common.cuh"
#include "kernels.h"

namespace m2 {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 64;
constexpr int kMaxPerCap = 32;   // hidden dim <= 1024

struct HeadsDev {
  const float* tok[3]; long long tok_bstride[3]; int ntok[3]; int dim[3];
  const float* w[3]; const float* b[3];
  int nheads, B, K, loss_kind;
  const void* labels; const float* pos_weight;
  float head_weight[3];
};

template <int kMaxPer>
__device__ __forceinline__ void pool_tokens(const float* t, int ntok, int dim, int lane, float (&pooled)[kMaxPer]) {
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) pooled[i] = 0.f;
  // four token rows are requested before the first one is added: with one row per trip the loop was a chain of ntok
  // dependent global-load round trips per sample and head (80 us for the whole MIMIC-H heads backward at batch 128)
  int n = 0;
  for (; n + 4 <= ntok; n += 4) {
    float v[4][kMaxPer];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* r = t + static_cast<long long>(n + k) * dim;
#pragma unroll
      for (int i = 0; i < kMaxPer; ++i) {
        const int d = lane + 32 * i;
        v[k][i] = d < dim ? r[d] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) pooled[i] += (v[0][i] + v[1][i]) + (v[2][i] + v[3][i]);
  }
  for (; n < ntok; ++n) {
    const float* r = t + static_cast<long long>(n) * dim;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int d = lane + 32 * i;
      if (d < dim) pooled[i] += r[d];
    }
  }
  const float inv = 1.f / ntok;
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) pooled[i] *= inv;
}

__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x))); }

template <int kMaxPer>
__global__ void __launch_bounds__(kThreads) heads_fwd_kernel(const HeadsDev a, float* __restrict__ logits,
                                                             float* __restrict__ losses, long long* __restrict__ preds) {
  __shared__ float s_logit[kWarps][kMaxK];
  __shared__ float s_loss[kWarps][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float lsum[3] = {0.f, 0.f, 0.f};
  for (int b = blockIdx.x * kWarps + warp; b < a.B; b += gridDim.x * kWarps) {
    for (int h = 0; h < a.nheads; ++h) {
      float pooled[kMaxPer];
      pool_tokens<kMaxPer>(a.tok[h] + b * a.tok_bstride[h], a.ntok[h], a.dim[h], lane, pooled);
      // eight classes at a time: their dot products are independent, so the eight warp reductions run interleaved
      // (one class after the other was a chain of K x 5 dependent shuffles per head and sample: the kernel was latency bound)
      for (int k0 = 0; k0 < a.K; k0 += 8) {
        float acc[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          acc[kk] = 0.f;
          if (k0 + kk < a.K) {
            const float* wr = a.w[h] + static_cast<long long>(k0 + kk) * a.dim[h];
#pragma unroll
            for (int i = 0; i < kMaxPer; ++i) {
              const int d = lane + 32 * i;
              if (d < a.dim[h]) acc[kk] = fmaf(__ldg(wr + d), pooled[i], acc[kk]);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) acc[kk] += __shfl_xor_sync(0xffffffffu, acc[kk], o);
        if (lane < 8 && k0 + lane < a.K) {
          float v = acc[0];
#pragma unroll
          for (int kk = 1; kk < 8; ++kk) v = lane == kk ? acc[kk] : v;
          v += a.b[h][k0 + lane];
          s_logit[warp][k0 + lane] = v;
          logits[(static_cast<long long>(h) * a.B + b) * a.K + k0 + lane] = v;
        }
      }
      __syncwarp();
      if (a.loss_kind == 0) {
        // cross entropy: lanes stride k
        float mx = -INFINITY;
        int arg = 0;
        for (int k = lane; k < a.K; k += 32) {
          const float v = s_logit[warp][k];
          if (v > mx) { mx = v; arg = k; }
        }
        // argmax with lowest-index tie-break (torch.argmax returns the first maximal index)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float om = __shfl_xor_sync(0xffffffffu, mx, o);
          const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
          if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float se = 0.f;
        for (int k = lane; k < a.K; k += 32) se += __expf(s_logit[warp][k] - mx);
        se = warp_sum(se);
        if (lane == 0) {
          const long long y = static_cast<const long long*>(a.labels)[b];
          lsum[h] += (mx + logf(se)) - s_logit[warp][y];
          preds[static_cast<long long>(h) * a.B + b] = arg;
        }
      } else {
        float l = 0.f;
        for (int k = lane; k < a.K; k += 32) {
          const float x = s_logit[warp][k];
          const float y = static_cast<const float*>(a.labels)[static_cast<long long>(b) * a.K + k];
          const float pw = a.pos_weight ? a.pos_weight[k] : 1.f;
          l += pw * y * softplus(-x) + (1.f - y) * softplus(x);
          preds[(static_cast<long long>(h) * a.B + b) * a.K + k] = x > 0.f ? 1 : 0;
        }
        l = warp_sum(l);
        if (lane == 0) lsum[h] += l;
      }
      __syncwarp();
    }
  }
  if (lane == 0) { s_loss[warp][0] = lsum[0]; s_loss[warp][1] = lsum[1]; s_loss[warp][2] = lsum[2]; }
  __syncthreads();
  if (threadIdx.x < a.nheads) {
    const int h = threadIdx.x;
    float t = 0.f;
    for (int w = 0; w < kWarps; ++w) t += s_loss[w][h];
    t *= (a.loss_kind == 0) ? 1.f / a.B : 1.f / (static_cast<float>(a.B) * a.K);
    atomicAdd(&losses[1 + h], t);
    atomicAdd(&losses[0], a.head_weight[h] * t);
  }
}

struct HeadsBwdDev {
  float* dtok[3]; long long dtok_bstride[3]; int accumulate[3];
  float* dw[3]; float* db[3];
  float grad_scale; const float* grad_scale_dev;
};

template <int kMaxPer>
__global__ void __launch_bounds__(kThreads) heads_bwd_kernel(const HeadsDev a, const HeadsBwdDev g,
                                                             const float* __restrict__ logits) {
  extern __shared__ float sm[];   // per-head dW partial [K][dim] and db partial [K], laid out back to back
  __shared__ float s_dl[kWarps][kMaxK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int off[4];
  off[0] = 0;
  for (int h = 0; h < a.nheads; ++h) off[h + 1] = off[h] + a.K * a.dim[h] + a.K;
  for (int i = threadIdx.x; i < off[a.nheads]; i += kThreads) sm[i] = 0.f;
  __syncthreads();
  for (int b = blockIdx.x * kWarps + warp; b < a.B; b += gridDim.x * kWarps) {
    for (int h = 0; h < a.nheads; ++h) {
      const int dim = a.dim[h];
      float pooled[kMaxPer];
      pool_tokens<kMaxPer>(a.tok[h] + b * a.tok_bstride[h], a.ntok[h], dim, lane, pooled);
      const float* lg = logits + (static_cast<long long>(h) * a.B + b) * a.K;
      const float hw = a.head_weight[h] * g.grad_scale * (g.grad_scale_dev ? g.grad_scale_dev[0] : 1.f);
      if (a.loss_kind == 0) {
        float mx = -INFINITY;
        for (int k = lane; k < a.K; k += 32) mx = fmaxf(mx, lg[k]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int k = lane; k < a.K; k += 32) se += __expf(lg[k] - mx);
        se = warp_sum(se);
        const long long y = static_cast<const long long*>(a.labels)[b];
        for (int k = lane; k < a.K; k += 32)
          s_dl[warp][k] = (__expf(lg[k] - mx) / se - (k == y ? 1.f : 0.f)) * hw / a.B;
      } else {
        for (int k = lane; k < a.K; k += 32) {
          const float x = lg[k];
          const float y = static_cast<const float*>(a.labels)[static_cast<long long>(b) * a.K + k];
          const float pw = a.pos_weight ? a.pos_weight[k] : 1.f;
          const float sg = 1.f / (1.f + __expf(-x));
          s_dl[warp][k] = (-pw * y * (1.f - sg) + (1.f - y) * sg) * hw / (static_cast<float>(a.B) * a.K);
        }
      }
      __syncwarp();
      float* sdw = sm + off[h];
      float* sdb = sdw + a.K * dim;
      float dp[kMaxPer];
#pragma unroll
      for (int i = 0; i < kMaxPer; ++i) dp[i] = 0.f;
      for (int k = 0; k < a.K; ++k) {
        const float dl = s_dl[warp][k];
        const float* wr = a.w[h] + static_cast<long long>(k) * dim;
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) {
          const int d = lane + 32 * i;
          if (d < dim) {
            atomicAdd(&sdw[k * dim + d], dl * pooled[i]);
            dp[i] = fmaf(wr[d], dl, dp[i]);
          }
        }
        if (lane == 0) atomicAdd(&sdb[k], dl);
      }
      if (g.dtok[h]) {
        const float inv = 1.f / a.ntok[h];
        float* dt = g.dtok[h] + b * g.dtok_bstride[h];
        for (int n = 0; n < a.ntok[h]; ++n) {
#pragma unroll
          for (int i = 0; i < kMaxPer; ++i) {
            const int d = lane + 32 * i;
            if (d < dim) {
              float* p = dt + static_cast<long long>(n) * dim + d;
              *p = g.accumulate[h] ? *p + dp[i] * inv : dp[i] * inv;
            }
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int h = 0; h < a.nheads; ++h) {
    const int nw = a.K * a.dim[h];
    for (int i = threadIdx.x; i < nw; i += kThreads) atomicAdd(&g.dw[h][i], sm[off[h] + i]);
    for (int i = threadIdx.x; i < a.K; i += kThreads) atomicAdd(&g.db[h][i], sm[off[h] + nw + i]);
  }
}

// Register-accumulating variant for the common small heads (K <= kKr classes, dim <= 32 * kMaxPer): the weight-gradient
// outer products of a warp's samples are summed in registers and leave the warp ONCE per head (plain stores to a
// per-warp slot, tree-reduced over the warps, one global atomic per element and CTA).  The generic kernel above pays a
// contended shared-memory atomic per (sample, class, column): 160 us of the M2-Mixer-B step for 64 MB of real traffic.
template <int kMaxPer, int kKr>
__global__ void __launch_bounds__(kThreads) heads_bwd_reg_kernel(const HeadsDev a, const HeadsBwdDev g,
                                                                 const float* __restrict__ logits) {
  extern __shared__ float red[];   // [kWarps][K * dim + K]
  __shared__ float s_dl[kWarps][kKr];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int h = blockIdx.y;   // one head per CTA row: the heads are independent, three times the CTAs in flight
    const int dim = a.dim[h];
    const int slot = a.K * dim + a.K;
    float dwacc[kKr][kMaxPer];
#pragma unroll
    for (int k = 0; k < kKr; ++k)
#pragma unroll
      for (int i = 0; i < kMaxPer; ++i) dwacc[k][i] = 0.f;
    float dbacc = 0.f;
    const float hw = a.head_weight[h] * g.grad_scale * (g.grad_scale_dev ? g.grad_scale_dev[0] : 1.f);
    for (int b = blockIdx.x * kWarps + warp; b < a.B; b += gridDim.x * kWarps) {
      float pooled[kMaxPer];
      pool_tokens<kMaxPer>(a.tok[h] + b * a.tok_bstride[h], a.ntok[h], dim, lane, pooled);
      const float* lg = logits + (static_cast<long long>(h) * a.B + b) * a.K;
      float dl_mine = 0.f;
      if (a.loss_kind == 0) {
        const float v = lane < a.K ? lg[lane] : -INFINITY;
        const float mx = warp_max(v);
        const float e = lane < a.K ? __expf(v - mx) : 0.f;
        const float se = warp_sum(e);
        const long long y = static_cast<const long long*>(a.labels)[b];
        if (lane < a.K) dl_mine = (e / se - (lane == y ? 1.f : 0.f)) * hw / a.B;
      } else if (lane < a.K) {
        const float x = lg[lane];
        const float y = static_cast<const float*>(a.labels)[static_cast<long long>(b) * a.K + lane];
        const float pw = a.pos_weight ? a.pos_weight[lane] : 1.f;
        const float sg = 1.f / (1.f + __expf(-x));
        dl_mine = (-pw * y * (1.f - sg) + (1.f - y) * sg) * hw / (static_cast<float>(a.B) * a.K);
      }
      dbacc += dl_mine;
      if (lane < kKr) s_dl[warp][lane] = dl_mine;
      __syncwarp();
      float dp[kMaxPer];
#pragma unroll
      for (int i = 0; i < kMaxPer; ++i) dp[i] = 0.f;
#pragma unroll
      for (int k = 0; k < kKr; ++k) {
        if (k < a.K) {
          const float dl = s_dl[warp][k];
          const float* wr = a.w[h] + static_cast<long long>(k) * dim;
#pragma unroll
          for (int i = 0; i < kMaxPer; ++i) {
            const int d = lane + 32 * i;
            if (d < dim) {
              dwacc[k][i] = fmaf(dl, pooled[i], dwacc[k][i]);
              dp[i] = fmaf(__ldg(wr + d), dl, dp[i]);
            }
          }
        }
      }
      if (g.dtok[h]) {
        const float inv = 1.f / a.ntok[h];
        float* dt = g.dtok[h] + b * g.dtok_bstride[h];
        for (int n = 0; n < a.ntok[h]; ++n) {
#pragma unroll
          for (int i = 0; i < kMaxPer; ++i) {
            const int d = lane + 32 * i;
            if (d < dim) {
              float* p = dt + static_cast<long long>(n) * dim + d;
              *p = g.accumulate[h] ? *p + dp[i] * inv : dp[i] * inv;
            }
          }
        }
      }
      __syncwarp();
    }
    // ---- leave the warp once per head
    float* mine = red + warp * slot;
#pragma unroll
    for (int k = 0; k < kKr; ++k)
      if (k < a.K) {
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) {
          const int d = lane + 32 * i;
          if (d < dim) mine[k * dim + d] = dwacc[k][i];
        }
      }
    if (lane < a.K) mine[a.K * dim + lane] = dbacc;
    __syncthreads();
    for (int i = threadIdx.x; i < slot; i += kThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += red[w * slot + i];
      if (i < a.K * dim) atomicAdd(&g.dw[h][i], t);
      else atomicAdd(&g.db[h][i - a.K * dim], t);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Vectorised forms for the shipped head shapes (hidden dims that are multiples of 4 and <= 128 kV, K <= 16 classes, weights of
// all heads <= 64 KB).  The kernels above spent 2760 / 970 warp instructions per sample (and head) on 4-byte token loads and on
// re-reading the K weight rows from L1 / L2 for every sample (ncu: long-scoreboard stalls on those loads, 11 M warp
// instructions per launch, 30 / 56 us at batch 4096 for 34 / 84 MB of traffic).  Here: the weight rows are staged in shared
// memory once per CTA, a lane owns 4 kV adjacent columns (one 16-byte load per token row), every token row of a sample is
// requested before the first is added, the K dot products are xor-reduced together so that every lane holds every logit
// (the softmax / argmax then needs no further shuffle), and the backward broadcasts the K logit gradients with shuffles.
template <int kV>
__device__ __forceinline__ void pool_tokens_vec(const float* t, int ntok, int dim4, int lane, float4 (&pooled)[kV]) {
#pragma unroll
  for (int i = 0; i < kV; ++i) pooled[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  for (int n0 = 0; n0 < ntok; n0 += 8) {
    float4 v[8][kV];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        const int c = lane + 32 * i;
        v[k][i] = (n0 + k < ntok && c < dim4) ? t4[static_cast<long long>(n0 + k) * dim4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        pooled[i].x += v[k][i].x; pooled[i].y += v[k][i].y; pooled[i].z += v[k][i].z; pooled[i].w += v[k][i].w;
      }
  }
  const float inv = 1.f / ntok;
#pragma unroll
  for (int i = 0; i < kV; ++i) { pooled[i].x *= inv; pooled[i].y *= inv; pooled[i].z *= inv; pooled[i].w *= inv; }
}

constexpr int kKv = 16;   // classes the vectorised kernels hold in registers

// 4 CTAs per SM (<= 64 registers): at batch 4096 the 512 CTAs (one sample per warp) must all be resident - with 80 registers
// 444 were, and the 68 left over ran as a second wave (53 us instead of 41 us for the launch).
template <int kV>
__global__ void __launch_bounds__(kThreads, 4) heads_fwd_vec_kernel(const HeadsDev a, float* __restrict__ logits,
                                                                 float* __restrict__ losses, long long* __restrict__ preds) {
  extern __shared__ float4 sW4[];   // per head [K][dim / 4]
  __shared__ float s_loss[kWarps][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int off[4];
  off[0] = 0;
  for (int h = 0; h < a.nheads; ++h) off[h + 1] = off[h] + a.K * (a.dim[h] >> 2);
  // the weight rows are staged asynchronously (cp.async) while the first sample's token rows are on their way
  for (int h = 0; h < a.nheads; ++h)
    for (int i = threadIdx.x; i < a.K * (a.dim[h] >> 2); i += kThreads)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sW4 + off[h] + i)),
                   "l"(reinterpret_cast<const float4*>(a.w[h]) + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  float lsum[3] = {0.f, 0.f, 0.f};
  // the token rows of ALL heads are requested before the first dot product: one HBM round trip per sample instead of one
  // per head (a warp's sample is a single dependency chain, and at one sample per warp the chain is the kernel)
  float4 pooled_all[3][kV];
  auto pool_all = [&](int b) {
#pragma unroll
    for (int h = 0; h < 3; ++h)
      if (h < a.nheads) pool_tokens_vec<kV>(a.tok[h] + b * a.tok_bstride[h], a.ntok[h], a.dim[h] >> 2, lane, pooled_all[h]);
  };
  int b = blockIdx.x * kWarps + warp;
  if (b < a.B) pool_all(b);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (; b < a.B;) {
#pragma unroll
    for (int h = 0; h < 3; ++h) {
      if (h >= a.nheads) break;
      const int dim4 = a.dim[h] >> 2;
      float4 (&pooled)[kV] = pooled_all[h];
      float acc[kKv];
#pragma unroll
      for (int k = 0; k < kKv; ++k) {
        acc[k] = 0.f;
        if (k < a.K) {
#pragma unroll
          for (int i = 0; i < kV; ++i) {
            const int c = lane + 32 * i;
            if (c < dim4) {
              const float4 w = sW4[off[h] + k * dim4 + c];
              acc[k] += w.x * pooled[i].x + w.y * pooled[i].y + w.z * pooled[i].z + w.w * pooled[i].w;
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < kKv; ++k)
          if (k < a.K) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      // every lane now holds every logit (minus the bias)
      float mine = 0.f;
#pragma unroll
      for (int k = 0; k < kKv; ++k) {
        if (k < a.K) acc[k] += __ldg(a.b[h] + k);
        mine = lane == k ? acc[k] : mine;
      }
      if (lane < a.K) logits[(static_cast<long long>(h) * a.B + b) * a.K + lane] = mine;
      if (a.loss_kind == 0) {
        float mx = acc[0];
        int arg = 0;
#pragma unroll
        for (int k = 1; k < kKv; ++k)
          if (k < a.K && acc[k] > mx) { mx = acc[k]; arg = k; }   // first maximal index, as torch.argmax
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < kKv; ++k)
          if (k < a.K) se += __expf(acc[k] - mx);
        if (lane == 0) {
          const int y = static_cast<int>(static_cast<const long long*>(a.labels)[b]);
          float ly = acc[0];
#pragma unroll
          for (int k = 1; k < kKv; ++k) ly = y == k ? acc[k] : ly;
          lsum[h] += (mx + logf(se)) - ly;
          preds[static_cast<long long>(h) * a.B + b] = arg;
        }
      } else {
        float l = 0.f;
        if (lane < a.K) {
          const float y = static_cast<const float*>(a.labels)[static_cast<long long>(b) * a.K + lane];
          const float pw = a.pos_weight ? a.pos_weight[lane] : 1.f;
          l = pw * y * softplus(-mine) + (1.f - y) * softplus(mine);
          preds[(static_cast<long long>(h) * a.B + b) * a.K + lane] = mine > 0.f ? 1 : 0;
        }
        l = warp_sum(l);
        if (lane == 0) lsum[h] += l;
      }
    }
    b += gridDim.x * kWarps;
    if (b < a.B) pool_all(b);
  }
  if (lane == 0) { s_loss[warp][0] = lsum[0]; s_loss[warp][1] = lsum[1]; s_loss[warp][2] = lsum[2]; }
  __syncthreads();
  if (threadIdx.x < a.nheads) {
    const int h = threadIdx.x;
    float t = 0.f;
    for (int w = 0; w < kWarps; ++w) t += s_loss[w][h];
    t *= (a.loss_kind == 0) ? 1.f / a.B : 1.f / (static_cast<float>(a.B) * a.K);
    atomicAdd(&losses[1 + h], t);
    atomicAdd(&losses[0], a.head_weight[h] * t);
  }
}

// Backward, one head per CTA row (blockIdx.y), dim <= 128: a lane owns 4 adjacent columns; the weight-gradient outer products
// of a warp's samples are summed in registers and leave the warp once (as heads_bwd_reg_kernel).
__global__ void __launch_bounds__(kThreads, 2) heads_bwd_vec_kernel(const HeadsDev a, const HeadsBwdDev g,
                                                                 const float* __restrict__ logits) {
  extern __shared__ float4 sm4[];   // [K][dim / 4] weights of the head, then [kWarps][K * dim + K] reduction slots
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int dim = a.dim[h], dim4 = dim >> 2;
  const int nred = a.K * dim + a.K;
  const int slot = (nred + 3) & ~3;   // 16-byte aligned per-warp slots (float4 stores below)
  float* red = reinterpret_cast<float*>(sm4 + a.K * dim4);
  for (int i = threadIdx.x; i < a.K * dim4; i += kThreads) sm4[i] = reinterpret_cast<const float4*>(a.w[h])[i];
  __syncthreads();
  const bool live = lane < dim4;
  float4 dwacc[kKv];
#pragma unroll
  for (int k = 0; k < kKv; ++k) dwacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dbacc = 0.f;
  const float hw = a.head_weight[h] * g.grad_scale * (g.grad_scale_dev ? g.grad_scale_dev[0] : 1.f);
  for (int b = blockIdx.x * kWarps + warp; b < a.B; b += gridDim.x * kWarps) {
    float4 pooled[1];
    pool_tokens_vec<1>(a.tok[h] + b * a.tok_bstride[h], a.ntok[h], dim4, lane, pooled);
    const float* lg = logits + (static_cast<long long>(h) * a.B + b) * a.K;
    float dl_mine = 0.f;
    if (a.loss_kind == 0) {
      const float v = lane < a.K ? lg[lane] : -INFINITY;
      const float mx = warp_max(v);
      const float e = lane < a.K ? __expf(v - mx) : 0.f;
      const float se = warp_sum(e);
      const long long y = static_cast<const long long*>(a.labels)[b];
      if (lane < a.K) dl_mine = (e / se - (lane == y ? 1.f : 0.f)) * hw / a.B;
    } else if (lane < a.K) {
      const float x = lg[lane];
      const float y = static_cast<const float*>(a.labels)[static_cast<long long>(b) * a.K + lane];
      const float pw = a.pos_weight ? a.pos_weight[lane] : 1.f;
      const float sg = 1.f / (1.f + __expf(-x));
      dl_mine = (-pw * y * (1.f - sg) + (1.f - y) * sg) * hw / (static_cast<float>(a.B) * a.K);
    }
    dbacc += dl_mine;
    float4 dp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kKv; ++k) {
      if (k < a.K) {
        const float dl = __shfl_sync(0xffffffffu, dl_mine, k);
        if (live) {
          const float4 w = sm4[k * dim4 + lane];
          dwacc[k].x = fmaf(dl, pooled[0].x, dwacc[k].x); dwacc[k].y = fmaf(dl, pooled[0].y, dwacc[k].y);
          dwacc[k].z = fmaf(dl, pooled[0].z, dwacc[k].z); dwacc[k].w = fmaf(dl, pooled[0].w, dwacc[k].w);
          dp.x = fmaf(w.x, dl, dp.x); dp.y = fmaf(w.y, dl, dp.y); dp.z = fmaf(w.z, dl, dp.z); dp.w = fmaf(w.w, dl, dp.w);
        }
      }
    }
    if (g.dtok[h] && live) {
      const float inv = 1.f / a.ntok[h];
      dp.x *= inv; dp.y *= inv; dp.z *= inv; dp.w *= inv;
      float4* dt = reinterpret_cast<float4*>(g.dtok[h] + b * g.dtok_bstride[h]) + lane;
      if (g.accumulate[h]) {
        for (int n0 = 0; n0 < a.ntok[h]; n0 += 4) {   // four rows requested before the first add
          float4 o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (n0 + k < a.ntok[h]) o[k] = dt[static_cast<long long>(n0 + k) * dim4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (n0 + k < a.ntok[h])
              dt[static_cast<long long>(n0 + k) * dim4] = make_float4(o[k].x + dp.x, o[k].y + dp.y, o[k].z + dp.z, o[k].w + dp.w);
        }
      } else {
        for (int n = 0; n < a.ntok[h]; ++n) dt[static_cast<long long>(n) * dim4] = dp;
      }
    }
  }
  // ---- leave the warp once
  float* mine = red + warp * slot;
#pragma unroll
  for (int k = 0; k < kKv; ++k)
    if (k < a.K && live) *reinterpret_cast<float4*>(mine + k * dim + 4 * lane) = dwacc[k];
  if (lane < a.K) mine[a.K * dim + lane] = dbacc;
  __syncthreads();
  for (int i = threadIdx.x; i < nred; i += kThreads) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[w * slot + i];
    if (i < a.K * dim) atomicAdd(&g.dw[h][i], t);
    else atomicAdd(&g.db[h][i - a.K * dim], t);
  }
}

// shapes the vectorised kernels cover: every head's dim a multiple of 4 (16-byte token rows), K <= 16
bool heads_vec_ok(const HeadsArgs& a, int max_dim, const float* const* dtok, const long long* dtok_bstride) {
  if (a.K > kKv) return false;
  for (int h = 0; h < a.nheads; ++h) {
    if (a.dim[h] % 4 || a.dim[h] > max_dim || a.tok_bstride[h] % 4) return false;
    if ((reinterpret_cast<uintptr_t>(a.tok[h]) | reinterpret_cast<uintptr_t>(a.w[h])) & 15) return false;
    if (dtok && dtok[h] && ((reinterpret_cast<uintptr_t>(dtok[h]) & 15) || dtok_bstride[h] % 4)) return false;
  }
  return true;
}

int fill_dev(const HeadsArgs& a, HeadsDev* d) {
  if (a.nheads < 1 || a.nheads > 3 || a.B <= 0 || a.K <= 0 || a.K > kMaxK || !a.labels) return M2_ERR_ARG;
  for (int h = 0; h < a.nheads; ++h) {
    if (!a.tok[h] || !a.w[h] || !a.b[h] || a.ntok[h] <= 0 || a.dim[h] <= 0 || a.dim[h] > 32 * kMaxPerCap) return M2_ERR_ARG;
    d->tok[h] = a.tok[h]; d->tok_bstride[h] = a.tok_bstride[h]; d->ntok[h] = a.ntok[h]; d->dim[h] = a.dim[h];
    d->w[h] = a.w[h]; d->b[h] = a.b[h]; d->head_weight[h] = a.head_weight[h];
  }
  d->nheads = a.nheads; d->B = a.B; d->K = a.K; d->loss_kind = a.loss_kind;
  d->labels = a.labels; d->pos_weight = a.pos_weight;
  return M2_OK;
}

}  // namespace

int heads_loss_fwd(const HeadsArgs& a, float* logits, float* losses, long long* preds, cudaStream_t s) {
  HeadsDev d = {};
  int rc = fill_dev(a, &d);
  if (rc) return rc;
  if (!logits || !losses || !preds) return M2_ERR_ARG;
  LaunchScope scope("heads_loss_fwd", s);
  if (cudaMemsetAsync(losses, 0, 4 * sizeof(float), s) != cudaSuccess) return M2_ERR_LAUNCH;
  int grid = ceil_div(a.B, kWarps);
  if (grid > 148 * 4) grid = 148 * 4;
  int per = 1;
  for (int h = 0; h < a.nheads; ++h) per = per > ceil_div(a.dim[h], 32) ? per : ceil_div(a.dim[h], 32);
  if (heads_vec_ok(a, 256, nullptr, nullptr)) {
    size_t wsm = 0;
    for (int h = 0; h < a.nheads; ++h) wsm += static_cast<size_t>(a.K) * a.dim[h] * sizeof(float);
    if (wsm <= 64 * 1024) {
      auto kern = per <= 4 ? heads_fwd_vec_kernel<1> : heads_fwd_vec_kernel<2>;
      if (wsm > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(wsm)) != cudaSuccess)
        return M2_ERR_LAUNCH;
      int vgrid = ceil_div(a.B, kWarps);      // one sample per warp while the CTAs are co-resident: the per-sample chain is the
      if (vgrid > 148 * 4) vgrid = 148 * 4;   // kernel's critical path (296 CTAs with two samples per warp measured 59 vs 41 us)
      kern<<<vgrid, kThreads, wsm, s>>>(d, logits, losses, preds);
      M2_LAUNCH_CHECK();
      return M2_OK;
    }
  }
  if (per <= 2) heads_fwd_kernel<2><<<grid, kThreads, 0, s>>>(d, logits, losses, preds);
  else if (per <= 4) heads_fwd_kernel<4><<<grid, kThreads, 0, s>>>(d, logits, losses, preds);
  else if (per <= 8) heads_fwd_kernel<8><<<grid, kThreads, 0, s>>>(d, logits, losses, preds);
  else heads_fwd_kernel<32><<<grid, kThreads, 0, s>>>(d, logits, losses, preds);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int heads_loss_bwd(const HeadsArgs& a, const float* logits, float grad_scale, const float* grad_scale_dev, float* dtok[3], long long dtok_bstride[3],
                   int accumulate_dtok[3], float* dw[3], float* db[3], cudaStream_t s) {
  HeadsDev d = {};
  int rc = fill_dev(a, &d);
  if (rc) return rc;
  HeadsBwdDev g = {};
  size_t smem = 0;
  for (int h = 0; h < a.nheads; ++h) {
    if (!dw[h] || !db[h]) return M2_ERR_ARG;
    g.dtok[h] = dtok[h]; g.dtok_bstride[h] = dtok_bstride[h]; g.accumulate[h] = accumulate_dtok[h];
    g.dw[h] = dw[h]; g.db[h] = db[h];
    smem += static_cast<size_t>(a.K) * a.dim[h] + a.K;
  }
  smem *= sizeof(float);
  g.grad_scale = grad_scale; g.grad_scale_dev = grad_scale_dev;
  if (smem > 200 * 1024) return M2_ERR_ARG;
  int per = 1;
  for (int h = 0; h < a.nheads; ++h) per = per > ceil_div(a.dim[h], 32) ? per : ceil_div(a.dim[h], 32);
  int grid = ceil_div(a.B, kWarps);
  if (grid > 148) grid = 148;
  LaunchScope scope("heads_loss_bwd", s);
  if (heads_vec_ok(a, 128, dtok, dtok_bstride)) {   // vectorised register-accumulating variant
    int maxdim = 0;
    for (int h = 0; h < a.nheads; ++h) maxdim = maxdim > a.dim[h] ? maxdim : a.dim[h];
    const size_t vsm = (static_cast<size_t>(a.K) * maxdim + static_cast<size_t>(kWarps) * ((static_cast<size_t>(a.K) * maxdim + a.K + 3) & ~size_t(3))) * sizeof(float);
    if (vsm > 48 * 1024 && cudaFuncSetAttribute(heads_bwd_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(vsm)) != cudaSuccess)
      return M2_ERR_LAUNCH;
    int rgrid = ceil_div(a.B, kWarps * 4);   // >= 4 samples per warp: the per-head flush is amortised
    if (rgrid > 148 * 2) rgrid = 148 * 2;
    if (rgrid < 1) rgrid = 1;
    heads_bwd_vec_kernel<<<dim3(rgrid, a.nheads), kThreads, vsm, s>>>(d, g, logits);
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
  if (a.K <= 16 && per <= 4) {   // register-accumulating variant
    int maxdim = 0;
    for (int h = 0; h < a.nheads; ++h) maxdim = maxdim > a.dim[h] ? maxdim : a.dim[h];
    const size_t rsm = static_cast<size_t>(kWarps) * (static_cast<size_t>(a.K) * maxdim + a.K) * sizeof(float);
    auto kern = per <= 2 ? heads_bwd_reg_kernel<2, 16> : heads_bwd_reg_kernel<4, 16>;
    if (rsm > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(rsm)) != cudaSuccess)
      return M2_ERR_LAUNCH;
    int rgrid = ceil_div(a.B, kWarps * 4);   // >= 4 samples per warp: the per-head flush is amortised
    if (rgrid > 148 * 2) rgrid = 148 * 2;
    if (rgrid < 1) rgrid = 1;
    kern<<<dim3(rgrid, a.nheads), kThreads, rsm, s>>>(d, g, logits);
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
#define M2_HB(P_)                                                                                                      \
  {                                                                                                                    \
    if (smem > 48 * 1024 && cudaFuncSetAttribute(heads_bwd_kernel<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                                 static_cast<int>(smem)) != cudaSuccess)                               \
      return M2_ERR_LAUNCH;                                                                                            \
    heads_bwd_kernel<P_><<<grid, kThreads, smem, s>>>(d, g, logits);                                                   \
  }
  if (per <= 2) M2_HB(2) else if (per <= 4) M2_HB(4) else if (per <= 8) M2_HB(8) else M2_HB(32)
#undef M2_HB
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace m2
