// Fused channel-mixing chains of a Mixer block on tcgen05 / TMEM / TMA (bf16 operands, fp32 accumulate).
//
// Reference arithmetic: MixerBlock.channel_mix, modules/mixer.py:37-40,45
//     y = u + Drop(W2 . Drop(GELU(W1 . LN(u) + b1)) + b2)     per token row (M = B*N rows, D hidden, C channel_dim)
//
// FORWARD (chain_fwd_kernel), one CTA per 128-row token tile, the [128 x C] hidden activation never leaves the SM:
//   prologue   all warps: LayerNorm the fp32 rows, round to bf16, store as the K-major SW128 A operand (sX)
//   C is walked in chunks of 64 channels.  Weights arrive by TMA into two independent rings (W1 chunk [64 x D],
//   W2 chunk [D x 64]).  The MMA issuer runs GEMM1 kNB chunks AHEAD of the epilogue:
//     GEMM1(j)   Hacc[j % kNB][128x64] = sX . W1_j^T              kNB TMEM accumulator buffers
//     epilogue   group g (4 warps) owns chunks j = g (mod 2): Hacc -> regs, +b1, GELU, (dropout), bf16 -> sG[g]
//     GEMM2(j)   Yacc[128xD] += sG[j&1] . W2_j^T
//   so that when a group finishes chunk j its next chunk's accumulator is already complete (no MMA round trip on
//   the epilogue's critical path) and the two groups hide each other's TMEM-load / MUFU / store latencies.
//   final      Yacc -> regs, +b2, (dropout), +u (residual), fp32 store
//
// BACKWARD part A (chain_bwd_kernel), same tiling, given dY:
//   prologue : LN(u) -> sX ; dY (masked by the output dropout) -> bf16 sdY ; both also written to HBM as bf16
//              (operands of the weight-gradient GEMMs)
//   per chunk: H   = sX  . W1_j^T     (recompute)                 -> TMEM
//              dG  = sdY . W2_j       (W2 tile consumed MN-major: no transposed weight copy)
//              epilogue: G = Drop(GELU(H+b1)), dH = dG * Drop'(.) * GELU'(H+b1); G,dH -> HBM (bf16), dH -> smem
//              dXn += sdH . W1_j      (W1 tile consumed MN-major)  -> TMEM accumulator
//   final    : dXn -> HBM fp32 (LayerNorm backward + residual is a separate bandwidth kernel)
//   The weight gradients dW2 = dY^T.G and dW1 = dH^T.LN(u) are token-axis contractions done by the generic
//   tcgen05 GEMM (umma_gemm.cu) with both operands MN-major.
//
// D is padded in shared memory to DP in {64,128,256}; C is arbitrary (TMA zero-fills the ragged last chunk, the
// bf16 weight copies are padded to a multiple of 8 columns so their row stride is 16-byte aligned).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"

namespace m2 {
namespace {

constexpr int kRows = 128;      // token rows per CTA (UMMA M)
constexpr int kCc = 64;         // channels per chunk
constexpr int kThreads = 320;   // warp0 TMA, warp1 MMA, warps 2-9 epilogue: two groups of 4 warps
constexpr int kGBytes = kRows * kCc * 2;   // one [128 x 64] bf16 tile
constexpr int kBarBytes = 512;

template <int DP>
struct Cfg {
  static constexpr int kPanels = DP / 64;
  static constexpr int kXBytes = kRows * DP * 2;
  static constexpr int kW1Bytes = kCc * DP * 2;      // [64 c-rows][DP d]
  static constexpr int kW2Bytes = DP * kCc * 2;      // [DP d-rows][64 c]
  // forward: kNB accumulator buffers, GEMM1 issued kNB chunks ahead
  static constexpr int kFwdNB = 4;
  static constexpr int kFwdS1 = DP == 256 ? 2 : 3;
  static constexpr int kFwdS2 = DP == 256 ? 2 : 3;
  static constexpr int kFwdSmem = kXBytes + kFwdS1 * kW1Bytes + kFwdS2 * kW2Bytes + 2 * kGBytes + kBarBytes + 1024;
  static constexpr int kFwdTmem = 512;               // kNB * 64 (H) + DP (Y) <= 512 for DP <= 256
  static constexpr int kFwdYCol = kFwdNB * kCc;
  // backward: H and dG triple buffered, W1 stays resident until its dXn GEMM (kNB + 2 slots)
  static constexpr int kBwdNB = 3;
  static constexpr int kBwdS1 = 5;
  static constexpr int kBwdS2 = 2;
  static constexpr int kBwdSmem = 2 * kXBytes + kBwdS1 * kW1Bytes + kBwdS2 * kW2Bytes + 2 * kGBytes + kBarBytes + 1024;
  static constexpr int kBwdTmem = 512;               // 3*64 (H) + 3*64 (dG) + DP (dXn), DP <= 128
  static constexpr int kBwdGCol = kBwdNB * kCc;
  static constexpr int kBwdXCol = 2 * kBwdNB * kCc;
  static constexpr int kMaxBiasFwd = DP > 128 ? 0 : 32 * 1024;
  static constexpr int kMaxBiasBwd = DP > 128 ? 0 : 14 * 1024;
};

struct ChainParams {
  const float* u;        // [M][D] block input (pre-LN residual stream)
  const float* ln_w; const float* ln_b;
  const float* b1;       // [C]
  const float* b2;       // [D]
  float* y;              // fwd: [M][D]
  const float* dy;       // bwd: [M][D]
  __nv_bfloat16* xn_b;   // bwd out: LN(u) bf16 [M][D]
  __nv_bfloat16* dy_b;   // bwd out: dY (masked) bf16 [M][D]
  __nv_bfloat16* g_b;    // bwd out: G  bf16    [M][ldh]
  __nv_bfloat16* dh_b;   // bwd out: dH bf16    [M][ldh]
  float* dxn;            // bwd out: dL/dLN(u) fp32 [M][D]
  int M, D, C, ldh;
  int bias_smem;         // b1 (zero padded to a multiple of 64) is staged in shared memory
  Drop dh, dout;         // dropout after GELU (index row*ldh + c) and after the second Linear (index row*D + d)
};

// LayerNorm the tile's rows into the swizzled bf16 A operand; optional bf16 copy to HBM.  kB rows are in flight
// per warp so that the global-load latency is paid once per batch, not once per row.
template <int DP>
__device__ __forceinline__ void ln_rows_to_smem(const ChainParams& p, int m0, uint8_t* sX, __nv_bfloat16* xn_b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kV = DP / 128 > 0 ? DP / 128 : 1;   // float4 per lane per row
  constexpr int kB = DP == 256 ? 2 : 4;             // rows in flight
  constexpr int kW = kThreads / 32;
  float4 gw[kV], gb[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int c = (i * 32 + lane) * 4;
    gw[i] = c < p.D ? *reinterpret_cast<const float4*>(p.ln_w + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    gb[i] = c < p.D ? *reinterpret_cast<const float4*>(p.ln_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r0 = warp * kB; r0 < kRows; r0 += kW * kB) {
    float4 v[kB][kV];
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      const int row = m0 + r0 + b;
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        const int c = (i * 32 + lane) * 4;
        v[b][i] = (r0 + b < kRows && row < p.M && c < p.D)
                      ? *reinterpret_cast<const float4*>(p.u + static_cast<long long>(row) * p.D + c)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      const int r = r0 + b, row = m0 + r;
      if (r < kRows) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kV; ++i) s += v[b][i].x + v[b][i].y + v[b][i].z + v[b][i].w;
        const float mean = warp_sum(s) / p.D;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < p.D) {
            const float a = v[b][i].x - mean, bb = v[b][i].y - mean, cc = v[b][i].z - mean, d = v[b][i].w - mean;
            ss += a * a + bb * bb + cc * cc + d * d;
          }
        }
        const float rstd = rsqrtf(warp_sum(ss) / p.D + kLnEps);
#pragma unroll
        for (int i = 0; i < kV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < DP) {
            uint2 o = make_uint2(0u, 0u);
            if (row < p.M && c < p.D) {
              o.x = pack_bf16((v[b][i].x - mean) * rstd * gw[i].x + gb[i].x, (v[b][i].y - mean) * rstd * gw[i].y + gb[i].y);
              o.y = pack_bf16((v[b][i].z - mean) * rstd * gw[i].z + gb[i].z, (v[b][i].w - mean) * rstd * gw[i].w + gb[i].w);
              if (xn_b) *reinterpret_cast<uint2*>(xn_b + static_cast<long long>(row) * p.D + c) = o;
            }
            // panel = c/64, 16-byte chunk = (c%64)/8, 8 bytes at (c%8)*2
            *reinterpret_cast<uint2*>(sX + (c >> 6) * (kRows * 128) + sw128_offset(r, (c & 63) >> 3) + (c & 7) * 2) = o;
          }
        }
      }
    }
  }
}

// Plain fp32 rows (optionally masked by the output-site dropout) -> swizzled bf16 A operand (+ bf16 copy to HBM).
template <int DP, bool kDrop>
__device__ __forceinline__ void rows_to_smem(const float* src, int M, int D, int m0, uint8_t* sA, __nv_bfloat16* dst_b,
                                             const Drop& drop) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kThreads / 32) {
    const int row = m0 + r;
    for (int c = lane * 4; c < DP; c += 128) {
      uint2 o = make_uint2(0u, 0u);
      if (row < M && c < D) {
        float4 v = *reinterpret_cast<const float4*>(src + static_cast<long long>(row) * D + c);
        if (kDrop) {   // gradient of the dropped branch output: dY * mask * scale
          const unsigned long long i0 = static_cast<unsigned long long>(row) * D + c;
          drop_apply2(drop, v.x, v.y, i0);
          drop_apply2(drop, v.z, v.w, i0 + 2);
        }
        o.x = pack_bf16(v.x, v.y);
        o.y = pack_bf16(v.z, v.w);
        if (dst_b) *reinterpret_cast<uint2*>(dst_b + static_cast<long long>(row) * D + c) = o;
      }
      *reinterpret_cast<uint2*>(sA + (c >> 6) * (kRows * 128) + sw128_offset(r, (c & 63) >> 3) + (c & 7) * 2) = o;
    }
  }
}

template <int DP>
__device__ __forceinline__ void load_w1(uint8_t* slot, const CUtensorMap* tmW1, uint64_t* bar, int c0) {
  mbar_arrive_expect_tx(bar, Cfg<DP>::kW1Bytes);
#pragma unroll
  for (int pnl = 0; pnl < Cfg<DP>::kPanels; ++pnl)          // [64 c-rows][64 d] panels
    tma_load_2d(slot + pnl * (kCc * 128), tmW1, bar, pnl * 64, c0);
}
template <int DP>
__device__ __forceinline__ void load_w2(uint8_t* slot, const CUtensorMap* tmW2, uint64_t* bar, int c0) {
  mbar_arrive_expect_tx(bar, Cfg<DP>::kW2Bytes);
  tma_load_2d(slot, tmW2, bar, c0, 0);                       // [DP d-rows][64 c]
}

__device__ __forceinline__ void load_bias8(const ChainParams& p, const float* sBias, int c, float (&b)[8]) {
  if (p.bias_smem) {
    const float4 b0 = *reinterpret_cast<const float4*>(sBias + c);
    const float4 b1v = *reinterpret_cast<const float4*>(sBias + c + 4);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1v.x; b[5] = b1v.y; b[6] = b1v.z; b[7] = b1v.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) b[e] = (c + e < p.C) ? __ldg(p.b1 + c + e) : 0.f;
  }
}

// ============================================================================================ forward
template <int DP, bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
chain_fwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const ChainParams p) {
  using C = Cfg<DP>;
  constexpr int S1 = C::kFwdS1, S2 = C::kFwdS2, NB = C::kFwdNB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sW1 = sX + C::kXBytes;
  uint8_t* sW2 = sW1 + S1 * C::kW1Bytes;
  uint8_t* sG = sW2 + S2 * C::kW2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 2 * kGBytes);
  uint64_t* w1full = bars;            // [S1]  TMA -> MMA
  uint64_t* w1empty = w1full + S1;    // [S1]  GEMM1 done -> TMA
  uint64_t* w2full = w1empty + S1;    // [S2]
  uint64_t* w2empty = w2full + S2;    // [S2]  GEMM2 done -> TMA
  uint64_t* hfull = w2empty + S2;     // [NB]  GEMM1 done -> epilogue
  uint64_t* hempty = hfull + NB;      // [NB]  epilogue has read Hacc -> MMA
  uint64_t* gfull = hempty + NB;      // [2]   epilogue wrote sG -> MMA
  uint64_t* gempty = gfull + 2;       // [2]   GEMM2 done reading sG -> epilogue
  uint64_t* yfull = gempty + 2;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yfull + 1);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kRows;
  const int nch = ceil_div(p.C, kCc);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S1; ++i) { mbar_init(&w1full[i], 1); mbar_init(&w1empty[i], 1); }
    for (int i = 0; i < S2; ++i) { mbar_init(&w2full[i], 1); mbar_init(&w2empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(&gfull[i], 128); mbar_init(&gempty[i], 1); }
    mbar_init(yfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kFwdTmem);
  if (p.bias_smem)
    for (int i = threadIdx.x; i < nch * kCc; i += kThreads) sBias[i] = i < p.C ? p.b1[i] : 0.f;
  __syncthreads();   // barriers initialised before the producer's early prefetch below

  // The weight rings do not depend on the activations: start filling them before the LayerNorm prologue.
  if (warp == 0 && lane == 0) {
    for (int j = 0; j < (nch < S1 ? nch : S1); ++j) load_w1<DP>(sW1 + j * C::kW1Bytes, &tmW1, &w1full[j], j * kCc);
    for (int j = 0; j < (nch < S2 ? nch : S2); ++j) load_w2<DP>(sW2 + j * C::kW2Bytes, &tmW2, &w2full[j], j * kCc);
  }
  ln_rows_to_smem<DP>(p, m0, sX, nullptr);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tY = tmem_base + C::kFwdYCol;

  if (warp == 0) {
    if (lane == 0) {
      // Refill both rings in the order the MMA issuer frees the slots: the prologue GEMM1s free W1 slots first, then
      // iteration `it` of the issuer completes GEMM2(it) (frees a W2 slot) and GEMM1(it + NB) (frees a W1 slot).
      auto refill_w1 = [&](int x) {
        mbar_wait(&w1empty[x % S1], ((x / S1) & 1) ^ 1);
        load_w1<DP>(sW1 + (x % S1) * C::kW1Bytes, &tmW1, &w1full[x % S1], x * kCc);
      };
      auto refill_w2 = [&](int y) {
        mbar_wait(&w2empty[y % S2], ((y / S2) & 1) ^ 1);
        load_w2<DP>(sW2 + (y % S2) * C::kW2Bytes, &tmW2, &w2full[y % S2], y * kCc);
      };
      for (int x = S1; x < S1 + NB && x < nch; ++x) refill_w1(x);
      for (int it = 0; it < nch; ++it) {
        if (it + S2 < nch) refill_w2(it + S2);
        if (it + S1 + NB < nch) refill_w1(it + S1 + NB);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(kRows, kCc, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(kRows, DP, 0, 0);
      const uint32_t x_addr = smem_u32(sX);
      auto gemm1 = [&](int j) {   // Hacc[j % NB] = sX . W1_j^T
        const int s = j % S1, hb = j % NB;
        mbar_wait(&w1full[s], (j / S1) & 1);
        mbar_wait(&hempty[hb], ((j / NB) & 1) ^ 1);
        tc_fence_after();
        const uint32_t w1_addr = smem_u32(sW1 + s * C::kW1Bytes);
        const uint32_t tH = tmem_base + hb * kCc;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tH, umma_desc_sw128(x_addr + (kk >> 2) * (kRows * 128) + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(w1_addr + (kk >> 2) * (kCc * 128) + (kk & 3) * 32, 16, 1024), idesc1, kk > 0 ? 1u : 0u);
        umma_commit(&w1empty[s]);
        umma_commit(&hfull[hb]);
      };
      for (int j = 0; j < (nch < NB ? nch : NB); ++j) gemm1(j);
      for (int j = 0; j < nch; ++j) {   // Yacc += sG[j&1] . W2_j^T, then run GEMM1 NB chunks ahead
        const int s = j % S2;
        mbar_wait(&w2full[s], (j / S2) & 1);
        mbar_wait(&gfull[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t g_addr = smem_u32(sG + (j & 1) * kGBytes);
        const uint32_t w2_addr = smem_u32(sW2 + s * C::kW2Bytes);
#pragma unroll
        for (int kk = 0; kk < kCc / 16; ++kk)
          umma_bf16(tY, umma_desc_sw128(g_addr + kk * 32, 16, 1024), umma_desc_sw128(w2_addr + kk * 32, 16, 1024), idesc2,
                    (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&w2empty[s]);
        umma_commit(&gempty[j & 1]);
        if (j + NB < nch) gemm1(j + NB);
      }
      umma_commit(yfull);
    }
  } else {
    const int q = warp & 3;                // TMEM lane quadrant this warp may access (warp id % 4)
    const int grp = (warp - 2) >> 2;       // epilogue group: chunks j = grp (mod 2), staging buffer sG[grp]
    const int r = q * 32 + lane;           // row inside the tile == TMEM lane
    const int row = m0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* g = sG + grp * kGBytes;
    for (int j = grp; j < nch; j += 2) {
      const int hb = j % NB;
      mbar_wait(&hfull[hb], (j / NB) & 1);
      tc_fence_after();
      uint32_t h[64];
      {
        uint32_t (&h0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&h[0]);
        uint32_t (&h1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&h[32]);
        tmem_ld32(tmem_base + lane_addr + hb * kCc, h0);
        tmem_ld32(tmem_base + lane_addr + hb * kCc + 32, h1);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&hempty[hb]);
      mbar_wait(&gempty[grp], ((j >> 1) & 1) ^ 1);
      const int c0 = j * kCc;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float b[8];
        load_bias8(p, sBias, c0 + ch * 8, b);
        float2 v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          v[e] = gelu2(__fadd2_rn(make_float2(__uint_as_float(h[ch * 8 + 2 * e]), __uint_as_float(h[ch * 8 + 2 * e + 1])),
                                  make_float2(b[2 * e], b[2 * e + 1])));
        if (kDrop) {
          const unsigned long long i0 = static_cast<unsigned long long>(row) * p.ldh + c0 + ch * 8;
#pragma unroll
          for (int e = 0; e < 4; ++e) drop_apply2(p.dh, v[e].x, v[e].y, i0 + 2 * e);
        }
        *reinterpret_cast<uint4*>(g + sw128_offset(r, ch)) =
            make_uint4(pack_bf16(v[0].x, v[0].y), pack_bf16(v[1].x, v[1].y), pack_bf16(v[2].x, v[2].y), pack_bf16(v[3].x, v[3].y));
      }
      fence_proxy_async();
      mbar_arrive(&gfull[grp]);
    }
    // final: y = u + Drop(Yacc + b2); each group drains half of the columns
    mbar_wait(yfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int d0 = grp * (DP / 2); d0 < (grp + 1) * (DP / 2); d0 += 32) {
      uint32_t a[32];
      tmem_ld32(tY + lane_addr + d0, a);
      tmem_ld_wait();
      if (row < p.M && d0 < p.D) {
        const float* urow = p.u + static_cast<long long>(row) * p.D + d0;
        float* yrow = p.y + static_cast<long long>(row) * p.D + d0;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          if (d0 + e < p.D) {
            const float4 uu = *reinterpret_cast<const float4*>(urow + e);
            const float4 bb = *reinterpret_cast<const float4*>(p.b2 + d0 + e);
            float4 o;
            o.x = bb.x + __uint_as_float(a[e]);
            o.y = bb.y + __uint_as_float(a[e + 1]);
            o.z = bb.z + __uint_as_float(a[e + 2]);
            o.w = bb.w + __uint_as_float(a[e + 3]);
            if (kDrop) {
              const unsigned long long i0 = static_cast<unsigned long long>(row) * p.D + d0 + e;
              drop_apply2(p.dout, o.x, o.y, i0);
              drop_apply2(p.dout, o.z, o.w, i0 + 2);
            }
            o.x += uu.x; o.y += uu.y; o.z += uu.z; o.w += uu.w;
            *reinterpret_cast<float4*>(yrow + e) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kFwdTmem);
}

// ============================================================================================ backward A
template <int DP, bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
chain_bwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const ChainParams p) {
  using C = Cfg<DP>;
  constexpr int S1 = C::kBwdS1, S2 = C::kBwdS2, NB = C::kBwdNB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sdY = sX + C::kXBytes;
  uint8_t* sW1 = sdY + C::kXBytes;
  uint8_t* sW2 = sW1 + S1 * C::kW1Bytes;
  uint8_t* sdH = sW2 + S2 * C::kW2Bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdH + 2 * kGBytes);
  uint64_t* w1full = bars;            // [S1]
  uint64_t* w1empty = w1full + S1;    // [S1]  dXn GEMM done with the W1 chunk -> TMA
  uint64_t* w2full = w1empty + S1;    // [S2]
  uint64_t* w2empty = w2full + S2;    // [S2]  dG GEMM done -> TMA
  uint64_t* hfull = w2empty + S2;     // [NB]  H and dG accumulators of chunk j ready
  uint64_t* hempty = hfull + NB;      // [NB]
  uint64_t* gfull = hempty + NB;      // [2]   sdH written
  uint64_t* gempty = gfull + 2;       // [2]   dXn GEMM done with sdH
  uint64_t* yfull = gempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yfull + 1);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kRows;
  const int nch = ceil_div(p.C, kCc);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S1; ++i) { mbar_init(&w1full[i], 1); mbar_init(&w1empty[i], 1); }
    for (int i = 0; i < S2; ++i) { mbar_init(&w2full[i], 1); mbar_init(&w2empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(&gfull[i], 128); mbar_init(&gempty[i], 1); }
    mbar_init(yfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kBwdTmem);
  if (p.bias_smem)
    for (int i = threadIdx.x; i < nch * kCc; i += kThreads) sBias[i] = i < p.C ? p.b1[i] : 0.f;
  __syncthreads();
  if (warp == 0 && lane == 0) {
    for (int j = 0; j < (nch < S1 ? nch : S1); ++j) load_w1<DP>(sW1 + j * C::kW1Bytes, &tmW1, &w1full[j], j * kCc);
    for (int j = 0; j < (nch < S2 ? nch : S2); ++j) load_w2<DP>(sW2 + j * C::kW2Bytes, &tmW2, &w2full[j], j * kCc);
  }
  ln_rows_to_smem<DP>(p, m0, sX, p.xn_b);
  rows_to_smem<DP, kDrop>(p.dy, p.M, p.D, m0, sdY, p.dy_b, p.dout);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tDX = tmem_base + C::kBwdXCol;

  if (warp == 0) {
    if (lane == 0) {
      // Slot release order of the MMA issuer: the prologue H/dG GEMMs free W2 slots, then iteration `it` completes the
      // dXn GEMM of chunk it (frees its W1 slot) and the H/dG GEMMs of chunk it + NB (frees a W2 slot).
      auto refill_w1 = [&](int x) {
        mbar_wait(&w1empty[x % S1], ((x / S1) & 1) ^ 1);
        load_w1<DP>(sW1 + (x % S1) * C::kW1Bytes, &tmW1, &w1full[x % S1], x * kCc);
      };
      auto refill_w2 = [&](int y) {
        mbar_wait(&w2empty[y % S2], ((y / S2) & 1) ^ 1);
        load_w2<DP>(sW2 + (y % S2) * C::kW2Bytes, &tmW2, &w2full[y % S2], y * kCc);
      };
      for (int y = S2; y < S2 + NB && y < nch; ++y) refill_w2(y);
      for (int it = 0; it < nch; ++it) {
        if (it + S1 < nch) refill_w1(it + S1);
        if (it + S2 + NB < nch) refill_w2(it + S2 + NB);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idescH = umma_idesc_bf16(kRows, kCc, 0, 0);    // A K-major, B (W1 chunk)  K-major
      constexpr uint32_t idescG = umma_idesc_bf16(kRows, kCc, 0, 1);    // A K-major, B (W2 tile)   MN-major
      constexpr uint32_t idescX = umma_idesc_bf16(kRows, DP, 0, 1);     // A K-major, B (W1 chunk)  MN-major
      const uint32_t x_addr = smem_u32(sX), dy_addr = smem_u32(sdY);
      auto gemm_hg = [&](int j) {   // H[j%NB] = sX . W1_j^T ; dG[j%NB] = sdY . W2_j
        const int s1 = j % S1, s2 = j % S2, hb = j % NB;
        mbar_wait(&w1full[s1], (j / S1) & 1);
        mbar_wait(&w2full[s2], (j / S2) & 1);
        mbar_wait(&hempty[hb], ((j / NB) & 1) ^ 1);
        tc_fence_after();
        const uint32_t w1_addr = smem_u32(sW1 + s1 * C::kW1Bytes);
        const uint32_t w2_addr = smem_u32(sW2 + s2 * C::kW2Bytes);
        const uint32_t tH = tmem_base + hb * kCc;
        const uint32_t tG = tmem_base + C::kBwdGCol + hb * kCc;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tH, umma_desc_sw128(x_addr + (kk >> 2) * (kRows * 128) + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(w1_addr + (kk >> 2) * (kCc * 128) + (kk & 3) * 32, 16, 1024), idescH, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)    // B = W2 tile [DP d-rows][64 c]: 16 d-rows per step = 2048 B
          umma_bf16(tG, umma_desc_sw128(dy_addr + (kk >> 2) * (kRows * 128) + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(w2_addr + kk * 2048, 8192, 1024), idescG, kk > 0 ? 1u : 0u);
        umma_commit(&w2empty[s2]);
        umma_commit(&hfull[hb]);
      };
      for (int j = 0; j < (nch < NB ? nch : NB); ++j) gemm_hg(j);
      for (int j = 0; j < nch; ++j) {   // dXn += sdH[j&1] . W1_j   (contraction over the 64 channels of the chunk)
        const int s1 = j % S1;
        mbar_wait(&gfull[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t dh_addr = smem_u32(sdH + (j & 1) * kGBytes);
        const uint32_t w1_addr = smem_u32(sW1 + s1 * C::kW1Bytes);
#pragma unroll
        for (int kk = 0; kk < kCc / 16; ++kk)   // B: 16 c-rows per step = 2048 B; d panels 8 KB apart (LBO)
          umma_bf16(tDX, umma_desc_sw128(dh_addr + kk * 32, 16, 1024), umma_desc_sw128(w1_addr + kk * 2048, kCc * 128, 1024),
                    idescX, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&w1empty[s1]);
        umma_commit(&gempty[j & 1]);
        if (j + NB < nch) gemm_hg(j + NB);
      }
      umma_commit(yfull);
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int row = m0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* sd = sdH + grp * kGBytes;
    for (int j = grp; j < nch; j += 2) {
      const int hb = j % NB;
      mbar_wait(&hfull[hb], (j / NB) & 1);
      tc_fence_after();
      const int c0 = j * kCc;
      bool waited = false;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t h[32], dg[32];
        tmem_ld32(tmem_base + lane_addr + hb * kCc + half * 32, h);
        tmem_ld32(tmem_base + lane_addr + C::kBwdGCol + hb * kCc + half * 32, dg);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&hempty[hb]);
        }
        if (!waited) {
          mbar_wait(&gempty[grp], ((j >> 1) & 1) ^ 1);
          waited = true;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cc = c0 + half * 32 + ch * 8;
          float bias[8];
          load_bias8(p, sBias, cc, bias);
          float2 gv[4], dv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 dgelu;
            gv[e] = gelu2_grad(__fadd2_rn(make_float2(__uint_as_float(h[ch * 8 + 2 * e]), __uint_as_float(h[ch * 8 + 2 * e + 1])),
                                          make_float2(bias[2 * e], bias[2 * e + 1])), dgelu);
            dv[e] = __fmul2_rn(make_float2(__uint_as_float(dg[ch * 8 + 2 * e]), __uint_as_float(dg[ch * 8 + 2 * e + 1])), dgelu);
          }
          if (kDrop) {   // G' = m*s*G ; dH = dG' * m*s*gelu'(h)
            const unsigned long long i0 = static_cast<unsigned long long>(row) * p.ldh + cc;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              drop_apply2(p.dh, gv[e].x, gv[e].y, i0 + 2 * e);
              drop_apply2(p.dh, dv[e].x, dv[e].y, i0 + 2 * e);
            }
          }
          const uint4 go = make_uint4(pack_bf16(gv[0].x, gv[0].y), pack_bf16(gv[1].x, gv[1].y), pack_bf16(gv[2].x, gv[2].y),
                                      pack_bf16(gv[3].x, gv[3].y));
          const uint4 dh = make_uint4(pack_bf16(dv[0].x, dv[0].y), pack_bf16(dv[1].x, dv[1].y), pack_bf16(dv[2].x, dv[2].y),
                                      pack_bf16(dv[3].x, dv[3].y));
          *reinterpret_cast<uint4*>(sd + sw128_offset(r, half * 4 + ch)) = dh;
          if (row < p.M && cc < p.ldh) {   // ldh is a multiple of 8 >= C: whole 16-byte chunks only
            *reinterpret_cast<uint4*>(p.g_b + static_cast<long long>(row) * p.ldh + cc) = go;
            *reinterpret_cast<uint4*>(p.dh_b + static_cast<long long>(row) * p.ldh + cc) = dh;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&gfull[grp]);
    }
    mbar_wait(yfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int d0 = grp * (DP / 2); d0 < (grp + 1) * (DP / 2); d0 += 32) {
      uint32_t a[32];
      tmem_ld32(tDX + lane_addr + d0, a);
      tmem_ld_wait();
      if (row < p.M && d0 < p.D) {
        float* o = p.dxn + static_cast<long long>(row) * p.D + d0;
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          if (d0 + e < p.D)
            *reinterpret_cast<float4*>(o + e) = make_float4(__uint_as_float(a[e]), __uint_as_float(a[e + 1]),
                                                            __uint_as_float(a[e + 2]), __uint_as_float(a[e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kBwdTmem);
}

template <int DP, bool kBwd, bool kDrop>
int launch_chain(const CUtensorMap& t1, const CUtensorMap& t2, const ChainParams& p, cudaStream_t s) {
  constexpr int base = kBwd ? Cfg<DP>::kBwdSmem : Cfg<DP>::kFwdSmem;
  constexpr int kMaxBias = kBwd ? Cfg<DP>::kMaxBiasBwd : Cfg<DP>::kMaxBiasFwd;
  const int bias_bytes = ceil_div(p.C, kCc) * kCc * 4;
  ChainParams pp = p;
  pp.bias_smem = bias_bytes <= kMaxBias ? 1 : 0;
  const int smem = base + (pp.bias_smem ? bias_bytes : 0);
  static int configured = 0;   // largest dynamic smem size opted into so far (idempotent attribute)
  if constexpr (kBwd) {
    if (smem > configured) {
      if (cudaFuncSetAttribute(chain_bwd_kernel<DP, kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return M2_ERR_LAUNCH;
      configured = smem;
    }
    LaunchScope scope("chain_bwd", s);
    chain_bwd_kernel<DP, kDrop><<<ceil_div(p.M, kRows), kThreads, smem, s>>>(t1, t2, pp);
  } else {
    if (smem > configured) {
      if (cudaFuncSetAttribute(chain_fwd_kernel<DP, kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return M2_ERR_LAUNCH;
      configured = smem;
    }
    LaunchScope scope("chain_fwd", s);
    chain_fwd_kernel<DP, kDrop><<<ceil_div(p.M, kRows), kThreads, smem, s>>>(t1, t2, pp);
  }
  M2_LAUNCH_CHECK();
  return M2_OK;
}

template <int DP, bool kBwd>
int launch_chain_d(const CUtensorMap& t1, const CUtensorMap& t2, const ChainParams& p, cudaStream_t s) {
  return (p.dh.thresh || p.dout.thresh) ? launch_chain<DP, kBwd, true>(t1, t2, p, s) : launch_chain<DP, kBwd, false>(t1, t2, p, s);
}

int make_weight_maps(CUtensorMap* t1, CUtensorMap* t2, const void* w1b, const void* w2b, int D, int C, int ldw2, int DP) {
  // W1 bf16 [C][D] (ld = D): box 64 c-rows x 64 d.   W2 bf16 [D][ldw2] (cols >= C zero): box DP d-rows x 64 c.
  int rc = make_tmap_bf16(t1, w1b, C, D, D, kCc, 64);
  if (rc) return rc;
  return make_tmap_bf16(t2, w2b, D, ldw2, ldw2, DP, kCc);
}

}  // namespace

int chain_generation() {
  static const int gen = []() {
    const char* e = getenv("M2B200_CHAIN_GEN");
    return e ? atoi(e) : 4;
  }();
  return gen;
}

int dh_l2_hint() {
  static const int on = []() {
    const char* e = getenv("M2B200_DH_L2HINT");
    return e ? atoi(e) : 1;
  }();
  return on;
}

// Which hidden sizes the fused chains cover (others take the unfused GEMM path in abi.cu).
bool chain_fwd_supported(int D) { return D >= 16 && D <= 256 && D % 8 == 0; }
bool chain_bwd_supported(int D) { return D >= 16 && D <= 128 && D % 8 == 0; }

int chain_fwd(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
              int ldw2, const float* b2, float* y, int M, int D, int C, int exact_gelu, float drop_p, unsigned long long seed,
              cudaStream_t s) {
  if (!chain_fwd_supported(D) || ldw2 % 8 || ldw2 < C || exact_gelu) return M2_ERR_ARG;
  if (chain_generation() != 1 && chain_fwd_ts_supported(D))
    return chain_fwd_ts(u, ln_w, ln_b, w1b, b1, w2b, ldw2, b2, y, M, D, C, drop_p, seed, s);
  const int DP = D <= 64 ? 64 : (D <= 128 ? 128 : 256);
  CUtensorMap t1, t2;
  int rc = make_weight_maps(&t1, &t2, w1b, w2b, D, C, ldw2, DP);
  if (rc) return rc;
  ChainParams p = {};
  p.u = u; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2; p.y = y;
  p.M = M; p.D = D; p.C = C; p.ldh = (C + 7) & ~7;
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden); p.dout = make_drop(drop_p, seed, kSiteChannelOut);
  if (DP == 64) return launch_chain_d<64, false>(t1, t2, p, s);
  if (DP == 128) return launch_chain_d<128, false>(t1, t2, p, s);
  return launch_chain_d<256, false>(t1, t2, p, s);
}

int chain_bwd(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
              int ldw2, const float* dy, void* xn_b, void* dy_b, void* g_b, void* dh_b, int ldh, float* dxn, int M, int D,
              int C, int exact_gelu, float drop_p, unsigned long long seed, cudaStream_t s) {
  if (!chain_bwd_supported(D) || ldw2 % 8 || ldw2 < C || ldh % 8 || ldh < C || exact_gelu) return M2_ERR_ARG;
  const int DP = D <= 64 ? 64 : 128;
  CUtensorMap t1, t2;
  int rc = make_weight_maps(&t1, &t2, w1b, w2b, D, C, ldw2, DP);
  if (rc) return rc;
  ChainParams p = {};
  p.u = u; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.dy = dy;
  p.xn_b = static_cast<__nv_bfloat16*>(xn_b); p.dy_b = static_cast<__nv_bfloat16*>(dy_b);
  p.g_b = static_cast<__nv_bfloat16*>(g_b); p.dh_b = static_cast<__nv_bfloat16*>(dh_b);
  p.dxn = dxn; p.M = M; p.D = D; p.C = C; p.ldh = ldh;
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden); p.dout = make_drop(drop_p, seed, kSiteChannelOut);
  if (DP == 64) return launch_chain_d<64, true>(t1, t2, p, s);
  return launch_chain_d<128, true>(t1, t2, p, s);
}

}  // namespace m2
