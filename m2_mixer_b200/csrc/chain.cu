// Fused channel-mixing chains of a Mixer block on tcgen05 / TMEM / TMA (bf16 operands, fp32 accumulate).
//
// Reference arithmetic: MixerBlock.channel_mix, modules/mixer.py:37-40,45
//     y = u + W2 . GELU(W1 . LN(u) + b1) + b2        per token row (M = B*N rows, D hidden, C channel_dim)
//
// FORWARD (chain_fwd_kernel), one CTA per 128-row token tile, the [128 x C] hidden activation never leaves the SM:
//   phase 0  all warps: LayerNorm the fp32 rows, round to bf16, store as the K-major SW128 A operand (sX)
//   loop over C in chunks of 64 channels (weights arrive by TMA into a ring, W1 chunk [64 x D], W2 chunk [D x 64]):
//     MMA warp : GEMM1(j)  Hacc[j%2][128x64]  = sX . W1_j^T           (TMEM, double buffered)
//                GEMM2(j-1) Yacc[128xD]      += sG[(j-1)%2] . W2_{j-1}^T
//     epilogue : Hacc -> regs, +b1, GELU, bf16 -> sG[j%2] (swizzled A operand of GEMM2)
//   final      : Yacc -> regs, +b2, +u (residual), fp32 store
//
// BACKWARD part A (chain_bwd_kernel), same tiling, given dY:
//   phase 0  : LN(u) -> sX ; dY -> bf16 sdY ; both also written to HBM as bf16 (operands of the wgrad GEMMs)
//   per chunk: H   = sX  . W1_j^T     (recompute)          -> TMEM
//              dG  = sdY . W2_j       (W2 tile consumed MN-major: no transposed weight copy)
//              epilogue: G = GELU(H+b1), dH = dG * GELU'(H+b1); G,dH -> HBM (bf16, for dW2/dW1), dH -> smem
//              dXn += sdH . W1_j      (W1 tile consumed MN-major)  -> TMEM accumulator
//   final    : dXn -> HBM fp32 (LayerNorm backward + residual is a separate bandwidth kernel)
//   The weight gradients dW2 = dY^T.G and dW1 = dH^T.LN(u) are token-axis contractions done by the generic
//   tcgen05 GEMM (umma_gemm.cu) with both operands MN-major.
//
// D is padded in shared memory to DP in {64,128,256}; C is arbitrary (TMA zero-fills the ragged last chunk, the
// bf16 weight copies are padded to a multiple of 8 columns so their row stride is 16-byte aligned).
#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"

namespace m2 {
namespace {

constexpr int kRows = 128;      // token rows per CTA (UMMA M)
constexpr int kCc = 64;         // channels per chunk
constexpr int kThreads = 320;   // warp0 TMA, warp1 MMA, warps 2-9 epilogue: two groups of 4 warps, group g owns chunks j = g (mod 2)
constexpr int kMaxBiasSmemFwd = 32 * 1024, kMaxBiasSmemBwd = 24 * 1024;
constexpr int kGBytes = kRows * kCc * 2;   // one [128 x 64] bf16 tile

template <int DP>
struct Cfg {
  static constexpr int kPanels = DP / 64;
  static constexpr int kXBytes = kRows * DP * 2;
  static constexpr int kW1Bytes = kCc * DP * 2;
  static constexpr int kW2Bytes = DP * kCc * 2;
  static constexpr int kStageBytes = kW1Bytes + kW2Bytes;
  static constexpr int kFwdStages = DP == 256 ? 2 : (DP == 128 ? 3 : 4);
  static constexpr int kBwdStages = DP == 128 ? 3 : 4;
  // + barriers (256 B) + 1024 B alignment slack; the b1 vector is staged behind it when it fits (DP <= 128)
  static constexpr int kFwdSmem = kXBytes + kFwdStages * kStageBytes + 2 * kGBytes + 256 + 1024;
  static constexpr int kBwdSmem = 2 * kXBytes + kBwdStages * kStageBytes + 2 * kGBytes + 256 + 1024;
  static constexpr int kFwdTmem = (128 + DP) <= 256 ? 256 : 512;          // 2 x 64 (H) + DP (Y)
  static constexpr int kBwdTmem = 512;                                   // 2 x 64 (H) + 2 x 64 (dG) + DP (dXn)
};

struct ChainParams {
  const float* u;        // [M][D] block input (pre-LN residual stream)
  const float* ln_w; const float* ln_b;
  const float* b1;       // [C]
  const float* b2;       // [D]
  float* y;              // fwd: [M][D]
  const float* dy;       // bwd: [M][D]
  __nv_bfloat16* xn_b;   // bwd out: LN(u) bf16 [M][D]
  __nv_bfloat16* dy_b;   // bwd out: dY bf16    [M][D]
  __nv_bfloat16* g_b;    // bwd out: G  bf16    [M][ldh]
  __nv_bfloat16* dh_b;   // bwd out: dH bf16    [M][ldh]
  float* dxn;            // bwd out: dL/dLN(u) fp32 [M][D]
  int M, D, C, ldh;
  int exact_gelu;
  int bias_smem;         // b1 (zero padded to a multiple of 64) is staged in shared memory
  Drop dh, dout;         // dropout after GELU (index row*ldh + c) and after the second Linear (index row*D + d)
};

// LayerNorm `rows` of the tile into the swizzled bf16 A operand; optional bf16 copy to HBM.
template <int DP>
__device__ __forceinline__ void ln_rows_to_smem(const ChainParams& p, int m0, uint8_t* sX, __nv_bfloat16* xn_b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kV = DP / 128 > 0 ? DP / 128 : 1;   // float4 per lane
  for (int r = warp; r < kRows; r += kThreads / 32) {
    const int row = m0 + r;
    float4 v[kV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = (row < p.M && c < p.D) ? *reinterpret_cast<const float4*>(p.u + static_cast<long long>(row) * p.D + c)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / p.D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < p.D) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        ss += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / p.D + kLnEps);
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c >= DP) continue;
      uint2 o = make_uint2(0u, 0u);
      if (row < p.M && c < p.D) {
        const float4 w = *reinterpret_cast<const float4*>(p.ln_w + c);
        const float4 bb = *reinterpret_cast<const float4*>(p.ln_b + c);
        o.x = pack_bf16((v[i].x - mean) * rstd * w.x + bb.x, (v[i].y - mean) * rstd * w.y + bb.y);
        o.y = pack_bf16((v[i].z - mean) * rstd * w.z + bb.z, (v[i].w - mean) * rstd * w.w + bb.w);
        if (xn_b) *reinterpret_cast<uint2*>(xn_b + static_cast<long long>(row) * p.D + c) = o;
      }
      // panel = c/64, 16-byte chunk = (c%64)/8, 8 bytes at (c%8)*2
      *reinterpret_cast<uint2*>(sX + (c >> 6) * (kRows * 128) + sw128_offset(r, (c & 63) >> 3) + (c & 7) * 2) = o;
    }
  }
}

// Plain fp32 rows -> swizzled bf16 A operand (+ bf16 copy to HBM).
template <int DP>
__device__ __forceinline__ void rows_to_smem(const float* src, int M, int D, int m0, uint8_t* sA, __nv_bfloat16* dst_b,
                                             const Drop& drop) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kRows; r += kThreads / 32) {
    const int row = m0 + r;
    for (int c = lane * 4; c < DP; c += 128) {
      uint2 o = make_uint2(0u, 0u);
      if (row < M && c < D) {
        float4 v = *reinterpret_cast<const float4*>(src + static_cast<long long>(row) * D + c);
        if (drop.thresh) {   // gradient of the dropped branch output: dY * mask * scale
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
__device__ __forceinline__ void load_weight_stage(uint8_t* stage, const CUtensorMap* tmW1, const CUtensorMap* tmW2,
                                                  uint64_t* bar, int c0) {
  mbar_arrive_expect_tx(bar, Cfg<DP>::kStageBytes);
#pragma unroll
  for (int pnl = 0; pnl < Cfg<DP>::kPanels; ++pnl)          // W1 chunk: [64 c-rows][64 d] panels
    tma_load_2d(stage + pnl * (kCc * 128), tmW1, bar, pnl * 64, c0);
  tma_load_2d(stage + Cfg<DP>::kW1Bytes, tmW2, bar, c0, 0);  // W2 chunk: [DP d-rows][64 c]
}

// ============================================================================================ forward
template <int DP>
__global__ void __launch_bounds__(kThreads, 1)
chain_fwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const ChainParams p) {
  using C = Cfg<DP>;
  constexpr int S = C::kFwdStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sW = sX + C::kXBytes;
  uint8_t* sG = sW + S * C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 2 * kGBytes);
  uint64_t* full = bars;            // [S]   TMA -> MMA
  uint64_t* empty = full + S;       // [S]   MMA -> TMA
  uint64_t* hfull = empty + S;      // [2]   GEMM1 done   -> epilogue
  uint64_t* hempty = hfull + 2;     // [2]   epilogue read Hacc -> MMA
  uint64_t* gfull = hempty + 2;     // [2]   epilogue wrote sG  -> MMA
  uint64_t* gempty = gfull + 2;     // [2]   GEMM2 done reading sG -> epilogue
  uint64_t* yfull = gempty + 2;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yfull + 1);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kRows;
  const int nchunks = ceil_div(p.C, kCc);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 128);
      mbar_init(&gfull[i], 128); mbar_init(&gempty[i], 1);
    }
    mbar_init(yfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kFwdTmem);
  if (p.bias_smem)
    for (int i = threadIdx.x; i < nchunks * kCc; i += kThreads) sBias[i] = i < p.C ? p.b1[i] : 0.f;
  __syncthreads();   // barriers initialised before the producer's early prefetch below

  // The weight ring does not depend on the activations: start filling it before the LayerNorm prologue.
  if (warp == 0 && lane == 0) {
    const int pre = nchunks < S ? nchunks : S;
    for (int j = 0; j < pre; ++j) load_weight_stage<DP>(sW + j * C::kStageBytes, &tmW1, &tmW2, &full[j], j * kCc);
  }
  ln_rows_to_smem<DP>(p, m0, sX, nullptr);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tY = tmem_base + 128;   // columns [128, 128+DP)

  if (warp == 0) {
    if (lane == 0) {
      for (int j = S; j < nchunks; ++j) {
        const int s = j % S;
        mbar_wait(&empty[s], ((j / S) & 1) ^ 1);
        load_weight_stage<DP>(sW + s * C::kStageBytes, &tmW1, &tmW2, &full[s], j * kCc);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(kRows, kCc, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(kRows, DP, 0, 0);
      const uint32_t x_addr = smem_u32(sX);
      auto gemm2 = [&](int j) {   // Yacc += sG[j%2] . W2_j^T
        const int s = j % S;
        mbar_wait(&gfull[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t g_addr = smem_u32(sG + (j & 1) * kGBytes);
        const uint32_t w2_addr = smem_u32(sW + s * C::kStageBytes + C::kW1Bytes);
#pragma unroll
        for (int kk = 0; kk < kCc / 16; ++kk)
          umma_bf16(tY, umma_desc_sw128(g_addr + kk * 32, 16, 1024), umma_desc_sw128(w2_addr + kk * 32, 16, 1024), idesc2,
                    (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);        // weight stage free
        umma_commit(&gempty[j & 1]);   // sG buffer free
      };
      for (int j = 0; j < nchunks; ++j) {
        const int s = j % S;
        mbar_wait(&full[s], (j / S) & 1);
        mbar_wait(&hempty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t w1_addr = smem_u32(sW + s * C::kStageBytes);
        const uint32_t tH = tmem_base + (j & 1) * kCc;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tH, umma_desc_sw128(x_addr + (kk >> 2) * (kRows * 128) + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(w1_addr + (kk >> 2) * (kCc * 128) + (kk & 3) * 32, 16, 1024), idesc1, kk > 0 ? 1u : 0u);
        umma_commit(&hfull[j & 1]);
        if (j > 0) gemm2(j - 1);
      }
      gemm2(nchunks - 1);
      umma_commit(yfull);
    }
  } else {
    const int q = warp & 3;                // TMEM lane quadrant this warp may access (warp id % 4)
    const int grp = (warp - 2) >> 2;       // epilogue group: chunks j = grp (mod 2), buffers Hacc[grp] / sG[grp]
    const int r = q * 32 + lane;           // row inside the tile == TMEM lane
    const int row = m0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    for (int j = grp; j < nchunks; j += 2) {
      mbar_wait(&hfull[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t h[64];
      {
        uint32_t (&h0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&h[0]);
        uint32_t (&h1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&h[32]);
        tmem_ld32(tmem_base + lane_addr + (j & 1) * kCc, h0);
        tmem_ld32(tmem_base + lane_addr + (j & 1) * kCc + 32, h1);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&hempty[j & 1]);
      mbar_wait(&gempty[j & 1], ((j >> 1) & 1) ^ 1);
      uint8_t* g = sG + (j & 1) * kGBytes;
      const int c0 = j * kCc;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float b[8];
        if (p.bias_smem) {
          const float4 b0 = *reinterpret_cast<const float4*>(sBias + c0 + ch * 8);
          const float4 b1v = *reinterpret_cast<const float4*>(sBias + c0 + ch * 8 + 4);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1v.x; b[5] = b1v.y; b[6] = b1v.z; b[7] = b1v.w;
        } else if (c0 + ch * 8 + 8 <= p.C) {
          const float4 b0 = *reinterpret_cast<const float4*>(p.b1 + c0 + ch * 8);
          const float4 b1v = *reinterpret_cast<const float4*>(p.b1 + c0 + ch * 8 + 4);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1v.x; b[5] = b1v.y; b[6] = b1v.z; b[7] = b1v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) b[e] = (c0 + ch * 8 + e < p.C) ? p.b1[c0 + ch * 8 + e] : 0.f;
        }
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = __uint_as_float(h[ch * 8 + e]) + b[e];
          v[e] = p.exact_gelu ? gelu_erf(x) : gelu_fast(x);
        }
        if (p.dh.thresh) {
          const unsigned long long i0 = static_cast<unsigned long long>(row) * p.ldh + c0 + ch * 8;
#pragma unroll
          for (int e = 0; e < 8; e += 2) drop_apply2(p.dh, v[e], v[e + 1], i0 + e);
        }
        *reinterpret_cast<uint4*>(g + sw128_offset(r, ch)) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
      fence_proxy_async();
      mbar_arrive(&gfull[j & 1]);
    }
    // final: y = u + Yacc + b2
    mbar_wait(yfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int d0 = grp * (DP / 2); d0 < (grp + 1) * (DP / 2); d0 += 32) {   // each group drains half of the columns
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
            if (p.dout.thresh) {
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
template <int DP>
__global__ void __launch_bounds__(kThreads, 1)
chain_bwd_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const ChainParams p) {
  using C = Cfg<DP>;
  constexpr int S = C::kBwdStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sdY = sX + C::kXBytes;
  uint8_t* sW = sdY + C::kXBytes;
  uint8_t* sdH = sW + S * C::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdH + 2 * kGBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + S;
  uint64_t* hfull = empty + S;      // [2]  H and dG accumulators of chunk j are ready
  uint64_t* hempty = hfull + 2;     // [2]
  uint64_t* gfull = hempty + 2;     // [2]  sdH written
  uint64_t* gempty = gfull + 2;     // [2]  dXn GEMM done with sdH
  uint64_t* yfull = gempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yfull + 1);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kRows;
  const int nchunks = ceil_div(p.C, kCc);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 128);
      mbar_init(&gfull[i], 128); mbar_init(&gempty[i], 1);
    }
    mbar_init(yfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kBwdTmem);
  if (p.bias_smem)
    for (int i = threadIdx.x; i < nchunks * kCc; i += kThreads) sBias[i] = i < p.C ? p.b1[i] : 0.f;
  __syncthreads();
  if (warp == 0 && lane == 0) {
    const int pre = nchunks < S ? nchunks : S;
    for (int j = 0; j < pre; ++j) load_weight_stage<DP>(sW + j * C::kStageBytes, &tmW1, &tmW2, &full[j], j * kCc);
  }
  ln_rows_to_smem<DP>(p, m0, sX, p.xn_b);
  rows_to_smem<DP>(p.dy, p.M, p.D, m0, sdY, p.dy_b, p.dout);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: H[2] at 0/64, dG[2] at 128/192, dXn at 256..256+DP
  const uint32_t tDX = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      for (int j = S; j < nchunks; ++j) {
        const int s = j % S;
        mbar_wait(&empty[s], ((j / S) & 1) ^ 1);
        load_weight_stage<DP>(sW + s * C::kStageBytes, &tmW1, &tmW2, &full[s], j * kCc);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idescH = umma_idesc_bf16(kRows, kCc, 0, 0);    // A K-major, B (W1 chunk)  K-major
      constexpr uint32_t idescG = umma_idesc_bf16(kRows, kCc, 0, 1);    // A K-major, B (W2 tile)   MN-major
      constexpr uint32_t idescX = umma_idesc_bf16(kRows, DP, 0, 1);     // A K-major, B (W1 chunk)  MN-major
      const uint32_t x_addr = smem_u32(sX), dy_addr = smem_u32(sdY);
      auto gemm_dx = [&](int j) {   // dXn += sdH[j%2] . W1_j   (contraction over the 64 channels of the chunk)
        const int s = j % S;
        mbar_wait(&gfull[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t dh_addr = smem_u32(sdH + (j & 1) * kGBytes);
        const uint32_t w1_addr = smem_u32(sW + s * C::kStageBytes);
#pragma unroll
        for (int kk = 0; kk < kCc / 16; ++kk)   // B: 16 c-rows per step = 2048 B; d panels 8 KB apart (LBO)
          umma_bf16(tDX, umma_desc_sw128(dh_addr + kk * 32, 16, 1024), umma_desc_sw128(w1_addr + kk * 2048, kCc * 128, 1024),
                    idescX, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        umma_commit(&gempty[j & 1]);
      };
      for (int j = 0; j < nchunks; ++j) {
        const int s = j % S;
        mbar_wait(&full[s], (j / S) & 1);
        mbar_wait(&hempty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t w1_addr = smem_u32(sW + s * C::kStageBytes);
        const uint32_t w2_addr = w1_addr + C::kW1Bytes;
        const uint32_t tH = tmem_base + (j & 1) * kCc;
        const uint32_t tG = tmem_base + 128 + (j & 1) * kCc;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tH, umma_desc_sw128(x_addr + (kk >> 2) * (kRows * 128) + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(w1_addr + (kk >> 2) * (kCc * 128) + (kk & 3) * 32, 16, 1024), idescH, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)    // B = W2 tile [DP d-rows][64 c]: 16 d-rows per step = 2048 B
          umma_bf16(tG, umma_desc_sw128(dy_addr + (kk >> 2) * (kRows * 128) + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(w2_addr + kk * 2048, 8192, 1024), idescG, kk > 0 ? 1u : 0u);
        umma_commit(&hfull[j & 1]);
        if (j > 0) gemm_dx(j - 1);
      }
      gemm_dx(nchunks - 1);
      umma_commit(yfull);
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int row = m0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    for (int j = grp; j < nchunks; j += 2) {
      mbar_wait(&hfull[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int c0 = j * kCc;
      uint8_t* sd = sdH + (j & 1) * kGBytes;
      bool waited = false;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t h[32], dg[32];
        tmem_ld32(tmem_base + lane_addr + (j & 1) * kCc + half * 32, h);
        tmem_ld32(tmem_base + lane_addr + 128 + (j & 1) * kCc + half * 32, dg);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&hempty[j & 1]);
        }
        if (!waited) {
          mbar_wait(&gempty[j & 1], ((j >> 1) & 1) ^ 1);
          waited = true;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int cc = c0 + half * 32 + ch * 8;
          float gv[8], dv[8], bias[8];
          if (p.bias_smem) {
            const float4 b0 = *reinterpret_cast<const float4*>(sBias + cc);
            const float4 b1v = *reinterpret_cast<const float4*>(sBias + cc + 4);
            bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
            bias[4] = b1v.x; bias[5] = b1v.y; bias[6] = b1v.z; bias[7] = b1v.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) bias[e] = (cc + e < p.C) ? __ldg(p.b1 + cc + e) : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float x = __uint_as_float(h[ch * 8 + e]) + bias[e];
            float dgelu;
            if (p.exact_gelu) {
              gv[e] = gelu_erf(x);
              dgelu = gelu_erf_grad(x);
            } else {
              gv[e] = gelu_fast_grad(x, dgelu);
            }
            dv[e] = __uint_as_float(dg[ch * 8 + e]) * dgelu;
          }
          if (p.dh.thresh) {   // G' = m*s*G ; dH = dG' * m*s*gelu'(h)
            const unsigned long long i0 = static_cast<unsigned long long>(row) * p.ldh + cc;
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              drop_apply2(p.dh, gv[e], gv[e + 1], i0 + e);
              drop_apply2(p.dh, dv[e], dv[e + 1], i0 + e);
            }
          }
          const uint4 go = make_uint4(pack_bf16(gv[0], gv[1]), pack_bf16(gv[2], gv[3]), pack_bf16(gv[4], gv[5]), pack_bf16(gv[6], gv[7]));
          const uint4 dh = make_uint4(pack_bf16(dv[0], dv[1]), pack_bf16(dv[2], dv[3]), pack_bf16(dv[4], dv[5]), pack_bf16(dv[6], dv[7]));
          *reinterpret_cast<uint4*>(sd + sw128_offset(r, half * 4 + ch)) = dh;
          if (row < p.M && cc < p.ldh) {   // ldh is a multiple of 8 >= C: whole 16-byte chunks only
            *reinterpret_cast<uint4*>(p.g_b + static_cast<long long>(row) * p.ldh + cc) = go;
            *reinterpret_cast<uint4*>(p.dh_b + static_cast<long long>(row) * p.ldh + cc) = dh;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&gfull[j & 1]);
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

template <int DP, bool kBwd>
int launch_chain(const CUtensorMap& t1, const CUtensorMap& t2, const ChainParams& p, cudaStream_t s) {
  auto kern = kBwd ? chain_bwd_kernel<DP> : chain_fwd_kernel<DP>;
  constexpr int base = kBwd ? Cfg<DP>::kBwdSmem : Cfg<DP>::kFwdSmem;
  constexpr int kMaxBias = DP > 128 ? 0 : (kBwd ? kMaxBiasSmemBwd : kMaxBiasSmemFwd);
  const int bias_bytes = ceil_div(p.C, kCc) * kCc * 4;
  ChainParams pp = p;
  pp.bias_smem = bias_bytes <= kMaxBias ? 1 : 0;
  const int smem = base + (pp.bias_smem ? bias_bytes : 0);
  static int configured = 0;   // largest dynamic smem size opted into so far
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return M2_ERR_LAUNCH;
    configured = smem;
  }
  LaunchScope scope(kBwd ? "chain_bwd" : "chain_fwd", s);
  kern<<<ceil_div(p.M, kRows), kThreads, smem, s>>>(t1, t2, pp);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int make_weight_maps(CUtensorMap* t1, CUtensorMap* t2, const void* w1b, const void* w2b, int D, int C, int ldw2, int DP) {
  // W1 bf16 [C][D] (ld = D): box 64 c-rows x 64 d.   W2 bf16 [D][ldw2] (cols >= C zero): box DP d-rows x 64 c.
  int rc = make_tmap_bf16(t1, w1b, C, D, D, kCc, 64);
  if (rc) return rc;
  return make_tmap_bf16(t2, w2b, D, ldw2, ldw2, DP, kCc);
}

}  // namespace

// Which hidden sizes the fused chains cover (others take the unfused GEMM path in abi.cu).
bool chain_fwd_supported(int D) { return D >= 16 && D <= 256 && D % 8 == 0; }
bool chain_bwd_supported(int D) { return D >= 16 && D <= 128 && D % 8 == 0; }

int chain_fwd(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
              int ldw2, const float* b2, float* y, int M, int D, int C, int exact_gelu, float drop_p, unsigned long long seed,
              cudaStream_t s) {
  if (!chain_fwd_supported(D) || ldw2 % 8 || ldw2 < C) return M2_ERR_ARG;
  const int DP = D <= 64 ? 64 : (D <= 128 ? 128 : 256);
  CUtensorMap t1, t2;
  int rc = make_weight_maps(&t1, &t2, w1b, w2b, D, C, ldw2, DP);
  if (rc) return rc;
  ChainParams p = {};
  p.u = u; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2; p.y = y;
  p.M = M; p.D = D; p.C = C; p.exact_gelu = exact_gelu; p.ldh = (C + 7) & ~7;
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden); p.dout = make_drop(drop_p, seed, kSiteChannelOut);
  if (DP == 64) return launch_chain<64, false>(t1, t2, p, s);
  if (DP == 128) return launch_chain<128, false>(t1, t2, p, s);
  return launch_chain<256, false>(t1, t2, p, s);
}

int chain_bwd(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
              int ldw2, const float* dy, void* xn_b, void* dy_b, void* g_b, void* dh_b, int ldh, float* dxn, int M, int D,
              int C, int exact_gelu, float drop_p, unsigned long long seed, cudaStream_t s) {
  if (!chain_bwd_supported(D) || ldw2 % 8 || ldw2 < C || ldh % 8 || ldh < C) return M2_ERR_ARG;
  const int DP = D <= 64 ? 64 : 128;
  CUtensorMap t1, t2;
  int rc = make_weight_maps(&t1, &t2, w1b, w2b, D, C, ldw2, DP);
  if (rc) return rc;
  ChainParams p = {};
  p.u = u; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.dy = dy;
  p.xn_b = static_cast<__nv_bfloat16*>(xn_b); p.dy_b = static_cast<__nv_bfloat16*>(dy_b);
  p.g_b = static_cast<__nv_bfloat16*>(g_b); p.dh_b = static_cast<__nv_bfloat16*>(dh_b);
  p.dxn = dxn; p.M = M; p.D = D; p.C = C; p.ldh = ldh; p.exact_gelu = exact_gelu;
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden); p.dout = make_drop(drop_p, seed, kSiteChannelOut);
  if (DP == 64) return launch_chain<64, true>(t1, t2, p, s);
  return launch_chain<128, true>(t1, t2, p, s);
}

}  // namespace m2
