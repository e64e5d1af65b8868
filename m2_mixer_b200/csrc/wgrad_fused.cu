// Channel-mixing weight gradients with the hidden activation RECOMPUTED on chip (no G / dH round trip through HBM).
//
// Reference arithmetic (autograd of MixerBlock.channel_mix, modules/mixer.py:37-40):
//     H = LN(u) W1^T + b1 ;  G = Drop(GELU(H)) ;  dG = dY W2 ;  dH = dG * Drop'(.) * GELU'(H)
//     dW1 += dH^T LN(u)      dW2 += dY^T G      db1 += colsum(dH)
//
// The dgrad chain (chain_ts.cu) walks token-row tiles and needs every channel of a row; the weight gradients need
// every ROW of a channel.  Instead of spilling G and dH ([M x C] bf16 each, 200 MB per block at B = 4096) this kernel
// is tiled the other way round: a CTA owns one 64-channel chunk (its W1 / W2 slices stay resident in shared memory,
// its dW1^T / dW2 slices accumulate in TMEM for the whole kernel) and streams the bf16 LN(u) / dY row tiles (8 MB per
// block, L2 resident) through a TMA ring:
//     per 128-row tile i:   H  = Xn_i . W1c^T        dG = dY_i . W2c                      (recompute, 2 x 8 MMA N=64)
//                           epilogue: G, dH -> bf16 -> shared memory (MN-major B operands), db1 partials in registers
//                           dW1c^T += Xn_i^T . dH    dW2c += dY_i^T . G                   (A = the SAME row tiles,
//                                                                                          consumed MN-major)
// grid = (C / 64 chunks) x R row splits, R chosen so that one wave fills the 148 SMs; the R partial sums of a chunk
// are combined with fp32 reductions into the running gradients.
#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"

namespace m2 {
namespace {

// Row tiles are 96 rows: three 48 KB ring stages fit beside the resident weight slices, and a third stage is what hides
// the TMA latency (with two, the load of tile i + 2 can only start when the gradient GEMMs of tile i retire and the
// in-order MMA issuer stalls on it: 5200 clk per tile measured, profiles/r01_ncu_chain_v9.md).  The recompute GEMMs
// still run as M = 128 instructions; accumulator rows 96..127 are garbage that no gradient GEMM (K = 96) ever reads.
constexpr int kRows = 96;       // token rows per tile (K of the gradient GEMMs)
constexpr int kMmaM = 128;      // UMMA M of the recompute GEMMs
constexpr int kCc = 64;         // channels per CTA
constexpr int kThreads = 320;   // warp0 TMA, warp1 MMA, warps 2-9 epilogue (group g = column half g of the chunk)
constexpr int kNS = 3;          // row-tile ring depth

template <int DP>
struct CfgW {
  static constexpr int kPanel = kRows * 128;            // one [128 rows][64 d] SW128 panel
  static constexpr int kTile = (DP / 64) * kPanel;      // LN(u) or dY row tile
  static constexpr int kStage = 2 * kTile;
  static constexpr int kW1Bytes = kCc * DP * 2;         // [64 c][DP d]
  static constexpr int kW2Bytes = DP * kCc * 2;         // [DP d][64 c]
  static constexpr int kGBytes = kMmaM * kCc * 2;       // [128 rows][64 c] (rows >= 96 are written but never read)
  static constexpr int kSmem = kNS * kStage + kW1Bytes + kW2Bytes + 2 * kGBytes + 1024 + 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kColW1 = 0, kColW2 = 64, kColH = 128, kColG = 256;
};

struct WgParams {
  const float* b1;
  float* dw1;   // [C][D]
  float* dw2;   // [D][C]
  float* db1;   // [C]
  int M, D, C, ldh;
  int ntiles, R;
  Drop dh;
};

template <int DP, bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                   const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const WgParams p) {
  using C = CfgW<DP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem;                              // [kNS][Xn tile | dY tile]
  uint8_t* sW1 = sStage + kNS * C::kStage;
  uint8_t* sW2 = sW1 + C::kW1Bytes;
  uint8_t* sG = sW2 + C::kW2Bytes;
  uint8_t* sdH = sG + C::kGBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdH + C::kGBytes);
  uint64_t* wfull = bars;             // [1]   W1c / W2c landed
  uint64_t* full = wfull + 1;         // [kNS] row tile landed
  uint64_t* empty = full + kNS;       // [kNS] gradient GEMMs done with the row tile -> TMA
  uint64_t* hfull = empty + kNS;      // [2]   H / dG accumulators ready -> epilogue
  uint64_t* hempty = hfull + 2;       // [2]   epilogue has read them -> MMA
  uint64_t* gfull = hempty + 2;       // [1]   epilogue wrote sG / sdH -> MMA
  uint64_t* gempty = gfull + 1;       // [1]   gradient GEMMs done with sG / sdH -> epilogue
  uint64_t* accfull = gempty + 1;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
  float* sDb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);   // [64] db1 partials (16-byte aligned)
  float* sB1 = sDb + kCc;                                  // [64] b1 of the chunk

  // logical roles 0 = TMA, 1 = MMA, 2.. = epilogue; physically the epilogue warps come first (see chain_ts.cu)
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = pwarp < 8 ? pwarp + 2 : pwarp - 8;
  const int c0 = blockIdx.x * kCc;
  const int t_lo = static_cast<int>(static_cast<long long>(p.ntiles) * blockIdx.y / p.R);
  const int t_hi = static_cast<int>(static_cast<long long>(p.ntiles) * (blockIdx.y + 1) / p.R);
  const int nt = t_hi - t_lo;
  if (nt <= 0) return;   // uniform for the CTA

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int i = 0; i < kNS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], 256); }
    mbar_init(gfull, 256);
    mbar_init(gempty, 1);
    mbar_init(accfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  if (threadIdx.x < kCc) {
    sDb[threadIdx.x] = 0.f;
    sB1[threadIdx.x] = (c0 + threadIdx.x < p.C) ? p.b1[c0 + threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- TMA producer (whole warp walks the loop, one elected lane issues)
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, C::kW1Bytes + C::kW2Bytes);
#pragma unroll
      for (int pnl = 0; pnl < DP / 64; ++pnl) tma_load_2d(sW1 + pnl * (kCc * 128), &tmW1, wfull, pnl * 64, c0);
      tma_load_2d(sW2, &tmW2, wfull, c0, 0);
    }
    __syncwarp();
    for (int i = 0; i < nt; ++i) {
      const int s = i % kNS;
      mbar_wait(&empty[s], ((i / kNS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], C::kStage);
        uint8_t* dst = sStage + s * C::kStage;
        const int row0 = (t_lo + i) * kRows;
#pragma unroll
        for (int pnl = 0; pnl < DP / 64; ++pnl) {
          tma_load_2d(dst + pnl * C::kPanel, &tmX, &full[s], pnl * 64, row0);
          tma_load_2d(dst + C::kTile + pnl * C::kPanel, &tmDY, &full[s], pnl * 64, row0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    constexpr uint32_t idescH = umma_idesc_bf16(kMmaM, kCc, 0, 0);   // A row tile K-major,  B = W1c K-major
    constexpr uint32_t idescG = umma_idesc_bf16(kMmaM, kCc, 0, 1);   // A row tile K-major,  B = W2c MN-major
    constexpr uint32_t idescW = umma_idesc_bf16(kMmaM, kCc, 1, 1);   // A row tile MN-major (M = d), B = sdH / sG MN-major
    constexpr uint32_t kLboA = DP == 128 ? C::kPanel : 0;            // DP = 64: M rows 64..127 alias the only panel
    const uint64_t xk0 = umma_desc_sw128(smem_u32(sStage), 16, 1024);
    const uint64_t xm0 = umma_desc_sw128(smem_u32(sStage), kLboA, 1024);
    const uint64_t w1d = umma_desc_sw128(smem_u32(sW1), 16, 1024);
    const uint64_t w2d = umma_desc_sw128(smem_u32(sW2), 8192, 1024);
    const uint64_t gd = umma_desc_sw128(smem_u32(sG), 8192, 1024);
    const uint64_t dhd = umma_desc_sw128(smem_u32(sdH), 8192, 1024);
    auto hg = [&](int i) {   // H[i&1] = Xn_i . W1c^T ; dG[i&1] = dY_i . W2c
      const int s = i % kNS, b = i & 1;
      mbar_wait(&full[s], (i / kNS) & 1);
      mbar_wait(&hempty[b], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xa = xk0 + static_cast<uint64_t>((s * C::kStage) >> 4);
        const uint64_t ya = xa + static_cast<uint64_t>(C::kTile >> 4);
        const uint32_t tH = tmem_base + C::kColH + b * kCc;
        const uint32_t tG = tmem_base + C::kColG + b * kCc;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tH, xa + (((kk >> 2) * C::kPanel + (kk & 3) * 32) >> 4), w1d + (((kk >> 2) * (kCc * 128) + (kk & 3) * 32) >> 4),
                    idescH, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tG, ya + (((kk >> 2) * C::kPanel + (kk & 3) * 32) >> 4), w2d + ((kk * 2048) >> 4), idescG, kk > 0 ? 1u : 0u);
        umma_commit(&hfull[b]);
      }
      __syncwarp();
    };
    mbar_wait(wfull, 0);
    hg(0);
    if (nt > 1) hg(1);
    // Steady state: ONE burst per tile - the gradient GEMMs of tile i and the recompute GEMMs of tile i + 2 are issued
    // together after all their waits (every wake-up of the issuer costs several hundred cycles, chain_ts.cu).
    for (int i = 0; i < nt; ++i) {   // dW1c^T += Xn_i^T . dH_i ; dW2c += dY_i^T . G_i   (contraction over the rows)
      const int s = i % kNS;
      const int in = i + 2, sn = in % kNS, b = i & 1;
      const bool more = in < nt;
      mbar_wait(gfull, i & 1);
      if (more) {
        mbar_wait(&full[sn], (in / kNS) & 1);
        mbar_wait(&hempty[b], ((in >> 1) & 1) ^ 1);
      }
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xa = xm0 + static_cast<uint64_t>((s * C::kStage) >> 4);
        const uint64_t ya = xa + static_cast<uint64_t>(C::kTile >> 4);
#pragma unroll
        for (int kk = 0; kk < kRows / 16; ++kk)   // 16 rows per step = 2048 B in both operands
          umma_bf16(tmem_base + C::kColW1, xa + ((kk * 2048) >> 4), dhd + ((kk * 2048) >> 4), idescW, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < kRows / 16; ++kk)
          umma_bf16(tmem_base + C::kColW2, ya + ((kk * 2048) >> 4), gd + ((kk * 2048) >> 4), idescW, (i > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        umma_commit(gempty);
        if (more) {
          const uint64_t xk = xk0 + static_cast<uint64_t>((sn * C::kStage) >> 4);
          const uint64_t yk = xk + static_cast<uint64_t>(C::kTile >> 4);
          const uint32_t tH = tmem_base + C::kColH + b * kCc;
          const uint32_t tG = tmem_base + C::kColG + b * kCc;
#pragma unroll
          for (int kk = 0; kk < DP / 16; ++kk)
            umma_bf16(tH, xk + (((kk >> 2) * C::kPanel + (kk & 3) * 32) >> 4), w1d + (((kk >> 2) * (kCc * 128) + (kk & 3) * 32) >> 4),
                      idescH, kk > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < DP / 16; ++kk)
            umma_bf16(tG, yk + (((kk >> 2) * C::kPanel + (kk & 3) * 32) >> 4), w2d + ((kk * 2048) >> 4), idescG, kk > 0 ? 1u : 0u);
          umma_commit(&hfull[b]);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accfull);
    __syncwarp();
  } else {
    // ---- epilogue: thread = row of the tile (TMEM lane), group g = columns [32 g, 32 g + 32) of the chunk
    const int q = pwarp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int cg = c0 + grp * 32;
    float dbp[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) dbp[k] = 0.f;
    const uint32_t bias_addr = smem_u32(sB1 + grp * 32);
    const bool live = r < kRows;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      mbar_wait(&hfull[b], (i >> 1) & 1);
      tc_fence_after();
      uint32_t h[32], dg[32];
      tmem_ld32(tmem_base + C::kColH + lane_addr + b * kCc + grp * 32, h);
      tmem_ld32(tmem_base + C::kColG + lane_addr + b * kCc + grp * 32, dg);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&hempty[b]);
      uint32_t gp[16], dp[16];
      const unsigned long long i0 = static_cast<unsigned long long>((t_lo + i) * kRows + r) * p.ldh + cg;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float bias[8];
#pragma unroll
        for (int e = 0; e < 2; ++e)
          asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
              : "=f"(bias[4 * e]), "=f"(bias[4 * e + 1]), "=f"(bias[4 * e + 2]), "=f"(bias[4 * e + 3])
              : "r"(bias_addr + (ch * 8 + 4 * e) * 4));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = ch * 8 + 2 * e;
          float2 dgelu;
          float2 gv = gelu2_grad(__fadd2_rn(make_float2(__uint_as_float(h[k]), __uint_as_float(h[k + 1])),
                                            make_float2(bias[2 * e], bias[2 * e + 1])), dgelu);
          float2 dv = __fmul2_rn(make_float2(__uint_as_float(dg[k]), __uint_as_float(dg[k + 1])), dgelu);
          if (kDrop) {
            drop_apply2(p.dh, gv.x, gv.y, i0 + k);
            drop_apply2(p.dh, dv.x, dv.y, i0 + k);
          }
          if (live) { dbp[k] += dv.x; dbp[k + 1] += dv.y; }
          gp[ch * 4 + e] = pack_bf16(gv.x, gv.y);
          dp[ch * 4 + e] = pack_bf16(dv.x, dv.y);
        }
      }
      mbar_wait(gempty, (i & 1) ^ 1);   // gradient GEMMs of tile i - 1 have consumed sG / sdH
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        *reinterpret_cast<uint4*>(sG + sw128_offset(r, grp * 4 + k)) = make_uint4(gp[4 * k], gp[4 * k + 1], gp[4 * k + 2], gp[4 * k + 3]);
        *reinterpret_cast<uint4*>(sdH + sw128_offset(r, grp * 4 + k)) = make_uint4(dp[4 * k], dp[4 * k + 1], dp[4 * k + 2], dp[4 * k + 3]);
      }
      fence_proxy_async();
      mbar_arrive(gfull);
    }
    // db1: reduce the per-row partials over the 32 rows of the warp, then over the warps (shared-memory atomics)
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float sres = warp_sum(dbp[k]);
      if (lane == k) mine = sres;
    }
    atomicAdd(&sDb[grp * 32 + lane], mine);
    // accumulators: TMEM lane = d.  group 0: dW1^T slice -> dw1[c][d] (lanes contiguous in d: coalesced reductions);
    //                               group 1: dW2 slice   -> dw2[d][c]
    mbar_wait(accfull, 0);
    tc_fence_after();
    const int d = r;
#pragma unroll 1
    for (int cb = 0; cb < kCc; cb += 32) {
      uint32_t a[32];
      tmem_ld32(tmem_base + (grp == 0 ? C::kColW1 : C::kColW2) + lane_addr + cb, a);
      tmem_ld_wait();
      if (d < p.D) {
        if (grp == 0) {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + cb + k < p.C) atomicAdd(p.dw1 + static_cast<long long>(c0 + cb + k) * p.D + d, __uint_as_float(a[k]));
        } else {
          float* dst = p.dw2 + static_cast<long long>(d) * p.C + c0 + cb;
          if ((p.C & 3) == 0 && c0 + cb + 32 <= p.C) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              atomicAdd(reinterpret_cast<float4*>(dst + k), make_float4(__uint_as_float(a[k]), __uint_as_float(a[k + 1]),
                                                                        __uint_as_float(a[k + 2]), __uint_as_float(a[k + 3])));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (c0 + cb + k < p.C) atomicAdd(dst + k, __uint_as_float(a[k]));
          }
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int t = (warp - 2) * 32 + lane;
    if (t < kCc && c0 + t < p.C) atomicAdd(p.db1 + c0 + t, sDb[t]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int DP, bool kDrop>
int launch_wg(const CUtensorMap& tx, const CUtensorMap& ty, const CUtensorMap& t1, const CUtensorMap& t2, const WgParams& p,
              cudaStream_t s) {
  auto kern = wgrad_fused_kernel<DP, kDrop>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgW<DP>::kSmem) != cudaSuccess) return M2_ERR_LAUNCH;
    configured = true;
  }
  LaunchScope scope("wgrad_fused", s);
  dim3 grid(ceil_div(p.C, kCc), p.R);
  kern<<<grid, kThreads, CfgW<DP>::kSmem, s>>>(tx, ty, t1, t2, p);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace

// dw1 [C][D], dw2 [D][C], db1 [C] are ACCUMULATED (fp32 reductions).  xn_b / dy_b: bf16 [M][D] written by chain_bwd_ts.
int wgrad_fused(const void* xn_b, const void* dy_b, const void* w1b, const void* w2b, int ldw2, const float* b1, float* dw1,
                float* db1, float* dw2, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s) {
  if (!chain_fwd_ts_supported(D) || ldw2 % 8 || ldw2 < C) return M2_ERR_ARG;
  const int DP = D <= 64 ? 64 : 128;
  CUtensorMap tx, ty, t1, t2;
  int rc = make_tmap_bf16(&tx, xn_b, M, D, D, kRows, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&ty, dy_b, M, D, D, kRows, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&t1, w1b, C, D, D, kCc, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&t2, w2b, D, ldw2, ldw2, DP, kCc);
  if (rc) return rc;
  WgParams p = {};
  p.b1 = b1; p.dw1 = dw1; p.dw2 = dw2; p.db1 = db1;
  p.M = M; p.D = D; p.C = C; p.ldh = (C + 7) & ~7;
  p.ntiles = ceil_div(M, kRows);
  const int nch = ceil_div(C, kCc);
  int R = 148 / nch;
  if (R < 1) R = 1;
  if (R > p.ntiles) R = p.ntiles;
  p.R = R;
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden);
  const bool drop = p.dh.thresh != 0;
  if (DP == 64) return drop ? launch_wg<64, true>(tx, ty, t1, t2, p, s) : launch_wg<64, false>(tx, ty, t1, t2, p, s);
  return drop ? launch_wg<128, true>(tx, ty, t1, t2, p, s) : launch_wg<128, false>(tx, ty, t1, t2, p, s);
}

}  // namespace m2
