// Channel-mixing weight gradients.  Two kernels:
//   wgrad_dh_kernel     (default, M2B200_CHAIN_GEN=4): dH comes back from the dgrad chain's spill by TMA, only G is recomputed
//   wgrad_fused_kernel  (M2B200_CHAIN_GEN=2, the A/B form): G AND dH recomputed on chip, nothing spilled
//
// Reference arithmetic (autograd of MixerBlock.channel_mix, modules/mixer.py:37-40):
//     H = LN(u) W1^T + b1 ;  G = Drop(GELU(H)) ;  dG = dY W2 ;  dH = dG * Drop'(.) * GELU'(H)
//     dW1 += dH^T LN(u)      dW2 += dY^T G      db1 += colsum(dH)
//
// The dgrad chain (chain_ts.cu) walks token-row tiles and needs every channel of a row; the weight gradients need
// every ROW of a channel.  Both kernels here are tiled the other way round: a CTA owns one 128-channel slice (its W1 / W2
// slices stay resident in shared memory, its dW1^T / dW2 slices accumulate in TMEM for the whole kernel) and streams the bf16
// LN(u) / dY row tiles (8 MB per block, L2 resident) through TMA:
//     per 96-row tile i:    H  = Xn_i . W1c^T      ( dG = dY_i . W2c  in wgrad_fused only )          recompute GEMMs, N = 128
//                           epilogue: G ( and dH ) -> bf16 -> shared memory (MN-major B operands)
//                           dW1c^T += Xn_i^T . dH    dW2c += dY_i^T . G                   (A = the SAME row tiles,
//                                                                                          consumed MN-major)
// grid = (C / 128 slices) x R row splits, R chosen so that one wave fills the 148 SMs; the R partial sums of a slice
// are combined with fp32 reductions into the running gradients.  wgrad_fused_kernel is described first (round 1), the
// generation-4 kernel and what its measurements changed follow below.
#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"

namespace m2 {
namespace {

// A CTA owns 128 channels and walks 96-row tiles.  Why these numbers (measured, profiles/r01_chain_microbench_v25.log, profiles/r01_trace_wgrad_final.log):
//  * every wake-up of the single MMA-issuing thread costs several hundred cycles (mbarrier waits, burst start-up), so the
//    work per wake-up must be large: 128-channel GEMMs (N = 128, 64 clk per MMA instead of 48 for N = 64: the same
//    wake-ups now feed twice the channels, and A is read from shared memory once per 128 instead of per 64 channels);
//  * 96-row tiles: two 48 KB ring stages + 64 KB of resident weight slices + 64 KB of G / dH operand tiles = 224 KB.  The
//    recompute GEMMs still run as M = 128 instructions; accumulator rows 96..127 are garbage that no gradient GEMM
//    (K = 96) ever reads.
constexpr int kRows = 96;       // token rows per tile (K of the gradient GEMMs)
constexpr int kMmaM = 128;      // UMMA M of the recompute GEMMs
constexpr int kCc = 128;        // channels per CTA
// 16 warps = 4 column groups (group g = columns [32 g, 32 g + 32)) x 4 TMEM lane quadrants.  A warp may only touch the
// TMEM lanes of quadrant (warp id % 4), which is also its scheduler: with 96-row tiles quadrant 3 (rows 96..127) holds
// garbage, so its four warps do no epilogue work - they are the two TMA producers (warps 3, 15) and the two MMA issuers
// (warps 7, 11), which then have scheduler 3 to themselves (no busy epilogue warp delays an MMA burst), and the twelve live warps
// put four warps on each of the other three schedulers (two before: the GELU' epilogue was latency bound at 0.5 IPC).
constexpr int kThreads = 512;
constexpr int kLiveThreads = 384;
constexpr int kNS = 2;          // row-tile buffers: [0] feeds the recompute GEMMs, [1] the gradient GEMMs (see the kernel)

template <int DP>
struct CfgW {
  static constexpr int kPanel = kRows * 128;            // one [96 rows][64 d] SW128 panel of a row tile
  static constexpr int kTile = (DP / 64) * kPanel;      // LN(u) or dY row tile
  static constexpr int kStage = 2 * kTile;
  static constexpr int kW1Panel = kCc * 128;            // [128 c][64 d]
  static constexpr int kW1Bytes = (DP / 64) * kW1Panel;
  static constexpr int kW2Panel = DP * 128;             // [DP d][64 c]
  static constexpr int kW2Bytes = 2 * kW2Panel;
  static constexpr int kGPanel = kMmaM * 128;           // [128 rows][64 c]
  static constexpr int kGBytes = 2 * kGPanel;
  static constexpr int kSmem = kNS * kStage + kW1Bytes + kW2Bytes + 2 * kGBytes + 1280 + 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kColW1 = 0, kColW2 = 128, kColH = 256, kColG = 384;
};

// Optional in-kernel timeline (debug builds: -DM2_TRACE, tools/trace_wgrad.cu): CTA (0,0) records (tag, tile, clock).
#ifdef M2_TRACE
__device__ long long g_wtrace[4096];
__device__ __forceinline__ void wtrace_evt(int slot, int tag, int j) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && slot < 1360) {
    g_wtrace[3 * slot] = tag; g_wtrace[3 * slot + 1] = j; g_wtrace[3 * slot + 2] = clock64();
  }
}
#define M2_WTR(slot, tag, j) wtrace_evt(slot, tag, j)
#else
#define M2_WTR(slot, tag, j)
#endif

struct WgParams {
  const float* b1;
  float* dw1;   // [C][D]
  float* dw2;   // [D][C]
  float* db1;   // [C]
  int M, D, C, ldh;
  int ntiles, R;
  int l2_hint;  // dH loads carry an L2 evict_first policy (M2B200_DH_L2HINT, default on)
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
  uint64_t* hfull = empty + kNS;      // [1]   H / dG accumulators ready -> epilogue
  uint64_t* hempty = hfull + 1;       // [1]   epilogue has them in registers -> MMA
  uint64_t* gfull = hempty + 1;       // [1]   epilogue wrote sG / sdH -> MMA
  uint64_t* gempty = gfull + 1;       // [1]   gradient GEMMs done with sG / sdH -> epilogue
  uint64_t* accfull = gempty + 1;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
  float* sDb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);   // [128] db1 partials (16-byte aligned)
  float* sB1 = sDb + kCc;                                                          // [128] b1 of the chunk

  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = pwarp & 3;             // TMEM lane quadrant (= scheduler)
  const int grp = pwarp >> 2;          // column group
  const bool is_tma = pwarp == 3, is_mma_hg = pwarp == 7, is_mma_wg = pwarp == 11, is_tma_b = pwarp == 15;
  const int c0 = blockIdx.x * kCc;
  const int t_lo = static_cast<int>(static_cast<long long>(p.ntiles) * blockIdx.y / p.R);
  const int t_hi = static_cast<int>(static_cast<long long>(p.ntiles) * (blockIdx.y + 1) / p.R);
  const int nt = t_hi - t_lo;
  if (nt <= 0) return;   // uniform for the CTA

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int i = 0; i < kNS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(hfull, 1);
    mbar_init(hempty, kLiveThreads);
    mbar_init(gfull, kLiveThreads);
    mbar_init(gempty, 1);
    mbar_init(accfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
  }
  if (pwarp == 0) tmem_alloc(tmem_slot, C::kTmemCols);
  if (threadIdx.x < kCc) {
    sDb[threadIdx.x] = 0.f;
    sB1[threadIdx.x] = (c0 + threadIdx.x < p.C) ? p.b1[c0 + threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (is_tma) {
    // ---- TMA producer (whole warp walks the loop, one elected lane issues)
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, C::kW1Bytes + C::kW2Bytes);
#pragma unroll
      for (int pnl = 0; pnl < DP / 64; ++pnl)
#pragma unroll
        for (int h = 0; h < 2; ++h)   // [64 c][64 d] boxes: d panel pnl, channel half h
          tma_load_2d(sW1 + pnl * C::kW1Panel + h * (64 * 128), &tmW1, wfull, pnl * 64, c0 + h * 64);
#pragma unroll
      for (int h = 0; h < 2; ++h)     // [DP d][64 c] boxes: channel panel h
        tma_load_2d(sW2 + h * C::kW2Panel, &tmW2, wfull, c0 + h * 64, 0);
    }
    __syncwarp();
    // Row tile i is needed twice: by the recompute GEMMs (early) and by the gradient GEMMs (after the epilogue of the tile).
    // Held in one 2-slot ring it stayed resident for two tile periods and the ring (wg(i) done -> TMA of tile i + 2 ->
    // hg(i + 2) -> epilogue -> wg(i + 2)) paced the kernel at 3400-3700 clk per tile for a 3100-clk epilogue
    // (profiles/r01_trace_wgrad.log).  So the tile is fetched TWICE from L2 into two single buffers with independent
    // lifetimes: buffer 0 (this warp) is refilled as soon as hg(i) completes, buffer 1 (warp 15) as soon as wg(i) completes.
    for (int i = 0; i < nt; ++i) {
      mbar_wait(&empty[0], (i & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[0], C::kStage);
        const int row0 = (t_lo + i) * kRows;
#pragma unroll
        for (int pnl = 0; pnl < DP / 64; ++pnl) {
          tma_load_2d(sStage + pnl * C::kPanel, &tmX, &full[0], pnl * 64, row0);
          tma_load_2d(sStage + C::kTile + pnl * C::kPanel, &tmDY, &full[0], pnl * 64, row0);
        }
      }
      __syncwarp();
    }
  } else if (is_tma_b) {
    for (int i = 0; i < nt; ++i) {
      mbar_wait(&empty[1], (i & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[1], C::kStage);
        uint8_t* dst = sStage + C::kStage;
        const int row0 = (t_lo + i) * kRows;
#pragma unroll
        for (int pnl = 0; pnl < DP / 64; ++pnl) {
          tma_load_2d(dst + pnl * C::kPanel, &tmX, &full[1], pnl * 64, row0);
          tma_load_2d(dst + C::kTile + pnl * C::kPanel, &tmDY, &full[1], pnl * 64, row0);
        }
      }
      __syncwarp();
    }
  } else if (is_mma_hg) {
    // ---- recompute issuer: H = Xn_i . W1c^T ; dG = dY_i . W2c   (N = 128), as soon as the epilogue has tile i - 1's
    // accumulators in registers.  Two issuing warps (this one and the gradient issuer below) so that one's barrier waits
    // overlap the other's MMAs (tcgen05.mma issue blocks at the execution rate and a completed-barrier wait costs ~130 clk:
    // see chain_fwd_ts_kernel); each accumulator has one writer, and every hand-over goes through the epilogue's barriers.
    constexpr uint32_t idescH = umma_idesc_bf16(kMmaM, kCc, 0, 0);   // A row tile K-major,  B = W1c K-major (N = 128 rows)
    constexpr uint32_t idescG = umma_idesc_bf16(kMmaM, kCc, 0, 1);   // A row tile K-major,  B = W2c MN-major
    const uint64_t xk0 = umma_desc_sw128(smem_u32(sStage), 16, 1024);
    const uint64_t w1d = umma_desc_sw128(smem_u32(sW1), 16, 1024);
    const uint64_t w2d = umma_desc_sw128(smem_u32(sW2), C::kW2Panel, 1024);
    mbar_wait(wfull, 0);
    for (int i = 0; i < nt; ++i) {
      constexpr int s = 0;
      M2_WTR(4 * i + 0, 1, i);
      mbar_wait2(&full[0], i & 1, hempty, (i & 1) ^ 1);
      M2_WTR(4 * i + 1, 2, i);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xa = xk0 + static_cast<uint64_t>((s * C::kStage) >> 4);
        const uint64_t ya = xa + static_cast<uint64_t>(C::kTile >> 4);
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tmem_base + C::kColH, xa + (((kk >> 2) * C::kPanel + (kk & 3) * 32) >> 4),
                    w1d + (((kk >> 2) * C::kW1Panel + (kk & 3) * 32) >> 4), idescH, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tmem_base + C::kColG, ya + (((kk >> 2) * C::kPanel + (kk & 3) * 32) >> 4), w2d + ((kk * 2048) >> 4), idescG,
                    kk > 0 ? 1u : 0u);
        umma_commit(&empty[0]);
        umma_commit(hfull);
        M2_WTR(4 * i + 2, 3, i);
      }
      __syncwarp();
    }
  } else if (is_mma_wg) {
    // ---- gradient issuer: dW1c^T += Xn_i^T . dH_i ; dW2c += dY_i^T . G_i   (contraction over the 96 rows), when the
    // epilogue has written G / dH and the tile's second copy has landed in buffer 1.
    constexpr uint32_t idescW = umma_idesc_bf16(kMmaM, kCc, 1, 1);   // A row tile MN-major (M = d), B = sdH / sG MN-major
    constexpr uint32_t kLboA = DP == 128 ? C::kPanel : 0;            // DP = 64: M rows 64..127 alias the only panel
    const uint64_t xm0 = umma_desc_sw128(smem_u32(sStage), kLboA, 1024);
    const uint64_t gd = umma_desc_sw128(smem_u32(sG), C::kGPanel, 1024);
    const uint64_t dhd = umma_desc_sw128(smem_u32(sdH), C::kGPanel, 1024);
    for (int i = 0; i < nt; ++i) {
      constexpr int s = 1;
      M2_WTR(200 + 4 * i + 0, 4, i);
      mbar_wait2(gfull, i & 1, &full[1], i & 1);
      M2_WTR(200 + 4 * i + 1, 5, i);
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
        umma_commit(&empty[1]);
        umma_commit(gempty);
        M2_WTR(200 + 4 * i + 2, 6, i);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accfull);
    __syncwarp();
  } else if (q < 3) {
    // ---- epilogue: thread = row of the tile (TMEM lane), group g = columns [32 g, 32 g + 32) in two 16-column pieces
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int cg = c0 + grp * 32;
    float2 dbp[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) dbp[k] = make_float2(0.f, 0.f);
    const uint32_t bias_addr = smem_u32(sB1 + grp * 32);
    const float hs = kDrop ? 0.5f * p.dh.scale : 0.5f;          // dropout scale folded into GELU / GELU'
    const float ninv_s = kDrop ? -1.f / p.dh.scale : -1.f;
    uint8_t* gdst = sG + (grp >> 1) * C::kGPanel;               // [128 rows][64 c] SW128 panel; this group: 16-byte chunks
    uint8_t* hdst = sdH + (grp >> 1) * C::kGPanel;              //   4 (grp & 1) .. 4 (grp & 1) + 3 of the row
    const int chunk0 = (grp & 1) * 4;
    auto ld_piece = [&](int pc, uint32_t (&hd)[16], uint32_t (&gd2)[16]) {
      tmem_ld16(tmem_base + C::kColH + lane_addr + grp * 32 + pc * 16, hd);
      tmem_ld16(tmem_base + C::kColG + lane_addr + grp * 32 + pc * 16, gd2);
    };
    const uint32_t dkey = drop_key(p.dh);
    bool ready = false;                              // hfull of the tile already observed by an early probe
    for (int i = 0; i < nt; ++i) {
      if (pwarp == 0) M2_WTR(400 + 4 * i + 0, 7, i);
      if (!ready) mbar_wait(hfull, i & 1);
      __syncwarp();
      if (pwarp == 0) M2_WTR(400 + 4 * i + 1, 8, i);
      tc_fence_after();
      // dropout: hash input of the quad at this thread's (row, first channel of the group)
      const uint32_t hin = ((static_cast<uint32_t>((t_lo + i) * kRows + r) * static_cast<uint32_t>(p.ldh) + static_cast<uint32_t>(cg)) >> 2) * kDropGolden + dkey;
      uint32_t hA[16], gA[16], hB[16], gB[16];
      // Both pieces are fetched before any math and the accumulators handed back at once: hg(i + 1) (1000 clk of MMAs plus the
      // commit round trip) then runs under this tile's ~3000 clk of epilogue math.  When hempty was only signalled after the
      // first piece's math the recompute GEMMs and the epilogue alternated (profiles/r01_trace_wgrad.log: 3700 clk per tile).
      ld_piece(0, hA, gA);
      ld_piece(1, hB, gB);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(hempty);
      if (pwarp == 0) M2_WTR(400 + 4 * i + 2, 9, i);
#pragma unroll
      for (int pc = 0; pc < 2; ++pc) {
        uint32_t (&h)[16] = pc ? hB : hA;
        uint32_t (&dg)[16] = pc ? gB : gA;
        float bias[16];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
              : "=f"(bias[4 * e]), "=f"(bias[4 * e + 1]), "=f"(bias[4 * e + 2]), "=f"(bias[4 * e + 3])
              : "r"(bias_addr + (pc * 16 + 4 * e) * 4));
        uint32_t gp[8], dp[8];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {   // quads of channels: one mask hash each
          uint32_t flags = 0;
          if (kDrop) flags = drop_flags_from_hash_input(p.dh, hin + static_cast<uint32_t>(pc * 4 + qd) * kDropGolden);
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int e = qd * 2 + e2;
            float2 dgelu;
            const float2 gv = gelu2_grad(__fadd2_rn(make_float2(__uint_as_float(h[2 * e]), __uint_as_float(h[2 * e + 1])),
                                                    make_float2(bias[2 * e], bias[2 * e + 1])), dgelu, hs, ninv_s);
            float2 dv = __fmul2_rn(make_float2(__uint_as_float(dg[2 * e]), __uint_as_float(dg[2 * e + 1])), dgelu);
            gp[e] = pack_bf16(gv.x, gv.y);
            if (kDrop) {   // scale folded into gelu2_grad: G is masked packed, dH in fp32 (db1 accumulates it)
              gp[e] &= e2 ? drop_mask_bf16x2<1>(flags) : drop_mask_bf16x2<0>(flags);
              dv.x = drop_and(dv.x, e2 ? drop_mask_b32<2>(flags) : drop_mask_b32<0>(flags));
              dv.y = drop_and(dv.y, e2 ? drop_mask_b32<3>(flags) : drop_mask_b32<1>(flags));
            }
            dbp[pc * 8 + e] = __fadd2_rn(dbp[pc * 8 + e], dv);
            dp[e] = pack_bf16(dv.x, dv.y);
          }
        }
        if (pc == 0) mbar_wait(gempty, (i & 1) ^ 1);   // gradient GEMMs of tile i - 1 have consumed sG / sdH
        else ready = (i + 1 < nt) && mbar_probe(hfull, (i + 1) & 1);   // next tile's accumulators: round trip hidden under the stores
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          *reinterpret_cast<uint4*>(gdst + sw128_offset(r, chunk0 + pc * 2 + k)) = make_uint4(gp[4 * k], gp[4 * k + 1], gp[4 * k + 2], gp[4 * k + 3]);
          *reinterpret_cast<uint4*>(hdst + sw128_offset(r, chunk0 + pc * 2 + k)) = make_uint4(dp[4 * k], dp[4 * k + 1], dp[4 * k + 2], dp[4 * k + 3]);
        }
      }
      fence_proxy_async();
      mbar_arrive(gfull);
      if (pwarp == 0) M2_WTR(400 + 4 * i + 3, 10, i);
    }
    // db1: reduce the per-row partials over the 32 rows of the warp, then over the warps (shared-memory atomics)
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {   // column k of the group's 32 columns
      const float s0 = warp_sum((k & 1) ? dbp[k >> 1].y : dbp[k >> 1].x);
      if (lane == k) mine = s0;
    }
    atomicAdd(&sDb[grp * 32 + lane], mine);
  }
  // ---- all 16 warps: accumulators -> global.  TMEM lane = d.  groups 0 / 1: dW1^T columns [0,64) / [64,128) -> dw1[c][d]
  // (lanes contiguous in d: coalesced reductions); groups 2 / 3: dW2 columns likewise -> dw2[d][c]
  __syncwarp();
  mbar_wait(accfull, 0);
  tc_fence_after();
  {
    const int d = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool first = grp < 2;
    const int cbase = (grp & 1) * 64;
#pragma unroll 1
    for (int cb = cbase; cb < cbase + 64; cb += 32) {
      uint32_t a[32];
      tmem_ld32(tmem_base + (first ? C::kColW1 : C::kColW2) + lane_addr + cb, a);
      tmem_ld_wait();
      if (d < p.D) {
        if (first) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < kCc && c0 + threadIdx.x < p.C) atomicAdd(p.db1 + c0 + threadIdx.x, sDb[threadIdx.x]);
  if (pwarp == 0) tmem_dealloc(tmem_base, C::kTmemCols);
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


// ---------------------------------------------------------------------------------------------------------------------------
// Generation 4: the dgrad chain spills dH (bf16, chunk-major [C / 64][M][64], TMA stores, chain_ts.cu kStore == 2) and this
// kernel only recomputes G.  Per 96-row tile:
//     H = Xn_i . W1c^T                                   one recompute GEMM instead of two
//     epilogue: G = Drop(GELU(H + b1)) -> bf16 -> shared memory   (forward GELU only: no GELU', no dG, no dH, no db1 partials:
//                                                                  about half the instructions of wgrad_fused's epilogue)
//     dH_i arrives by TMA straight in the MN-major B operand layout (from HBM / L2, requested two tiles ahead)
//     dW2c += dY_i^T . G_i      dW1c^T += Xn_i^T . dH_i      dB += 1 . dH_i
// db1 = colsum(dH) is a third gradient GEMM whose A operand is a constant tile of ones (2 KB: the two 64-row halves of the
// M = 128 operand and all six 16-row k-steps alias it), so every accumulator row of dB holds db1; lane 0 is read at the end.
// With the epilogue halved the kernel would be paced by load -> gradient GEMMs -> load on a single gradient-side row tile
// (3040 clk per tile measured in that form), so the gradient GEMMs' LN(u) tiles sit in a 2-slot ring that is loaded a tile
// ahead; dY has one buffer, released by the dW2 GEMM that is issued first.
// Algorithmic HBM bytes: the dH tensor once (2 M C bytes).
constexpr int kNSH = 2;         // dH ring slots
constexpr int kNSX = 3;         // LN(u) ring slots (a tile is fetched once and read by the recompute AND the dW1 GEMM)

template <int DP>
struct CfgD {
  static constexpr int kPanel = kRows * 128;            // [96 rows][64 d or c] SW128 panel
  static constexpr int kTile = (DP / 64) * kPanel;      // LN(u) or dY row tile
  static constexpr int kW1Panel = kCc * 128;            // [128 c][64 d]
  static constexpr int kW1Bytes = (DP / 64) * kW1Panel;
  static constexpr int kGPanel = kPanel;                // [96 rows][64 c] (epilogue-written G)
  static constexpr int kGBytes = 2 * kGPanel;
  static constexpr int kDhBytes = 2 * kPanel;           // [96 rows][128 c] as two 64-channel panels
  static constexpr int kAux = 768;                      // barriers (< 256 B) + b1 of the chunk (512 B)
  static constexpr int kSmem = (1 + kNSX) * kTile + kW1Bytes + 2 * kGBytes + kNSH * kDhBytes + kAux + 1024;
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
  static constexpr int kTmemCols = 512;
  static constexpr int kColW1 = 0, kColW2 = 128, kColH = 256;   // H: two buffers of 128 columns
};

template <int DP, bool kDrop>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_dh_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmDH, const WgParams p) {
  using C = CfgD<DP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sXg = smem;                                 // [kNSX] LN(u) tiles: K-major A of the recompute GEMM, MN-major A of dW1
  uint8_t* sDY = sXg + kNSX * C::kTile;                // dY tile of the dW2 GEMM
  uint8_t* sW1 = sDY + C::kTile;
  uint8_t* sG = sW1 + C::kW1Bytes;
  uint8_t* sDH = sG + 2 * C::kGBytes;                  // [kNSH] dH tiles (TMA); sG: [2] G tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDH + kNSH * C::kDhBytes);
  uint64_t* wfull = bars;             // [1]   W1c landed
  uint64_t* xgfull = wfull + 1;       // [kNSX]
  uint64_t* xgempty = xgfull + kNSX;  // [kNSX] recompute GEMM AND dW1 GEMM done with the tile (two arrivals)
  uint64_t* dyfull = xgempty + kNSX;  // [1]
  uint64_t* dyempty = dyfull + 1;     // [1]   dW2 GEMM done with the dY tile
  uint64_t* hfull = dyempty + 1;      // [2]   H accumulator i & 1 ready -> epilogue
  uint64_t* hempty = hfull + 2;       // [2]   epilogue has it in registers -> MMA
  uint64_t* gfull = hempty + 2;       // [2]   epilogue wrote sG[i & 1] -> MMA
  uint64_t* gempty = gfull + 2;       // [2]   dW2 GEMM done with sG[i & 1] -> epilogue
  uint64_t* dhfull = gempty + 2;      // [kNSH] dH tile landed
  uint64_t* dhempty = dhfull + kNSH;  // [kNSH] dW1 GEMM AND the db1 warp done with the dH tile (two arrivals) -> TMA
  uint64_t* accfull = dhempty + kNSH; // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
  float* sB1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [128] b1 of the chunk (kAux = 256 + 512)

  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = pwarp & 3;             // TMEM lane quadrant (= scheduler)
  const int grp = pwarp >> 2;          // column group
  const bool is_tma = pwarp == 3, is_mma_hg = pwarp == 7, is_mma_wg = pwarp == 11, is_db = pwarp == 15;
  const int c0 = blockIdx.x * kCc;
  const int t_lo = static_cast<int>(static_cast<long long>(p.ntiles) * blockIdx.y / p.R);
  const int t_hi = static_cast<int>(static_cast<long long>(p.ntiles) * (blockIdx.y + 1) / p.R);
  const int nt = t_hi - t_lo;
  if (nt <= 0) return;   // uniform for the CTA

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int i = 0; i < kNSX; ++i) { mbar_init(&xgfull[i], 1); mbar_init(&xgempty[i], 2); }
    mbar_init(dyfull, 1); mbar_init(dyempty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], kLiveThreads); }
    for (int i = 0; i < 2; ++i) { mbar_init(&gfull[i], kLiveThreads); mbar_init(&gempty[i], 1); }
    for (int i = 0; i < kNSH; ++i) { mbar_init(&dhfull[i], 1); mbar_init(&dhempty[i], 2); }
    mbar_init(accfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmDH);
  }
  if (pwarp == 0) tmem_alloc(tmem_slot, C::kTmemCols);
  if (threadIdx.x < kCc) sB1[threadIdx.x] = (c0 + threadIdx.x < p.C) ? p.b1[c0 + threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (is_tma) {
    // ---- the producer: W1c once, then per tile three loads, issued in the order in which their buffers are released:
    //   dY(i)      one buffer, freed by the dW2 GEMM of tile i - 1 (issued first, right after that tile's epilogue)
    //   LN(u)(i+2) fetched ONCE into a 3-slot ring and read twice - K-major by the recompute GEMM of its tile (which runs a tile
    //              ahead, under the previous epilogue) and MN-major by the dW1 GEMM; free when both GEMMs of tile i - 1 are done
    //   dH(i+1)    2-slot ring from HBM / L2 (evict_first), free when the dW1 GEMM and the db1 warp are done with tile i - 1
    const int zc = c0 / 64;
    const uint64_t dh_policy = l2_policy_evict_first();
    auto load_xg = [&](int i) {
      const int sx = i % kNSX;
      mbar_wait(&xgempty[sx], ((i / kNSX) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&xgfull[sx], C::kTile);
#pragma unroll
        for (int pnl = 0; pnl < DP / 64; ++pnl)
          tma_load_2d(sXg + sx * C::kTile + pnl * C::kPanel, &tmX, &xgfull[sx], pnl * 64, (t_lo + i) * kRows);
      }
      __syncwarp();
    };
    auto load_dh = [&](int i) {
      const int sh = i % kNSH;
      mbar_wait(&dhempty[sh], ((i / kNSH) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&dhfull[sh], C::kDhBytes);
        const int rr = (t_lo + i) * kRows;
        if (p.l2_hint) {
          tma_load_3d_hint(sDH + sh * C::kDhBytes, &tmDH, &dhfull[sh], 0, rr, zc, dh_policy);
          tma_load_3d_hint(sDH + sh * C::kDhBytes + C::kPanel, &tmDH, &dhfull[sh], 0, rr, zc + 1, dh_policy);
        } else {
          tma_load_3d(sDH + sh * C::kDhBytes, &tmDH, &dhfull[sh], 0, rr, zc);
          tma_load_3d(sDH + sh * C::kDhBytes + C::kPanel, &tmDH, &dhfull[sh], 0, rr, zc + 1);
        }
      }
      __syncwarp();
    };
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, C::kW1Bytes);
#pragma unroll
      for (int pnl = 0; pnl < DP / 64; ++pnl)
#pragma unroll
        for (int h = 0; h < 2; ++h)   // [64 c][64 d] boxes: d panel pnl, channel half h
          tma_load_2d(sW1 + pnl * C::kW1Panel + h * (64 * 128), &tmW1, wfull, pnl * 64, c0 + h * 64);
    }
    __syncwarp();
    load_xg(0);
    if (nt > 1) load_xg(1);
    load_dh(0);
    for (int i = 0; i < nt; ++i) {
      mbar_wait(dyempty, (i & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(dyfull, C::kTile);
#pragma unroll
        for (int pnl = 0; pnl < DP / 64; ++pnl) tma_load_2d(sDY + pnl * C::kPanel, &tmDY, dyfull, pnl * 64, (t_lo + i) * kRows);
      }
      __syncwarp();
      if (i + 2 < nt) load_xg(i + 2);
      if (i + 1 < nt) load_dh(i + 1);
    }
  } else if (is_db) {
    // ---- db1 = colsum(dH) on the CUDA cores of the scheduler that has no epilogue warp (quadrant 3): lane l owns channels
    // 4 l .. 4 l + 3 of the chunk and walks the 96 rows of every dH tile in shared memory (8-byte reads; the tile is in the
    // 128-byte swizzle: 16-byte chunk j of row r sits at j ^ (r & 7)).  It replaced a third gradient GEMM against a tile of
    // ones, which cost 6 of the 26 MMAs per tile and the 128 tensor-memory columns the second H accumulator now uses.
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int jch = (lane & 15) >> 1;
    uint32_t off[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) off[k] = static_cast<uint32_t>(k * 128 + ((jch ^ k) << 4));
    const uint32_t lane_base = smem_u32(sDH) + static_cast<uint32_t>((lane >> 4) * C::kPanel + (lane & 1) * 8);
    for (int i = 0; i < nt; ++i) {
      const int sh = i % kNSH;
      mbar_wait(&dhfull[sh], (i / kNSH) & 1);
      const uint32_t base = lane_base + static_cast<uint32_t>(sh * C::kDhBytes);
#pragma unroll
      for (int r8 = 0; r8 < kRows / 8; ++r8)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint32_t v0, v1;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(base + r8 * 1024 + off[k]));
          acc[0] += __uint_as_float(v0 << 16); acc[1] += __uint_as_float(v0 & 0xffff0000u);
          acc[2] += __uint_as_float(v1 << 16); acc[3] += __uint_as_float(v1 & 0xffff0000u);
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(&dhempty[sh]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (c0 + 4 * lane + k < p.C) atomicAdd(p.db1 + c0 + 4 * lane + k, acc[k]);
  } else if (is_mma_hg) {
    // ---- recompute issuer: H = Xn_i . W1c^T (N = 128) as soon as the epilogue has tile i - 1's accumulator in registers
    constexpr uint32_t idescH = umma_idesc_bf16(kMmaM, kCc, 0, 0);   // A row tile K-major,  B = W1c K-major (N = 128 rows)
    const uint64_t xk0 = umma_desc_sw128(smem_u32(sXg), 16, 1024);
    const uint64_t w1d = umma_desc_sw128(smem_u32(sW1), 16, 1024);
    mbar_wait(wfull, 0);
    for (int i = 0; i < nt; ++i) {
      const int sx = i % kNSX;
      M2_WTR(4 * i + 0, 1, i);
      mbar_wait2(&xgfull[sx], (i / kNSX) & 1, &hempty[i & 1], ((i >> 1) & 1) ^ 1);
      M2_WTR(4 * i + 1, 2, i);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xa = xk0 + static_cast<uint64_t>((sx * C::kTile) >> 4);
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16(tmem_base + C::kColH + (i & 1) * kCc, xa + (((kk >> 2) * C::kPanel + (kk & 3) * 32) >> 4),
                    w1d + (((kk >> 2) * C::kW1Panel + (kk & 3) * 32) >> 4), idescH, kk > 0 ? 1u : 0u);
        umma_commit(&xgempty[sx]);
        umma_commit(&hfull[i & 1]);
        M2_WTR(4 * i + 2, 3, i);
      }
      __syncwarp();
    }
  } else if (is_mma_wg) {
    // ---- gradient issuer (contraction over the 96 rows): dW2c += dY_i^T . G_i ; dW1c^T += Xn_i^T . dH_i
    constexpr uint32_t idescW = umma_idesc_bf16(kMmaM, kCc, 1, 1);   // A row tile MN-major (M = d), B = sDH / sG MN-major
    constexpr uint32_t kLboA = DP == 128 ? C::kPanel : 0;            // DP = 64: M rows 64..127 alias the only panel
    const uint64_t xg0 = umma_desc_sw128(smem_u32(sXg), kLboA, 1024);
    const uint64_t ya = umma_desc_sw128(smem_u32(sDY), kLboA, 1024);
    const uint64_t gd = umma_desc_sw128(smem_u32(sG), C::kGPanel, 1024);
    const uint64_t dh0 = umma_desc_sw128(smem_u32(sDH), C::kPanel, 1024);
    for (int i = 0; i < nt; ++i) {
      const int sh = i % kNSH, sx = i % kNSX;
      const uint32_t acc = i > 0 ? 1u : 0u;
      M2_WTR(200 + 4 * i + 0, 4, i);
      mbar_wait2(&gfull[i & 1], (i >> 1) & 1, dyfull, i & 1);
      M2_WTR(200 + 4 * i + 1, 5, i);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kRows / 16; ++kk)   // 16 rows per step = 2048 B in both operands
          umma_bf16(tmem_base + C::kColW2, ya + ((kk * 2048) >> 4), gd + (((i & 1) * C::kGBytes + kk * 2048) >> 4), idescW,
                    (kk > 0) ? 1u : acc);
        umma_commit(&gempty[i & 1]);
        umma_commit(dyempty);
      }
      __syncwarp();
      mbar_wait2(&xgfull[sx], (i / kNSX) & 1, &dhfull[sh], (i / kNSH) & 1);
      M2_WTR(200 + 4 * i + 2, 11, i);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xa = xg0 + static_cast<uint64_t>((sx * C::kTile) >> 4);
        const uint64_t dhd = dh0 + static_cast<uint64_t>((sh * C::kDhBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < kRows / 16; ++kk)
          umma_bf16(tmem_base + C::kColW1, xa + ((kk * 2048) >> 4), dhd + ((kk * 2048) >> 4), idescW, (kk > 0) ? 1u : acc);
        umma_commit(&xgempty[sx]);
        umma_commit(&dhempty[sh]);
        M2_WTR(200 + 4 * i + 3, 6, i);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accfull);
    __syncwarp();
  } else if (q < 3) {
    // ---- epilogue: thread = row of the tile (TMEM lane), group g = columns [32 g, 32 g + 32) in two 16-column pieces
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int cg = c0 + grp * 32;
    const uint32_t bias_addr = smem_u32(sB1 + grp * 32);
    const float hs = kDrop ? 0.5f * p.dh.scale : 0.5f;          // dropout scale folded into the GELU
    uint8_t* gdst0 = sG + (grp >> 1) * C::kGPanel;              // [96 rows][64 c] SW128 panel; this group: 16-byte chunks
    const int chunk0 = (grp & 1) * 4;                           //   4 (grp & 1) .. 4 (grp & 1) + 3 of the row
    const uint32_t dkey = drop_key(p.dh);
    bool ready = false;                              // hfull of the tile already observed by an early probe
    for (int i = 0; i < nt; ++i) {
      if (pwarp == 0) M2_WTR(400 + 4 * i + 0, 7, i);
      uint8_t* gdst = gdst0 + (i & 1) * C::kGBytes;
      // two G buffers: the dW2 GEMM of tile i - 2 released this one long ago (with ONE buffer the loop gfull -> issuer wakes ->
      // dW2 -> commit -> second half of the next epilogue -> gfull set the pace: 2100 clk per tile, profiles/r02_trace_wgrad_dh.log)
      if (!ready) mbar_wait(&hfull[i & 1], (i >> 1) & 1);
      __syncwarp();
      if (pwarp == 0) M2_WTR(400 + 4 * i + 1, 8, i);
      tc_fence_after();
      const uint32_t hin = ((static_cast<uint32_t>((t_lo + i) * kRows + r) * static_cast<uint32_t>(p.ldh) + static_cast<uint32_t>(cg)) >> 2) * kDropGolden + dkey;
      uint32_t hA[16], hB[16];
      tmem_ld16(tmem_base + C::kColH + (i & 1) * kCc + lane_addr + grp * 32, hA);
      tmem_ld16(tmem_base + C::kColH + (i & 1) * kCc + lane_addr + grp * 32 + 16, hB);
      // probed while the accumulator loads are in flight: an mbarrier probe has a ~200-clk round trip even when the phase
      // completed long ago, and the twelve epilogue warps run in lockstep (nothing else would hide it)
      const bool gfree = mbar_probe(&gempty[i & 1], ((i >> 1) & 1) ^ 1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&hempty[i & 1]);                   // two H accumulators: the recompute GEMM runs a whole tile ahead
      if (pwarp == 0) M2_WTR(600 + 2 * i + 1, 9, i);
#pragma unroll
      for (int pc = 0; pc < 2; ++pc) {
        uint32_t (&h)[16] = pc ? hB : hA;
        float bias[16];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
              : "=f"(bias[4 * e]), "=f"(bias[4 * e + 1]), "=f"(bias[4 * e + 2]), "=f"(bias[4 * e + 3])
              : "r"(bias_addr + (pc * 16 + 4 * e) * 4));
        uint32_t gp[8];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {   // quads of channels: one mask hash each
          uint32_t flags = 0;
          if (kDrop) flags = drop_flags_from_hash_input(p.dh, hin + static_cast<uint32_t>(pc * 4 + qd) * kDropGolden);
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int e = qd * 2 + e2;
            const float2 gv = gelu2(__fadd2_rn(make_float2(__uint_as_float(h[2 * e]), __uint_as_float(h[2 * e + 1])),
                                               make_float2(bias[2 * e], bias[2 * e + 1])), hs);
            gp[e] = pack_bf16(gv.x, gv.y);
            if (kDrop) gp[e] &= e2 ? drop_mask_bf16x2<1>(flags) : drop_mask_bf16x2<0>(flags);
          }
        }
        // (keeping all 32 values in registers and storing them after the gempty wait at the END of the tile measured
        // slower, 60.2 vs 56.4 us per launch: the kernel is bound by shared-memory bandwidth, not by this wait)
        if (pc == 0 && pwarp == 0) M2_WTR(600 + 2 * i, 13, i);
        if (pc == 0 && !gfree) mbar_wait(&gempty[i & 1], ((i >> 1) & 1) ^ 1);   // the dW2 GEMM of tile i - 2 has consumed the buffer
        if (pc == 0 && pwarp == 0) M2_WTR(400 + 4 * i + 2, 12, i);
        if (pc == 1) ready = (i + 1 < nt) && mbar_probe(&hfull[(i + 1) & 1], ((i + 1) >> 1) & 1);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          *reinterpret_cast<uint4*>(gdst + sw128_offset(r, chunk0 + pc * 2 + k)) = make_uint4(gp[4 * k], gp[4 * k + 1], gp[4 * k + 2], gp[4 * k + 3]);
      }
      fence_proxy_async();
      mbar_arrive(&gfull[i & 1]);
      if (pwarp == 0) M2_WTR(400 + 4 * i + 3, 10, i);
    }
  }
  // ---- all 16 warps: accumulators -> global.  TMEM lane = d.  groups 0 / 1: dW1^T columns [0,64) / [64,128) -> dw1[c][d]
  // (lanes contiguous in d: coalesced reductions); groups 2 / 3: dW2 columns likewise -> dw2[d][c]
  __syncwarp();
  mbar_wait(accfull, 0);
  tc_fence_after();
  {
    const int d = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool first = grp < 2;
    const int cbase = (grp & 1) * 64;
#pragma unroll 1
    for (int cb = cbase; cb < cbase + 64; cb += 32) {
      uint32_t a[32];
      tmem_ld32(tmem_base + (first ? C::kColW1 : C::kColW2) + lane_addr + cb, a);
      tmem_ld_wait();
      if (d < p.D) {
        if (first) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (pwarp == 0) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int DP, bool kDrop>
int launch_wd(const CUtensorMap& tx, const CUtensorMap& ty, const CUtensorMap& t1, const CUtensorMap& tdh, const WgParams& p,
              cudaStream_t s) {
  auto kern = wgrad_dh_kernel<DP, kDrop>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgD<DP>::kSmem) != cudaSuccess) return M2_ERR_LAUNCH;
    configured = true;
  }
  LaunchScope scope("wgrad_dh", s);
  dim3 grid(ceil_div(p.C, kCc), p.R);
  kern<<<grid, kThreads, CfgD<DP>::kSmem, s>>>(tx, ty, t1, tdh, p);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace

#ifdef M2_TRACE
int wgrad_trace_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_wtrace, sizeof(long long) * n) == cudaSuccess ? 0 : 1;
}
#endif

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
  rc = make_tmap_bf16(&t1, w1b, C, D, D, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&t2, w2b, D, ldw2, ldw2, DP, 64);
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

// Generation 4: dH (bf16, chunk-major [ceil(C / 64)][M][64], written by chain_bwd_ts through TMA stores) is an input; only G is
// recomputed.  ldh is the row stride of the dropout mask (up8(C)), not of the dH buffer.
int wgrad_dh(const void* xn_b, const void* dy_b, const void* dh_b, int ldh, const void* w1b, const float* b1, float* dw1,
             float* db1, float* dw2, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s) {
  if (!chain_fwd_ts_supported(D) || ldh % 8 || ldh < C) return M2_ERR_ARG;
  const int DP = D <= 64 ? 64 : 128;
  CUtensorMap tx, ty, t1, tdh;
  int rc = make_tmap_bf16(&tx, xn_b, M, D, D, kRows, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&ty, dy_b, M, D, D, kRows, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&t1, w1b, C, D, D, 64, 64);
  if (rc) return rc;
  rc = make_tmap_store3d(&tdh, dh_b, 2, ceil_div(C, 64), M, 64, 64, static_cast<uint64_t>(M) * 64, kRows, 64);   // chunk-major
  if (rc) return rc;
  WgParams p = {};
  p.b1 = b1; p.dw1 = dw1; p.dw2 = dw2; p.db1 = db1;
  p.M = M; p.D = D; p.C = C; p.ldh = ldh;
  p.ntiles = ceil_div(M, kRows);
  const int nch = ceil_div(C, kCc);
  int R = 148 / nch;
  if (R < 1) R = 1;
  if (R > p.ntiles) R = p.ntiles;
  p.R = R;
  p.l2_hint = dh_l2_hint();
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden);
  const bool drop = p.dh.thresh != 0;
  if (DP == 64) return drop ? launch_wd<64, true>(tx, ty, t1, tdh, p, s) : launch_wd<64, false>(tx, ty, t1, tdh, p, s);
  return drop ? launch_wd<128, true>(tx, ty, t1, tdh, p, s) : launch_wd<128, false>(tx, ty, t1, tdh, p, s);
}

}  // namespace m2
