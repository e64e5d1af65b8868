// Fused channel-mixing chains, generation 2: every row-tile operand lives in TENSOR MEMORY.
//
// Reference arithmetic: MixerBlock.channel_mix, modules/mixer.py:37-40,45
//     y = u + Drop(W2 . Drop(GELU(W1 . LN(u) + b1)) + b2)     per token row (M = B*N rows, D hidden, C channel_dim)
//
// Why a second generation: in chain.cu the A operands (LN(u), G) are read from shared memory by every tcgen05.mma.
// Per 64-channel chunk that is 32 KB (sX) + 16 KB (sG) of operand reads + 16 KB of epilogue stores on top of the
// 32 KB of weight tiles, i.e. ~1000 cycles of the 128 B/clk shared-memory pipe for ~512 cycles of tensor work, and
// the epilogue could not start a chunk before GEMM2 had drained its one staging tile.  Here
//   * LN(u) is written ONCE per CTA into TMEM (bf16, DP/2 columns) and is the TMEM A operand of every GEMM1,
//   * the epilogue writes G = GELU(H + b1) back to TMEM with tcgen05.st (4 rotating 32-column buffers) and GEMM2
//     consumes it as a TMEM A operand (.kind::f16 "TS" form): shared memory only carries the weight tiles,
//   * H has 5 accumulator buffers and GEMM1 runs 5 chunks ahead of the epilogue; G(j) overwrites the first 32 columns
//     of its own H buffer (a thread only touches its own lane; the tensor pipe executes GEMM2(j) before the GEMM1 of
//     chunk j + 5 that reuses the buffer), so no "empty" barriers exist at all,
//   * TMEM reads are the scarce resource of the epilogue (tcgen05.ld moves 16 B/clk per lane quadrant, 512 clk per
//     [128 x 64] fp32 chunk): the accumulator is read in 16-column pieces, the load of piece t + 1 is in flight while
//     piece t goes through the GELU,
//   * the output tile is staged through the (then idle) weight ring so that u is read and y written coalesced.
//
// TMEM map (512 columns): [0,DP) Y accumulator | [DP, DP+DP/2) LN(u) bf16 | 5 x 64 H accumulators (G aliased).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"

namespace m2 {
namespace {

constexpr int kRows = 128;      // token rows per CTA (UMMA M)
constexpr int kCc = 64;         // channels per chunk
// forward: 96 + 128 kG threads (TMA warp, two MMA-issuing warps, kG epilogue groups of 4 warps), see chain_fwd_ts_kernel
// backward: FOUR epilogue groups (16 warps, 4 per scheduler).  With two, the GELU' epilogue ran at 0.46 IPC per
// scheduler - dependent FMA chains and MUFU latency, no pipe above 40 % (profiles/r01_ncu_final.md): latency bound,
// so the cure is more warps in flight, each on a quarter (16 columns) of the chunk.
constexpr int kGroupsB = 4;
constexpr int kThreadsB = 96 + 128 * kGroupsB;
constexpr int kBarBytes = 512;

template <int DP>
struct CfgT {
  static constexpr int kW1Bytes = kCc * DP * 2;      // [64 c-rows][DP d]
  static constexpr int kW2Bytes = DP * kCc * 2;      // [DP d-rows][64 c]
  static constexpr int S1 = 4, S2 = 4;               // weight ring depths
  static constexpr int NB = 5;                       // H accumulator buffers (GEMM1 lookahead); G(j) aliases H(j)[0:32)
  static constexpr int kRingBytes = S1 * kW1Bytes + S2 * kW2Bytes;
  static constexpr int kXPitch = DP * 2 + 16;        // bf16 LN(u) staging row pitch (conflict-free row reads)
  static constexpr int kXStage = kRows * kXPitch;
  static constexpr int kOPitch = DP * 4 + 16;        // fp32 output staging row pitch (aliases the weight ring)
  static_assert(kRows * kOPitch <= kRingBytes, "output staging must fit in the weight ring");
  static constexpr int kSmem = kRingBytes + kXStage + kBarBytes + 1024;
  static constexpr int kMaxBias = 32 * 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kColY = 0;
  static constexpr int kColX = DP;
  static constexpr int kColH = DP + DP / 2;
  static_assert(kColH + NB * kCc <= 512, "TMEM budget");
};

// Optional in-kernel timeline (debug builds: -DM2_TRACE): CTA 0 records (tag, chunk, clock) triples of its MMA warp
// and first epilogue warp; tools/trace_chain.py prints them.  Compiled out by default.
#ifdef M2_TRACE
__device__ long long g_trace[4096];
__device__ __forceinline__ void trace_evt(int slot, int tag, int j) {
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && slot < 1360) {
    g_trace[3 * slot] = tag; g_trace[3 * slot + 1] = j; g_trace[3 * slot + 2] = clock64();
  }
}
#define M2_TR(slot, tag, j) trace_evt(slot, tag, j)
#else
#define M2_TR(slot, tag, j)
#endif

struct TsParams {
  const float* u;        // [M][D] block input (pre-LN residual stream)
  const float* ln_w; const float* ln_b;
  const float* b1;       // [C]
  const float* b2;       // [D]
  float* y;              // fwd: [M][D]
  // backward (dgrad kernel)
  const float* dy;       // [M][D]
  float* du;             // [M][D]  = dy + LayerNorm'(dXn)
  float* dln_w; float* dln_b; float* db2;   // [D] each, accumulated with atomics
  __nv_bfloat16* xn_b;   // out: LN(u) bf16 [M][D]           (operand of the weight-gradient kernels)
  __nv_bfloat16* dy_b;   // out: dY (masked) bf16 [M][D]
  __nv_bfloat16* g_b;    // optional out: G  bf16 [M][ldh]   (nullptr: the weight-gradient kernel recomputes it)
  __nv_bfloat16* dh_b;   // optional out: dH bf16 [M][ldh]
  int M, D, C, ldh;
  int bias_smem;
  int l2_hint;           // dH stores carry an L2 evict_first policy (M2B200_DH_L2HINT, default on)
  Drop dh, dout;
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// registers -> TMEM: this thread's lane, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// explicit ld.shared (the bias pointer is derived from an aligned-up dynamic smem base: ptxas falls back to generic LD)
__device__ __forceinline__ float4 lds_f4(const float* p) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// LayerNorm the tile's rows (warp per row, kB rows in flight) into a padded row-major bf16 staging tile.
template <int DP, int kB = 13>   // kB rows in flight per warp: one batch must cover the warp's share of the tile
__device__ __forceinline__ void ln_rows_to_stage(const TsParams& p, int m0, uint8_t* stage, float* s_mean = nullptr,
                                                 float* s_rstd = nullptr, __nv_bfloat16* xn_b = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kV = DP / 128 > 0 ? DP / 128 : 1;   // float4 per lane per row
  const int kW = blockDim.x >> 5;
  static_assert(kV == 1, "LayerNorm prologue layout");
  float4 gw[kV], gb[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int c = (i * 32 + lane) * 4;
    gw[i] = c < p.D ? *reinterpret_cast<const float4*>(p.ln_w + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    gb[i] = c < p.D ? *reinterpret_cast<const float4*>(p.ln_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_d = 1.f / static_cast<float>(p.D);
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
    // Statistics of all kB rows together: the butterfly steps of different rows are independent, so their shuffle
    // latencies overlap (row after row this prologue took 6900 clk of a 62 k clk kernel, profiles/r01_trace_fwd_*.log).
    float mean[kB], rstd[kB];
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kV; ++i) s += v[b][i].x + v[b][i].y + v[b][i].z + v[b][i].w;
      mean[b] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int b = 0; b < kB; ++b) mean[b] += __shfl_xor_sync(0xffffffffu, mean[b], o);
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      mean[b] *= inv_d;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < p.D) {
          const float a = v[b][i].x - mean[b], bb = v[b][i].y - mean[b], cc = v[b][i].z - mean[b], d = v[b][i].w - mean[b];
          ss += a * a + bb * bb + cc * cc + d * d;
        }
      }
      rstd[b] = ss;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int b = 0; b < kB; ++b) rstd[b] += __shfl_xor_sync(0xffffffffu, rstd[b], o);
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      const int r = r0 + b, row = m0 + r;
      if (r < kRows) {
        const float mu = mean[b], rs = rsqrtf(rstd[b] * inv_d + kLnEps);
        if (s_mean && lane == 0) { s_mean[r] = mu; s_rstd[r] = rs; }
#pragma unroll
        for (int i = 0; i < kV; ++i) {
          const int c = (i * 32 + lane) * 4;
          if (c < DP) {
            uint2 o = make_uint2(0u, 0u);
            if (row < p.M && c < p.D) {
              o.x = pack_bf16((v[b][i].x - mu) * rs * gw[i].x + gb[i].x, (v[b][i].y - mu) * rs * gw[i].y + gb[i].y);
              o.y = pack_bf16((v[b][i].z - mu) * rs * gw[i].z + gb[i].z, (v[b][i].w - mu) * rs * gw[i].w + gb[i].w);
              if (xn_b) *reinterpret_cast<uint2*>(xn_b + static_cast<long long>(row) * p.D + c) = o;
            }
            *reinterpret_cast<uint2*>(stage + r * CfgT<DP>::kXPitch + c * 2) = o;
          }
        }
      }
    }
  }
}

// b1 -> shared memory (zero padded to whole chunks), 16-byte loads.  Split into a load half and a store half so that the
// loads are in flight during the LayerNorm prologue (back to back they cost 640 clk of load latency,
// profiles/r01_trace_fwd_*.log).  One round covers 16 * blockDim floats (the launchers check it).
struct BiasRegs { float4 v[4]; };
__device__ __forceinline__ BiasRegs bias_load(const float* __restrict__ b1, int C, int n_pad) {
  BiasRegs br;
  const int nv = n_pad >> 2;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * blockDim.x;
    br.v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nv) {
      const int c = i * 4;
      if (c + 3 < C && (reinterpret_cast<uintptr_t>(b1) & 15) == 0) br.v[k] = *reinterpret_cast<const float4*>(b1 + c);
      else {
        if (c < C) br.v[k].x = b1[c];
        if (c + 1 < C) br.v[k].y = b1[c + 1];
        if (c + 2 < C) br.v[k].z = b1[c + 2];
        if (c + 3 < C) br.v[k].w = b1[c + 3];
      }
    }
  }
  return br;
}
__device__ __forceinline__ void bias_store(const BiasRegs& br, int n_pad, float* sBias) {
  const int nv = n_pad >> 2;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * blockDim.x;
    if (i < nv) *reinterpret_cast<float4*>(sBias + i * 4) = br.v[k];
  }
}

// One epilogue thread copies (its row) x (32 TMEM columns = 64 bf16) of a padded row-major bf16 staging tile into TMEM.
__device__ __forceinline__ void stage_row_to_tmem(const uint8_t* stage_row, uint32_t taddr) {
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 q = *reinterpret_cast<const uint4*>(stage_row + i * 16);
    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
  }
  tmem_st32(taddr, v);
}

template <int DP>
__device__ __forceinline__ void load_w1(uint8_t* slot, const CUtensorMap* tmW1, uint64_t* bar, int c0) {
  mbar_arrive_expect_tx(bar, CfgT<DP>::kW1Bytes);
#pragma unroll
  for (int pnl = 0; pnl < DP / 64; ++pnl)          // [64 c-rows][64 d] panels
    tma_load_2d(slot + pnl * (kCc * 128), tmW1, bar, pnl * 64, c0);
}
template <int DP>
__device__ __forceinline__ void load_w2(uint8_t* slot, const CUtensorMap* tmW2, uint64_t* bar, int c0) {
  mbar_arrive_expect_tx(bar, CfgT<DP>::kW2Bytes);
  tma_load_2d(slot, tmW2, bar, c0, 0);              // [DP d-rows][64 c]
}

// ============================================================================================ forward
// kG epilogue groups of 4 warps (group g owns chunks j = g mod kG).  With two groups (2 warps per scheduler) the epilogue
// issued 34 % of the time with no pipe above 30 % (profiles/r01_ncu_final.md): latency bound, like the backward before
// it went to four groups.
template <int DP, bool kDrop, int kG>
__global__ void __launch_bounds__(96 + 128 * kG, 1)
chain_fwd_ts_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const TsParams p) {
  using C = CfgT<DP>;
  constexpr int S1 = C::S1, S2 = C::S2, NB = C::NB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + S1 * C::kW1Bytes;
  uint8_t* sStage = sW2 + S2 * C::kW2Bytes;          // LN(u) bf16 staging
  uint8_t* sOut = smem;                              // fp32 output staging (aliases the ring once every MMA is done)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + C::kXStage);
  uint64_t* w1full = bars;            // [S1]  TMA -> MMA
  uint64_t* w1empty = w1full + S1;    // [S1]  GEMM1 done -> TMA
  uint64_t* w2full = w1empty + S1;    // [S2]
  uint64_t* w2empty = w2full + S2;    // [S2]  GEMM2 done -> TMA
  uint64_t* hfull = w2empty + S2;     // [NB]  GEMM1 done -> epilogue
  uint64_t* gfull = hfull + NB;       // [NB]  epilogue wrote G (bf16) over the first 32 columns of the buffer -> MMA
  uint64_t* yfull = gfull + NB;       // [1]
  uint64_t* g2done = yfull + 1;       // [NB]  GEMM2(j) has read G(j): the GEMM1 issuer may overwrite the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g2done + NB);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);

  // Roles by LOGICAL warp id (0 = TMA producer, 1 = GEMM2 issuer, 2.. = epilogue, -1 = GEMM1 issuer).  Physically the
  // epilogue warps come first and the single-thread roles LAST: the warp scheduler prefers the highest warp id on its
  // sub-partition, and a latency-critical issuer starved behind the busy epilogue warps when it was warp 1 (724-clk gaps
  // between a commit and the next wait in the in-kernel timeline, profiles/r01_trace_dgrad.log).
  //
  // TWO issuing warps: tcgen05.mma issue blocks at the execution rate (the queue holds only a few MMAs:
  // tools/umma_probe.cu, issue time == execution time) and every wait on an mbarrier costs the issuer ~200 clk even
  // when the phase completed long ago, so ONE issuer that waits for G(j), W2(j), W1(j+NB) and then issues 12 MMAs keeps
  // the tensor pipe idle during its waits: 1378 clk per chunk for 620 clk of tensor work, the epilogue starved for H
  // (in-kernel timeline, profiles/r01_trace_fwd_*.log).  With the GEMM1s and the GEMM2s on different warps one issuer's
  // waits overlap the other's MMAs.  Each accumulator has ONE writer (Y: GEMM2 warp, H buffers: GEMM1 warp); the only
  // cross-warp hazard is the G(j) -> H(j + NB) buffer reuse, ordered by the g2done barrier (tcgen05.commit of GEMM2(j)).
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = pwarp < 4 * kG ? pwarp + 2 : (pwarp == 4 * kG + 2 ? -1 : pwarp - 4 * kG);
  const int m0 = blockIdx.x * kRows;
  const int nch = ceil_div(p.C, kCc);
  if (pwarp == 0) M2_TR(1300, 20, 0);   // kernel entry

  if (threadIdx.x == 0) {
    for (int i = 0; i < S1; ++i) { mbar_init(&w1full[i], 1); mbar_init(&w1empty[i], 1); }
    for (int i = 0; i < S2; ++i) { mbar_init(&w2full[i], 1); mbar_init(&w2empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&hfull[i], 1); mbar_init(&gfull[i], 64 * kG); mbar_init(&g2done[i], 1); }
    mbar_init(yfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  __syncthreads();   // barriers initialised before the producer's early prefetch below
  if (pwarp == 0) M2_TR(1301, 21, 0);   // barriers / TMEM allocated

  // The weight rings do not depend on the activations: start filling them before the LayerNorm prologue.
  if (warp == 0 && lane == 0) {
    for (int j = 0; j < (nch < S1 ? nch : S1); ++j) load_w1<DP>(sW1 + j * C::kW1Bytes, &tmW1, &w1full[j], j * kCc);
    for (int j = 0; j < (nch < S2 ? nch : S2); ++j) load_w2<DP>(sW2 + j * C::kW2Bytes, &tmW2, &w2full[j], j * kCc);
  }
  if (pwarp == 0) M2_TR(1308, 28, 0);
  BiasRegs br;
  if (p.bias_smem) br = bias_load(p.b1, p.C, nch * kCc);      // in flight during the LayerNorm
  ln_rows_to_stage<DP, kG == 2 ? 13 : 8>(p, m0, sStage);
  if (pwarp == 0) M2_TR(1309, 29, 0);
  if (p.bias_smem) bias_store(br, nch * kCc, sBias);          // first read by the epilogue, two block barriers later
  if (pwarp == 0) M2_TR(1310, 30, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (pwarp == 0) M2_TR(1302, 22, 0);   // LayerNorm rows staged
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tY = tmem_base + C::kColY;
  const uint32_t tX = tmem_base + C::kColX;

  // staging -> TMEM (the A operand of every GEMM1): epilogue thread = TMEM lane = tile row; DP/2 columns
  if (warp >= 2) {
    const int q = pwarp & 3, grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    for (int cb = grp; cb < DP / 64; cb += kG)   // 32 TMEM columns (= 64 bf16 = 128 B of the staging row) per step
      stage_row_to_tmem(sStage + r * C::kXPitch + cb * 128, tX + lane_addr + cb * 32);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (pwarp == 0) M2_TR(1303, 23, 0);   // LN(u) in TMEM: roles start

  if (warp == 0) {
    // Refill both rings in the order the MMA issuer frees the slots: the prologue GEMM1s free W1 slots first, then
    // iteration `it` of the issuer completes GEMM2(it) (frees a W2 slot) and GEMM1(it + NB) (frees a W1 slot).
    auto refill_w1 = [&](int x) {
      mbar_wait(&w1empty[x % S1], ((x / S1) & 1) ^ 1);
      if (elect_one()) load_w1<DP>(sW1 + (x % S1) * C::kW1Bytes, &tmW1, &w1full[x % S1], x * kCc);
      __syncwarp();
    };
    auto refill_w2 = [&](int y) {
      mbar_wait(&w2empty[y % S2], ((y / S2) & 1) ^ 1);
      if (elect_one()) load_w2<DP>(sW2 + (y % S2) * C::kW2Bytes, &tmW2, &w2full[y % S2], y * kCc);
      __syncwarp();
    };
    int x = S1;   // next W1 chunk to load
    for (; x < S1 + NB && x < nch; ++x) refill_w1(x);
    for (int it = 0; it < nch; ++it) {
      if (it + S2 < nch) refill_w2(it + S2);
      if (x < nch) { refill_w1(x); ++x; }
    }
  } else if (warp == -1) {
    // ---- GEMM1 issuer: Hacc[j % NB] = LN(u) . W1_j^T (A from TMEM), NB chunks ahead of the epilogue.
    // The whole warp walks the loop (waits are warp-wide), one elected lane issues: see elect_one().
    constexpr uint32_t idesc1 = umma_idesc_bf16(kRows, kCc, 0, 0);
    const uint64_t w1_desc0 = umma_desc_sw128(smem_u32(sW1), 16, 1024);
    for (int j = 0; j < nch; ++j) {
      const int s = j % S1, hb = j % NB;
      // GEMM2(j - NB) must have consumed the G that aliases this buffer
      if (j >= NB) mbar_wait2(&w1full[s], (j / S1) & 1, &g2done[hb], ((j / NB) - 1) & 1);
      else mbar_wait(&w1full[s], (j / S1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bd = w1_desc0 + static_cast<uint64_t>((s * C::kW1Bytes) >> 4);
        const uint32_t tH = tmem_base + C::kColH + hb * kCc;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16_ts(tH, tX + kk * 8, bd + (((kk >> 2) * (kCc * 128) + (kk & 3) * 32) >> 4), idesc1, kk > 0 ? 1u : 0u);
        umma_commit(&w1empty[s]);
        umma_commit(&hfull[hb]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- GEMM2 issuer: Yacc += G_j . W2_j^T (A from TMEM)
    constexpr uint32_t idesc2 = umma_idesc_bf16(kRows, DP, 0, 0);
    const uint64_t w2_desc0 = umma_desc_sw128(smem_u32(sW2), 16, 1024);
    for (int j = 0; j < nch; ++j) {
      const int s = j % S2, gb = j % NB;
      M2_TR(4 * j + 0, 1, j);                         // issuer: about to wait
      mbar_wait2(&gfull[gb], (j / NB) & 1, &w2full[s], (j / S2) & 1);
      M2_TR(4 * j + 2, 3, j);                         // issuer: G(j) and the weights there
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bd = w2_desc0 + static_cast<uint64_t>((s * C::kW2Bytes) >> 4);
        const uint32_t tG = tmem_base + C::kColH + gb * kCc;
#pragma unroll
        for (int kk = 0; kk < kCc / 16; ++kk)   // A: k-step kk of G(j); kG = 4: the halves sit at columns 0 and 32 of the buffer
          umma_bf16_ts(tY, tG + (kG == 4 ? (kk >> 1) * 32 + (kk & 1) * 8 : kk * 8), bd + ((kk * 32) >> 4), idesc2,
                       (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&w2empty[s]);
        umma_commit(&g2done[gb]);
        M2_TR(4 * j + 3, 4, j);                       // issuer: burst issued
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(yfull);
    __syncwarp();
  } else {
    const int q = pwarp & 3;               // TMEM lane quadrant this warp may access (physical warp id % 4)
    const int grp = (warp - 2) >> 2;       // epilogue group: chunks j = grp (mod kG)
    const int r = q * 32 + lane;           // row inside the tile == TMEM lane
    const int row = m0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    // Software pipeline over 16-column pieces: the tcgen05.ld of piece t + 1 is in flight while piece t is computed.
    auto ld_piece = [&](int j, int pc, uint32_t (&dst)[16]) {
      tmem_ld16(tmem_base + C::kColH + lane_addr + (j % NB) * kCc + pc * 16, dst);
    };
    const float hs = kDrop ? 0.5f * p.dh.scale : 0.5f;   // dropout scale folded into the GELU
    // dropout: hash input of the quad at channel 0 of this thread's row (quad index = (row * ldh + c) / 4)
    const uint32_t hrow = (static_cast<uint32_t>(row) * static_cast<uint32_t>(p.ldh) >> 2) * kDropGolden + drop_key(p.dh);
    auto gelu_piece = [&](const uint32_t (&h)[16], int c, uint32_t* gout) {   // 16 columns starting at channel c
      const uint32_t hin = hrow + static_cast<uint32_t>(c >> 2) * kDropGolden;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float b[8];
        if (p.bias_smem) {
          const float4 b0 = lds_f4(sBias + c + hh * 8);
          const float4 b1v = lds_f4(sBias + c + hh * 8 + 4);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1v.x; b[5] = b1v.y; b[6] = b1v.z; b[7] = b1v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) b[e] = (c + hh * 8 + e < p.C) ? __ldg(p.b1 + c + hh * 8 + e) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 v = gelu2(__fadd2_rn(make_float2(__uint_as_float(h[hh * 8 + 2 * e]), __uint_as_float(h[hh * 8 + 2 * e + 1])),
                                            make_float2(b[2 * e], b[2 * e + 1])), hs);
          gout[hh * 4 + e] = pack_bf16(v.x, v.y);
        }
        if (kDrop) {   // the scale is already in the values (hs): AND the packed pairs with the keep masks
          const uint32_t f0 = drop_flags_from_hash_input(p.dh, hin + (2 * hh) * kDropGolden);
          const uint32_t f1 = drop_flags_from_hash_input(p.dh, hin + (2 * hh + 1) * kDropGolden);
          gout[hh * 4 + 0] &= drop_mask_bf16x2<0>(f0); gout[hh * 4 + 1] &= drop_mask_bf16x2<1>(f0);
          gout[hh * 4 + 2] &= drop_mask_bf16x2<0>(f1); gout[hh * 4 + 3] &= drop_mask_bf16x2<1>(f1);
        }
      }
    };
    // kG = 2: group g owns the chunks j = g (mod 2), four 16-column pieces per thread.
    // kG = 4: TWO groups share a chunk (chunks j = g / 2 (mod 2)); group half h = g & 1 owns columns [32 h, 32 h + 32),
    //         two pieces per thread.  At most two chunks are in the epilogue either way (that is what the 5-buffer ring
    //         sustains: with four chunks in flight every group starved for H, profiles/r01_trace_fwd_*.log), but a chunk
    //         is finished by 8 warps.  A group writes its bf16 G over the head of ITS OWN columns: G columns
    //         [32 h, 32 h + 16) of the buffer (the GEMM2 issuer picks its A operand k-steps from there).
    constexpr int kSplit = kG / 2;             // groups per chunk
    constexpr int kPieces = 4 / kSplit;        // 16-column pieces per thread and chunk
    const int cgrp = grp / kSplit, half = grp % kSplit;
    const int pc0 = half * kPieces;            // first piece of this group
    const bool tr = pwarp == 0 || pwarp == 4 * kSplit;   // first warp of the first group of each chunk parity (M2_TRACE)
    uint32_t hA[16], hB[16];
    if (cgrp < nch) {
      mbar_wait(&hfull[cgrp % NB], 0);
      tc_fence_after();
      ld_piece(cgrp, pc0, hA);
      tmem_ld_wait();
    }
    for (int j = cgrp; j < nch; j += 2) {
      const int c0 = j * kCc;
      const uint32_t tGj = tmem_base + C::kColH + lane_addr + (j % NB) * kCc + half * 32;
      uint32_t g[8];
      if (tr) M2_TR(400 + 4 * j + 0, 5, j);      // epilogue: chunk start (first piece loaded)
      // probe the next chunk's accumulator barrier now: in steady state GEMM1 is chunks ahead, and the wait's round trip
      // then hides under this chunk's math
      const bool next_ready = (j + 2 < nch) && mbar_probe(&hfull[(j + 2) % NB], ((j + 2) / NB) & 1);
#pragma unroll
      for (int pl = 0; pl < kPieces; ++pl) {
        const int pc = pc0 + pl;
        uint32_t (&cur)[16] = (pl & 1) ? hB : hA;
        uint32_t (&nxt)[16] = (pl & 1) ? hA : hB;
        if (pl < kPieces - 1) {
          ld_piece(j, pc + 1, nxt);
        } else if (j + 2 < nch) {
          if (tr) M2_TR(400 + 4 * j + 1, 6, j);  // epilogue: about to wait for the next H
          if (!next_ready) mbar_wait(&hfull[(j + 2) % NB], ((j + 2) / NB) & 1);
          __syncwarp();
          if (tr) M2_TR(400 + 4 * j + 2, 7, j);  // epilogue: next H ready
          tc_fence_after();
          ld_piece(j + 2, pc0, nxt);
        }
        // G piece pl (8 bf16x2 columns) goes over columns [8 pl, 8 pl + 8) of the group's part of the buffer: all read
        // already (the piece in flight starts at column 16 (pl + 1))
        gelu_piece(cur, c0 + pc * 16, g);
        tmem_st8(tGj + pl * 8, *reinterpret_cast<const uint32_t (*)[8]>(g));
        tmem_ld_wait();
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&gfull[j % NB]);
      if (tr) M2_TR(400 + 4 * j + 3, 8, j);      // epilogue: G stored, arrived
    }
    // final: y = u + Drop(Yacc + b2).  Pass 1: accumulator rows -> padded fp32 staging (each group half the columns).
    if (pwarp == 0) M2_TR(1304, 24, 0);   // epilogue loop done
    // final: y = u + Drop(Yacc + b2).  The u rows of pass 2 (warp per row, lanes along d) are fetched NOW, while the last
    // GEMM2s drain (the loads were the critical path of the 4200-clk tail, profiles/r01_trace_fwd_*.log).
    const int ew = warp - 2;
    constexpr int kRowsPerIter = DP == 128 ? 1 : 2;   // DP = 64: 16 lanes cover a row, two rows per warp step
    constexpr int kIters = kRows / (4 * kG * kRowsPerIter);
    const int d2 = DP == 128 ? lane * 4 : (lane & 15) * 4;
    const int rsub = DP == 128 ? 0 : (lane >> 4);
    float4 upre[kIters];
    float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d2 < p.D) {
      bq = *reinterpret_cast<const float4*>(p.b2 + d2);
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int grow = m0 + (ew + it * 4 * kG) * kRowsPerIter + rsub;
        upre[it] = grow < p.M ? *reinterpret_cast<const float4*>(p.u + static_cast<long long>(grow) * p.D + d2)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    mbar_wait(yfull, 0);
    tc_fence_after();
    if (pwarp == 0) M2_TR(1305, 25, 0);   // Y accumulator complete
    // Pass 1: accumulator rows -> padded fp32 staging (each group its share of the columns).
    constexpr int kColsPerGrp = DP / kG < 32 ? 32 : DP / kG;
#pragma unroll 1
    for (int d0 = grp * kColsPerGrp; d0 < (grp + 1) * kColsPerGrp && d0 < DP; d0 += 32) {
      uint32_t a[32];
      tmem_ld32(tY + lane_addr + d0, a);
      tmem_ld_wait();
      float* o = reinterpret_cast<float*>(sOut + r * C::kOPitch) + d0;
#pragma unroll
      for (int e = 0; e < 32; e += 4)
        *reinterpret_cast<float4*>(o + e) = make_float4(__uint_as_float(a[e]), __uint_as_float(a[e + 1]),
                                                        __uint_as_float(a[e + 2]), __uint_as_float(a[e + 3]));
    }
    named_bar_sync(1, 128 * kG);
    if (pwarp == 0) M2_TR(1306, 26, 0);   // Y staged in shared memory
    // Pass 2: warp per row, lanes along d: coalesced y write, bias and dropout applied here.
    if (d2 < p.D) {
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const int r2 = (ew + it * 4 * kG) * kRowsPerIter + rsub;
        const int grow = m0 + r2;
        if (grow < p.M) {
          const float4 acc = *reinterpret_cast<const float4*>(sOut + r2 * C::kOPitch + d2 * 4);
          float4 o = make_float4(acc.x + bq.x, acc.y + bq.y, acc.z + bq.z, acc.w + bq.w);
          if (kDrop) drop_apply4(p.dout, o, static_cast<unsigned long long>(grow) * p.D + d2);
          o.x += upre[it].x; o.y += upre[it].y; o.z += upre[it].z; o.w += upre[it].w;
          *reinterpret_cast<float4*>(p.y + static_cast<long long>(grow) * p.D + d2) = o;
        }
      }
    }
  }
  if (pwarp == 0) M2_TR(1307, 27, 0);     // output written
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ============================================================================================ backward (dgrad)
// Given dY: recompute H, form dG = dY.W2, dH = dG * GELU'(H + b1), accumulate dXn += dH.W1 in TMEM, then finish the
// LayerNorm backward in the same kernel:  du = dy + LN'(dXn), dln_w / dln_b / db2 column sums (atomics).
// TMEM map: [0,DP) dXn accumulator | LN(u) bf16 (DP/2) | dY bf16 (DP/2) | 2 x 64 H | 2 x 64 dG; the bf16 dH operand of
// the dXn GEMM overwrites columns 0..31 of its own dG buffer (each thread only touches its own lane, and the
// tensor pipe executes dXn(j) before the H/dG GEMMs of chunk j + 2 that reuse the buffer).
template <int DP>
struct CfgB {
  static constexpr int kW1Bytes = kCc * DP * 2;
  static constexpr int kW2Bytes = DP * kCc * 2;
  static constexpr int S1 = 4, S2 = 3;
  static constexpr int kRingBytes = S1 * kW1Bytes + S2 * kW2Bytes;
  static constexpr int kXPitch = CfgT<DP>::kXPitch;
  static constexpr int kXStage = kRows * kXPitch;
  static constexpr int kOPitch = DP * 4 + 16;
  static_assert(kRows * kOPitch <= kRingBytes, "output staging must fit in the weight ring");
  static constexpr int kStatBytes = 2 * kRows * 4 + 3 * DP * 4;   // mean, rstd, column partials [3][DP]
  // dH spill (kStore == 2): once the prologue has copied LN(u) / dY to tensor memory the two row staging tiles are dead;
  // the same region then holds kDhSlots [128 rows][64 c] bf16 tiles (128-byte swizzle) that leave through TMA stores.
  static constexpr int kDhSlot = kRows * 128;
  static constexpr int kDhSlots = 4;
  static constexpr int kStageRegion = 2 * kXStage > kDhSlots * kDhSlot ? 2 * kXStage : kDhSlots * kDhSlot;
  static_assert(kRingBytes % 1024 == 0, "the dH slots must be 1024-byte aligned");
  static constexpr int kSmem = kRingBytes + kStageRegion + kStatBytes + kBarBytes + 1024;
  static constexpr int kMaxBias = 16 * 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kColDX = 0;
  static constexpr int kColX = DP;
  static constexpr int kColDY = DP + DP / 2;
  static constexpr int kColH = 2 * DP;
  static constexpr int kColG = 2 * DP + 2 * kCc;
  static_assert(kColG + 2 * kCc <= 512, "TMEM budget");
};

// dY rows (masked by the output-site dropout) -> padded row-major bf16 staging (+ bf16 copy to HBM).
template <int DP, bool kDrop>
__device__ __forceinline__ void dy_rows_to_stage(const TsParams& p, int m0, uint8_t* stage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kLanes = DP / 4;                 // lanes that cover one row
  constexpr int kRowsPerIter = 32 / kLanes;      // 1 (DP = 128) or 2 (DP = 64)
  const int c = (lane % kLanes) * 4;
  for (int r0 = warp * kRowsPerIter; r0 < kRows; r0 += (blockDim.x >> 5) * kRowsPerIter) {
    const int r = r0 + lane / kLanes, row = m0 + r;
    uint2 o = make_uint2(0u, 0u);
    if (row < p.M && c < p.D) {
      float4 v = *reinterpret_cast<const float4*>(p.dy + static_cast<long long>(row) * p.D + c);
      if (kDrop) drop_apply4(p.dout, v, static_cast<unsigned long long>(row) * p.D + c);   // dY * mask * scale
      o.x = pack_bf16(v.x, v.y);
      o.y = pack_bf16(v.z, v.w);
      if (p.dy_b) *reinterpret_cast<uint2*>(p.dy_b + static_cast<long long>(row) * p.D + c) = o;
    }
    *reinterpret_cast<uint2*>(stage + r * CfgB<DP>::kXPitch + c * 2) = o;
  }
}

// kStore: 0 = nothing but du / LN(u) / dY leaves the kernel (wgrad_fused recomputes G and dH), 1 = G and dH with per-thread
// 16-byte stores (M2B200_CHAIN_GEN=3, the A/B reference), 2 = dH only, staged in shared memory in the 128-byte swizzle and
// written by TMA stores, chunk-major [C / 64][M][64] (wgrad_dh consumes it with TMA loads and recomputes only G).
template <int DP, bool kDrop, int kStore>
__global__ void __launch_bounds__(kThreadsB, 1)
chain_bwd_ts_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                    const __grid_constant__ CUtensorMap tmDH, const TsParams p) {
  constexpr bool kStoreGH = kStore == 1;
  using C = CfgB<DP>;
  constexpr int S1 = C::S1, S2 = C::S2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + S1 * C::kW1Bytes;
  uint8_t* sStageX = sW2 + S2 * C::kW2Bytes;
  uint8_t* sStageDY = sStageX + C::kXStage;
  uint8_t* sOut = smem;                              // fp32 dXn staging (aliases the ring once every MMA is done)
  uint8_t* sDH = sStageX;                            // kStore == 2: dH store slots (alias the row staging after the prologue)
  float* sMean = reinterpret_cast<float*>(sStageX + C::kStageRegion);
  float* sRstd = sMean + kRows;
  float* sCol = sRstd + kRows;                       // [3][DP]: dln_w, dln_b, db2 partial column sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + 3 * DP);
  uint64_t* w1full = bars;            // [S1]
  uint64_t* w1empty = w1full + S1;    // [S1]  dXn GEMM done with the W1 chunk -> TMA
  uint64_t* w2full = w1empty + S1;    // [S2]
  uint64_t* w2empty = w2full + S2;    // [S2]  dG GEMM done -> TMA
  uint64_t* hfull = w2empty + S2;     // [2]   H and dG accumulators of a chunk ready (one arrival per issuing warp) -> epilogue
  uint64_t* dhfull = hfull + 2;       // [2]   epilogue wrote dH (bf16, TMEM) -> MMA
  uint64_t* yfull = dhfull + 2;
  uint64_t* hloaded = yfull + 1;      // [2]   epilogue has H / dG of the chunk in registers: the H accumulator may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hloaded + 2);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);

  // Roles by LOGICAL warp id (0 = TMA producer, 1 = dXn / dG issuer, 2.. = epilogue, -1 = H issuer); physically the
  // epilogue warps come first and the single-thread roles last, and the MMAs are issued by TWO warps so that one
  // issuer's barrier waits overlap the other's MMAs (see chain_fwd_ts_kernel).  One writer per accumulator: dXn and the
  // dG buffers <- warp 1, H buffers <- warp -1.
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = pwarp < 4 * kGroupsB ? pwarp + 2 : (pwarp == 4 * kGroupsB + 2 ? -1 : pwarp - 4 * kGroupsB);
  const int m0 = blockIdx.x * kRows;
  const int nch = ceil_div(p.C, kCc);

  if (threadIdx.x == 0) {
    for (int i = 0; i < S1; ++i) { mbar_init(&w1full[i], 1); mbar_init(&w1empty[i], 1); }
    for (int i = 0; i < S2; ++i) { mbar_init(&w2full[i], 1); mbar_init(&w2empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&hfull[i], 2); mbar_init(&dhfull[i], 128 * kGroupsB); mbar_init(&hloaded[i], 128 * kGroupsB); }
    mbar_init(yfull, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    if (kStore == 2) tma_prefetch_desc(&tmDH);
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  for (int i = threadIdx.x; i < 3 * DP; i += blockDim.x) sCol[i] = 0.f;
  __syncthreads();
  if (warp == 0 && lane == 0) {
    for (int j = 0; j < (nch < S1 ? nch : S1); ++j) load_w1<DP>(sW1 + j * C::kW1Bytes, &tmW1, &w1full[j], j * kCc);
    for (int j = 0; j < (nch < S2 ? nch : S2); ++j) load_w2<DP>(sW2 + j * C::kW2Bytes, &tmW2, &w2full[j], j * kCc);
  }
  BiasRegs br;
  if (p.bias_smem) br = bias_load(p.b1, p.C, nch * kCc);           // in flight during the row prologue
  ln_rows_to_stage<DP, 8>(p, m0, sStageX, sMean, sRstd, p.xn_b);   // 19 warps x 8 rows >= 128
  dy_rows_to_stage<DP, kDrop>(p, m0, sStageDY);
  if (p.bias_smem) bias_store(br, nch * kCc, sBias);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tDX = tmem_base + C::kColDX;
  const uint32_t tX = tmem_base + C::kColX;
  const uint32_t tDY = tmem_base + C::kColDY;

  if (warp >= 2) {   // staging -> TMEM: groups 0/1 copy the two 32-column halves of LN(u), groups 2/3 those of dY
    const int q = pwarp & 3, grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int mat = grp >> 1, cb = grp & 1;
    if (cb < DP / 64) {
      const uint8_t* src = (mat == 0 ? sStageX : sStageDY) + r * C::kXPitch;
      const uint32_t dst = (mat == 0 ? tX : tDY) + lane_addr;
      stage_row_to_tmem(src + cb * 128, dst + cb * 32);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // Slot release order of the MMA issuer: hg(0), hg(1), then per iteration k: dx(k) frees W1 chunk k, hg(k + 2)
    // frees W2 chunk k + 2.
    auto refill_w1 = [&](int x) {
      if (x >= nch) return;
      mbar_wait(&w1empty[x % S1], ((x / S1) & 1) ^ 1);
      if (elect_one()) load_w1<DP>(sW1 + (x % S1) * C::kW1Bytes, &tmW1, &w1full[x % S1], x * kCc);
      __syncwarp();
    };
    auto refill_w2 = [&](int y) {
      if (y >= nch) return;
      mbar_wait(&w2empty[y % S2], ((y / S2) & 1) ^ 1);
      if (elect_one()) load_w2<DP>(sW2 + (y % S2) * C::kW2Bytes, &tmW2, &w2full[y % S2], y * kCc);
      __syncwarp();
    };
    refill_w2(S2);
    refill_w2(S2 + 1);
    for (int k = 0; S1 + k < nch || S2 + 2 + k < nch; ++k) {
      refill_w1(S1 + k);
      refill_w2(S2 + 2 + k);
    }
  } else if (warp == -1) {
    // ---- H issuer: H[j&1] = LN(u) . W1_j^T (A from TMEM), two chunks ahead: the H accumulator is free as soon as the
    // epilogue of chunk j - 2 has LOADED it (hloaded), so this GEMM overlaps the epilogue's math.
    // The whole warp walks the loop (waits are warp-wide), one elected lane issues: see elect_one().
    constexpr uint32_t idescH = umma_idesc_bf16(kRows, kCc, 0, 0);    // B (W1 chunk) K-major
    const uint64_t w1k_desc0 = umma_desc_sw128(smem_u32(sW1), 16, 1024);          // W1 chunk as K-major B (H GEMM)
    for (int j = 0; j < nch; ++j) {
      const int s1 = j % S1, b = j & 1;
      M2_TR(800 + 4 * j + 0, 11, j);
      if (j >= 2) mbar_wait2(&w1full[s1], (j / S1) & 1, &hloaded[b], ((j >> 1) - 1) & 1);
      else mbar_wait(&w1full[s1], (j / S1) & 1);
      M2_TR(800 + 4 * j + 1, 12, j);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bd1 = w1k_desc0 + static_cast<uint64_t>((s1 * C::kW1Bytes) >> 4);
        const uint32_t tH = tmem_base + C::kColH + b * kCc;
#pragma unroll
        for (int kk = 0; kk < DP / 16; ++kk)
          umma_bf16_ts(tH, tX + kk * 8, bd1 + (((kk >> 2) * (kCc * 128) + (kk & 3) * 32) >> 4), idescH, kk > 0 ? 1u : 0u);
        umma_commit(&hfull[b]);     // hfull counts two arrivals: this one and the dG GEMM's
        M2_TR(800 + 4 * j + 3, 14, j);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- dXn / dG issuer: dXn += dH_j . W1_j (A = dH from TMEM, the W1 chunk consumed MN-major), then in the SAME burst
    // dG[j&1] = dY . W2_{j+2}, which overwrites the accumulator whose head held dH_j: the tensor pipe executes one thread's
    // MMAs in issue order, so no completion round trip sits in the epilogue -> dXn -> dG -> epilogue ring (with the dG GEMM
    // on the other warp that ring cost 2 x ~300 clk of commit -> barrier -> wait latency per chunk,
    // profiles/r01_trace_bwd_two_issuers.log).  W1_j is known to have landed: the H GEMM of the chunk waited for it, and its
    // completion reached this warp through hfull -> epilogue -> dhfull.
    constexpr uint32_t idescX = umma_idesc_bf16(kRows, DP, 0, 1);     // B (W1 chunk) MN-major
    constexpr uint32_t idescG = umma_idesc_bf16(kRows, kCc, 0, 1);    // B (W2 tile)  MN-major
    const uint64_t w1m_desc0 = umma_desc_sw128(smem_u32(sW1), kCc * 128, 1024);   // W1 chunk as MN-major B (dXn GEMM)
    const uint64_t w2m_desc0 = umma_desc_sw128(smem_u32(sW2), 8192, 1024);        // W2 tile as MN-major B (dG GEMM)
    const uint64_t dh_policy = l2_policy_evict_first();
    auto issue_dg = [&](int jn) {   // elected lane only
      const int s2 = jn % S2, b = jn & 1;
      const uint64_t bd2 = w2m_desc0 + static_cast<uint64_t>((s2 * C::kW2Bytes) >> 4);
      const uint32_t tG = tmem_base + C::kColG + b * kCc;
#pragma unroll
      for (int kk = 0; kk < DP / 16; ++kk)    // B = W2 tile [DP d-rows][64 c]: 16 d-rows per step = 2048 B
        umma_bf16_ts(tG, tDY + kk * 8, bd2 + ((kk * 2048) >> 4), idescG, kk > 0 ? 1u : 0u);
      umma_commit(&w2empty[s2]);
      umma_commit(&hfull[b]);
    };
    for (int jn = 0; jn < 2 && jn < nch; ++jn) {
      mbar_wait(&w2full[jn % S2], 0);
      tc_fence_after();
      if (elect_one()) issue_dg(jn);
      __syncwarp();
    }
    for (int j = 0; j < nch; ++j) {
      const int s1 = j % S1, b = j & 1;
      const int jn = j + 2;
      M2_TR(4 * j + 0, 1, j);
      if (jn < nch) mbar_wait2(&dhfull[b], (j >> 1) & 1, &w2full[jn % S2], (jn / S2) & 1);
      else mbar_wait(&dhfull[b], (j >> 1) & 1);
      M2_TR(4 * j + 1, 2, j);
      tc_fence_after();
      if (elect_one()) {   // elect.sync names the same leader every time: the bulk groups below belong to ONE thread
        // kStore == 2: the dG GEMM issued below lets the epilogue of chunk j + 2 overwrite the store slot of chunk j - 2: that
        // store (two chunk periods old) must have finished READING shared memory.  No extra barrier carries the dH stores:
        // the slot of chunk j is complete because its writers arrived on dhfull after writing it, and hfull(j + 2) cannot
        // complete before this thread has passed the wait.
        if (kStore == 2) tma_store_wait_read_n<1>();
        const uint64_t bd = w1m_desc0 + static_cast<uint64_t>((s1 * C::kW1Bytes) >> 4);
        const uint32_t tG = tmem_base + C::kColG + b * kCc;
#pragma unroll
        for (int kk = 0; kk < kCc / 16; ++kk)   // B: 16 c-rows per step = 2048 B; d panels 8 KB apart (LBO); A: group kk's dH
          umma_bf16_ts(tDX, tG + kk * 16, bd + ((kk * 2048) >> 4), idescX, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&w1empty[s1]);
        if (jn < nch) issue_dg(jn);
        if (kStore == 2) {
          // ONE generic -> async proxy fence, by the thread that has acquired the writers' stores through dhfull (a fence in
          // each of the 512 writers cost 4-5 us per launch, profiles/r02_dh_spill.md)
          fence_proxy_async();
          if (p.l2_hint) tma_store_3d_hint(&tmDH, sDH + (j % C::kDhSlots) * C::kDhSlot, 0, m0, j, dh_policy);
          else tma_store_3d(&tmDH, sDH + (j % C::kDhSlots) * C::kDhSlot, 0, m0, j);   // one contiguous 16 KB piece
          tma_store_commit();
        }
        M2_TR(4 * j + 2, 10, j);
      }
      __syncwarp();
    }
    if (elect_one()) {
      umma_commit(yfull);
      if (kStore == 2) tma_store_wait_all();
    }
    __syncwarp();
  } else {
    const int q = pwarp & 3;
    const int grp = (warp - 2) >> 2;       // every group works on every chunk: group g owns columns [16 g, 16 g + 16)
    const int r = q * 32 + lane;
    const int row = m0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float hs = kDrop ? 0.5f * p.dh.scale : 0.5f;          // dropout scale folded into GELU / GELU'
    const float ninv_s = kDrop ? -1.f / p.dh.scale : -1.f;
    const uint32_t hrow = (static_cast<uint32_t>(row) * static_cast<uint32_t>(p.ldh) >> 2) * kDropGolden + drop_key(p.dh);
    bool ready = false;                              // hfull of the chunk already observed by an early probe
    for (int j = 0; j < nch; ++j) {
      const int b = j & 1;
      const uint32_t tH = tmem_base + C::kColH + lane_addr + b * kCc + grp * 16;
      const uint32_t tG = tmem_base + C::kColG + lane_addr + b * kCc + grp * 16;
      if (warp == 2) M2_TR(400 + 4 * j + 0, 5, j);   // epilogue: about to wait for H / dG
      if (!ready) mbar_wait(&hfull[b], (j >> 1) & 1);
      __syncwarp();
      if (warp == 2) M2_TR(400 + 4 * j + 1, 6, j);   // epilogue: accumulators ready
      tc_fence_after();
      const int c0 = j * kCc + grp * 16;
      uint32_t h[16], dg[16], dhp[8];
      tmem_ld16(tH, h);
      tmem_ld16(tG, dg);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&hloaded[b]);                      // the H accumulator of this buffer may be overwritten
      if (warp == 2) M2_TR(400 + 4 * j + 2, 7, j);   // epilogue: TMEM loads done
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int cc = c0 + ch * 8;
        float bias[8];
        if (p.bias_smem) {
          const float4 b0 = lds_f4(sBias + cc);
          const float4 b1v = lds_f4(sBias + cc + 4);
          bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
          bias[4] = b1v.x; bias[5] = b1v.y; bias[6] = b1v.z; bias[7] = b1v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) bias[e] = (cc + e < p.C) ? __ldg(p.b1 + cc + e) : 0.f;
        }
        uint32_t gpk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 dgelu;
          const float2 gv = gelu2_grad(__fadd2_rn(make_float2(__uint_as_float(h[ch * 8 + 2 * e]), __uint_as_float(h[ch * 8 + 2 * e + 1])),
                                                  make_float2(bias[2 * e], bias[2 * e + 1])), dgelu, hs, ninv_s);
          const float2 dv = __fmul2_rn(make_float2(__uint_as_float(dg[ch * 8 + 2 * e]), __uint_as_float(dg[ch * 8 + 2 * e + 1])), dgelu);
          dhp[ch * 4 + e] = pack_bf16(dv.x, dv.y);
          if (kStoreGH) gpk[e] = pack_bf16(gv.x, gv.y);
        }
        if (kDrop) {   // G' = m*s*G ; dH = dG' * m*s*gelu'(h): the scale is folded in, AND the packed pairs with the keep masks
          const uint32_t hin = hrow + static_cast<uint32_t>(cc >> 2) * kDropGolden;
          const uint32_t f0 = drop_flags_from_hash_input(p.dh, hin);
          const uint32_t f1 = drop_flags_from_hash_input(p.dh, hin + kDropGolden);
          const uint32_t m0 = drop_mask_bf16x2<0>(f0), m1 = drop_mask_bf16x2<1>(f0);
          const uint32_t m2 = drop_mask_bf16x2<0>(f1), m3 = drop_mask_bf16x2<1>(f1);
          dhp[ch * 4 + 0] &= m0; dhp[ch * 4 + 1] &= m1; dhp[ch * 4 + 2] &= m2; dhp[ch * 4 + 3] &= m3;
          if (kStoreGH) { gpk[0] &= m0; gpk[1] &= m1; gpk[2] &= m2; gpk[3] &= m3; }
        }
        if (kStoreGH) {
          if (row < p.M && cc < p.ldh) {   // ldh is a multiple of 8 >= C: whole 16-byte chunks only
            *reinterpret_cast<uint4*>(p.g_b + static_cast<long long>(row) * p.ldh + cc) = make_uint4(gpk[0], gpk[1], gpk[2], gpk[3]);
            *reinterpret_cast<uint4*>(p.dh_b + static_cast<long long>(row) * p.ldh + cc) =
                make_uint4(dhp[ch * 4], dhp[ch * 4 + 1], dhp[ch * 4 + 2], dhp[ch * 4 + 3]);
          }
        }
      }
      if (kStore == 2) {
        // the same 16 dH values -> this row's 128-byte line of the chunk's store slot (16-byte chunks 2 grp, 2 grp + 1), BEFORE
        // the arrival on dhfull: the dXn / dG issuer that dhfull wakes also issues the slot's TMA store
        uint8_t* dst = sDH + (j % C::kDhSlots) * C::kDhSlot;
        *reinterpret_cast<uint4*>(dst + sw128_offset(r, 2 * grp)) = make_uint4(dhp[0], dhp[1], dhp[2], dhp[3]);
        *reinterpret_cast<uint4*>(dst + sw128_offset(r, 2 * grp + 1)) = make_uint4(dhp[4], dhp[5], dhp[6], dhp[7]);
      }
      // dH (bf16, 8 columns) over the first half of the dG columns this thread has just read: the groups never touch
      // each other's columns, the tensor pipe reads them (dXn GEMM) before chunk j + 2 overwrites the buffer.
      tmem_st8(tG, dhp);
      // probe the next chunk's barrier (other buffer) while the store drains: its round trip hides under the store wait
      ready = (j + 1 < nch) && mbar_probe(&hfull[b ^ 1], ((j + 1) >> 1) & 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&dhfull[b]);
      if (warp == 2) M2_TR(400 + 4 * j + 3, 8, j);   // epilogue: dH stored, arrived
    }
    // ---- final: LayerNorm backward fused on the accumulator.  Pass 1: dXn rows -> padded fp32 staging.
    mbar_wait(yfull, 0);
    tc_fence_after();
    if (grp * 32 < DP) {   // 32 accumulator columns per group
      const int d0 = grp * 32;
      uint32_t a[32];
      tmem_ld32(tDX + lane_addr + d0, a);
      tmem_ld_wait();
      float* o = reinterpret_cast<float*>(sOut + r * C::kOPitch) + d0;
#pragma unroll
      for (int e = 0; e < 32; e += 4)
        *reinterpret_cast<float4*>(o + e) = make_float4(__uint_as_float(a[e]), __uint_as_float(a[e + 1]),
                                                        __uint_as_float(a[e + 2]), __uint_as_float(a[e + 3]));
    }
    named_bar_sync(1, 128 * kGroupsB);
    // Pass 2: kLanes lanes per row along d (coalesced u / dy reads, du writes):
    //   g = dXn * gamma ; du = dy + rstd * (g - mean_d(g) - xhat * mean_d(g * xhat))
    //   dln_w += sum_rows dXn * xhat ; dln_b += sum_rows dXn ; db2 += sum_rows dY(masked)
    constexpr int kLanes = DP / 4;
    constexpr int kRowsPerIter = 32 / kLanes;
    const int ew = warp - 2;
    const int d2 = (lane % kLanes) * 4;
    const int rsub = lane / kLanes;
    const float inv_d = 1.f / static_cast<float>(p.D);
    const bool dok = d2 < p.D;
    const float4 gam = dok ? *reinterpret_cast<const float4*>(p.ln_w + d2) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 aw = make_float4(0.f, 0.f, 0.f, 0.f), ab = aw, a2 = aw;
#pragma unroll 2
    for (int rr = ew * kRowsPerIter; rr < kRows; rr += 4 * kGroupsB * kRowsPerIter) {
      const int r2 = rr + rsub;
      const int grow = m0 + r2;
      const bool ok = dok && grow < p.M;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), uu = acc, dyv = acc;
      if (ok) {
        acc = *reinterpret_cast<const float4*>(sOut + r2 * C::kOPitch + d2 * 4);
        uu = *reinterpret_cast<const float4*>(p.u + static_cast<long long>(grow) * p.D + d2);
        dyv = *reinterpret_cast<const float4*>(p.dy + static_cast<long long>(grow) * p.D + d2);
      }
      const float mean = sMean[r2], rstd = sRstd[r2];
      float4 xh = make_float4((uu.x - mean) * rstd, (uu.y - mean) * rstd, (uu.z - mean) * rstd, (uu.w - mean) * rstd);
      if (!ok) xh = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 g = make_float4(acc.x * gam.x, acc.y * gam.y, acc.z * gam.z, acc.w * gam.w);
      float s1 = g.x + g.y + g.z + g.w;
      float s2 = g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
#pragma unroll
      for (int o = kLanes / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      s1 *= inv_d; s2 *= inv_d;
      if (ok) {
        float4 o4;
        o4.x = dyv.x + rstd * (g.x - s1 - xh.x * s2);
        o4.y = dyv.y + rstd * (g.y - s1 - xh.y * s2);
        o4.z = dyv.z + rstd * (g.z - s1 - xh.z * s2);
        o4.w = dyv.w + rstd * (g.w - s1 - xh.w * s2);
        *reinterpret_cast<float4*>(p.du + static_cast<long long>(grow) * p.D + d2) = o4;
        if (kDrop) drop_apply4(p.dout, dyv, static_cast<unsigned long long>(grow) * p.D + d2);
        aw.x += acc.x * xh.x; aw.y += acc.y * xh.y; aw.z += acc.z * xh.z; aw.w += acc.w * xh.w;
        ab.x += acc.x; ab.y += acc.y; ab.z += acc.z; ab.w += acc.w;
        a2.x += dyv.x; a2.y += dyv.y; a2.z += dyv.z; a2.w += dyv.w;
      }
    }
    if (dok) {
      atomicAdd(&sCol[d2], aw.x); atomicAdd(&sCol[d2 + 1], aw.y); atomicAdd(&sCol[d2 + 2], aw.z); atomicAdd(&sCol[d2 + 3], aw.w);
      atomicAdd(&sCol[DP + d2], ab.x); atomicAdd(&sCol[DP + d2 + 1], ab.y); atomicAdd(&sCol[DP + d2 + 2], ab.z); atomicAdd(&sCol[DP + d2 + 3], ab.w);
      atomicAdd(&sCol[2 * DP + d2], a2.x); atomicAdd(&sCol[2 * DP + d2 + 1], a2.y); atomicAdd(&sCol[2 * DP + d2 + 2], a2.z); atomicAdd(&sCol[2 * DP + d2 + 3], a2.w);
    }
    named_bar_sync(1, 128 * kGroupsB);
    for (int i = (warp - 2) * 32 + lane; i < 3 * DP; i += 128 * kGroupsB) {
      const int which = i / DP, d = i % DP;
      if (d < p.D) atomicAdd((which == 0 ? p.dln_w : which == 1 ? p.dln_b : p.db2) + d, sCol[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int DP, bool kDrop, int kG>
int launch_fwd_g(const CUtensorMap& t1, const CUtensorMap& t2, const TsParams& p, cudaStream_t s) {
  const int bias_bytes = ceil_div(p.C, kCc) * kCc * 4;
  TsParams pp = p;
  pp.bias_smem = (bias_bytes <= CfgT<DP>::kMaxBias && bias_bytes / 4 <= 16 * (96 + 128 * kG)) ? 1 : 0;   // one bias_load round
  const int smem = CfgT<DP>::kSmem + (pp.bias_smem ? bias_bytes : 0);
  static int configured = 0;   // largest dynamic smem size opted into so far (idempotent attribute)
  if (smem > configured) {
    if (cudaFuncSetAttribute(chain_fwd_ts_kernel<DP, kDrop, kG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return M2_ERR_LAUNCH;
    configured = smem;
  }
  LaunchScope scope("chain_fwd", s);
  chain_fwd_ts_kernel<DP, kDrop, kG><<<ceil_div(p.M, kRows), 96 + 128 * kG, smem, s>>>(t1, t2, pp);
  M2_LAUNCH_CHECK();
  return M2_OK;
}
// M2B200_FWD_GROUPS = 2 | 4 selects the number of epilogue groups (A/B runs); default 4.
inline int fwd_groups() {
  static const int g = [] {
    const char* e = getenv("M2B200_FWD_GROUPS");
    return (e && e[0] == '2') ? 2 : 4;
  }();
  return g;
}
template <int DP, bool kDrop>
int launch_fwd(const CUtensorMap& t1, const CUtensorMap& t2, const TsParams& p, cudaStream_t s) {
  return fwd_groups() == 2 ? launch_fwd_g<DP, kDrop, 2>(t1, t2, p, s) : launch_fwd_g<DP, kDrop, 4>(t1, t2, p, s);
}

template <int DP, bool kDrop, int kStore>
int launch_bwd(const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& tdh, const TsParams& p, cudaStream_t s) {
  const int bias_bytes = ceil_div(p.C, kCc) * kCc * 4;
  TsParams pp = p;
  pp.bias_smem = (bias_bytes <= CfgB<DP>::kMaxBias && bias_bytes / 4 <= 16 * kThreadsB) ? 1 : 0;   // one bias_load round
  const int smem = CfgB<DP>::kSmem + (pp.bias_smem ? bias_bytes : 0);
  static int configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(chain_bwd_ts_kernel<DP, kDrop, kStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return M2_ERR_LAUNCH;
    configured = smem;
  }
  LaunchScope scope("chain_bwd", s);
  chain_bwd_ts_kernel<DP, kDrop, kStore><<<ceil_div(p.M, kRows), kThreadsB, smem, s>>>(t1, t2, tdh, pp);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

template <int DP>
int launch_bwd_d(const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& tdh, const TsParams& p, cudaStream_t s) {
  const bool drop = p.dh.thresh || p.dout.thresh;
  const int store = p.g_b != nullptr ? 1 : (p.dh_b != nullptr ? 2 : 0);
  if (drop) {
    if (store == 2) return launch_bwd<DP, true, 2>(t1, t2, tdh, p, s);
    return store ? launch_bwd<DP, true, 1>(t1, t2, tdh, p, s) : launch_bwd<DP, true, 0>(t1, t2, tdh, p, s);
  }
  if (store == 2) return launch_bwd<DP, false, 2>(t1, t2, tdh, p, s);
  return store ? launch_bwd<DP, false, 1>(t1, t2, tdh, p, s) : launch_bwd<DP, false, 0>(t1, t2, tdh, p, s);
}

}  // namespace

#ifdef M2_TRACE
int chain_trace_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * n) == cudaSuccess ? 0 : 1;
}
#endif

bool chain_fwd_ts_supported(int D) { return D >= 16 && D <= 128 && D % 8 == 0; }

int chain_fwd_ts(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
                 int ldw2, const float* b2, float* y, int M, int D, int C, float drop_p, unsigned long long seed,
                 cudaStream_t s) {
  if (!chain_fwd_ts_supported(D) || ldw2 % 8 || ldw2 < C) return M2_ERR_ARG;
  const int DP = D <= 64 ? 64 : 128;
  CUtensorMap t1, t2;
  // W1 bf16 [C][D] (ld = D): box 64 c-rows x 64 d.   W2 bf16 [D][ldw2] (cols >= C zero): box DP d-rows x 64 c.
  int rc = make_tmap_bf16(&t1, w1b, C, D, D, kCc, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&t2, w2b, D, ldw2, ldw2, DP, kCc);
  if (rc) return rc;
  TsParams p = {};
  p.u = u; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2; p.y = y;
  p.M = M; p.D = D; p.C = C; p.ldh = (C + 7) & ~7;
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden); p.dout = make_drop(drop_p, seed, kSiteChannelOut);
  const bool drop = p.dh.thresh || p.dout.thresh;
  if (DP == 64) return drop ? launch_fwd<64, true>(t1, t2, p, s) : launch_fwd<64, false>(t1, t2, p, s);
  return drop ? launch_fwd<128, true>(t1, t2, p, s) : launch_fwd<128, false>(t1, t2, p, s);
}

// Backward dgrad chain + fused LayerNorm backward.  du = dy + LN'(dXn); dln_w, dln_b, db2 accumulate (atomics).
// xn_b / dy_b (bf16 [M][D]) are always written; g_b / dh_b (bf16 [M][ldh]) only when non-null: both = per-thread stores
// (generation 3), dh_b alone = the dH spill through TMA stores (generation 4, consumed by wgrad_dh): CHUNK-MAJOR
// [ceil(C / 64)][M][64] (every store is one contiguous 16 KB piece); ldh then only is the row stride of the dropout mask.
int chain_bwd_ts(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
                 int ldw2, const float* dy, float* du, float* dln_w, float* dln_b, float* db2, void* xn_b, void* dy_b,
                 void* g_b, void* dh_b, int ldh, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s) {
  if (!chain_fwd_ts_supported(D) || ldw2 % 8 || ldw2 < C || ldh % 8 || ldh < C) return M2_ERR_ARG;
  if (g_b != nullptr && dh_b == nullptr) return M2_ERR_ARG;
  const int DP = D <= 64 ? 64 : 128;
  CUtensorMap t1, t2, tdh;
  int rc = make_tmap_bf16(&t1, w1b, C, D, D, kCc, 64);
  if (rc) return rc;
  rc = make_tmap_bf16(&t2, w2b, D, ldw2, ldw2, DP, kCc);
  if (rc) return rc;
  tdh = t1;   // unused unless dH leaves through TMA stores: chunk-major bf16 [C / 64][M][64], box = one chunk of the row tile
  if (dh_b != nullptr && g_b == nullptr) {
    rc = make_tmap_store3d(&tdh, dh_b, 2, ceil_div(C, kCc), M, kCc, kCc, static_cast<uint64_t>(M) * kCc, kRows, kCc);
    if (rc) return rc;
  }
  TsParams p = {};
  p.u = u; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.dy = dy; p.du = du;
  p.dln_w = dln_w; p.dln_b = dln_b; p.db2 = db2;
  p.xn_b = static_cast<__nv_bfloat16*>(xn_b); p.dy_b = static_cast<__nv_bfloat16*>(dy_b);
  p.g_b = static_cast<__nv_bfloat16*>(g_b); p.dh_b = static_cast<__nv_bfloat16*>(dh_b);
  p.M = M; p.D = D; p.C = C; p.ldh = ldh;
  p.l2_hint = dh_l2_hint();
  p.dh = make_drop(drop_p, seed, kSiteChannelHidden); p.dout = make_drop(drop_p, seed, kSiteChannelOut);
  if (DP == 64) return launch_bwd_d<64>(t1, t2, tdh, p, s);
  return launch_bwd_d<128>(t1, t2, tdh, p, s);
}

}  // namespace m2
