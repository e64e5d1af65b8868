// Token-mixing half of a Mixer block on the warp-level tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32).
//
// Reference arithmetic: MixerBlock.token_mix, modules/mixer.py:30-35,43, transpose-free:
//     H^T[d,t] = sum_n LN(x)[n,d] W1[t,n] + b1[t]      U^T[d,n'] = sum_t Drop(GELU(H^T))[d,t] W2[n',t] + b2[n']
//     u[b,n',d] = x[b,n',d] + Drop(U^T[d,n'])
// The contraction lengths of the shipped configs are tiny (N = 4 or 8 tokens, T <= 32: SURVEY D2), so a tcgen05 tile would
// be >90 % padding and TMEM round trips per 4 KB sample; but on the CUDA cores the 2 x N x T FMAs per (sample, column)
// cost 3x the GELU they surround (token_mix.cu: 40-130 us per launch at B = 4096).  Here one WARP owns one sample:
//   * its [N x D] tile lives in registers in mma A-fragment order (rows = hidden columns d, k = tokens n), LayerNorm
//     statistics are reduced with three shuffles per row,
//   * GEMM1 (K = N zero-padded to 16) -> fp32 C fragments -> GELU -> packed straight into the A fragments of GEMM2
//     (the flash-attention register hand-off: two adjacent n8 C tiles are one k16 A tile), no shared memory at all,
//   * MMA row r of a 16-row tile is hidden column d0 + 2r (r < 8) / d0 + 2(r-8) + 1, so every lane owns PAIRS of adjacent
//     columns: 8-byte global accesses, and one dropout hash per pair.
// Backward (token_mix_mma_bwd_kernel): recomputes H, forms dG = dU W2, dH = dG * GELU'(H), dXn = dH W1 the same way,
// finishes the LayerNorm backward in the warp (dx = du + LN'(dXn): no dXn buffer, no separate ln_bwd launch), and gets
// the weight gradients (contractions over d) from movmatrix-transposed fragments accumulated in registers across samples.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace m2 {
namespace {

constexpr int kWarps = 8;

// WS warps share one sample (each owns DM / WS of its 16-column m-tiles).  One warp per sample (round 1) needs the whole
// [N x D] tile plus every fragment in registers (128 / 255 registers: 16 / 8 warps per SM) and walks a ~6000-instruction
// unrolled body per sample with 2 warps per scheduler: issue slots 35 % busy, stalls split between fixed-latency waits,
// scoreboards and instruction fetch (profiles/r01_token_mix_bwd_hotspots_v38.txt).  The m-tiles of a sample are independent
// once the LayerNorm row statistics are known, so WS warps split them and exchange the per-token partial sums through
// shared memory (two named-barrier rounds forward, three backward): a quarter of the registers and of the unrolled body
// per warp, four times the warps in flight.  The four warps of a sample sit on the four schedulers (wsub = warp % 4).
__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Sum `v` (already reduced over the 8 lanes that share tq, i.e. replicated in them) over the WS warps of the sample.
// buf: [WS][R][4] floats of this sample slot and exchange round; row = the value's index in [0, R).
template <int WS, int R>
__device__ __forceinline__ void xwarp_put(float* buf, int wsub, int row, int lane, float v) {
  if (WS > 1 && lane < 4) buf[(wsub * R + row) * 4 + lane] = v;
}
template <int WS, int R>
__device__ __forceinline__ float xwarp_get(const float* buf, int row, int tq, float own) {
  if (WS == 1) return own;
  float a = 0.f;
#pragma unroll
  for (int w = 0; w < WS; ++w) a += buf[(w * R + row) * 4 + tq];
  return a;
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// transpose an 8x8 b16 matrix held one row-pair per thread (thread 4r+q holds row r, cols 2q..2q+1)
__device__ __forceinline__ uint32_t movmatrix_t(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
// sum over the 8 lanes that share (lane & 3)
__device__ __forceinline__ float group_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// Which hidden column an MMA row stands for.  The contractions run over tokens (k), so ANY row <-> column labelling works as
// long as loads, stores, the LayerNorm affine and the dropout indices use the same one.
//   pair form (warps with an ODD number of m-tiles): row g / g + 8 of m-tile mt = columns 16 mt + 2 g, + 1: a lane owns PAIRS
//     of adjacent columns, and each dropout hash (one per quad of adjacent elements) serves one pair: half of it is unused;
//   quad form (even number of m-tiles per warp, the shipped D = 128 shapes): the two m-tiles 2 p, 2 p + 1 of a pair cover 32
//     columns and row g / g + 8 of tile 2 p + e = columns 32 p + 4 g + 2 e, + 1: a lane's four rows of the tile pair are ONE
//     aligned quad, so both tiles ask for the same hash (the compiler computes it once): the mask arithmetic, ~a third of
//     the forward's instructions (profiles/r02_token_mix_sass.md), halves.
// bytes of (a | b << 32) selected by the nibbles of `sel`; a nibble with bit 3 set replicates the byte's sign bit
__device__ __forceinline__ uint32_t prmt_rr(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
template <bool kQuad>
__device__ __forceinline__ int tm_col(int mtg, int g) {
  return kQuad ? 32 * (mtg >> 1) + 4 * g + 2 * (mtg & 1) : 16 * mtg + 2 * g;
}
template <bool kQuad>   // hash-input offset of the tile's quad within a row (the lane's part is tm_lane_quad)
__device__ __forceinline__ uint32_t tm_tile_quad(int mtg) {
  return static_cast<uint32_t>(kQuad ? 8 * (mtg >> 1) : 4 * mtg) * kDropGolden;
}
template <bool kQuad>
__device__ __forceinline__ uint32_t tm_lane_quad(int g) { return static_cast<uint32_t>(kQuad ? g : (g >> 1)) * kDropGolden; }
template <bool kQuad>   // which 16-bit half of the quad's flags belongs to this lane's pair of the tile
__device__ __forceinline__ uint32_t tm_shift(int mtg, int g) { return static_cast<uint32_t>(kQuad ? (mtg & 1) : (g & 1)) * 16u; }

// D = 16 * DM, T <= TP (multiple of 16), N <= NP (8 or 16).
template <int DM, int TP, int NP, bool kDrop, int WS>
__global__ void __launch_bounds__(kWarps * 32, WS == 4 ? 3 : 1)
token_mix_mma_fwd_kernel(const float* __restrict__ x, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                         const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                         const float* __restrict__ b2, float* __restrict__ u, int B, int N, int T, const Drop dh,
                         const Drop dout) {
  constexpr int D = 16 * DM;
  constexpr int KN = NP / 8;      // 8-token groups (k halves of GEMM1 / n tiles of GEMM2)
  constexpr int NT1 = TP / 8;     // n tiles of GEMM1
  constexpr int KS2 = TP / 16;    // k steps of GEMM2
  constexpr int DMW = DM / WS;    // m-tiles of this warp
  constexpr int SPB = kWarps / WS;   // samples per CTA pass
  constexpr int R = 2 * KN;       // token rows a lane quad position owns
  static_assert(DM % WS == 0 && kWarps % WS == 0, "warps per sample");
  __shared__ float sX[WS > 1 ? 2 * SPB * WS * R * 4 : 1];   // [round][slot][wsub][row][tq]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int slot = warp / WS, wsub = warp % WS;
  const int mt0 = wsub * DMW;
  float* xA = sX + (0 * SPB + slot) * (WS * R * 4);
  float* xB = sX + (1 * SPB + slot) * (WS * R * 4);

  // ---- constant B fragments and biases (registers, loaded once per warp)
  uint32_t w1f[NT1][KN];          // B1[k = n][col = t] = W1[t][n]
  float b1f[NT1][2];
#pragma unroll
  for (int j = 0; j < NT1; ++j) {
    const int t = 8 * j + g;
#pragma unroll
    for (int kh = 0; kh < KN; ++kh) {
      const int n0 = 2 * tq + 8 * kh;
      const float v0 = (t < T && n0 < N) ? w1[t * N + n0] : 0.f;
      const float v1 = (t < T && n0 + 1 < N) ? w1[t * N + n0 + 1] : 0.f;
      w1f[j][kh] = pack_bf16(v0, v1);
    }
    b1f[j][0] = (8 * j + 2 * tq < T) ? b1[8 * j + 2 * tq] : 0.f;
    b1f[j][1] = (8 * j + 2 * tq + 1 < T) ? b1[8 * j + 2 * tq + 1] : 0.f;
  }
  uint32_t w2f[KS2][KN][2];       // B2[k = t][col = n'] = W2[n'][t]
  float b2f[KN][2];
#pragma unroll
  for (int jn = 0; jn < KN; ++jn) {
    const int n = 8 * jn + g;
#pragma unroll
    for (int ks = 0; ks < KS2; ++ks)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int t0 = 16 * ks + 2 * tq + 8 * hh;
        const float v0 = (n < N && t0 < T) ? w2[n * T + t0] : 0.f;
        const float v1 = (n < N && t0 + 1 < T) ? w2[n * T + t0 + 1] : 0.f;
        w2f[ks][jn][hh] = pack_bf16(v0, v1);
      }
    b2f[jn][0] = (8 * jn + 2 * tq < N) ? b2[8 * jn + 2 * tq] : 0.f;
    b2f[jn][1] = (8 * jn + 2 * tq + 1 < N) ? b2[8 * jn + 2 * tq + 1] : 0.f;
  }
  const float inv_d = 1.f / D;
  const float hs = kDrop ? 0.5f * dh.scale : 0.5f;   // dropout scale folded into the GELU

  // dropout index arithmetic in 32 bits, incrementally: the hash input of the quad that holds (row, d_lo) is
  //   ((row * D + 16 mt + 2 g) >> 2) * golden + key  =  base(row) + 4 mt * golden,   pair select shift = 16 (g & 1);
  // rows are (b T + t) for the hidden site and (b N + n) for the output site, t / n = 2 tq + compile-time offsets.
  constexpr uint32_t kRowG = static_cast<uint32_t>(D / 4) * kDropGolden;       // one row further
  constexpr bool kQuad = DMW % 2 == 0;
  const uint32_t lane_h = tm_lane_quad<kQuad>(g) + static_cast<uint32_t>(2 * tq) * kRowG;
  const uint32_t key_h = kDrop ? drop_key(dh) : 0u, key_o = kDrop ? drop_key(dout) : 0u;
  for (int b = blockIdx.x * SPB + slot; b < B; b += gridDim.x * SPB) {   // the WS warps of a slot walk the same samples
    const float* xb = x + static_cast<long long>(b) * N * D;
    const uint32_t hin_h = static_cast<uint32_t>(b) * static_cast<uint32_t>(T) * kRowG + lane_h + key_h;   // row b T + 2 tq
    const uint32_t hin_o = static_cast<uint32_t>(b) * static_cast<uint32_t>(N) * kRowG + lane_h + key_o;   // row b N + 2 tq
    // raw tile: xr[mt][kh][nn] = x[n = 8 kh + 2 tq + nn][d = 16 (mt0 + mt) + 2 g .. +1]
    float2 xr[DMW][KN][2];
    float s[KN][2];
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        const int n = 8 * kh + 2 * tq + nn;
        float acc = 0.f;
#pragma unroll
        for (int mt = 0; mt < DMW; ++mt) {
          xr[mt][kh][nn] = n < N ? *reinterpret_cast<const float2*>(xb + n * D + tm_col<kQuad>(mt0 + mt, g)) : make_float2(0.f, 0.f);
          acc += xr[mt][kh][nn].x + xr[mt][kh][nn].y;
        }
        s[kh][nn] = group_sum(acc);
        xwarp_put<WS, R>(xA, wsub, kh * 2 + nn, lane, s[kh][nn]);
      }
    if (WS > 1) group_bar(1 + slot, 32 * WS);
    float mean[KN][2], rstd[KN][2];
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        mean[kh][nn] = xwarp_get<WS, R>(xA, kh * 2 + nn, tq, s[kh][nn]) * inv_d;
        float ss = 0.f;
#pragma unroll
        for (int mt = 0; mt < DMW; ++mt) {
          const float a = xr[mt][kh][nn].x - mean[kh][nn], c = xr[mt][kh][nn].y - mean[kh][nn];
          ss += a * a + c * c;
        }
        s[kh][nn] = group_sum(ss);
        xwarp_put<WS, R>(xB, wsub, kh * 2 + nn, lane, s[kh][nn]);
      }
    if (WS > 1) group_bar(1 + slot, 32 * WS);
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn)
        rstd[kh][nn] = rsqrtf(xwarp_get<WS, R>(xB, kh * 2 + nn, tq, s[kh][nn]) * inv_d + kLnEps);

#pragma unroll
    for (int mt = 0; mt < DMW; ++mt) {
      const int mtg = mt0 + mt;
      const int d_lo = tm_col<kQuad>(mtg, g);
      const uint32_t dsh = tm_shift<kQuad>(mtg, g);
      const uint32_t sel_lo = 0xCC88u + 0x1111u * (dsh >> 3);   // d_lo is byte dsh / 8 of the quad's flags, d_hi the next one
      const float2 gm = *reinterpret_cast<const float2*>(ln_w + d_lo);
      const float2 bt = *reinterpret_cast<const float2*>(ln_b + d_lo);
      // A1: a0 = (row g = d_lo, k = 2tq..+1), a1 = (row g+8 = d_hi, same k), a2 / a3 = k + 8
      uint32_t a1f[KN > 2 ? 2 * KN : 4] = {};   // KN = 1: the upper k half stays zero
#pragma unroll
      for (int kh = 0; kh < KN; ++kh) {
        const float x00 = (xr[mt][kh][0].x - mean[kh][0]) * rstd[kh][0] * gm.x + bt.x;
        const float x01 = (xr[mt][kh][1].x - mean[kh][1]) * rstd[kh][1] * gm.x + bt.x;
        const float x10 = (xr[mt][kh][0].y - mean[kh][0]) * rstd[kh][0] * gm.y + bt.y;
        const float x11 = (xr[mt][kh][1].y - mean[kh][1]) * rstd[kh][1] * gm.y + bt.y;
        const int n = 8 * kh + 2 * tq;
        a1f[2 * kh] = pack_bf16(n < N ? x00 : 0.f, n + 1 < N ? x01 : 0.f);
        a1f[2 * kh + 1] = pack_bf16(n < N ? x10 : 0.f, n + 1 < N ? x11 : 0.f);
      }
      float c[NT1][4];
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        c[j][0] = c[j][2] = b1f[j][0];
        c[j][1] = c[j][3] = b1f[j][1];
        mma_bf16(c[j], a1f[0], a1f[1], a1f[2], a1f[3], w1f[j][0], KN > 1 ? w1f[j][1] : 0u);
        if constexpr (KN > 2)      // N up to 32 tokens: a second k-step (MIMIC-H's 24 time steps / 25 fused tokens)
          mma_bf16(c[j], a1f[4], a1f[5], a1f[6], a1f[7], w1f[j][2], w1f[j][3]);
      }
      // GELU (+ dropout) -> A2 fragments; C tile j: c0 = (d_lo, t), c1 = (d_lo, t+1), c2 = (d_hi, t), c3 = (d_hi, t+1)
      uint32_t a2f[KS2][4];
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        const float2 lo = gelu2(make_float2(c[j][0], c[j][1]), hs);   // scale folded into the GELU (hs)
        const float2 hi = gelu2(make_float2(c[j][2], c[j][3]), hs);
        uint32_t plo = pack_bf16(lo.x, lo.y), phi = pack_bf16(hi.x, hi.y);
        if (kDrop) {   // hidden-site index (b T + t) D + d, t = 8 j + 2 tq (+ 1): the packed halves are two ROWS of one column
          const uint32_t h0 = hin_h + static_cast<uint32_t>(8 * j) * kRowG + tm_tile_quad<kQuad>(mtg);
          const uint32_t f0 = drop_flags_from_hash_input(dh, h0), f1 = drop_flags_from_hash_input(dh, h0 + kRowG);
          // AND masks straight on the packed pairs: one sign-replicating PRMT over the two rows' flag words per pair (byte k of
          // f0 -> low half, byte k of f1 -> high half) instead of a test, a compare and a select per element
          plo &= prmt_rr(f0, f1, sel_lo);
          phi &= prmt_rr(f0, f1, sel_lo + 0x1111u);
        }
        a2f[j >> 1][(j & 1) * 2] = plo;
        a2f[j >> 1][(j & 1) * 2 + 1] = phi;
      }
      float o[KN][4];
#pragma unroll
      for (int jn = 0; jn < KN; ++jn) {
        o[jn][0] = o[jn][2] = b2f[jn][0];
        o[jn][1] = o[jn][3] = b2f[jn][1];
#pragma unroll
        for (int ks = 0; ks < KS2; ++ks)
          mma_bf16(o[jn], a2f[ks][0], a2f[ks][1], a2f[ks][2], a2f[ks][3], w2f[ks][jn][0], w2f[ks][jn][1]);
      }
      // o[jn]: c0 = (d_lo, n'), c1 = (d_lo, n'+1), c2 = (d_hi, n'), c3 = (d_hi, n'+1),  n' = 8 jn + 2 tq
#pragma unroll
      for (int jn = 0; jn < KN; ++jn)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          const int n = 8 * jn + 2 * tq + nn;
          if (n < N) {
            float v_lo = o[jn][nn], v_hi = o[jn][2 + nn];
            const long long off = (static_cast<long long>(b) * N + n) * D + d_lo;
            if (kDrop)   // output-site index (b N + n) D + d, n = 8 jn + 2 tq + nn
              drop_apply2_hin(dout, v_lo, v_hi, hin_o + static_cast<uint32_t>(8 * jn + nn) * kRowG + tm_tile_quad<kQuad>(mtg), dsh);
            *reinterpret_cast<float2*>(u + off) = make_float2(xr[mt][jn][nn].x + v_lo, xr[mt][jn][nn].y + v_hi);
          }
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  dx = du + LN'(dXn);  dln_w, dln_b, dw1, db1, dw2, db2 are accumulated (per-CTA shared-memory partials, then
// one global atomic per element and CTA).  kW warps per CTA, WS of them per sample; sDxn is the lane-private spill of the dXn
// tile that the LayerNorm backward needs after the per-row sums over the whole hidden axis are known.
template <int DM, int TP, int NP, bool kDrop, int kW, int kMinB, int WS>
__global__ void __launch_bounds__(kW * 32, kMinB)
token_mix_mma_bwd_kernel(const float* __restrict__ du, const float* __restrict__ x, const float* __restrict__ ln_w,
                         const float* __restrict__ ln_b, const float* __restrict__ w1, const float* __restrict__ b1,
                         const float* __restrict__ w2, float* __restrict__ dx, float* __restrict__ dln_w,
                         float* __restrict__ dln_b, float* __restrict__ dw1, float* __restrict__ db1,
                         float* __restrict__ dw2, float* __restrict__ db2, int B, int N, int T, const Drop dh, const Drop dout) {
  constexpr int D = 16 * DM;
  constexpr int KN = NP / 8;      // 8-token groups
  constexpr int NT1 = TP / 8;     // 8-wide hidden-unit tiles
  constexpr int KS2 = TP / 16;    // 16-wide hidden-unit steps (k steps of the dXn GEMM, m tiles of the weight-gradient GEMMs)
  constexpr int DMW = DM / WS;    // m-tiles of this warp
  constexpr int SPB = kW / WS;    // samples per CTA pass
  constexpr int R = 2 * KN;       // token rows a lane quad position owns
  constexpr int DQ = (DMW + 3) / 4;
  static_assert(DM % WS == 0 && kW % WS == 0, "warps per sample");
  extern __shared__ float smem_f[];
  float* sAcc = smem_f;                               // [2][TP][NP] dw1^T / dw2^T partials, [TP] db1, [NP] db2, [2][D] dln_w / dln_b
  constexpr int kAcc = 2 * TP * NP + TP + NP + 2 * D;
  float* sDxn = sAcc + kAcc;                          // [kW][DMW * KN * 4][32] lane-private
  float* sX = sDxn + kW * (DMW * KN * 4 * 32);        // [3 rounds][SPB][WS][2 R][4] cross-warp partial sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int slot = warp / WS, wsub = warp % WS;
  const int mt0 = wsub * DMW;
  float* xA = sX + (0 * SPB + slot) * (WS * 2 * R * 4);
  float* xB = sX + (1 * SPB + slot) * (WS * 2 * R * 4);
  float* xC = sX + (2 * SPB + slot) * (WS * 2 * R * 4);
  for (int i = threadIdx.x; i < kAcc; i += kW * 32) sAcc[i] = 0.f;
  float* myDxn = sDxn + warp * (DMW * KN * 4 * 32) + lane;

  // ---- constant B fragments
  uint32_t w1f[NT1][KN];          // GEMM1  B[k = n][col = t]  = W1[t][n]
  uint32_t w2g[NT1][KN];          // GEMM3  B[k = n'][col = t] = W2[n'][t]
  float b1f[NT1][2];
#pragma unroll
  for (int j = 0; j < NT1; ++j) {
    const int t = 8 * j + g;
#pragma unroll
    for (int kh = 0; kh < KN; ++kh) {
      const int n0 = 2 * tq + 8 * kh;
      w1f[j][kh] = pack_bf16((t < T && n0 < N) ? w1[t * N + n0] : 0.f, (t < T && n0 + 1 < N) ? w1[t * N + n0 + 1] : 0.f);
      w2g[j][kh] = pack_bf16((t < T && n0 < N) ? w2[n0 * T + t] : 0.f, (t < T && n0 + 1 < N) ? w2[(n0 + 1) * T + t] : 0.f);
    }
    b1f[j][0] = (8 * j + 2 * tq < T) ? b1[8 * j + 2 * tq] : 0.f;
    b1f[j][1] = (8 * j + 2 * tq + 1 < T) ? b1[8 * j + 2 * tq + 1] : 0.f;
  }
  uint32_t w1g[KS2][KN][2];       // GEMM4  B[k = t][col = n]  = W1[t][n]
#pragma unroll
  for (int jn = 0; jn < KN; ++jn) {
    const int n = 8 * jn + g;
#pragma unroll
    for (int ks = 0; ks < KS2; ++ks)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int t0 = 16 * ks + 2 * tq + 8 * hh;
        w1g[ks][jn][hh] = pack_bf16((n < N && t0 < T) ? w1[t0 * N + n] : 0.f, (n < N && t0 + 1 < T) ? w1[(t0 + 1) * N + n] : 0.f);
      }
  }
  // ---- accumulators that live across samples
  float gw1[KS2][KN][4], gw2[KS2][KN][4];   // C fragments of dW1[t][n], dW2^T[t][n']
  float db1a[NT1][2], db2a[KN][2], dgam[DQ][2], dbet[DQ][2];
#pragma unroll
  for (int a = 0; a < KS2; ++a)
#pragma unroll
    for (int c = 0; c < KN; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) { gw1[a][c][e] = 0.f; gw2[a][c][e] = 0.f; }
#pragma unroll
  for (int j = 0; j < NT1; ++j) { db1a[j][0] = 0.f; db1a[j][1] = 0.f; }
#pragma unroll
  for (int c = 0; c < KN; ++c) { db2a[c][0] = 0.f; db2a[c][1] = 0.f; }
#pragma unroll
  for (int q = 0; q < DQ; ++q) { dgam[q][0] = dgam[q][1] = dbet[q][0] = dbet[q][1] = 0.f; }
  const float inv_d = 1.f / D;
  const float hs = kDrop ? 0.5f * dh.scale : 0.5f;          // dropout scale folded into GELU / GELU'
  const float ninv_s = kDrop ? -1.f / dh.scale : -1.f;
  __syncthreads();

  // dropout index arithmetic in 32 bits, incrementally (see the forward kernel)
  constexpr uint32_t kRowG = static_cast<uint32_t>(D / 4) * kDropGolden;
  constexpr bool kQuad = DMW % 2 == 0;
  const uint32_t lane_h = tm_lane_quad<kQuad>(g) + static_cast<uint32_t>(2 * tq) * kRowG;
  const uint32_t key_h = kDrop ? drop_key(dh) : 0u, key_o = kDrop ? drop_key(dout) : 0u;
  for (int b = blockIdx.x * SPB + slot; b < B; b += gridDim.x * SPB) {   // the WS warps of a slot walk the same samples
    const long long sbase = static_cast<long long>(b) * N * D;
    const uint32_t hin_h = static_cast<uint32_t>(b) * static_cast<uint32_t>(T) * kRowG + lane_h + key_h;   // row b T + 2 tq
    const uint32_t hin_o = static_cast<uint32_t>(b) * static_cast<uint32_t>(N) * kRowG + lane_h + key_o;   // row b N + 2 tq
    const float* xb = x + sbase;
    const float* dub = du + sbase;
    float2 xr[DMW][KN][2];
    float mean[KN][2], rstd[KN][2];
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        const int n = 8 * kh + 2 * tq + nn;
        float acc = 0.f;
#pragma unroll
        for (int mt = 0; mt < DMW; ++mt) {
          xr[mt][kh][nn] = n < N ? *reinterpret_cast<const float2*>(xb + n * D + tm_col<kQuad>(mt0 + mt, g)) : make_float2(0.f, 0.f);
          acc += xr[mt][kh][nn].x + xr[mt][kh][nn].y;
        }
        mean[kh][nn] = group_sum(acc);
        xwarp_put<WS, 2 * R>(xA, wsub, kh * 2 + nn, lane, mean[kh][nn]);
      }
    if (WS > 1) group_bar(1 + slot, 32 * WS);
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        mean[kh][nn] = xwarp_get<WS, 2 * R>(xA, kh * 2 + nn, tq, mean[kh][nn]) * inv_d;
        float ss = 0.f;
#pragma unroll
        for (int mt = 0; mt < DMW; ++mt) {
          const float a = xr[mt][kh][nn].x - mean[kh][nn], c = xr[mt][kh][nn].y - mean[kh][nn];
          ss += a * a + c * c;
        }
        rstd[kh][nn] = group_sum(ss);
        xwarp_put<WS, 2 * R>(xB, wsub, kh * 2 + nn, lane, rstd[kh][nn]);
      }
    if (WS > 1) group_bar(1 + slot, 32 * WS);
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn)
        rstd[kh][nn] = rsqrtf(xwarp_get<WS, 2 * R>(xB, kh * 2 + nn, tq, rstd[kh][nn]) * inv_d + kLnEps);
    float s1[KN][2], s2[KN][2];   // per-token sums over d of gamma dXn and gamma dXn xhat
#pragma unroll
    for (int kh = 0; kh < KN; ++kh) { s1[kh][0] = s1[kh][1] = s2[kh][0] = s2[kh][1] = 0.f; }

#pragma unroll
    for (int mt = 0; mt < DMW; ++mt) {
      const int mtg = mt0 + mt;
      const int d_lo = tm_col<kQuad>(mtg, g);
      const uint32_t dsh = tm_shift<kQuad>(mtg, g);
      const float2 gm = *reinterpret_cast<const float2*>(ln_w + d_lo);
      const float2 bt = *reinterpret_cast<const float2*>(ln_b + d_lo);
      uint32_t a1f[KN > 2 ? 2 * KN : 4] = {}, a3f[KN > 2 ? 2 * KN : 4] = {};
      float2 xh[KN][2];            // xhat of this tile
#pragma unroll
      for (int kh = 0; kh < KN; ++kh) {
        const int n = 8 * kh + 2 * tq;
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          xh[kh][nn].x = (xr[mt][kh][nn].x - mean[kh][nn]) * rstd[kh][nn];
          xh[kh][nn].y = (xr[mt][kh][nn].y - mean[kh][nn]) * rstd[kh][nn];
        }
        const bool v0 = n < N, v1 = n + 1 < N;
        a1f[2 * kh] = pack_bf16(v0 ? xh[kh][0].x * gm.x + bt.x : 0.f, v1 ? xh[kh][1].x * gm.x + bt.x : 0.f);
        a1f[2 * kh + 1] = pack_bf16(v0 ? xh[kh][0].y * gm.y + bt.y : 0.f, v1 ? xh[kh][1].y * gm.y + bt.y : 0.f);
        float2 u0 = v0 ? *reinterpret_cast<const float2*>(dub + n * D + d_lo) : make_float2(0.f, 0.f);
        float2 u1 = v1 ? *reinterpret_cast<const float2*>(dub + (n + 1) * D + d_lo) : make_float2(0.f, 0.f);
        if (kDrop) {   // gradient of the dropped branch output: rows b N + n, n = 8 kh + 2 tq (+ 1)
          const uint32_t h0 = hin_o + static_cast<uint32_t>(8 * kh) * kRowG + tm_tile_quad<kQuad>(mtg);
          drop_apply2_hin(dout, u0.x, u0.y, h0, dsh);
          drop_apply2_hin(dout, u1.x, u1.y, h0 + kRowG, dsh);
        }
        db2a[kh][0] += u0.x + u0.y;
        db2a[kh][1] += u1.x + u1.y;
        a3f[2 * kh] = pack_bf16(u0.x, u1.x);
        a3f[2 * kh + 1] = pack_bf16(u0.y, u1.y);
      }
      float c1[NT1][4], c3[NT1][4];
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        c1[j][0] = c1[j][2] = b1f[j][0];
        c1[j][1] = c1[j][3] = b1f[j][1];
        c3[j][0] = c3[j][1] = c3[j][2] = c3[j][3] = 0.f;
        mma_bf16(c1[j], a1f[0], a1f[1], a1f[2], a1f[3], w1f[j][0], KN > 1 ? w1f[j][1] : 0u);
        mma_bf16(c3[j], a3f[0], a3f[1], a3f[2], a3f[3], w2g[j][0], KN > 1 ? w2g[j][1] : 0u);
        if constexpr (KN > 2) {
          mma_bf16(c1[j], a1f[4], a1f[5], a1f[6], a1f[7], w1f[j][2], w1f[j][3]);
          mma_bf16(c3[j], a3f[4], a3f[5], a3f[6], a3f[7], w2g[j][2], w2g[j][3]);
        }
      }
      // G = Drop(GELU(H)), dH = dG * Drop'(.) * GELU'(H); packs: lo = (rows 0-7 = d_lo set), hi = (rows 8-15 = d_hi set)
      uint32_t pG[NT1][2], pH[NT1][2];
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        float2 dg_lo, dg_hi;
        float2 g_lo = gelu2_grad(make_float2(c1[j][0], c1[j][1]), dg_lo, hs, ninv_s);
        float2 g_hi = gelu2_grad(make_float2(c1[j][2], c1[j][3]), dg_hi, hs, ninv_s);
        float2 h_lo = __fmul2_rn(make_float2(c3[j][0], c3[j][1]), dg_lo);
        float2 h_hi = __fmul2_rn(make_float2(c3[j][2], c3[j][3]), dg_hi);
        if (kDrop) {   // rows b T + t, t = 8 j + 2 tq (+ 1)
          const uint32_t h0 = hin_h + static_cast<uint32_t>(8 * j) * kRowG + tm_tile_quad<kQuad>(mtg);
          drop_zero2x2_hin(dh, g_lo.x, g_hi.x, h_lo.x, h_hi.x, h0, dsh);   // scale folded into gelu2_grad
          drop_zero2x2_hin(dh, g_lo.y, g_hi.y, h_lo.y, h_hi.y, h0 + kRowG, dsh);
        }
        db1a[j][0] += h_lo.x + h_hi.x;
        db1a[j][1] += h_lo.y + h_hi.y;
        pG[j][0] = pack_bf16(g_lo.x, g_lo.y); pG[j][1] = pack_bf16(g_hi.x, g_hi.y);
        pH[j][0] = pack_bf16(h_lo.x, h_lo.y); pH[j][1] = pack_bf16(h_hi.x, h_hi.y);
      }
      // dXn^T[d, n] = sum_t dH[d,t] W1[t,n]
      float c4[KN][4];
#pragma unroll
      for (int jn = 0; jn < KN; ++jn) {
        c4[jn][0] = c4[jn][1] = c4[jn][2] = c4[jn][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS2; ++ks)
          mma_bf16(c4[jn], pH[2 * ks][0], pH[2 * ks][1], pH[2 * ks + 1][0], pH[2 * ks + 1][1], w1g[ks][jn][0], w1g[ks][jn][1]);
      }
      // LayerNorm backward partials; c4[jn]: c0 = (d_lo, n), c1 = (d_lo, n+1), c2 = (d_hi, n), c3 = (d_hi, n+1)
      float pgl = 0.f, pgh = 0.f, pbl = 0.f, pbh = 0.f;
#pragma unroll
      for (int jn = 0; jn < KN; ++jn)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          const float dlo = c4[jn][nn], dhi = c4[jn][2 + nn];
          s1[jn][nn] += gm.x * dlo + gm.y * dhi;
          s2[jn][nn] += gm.x * dlo * xh[jn][nn].x + gm.y * dhi * xh[jn][nn].y;
          pgl += dlo * xh[jn][nn].x; pgh += dhi * xh[jn][nn].y;
          pbl += dlo; pbh += dhi;
          myDxn[((mt * KN + jn) * 4 + nn * 2) * 32] = dlo;
          myDxn[((mt * KN + jn) * 4 + nn * 2 + 1) * 32] = dhi;
        }
      // dln_w / dln_b: sum over the tokens held by the 4 lanes of this row group; lane tq keeps the tiles with mt % 4 == tq
      pgl += __shfl_xor_sync(0xffffffffu, pgl, 1); pgl += __shfl_xor_sync(0xffffffffu, pgl, 2);
      pgh += __shfl_xor_sync(0xffffffffu, pgh, 1); pgh += __shfl_xor_sync(0xffffffffu, pgh, 2);
      pbl += __shfl_xor_sync(0xffffffffu, pbl, 1); pbl += __shfl_xor_sync(0xffffffffu, pbl, 2);
      pbh += __shfl_xor_sync(0xffffffffu, pbh, 1); pbh += __shfl_xor_sync(0xffffffffu, pbh, 2);
      if ((mt & 3) == tq) {
        dgam[mt >> 2][0] += pgl; dgam[mt >> 2][1] += pgh;
        dbet[mt >> 2][0] += pbl; dbet[mt >> 2][1] += pbh;
      }
      // weight gradients: contraction over the 16 hidden columns of the tile -> transposed fragments
      uint32_t bx[KN][2], bu[KN][2];
#pragma unroll
      for (int jn = 0; jn < KN; ++jn) {
        bx[jn][0] = movmatrix_t(a1f[2 * jn]); bx[jn][1] = movmatrix_t(a1f[2 * jn + 1]);
        bu[jn][0] = movmatrix_t(a3f[2 * jn]); bu[jn][1] = movmatrix_t(a3f[2 * jn + 1]);
      }
#pragma unroll
      for (int m5 = 0; m5 < KS2; ++m5) {
        const uint32_t h0 = movmatrix_t(pH[2 * m5][0]), h1 = movmatrix_t(pH[2 * m5 + 1][0]);
        const uint32_t h2 = movmatrix_t(pH[2 * m5][1]), h3 = movmatrix_t(pH[2 * m5 + 1][1]);
        const uint32_t g0 = movmatrix_t(pG[2 * m5][0]), g1 = movmatrix_t(pG[2 * m5 + 1][0]);
        const uint32_t g2 = movmatrix_t(pG[2 * m5][1]), g3 = movmatrix_t(pG[2 * m5 + 1][1]);
#pragma unroll
        for (int jn = 0; jn < KN; ++jn) {
          mma_bf16(gw1[m5][jn], h0, h1, h2, h3, bx[jn][0], bx[jn][1]);   // dW1[t][n]   += dH^T . Xn^T
          mma_bf16(gw2[m5][jn], g0, g1, g2, g3, bu[jn][0], bu[jn][1]);   // dW2^T[t][n'] += G^T . dU^T
        }
      }
    }
    // ---- LayerNorm backward: dx = du + rstd (gamma dXn - mean_d(gamma dXn) - xhat mean_d(gamma dXn xhat))
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        s1[kh][nn] = group_sum(s1[kh][nn]);
        s2[kh][nn] = group_sum(s2[kh][nn]);
        xwarp_put<WS, 2 * R>(xC, wsub, kh * 2 + nn, lane, s1[kh][nn]);
        xwarp_put<WS, 2 * R>(xC, wsub, R + kh * 2 + nn, lane, s2[kh][nn]);
      }
    if (WS > 1) group_bar(1 + slot, 32 * WS);
#pragma unroll
    for (int kh = 0; kh < KN; ++kh)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        s1[kh][nn] = xwarp_get<WS, 2 * R>(xC, kh * 2 + nn, tq, s1[kh][nn]) * inv_d;
        s2[kh][nn] = xwarp_get<WS, 2 * R>(xC, R + kh * 2 + nn, tq, s2[kh][nn]) * inv_d;
      }
#pragma unroll
    for (int mt = 0; mt < DMW; ++mt) {
      const int d_lo = tm_col<kQuad>(mt0 + mt, g);
      const float2 gm = *reinterpret_cast<const float2*>(ln_w + d_lo);
#pragma unroll
      for (int kh = 0; kh < KN; ++kh)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          const int n = 8 * kh + 2 * tq + nn;
          if (n < N) {
            const float dlo = myDxn[((mt * KN + kh) * 4 + nn * 2) * 32], dhi = myDxn[((mt * KN + kh) * 4 + nn * 2 + 1) * 32];
            const float xlo = (xr[mt][kh][nn].x - mean[kh][nn]) * rstd[kh][nn];
            const float xhi = (xr[mt][kh][nn].y - mean[kh][nn]) * rstd[kh][nn];
            const float2 ur = *reinterpret_cast<const float2*>(dub + n * D + d_lo);
            float2 o;
            o.x = ur.x + rstd[kh][nn] * (gm.x * dlo - s1[kh][nn] - xlo * s2[kh][nn]);
            o.y = ur.y + rstd[kh][nn] * (gm.y * dhi - s1[kh][nn] - xhi * s2[kh][nn]);
            *reinterpret_cast<float2*>(dx + sbase + n * D + d_lo) = o;
          }
        }
    }
  }

  // ---- per-CTA reduction of the accumulators, then one global atomic per element
  float* sW1 = sAcc;                    // [TP][NP]
  float* sW2 = sW1 + TP * NP;           // [TP][NP]
  float* sB1 = sW2 + TP * NP;           // [TP]
  float* sB2 = sB1 + TP;                // [NP]
  float* sGam = sB2 + NP;               // [D]
  float* sBet = sGam + D;               // [D]
#pragma unroll
  for (int m5 = 0; m5 < KS2; ++m5)
#pragma unroll
    for (int jn = 0; jn < KN; ++jn)
#pragma unroll
      for (int e = 0; e < 4; ++e) {     // c0 = (t = 16 m5 + g, n = 8 jn + 2 tq), c1 = n + 1, c2 / c3 = t + 8
        const int t = 16 * m5 + g + (e >> 1) * 8, n = 8 * jn + 2 * tq + (e & 1);
        atomicAdd(&sW1[t * NP + n], gw1[m5][jn][e]);
        atomicAdd(&sW2[t * NP + n], gw2[m5][jn][e]);
      }
#pragma unroll
  for (int j = 0; j < NT1; ++j) {
    const float v0 = group_sum(db1a[j][0]), v1 = group_sum(db1a[j][1]);
    if (g == 0) { atomicAdd(&sB1[8 * j + 2 * tq], v0); atomicAdd(&sB1[8 * j + 2 * tq + 1], v1); }
  }
#pragma unroll
  for (int kh = 0; kh < KN; ++kh) {
    const float v0 = group_sum(db2a[kh][0]), v1 = group_sum(db2a[kh][1]);
    if (g == 0) { atomicAdd(&sB2[8 * kh + 2 * tq], v0); atomicAdd(&sB2[8 * kh + 2 * tq + 1], v1); }
  }
#pragma unroll
  for (int mt = 0; mt < DMW; ++mt)
    if ((mt & 3) == tq) {
      const int d_lo = tm_col<kQuad>(mt0 + mt, g);
      atomicAdd(&sGam[d_lo], dgam[mt >> 2][0]); atomicAdd(&sGam[d_lo + 1], dgam[mt >> 2][1]);
      atomicAdd(&sBet[d_lo], dbet[mt >> 2][0]); atomicAdd(&sBet[d_lo + 1], dbet[mt >> 2][1]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < TP * NP; i += kW * 32) {
    const int t = i / NP, n = i - t * NP;
    if (t < T && n < N) {
      atomicAdd(dw1 + t * N + n, sW1[i]);
      atomicAdd(dw2 + n * T + t, sW2[i]);
    }
  }
  for (int i = threadIdx.x; i < TP; i += kW * 32) if (i < T) atomicAdd(db1 + i, sB1[i]);
  for (int i = threadIdx.x; i < NP; i += kW * 32) if (i < N) atomicAdd(db2 + i, sB2[i]);
  for (int i = threadIdx.x; i < D; i += kW * 32) { atomicAdd(dln_w + i, sGam[i]); atomicAdd(dln_b + i, sBet[i]); }
}

// Warps per sample: M2B200_TOKEN_WS = 1 | 2 | 4 overrides (A/B runs); default 4 for D >= 128, 2 for D = 64, else 1.
inline int token_ws(int dm) {
  static const int env = [] {
    const char* e = getenv("M2B200_TOKEN_WS");
    return e ? atoi(e) : 0;
  }();
  int ws = env > 0 ? env : (dm >= 8 ? 4 : (dm >= 4 ? 2 : 1));
  while (ws > 1 && (dm % ws || ws > 4)) ws >>= 1;
  return ws == 3 ? 2 : ws;
}

template <int DM, int TP, int NP, int WS, int kW = 8, int kMinB = 1>
int launch_bwd_ws(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
               const float* w2, float* dx, float* dln_w, float* dln_b, float* dw1, float* db1, float* dw2, float* db2, int B,
               int N, int T, const Drop& dh, const Drop& dout, cudaStream_t s) {
  constexpr int D = 16 * DM, KN = NP / 8, DMW = DM / WS, SPB = kW / WS;
  constexpr size_t smem = (2 * TP * NP + TP + NP + 2 * D + static_cast<size_t>(kW) * DMW * KN * 4 * 32 +
                           3 * SPB * WS * 4 * KN * 4) * sizeof(float);
  int grid = ceil_div(B, SPB);
  // persistent: the per-CTA gradient partials cost one atomic round per CTA; as many CTAs as can be resident
  const int per_sm = 2 * kMinB;
  if (grid > 148 * per_sm) grid = 148 * per_sm;
  LaunchScope scope("token_mix_mma_bwd", s);
  auto launch = [&](auto kern) -> int {
    static bool configured = false;   // per instantiation of this generic lambda
    if (!configured && smem > 48 * 1024) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return M2_ERR_LAUNCH;
      configured = true;
    }
    kern<<<grid, kW * 32, smem, s>>>(du, x, ln_w, ln_b, w1, b1, w2, dx, dln_w, dln_b, dw1, db1, dw2, db2, B, N, T, dh, dout);
    return M2_OK;
  };
  int rc;
  if (dh.thresh || dout.thresh) rc = launch(token_mix_mma_bwd_kernel<DM, TP, NP, true, kW, kMinB, WS>);
  else rc = launch(token_mix_mma_bwd_kernel<DM, TP, NP, false, kW, kMinB, WS>);
  if (rc) return rc;
  M2_LAUNCH_CHECK();
  return M2_OK;
}
template <int DM, int TP, int NP>
int launch_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
               const float* w2, float* dx, float* dln_w, float* dln_b, float* dw1, float* db1, float* dw2, float* db2, int B,
               int N, int T, const Drop& dh, const Drop& dout, cudaStream_t s) {
  const int ws = token_ws(DM);
  if constexpr (DM % 4 == 0)
    if (ws == 4) return launch_bwd_ws<DM, TP, NP, 4, 8, 2>(du, x, ln_w, ln_b, w1, b1, w2, dx, dln_w, dln_b, dw1, db1, dw2, db2, B, N, T, dh, dout, s);
  if constexpr (DM % 2 == 0)
    if (ws >= 2) return launch_bwd_ws<DM, TP, NP, 2>(du, x, ln_w, ln_b, w1, b1, w2, dx, dln_w, dln_b, dw1, db1, dw2, db2, B, N, T, dh, dout, s);
  return launch_bwd_ws<DM, TP, NP, 1>(du, x, ln_w, ln_b, w1, b1, w2, dx, dln_w, dln_b, dw1, db1, dw2, db2, B, N, T, dh, dout, s);
}

template <int DM, int TP, int NP, int WS>
int launch_fwd_ws(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* u, int B, int N, int T, const Drop& dh, const Drop& dout, cudaStream_t s) {
  int grid = ceil_div(B, kWarps / WS);
  if (grid > 148 * 8) grid = 148 * 8;
  LaunchScope scope("token_mix_mma_fwd", s);
  if (dh.thresh || dout.thresh)
    token_mix_mma_fwd_kernel<DM, TP, NP, true, WS><<<grid, kWarps * 32, 0, s>>>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, T, dh, dout);
  else
    token_mix_mma_fwd_kernel<DM, TP, NP, false, WS><<<grid, kWarps * 32, 0, s>>>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, T, dh, dout);
  M2_LAUNCH_CHECK();
  return M2_OK;
}
template <int DM, int TP, int NP>
int launch_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
               const float* b2, float* u, int B, int N, int T, const Drop& dh, const Drop& dout, cudaStream_t s) {
  const int ws = token_ws(DM);
  if constexpr (DM % 4 == 0)
    if (ws == 4) return launch_fwd_ws<DM, TP, NP, 4>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, T, dh, dout, s);
  if constexpr (DM % 2 == 0)
    if (ws >= 2) return launch_fwd_ws<DM, TP, NP, 2>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, T, dh, dout, s);
  return launch_fwd_ws<DM, TP, NP, 1>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, T, dh, dout, s);
}

}  // namespace

// Shapes the warp-level tensor-core path covers (bf16 mode only): the register tile is DM * (N / 8) * 4 floats per lane.
// N <= 32 (two k-steps in GEMM1 / the dG GEMM) for D <= 64: MIMIC-H's 24 time steps and 25 fused tokens.
bool token_mix_mma_supported(int N, int D, int T) {
  if (!(D == 32 || D == 64 || D == 128 || D == 256)) return false;
  if (N < 1 || N > 32 || T < 1 || T > 32) return false;
  const int kn = N <= 8 ? 1 : (N <= 16 ? 2 : 4);
  return (D / 16) * kn <= 16;
}

int token_mix_mma_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* u, int B, int N, int D, int T, float drop_p, unsigned long long seed,
                      cudaStream_t s) {
  if (!token_mix_mma_supported(N, D, T)) return M2_ERR_ARG;
  const Drop dh = make_drop(drop_p, seed, kSiteTokenHidden), dout = make_drop(drop_p, seed, kSiteTokenOut);
  const int dm = D / 16, tp = T <= 16 ? 16 : 32, np = N <= 8 ? 8 : (N <= 16 ? 16 : 32);
#define M2_TMF(DM_, TP_, NP_) \
  if (dm == DM_ && tp == TP_ && np == NP_) return launch_fwd<DM_, TP_, NP_>(x, ln_w, ln_b, w1, b1, w2, b2, u, B, N, T, dh, dout, s);
  M2_TMF(2, 16, 8) M2_TMF(2, 32, 8) M2_TMF(4, 16, 8) M2_TMF(4, 32, 8) M2_TMF(8, 16, 8) M2_TMF(8, 32, 8) M2_TMF(16, 16, 8) M2_TMF(16, 32, 8)
  M2_TMF(2, 16, 16) M2_TMF(2, 32, 16) M2_TMF(4, 16, 16) M2_TMF(4, 32, 16) M2_TMF(8, 16, 16) M2_TMF(8, 32, 16)
  M2_TMF(2, 16, 32) M2_TMF(2, 32, 32) M2_TMF(4, 16, 32) M2_TMF(4, 32, 32)
#undef M2_TMF
  return M2_ERR_ARG;
}

// dx = du + LN'(dXn) (complete: no separate ln_bwd pass); every parameter gradient is ACCUMULATED.
int token_mix_mma_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                      const float* w2, float* dx, float* dln_w, float* dln_b, float* dw1, float* db1, float* dw2, float* db2,
                      int B, int N, int D, int T, float drop_p, unsigned long long seed, cudaStream_t s) {
  if (!token_mix_mma_supported(N, D, T)) return M2_ERR_ARG;
  const Drop dh = make_drop(drop_p, seed, kSiteTokenHidden), dout = make_drop(drop_p, seed, kSiteTokenOut);
  const int dm = D / 16, tp = T <= 16 ? 16 : 32, np = N <= 8 ? 8 : (N <= 16 ? 16 : 32);
#define M2_TMB(DM_, TP_, NP_)              \
  if (dm == DM_ && tp == TP_ && np == NP_) \
    return launch_bwd<DM_, TP_, NP_>(du, x, ln_w, ln_b, w1, b1, w2, dx, dln_w, dln_b, dw1, db1, dw2, db2, B, N, T, dh, dout, s);
  M2_TMB(2, 16, 8) M2_TMB(2, 32, 8) M2_TMB(4, 16, 8) M2_TMB(4, 32, 8) M2_TMB(8, 16, 8) M2_TMB(8, 32, 8) M2_TMB(16, 16, 8) M2_TMB(16, 32, 8)
  M2_TMB(2, 16, 16) M2_TMB(2, 32, 16) M2_TMB(4, 16, 16) M2_TMB(4, 32, 16) M2_TMB(8, 16, 16) M2_TMB(8, 32, 16)
  M2_TMB(2, 16, 32) M2_TMB(2, 32, 32) M2_TMB(4, 16, 32) M2_TMB(4, 32, 32)
#undef M2_TMB
  return M2_ERR_ARG;
}

}  // namespace m2
