// Generic bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM), operands staged by TMA.
//
//   CTA tile 128 x 128 x 64, one CTA per (m-tile, n-tile, batch*split).  6 warps:
//     warp 0   : TMA producer   (one lane)   - fills a kStages-deep smem ring, full/empty mbarriers
//     warp 1   : MMA issuer     (one lane)   - 4 x tcgen05.mma (K=16 each) per stage, tcgen05.commit frees the slot
//     warps 2-5: epilogue                    - tcgen05.ld the 128x128 fp32 accumulator (warp w owns TMEM lanes
//                                              32*(w%4)..), bias / GELU / residual, store fp32 or bf16
//   Both operands may be K-major (row-major [rows][K]) or MN-major (row-major [K][rows]); the latter is what the
//   weight-gradient GEMMs (dW = dY^T . X, contraction over the token axis) and the transpose-free token-mixing
//   GEMMs need, so no transposed copy is ever materialised ("operand layouts swapped").
//   Ragged M/N/K edges rely on TMA out-of-bounds zero fill; stores are masked.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"

namespace m2 {

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr int kStages = 3;   // 96 KB of operand ring: two CTAs per SM, one's epilogue overlaps the other's main loop
constexpr int kTileBytes = kBM * kBK * 2;                  // 16 KB per operand per stage
constexpr int kSmemBytes = kStages * 2 * kTileBytes + 256 + 1024;   // + barriers + alignment slack
constexpr int kTmemCols = 128;

struct GemmDev {
  int M, N, K;
  int splitk, k_tiles_per_split;
  int a_batch_rows, b_batch_rows;
  const float* bias; int bias_mode; int act;
  const float* residual; long long ldr, r_batch_stride;
  void* C; int c_bf16; long long ldc, c_batch_stride;
  int accumulate;
  int atomic;
  Drop drop; long long drop_ld;
};

// One accumulator row x 32 columns: bias / activation / dropout / residual -> v[32] (columns past N hold garbage).
__device__ __forceinline__ void epilogue_math(const GemmDev& p, const uint32_t (&r)[32], float (&v)[32], int row, int ncol0,
                                              const float* res, float rbias, bool lead) {
  const bool full_chunk = (ncol0 + 32 <= p.N);
  const bool vec_ok = full_chunk && (p.bias_mode != 1 || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) &&
                      (!res || (reinterpret_cast<uintptr_t>(res + ncol0) & 15) == 0) && (!p.drop.thresh || (p.drop_ld & 3) == 0);
  if (vec_ok) {
    // whole chunk in range: 16-byte bias / residual loads, packed GELU, one mask hash per four columns, no predicates
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 x = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      if (lead) {
        if (p.bias_mode == 1) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + ncol0 + j));
          x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
        }
        x.x += rbias; x.y += rbias; x.z += rbias; x.w += rbias;
      }
      if (p.act == 1) {
        const float2 a = gelu2(make_float2(x.x, x.y)), b = gelu2(make_float2(x.z, x.w));
        x = make_float4(a.x, a.y, b.x, b.y);
      } else if (p.act == 2) {
        x = make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
      }
      if (p.drop.thresh) drop_apply4(p.drop, x, static_cast<unsigned long long>(row) * p.drop_ld + ncol0 + j);
      if (res) {
        const float4 rv = *reinterpret_cast<const float4*>(res + ncol0 + j);
        x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
      }
      v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = ncol0 + j;
      float x = __uint_as_float(r[j]);
      if (col < p.N) {
        if (lead) {
          if (p.bias_mode == 1) x += p.bias[col];
          x += rbias;
        }
        if (p.act == 1) x = gelu_fast(x);
        else if (p.act == 2) x = fmaxf(x, 0.f);
        if (p.drop.thresh) x = drop_apply(p.drop, x, static_cast<unsigned long long>(row) * p.drop_ld + col);
        if (res) x += res[col];
      }
      v[j] = x;
    }
  }
}

// v[32] -> global, one row per thread (16-byte stores; reductions for split-K / shared outputs)
__device__ __forceinline__ void epilogue_store(const GemmDev& p, const float (&v)[32], int ncol0, long long coff) {
  const bool full_chunk = (ncol0 + 32 <= p.N);
  if (p.c_bf16) {
    __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + coff + ncol0;
    if (full_chunk && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 o = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                             pack_bf16(v[j + 6], v[j + 7]));
        *reinterpret_cast<uint4*>(c + j) = o;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (ncol0 + j < p.N) c[j] = __float2bfloat16(v[j]);
    }
  } else {
    float* c = reinterpret_cast<float*>(p.C) + coff + ncol0;
    if (p.splitk > 1 || p.atomic) {
      if (full_chunk && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)   // 128-bit reductions (red.global.add.v4.f32)
          atomicAdd(reinterpret_cast<float4*>(c + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      } else {
        for (int j = 0; j < 32; ++j)
          if (ncol0 + j < p.N) atomicAdd(c + j, v[j]);
      }
    } else if (full_chunk && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (p.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(c + j);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(c + j) = o;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (ncol0 + j < p.N) c[j] = p.accumulate ? c[j] + v[j] : v[j];
    }
  }
}

__device__ __forceinline__ void epilogue_chunk(const GemmDev& p, const uint32_t (&r)[32], int row, bool row_ok, int ncol0,
                                               long long coff, const float* res, float rbias, bool lead) {
  if (!row_ok || ncol0 >= p.N) return;
  float v[32];
  epilogue_math(p, r, v, row, ncol0, res, rbias, lead);
  epilogue_store(p, v, ncol0, coff);
}

template <bool kAMn, bool kBMn>
__global__ void __launch_bounds__(192, 2)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTileBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kStages * kTileBytes);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  // logical roles 0 = TMA, 1 = MMA, 2-5 = epilogue; physically the epilogue warps come first (see chain_ts.cu)
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = pwarp < 4 ? pwarp + 2 : pwarp - 4;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
  const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
  const int k_tiles = ceil_div(p.K, kBK);
  const int kt_begin = split * p.k_tiles_per_split;
  const int kt_end = min(k_tiles, kt_begin + p.k_tiles_per_split);
  const int nkt = kt_end - kt_begin;
  if (nkt <= 0) return;   // uniform for the whole CTA (only possible for trailing splits)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and issuer walk their loops with the whole warp; one elected lane issues (see elect_one()).
  if (warp == 0) {
    const int a_row0 = batch * p.a_batch_rows, b_row0 = batch * p.b_batch_rows;
    for (int it = 0; it < nkt; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], 2 * kTileBytes);
        const int k0 = (kt_begin + it) * kBK;
        uint8_t* a = sA + s * kTileBytes;
        uint8_t* b = sB + s * kTileBytes;
        if (!kAMn) {
          tma_load_2d(a, &tmA, &full[s], k0, a_row0 + m0);
        } else {   // [64 k-rows][64 m] panels, two per 128-wide tile
          tma_load_2d(a, &tmA, &full[s], m0, a_row0 + k0);
          tma_load_2d(a + kTileBytes / 2, &tmA, &full[s], m0 + 64, a_row0 + k0);
        }
        if (!kBMn) {
          tma_load_2d(b, &tmB, &full[s], k0, b_row0 + n0);
        } else {
          tma_load_2d(b, &tmB, &full[s], n0, b_row0 + k0);
          tma_load_2d(b + kTileBytes / 2, &tmB, &full[s], n0 + 64, b_row0 + k0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN, kAMn ? 1 : 0, kBMn ? 1 : 0);
    // K-major: +16 elements = +32 B inside the 128-B swizzled row.  MN-major: +16 k-rows = +2048 B.
    const uint64_t a_desc0 = kAMn ? umma_desc_sw128(smem_u32(sA), kTileBytes / 2, 1024) : umma_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t b_desc0 = kBMn ? umma_desc_sw128(smem_u32(sB), kTileBytes / 2, 1024) : umma_desc_sw128(smem_u32(sB), 16, 1024);
    for (int it = 0; it < nkt; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = a_desc0 + static_cast<uint64_t>((s * kTileBytes) >> 4);
        const uint64_t bd = b_desc0 + static_cast<uint64_t>((s * kTileBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < kBK / 16; ++kk)
          umma_bf16(tmem_base, ad + ((kk * (kAMn ? 2048 : 32)) >> 4), bd + ((kk * (kBMn ? 2048 : 32)) >> 4), idesc,
                    (it > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else {
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = pwarp & 3;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const bool lead = (split == 0);
    const float rbias = (p.bias_mode == 2 && row_ok && lead) ? p.bias[row] : 0.f;
    const long long coff = static_cast<long long>(batch) * p.c_batch_stride + static_cast<long long>(row) * p.ldc;
    const float* res = (p.residual && lead) ? p.residual + static_cast<long long>(batch) * p.r_batch_stride +
                                                  static_cast<long long>(row) * p.ldr : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < kBN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
      tmem_ld_wait();
      epilogue_chunk(p, r, row, row_ok, n0 + c0, coff, res, rbias, lead);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

template <bool kAMn, bool kBMn>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& d, dim3 grid, cudaStream_t s) {
  auto kern = umma_gemm_kernel<kAMn, kBMn>;
  static bool configured = false;   // immutable one-time attribute (idempotent; benign if raced)
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return M2_ERR_LAUNCH;
    configured = true;
  }
  LaunchScope scope(kAMn ? (kBMn ? "umma_gemm_tn" : "umma_gemm_tk") : (kBMn ? "umma_gemm_kn" : "umma_gemm_kk"), s);
  kern<<<grid, 192, kSmemBytes, s>>>(ta, tb, d);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Generation 2: PERSISTENT CTAs, 128 x 256 x 64 tiles, two accumulator buffers in tensor memory.
//
// The kernel above is one tile per CTA: TMEM allocation, barrier set-up, the pipeline fill and the 64 KB epilogue of
// every 128 x 128 tile are serial with its (often short: K = 768 is 12 k-steps) main loop, and a 128 x 128 x 64 step needs
// 32 KB of operands for 256 clk of tensor work = 128 B/clk/SM out of L2, which is what bounds it (BASELINE configs 4 / 5 ran
// at 16 % of the bf16 peak on it).  Here one CTA per SM walks tiles (m fastest: the CTAs of a wave share the B panel):
//   warp 0     TMA producer: a 4-stage ring of (16 KB A + 32 KB B), filled across tile boundaries
//   warp 1     MMA issuer:   4 x tcgen05.mma M128 N256 K16 per stage into accumulator buffer (tile & 1)
//   warps 2-17 epilogue:     buffer (tile & 1) -> registers -> bias / GELU / dropout / residual -> global, while the
//                            MMA warp fills the other buffer (acc_full / acc_empty barriers); four warps per TMEM lane
//                            quadrant, 64 columns each
// 48 KB of operands per 512 clk = 94 B/clk/SM.  Same operand layouts, batching, split-K and epilogue as generation 1.
constexpr int kBN2 = 256;
// four ring stages, or three + 64 KB of per-warp output staging for the TMA-store variant (short K: the epilogue bounds it)
constexpr int kTileA2 = kBM * kBK * 2;                     // 16 KB
constexpr int kTileB2 = kBN2 * kBK * 2;                    // 32 KB
constexpr int kStageOut2 = 4096;                           // one [32 rows][128 B] swizzled box per epilogue warp
constexpr int smem2_bytes(bool tma_store) { return (tma_store ? 3 : 4) * (kTileA2 + kTileB2) + (tma_store ? 16 * kStageOut2 : 0) + 256 + 1024; }
constexpr int kEpiWarps2 = 16;               // four per TMEM lane quadrant, 64 accumulator columns each
constexpr int kThreads2 = (2 + kEpiWarps2) * 32;

struct Sched2 { int tiles_m, tiles_n, total; };

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// kTmaStore: the output tile leaves through shared memory and cp.async.bulk.tensor stores (plain overwrite outputs only).
// One row per thread is what TMEM hands out, and 16-byte global stores from it scatter over 32 rows per instruction: 8 k L2
// sector writes per fp32 tile, more than the 6 k clk of MMAs a K = 768 tile has (the "no epilogue work" GEMM ran at 53 % of
// the rate of the 8192^3 one).  Each epilogue warp stages its [32 rows x 128 B] piece in the 128-byte swizzle and one lane
// stores the box; ragged edges are clipped by the tensor map.
template <bool kAMn, bool kBMn, bool kTmaStore>
__global__ void __launch_bounds__(kThreads2, 1)
umma_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const GemmDev p, const Sched2 sc) {
  constexpr int kStages2 = kTmaStore ? 3 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages2 * kTileA2;
  uint8_t* sOut = smem + kStages2 * (kTileA2 + kTileB2);      // [16 warps][4 KB], 1024-byte aligned (TMA-store variant only)
  uint64_t* full = reinterpret_cast<uint64_t*>(sOut + (kTmaStore ? 16 * kStageOut2 : 0));
  uint64_t* empty = full + kStages2;
  uint64_t* acc_full = empty + kStages2;      // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_tiles = ceil_div(p.K, kBK);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps2); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (kTmaStore) tma_prefetch_desc(&tmC);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_z = sc.tiles_m * sc.tiles_n;

  if (warp == 0) {
    int it = 0;
    for (int tile = blockIdx.x; tile < sc.total; tile += gridDim.x) {
      const int z = tile / per_z, r = tile - z * per_z;
      const int m0 = (r % sc.tiles_m) * kBM, n0 = (r / sc.tiles_m) * kBN2;
      const int batch = z / p.splitk, split = z - batch * p.splitk;
      const int kt_begin = split * p.k_tiles_per_split, kt_end = min(k_tiles, kt_begin + p.k_tiles_per_split);
      const int a_row0 = batch * p.a_batch_rows, b_row0 = batch * p.b_batch_rows;
      for (int kt = kt_begin; kt < kt_end; ++kt, ++it) {
        const int s = it % kStages2;
        mbar_wait(&empty[s], ((it / kStages2) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], kTileA2 + kTileB2);
          const int k0 = kt * kBK;
          uint8_t* a = sA + s * kTileA2;
          uint8_t* b = sB + s * kTileB2;
          if (!kAMn) {
            tma_load_2d(a, &tmA, &full[s], k0, a_row0 + m0);
          } else {
            tma_load_2d(a, &tmA, &full[s], m0, a_row0 + k0);
            tma_load_2d(a + kTileA2 / 2, &tmA, &full[s], m0 + 64, a_row0 + k0);
          }
          if (!kBMn) {
            tma_load_2d(b, &tmB, &full[s], k0, b_row0 + n0);            // [256 n][64 k] in one box
          } else {
#pragma unroll
            for (int pnl = 0; pnl < 4; ++pnl) tma_load_2d(b + pnl * (kTileB2 / 4), &tmB, &full[s], n0 + 64 * pnl, b_row0 + k0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN2, kAMn ? 1 : 0, kBMn ? 1 : 0);
    const uint64_t a_desc0 = kAMn ? umma_desc_sw128(smem_u32(sA), kTileA2 / 2, 1024) : umma_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t b_desc0 = kBMn ? umma_desc_sw128(smem_u32(sB), kTileB2 / 4, 1024) : umma_desc_sw128(smem_u32(sB), 16, 1024);
    int it = 0, t = 0;
    for (int tile = blockIdx.x; tile < sc.total; tile += gridDim.x, ++t) {
      const int z = tile / per_z;
      const int split = z % p.splitk;
      const int kt_begin = split * p.k_tiles_per_split, kt_end = min(k_tiles, kt_begin + p.k_tiles_per_split);
      const int buf = t & 1;
      mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);     // the epilogue has drained this buffer (tile t - 2)
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * kBN2;
      for (int kt = kt_begin; kt < kt_end; ++kt, ++it) {
        const int s = it % kStages2;
        mbar_wait(&full[s], (it / kStages2) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = a_desc0 + static_cast<uint64_t>((s * kTileA2) >> 4);
          const uint64_t bd = b_desc0 + static_cast<uint64_t>((s * kTileB2) >> 4);
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk)
            umma_bf16(acc, ad + ((kk * (kAMn ? 2048 : 32)) >> 4), bd + ((kk * (kBMn ? 2048 : 32)) >> 4), idesc,
                      (kt > kt_begin || kk > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                       // TMEM lane quadrant of this warp
    const int part = (warp - 2) >> 2;             // which 64 of the 256 accumulator columns
    int t = 0;
    for (int tile = blockIdx.x; tile < sc.total; tile += gridDim.x, ++t) {
      const int z = tile / per_z, r = tile - z * per_z;
      const int m0 = (r % sc.tiles_m) * kBM, n0 = (r / sc.tiles_m) * kBN2;
      const int batch = z / p.splitk, split = z - batch * p.splitk;
      const int buf = t & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool lead = (split == 0);
      const float rbias = (p.bias_mode == 2 && row_ok && lead) ? p.bias[row] : 0.f;
      const long long coff = static_cast<long long>(batch) * p.c_batch_stride + static_cast<long long>(row) * p.ldc;
      const float* res = (p.residual && lead) ? p.residual + static_cast<long long>(batch) * p.r_batch_stride +
                                                    static_cast<long long>(row) * p.ldr : nullptr;
      mbar_wait(&acc_full[buf], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t stage = smem_u32(sOut + (warp - 2) * kStageOut2);
#pragma unroll 1
      for (int c0 = part * 64; c0 < part * 64 + 64; c0 += 32) {
        uint32_t rg[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kBN2 + c0, rg);
        tmem_ld_wait();
        if (c0 + 32 == part * 64 + 64) {          // last piece in registers: hand the buffer back before the math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (!kTmaStore) {
          epilogue_chunk(p, rg, row, row_ok, n0 + c0, coff, res, rbias, lead);
        } else {
          const bool live = n0 + c0 < p.N;        // warp-uniform
          float v[32];
          if (live && row_ok) epilogue_math(p, rg, v, row, n0 + c0, res, rbias, lead);
          if (p.c_bf16) {
            // the warp's two 32-column pieces are the halves of ONE [32 x 64] bf16 box
            const int hpc = (c0 - part * 64) >> 5;
            if (hpc == 0) {                       // the previous tile's store must have read the staging buffer
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
            }
            if (live && row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 o = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                                           pack_bf16(v[j + 6], v[j + 7]));
                st_shared_v4(stage + sw128_offset(lane, hpc * 4 + (j >> 3)), o);
              }
            }
            if (hpc == 1) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0 && n0 + part * 64 < p.N) {
                tma_store_3d(&tmC, sOut + (warp - 2) * kStageOut2, n0 + part * 64, m0 + q * 32, batch);
                tma_store_commit();
              }
            }
          } else {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
            if (live && row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                st_shared_v4(stage + sw128_offset(lane, j >> 2),
                             make_uint4(__float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]), __float_as_uint(v[j + 3])));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && live) {
              tma_store_3d(&tmC, sOut + (warp - 2) * kStageOut2, n0 + c0, m0 + q * 32, batch);
              tma_store_commit();
            }
          }
        }
      }
    }
    if (kTmaStore) {                              // the bulk stores must have left shared memory before the CTA exits
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <bool kAMn, bool kBMn, bool kTmaStore>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmDev& d, const Sched2& sc, cudaStream_t s) {
  auto kern = umma_gemm2_kernel<kAMn, kBMn, kTmaStore>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2_bytes(kTmaStore)) != cudaSuccess) return M2_ERR_LAUNCH;
    configured = true;
  }
  LaunchScope scope(kAMn ? (kBMn ? "umma_gemm2_tn" : "umma_gemm2_tk") : (kBMn ? "umma_gemm2_kn" : "umma_gemm2_kk"), s);
  const int grid = sc.total < 148 ? sc.total : 148;
  kern<<<grid, kThreads2, smem2_bytes(kTmaStore), s>>>(ta, tb, tc, d, sc);
  M2_LAUNCH_CHECK();
  return M2_OK;
}
template <bool kTmaStore>
int launch2_layout(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmDev& d,
                   const Sched2& sc, cudaStream_t s) {
  if (a_mn) return b_mn ? launch2<true, true, kTmaStore>(ta, tb, tc, d, sc, s) : launch2<true, false, kTmaStore>(ta, tb, tc, d, sc, s);
  return b_mn ? launch2<false, true, kTmaStore>(ta, tb, tc, d, sc, s) : launch2<false, false, kTmaStore>(ta, tb, tc, d, sc, s);
}


// ---------------------------------------------------------------------------------------------------------------
// Generation 3: the same persistent kernel on CTA PAIRS (thread-block cluster of 2 = the two SMs of a TPC,
// tcgen05.mma.cta_group::2).  A pair owns a 256 x 256 tile: CTA r loads ITS 128 rows of A and ITS 128 columns of B per
// k-step (32 KB instead of the 48 KB a single CTA needs for 128 x 256: 64 instead of 94 B/clk/SM out of L2), the leader's
// one MMA thread issues M256 N256 K16 instructions that read both CTAs' shared memory and write both CTAs' tensor memory
// (each CTA keeps the accumulator rows of its own A half: 128 lanes x 256 columns, two buffers), and each CTA runs the
// epilogue of its half.  Barriers: the leader's `full` counts the bytes of BOTH CTAs' TMA loads (cp.async.bulk.tensor
// .cta_group::2 with the barrier address of the even CTA), `empty` / `acc_full` are signalled in both CTAs by multicast
// commits, the peer's epilogue warps arrive on the leader's `acc_empty` through the cluster address of its barrier.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared-window address: the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {   // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// box -> THIS CTA's shared memory, bytes counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

constexpr int kTileBh = kTileB2 / 2;                       // this CTA's 128 columns of B: 16 KB
constexpr int smem2c_bytes(bool tma_store) {
  return (tma_store ? 5 : 6) * (kTileA2 + kTileBh) + (tma_store ? 16 * kStageOut2 : 0) + 256 + 1024;
}

template <bool kAMn, bool kBMn, bool kTmaStore>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
umma_gemm2c_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const GemmDev p, const Sched2 sc) {
  constexpr int kStagesC = kTmaStore ? 5 : 6;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStagesC * kTileA2;
  uint8_t* sOut = smem + kStagesC * (kTileA2 + kTileBh);
  uint64_t* full = reinterpret_cast<uint64_t*>(sOut + (kTmaStore ? 16 * kStageOut2 : 0));
  uint64_t* empty = full + kStagesC;
  uint64_t* acc_full = empty + kStagesC;      // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]  (the leader's are the live ones)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int k_tiles = ceil_div(p.K, kBK);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStagesC; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * kEpiWarps2); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (kTmaStore) tma_prefetch_desc(&tmC);
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();           // both CTAs' barriers exist before any remote arrive / remote transaction count
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_z = sc.tiles_m * sc.tiles_n;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {
    int it = 0;
    for (int tile = pair; tile < sc.total; tile += npairs) {
      const int z = tile / per_z, r = tile - z * per_z;
      const int m0 = (r % sc.tiles_m) * 256 + rank * 128, n0 = (r / sc.tiles_m) * 256 + rank * 128;
      const int batch = z / p.splitk, split = z - batch * p.splitk;
      const int kt_begin = split * p.k_tiles_per_split, kt_end = min(k_tiles, kt_begin + p.k_tiles_per_split);
      const int a_row0 = batch * p.a_batch_rows, b_row0 = batch * p.b_batch_rows;
      for (int kt = kt_begin; kt < kt_end; ++kt, ++it) {
        const int s = it % kStagesC;
        mbar_wait(&empty[s], ((it / kStagesC) & 1) ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * (kTileA2 + kTileBh));   // both CTAs' boxes land on this barrier
          const int k0 = kt * kBK;
          uint8_t* a = sA + s * kTileA2;
          uint8_t* b = sB + s * kTileBh;
          if (!kAMn) {
            tma_load_2d_pair(a, &tmA, &full[s], k0, a_row0 + m0);
          } else {
            tma_load_2d_pair(a, &tmA, &full[s], m0, a_row0 + k0);
            tma_load_2d_pair(a + kTileA2 / 2, &tmA, &full[s], m0 + 64, a_row0 + k0);
          }
          if (!kBMn) {
            tma_load_2d_pair(b, &tmB, &full[s], k0, b_row0 + n0);              // [128 n][64 k]
          } else {
            tma_load_2d_pair(b, &tmB, &full[s], n0, b_row0 + k0);
            tma_load_2d_pair(b + kTileBh / 2, &tmB, &full[s], n0 + 64, b_row0 + k0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, kBN2, kAMn ? 1 : 0, kBMn ? 1 : 0);
      const uint64_t a_desc0 = kAMn ? umma_desc_sw128(smem_u32(sA), kTileA2 / 2, 1024) : umma_desc_sw128(smem_u32(sA), 16, 1024);
      const uint64_t b_desc0 = kBMn ? umma_desc_sw128(smem_u32(sB), kTileBh / 2, 1024) : umma_desc_sw128(smem_u32(sB), 16, 1024);
      int it = 0, t = 0;
      for (int tile = pair; tile < sc.total; tile += npairs, ++t) {
        const int z = tile / per_z;
        const int split = z % p.splitk;
        const int kt_begin = split * p.k_tiles_per_split, kt_end = min(k_tiles, kt_begin + p.k_tiles_per_split);
        const int buf = t & 1;
        mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);     // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * kBN2;
        for (int kt = kt_begin; kt < kt_end; ++kt, ++it) {
          const int s = it % kStagesC;
          mbar_wait(&full[s], (it / kStagesC) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = a_desc0 + static_cast<uint64_t>((s * kTileA2) >> 4);
            const uint64_t bd = b_desc0 + static_cast<uint64_t>((s * kTileBh) >> 4);
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk)
              umma_bf16_pair(acc, ad + ((kk * (kAMn ? 2048 : 32)) >> 4), bd + ((kk * (kBMn ? 2048 : 32)) >> 4), idesc,
                             (kt > kt_begin || kk > 0) ? 1u : 0u);
            umma_commit_pair(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_pair(&acc_full[buf]);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    int t = 0;
    for (int tile = pair; tile < sc.total; tile += npairs, ++t) {
      const int z = tile / per_z, r = tile - z * per_z;
      const int m0 = (r % sc.tiles_m) * 256 + rank * 128, n0 = (r / sc.tiles_m) * 256;
      const int batch = z / p.splitk, split = z - batch * p.splitk;
      const int buf = t & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool lead = (split == 0);
      const float rbias = (p.bias_mode == 2 && row_ok && lead) ? p.bias[row] : 0.f;
      const long long coff = static_cast<long long>(batch) * p.c_batch_stride + static_cast<long long>(row) * p.ldc;
      const float* res = (p.residual && lead) ? p.residual + static_cast<long long>(batch) * p.r_batch_stride +
                                                    static_cast<long long>(row) * p.ldr : nullptr;
      mbar_wait(&acc_full[buf], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t stage = smem_u32(sOut + (warp - 2) * kStageOut2);
#pragma unroll 1
      for (int c0 = part * 64; c0 < part * 64 + 64; c0 += 32) {
        uint32_t rg[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kBN2 + c0, rg);
        tmem_ld_wait();
        if (c0 + 32 == part * 64 + 64) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);
        }
        if (!kTmaStore) {
          epilogue_chunk(p, rg, row, row_ok, n0 + c0, coff, res, rbias, lead);
        } else {
          const bool live = n0 + c0 < p.N;
          float v[32];
          if (live && row_ok) epilogue_math(p, rg, v, row, n0 + c0, res, rbias, lead);
          if (p.c_bf16) {
            const int hpc = (c0 - part * 64) >> 5;
            if (hpc == 0) {
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
            }
            if (live && row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 o = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                                           pack_bf16(v[j + 6], v[j + 7]));
                st_shared_v4(stage + sw128_offset(lane, hpc * 4 + (j >> 3)), o);
              }
            }
            if (hpc == 1) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0 && n0 + part * 64 < p.N) {
                tma_store_3d(&tmC, sOut + (warp - 2) * kStageOut2, n0 + part * 64, m0 + q * 32, batch);
                tma_store_commit();
              }
            }
          } else {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
            if (live && row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                st_shared_v4(stage + sw128_offset(lane, j >> 2),
                             make_uint4(__float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]), __float_as_uint(v[j + 3])));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && live) {
              tma_store_3d(&tmC, sOut + (warp - 2) * kStageOut2, n0 + c0, m0 + q * 32, batch);
              tma_store_commit();
            }
          }
        }
      }
    }
    if (kTmaStore) {
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
    }
  }

  tc_fence_before();
  cluster_sync_all();           // the peer's shared memory and barriers stay alive until both CTAs are done
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

template <bool kAMn, bool kBMn, bool kTmaStore>
int launch2c(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmDev& d, const Sched2& sc, cudaStream_t s) {
  auto kern = umma_gemm2c_kernel<kAMn, kBMn, kTmaStore>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2c_bytes(kTmaStore)) != cudaSuccess) return M2_ERR_LAUNCH;
    configured = true;
  }
  LaunchScope scope(kAMn ? (kBMn ? "umma_gemm2c_tn" : "umma_gemm2c_tk") : (kBMn ? "umma_gemm2c_kn" : "umma_gemm2c_kk"), s);
  int grid = 2 * sc.total < 148 ? 2 * sc.total : 148;
  kern<<<grid, kThreads2, smem2c_bytes(kTmaStore), s>>>(ta, tb, tc, d, sc);
  M2_LAUNCH_CHECK();
  return M2_OK;
}
template <bool kTmaStore>
int launch2c_layout(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmDev& d,
                    const Sched2& sc, cudaStream_t s) {
  if (a_mn) return b_mn ? launch2c<true, true, kTmaStore>(ta, tb, tc, d, sc, s) : launch2c<true, false, kTmaStore>(ta, tb, tc, d, sc, s);
  return b_mn ? launch2c<false, true, kTmaStore>(ta, tb, tc, d, sc, s) : launch2c<false, false, kTmaStore>(ta, tb, tc, d, sc, s);
}

// M2B200_GEMM_GEN=1 keeps every GEMM on the one-tile-per-CTA kernel, 2 on single-CTA persistent tiles (A/B measurements)
int gemm_generation() {
  static const int gen = [] {
    const char* e = getenv("M2B200_GEMM_GEN");
    return e ? atoi(e) : 3;
  }();
  return gen;
}

}  // namespace

int gemm_bf16_umma(const GemmArgs& g, cudaStream_t s) {
  if (!g.A || !g.B || !g.C || g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batch <= 0) return M2_ERR_ARG;
  if (g.splitk < 1 || (g.splitk > 1 && (g.c_bf16 || g.act))) return M2_ERR_ARG;
  if (g.atomic_out && (g.c_bf16 || g.act || g.bias_mode || g.residual)) return M2_ERR_ARG;
  if (g.bias_mode && !g.bias) return M2_ERR_ARG;
  // wide persistent tiles when the output is wide enough to fill them and there is more than a wave of work
  const bool wide = gemm_generation() >= 2 && g.N > 128 &&
                    static_cast<long long>(ceil_div(g.M, kBM)) * ceil_div(g.N, kBN2) * g.batch * g.splitk >= 64;
  // CTA pairs (256 x 256 tiles) when at least one wave of pairs exists
  const bool pairs = wide && gemm_generation() >= 3 && g.M > 128 &&
                     static_cast<long long>(ceil_div(g.M, 256)) * ceil_div(g.N, kBN2) * g.batch * g.splitk >= 74;
  CUtensorMap ta, tb;
  int rc;
  {
    const uint64_t ext = g.a_mn ? g.K : g.M, cols = g.a_mn ? g.M : g.K;
    const uint64_t rows = static_cast<uint64_t>(g.batch - 1) * g.a_batch_rows + ext;
    rc = make_tmap_bf16(&ta, g.A, rows, cols, g.lda, g.a_mn ? 64 : 128, 64);
    if (rc) return rc;
  }
  {
    const uint64_t ext = g.b_mn ? g.K : g.N, cols = g.b_mn ? g.N : g.K;
    const uint64_t rows = static_cast<uint64_t>(g.batch - 1) * g.b_batch_rows + ext;
    rc = make_tmap_bf16(&tb, g.B, rows, cols, g.ldb, g.b_mn ? 64 : ((wide && !pairs) ? 256 : 128), 64);
    if (rc) return rc;
  }
  GemmDev d;
  d.M = g.M; d.N = g.N; d.K = g.K;
  const int k_tiles = ceil_div(g.K, kBK);
  d.splitk = g.splitk > k_tiles ? k_tiles : g.splitk;
  d.k_tiles_per_split = ceil_div(k_tiles, d.splitk);
  d.splitk = ceil_div(k_tiles, d.k_tiles_per_split);   // no empty trailing split
  d.a_batch_rows = static_cast<int>(g.a_batch_rows); d.b_batch_rows = static_cast<int>(g.b_batch_rows);
  d.bias = g.bias; d.bias_mode = g.bias_mode; d.act = g.act;
  d.residual = g.residual; d.ldr = g.ldr; d.r_batch_stride = g.r_batch_stride;
  d.C = g.C; d.c_bf16 = g.c_bf16; d.ldc = g.ldc; d.c_batch_stride = g.c_batch_stride;
  d.accumulate = g.accumulate;
  d.atomic = (g.atomic_out && !g.c_bf16) ? 1 : 0;
  d.drop = make_drop(g.drop_p, g.drop_seed, g.drop_site); d.drop_ld = g.drop_ld;
  if (wide) {
    Sched2 sc;
    sc.tiles_m = ceil_div(g.M, pairs ? 256 : kBM); sc.tiles_n = ceil_div(g.N, kBN2);
    const long long total = static_cast<long long>(sc.tiles_m) * sc.tiles_n * g.batch * d.splitk;
    if (total >= (1ll << 31)) return M2_ERR_ARG;
    sc.total = static_cast<int>(total);
    // plain overwrite outputs leave through TMA stores (M2B200_GEMM_TMA_STORE=0: per-thread stores, A/B measurements)
    static const bool tma_env = [] { const char* e = getenv("M2B200_GEMM_TMA_STORE"); return !e || atoi(e) != 0; }();
    CUtensorMap tc = ta;
    // short K: the tile's epilogue, not its main loop, is what the SM waits for - worth a ring stage (measured: tools/bench_gemm.py)
    bool tma_store = tma_env && d.splitk == 1 && !d.atomic && !d.accumulate && k_tiles <= 24;
    if (tma_store) {
      const int eb = g.c_bf16 ? 2 : 4;
      if (make_tmap_store3d(&tc, g.C, eb, g.batch, g.M, g.N, g.ldc, g.c_batch_stride, 32, g.c_bf16 ? 64 : 32) != M2_OK)
        tma_store = false;   // unaligned output: per-thread stores
    }
    if (pairs)
      return tma_store ? launch2c_layout<true>(g.a_mn, g.b_mn, ta, tb, tc, d, sc, s) : launch2c_layout<false>(g.a_mn, g.b_mn, ta, tb, tc, d, sc, s);
    return tma_store ? launch2_layout<true>(g.a_mn, g.b_mn, ta, tb, tc, d, sc, s) : launch2_layout<false>(g.a_mn, g.b_mn, ta, tb, tc, d, sc, s);
  }
  dim3 grid(ceil_div(g.M, kBM), ceil_div(g.N, kBN), g.batch * d.splitk);
  if (grid.y > 65535 || grid.z > 65535) return M2_ERR_ARG;
  if (g.a_mn) return g.b_mn ? launch<true, true>(ta, tb, d, grid, s) : launch<true, false>(ta, tb, d, grid, s);
  return g.b_mn ? launch<false, true>(ta, tb, d, grid, s) : launch<false, false>(ta, tb, d, grid, s);
}

}  // namespace m2
