// Patch embedding straight from the image: Conv2d(cin, D, k = stride = P) + 'b c h w -> b (h w) c'
// (reference modules/mixer.py:143-146) as ONE tcgen05 GEMM per direction whose image-side operand tile is gathered from
// the [B][cin][H][W] pixels by the CTA itself - no im2col buffer in HBM ("cols" of the round-1 design: 103 MB written and
// read twice per step for the AV-MNIST audio branch), and the pixels may arrive as fp32 or as bf16 (what
// data.DevicePrefetcher stages when asked to halve the host link traffic; bf16 is what the operand is rounded to anyway).
//
//   forward   y[m][d]  = sum_k pix[m][k] W[d][k] + bias[d]        m = (b, gy, gx),  k = (c, py, px)
//             A tile [128 m][64 k]  K-major, 128-B swizzle: 16-byte chunks of 8 consecutive px gathered by 8 warps
//             B tile [128 d][64 k]  K-major: bf16 weight copy by TMA
//   backward  dW[d][k] += sum_m dY[m][d] pix[m][k]                 (the image needs no gradient)
//             A tile [64 m][128 d]  MN-major: bf16 dY by TMA (two 64-wide panels)
//             B tile [64 m][128 k]  MN-major: the same gather, rows = K index; split over m, fp32 vector atomics
//
// A chunk is 8 px of one patch row: P % 8 == 0 keeps it inside the row and 16-byte aligned (W = gw P).  Other patch
// sizes (AV-MNIST's 28 x 28 image, P = 14: 12.8 MB per step) take the gather + GEMM fallback of abi.cu.
// The gathered tile is written with st.shared (generic proxy) and read by tcgen05.mma (async proxy): every producer
// thread fences (fence.proxy.async) before its warp's elected lane arrives on the stage's full barrier.
#include "common.cuh"
#include "kernels.h"
#include "tmap.cuh"

namespace m2 {
namespace {

constexpr int kStages = 4;
constexpr int kTile = 128 * 64 * 2;   // 16 KB per operand per stage
constexpr int kSmem = kStages * 2 * kTile + 256 + 1024;
constexpr int kTmemCols = 128;
constexpr int kGatherWarps = 8;            // + one TMA warp + one MMA warp
constexpr int kThreads = (kGatherWarps + 2) * 32;

struct PatchDev {
  const void* img;
  int B, cin, H, W, P, gh, gw;
  int M, D, K;
  const float* bias;
  float* out;            // y [M][D] (forward) / dW [D][K] (backward)
  int splits, kt_per_split;
};

template <typename TIn>
struct Chunk;            // 8 consecutive pixels -> one 16-byte bf16 chunk
template <>
struct Chunk<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ uint4 pack() const {
    return make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
};
template <>
struct Chunk<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ uint4 pack() const { return v; }
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// element offset of pixel (row m, k = 0) / of k inside a patch
__device__ __forceinline__ long long row_offset(const PatchDev& p, int m) {
  const int np = p.gh * p.gw;
  const int b = m / np, r = m - b * np;
  const int gy = r / p.gw, gx = r - gy * p.gw;
  return (static_cast<long long>(b) * p.cin * p.H + gy * p.P) * p.W + gx * p.P;
}
__device__ __forceinline__ int k_offset(const PatchDev& p, int k) {
  const int pp = p.P * p.P;
  const int c = k / pp, r = k - c * pp;
  const int py = r / p.P, px = r - py * p.P;
  return (c * p.H + py) * p.W + px;
}

// ------------------------------------------------------------------------------------------------- forward
// grid (ceil(M / 128), ceil(D / 128)); warps 0-7 gather + epilogue, warp 8 weight TMA, warp 9 MMA issue + TMEM owner.
template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1)
patch_embed_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const PatchDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTile;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kStages * kTile);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int nkt = p.K / 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], kGatherWarps + 1); mbar_init(&empty[i], 1); }   // 4 gather warps + the TMA issue
    mbar_init(acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
  }
  if (warp == kGatherWarps + 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kGatherWarps) {
    // chunk column c of the 64-k tile is fixed per thread, rows r = t / 8 + 32 i
    const int t = threadIdx.x, c = t & 7, r0 = t >> 3;
    const TIn* img = static_cast<const TIn*>(p.img);
    long long roff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) roff[i] = row_offset(p, min(m0 + r0 + 32 * i, p.M - 1));
    const uint32_t a_base = smem_u32(sA);
    // THREE stages of loads in flight per thread (a stage is 32 KB of fp32 pixels per CTA; HBM needs ~50 KB per SM)
    auto load = [&](Chunk<TIn> (&v)[4], int it) {
      const int koff = k_offset(p, it * 64 + 8 * c);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i].load(img + roff[i] + koff);
    };
    auto store = [&](const Chunk<TIn> (&v)[4], int it) {
      const int s = it % kStages;
      mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) st_shared_v4(a_base + s * kTile + sw128_offset(r0 + 32 * i, c), v[i].pack());
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    };
    Chunk<TIn> v0[4], v1[4], v2[4];
    load(v0, 0);
    if (nkt > 1) load(v1, 1);
    for (int it = 0; it < nkt; it += 3) {
      if (it + 2 < nkt) load(v2, it + 2);
      store(v0, it);
      if (it + 3 < nkt) load(v0, it + 3);
      if (it + 1 < nkt) store(v1, it + 1);
      if (it + 4 < nkt) load(v1, it + 4);
      if (it + 2 < nkt) store(v2, it + 2);
    }
    // ---- epilogue: thread = accumulator row
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3;                       // TMEM lane quadrant; warps q and q + 4 share it, 64 columns each
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 64; c0 < (warp >> 2) * 64 + 64; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
      tmem_ld_wait();
      if (row >= p.M || n0 + c0 >= p.D) continue;
      float* y = p.out + static_cast<long long>(row) * p.D + n0 + c0;
      if (n0 + c0 + 32 <= p.D && (p.D & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          if (p.bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          }
          *reinterpret_cast<float4*>(y + j) = o;
        }
      } else {
        for (int j = 0; j < 32; ++j)
          if (n0 + c0 + j < p.D) y[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n0 + c0 + j] : 0.f);
      }
    }
  } else if (warp == kGatherWarps) {
    for (int it = 0; it < nkt; ++it) {
      const int s = it % kStages;
      mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], kTile);
        tma_load_2d(sB + s * kTile, &tmW, &full[s], it * 64, n0);
      }
      __syncwarp();
    }
  } else {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB), 16, 1024);
    for (int it = 0; it < nkt; ++it) {
      const int s = it % kStages;
      mbar_wait(&full[s], (it / kStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = a_desc0 + static_cast<uint64_t>((s * kTile) >> 4);
        const uint64_t bd = b_desc0 + static_cast<uint64_t>((s * kTile) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, ad + ((kk * 32) >> 4), bd + ((kk * 32) >> 4), idesc, (it > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kGatherWarps + 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------- weight gradient
// grid (ceil(D / 128), ceil(K / 128), splits over the M / 64 row tiles).  dW is ACCUMULATED (fp32 vector atomics).
template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1)
patch_embed_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const PatchDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kTile;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kStages * kTile);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int kt_all = ceil_div(p.M, 64);
  const int kt0 = blockIdx.z * p.kt_per_split;
  const int nkt = min(kt_all, kt0 + p.kt_per_split) - kt0;
  if (nkt <= 0) return;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], kGatherWarps + 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmDy);
  }
  if (warp == kGatherWarps + 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kGatherWarps) {
    // pixel chunk c (16 per 128-px tile: panel c / 8, chunk c % 8) is fixed per thread, K rows r = t / 16 + 16 i
    const int t = threadIdx.x, c = t & 15, r0 = t >> 4;
    const TIn* img = static_cast<const TIn*>(p.img);
    const int koff = k_offset(p, min(n0 + 8 * c, p.K - 8));
    const uint32_t b_base = smem_u32(sB) + (c >> 3) * (kTile / 2);
    auto load = [&](Chunk<TIn> (&v)[4], int it) {
      const int mrow = (kt0 + it) * 64 + r0;
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i].load(img + row_offset(p, min(mrow + 16 * i, p.M - 1)) + koff);
    };
    auto store = [&](const Chunk<TIn> (&v)[4], int it) {
      const int s = it % kStages;
      mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) st_shared_v4(b_base + s * kTile + sw128_offset(r0 + 16 * i, c & 7), v[i].pack());
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    };
    Chunk<TIn> v0[4], v1[4], v2[4];
    load(v0, 0);
    if (nkt > 1) load(v1, 1);
    for (int it = 0; it < nkt; it += 3) {
      if (it + 2 < nkt) load(v2, it + 2);
      store(v0, it);
      if (it + 3 < nkt) load(v0, it + 3);
      if (it + 1 < nkt) store(v1, it + 1);
      if (it + 4 < nkt) load(v1, it + 4);
      if (it + 2 < nkt) store(v2, it + 2);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int d = d0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 64; c0 < (warp >> 2) * 64 + 64; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
      tmem_ld_wait();
      if (d >= p.D || n0 + c0 >= p.K) continue;
      float* o = p.out + static_cast<long long>(d) * p.K + n0 + c0;
      if (n0 + c0 + 32 <= p.K && (p.K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                  __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
      } else {
        for (int j = 0; j < 32; ++j)
          if (n0 + c0 + j < p.K) atomicAdd(o + j, __uint_as_float(r[j]));
      }
    }
  } else if (warp == kGatherWarps) {
    for (int it = 0; it < nkt; ++it) {
      const int s = it % kStages;
      mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], kTile);
        const int mk = (kt0 + it) * 64;
        tma_load_2d(sA + s * kTile, &tmDy, &full[s], d0, mk);                  // rows past M read as zero
        tma_load_2d(sA + s * kTile + kTile / 2, &tmDy, &full[s], d0 + 64, mk);
      }
      __syncwarp();
    }
  } else {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 1, 1);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA), kTile / 2, 1024);
    const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB), kTile / 2, 1024);
    for (int it = 0; it < nkt; ++it) {
      const int s = it % kStages;
      mbar_wait(&full[s], (it / kStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = a_desc0 + static_cast<uint64_t>((s * kTile) >> 4);
        const uint64_t bd = b_desc0 + static_cast<uint64_t>((s * kTile) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, ad + ((kk * 2048) >> 4), bd + ((kk * 2048) >> 4), idesc, (it > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kGatherWarps + 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <typename K>
int configure(K kern) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) == cudaSuccess ? M2_OK : M2_ERR_LAUNCH;
}

PatchDev make_dev(const void* img, int B, int cin, int H, int W, int P, int D) {
  PatchDev d{};
  d.img = img; d.B = B; d.cin = cin; d.H = H; d.W = W; d.P = P; d.gh = H / P; d.gw = W / P;
  d.M = B * d.gh * d.gw; d.D = D; d.K = cin * P * P;
  return d;
}

}  // namespace

bool patch_gemm_supported(const void* img, int img_bf16, int B, int cin, int H, int W, int P) {
  if (B <= 0 || cin <= 0 || P <= 0 || H % P || W % P || P % 8) return false;
  if (reinterpret_cast<uintptr_t>(img) & 15) return false;
  (void)img_bf16;
  return static_cast<long long>(B) * (H / P) * (W / P) < (1ll << 31) - 256;
}

int patch_gemm_fwd(const void* img, int img_bf16, const void* w_bf16, int ldwb, const float* bias, float* y, int B, int cin,
                   int H, int W, int P, int D, cudaStream_t s) {
  if (!patch_gemm_supported(img, img_bf16, B, cin, H, W, P) || !w_bf16 || !y || D <= 0) return M2_ERR_ARG;
  PatchDev d = make_dev(img, B, cin, H, W, P, D);
  d.bias = bias; d.out = y;
  CUtensorMap tw;
  int rc = make_tmap_bf16(&tw, w_bf16, D, d.K, ldwb, 128, 64);
  if (rc) return rc;
  dim3 grid(ceil_div(d.M, 128), ceil_div(D, 128));
  LaunchScope scope("patch_embed_fwd", s);
  static bool conf_f = false, conf_b = false;
  if (img_bf16) {
    if (!conf_b) { if (configure(patch_embed_fwd_kernel<__nv_bfloat16>)) return M2_ERR_LAUNCH; conf_b = true; }
    patch_embed_fwd_kernel<__nv_bfloat16><<<grid, kThreads, kSmem, s>>>(tw, d);
  } else {
    if (!conf_f) { if (configure(patch_embed_fwd_kernel<float>)) return M2_ERR_LAUNCH; conf_f = true; }
    patch_embed_fwd_kernel<float><<<grid, kThreads, kSmem, s>>>(tw, d);
  }
  M2_LAUNCH_CHECK();
  return M2_OK;
}

// dy_b: bf16 [M][ldd] copy of dY (ldd % 8 == 0); dw [D][K] fp32 is accumulated into.
int patch_gemm_wgrad(const void* img, int img_bf16, const void* dy_b, int ldd, float* dw, int B, int cin, int H, int W, int P,
                     int D, cudaStream_t s) {
  if (!patch_gemm_supported(img, img_bf16, B, cin, H, W, P) || !dy_b || !dw || D <= 0) return M2_ERR_ARG;
  PatchDev d = make_dev(img, B, cin, H, W, P, D);
  d.out = dw;
  CUtensorMap td;
  int rc = make_tmap_bf16(&td, dy_b, d.M, D, ldd, 64, 64);
  if (rc) return rc;
  const int tiles = ceil_div(D, 128) * ceil_div(d.K, 128);
  const int kt_all = ceil_div(d.M, 64);
  int splits = 148 / tiles;               // one CTA per SM (128 KB of operand ring)
  if (splits < 1) splits = 1;
  if (splits > kt_all) splits = kt_all;
  d.kt_per_split = ceil_div(kt_all, splits);
  d.splits = ceil_div(kt_all, d.kt_per_split);
  dim3 grid(ceil_div(D, 128), ceil_div(d.K, 128), d.splits);
  if (grid.y > 65535 || grid.z > 65535) return M2_ERR_ARG;
  LaunchScope scope("patch_embed_wgrad", s);
  static bool conf_f = false, conf_b = false;
  if (img_bf16) {
    if (!conf_b) { if (configure(patch_embed_wgrad_kernel<__nv_bfloat16>)) return M2_ERR_LAUNCH; conf_b = true; }
    patch_embed_wgrad_kernel<__nv_bfloat16><<<grid, kThreads, kSmem, s>>>(td, d);
  } else {
    if (!conf_f) { if (configure(patch_embed_wgrad_kernel<float>)) return M2_ERR_LAUNCH; conf_f = true; }
    patch_embed_wgrad_kernel<float><<<grid, kThreads, kSmem, s>>>(td, d);
  }
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace m2
