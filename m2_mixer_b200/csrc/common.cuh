// Shared device helpers for the m2b200 kernels (sm_100a only).
// Thin wrappers over the PTX this repo is built on: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld) and the UMMA shared-memory / instruction descriptors.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define M2_OK 0
#define M2_ERR_ARG 1          // bad shape / null pointer / unsupported combination
#define M2_ERR_ALIGN 2        // pointer or leading-dimension alignment
#define M2_ERR_WORKSPACE 3    // workspace too small
#define M2_ERR_LAUNCH 4       // cudaGetLastError() after launch
#define M2_ERR_DRIVER 5       // could not resolve cuTensorMapEncodeTiled / encode failed

#define M2_LAUNCH_CHECK()                              \
  do {                                                 \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return M2_ERR_LAUNCH;      \
  } while (0)

namespace m2 {

// RAII launch accounting (profile.cu): counts the launch, and brackets it with CUDA events when profiling is on.
struct LaunchScope {
  LaunchScope(const char* name, cudaStream_t s, int nkernels = 1);
  ~LaunchScope();
  const char* name_; cudaStream_t stream_; void* start_;
};

constexpr float kLnEps = 1e-5f;   // nn.LayerNorm default (reference modules/mixer.py:31)

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ------------------------------------------------------------------------------------------ math
// Exact (erf) GELU and its derivative, fp32.  Reference: nn.GELU() default, modules/mixer.py:15.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Scalar tanh-form GELU of the bf16 mode (same cubic inner polynomial as the packed version below, so that every
// forward / backward pair in the library differentiates the same function).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = x * fmaf(x * x, 0.03470089342f, 0.80015707848f);
  const float w = fmaf(tanh_approx(u), 0.5f, 0.5f);
  return x * w;
}
// returns GELU(x), writes d/dx
__device__ __forceinline__ float gelu_fast_grad(float x, float& dgelu) {
  const float x2 = x * x;
  const float u = x * fmaf(x2, 0.03470089342f, 0.80015707848f);
  const float du2 = fmaf(x2, 6.f * 0.03470089342f, 2.f * 0.80015707848f);
  const float w = fmaf(tanh_approx(u), 0.5f, 0.5f);
  const float g = x * w;
  dgelu = fmaf(g * (1.f - w), du2, w);
  return g;
}

// ---- packed fp32x2 GELU used by every bf16-mode tensor-core epilogue.  The epilogues are CUDA-core ISSUE bound
// (0.73 warp-instructions per hidden element in the weight-gradient kernel, profiles/r01_ncu_final.md), so the form
// is chosen for instruction count: tanh form with a CUBIC inner polynomial u = x (a + b x^2) fitted to the erf form
// (max abs error 2.7e-4 on the value, 8.7e-4 on the derivative: below the bf16 rounding of the result for |G| > 0.1),
// monotonic, so no clamping and no saturation fix-ups are needed (tanh.approx saturates to +-1):
//     w = (1 + tanh u) / 2        GELU = x w        GELU' = w + 2 (x w) (1 - w) u'
//   forward 7 instructions per pair (FMUL2, FFMA2, FMUL2, 2 MUFU, FFMA2, FMUL2), backward 11.
constexpr float kG3a = 0.80015707848f, kG3b = 0.03470089342f;
__device__ __forceinline__ float tanh_ap(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// `hs` = s / 2 folds the dropout scale s = 1 / (1 - p) of the following nn.Dropout into the result for free
// (the epilogue then only SELECTS kept values, drop_zero2): returns s * GELU(x).
__device__ __forceinline__ float2 gelu2(float2 x, float hs = 0.5f) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 p = __ffma2_rn(x2, make_float2(kG3b, kG3b), make_float2(kG3a, kG3a));
  const float2 u = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_ap(u.x), tanh_ap(u.y));
  const float2 w = __ffma2_rn(t, make_float2(hs, hs), make_float2(hs, hs));
  return __fmul2_rn(x, w);
}
// returns s * GELU(x); dg = s * GELU'(x)   (s = 2 hs, ninv_s = -1 / s; defaults: s = 1)
__device__ __forceinline__ float2 gelu2_grad(float2 x, float2& dg, float hs = 0.5f, float ninv_s = -1.f) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 p = __ffma2_rn(x2, make_float2(kG3b, kG3b), make_float2(kG3a, kG3a));
  const float2 du2 = __ffma2_rn(x2, make_float2(6.f * kG3b, 6.f * kG3b), make_float2(2.f * kG3a, 2.f * kG3a));   // 2 u'
  const float2 u = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_ap(u.x), tanh_ap(u.y));
  const float2 w = __ffma2_rn(t, make_float2(hs, hs), make_float2(hs, hs));           // s w0
  const float2 g = __fmul2_rn(x, w);                                                  // s GELU
  const float2 omw = __ffma2_rn(w, make_float2(ninv_s, ninv_s), make_float2(1.f, 1.f));   // 1 - w0
  const float2 tmp = __fmul2_rn(g, omw);
  dg = __ffma2_rn(tmp, du2, w);                                                       // s (w0 + 2 GELU (1 - w0) u')
  return g;
}

// ------------------------------------------------------------------------------------------ dropout
// Counter-based mask shared by every kernel (forward and the recomputing backward evaluate the same function):
// one 32-bit hash per QUAD of consecutive elements, 7 random bits per element (one byte each, top bit unused), keep iff
// bits >= thresh with thresh = round(p * 128): the drop probability is quantised to 1/128 (exact for p = 0.5, 0.25, ...;
// 0.1 -> 0.1016, 0.3 -> 0.2969) and scale = 128 / (128 - thresh) is the inverse of the REALISED keep probability, so
// the estimator stays unbiased.  The four compares are ONE add: (h & 0x7f7f7f7f) + (0x80 - thresh per byte) leaves
// each element's keep flag in bit 7 of its byte, and a sign-replicating PRMT turns two flags into an AND mask for a
// packed bf16x2 pair.  With the previous layout (16 bits per element, one hash per pair, ISETP + FSEL per element) the
// mask arithmetic was HALF of the dgrad epilogue's instructions (56 of 114 per 8 hidden elements in the SASS of
// chain_bwd_ts_kernel); this form costs 11 per quad.  Not bit-compatible with torch's Philox stream by design
// (SURVEY H7): tested statistically and through m2b200_dropout_mask(), which exports exactly this function.
struct Drop {
  uint32_t key;      // per call-site key derived from (seed, site) on the host
  uint32_t thresh;   // round(p * 128), 0 = dropout off
  float scale;       // 128 / (128 - thresh)
  uint32_t kadd;     // 0x80808080 - thresh * 0x01010101
  const uint32_t* epoch;   // optional device counter folded into the key at run time (nullptr: off), see drop_key()
};
// The (seed, site) key is a launch PARAMETER, so a captured CUDA graph would replay the same masks for ever.  When the host
// has registered a device-resident epoch counter (m2b200_set_dropout_epoch_ptr; the graphed training step advances it once
// per replay) every kernel folds it into its key: same kernels, same parameters, fresh masks per replay.
extern const uint32_t* g_drop_epoch_ptr;   // profile.cu
__device__ __forceinline__ uint32_t drop_key(const Drop& d) {
  return d.epoch ? d.key + __ldg(d.epoch) * 0x85EBCA6BU : d.key;
}
// Two-round multiply / xorshift finisher (the first two rounds of lowbias32); the key is already a full-avalanche
// splitmix64 of (seed, site).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15;
  return x;
}
constexpr uint32_t kDropGolden = 0x9E3779B9U;
// Only the LOW 32 bits of the element index enter the hash (the mask of a tensor with more than 2^32 elements repeats with
// that period): all index arithmetic at the call sites narrows to 32-bit integer instructions.
// Keep flags of elements 4q .. 4q+3: bit 7 of byte i <=> element 4q + i is kept.
__device__ __forceinline__ uint32_t drop_flags_from_hash_input(const Drop& d, uint32_t hin) {
  return (mix32(hin) & 0x7f7f7f7fU) + d.kadd;
}
__device__ __forceinline__ uint32_t drop_quad_flags(const Drop& d, uint32_t quad_idx) {
  return drop_flags_from_hash_input(d, quad_idx * kDropGolden + drop_key(d));
}
// AND masks from the flags (prmt with the sign-replicate bit set in every selector nibble).
template <int kPair>   // elements 2 kPair, 2 kPair + 1 of the quad as the two halves of a packed bf16x2
__device__ __forceinline__ uint32_t drop_mask_bf16x2(uint32_t flags) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(flags), "n"(kPair ? 0xBBAA : 0x9988));
  return m;
}
template <int kElem>   // element kElem of the quad as a full 32-bit mask
__device__ __forceinline__ uint32_t drop_mask_b32(uint32_t flags) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(flags), "n"(0x8888 + 0x1111 * kElem));
  return m;
}
__device__ __forceinline__ float drop_and(float v, uint32_t m) { return __uint_as_float(__float_as_uint(v) & m); }

__device__ __forceinline__ bool drop_keep(const Drop& d, unsigned long long idx) {
  const uint32_t lo = static_cast<uint32_t>(idx);
  return (drop_quad_flags(d, lo >> 2) >> (8 * (lo & 3u) + 7)) & 1u;
}
__device__ __forceinline__ float drop_apply(const Drop& d, float v, unsigned long long idx) {
  return drop_keep(d, idx) ? v * d.scale : 0.f;
}
// two consecutive elements starting at an EVEN index: one hash
__device__ __forceinline__ uint32_t drop_pair_flags(const Drop& d, unsigned long long even_idx) {   // bits 7 and 15
  const uint32_t lo = static_cast<uint32_t>(even_idx);
  return drop_quad_flags(d, lo >> 2) >> ((lo & 2u) << 3);
}
__device__ __forceinline__ void drop_apply2(const Drop& d, float& v0, float& v1, unsigned long long even_idx) {
  const uint32_t f = drop_pair_flags(d, even_idx);
  v0 = (f & 0x80u) ? v0 * d.scale : 0.f;
  v1 = (f & 0x8000u) ? v1 * d.scale : 0.f;
}
// same mask, values already carry the scale (gelu2 / gelu2_grad with hs = scale / 2): select only
__device__ __forceinline__ void drop_zero2(const Drop& d, float& v0, float& v1, unsigned long long even_idx) {
  const uint32_t f = drop_pair_flags(d, even_idx);
  v0 = (f & 0x80u) ? v0 : 0.f;
  v1 = (f & 0x8000u) ? v1 : 0.f;
}
// the same hash masks TWO value pairs (G and dH of the backward kernels)
__device__ __forceinline__ void drop_zero2x2(const Drop& d, float& a0, float& a1, float& b0, float& b1, unsigned long long even_idx) {
  const uint32_t f = drop_pair_flags(d, even_idx);
  const bool k0 = f & 0x80u, k1 = f & 0x8000u;
  a0 = k0 ? a0 : 0.f; b0 = k0 ? b0 : 0.f;
  a1 = k1 ? a1 : 0.f; b1 = k1 ? b1 : 0.f;
}
// The same pair functions for callers that keep the quad's hash input (quad_idx * kDropGolden + drop_key(d)) incrementally
// in 32-bit registers: `sh` = 0 for the quad's first pair, 16 for its second.
__device__ __forceinline__ void drop_zero2_hin(const Drop& d, float& v0, float& v1, uint32_t hin, uint32_t sh) {
  const uint32_t f = drop_flags_from_hash_input(d, hin) >> sh;
  v0 = (f & 0x80u) ? v0 : 0.f;
  v1 = (f & 0x8000u) ? v1 : 0.f;
}
__device__ __forceinline__ void drop_zero2x2_hin(const Drop& d, float& a0, float& a1, float& b0, float& b1, uint32_t hin, uint32_t sh) {
  const uint32_t f = drop_flags_from_hash_input(d, hin) >> sh;
  const bool k0 = f & 0x80u, k1 = f & 0x8000u;
  a0 = k0 ? a0 : 0.f; b0 = k0 ? b0 : 0.f;
  a1 = k1 ? a1 : 0.f; b1 = k1 ? b1 : 0.f;
}
__device__ __forceinline__ void drop_apply2_hin(const Drop& d, float& v0, float& v1, uint32_t hin, uint32_t sh) {
  const uint32_t f = drop_flags_from_hash_input(d, hin) >> sh;
  v0 = (f & 0x80u) ? v0 * d.scale : 0.f;
  v1 = (f & 0x8000u) ? v1 * d.scale : 0.f;
}
// four consecutive elements starting at a multiple of 4 (fp32 values, scale applied)
__device__ __forceinline__ void drop_apply4(const Drop& d, float4& v, unsigned long long idx4) {
  const uint32_t f = drop_quad_flags(d, static_cast<uint32_t>(idx4) >> 2);
  v.x = drop_and(v.x * d.scale, drop_mask_b32<0>(f)); v.y = drop_and(v.y * d.scale, drop_mask_b32<1>(f));
  v.z = drop_and(v.z * d.scale, drop_mask_b32<2>(f)); v.w = drop_and(v.w * d.scale, drop_mask_b32<3>(f));
}
inline Drop make_drop(float p, unsigned long long seed, uint32_t site) {
  Drop d;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (site + 1);   // splitmix64 of (seed, site)
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  d.key = static_cast<uint32_t>(z ^ (z >> 32));
  long t = p > 0.f ? static_cast<long>(p * 128.0f + 0.5f) : 0;
  if (p > 0.f && t < 1) t = 1;
  if (t > 127) t = 127;
  d.thresh = static_cast<uint32_t>(t);
  d.scale = 128.0f / (128.0f - static_cast<float>(t));
  d.kadd = 0x80808080U - d.thresh * 0x01010101U;
  d.epoch = g_drop_epoch_ptr;
  return d;
}
enum DropSite { kSiteTokenHidden = 0, kSiteTokenOut = 1, kSiteChannelHidden = 2, kSiteChannelOut = 3, kSiteLinear = 4 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------ smem / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error on the host) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// Early probe of a barrier whose result is consumed later (`if (!ready) mbar_wait(...)`): the ~130-clk SYNCS round trip of a
// wait on an already completed phase then overlaps the math between the probe and its use.
__device__ __forceinline__ bool mbar_probe(uint64_t* bar, uint32_t parity) { return mbar_try_wait(bar, parity); }

// Two barriers at once.  A wait costs ~120-150 clk even when the phase completed long ago (tools/wait_probe.cu: the
// SYNCS.TRYWAIT round trip); issuing both probes back to back overlaps the two round trips.
__device__ __forceinline__ bool mbar_try_wait2(uint64_t* bar_a, uint32_t parity_a, uint64_t* bar_b, uint32_t parity_b) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
      "and.pred p, p, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar_a)), "r"(parity_a), "r"(smem_u32(bar_b)), "r"(parity_b)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait2(uint64_t* bar_a, uint32_t parity_a, uint64_t* bar_b, uint32_t parity_b) {
  uint32_t spins = 0;
  while (!mbar_try_wait2(bar_a, parity_a, bar_b, parity_b)) {
    if (++spins > (1u << 26)) __trap();
  }
}

__device__ __forceinline__ void mbar_wait3(uint64_t* bar_a, uint32_t parity_a, uint64_t* bar_b, uint32_t parity_b, uint64_t* bar_c,
                                           uint32_t parity_c) {
  uint32_t spins = 0, ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p, q, r;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 r, [%5], %6;\n\t"
        "and.pred p, p, q;\n\t"
        "and.pred p, p, r;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar_a)), "r"(parity_a), "r"(smem_u32(bar_b)), "r"(parity_b), "r"(smem_u32(bar_c)), "r"(parity_c)
        : "memory");
    if (++spins > (1u << 26)) __trap();
  } while (!ok);
}

// One lane of a CONVERGED warp (the lowest).  Roles that issue tcgen05.mma / TMA from a single thread walk their loop
// with the whole warp and predicate only the issue on this: inside an `if (lane == 0)` region ptxas cannot prove
// warp-uniformity and wraps every UTCHMMA in an elect / R2UR.BROADCAST / BRA.U.ANY serialisation loop (~250 clk per
// MMA measured, profiles/r01_ncu_final.md), with it the operands stay in uniform registers.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2D tiled load: box lands in smem (swizzled per the tensor map) and completes `bytes` on the mbarrier.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(x), "r"(y), "r"(z)
               : "memory");
}
// L2 eviction-priority policy for streams that pass through L2 once (the spilled dH: 100-200 MB per block between the
// dgrad chain and the weight-gradient kernel): evict_first keeps them from displacing the block's inputs and weights.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, const void* smem_src, int x, int y, int z, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(x), "r"(y), "r"(z), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int kPending>   // all but the newest kPending bulk groups of this thread have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }

// ------------------------------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32.  One thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this thread's lane (= accumulator row), 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA shared-memory matrix descriptor, 128-byte swizzle (layout_type = 2), descriptor version 1 (sm_100).
// K-major operand tile  [rows = M or N][64 bf16 = 128 B]: 8-row groups are 1024 B apart (SBO); LBO unused.
// MN-major operand tile [rows = K][64 bf16 of M or N = 128 B]: 8-row K groups 1024 B apart (SBO); the next
// 64-element MN panel is `lbo_bytes` away (LBO).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);            // [0,14)  start address >> 4
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;       // [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;       // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                                // [46,48) descriptor version = 1
  d |= static_cast<uint64_t>(2) << 61;                                // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16, A/B = bf16, D = fp32, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4)                                    // c_format  = F32
         | (1u << 7)                                  // a_format  = BF16
         | (1u << 10)                                 // b_format  = BF16
         | (static_cast<uint32_t>(a_mn_major) << 15)  // a_major   (0 = K, 1 = MN)
         | (static_cast<uint32_t>(b_mn_major) << 16)  // b_major
         | (static_cast<uint32_t>(n >> 3) << 17)      // n_dim
         | (static_cast<uint32_t>(m >> 4) << 24);     // m_dim
}

// Byte offset of element (row r, 16-byte chunk c) inside a [rows][128 B] tile with the 128-byte swizzle
// (what TMA SWIZZLE_128B writes and the SW128 UMMA descriptor reads).  Tile base must be 1024-B aligned.
__device__ __forceinline__ uint32_t sw128_offset(int r, int chunk16) { return r * 128 + ((chunk16 ^ (r & 7)) << 4); }

}  // namespace m2
