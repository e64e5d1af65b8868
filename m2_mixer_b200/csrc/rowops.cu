// Bandwidth-bound row kernels around the GEMM chains: LayerNorm fwd/bwd, casts, column sums, GELU fwd/bwd for the
// unfused path, patch gather (Conv2d with kernel == stride as a GEMM operand), token concat/split.
// All fp32 math; one warp per row where a row reduction is needed, 128-bit accesses where alignment allows.
#include "common.cuh"
#include "kernels.h"

namespace m2 {
namespace {

constexpr int kNumSms = 148;

// ------------------------------------------------------------------------------------------ cast
__global__ void cast_pad_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                     long long ldd, int rows, int cols) {
  const long long total = static_cast<long long>(rows) * ldd;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / ldd;
    const int c = static_cast<int>(i - r * ldd);
    dst[i] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.f);
  }
}
// 8 elements per thread (two 16-byte loads, one 16-byte store); row = blockIdx.y * rows_per_block + ..: no divisions.
// Requires lds % 4 == 0, ldd % 8 == 0 and 16-byte aligned bases; columns >= cols are written as zero.
__global__ void __launch_bounds__(256) cast_pad_bf16_vec_kernel(const float* __restrict__ src, long long lds,
                                                                __nv_bfloat16* __restrict__ dst, long long ldd, int rows,
                                                                int cols) {
  const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 8;
  if (c >= ldd) return;
  for (int r = blockIdx.y * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.y * 8) {
    const float* s = src + r * lds + c;
    float v[8];
    if (c + 8 <= cols) {
      const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = c + e < cols ? s[e] : 0.f;
    }
    *reinterpret_cast<uint4*>(dst + r * ldd + c) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// All bf16 operand copies of a model in ONE launch: table[e] = {src, dst, rows, cols, lds, ldd} (device int64 sextuples),
// blockIdx.y = entry.  Replaces one small launch per weight matrix (22 per M2-Mixer-B step, 0.17 ms).
__global__ void __launch_bounds__(256) cast_pad_bf16_multi_kernel(const long long* __restrict__ table) {
  const long long* e = table + 6 * blockIdx.y;
  const float* __restrict__ src = reinterpret_cast<const float*>(e[0]);
  __nv_bfloat16* __restrict__ dst = reinterpret_cast<__nv_bfloat16*>(e[1]);
  const int rows = static_cast<int>(e[2]), cols = static_cast<int>(e[3]);
  const long long lds = e[4], ldd = e[5];
  const bool vec = (lds % 4 == 0) && (ldd % 8 == 0) && (((e[0] | e[1]) & 15) == 0);
  if (vec) {
    const int gpr = static_cast<int>(ldd / 8);
    const long long total = static_cast<long long>(rows) * gpr;
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int r = static_cast<int>(g / gpr), c = static_cast<int>(g - static_cast<long long>(r) * gpr) * 8;
      const float* sp = src + r * lds + c;
      float v[8];
      if (c + 8 <= cols) {
        const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = c + k < cols ? sp[k] : 0.f;
      }
      *reinterpret_cast<uint4*>(dst + r * ldd + c) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  } else {
    const long long total = static_cast<long long>(rows) * ldd;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long r = i / ldd;
      const int c = static_cast<int>(i - r * ldd);
      dst[i] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm forward
// rows = B*N tokens; output row (b, n) goes to out + b*out_bstride + n*D  (lets an encoder's final LN write
// straight into its slice of the fused-token buffer: zero-copy ConcatFusion, reference modules/fusion.py:117).
template <bool kBf16Out>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ b, void* __restrict__ out, int rows, int D,
                                                     int N, long long out_bstride, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < rows; row += nwarps) {
    const float* xr = x + static_cast<long long>(row) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = xr[c] - mean; ss += d * d; }
    const float rstd = rsqrtf(warp_sum(ss) / D + kLnEps);
    const long long o = static_cast<long long>(row / N) * out_bstride + static_cast<long long>(row % N) * D;
    for (int c = lane; c < D; c += 32) {
      const float v = (xr[c] - mean) * rstd * w[c] + b[c];
      if (kBf16Out) static_cast<__nv_bfloat16*>(out)[o + c] = __float2bfloat16(v);
      else static_cast<float*>(out)[o + c] = v;
    }
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}


// float4 form (D % 4 == 0, 16-byte aligned rows, D <= 128 * kV): the row is read ONCE into registers (the scalar kernel
// above walks it three times through L1) with 16-byte loads; one warp per row.
template <bool kBf16Out, int kV>
__global__ void __launch_bounds__(256) ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, void* __restrict__ out, int rows, int D,
                                                         int N, long long out_bstride, float* __restrict__ mean_out,
                                                         float* __restrict__ rstd_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.f / D;
  for (int row = warp; row < rows; row += nwarps) {
    const float* xr = x + static_cast<long long>(row) * D;
    float4 v[kV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = (lane + 32 * i) * 4;
      v[i] = c < D ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i)
      if ((lane + 32 * i) * 4 < D) {
        const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
        ss += a * a + bb * bb + cc * cc + dd * dd;
      }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + kLnEps);
    const long long o = static_cast<long long>(row / N) * out_bstride + static_cast<long long>(row % N) * D;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < D) {
        const float4 ww = *reinterpret_cast<const float4*>(w + c), bv = *reinterpret_cast<const float4*>(b + c);
        const float4 r = make_float4((v[i].x - mean) * rstd * ww.x + bv.x, (v[i].y - mean) * rstd * ww.y + bv.y,
                                     (v[i].z - mean) * rstd * ww.z + bv.z, (v[i].w - mean) * rstd * ww.w + bv.w);
        if (kBf16Out)
          *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + o + c) = make_uint2(pack_bf16(r.x, r.y), pack_bf16(r.z, r.w));
        else
          *reinterpret_cast<float4*>(static_cast<float*>(out) + o + c) = r;
      }
    }
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm backward
// dy = grad wrt LN output (row (b,n) at dy + b*dy_bstride + n*D), x = LN input.
// dx[row] = (dres ? dres[row] : 0) + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * w
// dw += sum_rows dy*xhat, db += sum_rows dy   (per-CTA partials in smem, then one atomicAdd per column per CTA)
template <int kMaxPer>   // columns per lane: D <= 32 * kMaxPer
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, long long dy_bstride, int N,
                                                     const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ dres, float* __restrict__ dx,
                                                     float* __restrict__ dw, float* __restrict__ db, int rows, int D) {
  extern __shared__ float sm[];   // [2][D] per-CTA partial dw/db
  float* sdw = sm;
  float* sdb = sm + D;
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  // per-lane register partials for the columns this lane owns (c = lane + 32*i), flushed to smem at the end
  float pdw[kMaxPer], pdb[kMaxPer];
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) { pdw[i] = 0.f; pdb[i] = 0.f; }
  for (int row = warp; row < rows; row += nwarps) {
    const float* xr = x + static_cast<long long>(row) * D;
    const float* gr = dy + static_cast<long long>(row / N) * dy_bstride + static_cast<long long>(row % N) * D;
    float xv[kMaxPer], gv[kMaxPer];   // the row stays in registers: x and dy are read exactly once
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int c = lane + 32 * i;
      xv[i] = c < D ? xr[c] : 0.f;
      gv[i] = c < D ? gr[c] : 0.f;
      s += xv[i];
    }
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const float d = (lane + 32 * i < D) ? xv[i] - mean : 0.f;
      ss += d * d;
    }
    const float rstd = rsqrtf(warp_sum(ss) / D + kLnEps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        xv[i] = (xv[i] - mean) * rstd;            // xhat
        const float g = gv[i] * w[c];
        sg += g; sgx += g * xv[i];
      }
    }
    sg = warp_sum(sg) / D; sgx = warp_sum(sgx) / D;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        float v = rstd * (gv[i] * w[c] - sg - xv[i] * sgx);
        if (dres) v += dres[static_cast<long long>(row) * D + c];
        dx[static_cast<long long>(row) * D + c] = v;
        pdw[i] += gv[i] * xv[i];
        pdb[i] += gv[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) {
    const int c = lane + 32 * i;
    if (c < D) { atomicAdd(&sdw[c], pdw[i]); atomicAdd(&sdb[c], pdb[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) { atomicAdd(&dw[c], sdw[c]); atomicAdd(&db[c], sdb[c]); }
}


// float4 form of the above (D % 4 == 0, 16-byte aligned rows, D <= 128 * kV).  The scalar kernel keeps a D = 768 row in
// 24 + 24 scalar registers per lane plus 48 accumulators and issues 4-byte accesses: 160-236 registers, 8 warps per SM,
// 6x off its HBM floor on the Scaled config (profiles/r01_large_configs.json).  Here: 16-byte accesses (a quarter of the
// memory instructions), at most two CTAs' worth of registers, x and dy still read exactly once.
template <int kV>
__global__ void __launch_bounds__(256, kV <= 2 ? 4 : (kV <= 6 ? 2 : 1))
ln_bwd_vec_kernel(const float* __restrict__ dy, long long dy_bstride, int N, const float* __restrict__ x,
                  const float* __restrict__ w, const float* __restrict__ dres, float* __restrict__ dx,
                  float* __restrict__ dw, float* __restrict__ db, int rows, int D) {
  extern __shared__ float sm[];   // [2][D] per-CTA partial dw/db
  float* sdw = sm;
  float* sdb = sm + D;
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.f / D;
  float4 pdw[kV], pdb[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) pdw[i] = pdb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto wld = [&](int i) {   // gamma stays in L1 (D <= 1024 floats): re-read instead of holding kV float4 registers per lane
    const int c = (lane + 32 * i) * 4;
    return c < D ? __ldg(reinterpret_cast<const float4*>(w + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  for (int row = warp; row < rows; row += nwarps) {
    const float* xr = x + static_cast<long long>(row) * D;
    const float* gr = dy + static_cast<long long>(row / N) * dy_bstride + static_cast<long long>(row % N) * D;
    float4 xv[kV], gv[kV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = (lane + 32 * i) * 4;
      const bool ok = c < D;
      xv[i] = ok ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      gv[i] = ok ? *reinterpret_cast<const float4*>(gr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
    }
    const float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i)
      if ((lane + 32 * i) * 4 < D) {
        xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
        ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
      }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + kLnEps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {   // xv <- xhat (pad lanes hold zeros: they add nothing)
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;
      const float4 wv = wld(i);
      const float g0 = gv[i].x * wv.x, g1 = gv[i].y * wv.y, g2 = gv[i].z * wv.z, g3 = gv[i].w * wv.w;
      sg += g0 + g1 + g2 + g3;
      sgx += g0 * xv[i].x + g1 * xv[i].y + g2 * xv[i].z + g3 * xv[i].w;
    }
    sg = warp_sum(sg) * inv_d; sgx = warp_sum(sgx) * inv_d;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < D) {
        const float4 wv = wld(i);
        float4 v = make_float4(rstd * (gv[i].x * wv.x - sg - xv[i].x * sgx), rstd * (gv[i].y * wv.y - sg - xv[i].y * sgx),
                               rstd * (gv[i].z * wv.z - sg - xv[i].z * sgx), rstd * (gv[i].w * wv.w - sg - xv[i].w * sgx));
        if (dres) {
          const float4 r = *reinterpret_cast<const float4*>(dres + static_cast<long long>(row) * D + c);
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        *reinterpret_cast<float4*>(dx + static_cast<long long>(row) * D + c) = v;
        pdw[i].x += gv[i].x * xv[i].x; pdw[i].y += gv[i].y * xv[i].y; pdw[i].z += gv[i].z * xv[i].z; pdw[i].w += gv[i].w * xv[i].w;
        pdb[i].x += gv[i].x; pdb[i].y += gv[i].y; pdb[i].z += gv[i].z; pdb[i].w += gv[i].w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) {
      atomicAdd(&sdw[c], pdw[i].x); atomicAdd(&sdw[c + 1], pdw[i].y); atomicAdd(&sdw[c + 2], pdw[i].z); atomicAdd(&sdw[c + 3], pdw[i].w);
      atomicAdd(&sdb[c], pdb[i].x); atomicAdd(&sdb[c + 1], pdb[i].y); atomicAdd(&sdb[c + 2], pdb[i].z); atomicAdd(&sdb[c + 3], pdb[i].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) { atomicAdd(&dw[c], sdw[c]); atomicAdd(&db[c], sdb[c]); }
}

// ------------------------------------------------------------------------------------------ column sums
// out[c] += sum_r src[r][c].  A thread owns kV consecutive columns (one 16-byte load per row), 8 row-lanes per CTA,
// grid.y splits the rows; per-CTA partials are combined in shared memory and added with one atomic per column.
template <typename T, int kV>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ src, long long ld, int rows, int cols,
                                                         float* __restrict__ out) {
  __shared__ float part[8][32 * kV + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + tx) * kV;
  float acc[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) acc[i] = 0.f;
  if (c0 < ld) {   // whole vectors lie inside the padded row (ld % kV == 0)
    const int step = gridDim.y * 8;
    for (int r = blockIdx.y * 8 + ty; r < rows; r += 4 * step) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)   // four independent 16-byte loads in flight per thread
        v[k] = (r + k * step < rows) ? *reinterpret_cast<const uint4*>(src + static_cast<long long>(r + k * step) * ld + c0)
                                     : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (sizeof(T) == 2) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[k]);
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
        } else {
          acc[0] += __uint_as_float(v[k].x); acc[1] += __uint_as_float(v[k].y);
          acc[2] += __uint_as_float(v[k].z); acc[3] += __uint_as_float(v[k].w);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) part[ty][tx * kV + i] = acc[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * kV; i += 256) {
    const int c = blockIdx.x * 32 * kV + i;
    if (c < cols) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += part[j][i];
      atomicAdd(&out[c], t);
    }
  }
}

// scalar fallback (unaligned base / leading dimension)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ src, long long ld, int rows, int cols,
                                                     float* __restrict__ out) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cols)
    for (int r = blockIdx.y * 8 + ty; r < rows; r += gridDim.y * 8) s += static_cast<float>(src[r * ld + c]);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    atomicAdd(&out[c], t);
  }
}

// ------------------------------------------------------------------------------------------ GELU fwd+bwd (unfused path)
// out[r % period] += sum_c src[r][c]: warp per row, per-CTA shared-memory partials (period <= 1024), one atomic per slot.
__global__ void __launch_bounds__(256) rowsum_mod_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long ld, int rows,
                                                              int cols, int period, float* __restrict__ out) {
  extern __shared__ float part[];   // [period]
  for (int i = threadIdx.x; i < period; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const __nv_bfloat16* p = src + static_cast<long long>(r) * ld;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += __bfloat162float(p[c]);
    s = warp_sum(s);
    if (lane == 0) atomicAdd(&part[r % period], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < period; i += blockDim.x)
    if (part[i] != 0.f) atomicAdd(&out[i], part[i]);
}

// h = pre-activation (bias already added). g_out = gelu(h) ; dh = dg * gelu'(h)  (dh may alias dg)
template <typename TO, typename TI = float>
__global__ void gelu_fwd_bwd_kernel(const TI* __restrict__ h, const TI* dg, long long n, int cols,
                                    long long ld_in, TO* __restrict__ g_out, TO* dh_out, long long ld_out, const Drop drop,
                                    long long drop_ld) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float x = static_cast<float>(h[r * ld_in + c]);
    const float dgv = static_cast<float>(dg[r * ld_in + c]);
    float g, d;
    if (sizeof(TO) == 2) {      // bf16 mode: the tanh form every bf16 kernel of the library differentiates (common.cuh)
      float dgel;
      g = gelu_fast_grad(x, dgel);
      d = dgv * dgel;
    } else {                    // fp32 parity mode: exact erf GELU
      g = gelu_erf(x); d = dgv * gelu_erf_grad(x);
    }
    if (drop.thresh) {
      const bool keep = drop_keep(drop, static_cast<unsigned long long>(r) * drop_ld + c);
      g = keep ? g * drop.scale : 0.f;
      d = keep ? d * drop.scale : 0.f;
    }
    g_out[r * ld_out + c] = static_cast<TO>(g);
    dh_out[r * ld_out + c] = static_cast<TO>(d);
  }
}

// Four consecutive columns per thread (cols, leading dimensions and the mask row stride multiples of 4, 16-byte aligned
// bases): 16-byte loads, one mask hash per thread, 8- / 16-byte stores.  The scalar kernel above ran at a sixth of the HBM
// rate on the Scaled config's [tokens x 3072] tensors (profiles/r01_large_configs.json).
template <typename TO>
__global__ void __launch_bounds__(256) gelu_fwd_bwd_vec4_kernel(const float* __restrict__ h, const float* dg, long long n4,
                                                                int cols4, long long ld_in, TO* __restrict__ g_out, TO* dh_out,
                                                                long long ld_out, const Drop drop, long long drop_ld) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols4;
    const int c = static_cast<int>(i - r * cols4) * 4;
    const float4 x = *reinterpret_cast<const float4*>(h + r * ld_in + c);
    const float4 dgv = *reinterpret_cast<const float4*>(dg + r * ld_in + c);
    float4 g, d;
    if (sizeof(TO) == 2) {      // bf16 mode: packed tanh-form GELU and derivative (common.cuh gelu2_grad)
      float2 d0, d1;
      const float2 g0 = gelu2_grad(make_float2(x.x, x.y), d0), g1 = gelu2_grad(make_float2(x.z, x.w), d1);
      g = make_float4(g0.x, g0.y, g1.x, g1.y);
      d = make_float4(dgv.x * d0.x, dgv.y * d0.y, dgv.z * d1.x, dgv.w * d1.y);
    } else {                    // fp32 parity mode: exact erf GELU
      g = make_float4(gelu_erf(x.x), gelu_erf(x.y), gelu_erf(x.z), gelu_erf(x.w));
      d = make_float4(dgv.x * gelu_erf_grad(x.x), dgv.y * gelu_erf_grad(x.y), dgv.z * gelu_erf_grad(x.z),
                      dgv.w * gelu_erf_grad(x.w));
    }
    if (drop.thresh) {
      const unsigned long long idx = static_cast<unsigned long long>(r) * drop_ld + c;
      drop_apply4(drop, g, idx);
      drop_apply4(drop, d, idx);
    }
    if (sizeof(TO) == 2) {
      *reinterpret_cast<uint2*>(g_out + r * ld_out + c) = make_uint2(pack_bf16(g.x, g.y), pack_bf16(g.z, g.w));
      *reinterpret_cast<uint2*>(dh_out + r * ld_out + c) = make_uint2(pack_bf16(d.x, d.y), pack_bf16(d.z, d.w));
    } else {
      *reinterpret_cast<float4*>(g_out + r * ld_out + c) = g;
      *reinterpret_cast<float4*>(dh_out + r * ld_out + c) = d;
    }
  }
}

// ------------------------------------------------------------------------------------------ patch gather
// img [B][cin][H][W] -> cols [(b, gy, gx)][(ci, py, px)], row stride ld (pad columns zeroed).
// Reference: Conv2d(cin, D, p, p) + 'b c h w -> b (h w) c', modules/mixer.py:143-146.
// (patch sizes with P % 8 == 0 never come here in BF16 mode: patch_gemm.cu gathers inside the GEMM)
// One block walks whole patch rows; the k -> pixel offset table of a patch is built once per block in shared memory, so
// the per-element work is one table read, one load and one store (the first version spent ~300 instructions per element
// on 64-bit index divisions: 38 us for the 3.3 M elements of the AV-MNIST image branch).
template <typename TI, typename TO>
__global__ void patch_gather_kernel(const TI* __restrict__ img, TO* __restrict__ cols, int rows, int cin, int H, int W,
                                    int P, int ld) {
  extern __shared__ int koff[];   // [ld], -1 = pad column
  const int K = cin * P * P, gw = W / P, gh = H / P;
  for (int k = threadIdx.x; k < ld; k += blockDim.x) {
    int o = -1;
    if (k < K) {
      const int px = k % P, py = (k / P) % P, ci = k / (P * P);
      o = (ci * H + py) * W + px;
    }
    koff[k] = o;
  }
  __syncthreads();
  // four patch rows per pass: their loads are independent, so four times the bytes are in flight per thread
  for (int row0 = blockIdx.x * 4; row0 < rows; row0 += gridDim.x * 4) {
    const TI* base[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = min(row0 + r, rows - 1);
      const int gx = row % gw, gy = (row / gw) % gh, b = row / (gw * gh);
      base[r] = img + (static_cast<long long>(b) * cin * H + gy * P) * W + gx * P;
    }
    for (int k = threadIdx.x; k < ld; k += blockDim.x) {
      const int o = koff[k];
      float v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r] = o >= 0 ? static_cast<float>(base[r][o]) : 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (row0 + r < rows) cols[static_cast<long long>(row0 + r) * ld + k] = static_cast<TO>(v[r]);
    }
  }
}

// ------------------------------------------------------------------------------------------ token concat / split
// dst[b][n_off + n][d] = src[b][n][d]   (copy == 1)    or   dst[b][n][d] (+)= src[b][n_off + n][d]  (copy == 0)
__global__ void concat_copy_kernel(const float* __restrict__ src, long long src_bstride, float* __restrict__ dst,
                                   long long dst_bstride, int B, long long per_batch, int accumulate) {
  const long long total = static_cast<long long>(B) * per_batch;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / per_batch, e = i - b * per_batch;
    const float v = src[b * src_bstride + e];
    float* d = dst + b * dst_bstride + e;
    *d = accumulate ? *d + v : v;
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    o[i] = a[i] + b[i];
}

// Elementwise fusion of two token tensors (reference modules/fusion.py:190-204 MaxFusion, :258-272 MeanFusion):
//   mode 1: o = max(a, b)      mode 2: o = (a + b) / 2
// backward of max (torch.maximum semantics: ties split the gradient evenly): da = g * [a > b] + g/2 * [a == b], db = g - da
__global__ void fuse2_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n, int mode) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    o[i] = mode == 1 ? fmaxf(a[i], b[i]) : 0.5f * (a[i] + b[i]);
}
__global__ void fuse2_max_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g,
                                     float* __restrict__ da, float* __restrict__ db, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = a[i], y = b[i], gi = g[i];
    const float ga = x > y ? gi : (x == y ? 0.5f * gi : 0.f);
    da[i] = ga;
    db[i] = gi - ga;
  }
}

// Gate of BiModalGatedUnit.forward (reference modules/fusion.py:16-23), exact tanh / sigmoid in fp32:
//   out = z tanh(h1) + (1 - z) tanh(h2),  z = sigmoid(zh)
//   dh1 = g z (1 - t1^2)   dh2 = g (1 - z) (1 - t2^2)   dzh = g (t1 - t2) z (1 - z)
__global__ void gate_fwd_kernel(const float* __restrict__ h1, const float* __restrict__ h2, const float* __restrict__ zh,
                                float* __restrict__ o, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float z = 1.f / (1.f + expf(-zh[i]));
    o[i] = z * tanhf(h1[i]) + (1.f - z) * tanhf(h2[i]);
  }
}
__global__ void gate_bwd_kernel(const float* __restrict__ h1, const float* __restrict__ h2, const float* __restrict__ zh,
                                const float* __restrict__ g, float* __restrict__ dh1, float* __restrict__ dh2,
                                float* __restrict__ dzh, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float z = 1.f / (1.f + expf(-zh[i]));
    const float t1 = tanhf(h1[i]), t2 = tanhf(h2[i]), gi = g[i];
    dh1[i] = gi * z * (1.f - t1 * t1);
    dh2[i] = gi * (1.f - z) * (1.f - t2 * t2);
    dzh[i] = gi * (t1 - t2) * z * (1.f - z);
  }
}

// pooled[b][d] = mean_n x[b][n][d]   |   dx[b][n][d] = dpooled[b][d] / N
__global__ void mean_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int N, int D) {
  const long long total = static_cast<long long>(B) * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / D;
    const int d = static_cast<int>(i - b * D);
    const float* p = x + b * N * D + d;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += p[static_cast<long long>(n) * D];
    out[i] = s / N;
  }
}
__global__ void mean_pool_bwd_kernel(const float* __restrict__ dp, float* __restrict__ dx, int B, int N, int D) {
  const long long total = static_cast<long long>(B) * N * D;
  const float inv = 1.f / N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / (static_cast<long long>(N) * D);
    const int d = static_cast<int>(i % D);
    dx[i] = dp[b * D + d] * inv;
  }
}

template <typename TO>
__global__ void mask_scale_kernel(const float* src, long long lds, TO* dst, long long ldd, int rows, int cols, const Drop drop,
                                  long long drop_ld) {
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    float v = src[r * lds + c];
    if (drop.thresh) v = drop_apply(drop, v, static_cast<unsigned long long>(r) * drop_ld + c);
    dst[r * ldd + c] = static_cast<TO>(v);
  }
}
__global__ void dropout_mask_kernel(float* out, int rows, int cols, long long ld, const Drop drop) {
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    out[i] = (!drop.thresh || drop_keep(drop, static_cast<unsigned long long>(r) * ld + c)) ? 1.f : 0.f;
  }
}

// dy *= (y > 0)   (ReLU backward, in place)
__global__ void relu_bwd_kernel(float* dy, const float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}

inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = static_cast<long long>(kNumSms) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int cast_pad_bf16(const float* src, long long lds, void* dst, long long ldd, int rows, int cols, cudaStream_t s) {
  LaunchScope scope("cast_pad_bf16", s);
  if (rows <= 0 || cols <= 0 || ldd < cols) return M2_ERR_ARG;
  if (lds % 4 == 0 && ldd % 8 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int gx = ceil_div(static_cast<int>(ldd), 256);
    int gy = ceil_div(rows, 8);
    const int cap = ceil_div(148 * 8, gx);
    if (gy > cap) gy = cap;
    cast_pad_bf16_vec_kernel<<<dim3(gx, gy), 256, 0, s>>>(src, lds, static_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  } else {
    cast_pad_bf16_kernel<<<grid_for(static_cast<long long>(rows) * ldd, 256), 256, 0, s>>>(src, lds, static_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  }
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int cast_pad_bf16_multi(const long long* table_dev, int n, cudaStream_t s) {
  LaunchScope scope("cast_pad_bf16_multi", s);
  if (!table_dev || n <= 0 || n > 65535) return M2_ERR_ARG;
  cast_pad_bf16_multi_kernel<<<dim3(ceil_div(148 * 4, n) < 1 ? 1 : ceil_div(148 * 4, n), n), 256, 0, s>>>(table_dev);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int ln_fwd(const float* x, const float* w, const float* b, void* out, int out_bf16, int rows, int D, int N,
           long long out_bstride, float* mean, float* rstd, cudaStream_t s) {
  LaunchScope scope("ln_fwd", s);
  if (rows <= 0 || D <= 0 || N <= 0) return M2_ERR_ARG;
  const int grid = grid_for(static_cast<long long>(rows) * 32, 256);
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(b) |
                    reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (D % 4 == 0 && D <= 1024 && al && out_bstride % 4 == 0) {
    const int kv = ceil_div(D, 128);
#define M2_LNF(V_)                                                                                                       \
  do {                                                                                                                   \
    if (out_bf16) ln_fwd_vec_kernel<true, V_><<<grid, 256, 0, s>>>(x, w, b, out, rows, D, N, out_bstride, mean, rstd);     \
    else ln_fwd_vec_kernel<false, V_><<<grid, 256, 0, s>>>(x, w, b, out, rows, D, N, out_bstride, mean, rstd);             \
  } while (0)
    if (kv <= 1) M2_LNF(1); else if (kv <= 2) M2_LNF(2); else if (kv <= 4) M2_LNF(4); else if (kv <= 6) M2_LNF(6); else M2_LNF(8);
#undef M2_LNF
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
  if (out_bf16) ln_fwd_kernel<true><<<grid, 256, 0, s>>>(x, w, b, out, rows, D, N, out_bstride, mean, rstd);
  else ln_fwd_kernel<false><<<grid, 256, 0, s>>>(x, w, b, out, rows, D, N, out_bstride, mean, rstd);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int ln_bwd(const float* dy, long long dy_bstride, int N, const float* x, const float* w, const float* dres, float* dx,
           float* dw, float* db, int rows, int D, cudaStream_t s) {
  LaunchScope scope("ln_bwd", s);
  if (rows <= 0 || D <= 0 || D > 1024 || N <= 0) return M2_ERR_ARG;
  int grid = grid_for(static_cast<long long>(rows) * 32, 256);
  if (grid > kNumSms * 8) grid = kNumSms * 8;   // 64 resident warps per SM: the kernel is load-latency bound
  const size_t sm = 2 * D * sizeof(float);
  const bool al = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) |
                    reinterpret_cast<uintptr_t>(dres) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  if (D % 4 == 0 && al && dy_bstride % 4 == 0) {
    const int kv = ceil_div(D, 128);
    // exactly the CTAs that are resident (launch bounds of ln_bwd_vec_kernel): every CTA ends with 2 D global atomics on the
    // same 2 D addresses, four waves of CTAs only multiplied those
    const int resident = kNumSms * (kv <= 2 ? 4 : (kv <= 6 ? 2 : 1));
    if (grid > resident) grid = resident;
#define M2_LNBV(V_) ln_bwd_vec_kernel<V_><<<grid, 256, sm, s>>>(dy, dy_bstride, N, x, w, dres, dx, dw, db, rows, D)
    if (kv <= 1) M2_LNBV(1); else if (kv <= 2) M2_LNBV(2); else if (kv <= 4) M2_LNBV(4); else if (kv <= 6) M2_LNBV(6); else M2_LNBV(8);
#undef M2_LNBV
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
  const int per = ceil_div(D, 32);
#define M2_LNB(P_) ln_bwd_kernel<P_><<<grid, 256, sm, s>>>(dy, dy_bstride, N, x, w, dres, dx, dw, db, rows, D)
  if (per <= 1) M2_LNB(1); else if (per <= 2) M2_LNB(2); else if (per <= 4) M2_LNB(4); else if (per <= 8) M2_LNB(8);
  else if (per <= 12) M2_LNB(12); else if (per <= 16) M2_LNB(16); else if (per <= 24) M2_LNB(24); else M2_LNB(32);
#undef M2_LNB
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int colsum_f32(const float* src, long long ld, int rows, int cols, float* out, cudaStream_t s) {
  LaunchScope scope("colsum_f32", s);
  if (rows <= 0 || cols <= 0) return M2_ERR_ARG;
  if (ld % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int gx = ceil_div(cols, 128);
    dim3 grid(gx, max(1, min(ceil_div(rows, 32), ceil_div(148 * 4, gx))));
    colsum_vec_kernel<float, 4><<<grid, 256, 0, s>>>(src, ld, rows, cols, out);
  } else {
    dim3 grid(ceil_div(cols, 32), min(ceil_div(rows, 8), 64));
    colsum_kernel<float><<<grid, 256, 0, s>>>(src, ld, rows, cols, out);
  }
  M2_LAUNCH_CHECK();
  return M2_OK;
}
int colsum_bf16(const void* src, long long ld, int rows, int cols, float* out, cudaStream_t s) {
  LaunchScope scope("colsum_bf16", s);
  if (rows <= 0 || cols <= 0) return M2_ERR_ARG;
  if (ld % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int gx = ceil_div(cols, 256);
    dim3 grid(gx, max(1, min(ceil_div(rows, 32), ceil_div(148 * 4, gx))));
    colsum_vec_kernel<__nv_bfloat16, 8><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src), ld, rows, cols, out);
  } else {
    dim3 grid(ceil_div(cols, 32), min(ceil_div(rows, 8), 64));
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src), ld, rows, cols, out);
  }
  M2_LAUNCH_CHECK();
  return M2_OK;
}

// bf16 in, bf16 out, EIGHT consecutive columns per thread (16-byte accesses): the unfused bf16 backward keeps its
// [rows x hidden] intermediates H and dG in bf16 (half the bytes of the fp32 round trip of round 1).
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__global__ void __launch_bounds__(256) gelu_fwd_bwd_bf16x8_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* dg,
                                                                  long long n8, int cols8, long long ld_in,
                                                                  __nv_bfloat16* __restrict__ g_out, __nv_bfloat16* dh_out,
                                                                  long long ld_out, const Drop drop, long long drop_ld) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols8;
    const int c = static_cast<int>(i - r * cols8) * 8;
    const uint4 hv = *reinterpret_cast<const uint4*>(h + r * ld_in + c);
    const uint4 dv = *reinterpret_cast<const uint4*>(dg + r * ld_in + c);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
    float g[8], d[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 dgel;
      const float2 gg = gelu2_grad(bf16x2_to_f2(hw[e]), dgel);
      const float2 dd = bf16x2_to_f2(dw[e]);
      g[2 * e] = gg.x; g[2 * e + 1] = gg.y;
      d[2 * e] = dd.x * dgel.x; d[2 * e + 1] = dd.y * dgel.y;
    }
    if (drop.thresh) {
      const unsigned long long idx = static_cast<unsigned long long>(r) * drop_ld + c;
      float4 g0 = make_float4(g[0], g[1], g[2], g[3]), g1 = make_float4(g[4], g[5], g[6], g[7]);
      float4 d0 = make_float4(d[0], d[1], d[2], d[3]), d1 = make_float4(d[4], d[5], d[6], d[7]);
      drop_apply4(drop, g0, idx); drop_apply4(drop, g1, idx + 4);
      drop_apply4(drop, d0, idx); drop_apply4(drop, d1, idx + 4);
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
      d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
    }
    *reinterpret_cast<uint4*>(g_out + r * ld_out + c) =
        make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
    *reinterpret_cast<uint4*>(dh_out + r * ld_out + c) =
        make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
  }
}

int mask_scale(const float* src, long long lds, void* dst, int dst_bf16, long long ldd, int rows, int cols, float drop_p,
               unsigned long long seed, int site, long long drop_ld, cudaStream_t s) {
  LaunchScope scope("mask_scale", s);
  const long long n = static_cast<long long>(rows) * cols;
  if (n <= 0) return M2_ERR_ARG;
  const Drop d = make_drop(drop_p, seed, site);
  if (dst_bf16) mask_scale_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, s>>>(src, lds, static_cast<__nv_bfloat16*>(dst), ldd, rows, cols, d, drop_ld);
  else mask_scale_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(src, lds, static_cast<float*>(dst), ldd, rows, cols, d, drop_ld);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int dropout_mask(float* out, int rows, int cols, long long ld, float drop_p, unsigned long long seed, int site, cudaStream_t s) {
  LaunchScope scope("dropout_mask", s);
  const long long n = static_cast<long long>(rows) * cols;
  if (n <= 0 || !out) return M2_ERR_ARG;
  dropout_mask_kernel<<<grid_for(n, 256), 256, 0, s>>>(out, rows, cols, ld, make_drop(drop_p, seed, site));
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int gelu_fwd_bwd(const void* h_, const void* dg_, int in_bf16, int rows, int cols, long long ld_in, void* g_out, void* dh_out,
                 long long ld_out, int out_bf16, float drop_p, unsigned long long seed, int site, long long drop_ld,
                 cudaStream_t s) {
  const Drop drop = make_drop(drop_p, seed, site);
  LaunchScope scope("gelu_fwd_bwd", s);
  const long long n = static_cast<long long>(rows) * cols;
  if (n <= 0 || (in_bf16 && !out_bf16)) return M2_ERR_ARG;
  if (in_bf16) {
    const __nv_bfloat16* hb = static_cast<const __nv_bfloat16*>(h_);
    const __nv_bfloat16* dgb = static_cast<const __nv_bfloat16*>(dg_);
    const bool al8 = ((reinterpret_cast<uintptr_t>(h_) | reinterpret_cast<uintptr_t>(dg_) | reinterpret_cast<uintptr_t>(g_out) |
                       reinterpret_cast<uintptr_t>(dh_out)) & 15) == 0;
    if (al8 && cols % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && drop_ld % 4 == 0)
      gelu_fwd_bwd_bf16x8_kernel<<<grid_for(n / 8, 256), 256, 0, s>>>(hb, dgb, n / 8, cols / 8, ld_in, static_cast<__nv_bfloat16*>(g_out),
                                                                     static_cast<__nv_bfloat16*>(dh_out), ld_out, drop, drop_ld);
    else
      gelu_fwd_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid_for(n, 256), 256, 0, s>>>(
          hb, dgb, n, cols, ld_in, static_cast<__nv_bfloat16*>(g_out), static_cast<__nv_bfloat16*>(dh_out), ld_out, drop, drop_ld);
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
  const float* h = static_cast<const float*>(h_);
  const float* dg = static_cast<const float*>(dg_);
  const bool al = ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dg) | reinterpret_cast<uintptr_t>(g_out) |
                    reinterpret_cast<uintptr_t>(dh_out)) & 15) == 0;
  if (al && cols % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && drop_ld % 4 == 0) {
    if (out_bf16)
      gelu_fwd_bwd_vec4_kernel<__nv_bfloat16><<<grid_for(n / 4, 256), 256, 0, s>>>(
          h, dg, n / 4, cols / 4, ld_in, static_cast<__nv_bfloat16*>(g_out), static_cast<__nv_bfloat16*>(dh_out), ld_out, drop, drop_ld);
    else
      gelu_fwd_bwd_vec4_kernel<float><<<grid_for(n / 4, 256), 256, 0, s>>>(h, dg, n / 4, cols / 4, ld_in, static_cast<float*>(g_out),
                                                                         static_cast<float*>(dh_out), ld_out, drop, drop_ld);
    M2_LAUNCH_CHECK();
    return M2_OK;
  }
  if (out_bf16)
    gelu_fwd_bwd_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, s>>>(h, dg, n, cols, ld_in, static_cast<__nv_bfloat16*>(g_out),
                                                                       static_cast<__nv_bfloat16*>(dh_out), ld_out, drop, drop_ld);
  else
    gelu_fwd_bwd_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(h, dg, n, cols, ld_in, static_cast<float*>(g_out),
                                                               static_cast<float*>(dh_out), ld_out, drop, drop_ld);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int patch_gather(const void* img, int img_bf16, void* cols, int out_bf16, int B, int cin, int H, int W, int P, long long ld,
                 cudaStream_t s) {
  LaunchScope scope("patch_gather", s);
  if (B <= 0 || cin <= 0 || P <= 0 || H % P || W % P || ld < static_cast<long long>(cin) * P * P || ld > 12000) return M2_ERR_ARG;
  const long long rows_ll = static_cast<long long>(B) * (H / P) * (W / P);
  if (rows_ll >= (1ll << 31)) return M2_ERR_ARG;
  const int rows = static_cast<int>(rows_ll), ldi = static_cast<int>(ld);
  const int threads = ld >= 1024 ? 256 : 128;
  const int grid = (rows + 3) / 4 < 148 * 16 ? (rows + 3) / 4 : 148 * 16;
  const size_t smem = ld * sizeof(int);
  const float* f = static_cast<const float*>(img);
  const __nv_bfloat16* h = static_cast<const __nv_bfloat16*>(img);
  if (out_bf16 && img_bf16) patch_gather_kernel<<<grid, threads, smem, s>>>(h, static_cast<__nv_bfloat16*>(cols), rows, cin, H, W, P, ldi);
  else if (out_bf16) patch_gather_kernel<<<grid, threads, smem, s>>>(f, static_cast<__nv_bfloat16*>(cols), rows, cin, H, W, P, ldi);
  else if (img_bf16) patch_gather_kernel<<<grid, threads, smem, s>>>(h, static_cast<float*>(cols), rows, cin, H, W, P, ldi);
  else patch_gather_kernel<<<grid, threads, smem, s>>>(f, static_cast<float*>(cols), rows, cin, H, W, P, ldi);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int concat_copy(const float* src, long long src_bstride, float* dst, long long dst_bstride, int B, long long per_batch,
                int accumulate, cudaStream_t s) {
  LaunchScope scope("concat_copy", s);
  if (B <= 0 || per_batch <= 0) return M2_ERR_ARG;
  concat_copy_kernel<<<grid_for(B * per_batch, 256), 256, 0, s>>>(src, src_bstride, dst, dst_bstride, B, per_batch, accumulate);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int mean_pool_fwd(const float* x, float* out, int B, int N, int D, cudaStream_t s) {
  LaunchScope scope("mean_pool_fwd", s);
  if (B <= 0 || N <= 0 || D <= 0) return M2_ERR_ARG;
  mean_pool_fwd_kernel<<<grid_for(static_cast<long long>(B) * D, 256), 256, 0, s>>>(x, out, B, N, D);
  M2_LAUNCH_CHECK();
  return M2_OK;
}
int mean_pool_bwd(const float* dp, float* dx, int B, int N, int D, cudaStream_t s) {
  LaunchScope scope("mean_pool_bwd", s);
  if (B <= 0 || N <= 0 || D <= 0) return M2_ERR_ARG;
  mean_pool_bwd_kernel<<<grid_for(static_cast<long long>(B) * N * D, 256), 256, 0, s>>>(dp, dx, B, N, D);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int relu_bwd(float* dy, const float* y, long long n, cudaStream_t s) {
  LaunchScope scope("relu_bwd", s);
  if (n <= 0) return M2_ERR_ARG;
  relu_bwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(dy, y, n);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int fuse2_fwd(const float* a, const float* b, float* o, long long n, int mode, cudaStream_t s) {
  LaunchScope scope("fuse2_fwd", s);
  if (n <= 0 || (mode != 1 && mode != 2)) return M2_ERR_ARG;
  fuse2_fwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(a, b, o, n, mode);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int fuse2_max_bwd(const float* a, const float* b, const float* g, float* da, float* db, long long n, cudaStream_t s) {
  LaunchScope scope("fuse2_max_bwd", s);
  if (n <= 0) return M2_ERR_ARG;
  fuse2_max_bwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(a, b, g, da, db, n);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int gate_fwd(const float* h1, const float* h2, const float* zh, float* o, long long n, cudaStream_t s) {
  LaunchScope scope("gate_fwd", s);
  if (n <= 0) return M2_ERR_ARG;
  gate_fwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(h1, h2, zh, o, n);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int gate_bwd(const float* h1, const float* h2, const float* zh, const float* g, float* dh1, float* dh2, float* dzh, long long n,
             cudaStream_t s) {
  LaunchScope scope("gate_bwd", s);
  if (n <= 0) return M2_ERR_ARG;
  gate_bwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(h1, h2, zh, g, dh1, dh2, dzh, n);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int rowsum_mod_bf16(const void* src, long long ld, int rows, int cols, int period, float* out, cudaStream_t s) {
  LaunchScope scope("rowsum_mod_bf16", s);
  if (rows <= 0 || cols <= 0 || period <= 0 || period > 4096) return M2_ERR_ARG;
  int grid = grid_for(static_cast<long long>(rows) * 32, 256);
  if (grid > kNumSms * 4) grid = kNumSms * 4;
  rowsum_mod_bf16_kernel<<<grid, 256, period * sizeof(float), s>>>(static_cast<const __nv_bfloat16*>(src), ld, rows, cols, period, out);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

int add_f32(const float* a, const float* b, float* o, long long n, cudaStream_t s) {
  LaunchScope scope("add_f32", s);
  if (n <= 0) return M2_ERR_ARG;
  add_kernel<<<grid_for(n, 256), 256, 0, s>>>(a, b, o, n);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace m2
