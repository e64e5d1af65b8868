// extern "C" surface of libm2b200.so (include/m2b200.h): argument validation, workspace carving and the kernel
// sequences behind each composite op.  No allocation, no sync, status codes only.
#include "../../include/m2b200.h"

#include "common.cuh"
#include "kernels.h"

using namespace m2;

namespace {

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline size_t up256(size_t b) { return (b + 255) & ~size_t(255); }
inline int up8(int v) { return (v + 7) & ~7; }

struct Carver {
  uint8_t* base; size_t size; size_t off = 0; bool ok = true;
  Carver(void* b, size_t s) : base(static_cast<uint8_t*>(b)), size(s) {}
  template <typename T> T* take(size_t count) {
    const size_t bytes = up256(count * sizeof(T));
    if (!base || off + bytes > size) { ok = false; return nullptr; }
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

#define M2_TRY(expr)        \
  do {                      \
    int rc__ = (expr);      \
    if (rc__) return rc__;  \
  } while (0)

GemmArgs gemm_args(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, int M, int N, int K,
                   void* C, int c_bf16, long long ldc) {
  GemmArgs g = {};
  g.A = A; g.a_mn = a_mn; g.lda = lda; g.B = B; g.b_mn = b_mn; g.ldb = ldb;
  g.M = M; g.N = N; g.K = K; g.batch = 1;
  g.C = C; g.c_bf16 = c_bf16; g.ldc = ldc; g.splitk = 1;
  return g;
}

// split-K factor for a token-axis (K = M rows) weight-gradient GEMM so that ~one wave of CTAs is launched
int wgrad_splitk(int m_out, int n_out, int k) {
  const int tiles = ceil_div(m_out, 128) * ceil_div(n_out, 128);
  int sk = (148 * 2) / tiles;   // two resident CTAs per SM, never a partial extra wave
  const int k_tiles = ceil_div(k, 64);
  if (sk > k_tiles) sk = k_tiles;
  return sk < 1 ? 1 : sk;
}

enum ChainPath { kPathF32 = 0, kPathFused = 1, kPathUnfusedBf16 = 2 };
ChainPath pick_path(int D, int precision, bool backward) {
  if (precision == M2B200_FP32) return kPathF32;
  return (backward ? chain_bwd_supported(D) : chain_fwd_supported(D)) ? kPathFused : kPathUnfusedBf16;
}

}  // namespace

extern "C" {

int m2b200_abi_version(void) { return 1; }

const char* m2b200_status_string(int st) {
  switch (st) {
    case M2B200_OK: return "ok";
    case M2B200_ERR_ARG: return "invalid argument / unsupported shape";
    case M2B200_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case M2B200_ERR_WORKSPACE: return "workspace missing or too small";
    case M2B200_ERR_LAUNCH: return "CUDA launch failed";
    case M2B200_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    default: return "unknown status";
  }
}

int m2b200_cast_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, void* stream) {
  if (!src || !dst) return M2_ERR_ARG;
  return cast_pad_bf16(src, lds, dst, ldd, rows, cols, S(stream));
}

int m2b200_cast_bf16_multi(const int64_t* table_dev, int n, void* stream) {
  return cast_pad_bf16_multi(reinterpret_cast<const long long*>(table_dev), n, S(stream));
}

int m2b200_gemm(int precision, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N,
                int K, int batch, int64_t a_batch_rows, int64_t b_batch_rows, const float* bias, int bias_mode, int act,
                const float* residual, int64_t ldr, int64_t r_batch_stride, void* C, int c_bf16, int64_t ldc,
                int64_t c_batch_stride, int accumulate, int splitk, void* stream) {
  GemmArgs g = {};
  g.A = A; g.a_mn = a_mn; g.lda = lda; g.B = B; g.b_mn = b_mn; g.ldb = ldb;
  g.M = M; g.N = N; g.K = K; g.batch = batch; g.a_batch_rows = a_batch_rows; g.b_batch_rows = b_batch_rows;
  g.bias = bias; g.bias_mode = bias_mode; g.act = act;
  g.residual = residual; g.ldr = ldr; g.r_batch_stride = r_batch_stride;
  g.C = C; g.c_bf16 = c_bf16; g.ldc = ldc; g.c_batch_stride = c_batch_stride;
  g.accumulate = accumulate; g.splitk = splitk < 1 ? 1 : splitk;
  return precision == M2B200_FP32 ? gemm_f32_simt(g, S(stream)) : gemm_bf16_umma(g, S(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ token mixing
namespace {

// Token mixing for contraction lengths the register-tile kernels do not cover (MM-IMDB-shaped and Scaled configs: N = 196..708
// tokens, T = 16..512): LayerNorm -> bf16, then the two contractions as BATCHED tcgen05 GEMMs with the activation tile
// [N x D] of a sample as the MN-major B operand in its native row-major layout (no permute) and the weights as
// the shared A operand; exact erf GELU in the first GEMM's epilogue, bias per output row, residual in the second's.
// The backward recomputes LN and H, and reduces the per-sample weight-gradient GEMMs into ONE gradient with atomics
// (c_batch_stride = 0).  On the CUDA-core fallback these shapes took 350-430 ms per step (profiles/r01_large_configs.json).
bool token_mix_gemm_path(int precision, int N, int D, int T) {
  if (precision == M2B200_FP32 || token_mix_mma_supported(N, D, T)) return false;
  return D % 8 == 0 && static_cast<long long>(N) * T >= 2048;
}

struct TokWs {
  __nv_bfloat16 *xn_b, *w1b, *w2b, *g_b, *du_b, *dh_b, *h_b, *dg_b;
  float *f1, *f2, *dxn;
};
size_t token_mix_gemm_ws(int B, int N, int D, int T, bool backward, TokWs* out, void* base, size_t bytes, bool* ok) {
  Carver ws(base, bytes);
  const size_t bn = static_cast<size_t>(B) * N * D, bt = static_cast<size_t>(B) * T * D;
  TokWs w = {};
  w.xn_b = ws.take<__nv_bfloat16>(bn);
  w.w1b = ws.take<__nv_bfloat16>(static_cast<size_t>(T) * up8(N));
  w.w2b = ws.take<__nv_bfloat16>(static_cast<size_t>(N) * up8(T));
  w.g_b = ws.take<__nv_bfloat16>(bt);
  if (backward) {
    w.h_b = ws.take<__nv_bfloat16>(bt);     // H pre-activation and dG: bf16 (rounded to bf16 as G / dH right after anyway)
    w.dg_b = ws.take<__nv_bfloat16>(bt);
    w.du_b = ws.take<__nv_bfloat16>(bn);
    w.dh_b = ws.take<__nv_bfloat16>(bt);
    w.dxn = ws.take<float>(bn);
  } else {
    w.f1 = ws.take<float>(bt);        // GELU output before the mask (dropout only)
    w.f2 = ws.take<float>(bn);        // branch output before the mask (dropout only)
  }
  if (out) *out = w;
  if (ok) *ok = ws.ok;
  return ws.off;
}
size_t token_mix_gemm_ws_bytes(int B, int N, int D, int T, bool backward) {
  // dry run over a fake base: Carver only does arithmetic
  Carver ws(reinterpret_cast<void*>(256), ~size_t(0) >> 1);
  const size_t bn = static_cast<size_t>(B) * N * D, bt = static_cast<size_t>(B) * T * D;
  ws.take<__nv_bfloat16>(bn); ws.take<__nv_bfloat16>(static_cast<size_t>(T) * up8(N)); ws.take<__nv_bfloat16>(static_cast<size_t>(N) * up8(T));
  ws.take<__nv_bfloat16>(bt);
  if (backward) { ws.take<__nv_bfloat16>(bt); ws.take<__nv_bfloat16>(bt); ws.take<__nv_bfloat16>(bn); ws.take<__nv_bfloat16>(bt); ws.take<float>(bn); }
  else { ws.take<float>(bt); ws.take<float>(bn); }
  return ws.off;
}

int token_mix_gemm_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
                       const float* b2, float* u, int B, int N, int D, int T, float p, uint64_t seed, void* workspace,
                       size_t workspace_bytes, cudaStream_t s) {
  TokWs w; bool ok;
  token_mix_gemm_ws(B, N, D, T, false, &w, workspace, workspace_bytes, &ok);
  if (!ok) return M2_ERR_WORKSPACE;
  const int n8 = up8(N), t8 = up8(T);
  const bool drop = p > 0.f;
  M2_TRY(ln_fwd(x, ln_w, ln_b, w.xn_b, 1, B * N, D, B * N, 0, nullptr, nullptr, s));
  M2_TRY(cast_pad_bf16(w1, N, w.w1b, n8, T, N, s));
  M2_TRY(cast_pad_bf16(w2, T, w.w2b, t8, N, T, s));
  // G_b [T x D] = GELU(W1 [T x N] . Xn_b [N x D] + b1 1^T)
  GemmArgs g1 = gemm_args(w.w1b, 0, n8, w.xn_b, 1, D, T, D, N, drop ? static_cast<void*>(w.f1) : static_cast<void*>(w.g_b), drop ? 0 : 1, D);
  g1.batch = B; g1.a_batch_rows = 0; g1.b_batch_rows = N; g1.c_batch_stride = static_cast<long long>(T) * D;
  g1.bias = b1; g1.bias_mode = 2; g1.act = M2B200_ACT_GELU;
  M2_TRY(gemm_bf16_umma(g1, s));
  if (drop) M2_TRY(mask_scale(w.f1, D, w.g_b, 1, D, B * T, D, p, seed, kSiteTokenHidden, D, s));
  // u_b [N x D] = x_b + Drop(W2 [N x T] . G_b [T x D] + b2 1^T)
  GemmArgs g2 = gemm_args(w.w2b, 0, t8, w.g_b, 1, D, N, D, T, drop ? static_cast<void*>(w.f2) : static_cast<void*>(u), 0, D);
  g2.batch = B; g2.a_batch_rows = 0; g2.b_batch_rows = T; g2.c_batch_stride = static_cast<long long>(N) * D;
  g2.bias = b2; g2.bias_mode = 2;
  if (!drop) { g2.residual = x; g2.ldr = D; g2.r_batch_stride = static_cast<long long>(N) * D; }
  M2_TRY(gemm_bf16_umma(g2, s));
  if (drop) {
    M2_TRY(mask_scale(w.f2, D, w.f2, 0, D, B * N, D, p, seed, kSiteTokenOut, D, s));
    M2_TRY(add_f32(x, w.f2, u, static_cast<long long>(B) * N * D, s));
  }
  return M2_OK;
}

int token_mix_gemm_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                       const float* w2, float* dx, float* dln_w, float* dln_b, float* dw1, float* db1, float* dw2, float* db2,
                       int B, int N, int D, int T, float p, uint64_t seed, void* workspace, size_t workspace_bytes,
                       cudaStream_t s) {
  TokWs w; bool ok;
  token_mix_gemm_ws(B, N, D, T, true, &w, workspace, workspace_bytes, &ok);
  if (!ok) return M2_ERR_WORKSPACE;
  const int n8 = up8(N), t8 = up8(T);
  const long long nd = static_cast<long long>(N) * D, td = static_cast<long long>(T) * D;
  M2_TRY(ln_fwd(x, ln_w, ln_b, w.xn_b, 1, B * N, D, B * N, 0, nullptr, nullptr, s));
  M2_TRY(cast_pad_bf16(w1, N, w.w1b, n8, T, N, s));
  M2_TRY(cast_pad_bf16(w2, T, w.w2b, t8, N, T, s));
  // gradient of the dropped branch output (mask + scale), as the bf16 GEMM operand
  if (p > 0.f) M2_TRY(mask_scale(du, D, w.du_b, 1, D, B * N, D, p, seed, kSiteTokenOut, D, s));
  else M2_TRY(cast_pad_bf16(du, D, w.du_b, D, B * N, D, s));
  // recompute H_b = W1 . Xn_b + b1 (pre-activation, bf16)
  GemmArgs gh = gemm_args(w.w1b, 0, n8, w.xn_b, 1, D, T, D, N, w.h_b, 1, D);
  gh.batch = B; gh.b_batch_rows = N; gh.c_batch_stride = td; gh.bias = b1; gh.bias_mode = 2;
  M2_TRY(gemm_bf16_umma(gh, s));
  // dG_b [T x D] = W2^T [T x N] . dU_b [N x D]      (A = W2 [N][T] consumed MN-major)
  GemmArgs gg = gemm_args(w.w2b, 1, t8, w.du_b, 1, D, T, D, N, w.dg_b, 1, D);
  gg.batch = B; gg.b_batch_rows = N; gg.c_batch_stride = td;
  M2_TRY(gemm_bf16_umma(gg, s));
  // G = Drop(GELU(H)), dH = dG * Drop'(.) * GELU'(H)
  M2_TRY(gelu_fwd_bwd(w.h_b, w.dg_b, 1, B * T, D, D, w.g_b, w.dh_b, D, 1, p, seed, kSiteTokenHidden, D, s));
  // dW2 [N x T] += sum_b dU_b [N x D] . G_b^T [D x T] ;  dW1 [T x N] += sum_b dH_b [T x D] . Xn_b^T [D x N]
  GemmArgs gw2 = gemm_args(w.du_b, 0, D, w.g_b, 0, D, N, T, D, dw2, 0, T);
  gw2.batch = B; gw2.a_batch_rows = N; gw2.b_batch_rows = T; gw2.c_batch_stride = 0; gw2.atomic_out = 1;
  M2_TRY(gemm_bf16_umma(gw2, s));
  GemmArgs gw1 = gemm_args(w.dh_b, 0, D, w.xn_b, 0, D, T, N, D, dw1, 0, N);
  gw1.batch = B; gw1.a_batch_rows = T; gw1.b_batch_rows = N; gw1.c_batch_stride = 0; gw1.atomic_out = 1;
  M2_TRY(gemm_bf16_umma(gw1, s));
  M2_TRY(rowsum_mod_bf16(w.dh_b, D, B * T, D, T, db1, s));
  M2_TRY(rowsum_mod_bf16(w.du_b, D, B * N, D, N, db2, s));
  // dXn_b [N x D] = W1^T [N x T] . dH_b [T x D]     (A = W1 [T][N] consumed MN-major)
  GemmArgs gx = gemm_args(w.w1b, 1, n8, w.dh_b, 1, D, N, D, T, w.dxn, 0, D);
  gx.batch = B; gx.b_batch_rows = T; gx.c_batch_stride = nd;
  M2_TRY(gemm_bf16_umma(gx, s));
  // dx = du (residual) + LayerNorm'(dxn);  dln_w / dln_b accumulate
  return ln_bwd(w.dxn, nd, N, x, ln_w, du, dx, dln_w, dln_b, B * N, D, s);
}

}  // namespace

extern "C" {

size_t m2b200_token_mix_fwd_workspace_bytes(int B, int N, int D, int T, int precision) {
  return token_mix_gemm_path(precision, N, D, T) ? token_mix_gemm_ws_bytes(B, N, D, T, false) : 0;
}

int m2b200_token_mix_fwd(const float* x, const float* ln_w, const float* ln_b, const float* wt1, const float* bt1,
                         const float* wt2, const float* bt2, float* u, int B, int N, int D, int T, int precision,
                         float dropout_p, uint64_t seed, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !ln_w || !ln_b || !wt1 || !bt1 || !wt2 || !bt2 || !u) return M2_ERR_ARG;
  if (dropout_p < 0.f || dropout_p >= 1.f) return M2_ERR_ARG;
  if (token_mix_gemm_path(precision, N, D, T))
    return token_mix_gemm_fwd(x, ln_w, ln_b, wt1, bt1, wt2, bt2, u, B, N, D, T, dropout_p, seed, workspace, workspace_bytes,
                              S(stream));
  return token_mix_fwd(x, ln_w, ln_b, wt1, bt1, wt2, bt2, u, B, N, D, T, precision == M2B200_FP32, dropout_p, seed, S(stream));
}

size_t m2b200_token_mix_bwd_workspace_bytes(int B, int N, int D, int T, int precision) {
  if (token_mix_gemm_path(precision, N, D, T)) return token_mix_gemm_ws_bytes(B, N, D, T, true);
  return up256(static_cast<size_t>(B) * N * D * sizeof(float));
}

int m2b200_token_mix_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* wt1,
                         const float* bt1, const float* wt2, float* dx, float* dln_w, float* dln_b, float* dwt1,
                         float* dbt1, float* dwt2, float* dbt2, int B, int N, int D, int T, int precision, float dropout_p,
                         uint64_t seed, void* workspace, size_t workspace_bytes, void* stream) {
  if (!du || !x || !ln_w || !ln_b || !wt1 || !bt1 || !wt2 || !dx || !dln_w || !dln_b || !dwt1 || !dbt1 || !dwt2 || !dbt2)
    return M2_ERR_ARG;
  if (dropout_p < 0.f || dropout_p >= 1.f) return M2_ERR_ARG;
  if (token_mix_gemm_path(precision, N, D, T))
    return token_mix_gemm_bwd(du, x, ln_w, ln_b, wt1, bt1, wt2, dx, dln_w, dln_b, dwt1, dbt1, dwt2, dbt2, B, N, D, T, dropout_p,
                              seed, workspace, workspace_bytes, S(stream));
  if (precision != M2B200_FP32 && token_generation() != 1 && token_mix_mma_supported(N, D, T))   // LN backward fused in
    return token_mix_mma_bwd(du, x, ln_w, ln_b, wt1, bt1, wt2, dx, dln_w, dln_b, dwt1, dbt1, dwt2, dbt2, B, N, D, T, dropout_p,
                             seed, S(stream));
  Carver ws(workspace, workspace_bytes);
  float* dxn = ws.take<float>(static_cast<size_t>(B) * N * D);
  if (!ws.ok) return M2_ERR_WORKSPACE;
  M2_TRY(token_mix_bwd(du, x, ln_w, ln_b, wt1, bt1, wt2, dxn, dwt1, dbt1, dwt2, dbt2, B, N, D, T,
                       precision == M2B200_FP32, dropout_p, seed, S(stream)));
  // dx = du (residual) + LayerNorm'(dxn);  dln_w/dln_b accumulate
  return ln_bwd(dxn, static_cast<long long>(N) * D, N, x, ln_w, du, dx, dln_w, dln_b, B * N, D, S(stream));
}

// ------------------------------------------------------------------------------------------ channel mixing
size_t m2b200_channel_mix_workspace_bytes(int M, int D, int C, int precision, int backward) {
  const size_t m = M, d = D, c = C, c8 = up8(C);
  const ChainPath path = pick_path(D, precision, backward != 0);
  size_t b = 0;
  if (!backward) {
    if (path == kPathF32) b = up256(m * d * 4) + up256(m * c * 4);
    else if (path == kPathUnfusedBf16) b = up256(m * d * 2) + up256(m * c8 * 2);
  } else {
    if (path == kPathF32) b = 3 * up256(m * d * 4) + 3 * up256(m * c * 4);
    else if (path == kPathFused && chain_generation() == 4 && chain_fwd_ts_supported(D))
      b = 2 * up256(m * d * 2) + up256(m * ((c + 63) / 64 * 64) * 2);   // bf16 LN(u), dY and the spilled dH (chunk-major)
    else if (path == kPathFused && chain_generation() != 1 && chain_generation() != 3 && chain_fwd_ts_supported(D))
      b = 2 * up256(m * d * 2);   // bf16 LN(u) and dY only: the weight-gradient kernel recomputes G / dH on chip
    else if (path == kPathFused) b = 2 * up256(m * d * 2) + 2 * up256(m * c8 * 2) + up256(m * d * 4);
    else b = 2 * up256(m * d * 2) + 4 * up256(m * c8 * 2) + up256(m * d * 4);   // bf16 LN(u), dY; bf16 G, dH, H, dG; fp32 dXn
  }
  return b;
}

int m2b200_channel_mix_fwd(const float* u, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                           const float* w2, const float* b2, const void* w1b, const void* w2b, int ldw2, float* y, int M,
                           int D, int C, int precision, float dropout_p, uint64_t seed, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!u || !ln_w || !ln_b || !b1 || !b2 || !y || M <= 0 || D <= 0 || C <= 0) return M2_ERR_ARG;
  if (!al16(u) || !al16(y) || !al16(ln_w) || !al16(ln_b) || !al16(b1) || !al16(b2) || D % 4) return M2_ERR_ALIGN;
  if (dropout_p < 0.f || dropout_p >= 1.f) return M2_ERR_ARG;
  cudaStream_t s = S(stream);
  const ChainPath path = pick_path(D, precision, false);
  const int c8 = up8(C);
  Carver ws(workspace, workspace_bytes);
  if (path == kPathF32) {
    if (!w1 || !w2) return M2_ERR_ARG;
    float* xn = ws.take<float>(static_cast<size_t>(M) * D);
    float* h = ws.take<float>(static_cast<size_t>(M) * C);
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(ln_fwd(u, ln_w, ln_b, xn, 0, M, D, M, 0, nullptr, nullptr, s));
    GemmArgs g1 = gemm_args(xn, 0, D, w1, 0, D, M, C, D, h, 0, C);
    g1.bias = b1; g1.bias_mode = 1; g1.act = 1;
    g1.drop_p = dropout_p; g1.drop_seed = seed; g1.drop_site = kSiteChannelHidden; g1.drop_ld = c8;
    M2_TRY(gemm_f32_simt(g1, s));
    GemmArgs g2 = gemm_args(h, 0, C, w2, 0, C, M, D, C, y, 0, D);
    g2.bias = b2; g2.bias_mode = 1; g2.residual = u; g2.ldr = D;
    g2.drop_p = dropout_p; g2.drop_seed = seed; g2.drop_site = kSiteChannelOut; g2.drop_ld = D;
    return gemm_f32_simt(g2, s);
  }
  if (!w1b || !w2b || D % 8 || ldw2 % 8 || ldw2 < C) return M2_ERR_ARG;
  if (path == kPathFused) return chain_fwd(u, ln_w, ln_b, w1b, b1, w2b, ldw2, b2, y, M, D, C, 0, dropout_p, seed, s);
  // unfused bf16: LN -> GEMM(+b1, GELU) -> GEMM(+b2, +u)
  __nv_bfloat16* xn = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * D);
  __nv_bfloat16* h = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * c8);
  if (!ws.ok) return M2_ERR_WORKSPACE;
  M2_TRY(ln_fwd(u, ln_w, ln_b, xn, 1, M, D, M, 0, nullptr, nullptr, s));
  GemmArgs g1 = gemm_args(xn, 0, D, w1b, 0, D, M, C, D, h, 1, c8);
  g1.bias = b1; g1.bias_mode = 1; g1.act = 1;
  g1.drop_p = dropout_p; g1.drop_seed = seed; g1.drop_site = kSiteChannelHidden; g1.drop_ld = c8;
  M2_TRY(gemm_bf16_umma(g1, s));
  GemmArgs g2 = gemm_args(h, 0, c8, w2b, 0, ldw2, M, D, C, y, 0, D);
  g2.bias = b2; g2.bias_mode = 1; g2.residual = u; g2.ldr = D;
  g2.drop_p = dropout_p; g2.drop_seed = seed; g2.drop_site = kSiteChannelOut; g2.drop_ld = D;
  return gemm_bf16_umma(g2, s);
}

int m2b200_channel_mix_bwd(const float* dy, const float* u, const float* ln_w, const float* ln_b, const float* w1,
                           const float* b1, const float* w2, const void* w1b, const void* w2b, int ldw2, float* du,
                           float* dln_w, float* dln_b, float* dw1, float* db1, float* dw2, float* db2, int M, int D, int C,
                           int precision, float dropout_p, uint64_t seed, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (!dy || !u || !ln_w || !ln_b || !b1 || !du || !dln_w || !dln_b || !dw1 || !db1 || !dw2 || !db2 || M <= 0 || D <= 0 ||
      C <= 0)
    return M2_ERR_ARG;
  if (!al16(dy) || !al16(u) || !al16(du) || !al16(ln_w) || !al16(ln_b) || !al16(b1) || D % 4) return M2_ERR_ALIGN;
  if (dropout_p < 0.f || dropout_p >= 1.f) return M2_ERR_ARG;
  const bool drop = dropout_p > 0.f;
  const int c8 = up8(C);
  cudaStream_t s = S(stream);
  const ChainPath path = pick_path(D, precision, true);
  Carver ws(workspace, workspace_bytes);
  const size_t md = static_cast<size_t>(M) * D, mc = static_cast<size_t>(M) * C;
  if (path == kPathF32) {
    if (!w1 || !w2) return M2_ERR_ARG;
    float* xn = ws.take<float>(md);
    float* dxn = ws.take<float>(md);
    float* dym = ws.take<float>(md);
    float* h = ws.take<float>(mc);
    float* gbuf = ws.take<float>(mc);
    float* dh = ws.take<float>(mc);
    if (!ws.ok) return M2_ERR_WORKSPACE;
    const float* dyb = dy;   // gradient of the (dropped) branch output
    if (drop) {
      M2_TRY(mask_scale(dy, D, dym, 0, D, M, D, dropout_p, seed, kSiteChannelOut, D, s));
      dyb = dym;
    }
    M2_TRY(ln_fwd(u, ln_w, ln_b, xn, 0, M, D, M, 0, nullptr, nullptr, s));
    GemmArgs gh = gemm_args(xn, 0, D, w1, 0, D, M, C, D, h, 0, C);          // H = Xn W1^T + b1
    gh.bias = b1; gh.bias_mode = 1;
    M2_TRY(gemm_f32_simt(gh, s));
    GemmArgs gg = gemm_args(dyb, 0, D, w2, 1, C, M, C, D, dh, 0, C);         // dG = dY W2   (W2 [D][C] as [K][N])
    M2_TRY(gemm_f32_simt(gg, s));
    M2_TRY(gelu_fwd_bwd(h, dh, 0, M, C, C, gbuf, dh, C, 0, dropout_p, seed, kSiteChannelHidden, c8, s));   // G, dH in place
    GemmArgs gw2 = gemm_args(dyb, 1, D, gbuf, 1, C, D, C, M, dw2, 0, C);     // dW2 += dY^T G
    gw2.accumulate = 1;
    M2_TRY(gemm_f32_simt(gw2, s));
    GemmArgs gw1 = gemm_args(dh, 1, C, xn, 1, D, C, D, M, dw1, 0, D);        // dW1 += dH^T Xn
    gw1.accumulate = 1;
    M2_TRY(gemm_f32_simt(gw1, s));
    M2_TRY(colsum_f32(dh, C, M, C, db1, s));
    M2_TRY(colsum_f32(dyb, D, M, D, db2, s));
    GemmArgs gx = gemm_args(dh, 0, C, w1, 1, D, M, D, C, dxn, 0, D);         // dXn = dH W1  (W1 [C][D] as [K][N])
    M2_TRY(gemm_f32_simt(gx, s));
    return ln_bwd(dxn, static_cast<long long>(M) * D, M, u, ln_w, dy, du, dln_w, dln_b, M, D, s);
  }
  if (!w1b || !w2b || D % 8 || ldw2 % 8 || ldw2 < C) return M2_ERR_ARG;
  const size_t mc8 = static_cast<size_t>(M) * c8;
  __nv_bfloat16* xn_b = ws.take<__nv_bfloat16>(md);
  __nv_bfloat16* dy_b = ws.take<__nv_bfloat16>(md);
  const bool gen2 = path == kPathFused && chain_generation() != 1 && chain_fwd_ts_supported(D);
  if (gen2 && chain_generation() == 4) {
    // generation 4: the dgrad chain also spills dH (bf16, TMA stores); the weight-gradient kernel reads it back with TMA and
    // recomputes only G (wgrad_dh): half the epilogue work of generation 2 for 2 M C bytes written and read once
    __nv_bfloat16* dh_sp = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * ((C + 63) / 64 * 64));
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(chain_bwd_ts(u, ln_w, ln_b, w1b, b1, w2b, ldw2, dy, du, dln_w, dln_b, db2, xn_b, dy_b, nullptr, dh_sp, c8, M, D, C,
                        dropout_p, seed, s));
    return wgrad_dh(xn_b, dy_b, dh_sp, c8, w1b, b1, dw1, db1, dw2, M, D, C, dropout_p, seed, s);
  }
  if (gen2 && chain_generation() != 3) {
    // generation 2: dgrad chain with the LayerNorm backward, dln_w / dln_b / db2 fused in (chain_ts.cu); no G / dH spill:
    // the weight-gradient kernel recomputes them (wgrad_fused.cu), so the workspace is the two bf16 [M][D] operand copies
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(chain_bwd_ts(u, ln_w, ln_b, w1b, b1, w2b, ldw2, dy, du, dln_w, dln_b, db2, xn_b, dy_b, nullptr, nullptr, c8, M, D,
                        C, dropout_p, seed, s));
    return wgrad_fused(xn_b, dy_b, w1b, w2b, ldw2, b1, dw1, db1, dw2, M, D, C, dropout_p, seed, s);
  }
  __nv_bfloat16* g_b = ws.take<__nv_bfloat16>(mc8);
  __nv_bfloat16* dh_b = ws.take<__nv_bfloat16>(mc8);
  float* dxn = ws.take<float>(md);
  if (gen2) {
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(chain_bwd_ts(u, ln_w, ln_b, w1b, b1, w2b, ldw2, dy, du, dln_w, dln_b, db2, xn_b, dy_b, g_b, dh_b, c8, M, D, C,
                        dropout_p, seed, s));
  } else if (path == kPathFused) {
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(chain_bwd(u, ln_w, ln_b, w1b, b1, w2b, ldw2, dy, xn_b, dy_b, g_b, dh_b, c8, dxn, M, D, C, 0, dropout_p, seed, s));
  } else {
    // the [M x C] intermediates H and dG stay bf16 (they are rounded to bf16 as G / dH right after anyway): half the bytes
    // of the fp32 round trip, which is what bounds the K = D GEMMs that write them
    __nv_bfloat16* h = ws.take<__nv_bfloat16>(mc8);
    __nv_bfloat16* dg = ws.take<__nv_bfloat16>(mc8);
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(ln_fwd(u, ln_w, ln_b, xn_b, 1, M, D, M, 0, nullptr, nullptr, s));
    if (drop) M2_TRY(mask_scale(dy, D, dy_b, 1, D, M, D, dropout_p, seed, kSiteChannelOut, D, s));
    else M2_TRY(cast_pad_bf16(dy, D, dy_b, D, M, D, s));
    GemmArgs gh = gemm_args(xn_b, 0, D, w1b, 0, D, M, C, D, h, 1, c8);
    gh.bias = b1; gh.bias_mode = 1;
    M2_TRY(gemm_bf16_umma(gh, s));
    GemmArgs gg = gemm_args(dy_b, 0, D, w2b, 1, ldw2, M, C, D, dg, 1, c8);
    M2_TRY(gemm_bf16_umma(gg, s));
    if (c8 != C) {   // pad columns of the bf16 G / dH buffers must be finite (they are never contracted over)
      if (cudaMemsetAsync(g_b, 0, mc8 * 2, s) != cudaSuccess || cudaMemsetAsync(dh_b, 0, mc8 * 2, s) != cudaSuccess)
        return M2_ERR_LAUNCH;
    }
    M2_TRY(gelu_fwd_bwd(h, dg, 1, M, C, c8, g_b, dh_b, c8, 1, dropout_p, seed, kSiteChannelHidden, c8, s));
    GemmArgs gx = gemm_args(dh_b, 0, c8, w1b, 1, D, M, D, C, dxn, 0, D);
    M2_TRY(gemm_bf16_umma(gx, s));
  }
  // token-axis contractions (both operands MN-major), split-K with fp32 atomics into the running gradients
  GemmArgs gw2 = gemm_args(dy_b, 1, D, g_b, 1, c8, D, C, M, dw2, 0, C);
  gw2.splitk = wgrad_splitk(D, C, M);
  if (gw2.splitk == 1) gw2.accumulate = 1;
  M2_TRY(gemm_bf16_umma(gw2, s));
  GemmArgs gw1 = gemm_args(dh_b, 1, c8, xn_b, 1, D, C, D, M, dw1, 0, D);
  gw1.splitk = wgrad_splitk(C, D, M);
  if (gw1.splitk == 1) gw1.accumulate = 1;
  M2_TRY(gemm_bf16_umma(gw1, s));
  M2_TRY(colsum_bf16(dh_b, c8, M, C, db1, s));
  if (gen2) return M2_OK;
  if (drop) M2_TRY(colsum_bf16(dy_b, D, M, D, db2, s));   // the masked gradient only exists as the bf16 operand copy
  else M2_TRY(colsum_f32(dy, D, M, D, db2, s));
  return ln_bwd(dxn, static_cast<long long>(M) * D, M, u, ln_w, dy, du, dln_w, dln_b, M, D, s);
}

// ------------------------------------------------------------------------------------------ LayerNorm
int m2b200_layernorm_fwd(const float* x, const float* w, const float* b, float* out, int B, int N, int D,
                         int64_t out_bstride, void* stream) {
  if (!x || !w || !b || !out || B <= 0) return M2_ERR_ARG;
  return ln_fwd(x, w, b, out, 0, B * N, D, N, out_bstride, nullptr, nullptr, S(stream));
}

int m2b200_layernorm_bwd(const float* dy, int64_t dy_bstride, const float* x, const float* w, const float* dres, float* dx,
                         float* dw, float* db, int B, int N, int D, void* stream) {
  if (!dy || !x || !w || !dx || !dw || !db || B <= 0) return M2_ERR_ARG;
  return ln_bwd(dy, dy_bstride, N, x, w, dres, dx, dw, db, B * N, D, S(stream));
}

// ------------------------------------------------------------------------------------------ linear
size_t m2b200_linear_workspace_bytes(int M, int N, int K, int precision, int backward) {
  if (precision == M2B200_FP32) return 0;
  const size_t m = M, k8 = up8(K), n8 = up8(N);
  return backward ? up256(m * k8 * 2) + up256(m * n8 * 2) : up256(m * k8 * 2);
}

int m2b200_linear_fwd(const float* x, const float* w, const void* w_bf16, int ldwb, const float* bias, int act, float* y,
                      int M, int N, int K, int precision, float dropout_p, uint64_t seed, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (!x || !y || M <= 0 || N <= 0 || K <= 0) return M2_ERR_ARG;
  cudaStream_t s = S(stream);
  if (precision == M2B200_FP32) {
    if (!w) return M2_ERR_ARG;
    GemmArgs g = gemm_args(x, 0, K, w, 0, K, M, N, K, y, 0, N);
    g.bias = bias; g.bias_mode = bias ? 1 : 0; g.act = act;
    g.drop_p = dropout_p; g.drop_seed = seed; g.drop_site = kSiteLinear; g.drop_ld = N;
    return gemm_f32_simt(g, s);
  }
  if (!w_bf16 || ldwb % 8 || ldwb < K) return M2_ERR_ARG;
  const int k8 = up8(K);
  Carver ws(workspace, workspace_bytes);
  __nv_bfloat16* xb = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * k8);
  if (!ws.ok) return M2_ERR_WORKSPACE;
  M2_TRY(cast_pad_bf16(x, K, xb, k8, M, K, s));
  GemmArgs g = gemm_args(xb, 0, k8, w_bf16, 0, ldwb, M, N, K, y, 0, N);
  g.bias = bias; g.bias_mode = bias ? 1 : 0; g.act = act;
  g.drop_p = dropout_p; g.drop_seed = seed; g.drop_site = kSiteLinear; g.drop_ld = N;
  return gemm_bf16_umma(g, s);
}

int m2b200_linear_bwd(float* dy, const float* x, const float* y, const float* w, const void* w_bf16, int ldwb, int act,
                      float* dx, float* dw, float* db, int M, int N, int K, int precision, float dropout_p, uint64_t seed,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!dy || !x || !dw || M <= 0 || N <= 0 || K <= 0) return M2_ERR_ARG;
  if (act == M2B200_ACT_GELU || dropout_p < 0.f || dropout_p >= 1.f) return M2_ERR_ARG;
  cudaStream_t s = S(stream);
  if (dropout_p > 0.f) M2_TRY(mask_scale(dy, N, dy, 0, N, M, N, dropout_p, seed, kSiteLinear, N, s));
  if (act == M2B200_ACT_RELU) {
    if (!y) return M2_ERR_ARG;
    M2_TRY(relu_bwd(dy, y, static_cast<long long>(M) * N, s));
  }
  if (db) M2_TRY(colsum_f32(dy, N, M, N, db, s));
  if (precision == M2B200_FP32) {
    if (!w) return M2_ERR_ARG;
    GemmArgs gw = gemm_args(dy, 1, N, x, 1, K, N, K, M, dw, 0, K);     // dW[N][K] += dY^T X
    gw.accumulate = 1;
    M2_TRY(gemm_f32_simt(gw, s));
    if (dx) {
      GemmArgs gx = gemm_args(dy, 0, N, w, 1, K, M, K, N, dx, 0, K);   // dX = dY W  (W [N][K] as [K'][N'])
      M2_TRY(gemm_f32_simt(gx, s));
    }
    return M2_OK;
  }
  if (!w_bf16 || ldwb % 8 || ldwb < K) return M2_ERR_ARG;
  const int k8 = up8(K), n8 = up8(N);
  Carver ws(workspace, workspace_bytes);
  __nv_bfloat16* xb = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * k8);
  __nv_bfloat16* dyb = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * n8);
  if (!ws.ok) return M2_ERR_WORKSPACE;
  M2_TRY(cast_pad_bf16(x, K, xb, k8, M, K, s));
  M2_TRY(cast_pad_bf16(dy, N, dyb, n8, M, N, s));
  GemmArgs gw = gemm_args(dyb, 1, n8, xb, 1, k8, N, K, M, dw, 0, K);
  gw.splitk = wgrad_splitk(N, K, M);
  if (gw.splitk == 1) gw.accumulate = 1;
  M2_TRY(gemm_bf16_umma(gw, s));
  if (dx) {
    GemmArgs gx = gemm_args(dyb, 0, n8, w_bf16, 1, ldwb, M, K, N, dx, 0, K);
    M2_TRY(gemm_bf16_umma(gx, s));
  }
  return M2_OK;
}

// Patch embedding.  BF16 mode with P % 8 == 0: the GEMMs gather their image-side operand themselves (patch_gemm.cu), no
// im2col buffer exists.  Otherwise (FP32 parity mode, odd patch sizes): gather into the workspace + GEMM, and the backward
// gathers again - the image is the only thing the forward keeps for the backward (there is no input gradient).
static bool patch_fused(const void* img, int img_bf16, int B, int cin, int H, int W, int P, int precision) {
  return precision == M2B200_BF16 && patch_gemm_supported(img, img_bf16, B, cin, H, W, P);
}
size_t m2b200_patch_embed_workspace_bytes(const void* img, int img_bf16, int B, int cin, int H, int W, int P, int D,
                                          int precision, int backward) {
  if (P <= 0 || H % P || W % P) return 0;
  const size_t rows = static_cast<size_t>(B) * (H / P) * (W / P), k = static_cast<size_t>(cin) * P * P;
  const bool fused = patch_fused(img, img_bf16, B, cin, H, W, P, precision);
  size_t n = 0;
  if (!fused) n += up256(precision == M2B200_FP32 ? rows * k * 4 : rows * static_cast<size_t>(up8(static_cast<int>(k))) * 2);
  if (backward && precision != M2B200_FP32) n += up256(rows * up8(D) * 2);   // bf16 dY
  return n;
}

int m2b200_patch_embed_fwd(const void* img, int img_bf16, const float* w, const void* w_bf16, int ldwb, const float* bias,
                           float* y, int B, int cin, int H, int W, int P, int D, int precision, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!img || !y || B <= 0 || cin <= 0 || P <= 0 || D <= 0 || H % P || W % P) return M2_ERR_ARG;
  cudaStream_t s = S(stream);
  const int M = B * (H / P) * (W / P), K = cin * P * P;
  if (patch_fused(img, img_bf16, B, cin, H, W, P, precision)) {
    if (!w_bf16 || ldwb % 8 || ldwb < K) return M2_ERR_ARG;
    return patch_gemm_fwd(img, img_bf16, w_bf16, ldwb, bias, y, B, cin, H, W, P, D, s);
  }
  Carver ws(workspace, workspace_bytes);
  if (precision == M2B200_FP32) {
    if (!w) return M2_ERR_ARG;
    float* cols = ws.take<float>(static_cast<size_t>(M) * K);
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(patch_gather(img, img_bf16, cols, 0, B, cin, H, W, P, K, s));
    GemmArgs g = gemm_args(cols, 0, K, w, 0, K, M, D, K, y, 0, D);
    g.bias = bias; g.bias_mode = bias ? 1 : 0;
    return gemm_f32_simt(g, s);
  }
  const int k8 = up8(K);
  if (!w_bf16 || ldwb % 8 || ldwb < K) return M2_ERR_ARG;
  __nv_bfloat16* cols = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * k8);
  if (!ws.ok) return M2_ERR_WORKSPACE;
  M2_TRY(patch_gather(img, img_bf16, cols, 1, B, cin, H, W, P, k8, s));
  GemmArgs g = gemm_args(cols, 0, k8, w_bf16, 0, ldwb, M, D, K, y, 0, D);
  g.bias = bias; g.bias_mode = bias ? 1 : 0;
  return gemm_bf16_umma(g, s);
}

int m2b200_patch_embed_bwd(const float* dy, const void* img, int img_bf16, float* dw, float* db, int B, int cin, int H, int W,
                           int P, int D, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dy || !img || !dw || B <= 0 || cin <= 0 || P <= 0 || D <= 0 || H % P || W % P) return M2_ERR_ARG;
  cudaStream_t s = S(stream);
  const int M = B * (H / P) * (W / P), K = cin * P * P;
  if (db) M2_TRY(colsum_f32(dy, D, M, D, db, s));
  Carver ws(workspace, workspace_bytes);
  if (precision == M2B200_FP32) {
    float* cols = ws.take<float>(static_cast<size_t>(M) * K);
    if (!ws.ok) return M2_ERR_WORKSPACE;
    M2_TRY(patch_gather(img, img_bf16, cols, 0, B, cin, H, W, P, K, s));
    GemmArgs gw = gemm_args(dy, 1, D, cols, 1, K, D, K, M, dw, 0, K);     // dW[D][K] += dY^T cols
    gw.accumulate = 1;
    return gemm_f32_simt(gw, s);
  }
  const int k8 = up8(K), d8 = up8(D);
  const bool fused = patch_fused(img, img_bf16, B, cin, H, W, P, precision);
  __nv_bfloat16* cols = fused ? nullptr : ws.take<__nv_bfloat16>(static_cast<size_t>(M) * k8);
  __nv_bfloat16* dyb = ws.take<__nv_bfloat16>(static_cast<size_t>(M) * d8);
  if (!ws.ok) return M2_ERR_WORKSPACE;
  M2_TRY(cast_pad_bf16(dy, D, dyb, d8, M, D, s));
  if (fused) return patch_gemm_wgrad(img, img_bf16, dyb, d8, dw, B, cin, H, W, P, D, s);
  M2_TRY(patch_gather(img, img_bf16, cols, 1, B, cin, H, W, P, k8, s));
  GemmArgs gw = gemm_args(dyb, 1, d8, cols, 1, k8, D, K, M, dw, 0, K);
  gw.splitk = wgrad_splitk(D, K, M);
  if (gw.splitk == 1) gw.accumulate = 1;
  return gemm_bf16_umma(gw, s);
}

int m2b200_patch_gather(const float* img, float* cols, int B, int cin, int H, int W, int P, void* stream) {
  if (!img || !cols) return M2_ERR_ARG;
  return patch_gather(img, 0, cols, 0, B, cin, H, W, P, static_cast<long long>(cin) * P * P, S(stream));
}

int m2b200_dropout_mask(float* out, int rows, int cols, int64_t ld, float dropout_p, uint64_t seed, int site, void* stream) {
  return dropout_mask(out, rows, cols, ld, dropout_p, seed, site, S(stream));
}

// ------------------------------------------------------------------------------------------ fusion
int m2b200_copy_tokens(const float* src, int64_t src_bstride, float* dst, int64_t dst_bstride, int B, int64_t per_batch,
                       int accumulate, void* stream) {
  if (!src || !dst) return M2_ERR_ARG;
  return concat_copy(src, src_bstride, dst, dst_bstride, B, per_batch, accumulate, S(stream));
}

int m2b200_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (!a || !b || !out) return M2_ERR_ARG;
  return add_f32(a, b, out, n, S(stream));
}

int m2b200_fuse2_fwd(const float* a, const float* b, float* out, int64_t n, int mode, void* stream) {
  if (!a || !b || !out) return M2_ERR_ARG;
  return fuse2_fwd(a, b, out, n, mode, S(stream));
}

int m2b200_fuse2_max_bwd(const float* a, const float* b, const float* g, float* da, float* db, int64_t n, void* stream) {
  if (!a || !b || !g || !da || !db) return M2_ERR_ARG;
  return fuse2_max_bwd(a, b, g, da, db, n, S(stream));
}

int m2b200_gate_fwd(const float* h1, const float* h2, const float* zh, float* out, int64_t n, void* stream) {
  if (!h1 || !h2 || !zh || !out) return M2_ERR_ARG;
  return gate_fwd(h1, h2, zh, out, n, S(stream));
}

int m2b200_gate_bwd(const float* h1, const float* h2, const float* zh, const float* g, float* dh1, float* dh2, float* dzh,
                    int64_t n, void* stream) {
  if (!h1 || !h2 || !zh || !g || !dh1 || !dh2 || !dzh) return M2_ERR_ARG;
  return gate_bwd(h1, h2, zh, g, dh1, dh2, dzh, n, S(stream));
}

int m2b200_mean_pool_fwd(const float* x, float* out, int B, int N, int D, void* stream) {
  if (!x || !out) return M2_ERR_ARG;
  return mean_pool_fwd(x, out, B, N, D, S(stream));
}
int m2b200_mean_pool_bwd(const float* dpooled, float* dx, int B, int N, int D, void* stream) {
  if (!dpooled || !dx) return M2_ERR_ARG;
  return mean_pool_bwd(dpooled, dx, B, N, D, S(stream));
}

// ------------------------------------------------------------------------------------------ heads + loss
static int fill_heads(HeadsArgs* a, const float* const* tok, const int64_t* tok_bstride, const int* ntok, const int* dim,
                      const float* const* w, const float* const* b, int nheads, int B, int K, int loss_kind,
                      const void* labels, const float* pos_weight, const float* head_weight) {
  if (!tok || !tok_bstride || !ntok || !dim || !w || !b || !head_weight || nheads < 1 || nheads > 3) return M2_ERR_ARG;
  for (int h = 0; h < nheads; ++h) {
    a->tok[h] = tok[h]; a->tok_bstride[h] = tok_bstride[h]; a->ntok[h] = ntok[h]; a->dim[h] = dim[h];
    a->w[h] = w[h]; a->b[h] = b[h]; a->head_weight[h] = head_weight[h];
  }
  a->nheads = nheads; a->B = B; a->K = K; a->loss_kind = loss_kind; a->labels = labels; a->pos_weight = pos_weight;
  return M2_OK;
}

int m2b200_heads_loss_fwd(const float* const* tok, const int64_t* tok_bstride, const int* ntok, const int* dim,
                          const float* const* w, const float* const* b, int nheads, int B, int K, int loss_kind,
                          const void* labels, const float* pos_weight, const float* head_weight, float* logits,
                          float* losses, int64_t* preds, void* stream) {
  HeadsArgs a = {};
  M2_TRY(fill_heads(&a, tok, tok_bstride, ntok, dim, w, b, nheads, B, K, loss_kind, labels, pos_weight, head_weight));
  return heads_loss_fwd(a, logits, losses, reinterpret_cast<long long*>(preds), S(stream));
}

int m2b200_heads_loss_bwd(const float* const* tok, const int64_t* tok_bstride, const int* ntok, const int* dim,
                          const float* const* w, const float* const* b, int nheads, int B, int K, int loss_kind,
                          const void* labels, const float* pos_weight, const float* head_weight, const float* logits,
                          float grad_scale, const float* grad_scale_dev, float* const* dtok, const int64_t* dtok_bstride,
                          const int* accumulate_dtok, float* const* dw, float* const* db, void* stream) {
  HeadsArgs a = {};
  M2_TRY(fill_heads(&a, tok, tok_bstride, ntok, dim, w, b, nheads, B, K, loss_kind, labels, pos_weight, head_weight));
  if (!logits || !dtok || !dtok_bstride || !accumulate_dtok || !dw || !db) return M2_ERR_ARG;
  float* dt[3] = {nullptr, nullptr, nullptr};
  long long dts[3] = {0, 0, 0};
  int acc[3] = {0, 0, 0};
  float* dwp[3] = {nullptr, nullptr, nullptr};
  float* dbp[3] = {nullptr, nullptr, nullptr};
  for (int h = 0; h < nheads; ++h) { dt[h] = dtok[h]; dts[h] = dtok_bstride[h]; acc[h] = accumulate_dtok[h]; dwp[h] = dw[h]; dbp[h] = db[h]; }
  return heads_loss_bwd(a, logits, grad_scale, grad_scale_dev, dt, dts, acc, dwp, dbp, S(stream));
}

// ------------------------------------------------------------------------------------------ optimiser
int m2b200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, float* state_dev, void* stream) {
  return adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, state_dev,
                   S(stream));
}

}  // extern "C"
