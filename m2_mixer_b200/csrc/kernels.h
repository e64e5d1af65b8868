// Internal (C++) launcher declarations shared between the .cu files and the C-ABI layer (abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace m2 {

// C[b] = epilogue( A[b] (x) B[b] ),  contraction over K.
//   a_mn == 0: A is row-major [M][K] (K-major);   a_mn == 1: A is row-major [K][M] (MN-major)
//   b_mn == 0: B is row-major [N][K] (K-major);   b_mn == 1: B is row-major [K][N] (MN-major)
//   batch b adds a_batch_rows / b_batch_rows ROWS to the operand (0 = operand shared by all batches).
//   epilogue: v = acc; v += bias (bias_mode 1: bias[n], 2: bias[m]); v = act(v) (1: erf GELU, 2: ReLU); v = drop(v);
//             v += residual[b][m][n];  C = accumulate ? C + v : v.
//   splitk > 1 (fp32 C only): K is cut into `splitk` ranges reduced with fp32 atomics INTO C (caller zeroes C
//   or wants accumulation); bias/residual are applied by split 0; act must be 0.
struct GemmArgs {
  const void* A; int a_mn; long long lda;
  const void* B; int b_mn; long long ldb;
  int M, N, K, batch;
  long long a_batch_rows, b_batch_rows;
  const float* bias; int bias_mode; int act;
  const float* residual; long long ldr; long long r_batch_stride;
  void* C; int c_bf16; long long ldc; long long c_batch_stride;
  int accumulate; int splitk;
  int atomic_out;   // fp32 C only: reduce into C with atomics even without split-K (batches that share one C: c_batch_stride = 0)
  // optional dropout applied after `act` and before `residual`; element index = m * drop_ld + n (batch ignored)
  float drop_p; unsigned long long drop_seed; int drop_site; long long drop_ld;
};

int gemm_bf16_umma(const GemmArgs& g, cudaStream_t s);   // A,B bf16; tcgen05 + TMA
int gemm_f32_simt(const GemmArgs& g, cudaStream_t s);    // A,B fp32; CUDA-core FMA (fp32-parity mode)


// ---- fused channel-mix chains (chain.cu)
bool chain_fwd_supported(int D);
bool chain_bwd_supported(int D);
int chain_fwd(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
              int ldw2, const float* b2, float* y, int M, int D, int C, int exact_gelu, float drop_p, unsigned long long seed,
              cudaStream_t s);
int chain_bwd(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
              int ldw2, const float* dy, void* xn_b, void* dy_b, void* g_b, void* dh_b, int ldh, float* dxn, int M, int D,
              int C, int exact_gelu, float drop_p, unsigned long long seed, cudaStream_t s);

// generation 2 (chain_ts.cu): row-tile operands in tensor memory, D <= 128
bool chain_fwd_ts_supported(int D);
int chain_fwd_ts(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
                 int ldw2, const float* b2, float* y, int M, int D, int C, float drop_p, unsigned long long seed,
                 cudaStream_t s);
int chain_bwd_ts(const float* u, const float* ln_w, const float* ln_b, const void* w1b, const float* b1, const void* w2b,
                 int ldw2, const float* dy, float* du, float* dln_w, float* dln_b, float* db2, void* xn_b, void* dy_b,
                 void* g_b, void* dh_b, int ldh, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s);
// fused recompute weight gradients (wgrad_fused.cu): dw1 / db1 / dw2 accumulate; xn_b / dy_b from chain_bwd_ts
int wgrad_fused(const void* xn_b, const void* dy_b, const void* w1b, const void* w2b, int ldw2, const float* b1, float* dw1,
                float* db1, float* dw2, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s);
// generation 4 (dH spilled by chain_bwd_ts through TMA stores, only G recomputed): dw1 / db1 / dw2 accumulate
int wgrad_dh(const void* xn_b, const void* dy_b, const void* dh_b, int ldh, const void* w1b, const float* b1, float* dw1,
             float* db1, float* dw2, int M, int D, int C, float drop_p, unsigned long long seed, cudaStream_t s);
// env M2B200_CHAIN_GEN (A/B measurements): 1 = generation-1 kernels, 2 = generation-2 dgrad + weight gradients that recompute
// G and dH (wgrad_fused), 3 = generation-2 dgrad + G/dH spill with per-thread stores + GEMM weight gradients,
// default 4 = generation-2 dgrad + dH spill through TMA stores + weight gradients that recompute only G (wgrad_dh).
int chain_generation();
// env M2B200_DH_L2HINT (A/B): 0 = the spilled dH moves without an L2 eviction policy; default 1 = evict_first
int dh_l2_hint();

// ---- row kernels (rowops.cu)
int cast_pad_bf16(const float* src, long long lds, void* dst, long long ldd, int rows, int cols, cudaStream_t s);
// table_dev: n x {src ptr, dst ptr, rows, cols, lds, ldd} as int64 in device memory
int cast_pad_bf16_multi(const long long* table_dev, int n, cudaStream_t s);
int ln_fwd(const float* x, const float* w, const float* b, void* out, int out_bf16, int rows, int D, int N,
           long long out_bstride, float* mean, float* rstd, cudaStream_t s);
int ln_bwd(const float* dy, long long dy_bstride, int N, const float* x, const float* w, const float* dres, float* dx,
           float* dw, float* db, int rows, int D, cudaStream_t s);
int colsum_f32(const float* src, long long ld, int rows, int cols, float* out, cudaStream_t s);
int colsum_bf16(const void* src, long long ld, int rows, int cols, float* out, cudaStream_t s);
// out[r % period] += sum_c src[r][c]   (bf16 rows; bias gradients of the GEMM-composed token mixing)
int rowsum_mod_bf16(const void* src, long long ld, int rows, int cols, int period, float* out, cudaStream_t s);
// h / dg: fp32, or bf16 (in_bf16 = 1, bf16 outputs only) with leading dimension ld_in
int gelu_fwd_bwd(const void* h, const void* dg, int in_bf16, int rows, int cols, long long ld_in, void* g_out, void* dh_out,
                 long long ld_out, int out_bf16, float drop_p, unsigned long long seed, int site, long long drop_ld,
                 cudaStream_t s);
// dst = src * mask * scale (dst fp32 or bf16; dst may alias src when fp32); index = r * drop_ld + c
int mask_scale(const float* src, long long lds, void* dst, int dst_bf16, long long ldd, int rows, int cols, float drop_p,
               unsigned long long seed, int site, long long drop_ld, cudaStream_t s);
// out[r][c] = 1/0 keep mask of (seed, site) at index r*ld + c  (test / debugging export of the kernels' mask function)
int dropout_mask(float* out, int rows, int cols, long long ld, float drop_p, unsigned long long seed, int site, cudaStream_t s);
int patch_gather(const void* img, int img_bf16, void* cols, int out_bf16, int B, int cin, int H, int W, int P, long long ld,
                 cudaStream_t s);
// patch embedding with the image-side operand gathered inside the GEMM (patch_gemm.cu): bf16 mode, P % 8 == 0
bool patch_gemm_supported(const void* img, int img_bf16, int B, int cin, int H, int W, int P);
int patch_gemm_fwd(const void* img, int img_bf16, const void* w_bf16, int ldwb, const float* bias, float* y, int B, int cin,
                   int H, int W, int P, int D, cudaStream_t s);
int patch_gemm_wgrad(const void* img, int img_bf16, const void* dy_b, int ldd, float* dw, int B, int cin, int H, int W, int P,
                     int D, cudaStream_t s);
int concat_copy(const float* src, long long src_bstride, float* dst, long long dst_bstride, int B, long long per_batch,
                int accumulate, cudaStream_t s);
int add_f32(const float* a, const float* b, float* o, long long n, cudaStream_t s);
int gate_fwd(const float* h1, const float* h2, const float* zh, float* o, long long n, cudaStream_t s);
int gate_bwd(const float* h1, const float* h2, const float* zh, const float* g, float* dh1, float* dh2, float* dzh, long long n,
             cudaStream_t s);
int fuse2_fwd(const float* a, const float* b, float* o, long long n, int mode, cudaStream_t s);
int fuse2_max_bwd(const float* a, const float* b, const float* g, float* da, float* db, long long n, cudaStream_t s);
int relu_bwd(float* dy, const float* y, long long n, cudaStream_t s);
int mean_pool_fwd(const float* x, float* out, int B, int N, int D, cudaStream_t s);
int mean_pool_bwd(const float* dp, float* dx, int B, int N, int D, cudaStream_t s);

// ---- token mixing (token_mix.cu)
int token_mix_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
                  const float* b2, float* u, int B, int N, int D, int T, int exact_gelu, float drop_p,
                  unsigned long long seed, cudaStream_t s);
int token_mix_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                  const float* w2, float* dxn, float* dw1, float* db1, float* dw2, float* db2,
                  int B, int N, int D, int T, int exact_gelu, float drop_p, unsigned long long seed, cudaStream_t s);

// warp-level tensor-core path (token_mix_mma.cu): bf16 mode, N <= 16, T <= 32, D in {32,64,128,256}
bool token_mix_mma_supported(int N, int D, int T);
int token_mix_mma_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* u, int B, int N, int D, int T, float drop_p, unsigned long long seed,
                      cudaStream_t s);
int token_mix_mma_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                      const float* w2, float* dx, float* dln_w, float* dln_b, float* dw1, float* db1, float* dw2, float* db2,
                      int B, int N, int D, int T, float drop_p, unsigned long long seed, cudaStream_t s);
int token_generation();   // env M2B200_TOKEN_GEN: 1 = CUDA-core kernels only (A/B measurements)

// ---- heads + multi-head loss (heads.cu)
struct HeadsArgs {
  const float* tok[3]; long long tok_bstride[3]; int ntok[3]; int dim[3];   // pooled inputs per head
  const float* w[3]; const float* b[3];                                      // [K][dim], [K]
  int nheads, B, K;
  int loss_kind;                 // 0 = cross-entropy (int64 labels [B]), 1 = BCE-with-logits (float labels [B][K])
  const void* labels; const float* pos_weight;
  float head_weight[3];          // loss = sum_h head_weight[h] * L_h
};
int heads_loss_fwd(const HeadsArgs& a, float* logits /*[3][B][K]*/, float* losses /*[4]: total, L0, L1, L2*/,
                   long long* preds /*[3][B] (CE) or [3][B][K] (BCE)*/, cudaStream_t s);
int heads_loss_bwd(const HeadsArgs& a, const float* logits, float grad_scale, const float* grad_scale_dev, float* dtok[3], long long dtok_bstride[3],
                   int accumulate_dtok[3], float* dw[3], float* db[3], cudaStream_t s);

// ---- optimiser (adam.cu)
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, float grad_scale, float* state_dev, cudaStream_t s);

}  // namespace m2
