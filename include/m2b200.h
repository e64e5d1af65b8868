/* m2b200 - C ABI of the B200-native M2-Mixer training hot path (libm2b200.so).
 *
 * The reference (bezirganyan/m2-mixer) is pure Python with no FFI: its only seam for this path is the name registry
 * modules.get_block_by_name / get_fusion_by_name / get_classifier_by_name (modules/__init__.py:12-26).  The drop-in
 * nn.Modules in m2_mixer_b200/modules bind to these entry points through ctypes (m2_mixer_b200/_lib.py) and register
 * them as torch.library ops (m2_mixer_b200/ops.py).  Each entry point below names the reference code it replaces.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers, explicit sizes, no torch types; fp32 unless a name ends in _bf16 / says bf16.
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it: no allocation, no host sync,
 *    no global mutable state (only one-time cudaFuncSetAttribute).  The caller owns every buffer, including
 *    `workspace` (size from the matching *_workspace_bytes query), and keeps them alive until the stream drains.
 *  - return value: 0 = ok, otherwise an M2B200_ERR_* code (never throws, never aborts).
 *  - precision: M2B200_FP32 = CUDA-core FP32 FMA + exact erf GELU (parity mode, <=1e-4 vs the fp32 reference);
 *               M2B200_BF16 = tcgen05 tensor cores, bf16 operands / fp32 accumulate (<=2e-2 on logits).
 *  - parameter-gradient outputs (dw*, db*, dln_*) are ACCUMULATED INTO (+=): zero them first or point them at a
 *    running gradient buffer.  Activation-gradient outputs (dx, du) are overwritten unless stated otherwise.
 *  - all fp32 pointers must be 16-byte aligned, hidden sizes D multiples of 8 in bf16 mode (4 in fp32 mode).
 */
#ifndef M2B200_H_
#define M2B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M2B200_OK 0
#define M2B200_ERR_ARG 1
#define M2B200_ERR_ALIGN 2
#define M2B200_ERR_WORKSPACE 3
#define M2B200_ERR_LAUNCH 4
#define M2B200_ERR_DRIVER 5

#define M2B200_FP32 0
#define M2B200_BF16 1

#define M2B200_ACT_NONE 0
#define M2B200_ACT_GELU 1
#define M2B200_ACT_RELU 2

int m2b200_abi_version(void);
const char* m2b200_status_string(int status);

/* bf16 operand cache of an fp32 [rows][cols] matrix, row stride ldd >= cols (pad columns zeroed).  The modules keep
 * fp32 master weights in the reference's [out,in] state-dict layout and refresh this cache after every
 * optimiser step. */
int m2b200_cast_bf16(const float* src, int64_t lds, void* dst_bf16, int64_t ldd, int rows, int cols, void* stream);
/* The same for every GEMM weight of a model in ONE launch.  table_dev: n sextuples {src ptr, dst ptr, rows, cols, lds,
 * ldd} of int64 in DEVICE memory (built once by the caller; the pointers are stable because FusedAdam keeps the
 * parameters in one flat buffer).  Replaces the per-matrix refresh after each optimiser step (reference: the weights
 * are nn.Linear / nn.Conv2d parameters, modules/mixer.py:13-19,143-146). */
int m2b200_cast_bf16_multi(const int64_t* table_dev, int n, void* stream);

/* Generic GEMM  C[b] = act(A[b] (x) B[b] + bias) + residual  (see csrc/kernels.h GemmArgs for the full contract).
 * precision FP32: A,B are fp32; BF16: A,B are bf16 (tcgen05 + TMA).  a_mn/b_mn = 1 means the operand is stored
 * [K][M] / [K][N] (contraction axis outermost) - this is how the transposes of modules/mixer.py:32,34 and every
 * weight-gradient contraction are expressed without materialising a permute. */
int m2b200_gemm(int precision, const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N,
                int K, int batch, int64_t a_batch_rows, int64_t b_batch_rows, const float* bias, int bias_mode, int act,
                const float* residual, int64_t ldr, int64_t r_batch_stride, void* C, int c_bf16, int64_t ldc,
                int64_t c_batch_stride, int accumulate, int splitk, void* stream);

/* ---- Dropout (nn.Dropout(p) twice per FeedForward, modules/mixer.py:16,18; once per MLP block, modules/mlp.py:17).
 * Every op that contains a dropout site takes (dropout_p, seed): p = 0 disables it; otherwise the mask is a
 * counter-based hash of (seed, site, element index) evaluated inside the fused epilogues and RE-evaluated by the
 * backward kernels (nothing is stored).  The stream differs from torch's Philox by design (no bit parity);
 * m2b200_dropout_mask exports the exact keep-mask (1/0) of a site for testing: index = r*ld + c with
 *   site 0 token hidden   rows = B*T, ld = D        site 1 token out     rows = B*N, ld = D
 *   site 2 channel hidden rows = M,   ld = up8(C)   site 3 channel out   rows = M,   ld = D      site 4 linear ld = N
 * The drop probability is quantised to t/128, t = round(p*128) (one hash per 4 consecutive elements, 7 bits each);
 * kept values are scaled by 128 / (128 - t), the inverse of the realised keep probability.                            */
int m2b200_dropout_mask(float* out, int rows, int cols, int64_t ld, float dropout_p, uint64_t seed, int site, void* stream);
/* CUDA-graph support for the dropout sites: (dropout_p, seed) are launch parameters, so a captured graph would replay the
 * same masks.  After m2b200_set_dropout_epoch_ptr(p) (p: one uint32 in device memory, NULL switches it off again) every
 * subsequently LAUNCHED kernel folds *p into its mask key at run time; m2b200_dropout_epoch_advance(p, stream) increments
 * it (one tiny launch, capturable).  Process-global setting, read at launch time.                                         */
void m2b200_set_dropout_epoch_ptr(const void* dev_u32);
int m2b200_dropout_epoch_advance(void* dev_u32, void* stream);

/* ---- MixerBlock.token_mix + residual: modules/mixer.py:30-35,43
 *   u[b] = x[b] + Wt2 . GELU(Wt1 . LN(x[b]) + bt1) + bt2,   x,u [B][N][D], wt1 [T][N], wt2 [N][T]
 * Three implementations behind the same call: register-tile mma.sync kernels (N <= 16, T <= 32, bf16), batched tcgen05
 * GEMMs with the sample tile as the MN-major operand (bf16, N * T >= 2048: MM-IMDB-shaped / Scaled configs; exact erf GELU)
 * and CUDA-core kernels (everything else, and the fp32 parity mode).  The workspace queries return 0 where none is needed. */
size_t m2b200_token_mix_fwd_workspace_bytes(int B, int N, int D, int T, int precision);
int m2b200_token_mix_fwd(const float* x, const float* ln_w, const float* ln_b, const float* wt1, const float* bt1,
                         const float* wt2, const float* bt2, float* u, int B, int N, int D, int T, int precision,
                         float dropout_p, uint64_t seed, void* workspace, size_t workspace_bytes, void* stream);
size_t m2b200_token_mix_bwd_workspace_bytes(int B, int N, int D, int T, int precision);
int m2b200_token_mix_bwd(const float* du, const float* x, const float* ln_w, const float* ln_b, const float* wt1,
                         const float* bt1, const float* wt2, float* dx, float* dln_w, float* dln_b, float* dwt1,
                         float* dbt1, float* dwt2, float* dbt2, int B, int N, int D, int T, int precision, float dropout_p,
                         uint64_t seed, void* workspace, size_t workspace_bytes, void* stream);

/* ---- MixerBlock.channel_mix + residual: modules/mixer.py:37-40,45
 *   y = u + W2 . GELU(W1 . LN(u) + b1) + b2,   u,y [M][D] (M = B*N token rows), w1 [C][D], w2 [D][C]
 * BF16 mode reads the bf16 caches w1_bf16 [C][D] and w2_bf16 [D][ldw2] (ldw2 = C rounded up to 8, pad zero).
 * Workspace: the fused forward (D <= 256) needs none; the fused backward (D <= 128) needs the bf16 copies of LN(u) and dY
 * (2 M D bf16) plus the bf16 dH that the dgrad chain spills for the weight-gradient kernel (M * ceil(C / 64) * 64 bf16,
 * written and read once); other shapes / FP32 mode materialise the [M][C] intermediates.  Always ask the query.   */
size_t m2b200_channel_mix_workspace_bytes(int M, int D, int C, int precision, int backward);
int m2b200_channel_mix_fwd(const float* u, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                           const float* w2, const float* b2, const void* w1_bf16, const void* w2_bf16, int ldw2, float* y,
                           int M, int D, int C, int precision, float dropout_p, uint64_t seed, void* workspace,
                           size_t workspace_bytes, void* stream);
int m2b200_channel_mix_bwd(const float* dy, const float* u, const float* ln_w, const float* ln_b, const float* w1,
                           const float* b1, const float* w2, const void* w1_bf16, const void* w2_bf16, int ldw2, float* du,
                           float* dln_w, float* dln_b, float* dw1, float* db1, float* dw2, float* db2, int M, int D, int C,
                           int precision, float dropout_p, uint64_t seed, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ---- nn.LayerNorm(hidden_dim) closing every stack: modules/mixer.py:131,161,185,263.
 * Token row (b,n) is written to out + b*out_bstride + n*D so that an encoder can normalise straight into its slice
 * of the fused-token buffer (zero-copy ConcatFusion, modules/fusion.py:117).  bwd: dx = LN'(dy) (+ dres).           */
int m2b200_layernorm_fwd(const float* x, const float* w, const float* b, float* out, int B, int N, int D,
                         int64_t out_bstride, void* stream);
int m2b200_layernorm_bwd(const float* dy, int64_t dy_bstride, const float* x, const float* w, const float* dres, float* dx,
                         float* dw, float* db, int B, int N, int D, void* stream);

/* ---- Linear layers feeding the stacks: Conv2d patch embedding as a GEMM over gathered patches
 * (modules/mixer.py:143-146), MLPMixerNoPatching.proj (:171), PNLPMixer.bottleneck (:244), MLP (modules/mlp.py).
 *   y[M][N] = act(x[M][K] . w[N][K]^T + bias)                                                                     */
size_t m2b200_linear_workspace_bytes(int M, int N, int K, int precision, int backward);
int m2b200_linear_fwd(const float* x, const float* w, const void* w_bf16, int ldwb, const float* bias, int act, float* y,
                      int M, int N, int K, int precision, float dropout_p, uint64_t seed, void* workspace,
                      size_t workspace_bytes, void* stream);
/* dy is modified in place when act == RELU (masked by y > 0).  dx may be NULL (patch embedding: input needs no grad) */
int m2b200_linear_bwd(float* dy, const float* x, const float* y, const float* w, const void* w_bf16, int ldwb, int act,
                      float* dx, float* dw, float* db, int M, int N, int K, int precision, float dropout_p, uint64_t seed,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Patch embedding, Conv2d(cin, D, k = stride = P) + Rearrange('b c h w -> b (h w) c') (modules/mixer.py:143-146), as a
 * GEMM over the patch rows: y [M][D] in 'b (h w) c' order, M = B*(H/P)*(W/P), K = cin*P*P, w [D][K] = the Conv2d weight
 * [D][cin][P][P] viewed flat.  img is [B][cin][H][W], fp32 (img_bf16 = 0) or bf16 (img_bf16 = 1: BF16 mode only; the bf16
 * GEMM operand is the pixel rounded to bf16 either way, so both give the same bits).  BF16 mode with P % 8 == 0 and a
 * 16-byte aligned image: the tcgen05 GEMMs gather their image-side operand tile from the pixels themselves (no im2col
 * buffer, workspace = the bf16 dY of the backward only).  Other shapes / FP32 mode: gather into the workspace + GEMM.
 * The backward takes the IMAGE again (nothing else is kept from the forward; there is no input gradient), accumulates
 * dw [D][K] and db [D] (db may be NULL).                                                                              */
size_t m2b200_patch_embed_workspace_bytes(const void* img, int img_bf16, int B, int cin, int H, int W, int P, int D,
                                          int precision, int backward);
int m2b200_patch_embed_fwd(const void* img, int img_bf16, const float* w, const void* w_bf16, int ldwb, const float* bias,
                           float* y, int B, int cin, int H, int W, int P, int D, int precision, void* workspace,
                           size_t workspace_bytes, void* stream);
int m2b200_patch_embed_bwd(const float* dy, const void* img, int img_bf16, float* dw, float* db, int B, int cin, int H, int W,
                           int P, int D, int precision, void* workspace, size_t workspace_bytes, void* stream);
/* img [B][cin][H][W] -> cols [B*(H/P)*(W/P)][cin*P*P]  (row = patch in (h w) order, col = (c, py, px))              */
int m2b200_patch_gather(const float* img, float* cols, int B, int cin, int H, int W, int P, void* stream);

/* ---- ConcatFusion / SumFusion: modules/fusion.py:112-146, 207-221
 * concat: dst[b][n_off + n][:] = src[b][n][:]  (accumulate=0)  |  split-add for the backward (swap src/dst strides) */
int m2b200_copy_tokens(const float* src, int64_t src_bstride, float* dst, int64_t dst_bstride, int B, int64_t per_batch,
                       int accumulate, void* stream);
int m2b200_add(const float* a, const float* b, float* out, int64_t n, void* stream);
/* MaxFusion (modules/fusion.py:190-204, torch.maximum) and MeanFusion of two modalities (modules/fusion.py:258-272):
 * mode 1: out = max(a, b), mode 2: out = (a + b) / 2.  Backward of max with torch's tie rule (equal inputs share the
 * gradient evenly): da = g [a > b] + g/2 [a == b], db = g - da.  The backward of mean is g/2 for both (host side).     */
int m2b200_fuse2_fwd(const float* a, const float* b, float* out, int64_t n, int mode, void* stream);
int m2b200_fuse2_max_bwd(const float* a, const float* b, const float* g, float* da, float* db, int64_t n, void* stream);
/* Gate of BiModalGatedUnit.forward (modules/fusion.py:16-23): out = z tanh(h1) + (1 - z) tanh(h2), z = sigmoid(zh); the three
 * linears around it are m2b200_linear_fwd/bwd.  Backward: dh1 = g z (1 - tanh(h1)^2), dh2 = g (1 - z)(1 - tanh(h2)^2),
 * dzh = g (tanh(h1) - tanh(h2)) z (1 - z).                                                                              */
int m2b200_gate_fwd(const float* h1, const float* h2, const float* zh, float* out, int64_t n, void* stream);
int m2b200_gate_bwd(const float* h1, const float* h2, const float* zh, const float* g, float* dh1, float* dh2, float* dzh,
                    int64_t n, void* stream);

/* ---- token mean-pool of a standalone StandardClassifier.forward (modules/classification.py:90):
 *   out[b][d] = mean_n x[b][n][d]; the task modules use the fused heads kernel below instead.                      */
int m2b200_mean_pool_fwd(const float* x, float* out, int B, int N, int D, void* stream);
int m2b200_mean_pool_bwd(const float* dpooled, float* dx, int B, int N, int D, void* stream);

/* ---- heads + multi-head loss: models/avmnist.py:267-312, models/mimic.py:106-142, models/mmimdb.py:106-147,
 * modules/classification.py:84-90.  Up to 3 heads; head h mean-pools tok[h] ([B][ntok][dim], batch stride given)
 * and applies Linear(dim -> K).  loss_kind 0: cross-entropy, int64 labels [B]; 1: BCE-with-logits with pos_weight,
 * float labels [B][K].  loss = sum_h head_weight[h] * L_h  (weights are runtime scalars: the reference changes the
 * fusion weight per epoch, avmnist.py:338-339).
 *   logits [3][B][K], losses [4] = {loss, L_0, L_1, L_2}, preds int64 [3][B] (CE: argmax) or [3][B][K] (BCE: >0)   */
int m2b200_heads_loss_fwd(const float* const* tok, const int64_t* tok_bstride, const int* ntok, const int* dim,
                          const float* const* w, const float* const* b, int nheads, int B, int K, int loss_kind,
                          const void* labels, const float* pos_weight, const float* head_weight, float* logits,
                          float* losses, int64_t* preds, void* stream);
/* dtok[h] may be NULL; accumulate_dtok[h] != 0 adds into dtok[h] (the fused-token gradient slices).  The upstream
 * gradient of the total loss is grad_scale * (grad_scale_dev ? *grad_scale_dev : 1): the device scalar lets autograd
 * hand over d(loss) without a host sync.                                                                            */
int m2b200_heads_loss_bwd(const float* const* tok, const int64_t* tok_bstride, const int* ntok, const int* dim,
                          const float* const* w, const float* const* b, int nheads, int B, int K, int loss_kind,
                          const void* labels, const float* pos_weight, const float* head_weight, const float* logits,
                          float grad_scale, const float* grad_scale_dev, float* const* dtok, const int64_t* dtok_bstride,
                          const int* accumulate_dtok, float* const* dw, float* const* db, void* stream);

/* ---- torch.optim.Adam as configured by the reference (models/avmnist.py:413-415) over one flat buffer.
 * state_dev (optional, device float[2] = {lr, step}) makes lr/step device-resident (graph replay, LR scheduler): the
 * call first advances state_dev[1], unless step < 0 (further ranges of the same optimiser step: torch.optim.Adam skips
 * parameters without a gradient, so a model with frozen encoders is updated range by range).                       */
int m2b200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, float* state_dev, void* stream);

/* ---- launch accounting (bench.py): total kernel launches since load; optional CUDA-event timing per kernel name.
 * m2b200_profile_collect synchronises the recorded events and writes "name,launches,total_ms\n" lines.           */
unsigned long long m2b200_launch_count(void);
void m2b200_profile_enable(int on);
size_t m2b200_profile_collect(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* M2B200_H_ */
