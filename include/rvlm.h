/* rvlm.h — C ABI of librvlm_b200.so
 *
 * The drop-in boundary of the B200-native adversarial-perturbation hot path
 * (PGD / APGD around the CLIP ViT vision encoder).  The reference
 * (chs20/RobustVLM) is pure Python and has no FFI of its own:
 * its "operator API" for this path is the Python call surface
 *   pgd()                    train/pgd_train.py:5-68
 *   apgd_train()             train/apgd_train.py:125-373
 *   APGDAttack               autoattack/autopgd_base.py:89-582
 *   ClipVisionModel          train/adversarial_training_clip.py:246-257
 *   ComputeLossWrapper/l2/ce train/adversarial_training_clip.py:260-274,495-528
 * kept, signature for signature, by the Python package `robustvlm_b200`.  This
 * header is what those shims bind (via ctypes) — and what a maintainer of the
 * reference would bind from `train/pgd_train.py` etc..
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - functions return 0 on success, a negative RVLM_ERR_* code otherwise, or
 *     (-1000 - cudaError_t) for CUDA runtime failures; no exceptions, no hidden
 *     allocation (workspaces are sized by query calls and owned by the caller);
 *   - all calls are asynchronous on `stream` unless stated otherwise;
 *   - fmt: 0 = fp16, 1 = bf16 (16-bit operand / activation format).
 *   - the per-function cudaFuncSetAttribute opt-ins are cached per device; a handle must be
 *     used with the device it was created on current.
 */
#ifndef RVLM_H_
#define RVLM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVLM_ABI_VERSION 2

#define RVLM_ERR_INVAL (-22)
#define RVLM_ERR_ALIGN (-23)
#define RVLM_ERR_DRIVER (-24)
#define RVLM_ERR_TMAP (-25)
#define RVLM_ERR_NOMEM (-12)
#define RVLM_ERR_UNSUPPORTED (-38)

#define RVLM_FMT_F16 0
#define RVLM_FMT_BF16 1
/* forward operands as fp16 hi + lo pairs, three tcgen05 MMAs per product (hi*hi + lo*hi + hi*lo):
 * the fp32-class forward for FARE's cancellation site (train/adversarial_training_clip.py:509-521,
 * SURVEY.md Appendix E).  Valid for rvlm_vit_config.fmt_fwd only. */
#define RVLM_FMT_F16X2 2

int rvlm_abi_version(void);
const char* rvlm_strerror(int code);

/* Accounting for the benchmark: number of kernels this library has launched, and an
 * optional mode that brackets every tcgen05 GEMM launch (enable = 1) or every GEMM, attention and
 * LayerNorm launch (enable = 2) with CUDA events on its stream (collect() synchronises and returns
 * the summed device time and algorithmic FLOPs of the GEMM launches). */
long long rvlm_launch_count(void);
void rvlm_launch_count_reset(void);
int rvlm_profile_gemm(int enable);
int rvlm_profile_gemm_collect(double* total_ms, double* total_flops, long long* launches);
/* Per-kernel breakdown of the last collect(): the same mode also brackets the attention and
 * LayerNorm launches; writes a JSON object {"<kernel>": {"ms", "launches", "gflop"}, ...} into
 * `out` (truncated to `cap`) and returns the size needed. */
int rvlm_profile_breakdown(char* out, int cap);

/* ------------------------------------------------------------------------ *
 * 1. tcgen05 GEMM (the contraction under every nn.Linear / Conv2d-as-GEMM of
 *    the open_clip VisionTransformer the reference wraps,
 *    train/adversarial_training_clip.py:253-257).
 *
 *    C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ resid[M,N])
 *    A, B: 16-bit (fmt), row-major, pitches lda/ldb in elements (16-byte
 *    aligned rows).  N must be a multiple of 32.
 *    flags: bit0 = B is given as [K,N] row-major (MN-major operand)
 *           bit1 = A is given as [K,M] row-major (MN-major operand)
 *           bit2 = 128-column tiles instead of 256
 *           bit3 = 2-CTA kernel (256 x 256 tile per SM pair)
 * ------------------------------------------------------------------------ */
#define RVLM_GEMM_B_MN 1
#define RVLM_GEMM_A_MN 2
#define RVLM_GEMM_BN128 4
#define RVLM_GEMM_2CTA 8 /* cluster of two CTAs, tcgen05.mma.cta_group::2 (K-major operands) */

/* 16-bit output */
int rvlm_gemm16_out16(const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                      const float* bias, int M, int N, int K, int fmt, int flags, void* stream);
/* fp32 output with optional bias and residual */
int rvlm_gemm16_out32(const void* A, int lda, const void* B, int ldb, float* C, int ldc,
                      const float* bias, const float* resid, int ldr, int M, int N, int K, int fmt,
                      int flags, void* stream);
/* Split operands: A = A_hi + A_lo, B = B_hi + B_lo (fp16 pairs, lo = fp16(x - hi)); the product
 * is formed as hi*hi + lo*hi + hi*lo by three tcgen05 MMAs into one fp32 TMEM accumulator
 * (2-CTA kernel, K-major operands; the forward GEMM of RVLM_FMT_F16X2). */
int rvlm_gemm16_split_out32(const void* A_hi, const void* A_lo, int lda, const void* B_hi,
                            const void* B_lo, int ldb, float* C, int ldc, const float* bias,
                            const float* resid, int ldr, int M, int N, int K, void* stream);
/* h = A*B^T + bias; Act = QuickGELU(h), G = QuickGELU'(h) - 1/2 (may be NULL; what the dgrad
 * below multiplies by, offset so that the 16-bit value stays within +-0.6 -- the pre-activation
 * itself is never stored); both 16-bit [M,N] pitch ld */
int rvlm_gemm16_gelu(const void* A, int lda, const void* B, int ldb, void* G, void* Act, int ld,
                     const float* bias, int M, int N, int K, int fmt, void* stream);
/* out = (A*B^T) .* (G + 1/2) ; 16-bit [M,N] pitch ld; G (as written by rvlm_gemm16_gelu) in
 * fmt_g, A/B/out in fmt */
int rvlm_gemm16_gelu_bwd(const void* A, int lda, const void* B, int ldb, void* out, const void* G,
                         int ld, int M, int N, int K, int fmt, int fmt_g, void* stream);

/* ------------------------------------------------------------------------ *
 * 2. Perturbation updates.  Tensors are contiguous fp32 [B, n] (n % 4 == 0).
 *    `workspace` (ell_2 variants): rvlm_update_workspace_bytes(B) bytes.
 * ------------------------------------------------------------------------ */
size_t rvlm_update_workspace_bytes(int B);

/* One PGD iteration's update given the loss gradient (train/pgd_train.py:37-59 with
 * normalize_grad / project_perturbation of vlm_eval/attacks/utils.py:8-26):
 * NaN scrub, sign, momentum, +-stepsize, clamp(+-eps), delta = clamp(x+delta,0,1)-x.
 * In place on delta and velocity.  Bit-exact w.r.t. the reference's fp32 sequence. */
int rvlm_pgd_update_linf(float* delta, float* velocity, const float* grad, const float* x, int B,
                         long long n, float stepsize, float eps, int maximize, float momentum,
                         void* stream);
/* ell_2 variant (F.normalize with eps 1e-12, torch.renorm with maxnorm eps). */
int rvlm_pgd_update_l2(float* delta, float* velocity, const float* grad, const float* x, int B,
                       long long n, float stepsize, float eps, int maximize, float momentum,
                       float* workspace, void* stream);

/* The same two updates for the VLM-evaluation attack (`PGD.perturb`, vlm_eval/attacks/pgd.py:21-73):
 * `mask` (nullable, [B, n]) multiplies the gradient before the NaN scrub (:39, built by
 * `_set_mask` :75-87); image_space = 0 skips delta = clamp(x+delta,0,1)-x (:57-59). */
int rvlm_pgd_update_linf_masked(float* delta, float* velocity, const float* grad, const float* x,
                                const float* mask, int image_space, int B, long long n,
                                float stepsize, float eps, int maximize, float momentum,
                                void* stream);
int rvlm_pgd_update_l2_masked(float* delta, float* velocity, const float* grad, const float* x,
                              const float* mask, int image_space, int B, long long n,
                              float stepsize, float eps, int maximize, float momentum,
                              float* workspace, void* stream);

/* One APGD proposal step (train/apgd_train.py:205-254 == autoattack/autopgd_base.py:328-350).
 * In place: x_adv_old <- x_adv; x_adv <- new iterate.  step_size is per sample [B];
 * a = 1.0 on the first iteration, 0.75 afterwards. */
int rvlm_apgd_step_linf(float* x_adv, float* x_adv_old, const float* x, const float* grad,
                        const float* step_size, int B, long long n, float eps, float a,
                        void* stream);
int rvlm_apgd_step_l2(float* x_adv, float* x_adv_old, const float* x, const float* grad,
                      const float* step_size, int B, long long n, float eps, float a,
                      float* workspace, void* stream);

/* APGD bookkeeping after the model evaluation of iteration `iter`
 * (train/apgd_train.py:301-355 == autoattack/autopgd_base.py:388-446), all on device:
 *   acc &= pred; x_best_adv[!pred] = x_adv; best-loss tracking (x_best, grad_best,
 *   loss_best, loss_steps[iter]); and, when k_check > 0 (a check-point of the
 *   data-independent schedule), the oscillation / no-improvement test with
 *   k_thr = (float)(k_check * rho): halve step_size and restart from x_best/grad_best.
 * flags[B] is scratch. */
int rvlm_apgd_bookkeep(float* x_adv, float* grad, float* x_best, float* grad_best,
                       float* x_best_adv, const float* loss_indiv, const int* pred,
                       float* loss_steps, float* loss_best, float* loss_best_last,
                       float* reduced_last, float* step_size, int* acc, int* flags, int B,
                       long long n, int n_iter, int iter, int k_check, float k_thr, void* stream);

/* out = a * b elementwise (grad *= mask, vlm_eval/attacks/apgd.py:225,297); out may alias a */
int rvlm_mul_f32(float* out, const float* a, const float* b, long long n, void* stream);
/* out = a + b (x_adv = x + delta, train/pgd_train.py:68) */
int rvlm_add_f32(float* out, const float* a, const float* b, long long n, void* stream);

/* ------------------------------------------------------------------------ *
 * 3. ViT building blocks (exported for unit tests and for callers that compose
 *    their own pipelines).  Residual stream fp32 [M, W]; 16-bit operands in fmt.
 * ------------------------------------------------------------------------ */
/* y = LayerNorm(x) (eps 1e-5, affine); y16 and/or y32 may be NULL; stats [M,2] = (mean, rstd) */
int rvlm_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y16,
                       float* y32, float* stats, int M, int W, int fmt, void* stream);
/* dx = dres + LayerNorm'(dy); dres/dx32/dx16 may be NULL; dx32 may alias dres.
 * drop_T > 0: rows with row % drop_T == 0 are skipped and the rest compacted. */
int rvlm_layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma,
                       const float* dres, float* dx32, void* dx16, int M, int W, int fmt,
                       int drop_T, void* stream);
/* Same with dy given as a 16-bit matrix in `fmt` (also the format of dx16). */
int rvlm_layernorm_bwd16(const void* dy16, const float* x, const float* stats, const float* gamma,
                         const float* dres, float* dx32, void* dx16, int M, int W, int fmt,
                         int drop_T, void* stream);
/* Multi-head attention core, head dim 64, no mask: softmax(q k^T / 8) v.
 *   qkv16 [B*T, 3W] (q | k | v per token; head h at columns 64h..64h+63), W = 64 H
 *   o16 [B*T, W]; lse [B, H, T] (log-sum-exp of the scaled scores)
 * tcgen05 kernels (production) and CUDA-core kernels (on-device checker).  The T = 256 / 257
 * kernels write rows 256 bits at a time: o16 (forward) and dqkv16 (rvlm_attn_bwd_lse) must be
 * 32-byte aligned (RVLM_ERR_ALIGN otherwise). */
int rvlm_attn_fwd(const void* qkv16, void* o16, float* lse, int B, int T, int H, int fmt,
                  void* stream);
/* Backward: qkv16 is the backward's copy of q|k|v in fmt_qkv, do16 the output gradient and
 * dqkv16 the result in fmt_g, o16 the forward output in fmt_o.  The softmax statistics
 * are recomputed from the operands given (lse_tmp [B,H,T] is scratch).  The tcgen05
 * kernels need fmt_qkv == fmt_g (one operand format per MMA). */
int rvlm_attn_bwd(const void* qkv16, const void* o16, const void* do16, float* lse_tmp,
                  void* dqkv16, int B, int T, int H, int fmt_qkv, int fmt_o, int fmt_g,
                  void* stream);
/* The same given the forward's log-sum-exp (rvlm_attn_fwd's `lse`): valid when qkv16 holds the very
 * bits the forward multiplied (same 16-bit format, forward not split), so P = exp(S - lse) is the
 * forward's softmax exactly and nothing is recomputed (one TMEM sweep per tile less; the fused
 * kernel for T in {256, 257}; other T fall back to rvlm_attn_bwd with `lse` as scratch). */
int rvlm_attn_bwd_lse(const void* qkv16, const void* o16, const void* do16, float* lse,
                      void* dqkv16, int B, int T, int H, int fmt_qkv, int fmt_o, int fmt_g,
                      void* stream);
/* Split-operand forward (the attention part of RVLM_FMT_F16X2): q | k | v and o as fp16 hi + lo
 * pairs in separate buffers of the same layout; softmax(q k^T / 8) v with every product formed as
 * hi*hi + lo*hi + hi*lo.  o_lo and lse may be NULL.  T <= 257. */
int rvlm_attn_fwd_split(const void* qkv_hi, const void* qkv_lo, void* o_hi, void* o_lo, float* lse,
                        int B, int T, int H, void* stream);
/* Class-token row only (the last block of the tower needs nothing else):
 * forward writes o16[b*T, :] for every image b; backward takes the class-token output
 * gradient do_cls16 [B, W] and writes the dense dqkv16 [B*T, 3W] (dQ rows t > 0 are zero).
 * One 16-bit format for all operands of a call. */
int rvlm_attn_cls_fwd(const void* qkv16, void* o16, int B, int T, int H, int fmt, void* stream);
int rvlm_attn_cls_bwd(const void* qkv16, const void* do_cls16, void* dqkv16, int B, int T, int H,
                      int fmt, void* stream);
int rvlm_attn_fwd_simt(const void* qkv16, void* o16, float* lse, int B, int T, int H, int fmt,
                       void* stream);
int rvlm_attn_bwd_simt(const void* qkv16, const void* o16, const void* do16, float* lse_tmp,
                       void* dqkv16, int B, int T, int H, int fmt_qkv, int fmt_o, int fmt_g,
                       void* stream);

/* ------------------------------------------------------------------------ *
 * 4. The fused engine.  Replaces, for a CLIP vision tower in open_clip
 *    `VisionTransformer` layout (state-dict keys in the field names below):
 *      ClipVisionModel.forward            train/adversarial_training_clip.py:253-257
 *      loss + torch.autograd.grad(loss,x) train/pgd_train.py:32-38, train/apgd_train.py:181-185
 *      the whole pgd() loop               train/pgd_train.py:29-68
 * ------------------------------------------------------------------------ */
typedef void* rvlm_vit_handle;

#define RVLM_VIT_ATTN_SIMT 1 /* flags: use the CUDA-core attention (debug / checker) */
#define RVLM_VIT_FULL_LAST_BLOCK 2 /* flags: evaluate all T rows of the last block (default:
                                      only the class-token row, the only one the output uses) */

typedef struct {
  int image, patch, width, layers, heads, mlp, out_dim;
  float mean[3], std[3]; /* Normalize() inside the model (:116,254) */
  int fmt_fwd;           /* format of forward operands: RVLM_FMT_F16 (fast), RVLM_FMT_F16X2 (fp32-class)
                            or RVLM_FMT_BF16 (the outer training step) */
  int fmt_bwd;           /* 16-bit format of gradient operands (RVLM_FMT_BF16 recommended) */
  int flags;
} rvlm_vit_config;

/* Device fp32 pointers, open_clip VisionTransformer.state_dict() layout. Per-layer
 * entries are host arrays of `layers` device pointers. */
typedef struct {
  const float* conv1_weight;         /* [W, 3, p, p] */
  const float* class_embedding;      /* [W] */
  const float* positional_embedding; /* [T, W] */
  const float* ln_pre_weight;
  const float* ln_pre_bias;
  const float* ln_post_weight;
  const float* ln_post_bias;
  const float* proj;                         /* [W, out_dim] */
  const float* const* ln_1_weight;           /* transformer.resblocks.{i}.ln_1.weight */
  const float* const* ln_1_bias;
  const float* const* attn_in_proj_weight;   /* [3W, W] */
  const float* const* attn_in_proj_bias;     /* [3W] */
  const float* const* attn_out_proj_weight;  /* [W, W] */
  const float* const* attn_out_proj_bias;
  const float* const* ln_2_weight;
  const float* const* ln_2_bias;
  const float* const* mlp_c_fc_weight;       /* [mlp, W] */
  const float* const* mlp_c_fc_bias;
  const float* const* mlp_c_proj_weight;     /* [W, mlp] */
  const float* const* mlp_c_proj_bias;
} rvlm_vit_weights;

/* The attack loss on the embedding (ComputeLossWrapper / compute_loss,
 * train/adversarial_training_clip.py:260-274,495-528) or on zero-shot logits
 * (ClassificationModel, CLIP_eval/clip_robustbench.py:50-69 + CE,
 * autoattack/autopgd_base.py:244). */
#define RVLM_LOSS_L2 1 /* sum_d (emb - emb_orig)^2            (FARE)  */
#define RVLM_LOSS_CE 2 /* CE(emb @ (logit_scale * text), y)   (TeCoA) */
/* AutoAttack's APGD losses on the zero-shot logits z = emb @ (logit_scale * text), z_(k) the k-th
 * largest (rvlm_vit_loss_grad only; n_classes >= 4, pred on the logits):
 *   DLR          -(z_y - (argmax == y ? z_(2) : z_(1))) / (z_(1) - z_(3) + 1e-12)   autopgd_base.py:195-201
 *   targeted DLR -(z_y - z_t) / (z_(1) - (z_(3) + z_(4)) / 2 + 1e-12)               autopgd_base.py:616-621 */
#define RVLM_LOSS_DLR 4
#define RVLM_LOSS_DLR_TARGETED 5
typedef struct {
  int mode;
  int output_normalize;     /* F.normalize(emb, dim=-1) before the loss */
  int pred_on_logits;       /* pred = argmax(class logits)==y (1) or argmax(emb)==y (0) */
  const float* emb_orig;    /* [B, out_dim]        (RVLM_LOSS_L2) */
  const float* text;        /* [out_dim, n_classes] (RVLM_LOSS_CE) */
  int n_classes;
  float logit_scale;
  const long long* targets; /* [B] int64, may be NULL for RVLM_LOSS_L2 */
  const long long* y_target; /* [B] int64 target classes (RVLM_LOSS_DLR_TARGETED), else NULL */
} rvlm_loss_config;

#define RVLM_NORM_LINF 0
#define RVLM_NORM_L2 2

/* Converts the fp32 weights into the engine's own 16-bit operand copies (allocates
 * device memory once, here; no allocation in any later call). */
int rvlm_vit_create(const rvlm_vit_config* cfg, const rvlm_vit_weights* weights, void* stream,
                    rvlm_vit_handle* out);
/* Re-read the weights (after an optimiser step of the outer training loop). */
int rvlm_vit_update_weights(rvlm_vit_handle h, const rvlm_vit_weights* weights, void* stream);
int rvlm_vit_destroy(rvlm_vit_handle h);
/* Caller-owned scratch for a batch of B images; with_backward = 1 also holds the
 * per-layer activations the input-gradient backward needs, 2 those of the outer
 * training step as well. */
size_t rvlm_vit_workspace_bytes(rvlm_vit_handle h, int B, int with_backward);

/* emb_out[B, out_dim] = model(x + delta) (delta may be NULL); x, delta: [B,3,H,W] in [0,1].
 * keep_for_backward = 1 leaves the activations in the workspace (which must then be sized
 * with_backward = 1) for a following rvlm_vit_backward_input on the same workspace. */
int rvlm_vit_forward(rvlm_vit_handle h, void* workspace, size_t workspace_bytes, const float* x,
                     const float* delta, int B, int output_normalize, int keep_for_backward,
                     float* emb_out, void* stream);
/* grad_out[B,3,H,W] = (d emb / d x)^T d_emb for the forward just run with
 * keep_for_backward = 1 (the vector-Jacobian product torch.autograd.grad(loss, x) needs for
 * an arbitrary loss on the embedding; d_emb: [B, out_dim] w.r.t. the returned, possibly
 * normalised, embedding).  The result carries the true scale also when fmt_bwd is fp16 (the
 * internal per-sample rescaling is undone on output). */
int rvlm_vit_backward_input(rvlm_vit_handle h, void* workspace, size_t workspace_bytes, int B,
                            const float* d_emb, int output_normalize, float* grad_out,
                            void* stream);
/* Per-sample loss, pred flag, embedding / logits, and d(sum_b loss_b)/d(x+delta) as
 * grad_out[B,3,H,W] (true scale; with fp16 gradient operands each sample's gradient is carried
 * at a fixed max-abs internally and rescaled on output).  Outputs other than grad_out may be NULL. */
int rvlm_vit_loss_grad(rvlm_vit_handle h, void* workspace, size_t workspace_bytes, const float* x,
                       const float* delta, int B, const rvlm_loss_config* loss, float* grad_out,
                       float* loss_out, int* pred_out, float* emb_out, float* logits_out,
                       int want_grad, void* stream);
/* The whole pgd() call: x_adv_out = x + delta_final.  delta0 = initial perturbation
 * (may be NULL = zeros).  loss_trace (may be NULL): [iterations, B] per-sample loss at
 * each iteration's forward.  No host synchronisation inside. */
int rvlm_pgd_run(rvlm_vit_handle h, void* workspace, size_t workspace_bytes, const float* x,
                 const float* delta0, int B, const rvlm_loss_config* loss, int norm, float eps,
                 float stepsize, int iterations, int maximize, float momentum, float* x_adv_out,
                 float* loss_trace, void* stream);
/* ------------------------------------------------------------------------ *
 * 5. Outer training step (the model update of FARE / TeCoA,
 *    train/adversarial_training_clip.py:340-366): forward of x (the adversarial batch),
 *    loss, full backward.  Writes d(loss_scale * sum_b loss_b)/d(parameter) for every
 *    parameter into `grads` (same layout as rvlm_vit_weights, fp32, overwritten); pass
 *    loss_scale = 1 / global_batch for the reference's batch-mean loss.  The engine must
 *    have fmt_fwd == fmt_bwd (weight-gradient GEMMs need one operand format; bf16
 *    recommended) and the workspace must be sized with with_backward = 2.
 * ------------------------------------------------------------------------ */
typedef struct {
  float* conv1_weight;
  float* class_embedding;
  float* positional_embedding;
  float* ln_pre_weight;
  float* ln_pre_bias;
  float* ln_post_weight;
  float* ln_post_bias;
  float* proj;
  float* const* ln_1_weight;
  float* const* ln_1_bias;
  float* const* attn_in_proj_weight;
  float* const* attn_in_proj_bias;
  float* const* attn_out_proj_weight;
  float* const* attn_out_proj_bias;
  float* const* ln_2_weight;
  float* const* ln_2_bias;
  float* const* mlp_c_fc_weight;
  float* const* mlp_c_fc_bias;
  float* const* mlp_c_proj_weight;
  float* const* mlp_c_proj_bias;
} rvlm_vit_grads;

int rvlm_vit_train_grads(rvlm_vit_handle h, void* workspace, size_t workspace_bytes,
                         const float* x, int B, const rvlm_loss_config* loss, float loss_scale,
                         const rvlm_vit_grads* grads, float* loss_out, float* emb_out,
                         void* stream);
/* The same with completion events for gradient bucketing: events = cudaEvent_t[layers + 1] (as
 * void*), events[l] recorded on `stream` once every parameter gradient of block l is enqueued
 * (blocks finish in the order layers-1 .. 0), events[layers] at the end (stem + head).  The
 * caller starts the NCCL all-reduce of a block's gradients on another stream behind its event
 * while the rest of the backward runs (replaces DataParallel's gradient gather,
 * train/adversarial_training_clip.py:184-190,362). */
int rvlm_vit_train_grads_ev(rvlm_vit_handle h, void* workspace, size_t workspace_bytes,
                            const float* x, int B, const rvlm_loss_config* loss, float loss_scale,
                            const rvlm_vit_grads* grads, float* loss_out, float* emb_out,
                            void* const* events, int n_events, void* stream);
/* torch.optim.AdamW semantics on a flat fp32 buffer (train/adversarial_training_clip.py:196-197,363):
 * decoupled weight decay, bias correction with the 1-based step count.
 * Hyper-parameters are doubles (what torch.optim holds); bias corrections formed in double. */
int rvlm_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                    long long n, double lr, double beta1, double beta2, double eps,
                    double weight_decay, int step, void* stream);
/* torch.optim.SGD semantics (`--opt sgd`, train/adversarial_training_clip.py:198-204): coupled
 * weight decay, momentum buffer initialised with the first gradient (step == 1), no dampening. */
int rvlm_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, double lr,
                  double momentum, double weight_decay, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RVLM_H_ */
