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
 */
#ifndef RVLM_H_
#define RVLM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVLM_ABI_VERSION 1

#define RVLM_ERR_INVAL (-22)
#define RVLM_ERR_ALIGN (-23)
#define RVLM_ERR_DRIVER (-24)
#define RVLM_ERR_TMAP (-25)
#define RVLM_ERR_NOMEM (-12)
#define RVLM_ERR_UNSUPPORTED (-38)

#define RVLM_FMT_F16 0
#define RVLM_FMT_BF16 1

int rvlm_abi_version(void);
const char* rvlm_strerror(int code);

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
 * ------------------------------------------------------------------------ */
#define RVLM_GEMM_B_MN 1
#define RVLM_GEMM_A_MN 2
#define RVLM_GEMM_BN128 4

/* 16-bit output */
int rvlm_gemm16_out16(const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                      const float* bias, int M, int N, int K, int fmt, int flags, void* stream);
/* fp32 output with optional bias and residual */
int rvlm_gemm16_out32(const void* A, int lda, const void* B, int ldb, float* C, int ldc,
                      const float* bias, const float* resid, int ldr, int M, int N, int K, int fmt,
                      int flags, void* stream);
/* H = A*B^T + bias (pre-activation, may be NULL), Act = QuickGELU(H); both 16-bit [M,N] pitch ld */
int rvlm_gemm16_gelu(const void* A, int lda, const void* B, int ldb, void* H, void* Act, int ld,
                     const float* bias, int M, int N, int K, int fmt, void* stream);
/* out = (A*B^T) .* QuickGELU'(H) ; 16-bit [M,N] pitch ld; H in fmt_h, A/B/out in fmt */
int rvlm_gemm16_gelu_bwd(const void* A, int lda, const void* B, int ldb, void* out, const void* H,
                         int ld, int M, int N, int K, int fmt, int fmt_h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RVLM_H_ */
