/* Test hooks of librvlm_b200.so that are NOT part of the product ABI (include/rvlm.h): exported
 * for the GPU tests only. */
#ifndef RVLM_DEBUG_H_
#define RVLM_DEBUG_H_
#include "../../include/rvlm.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Self-test of tcgen05.mma with the A operand in TMEM: C[128,64] = A[128,64] * B[64,64]^T
 * (16-bit A, B row-major; fp32 C). */
int rvlm_selftest_ts_mma(const void* A, const void* B, float* C, int fmt, void* stream);

/* Test hook: pointer of an internal workspace buffer (see vit_engine.cu). */
int rvlm_vit_debug_buffer(rvlm_vit_handle h, void* workspace, int B, int with_backward, int which,
                          int layer, void** ptr_out);

#ifdef __cplusplus
}
#endif
#endif /* RVLM_DEBUG_H_ */
