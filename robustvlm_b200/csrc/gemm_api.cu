// C-ABI entry points for the tcgen05 GEMM family (include/rvlm.h §1).
#include "gemm.cuh"
#include "epilogues.cuh"
#include "../../include/rvlm.h"

using namespace rvlm;

namespace {

template <class Epi>
int dispatch(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int fmt,
             int flags, const Epi& epi, cudaStream_t st) {
  const bool bmn = flags & RVLM_GEMM_B_MN, amn = flags & RVLM_GEMM_A_MN,
             bn128 = flags & RVLM_GEMM_BN128;
  if (fmt != 0 && fmt != 1) return RVLM_ERR_INVAL;
  if (flags & RVLM_GEMM_2CTA) {
    if (amn || bmn || bn128) return RVLM_ERR_UNSUPPORTED;
    return launch_gemm16_2cta(A, lda, B, ldb, M, N, K, fmt, epi, st);
  }
  if (!amn && !bmn) {
    return bn128 ? launch_gemm16<128, false, false>(A, lda, B, ldb, M, N, K, fmt, epi, st)
                 : launch_gemm16<256, false, false>(A, lda, B, ldb, M, N, K, fmt, epi, st);
  }
  if (!amn && bmn) {
    return bn128 ? launch_gemm16<128, false, true>(A, lda, B, ldb, M, N, K, fmt, epi, st)
                 : launch_gemm16<256, false, true>(A, lda, B, ldb, M, N, K, fmt, epi, st);
  }
  if (amn && !bmn) return launch_gemm16<128, true, false>(A, lda, B, ldb, M, N, K, fmt, epi, st);
  return launch_gemm16<128, true, true>(A, lda, B, ldb, M, N, K, fmt, epi, st);
}

}  // namespace

extern "C" int rvlm_gemm16_out16(const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                                 const float* bias, int M, int N, int K, int fmt, int flags,
                                 void* stream) {
  if ((reinterpret_cast<uintptr_t>(C) & 15) || (ldc & 7)) return RVLM_ERR_ALIGN;
  EpiStore16 epi{reinterpret_cast<uint16_t*>(C), ldc, bias, fmt};
  return dispatch(A, lda, B, ldb, M, N, K, fmt, flags, epi, (cudaStream_t)stream);
}

extern "C" int rvlm_gemm16_out32(const void* A, int lda, const void* B, int ldb, float* C, int ldc,
                                 const float* bias, const float* resid, int ldr, int M, int N,
                                 int K, int fmt, int flags, void* stream) {
  if ((reinterpret_cast<uintptr_t>(C) & 15) || (ldc & 3)) return RVLM_ERR_ALIGN;
  if (resid && ((reinterpret_cast<uintptr_t>(resid) & 15) || (ldr & 3))) return RVLM_ERR_ALIGN;
  EpiStore32 epi{C, ldc, bias, resid, ldr};
  return dispatch(A, lda, B, ldb, M, N, K, fmt, flags, epi, (cudaStream_t)stream);
}

extern "C" int rvlm_gemm16_gelu(const void* A, int lda, const void* B, int ldb, void* G, void* Act,
                                int ld, const float* bias, int M, int N, int K, int fmt,
                                void* stream) {
  if ((reinterpret_cast<uintptr_t>(Act) & 15) || (reinterpret_cast<uintptr_t>(G) & 15) || (ld & 7))
    return RVLM_ERR_ALIGN;
  EpiGelu epi{reinterpret_cast<uint16_t*>(G), reinterpret_cast<uint16_t*>(Act), ld, bias, fmt, nullptr};
  return launch_gemm16<256, false, false>(A, lda, B, ldb, M, N, K, fmt, epi, (cudaStream_t)stream);
}

extern "C" int rvlm_gemm16_gelu_bwd(const void* A, int lda, const void* B, int ldb, void* out,
                                    const void* G, int ld, int M, int N, int K, int fmt, int fmt_g,
                                    void* stream) {
  if ((reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(G) & 15) || (ld & 7))
    return RVLM_ERR_ALIGN;
  EpiGeluBwd epi{reinterpret_cast<uint16_t*>(out), ld, reinterpret_cast<const uint16_t*>(G), fmt_g,
                 fmt};
  return launch_gemm16<256, false, false>(A, lda, B, ldb, M, N, K, fmt, epi, (cudaStream_t)stream);
}

extern "C" int rvlm_gemm16_split_out32(const void* A_hi, const void* A_lo, int lda, const void* B_hi,
                                       const void* B_lo, int ldb, float* C, int ldc,
                                       const float* bias, const float* resid, int ldr, int M,
                                       int N, int K, void* stream) {
  if (!A_hi || !A_lo || !B_hi || !B_lo) return RVLM_ERR_INVAL;
  if ((reinterpret_cast<uintptr_t>(C) & 15) || (ldc & 3)) return RVLM_ERR_ALIGN;
  if (resid && ((reinterpret_cast<uintptr_t>(resid) & 15) || (ldr & 3))) return RVLM_ERR_ALIGN;
  EpiStore32 epi{C, ldc, bias, resid, ldr};
  return launch_gemm_fwd(A_hi, A_lo, lda, B_hi, B_lo, ldb, M, N, K, RVLM_FMT_F16, epi,
                         (cudaStream_t)stream);
}
