// Micro self-test of the "TS" form of tcgen05.mma (A operand read from TMEM): checks the
// operand packing assumption the attention kernels rely on before they depend on it:
//   A[i, k] (16-bit) lives in TMEM lane i, 32-bit column k/2, low half = even k.
// D[128, 64] = A[128, 64] * B[64, 64]^T, A written by tcgen05.st from the row-owner threads.
#include "gemm.cuh"
#include "rvlm_debug.h"

namespace rvlm {

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
ts_mma_test_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B,
                   float* __restrict__ C, int fmt) {
  __shared__ __align__(1024) uint8_t sB[64 * 128];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int r = threadIdx.x, warp = r >> 5;
  if (r == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
  // B[n][k] -> K-major 128B-swizzled rows
  for (int i = r; i < 64 * 8; i += 128) {
    const int n = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + n * 128 + ((c ^ (n & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(B + n * 64 + c * 8);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  uint32_t a[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) a[j] = reinterpret_cast<const uint32_t*>(A + r * 64)[j];
  tmem_st32(tmem + lane_off + 0, a);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (r == 0) {
    const uint32_t idesc = make_idesc_f16((uint32_t)fmt, 128, 64, 0, 0);
    for (int s = 0; s < 4; ++s)
      umma_f16_ts(tmem + 64, tmem + 8 * s, make_smem_desc_sw128(smem_u32(sB) + s * 32, 16, 1024),
                  idesc, s > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 90);
  tc_fence_after();
  uint32_t v0[32], v1[32];
  tmem_ld32(tmem + lane_off + 64, v0);
  tmem_ld32(tmem + lane_off + 96, v1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    C[r * 64 + j] = __uint_as_float(v0[j]);
    C[r * 64 + 32 + j] = __uint_as_float(v1[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

}  // namespace rvlm

extern "C" int rvlm_selftest_ts_mma(const void* A, const void* B, float* C, int fmt, void* stream) {
  rvlm::ts_mma_test_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint16_t*>(A), reinterpret_cast<const uint16_t*>(B), C, fmt);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}
