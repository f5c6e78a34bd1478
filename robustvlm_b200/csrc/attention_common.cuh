// Shared pieces of the tcgen05 attention kernels (attention.cu: the T = 256 / 257 production
// kernels; attention_generic.cu: the generic-T kernels): descriptors, MMA issue helpers, swizzled
// shared-memory accessors, 16-bit pack / unpack, parameter blocks.
#pragma once
#include "gemm.cuh"
#include <cstdlib>
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);


constexpr int AT_THREADS = 288;  // 8 worker warps + 1 control warp
constexpr int AT_CTRL_WARP = 8;
enum { AT_FWD = 0, AT_DQ = 1, AT_DKV = 2 };
constexpr uint32_t AT_TMEM_S = 0, AT_TMEM_ACC = 320, AT_TMEM_ACC2 = 384;

struct AttnParams {
  int B, T, H, W;
  int kvp, nblk, ntile, nbox, boxrows;
  int fmt;               // operand format of every MMA in this launch (FWD: the qkv format;
                         // DQ / DKV: the gradient format -- qkv here is the backward's copy)
  int fmt_o;             // format of the forward output o (read by threads only)
  int fmt_g;             // format of the dqkv output
  const uint16_t* o;     // [B*T, W]   (DQ / DKV: forward output, for D = rowsum(dO * O))
  const uint16_t* d_o;   // [B*T, W]
  float* lse_tmp;        // [B, H, T]  softmax log2-sum-exp recomputed by DQ, read by DKV
  uint16_t* o_out;       // FWD
  float* lse_out;        // FWD
  uint16_t* dqkv;        // [B*T, 3W]  (DQ / DKV)
};

__host__ __device__ constexpr uint32_t idesc_ab(uint32_t afmt, uint32_t bfmt, uint32_t M,
                                                uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | (a_mn << 15) | (b_mn << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[128 x kvp] = A[128 x 64] * B[kvp x 64]^T, both K-major
__device__ __forceinline__ void issue_mma1(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr,
                                           int kvp, uint32_t afmt, uint32_t bfmt) {
  const int n0 = kvp > 256 ? 256 : kvp;
  const uint32_t id0 = idesc_ab(afmt, bfmt, 128, n0, 0, 0);
  const uint32_t id1 = idesc_ab(afmt, bfmt, 128, kvp - n0, 0, 0);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
    umma_f16(tmem_d, ad, make_smem_desc_sw128(b_addr + k * 32, 16, 1024), id0, k > 0);
    if (kvp > 256)
      umma_f16(tmem_d + 256, ad, make_smem_desc_sw128(b_addr + 256 * 128 + k * 32, 16, 1024), id1,
               k > 0);
  }
}

// D[128 x 64] = A[128 x kvp] (K-major, 64-wide 16 KB blocks) * B[kvp x 64] (MN-major)
__device__ __forceinline__ void issue_mma2(uint32_t tmem_d, uint32_t p_addr, uint32_t b_addr,
                                           int kvp, uint32_t afmt, uint32_t bfmt) {
  const uint32_t id = idesc_ab(afmt, bfmt, 128, 64, 0, 1);
  const int steps = kvp >> 4;
  for (int s = 0; s < steps; ++s) {
    const uint64_t ad =
        make_smem_desc_sw128(p_addr + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024);
    const uint64_t bd = make_smem_desc_sw128(b_addr + s * 2048, 8192, 1024);
    umma_f16(tmem_d, ad, bd, id, s > 0);
  }
}

// byte offset of 8 consecutive 16-bit elements (row r, columns c0..c0+7, c0 % 8 == 0) in
// the K-major 128B-swizzled block layout
__device__ __forceinline__ uint32_t p_off(int r, int c0) {
  return (uint32_t)((c0 >> 6) * 16384 + r * 128 + ((((c0 & 63) >> 3) ^ (r & 7)) << 4));
}

__device__ __forceinline__ float dot_rows16(const uint16_t* a, int fa, const uint16_t* b, int fb) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 ua = reinterpret_cast<const uint4*>(a)[j];
    const uint4 ub = reinterpret_cast<const uint4*>(b)[j];
    const uint32_t* pa = &ua.x;
    const uint32_t* pb = &ub.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack2(pa[i], fa), y = unpack2(pb[i], fb);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  return acc;
}

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 32 (or, for the ragged tail chunk, 16) fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void ld_chunk(uint32_t taddr, int c, int n32, uint32_t (&v)[32]) {
  if (c < n32) tmem_ld32(taddr + c * 32, v);
  else tmem_ld16(taddr + c * 32, *reinterpret_cast<uint32_t(*)[16]>(v));
  tmem_ld_wait();
}

template <int FMT>
__device__ __forceinline__ uint32_t packc(float a, float b) {
  return FMT ? pack2_bf16(a, b) : pack2_f16(a, b);
}
template <int FMT>
__device__ __forceinline__ float2 unpackc(uint32_t u) {
  return unpack2(u, FMT);
}
template <int FMT>
__device__ __forceinline__ float ld1(const uint16_t* p) {
  return FMT ? __uint_as_float((uint32_t)*p << 16)
             : __half2float(*reinterpret_cast<const __half*>(p));
}
struct AttnFwd2Params {
  int B, T, H, W;
  int tm;      // rows / keys handled on the tensor cores: min(T, 256)
  int kvm;     // tm rounded up to 16 (UMMA N extent, P / V K-extent)
  int nblk;    // 64-column blocks of P
  int ntile;   // 128-row tiles: 1 or 2
  int tail;    // 1 if key index 256 exists (T == 257)
  int fmt;
  const uint16_t* qkv;
  uint16_t* o_out;
  float* lse_out;
};

struct AttnBwd2Params {
  int B, T, H, W;
  int tm, kvm, nblk, ntile, tail;
  int fmt_o;             // format of the forward output o (read by threads only)
  const uint16_t* qkv;   // [B*T, 3W] backward copy of q|k|v (operand format)
  const uint16_t* o;     // [B*T, W]
  const uint16_t* d_o;   // [B*T, W]  (operand format)
  float* lse_tmp;        // [B, H, T] log2-sum-exp, written by DQ, read by DKV
  uint16_t* dqkv;        // [B*T, 3W] (operand format)
};

// dot product of a 64-element row of a 128B-swizzled K-major smem tile with a 64-element
// vector held in plain shared memory
template <int FMT>
__device__ __forceinline__ float dot_swz_row(const uint8_t* tile, int r, const uint16_t* vec) {
  float acc = 0.f;
  const uint8_t* row = tile + r * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 a = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
    const uint4 b = *reinterpret_cast<const uint4*>(vec + c * 8);
    const uint32_t* pa = &a.x;
    const uint32_t* pb = &b.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpackc<FMT>(pa[i]), y = unpackc<FMT>(pb[i]);
      acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
    }
  }
  return acc;
}
template <int FMT>
__device__ __forceinline__ float dot_vec(const uint16_t* a, const uint16_t* b) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 ua = *reinterpret_cast<const uint4*>(a + c * 8);
    const uint4 ub = *reinterpret_cast<const uint4*>(b + c * 8);
    const uint32_t* pa = &ua.x;
    const uint32_t* pb = &ub.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpackc<FMT>(pa[i]), y = unpackc<FMT>(pb[i]);
      acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
    }
  }
  return acc;
}
// element (row kv, column d) of a 128B-swizzled [rows x 64] 16-bit tile
template <int FMT>
__device__ __forceinline__ float swz_at(const uint8_t* tile, int kv, int d) {
  return ld1<FMT>(reinterpret_cast<const uint16_t*>(tile + kv * 128 + (((d >> 3) ^ (kv & 7)) << 4) +
                                                    (d & 7) * 2));
}
__device__ __forceinline__ float blk256_max(float v, float* red, int warp, int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffff, v, o));
  if (lane == 0) red[warp] = v;
  asm volatile("bar.sync 3, 256;" ::: "memory");
  float t = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
  asm volatile("bar.sync 3, 256;" ::: "memory");
  return t;
}
__device__ __forceinline__ float blk256_sum(float v, float* red, int warp, int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  if (lane == 0) red[warp] = v;
  asm volatile("bar.sync 3, 256;" ::: "memory");
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  asm volatile("bar.sync 3, 256;" ::: "memory");
  return t;
}

// generic-T launchers (attention_generic.cu)
int attn_fwd2_launch(const CUtensorMap& tmTile, const CUtensorMap& tmFull, const AttnFwd2Params& P,
                     cudaStream_t st);
int attn_bwd_generic(const void* qkv16, const void* o16, const void* do16, float* lse_tmp,
                     void* dqkv16, int B, int T, int H, int fmt_o, int fmt_g, cudaStream_t st);

}  // namespace rvlm
