// tcgen05 multi-head attention for the CLIP ViT (head dim 64, T <= 272 tokens, no
// mask): forward and the input-gradient backward (nn.MultiheadAttention inside
// open_clip's ResidualAttentionBlock; reached from
// train/pgd_train.py:32,38 through ClipVisionModel.forward).
//
// One skeleton, three modes.  Per work item (image b, head h) a persistent CTA runs,
// for each 128-row tile:
//
//   FWD  : S = Q_i K^T            -> P = softmax rows (fp32, exp2)      -> O_i  = P V
//   DQ   : S = Q_i K^T            -> P = exp(S/8 - lse_i)
//          dP = dO_i V^T          -> dS = P (dP - D_i) / 8              -> dQ_i = dS K
//   DKV  : S^T = K_j Q^T          -> P^T = exp(S^T/8 - lse)             -> dV_j = P^T dO
//          dP^T = V_j dO^T        -> dS^T = P^T (dP^T - D) / 8          -> dK_j = dS^T Q
//
// "mma1" is a [128 x kvp] product of two K-major operands (kvp = T rounded up to 16,
// the full token range is one UMMA N extent, 256 + 16 for T = 257); its fp32 result
// lives in TMEM, one row per thread of the four worker warps, so the row softmax needs
// no cross-thread reduction.  The 16-bit P / dS tile is written by those threads
// straight into 128B-swizzled shared memory as the K-major A operand of "mma2", whose
// B operand (V, K, dO or Q: [tokens, 64] row-major) is consumed MN-major, so nothing is
// ever transposed in memory.  TMA loads everything (3-D tensor maps over [B, T, cols] so
// rows beyond T are zero-filled per image).
//
// Warp roles: warps 0-3 = row workers (TMEM lane == tile row), warp 4 = TMEM allocator
// and, lane 0, the TMA + MMA issuer.
#include "gemm.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

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

template <int MODE>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmTile,   // qkv, box 64 x 128
               const __grid_constant__ CUtensorMap tmFull,   // qkv, box 64 x boxrows
               const __grid_constant__ CUtensorMap tmDoTile, // dO,  box 64 x 128
               const __grid_constant__ CUtensorMap tmDoFull, // dO,  box 64 x boxrows
               AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int kvp = P.kvp, T = P.T, W = P.W;
  const int full_bytes = kvp * 128;
  uint8_t* s_full0 = smem;
  uint8_t* s_full1 = s_full0 + full_bytes;
  uint8_t* s_tile0 = s_full1 + full_bytes;  // [2][16384]
  uint8_t* s_tile1 = s_tile0 + 2 * 16384;   // [2][16384]
  uint8_t* s_p = s_tile1 + 2 * 16384;       // [nblk][16384]
  float* s_lse = reinterpret_cast<float*>(s_p + P.nblk * 16384);  // [kvp] (DKV)
  float* s_dvec = s_lse + kvp;                                    // [kvp] (DKV)
  float* s_red = s_dvec + kvp;   // [2][128] row max of each column half
  float* s_red2 = s_red + 256;   // [2][128] row sum of each column half
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red2 + 256);
  uint64_t* bar_full = bars + 0;     // full0 + full1 landed
  uint64_t* bar_tile = bars + 1;     // [2]
  uint64_t* bar_s = bars + 3;        // mma1 result ready
  uint64_t* bar_p = bars + 4;        // P / dS written (128 arrivals)
  uint64_t* bar_acc = bars + 5;      // mma2 result(s) ready
  uint64_t* bar_acc_free = bars + 6; // accumulators drained (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(&bar_tile[0], 1);
    mbar_init(&bar_tile[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc_free, 256);
    fence_barrier_init();
  }
  if (warp == AT_CTRL_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_items = P.B * P.H;
  const uint32_t f = (uint32_t)P.fmt;
  // column offsets (elements) of q / k / v inside a qkv row
  const int colQ = 0, colK = W, colV = 2 * W;

  if (warp == AT_CTRL_WARP) {
    if (lane == 0) {
      // ===================== TMA + MMA issuer =====================
      uint32_t g = 0;  // global tile counter
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int b = item / P.H, h = item % P.H;
        if (it > 0) mbar_wait(bar_acc, (g - 1) & 1, 10);  // every MMA of the previous item retired
        mbar_arrive_expect_tx(bar_full, 2 * full_bytes);
        for (int i = 0; i < P.nbox; ++i) {
          const int r0 = i * P.boxrows;
          if (MODE == AT_DKV) {
            tma_load_3d(s_full0 + r0 * 128, &tmFull, bar_full, colQ + h * 64, r0, b);
            tma_load_3d(s_full1 + r0 * 128, &tmDoFull, bar_full, h * 64, r0, b);
          } else {
            tma_load_3d(s_full0 + r0 * 128, &tmFull, bar_full, colK + h * 64, r0, b);
            tma_load_3d(s_full1 + r0 * 128, &tmFull, bar_full, colV + h * 64, r0, b);
          }
        }
        auto load_tile = [&](int j, uint32_t gg) {
          const int buf = gg & 1;
          mbar_arrive_expect_tx(&bar_tile[buf], MODE == AT_FWD ? 16384 : 32768);
          if (MODE == AT_DKV) {
            tma_load_3d(s_tile0 + buf * 16384, &tmTile, &bar_tile[buf], colK + h * 64, j * 128, b);
            tma_load_3d(s_tile1 + buf * 16384, &tmTile, &bar_tile[buf], colV + h * 64, j * 128, b);
          } else {
            tma_load_3d(s_tile0 + buf * 16384, &tmTile, &bar_tile[buf], colQ + h * 64, j * 128, b);
            if (MODE == AT_DQ)
              tma_load_3d(s_tile1 + buf * 16384, &tmDoTile, &bar_tile[buf], h * 64, j * 128, b);
          }
        };
        load_tile(0, g);
        mbar_wait(bar_full, it & 1, 11);
        for (int j = 0; j < P.ntile; ++j, ++g) {
          const int buf = g & 1;
          if (j + 1 < P.ntile) load_tile(j + 1, g + 1);
          mbar_wait(&bar_tile[buf], (g >> 1) & 1, 12);
          tc_fence_after();
          const uint32_t t0 = smem_u32(s_tile0 + buf * 16384);
          const uint32_t t1 = smem_u32(s_tile1 + buf * 16384);
          const uint32_t f0 = smem_u32(s_full0), f1 = smem_u32(s_full1), pa = smem_u32(s_p);
          if (MODE == AT_FWD) {
            issue_mma1(tmem + AT_TMEM_S, t0, f0, kvp, f, f);          // S = Q K^T
            umma_commit(bar_s);
            mbar_wait(bar_p, g & 1, 13);
            tc_fence_after();
            if (g > 0) mbar_wait(bar_acc_free, (g - 1) & 1, 14);
            tc_fence_after();
            issue_mma2(tmem + AT_TMEM_ACC, pa, f1, kvp, f, f);        // O = P V
            umma_commit(bar_acc);
          } else if (MODE == AT_DQ) {
            issue_mma1(tmem + AT_TMEM_S, t0, f0, kvp, f, f);          // S = Q K^T
            umma_commit(bar_s);
            mbar_wait(bar_p, (2 * g) & 1, 13);
            tc_fence_after();
            issue_mma1(tmem + AT_TMEM_S, t1, f1, kvp, f, f);          // dP = dO V^T
            umma_commit(bar_s);
            mbar_wait(bar_p, (2 * g + 1) & 1, 15);
            tc_fence_after();
            if (g > 0) mbar_wait(bar_acc_free, (g - 1) & 1, 14);
            tc_fence_after();
            issue_mma2(tmem + AT_TMEM_ACC, pa, f0, kvp, f, f);        // dQ = dS K
            umma_commit(bar_acc);
          } else {
            issue_mma1(tmem + AT_TMEM_S, t0, f0, kvp, f, f);          // S^T = K Q^T
            umma_commit(bar_s);
            mbar_wait(bar_p, (2 * g) & 1, 13);
            tc_fence_after();
            if (g > 0) mbar_wait(bar_acc_free, (g - 1) & 1, 14);
            tc_fence_after();
            issue_mma2(tmem + AT_TMEM_ACC, pa, f1, kvp, f, f);        // dV = P^T dO
            issue_mma1(tmem + AT_TMEM_S, t1, f1, kvp, f, f);          // dP^T = V dO^T
            umma_commit(bar_s);                                       // (also: dV retired)
            mbar_wait(bar_p, (2 * g + 1) & 1, 15);
            tc_fence_after();
            issue_mma2(tmem + AT_TMEM_ACC2, pa, f0, kvp, f, f);       // dK = dS^T Q
            umma_commit(bar_acc);
          }
        }
      }
    }
  } else {
    // ===================== row workers =====================
    // 8 warps: thread (r, hsel) owns tile row r (TMEM lane) and one half of the columns;
    // the two halves exchange the row max / row sum through shared memory.
    const int r = threadIdx.x & 127, hsel = threadIdx.x >> 7;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float c2 = 0.125f * 1.4426950408889634f;  // scale * log2(e)
    const int n32 = kvp >> 5, nch = n32 + ((kvp & 31) >> 4);
    const int c_lo = hsel ? (nch + 1) / 2 : 0, c_hi = hsel ? nch : (nch + 1) / 2;
    const uint32_t ts = tmem + lane_off + AT_TMEM_S;
    uint32_t g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / P.H, h = item % P.H;
      if (MODE == AT_DKV) {
        // per-token log2-sum-exp (from the DQ pass) and D = rowsum(dO * O) of this (b, h)
        for (int q = threadIdx.x; q < kvp; q += 256) {
          float L = INFINITY, Dv = 0.f;
          if (q < T) {
            L = P.lse_tmp[((size_t)b * P.H + h) * T + q];
            const size_t ro = ((size_t)b * T + q) * W + h * 64;
            Dv = dot_rows16(P.o + ro, P.fmt_o, P.d_o + ro, P.fmt);
          }
          s_lse[q] = L;
          s_dvec[q] = Dv;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      for (int j = 0; j < P.ntile; ++j, ++g) {
        const int row = j * 128 + r;  // q row (FWD/DQ) or kv row (DKV) inside the image
        const bool valid = row < T;
        const bool warp_valid = (j * 128 + (warp & 3) * 32) < T;
        const size_t grow = (size_t)b * T + row;
        float Dr = 0.f;
        if (MODE == AT_DQ && valid)
          Dr = dot_rows16(P.o + grow * W + h * 64, P.fmt_o, P.d_o + grow * W + h * 64, P.fmt);
        // ---------------- pass 1: P from the mma1 result ----------------
        mbar_wait(bar_s, (MODE == AT_FWD ? g : 2 * g) & 1, 20);
        tc_fence_after();
        float mc = 0.f, rsum = 0.f;
        if (MODE != AT_DKV) {
          // row max over this thread's columns, then across the two halves
          float mx = -INFINITY;
          if (warp_valid) {
            for (int c = c_lo; c < c_hi; ++c) {
              uint32_t v[32];
              ld_chunk(ts, c, n32, v);
              const int nv = c < n32 ? 32 : 16;
              if (c * 32 + nv <= T) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < nv) mx = fmaxf(mx, __uint_as_float(v[i]));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < nv && c * 32 + i < T) mx = fmaxf(mx, __uint_as_float(v[i]));
              }
            }
          }
          s_red[hsel * 128 + r] = mx;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          mc = fmaxf(s_red[r], s_red[128 + r]) * c2;
        }
        if (warp_valid) {
          for (int c = c_lo; c < c_hi; ++c) {
            uint32_t v[32];
            ld_chunk(ts, c, n32, v);
            const int nv = c < n32 ? 32 : 16;
            float p[32];
            if (MODE == AT_DKV) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                p[i] = (i < nv && valid)
                           ? fast_ex2(fmaf(__uint_as_float(v[i]), c2, -s_lse[c * 32 + i]))
                           : 0.f;
            } else if (c * 32 + nv <= T) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                p[i] = (i < nv) ? fast_ex2(fmaf(__uint_as_float(v[i]), c2, -mc)) : 0.f;
                rsum += p[i];
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                p[i] = (i < nv && c * 32 + i < T)
                           ? fast_ex2(fmaf(__uint_as_float(v[i]), c2, -mc))
                           : 0.f;
                rsum += p[i];
              }
            }
            if (MODE == AT_DQ && !valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) p[i] = 0.f;
            }
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {
              if (q8 * 8 < nv) {
                uint4 u;
                u.x = pack2(p[q8 * 8 + 0], p[q8 * 8 + 1], P.fmt);
                u.y = pack2(p[q8 * 8 + 2], p[q8 * 8 + 3], P.fmt);
                u.z = pack2(p[q8 * 8 + 4], p[q8 * 8 + 5], P.fmt);
                u.w = pack2(p[q8 * 8 + 6], p[q8 * 8 + 7], P.fmt);
                *reinterpret_cast<uint4*>(s_p + p_off(r, c * 32 + q8 * 8)) = u;
              }
            }
          }
        }
        if (MODE != AT_DKV) s_red2[hsel * 128 + r] = rsum;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_p);
        float rinv = 1.f;
        if (MODE != AT_DKV) {
          asm volatile("bar.sync 2, 256;" ::: "memory");
          rsum = s_red2[r] + s_red2[128 + r];
          rinv = 1.f / rsum;
          if (MODE == AT_DQ && valid && hsel == 0)
            P.lse_tmp[((size_t)b * P.H + h) * T + row] = mc + log2f(rsum);
        }
        if (MODE != AT_FWD) {
          // ------------- pass 2: dS = P (dP - D) / 8, in place over P -------------
          mbar_wait(bar_s, (2 * g + 1) & 1, 23);
          tc_fence_after();
          if (warp_valid) {
            const float sc = 0.125f * rinv;  // DQ: P is normalised here by 1 / rowsum
            for (int c = c_lo; c < c_hi; ++c) {
              uint32_t v[32];
              ld_chunk(ts, c, n32, v);
              const int nv = c < n32 ? 32 : 16;
#pragma unroll
              for (int q8 = 0; q8 < 4; ++q8) {
                if (q8 * 8 < nv) {
                  uint4* sp = reinterpret_cast<uint4*>(s_p + p_off(r, c * 32 + q8 * 8));
                  const uint4 pu = *sp;
                  const uint32_t* pw = &pu.x;
                  float d[8];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 pp = unpack2(pw[i], P.fmt);
                    const int col = c * 32 + q8 * 8 + 2 * i;
                    const float D0 = (MODE == AT_DQ) ? Dr : s_dvec[col];
                    const float D1 = (MODE == AT_DQ) ? Dr : s_dvec[col + 1];
                    d[2 * i] = pp.x * sc * (__uint_as_float(v[q8 * 8 + 2 * i]) - D0);
                    d[2 * i + 1] = pp.y * sc * (__uint_as_float(v[q8 * 8 + 2 * i + 1]) - D1);
                  }
                  uint4 u;
                  u.x = pack2(d[0], d[1], P.fmt);
                  u.y = pack2(d[2], d[3], P.fmt);
                  u.z = pack2(d[4], d[5], P.fmt);
                  u.w = pack2(d[6], d[7], P.fmt);
                  *sp = u;
                }
              }
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(bar_p);
        }
        // ---------------- epilogue: accumulators -> HBM ----------------
        mbar_wait(bar_acc, g & 1, 24);
        tc_fence_after();
        if (warp_valid) {
          if (MODE == AT_DKV) {
            // hsel 0: dV (ACC, columns 2W..), hsel 1: dK (ACC2, columns W..)
            uint32_t v0[32], v1[32];
            const uint32_t col = hsel ? AT_TMEM_ACC2 : AT_TMEM_ACC;
            tmem_ld32(tmem + lane_off + col, v0);
            tmem_ld32(tmem + lane_off + col + 32, v1);
            tmem_ld_wait();
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(P.dqkv + grow * 3 * W + (hsel ? W : 2 * W) +
                                                    h * 64);
#pragma unroll
              for (int q8 = 0; q8 < 8; ++q8) {
                const uint32_t* sv = q8 < 4 ? &v0[q8 * 8] : &v1[(q8 - 4) * 8];
                uint4 u;
                u.x = pack2(__uint_as_float(sv[0]), __uint_as_float(sv[1]), P.fmt_g);
                u.y = pack2(__uint_as_float(sv[2]), __uint_as_float(sv[3]), P.fmt_g);
                u.z = pack2(__uint_as_float(sv[4]), __uint_as_float(sv[5]), P.fmt_g);
                u.w = pack2(__uint_as_float(sv[6]), __uint_as_float(sv[7]), P.fmt_g);
                dst[q8] = u;
              }
            }
          } else {
            // FWD: O * (1 / rowsum); DQ: dQ.  Each half stores 32 of the 64 columns.
            uint32_t v0[32];
            tmem_ld32(tmem + lane_off + AT_TMEM_ACC + hsel * 32, v0);
            tmem_ld_wait();
            if (valid) {
              const float sc = (MODE == AT_FWD) ? rinv : 1.f;
              const int ofmt = (MODE == AT_FWD) ? P.fmt : P.fmt_g;
              uint16_t* base = (MODE == AT_FWD) ? P.o_out + grow * W + h * 64
                                                : P.dqkv + grow * 3 * W + h * 64;
              uint4* dst = reinterpret_cast<uint4*>(base + hsel * 32);
#pragma unroll
              for (int q8 = 0; q8 < 4; ++q8) {
                const uint32_t* sv = &v0[q8 * 8];
                uint4 u;
                u.x = pack2(__uint_as_float(sv[0]) * sc, __uint_as_float(sv[1]) * sc, ofmt);
                u.y = pack2(__uint_as_float(sv[2]) * sc, __uint_as_float(sv[3]) * sc, ofmt);
                u.z = pack2(__uint_as_float(sv[4]) * sc, __uint_as_float(sv[5]) * sc, ofmt);
                u.w = pack2(__uint_as_float(sv[6]) * sc, __uint_as_float(sv[7]) * sc, ofmt);
                dst[q8] = u;
              }
              if (MODE == AT_FWD && hsel == 0)
                P.lse_out[((size_t)b * P.H + h) * T + row] =
                    (mc + log2f(rsum)) * 0.6931471805599453f;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar_acc_free);
      }
      if (MODE == AT_DKV) asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AT_CTRL_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// =====================================================================================
// Forward, version 2: two row groups in flight, one thread per row.
//
// The 257 = 256 + 1 token count of ViT-L/14 is peeled: the tensor cores handle query rows
// and keys [0, 256) -- two 128-row tiles per (image, head), processed CONCURRENTLY by two
// groups of four warps, each thread owning one whole row (no cross-thread softmax
// reduction) -- while the one remaining key is a 64-MAC dot product per row on the CUDA
// cores, folded into the same softmax, and the one remaining query row is computed by the
// CUDA-core kernel.  This removes the 1-valid-row third tile and fits two S accumulators
// (2 x 256 columns, the O accumulator aliasing the consumed S) into the 512 TMEM columns.
// =====================================================================================
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

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmTile,   // qkv, box 64 x 128
                 const __grid_constant__ CUtensorMap tmFull,   // qkv, box 64 x kvm
                 AttnFwd2Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int kvm = P.kvm, T = P.T, W = P.W;
  const int full_bytes = kvm * 128;
  uint8_t* s_k = smem;
  uint8_t* s_v = s_k + full_bytes;
  uint8_t* s_q = s_v + full_bytes;                 // [2][16384]
  uint8_t* s_p = s_q + 2 * 16384;                  // [2][nblk * 16384]
  uint16_t* s_tail = reinterpret_cast<uint16_t*>(s_p + 2 * P.nblk * 16384);  // [2][2][64]
  uint16_t* s_qtail = s_tail + 256;                // [64] the peeled query row
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_qtail + 64);
  uint64_t* bar_k = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_q = bars + 2;       // [2]
  uint64_t* bar_s = bars + 4;       // [2]
  uint64_t* bar_p = bars + 6;       // [2] 128 arrivals
  uint64_t* bar_o = bars + 8;       // [2]
  uint64_t* bar_ofree = bars + 10;  // [2] 128 arrivals
  uint64_t* bar_tail = bars + 12;   // 128 arrivals: peeled query row done with K / V
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_tail, 128);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bar_q[g], 1);
      mbar_init(&bar_s[g], 1);
      mbar_init(&bar_p[g], 128);
      mbar_init(&bar_o[g], 1);
      mbar_init(&bar_ofree[g], 128);
    }
    fence_barrier_init();
  }
  if (warp == AT_CTRL_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_items = P.B * P.H;
  const uint32_t f = (uint32_t)P.fmt;
  const int ng = P.ntile;  // active groups (group g <-> tile g)

  if (warp == AT_CTRL_WARP) {
    if (lane == 0) {
      // ===================== TMA + MMA issuer =====================
      auto load_k = [&](int item) {
        mbar_arrive_expect_tx(bar_k, full_bytes);
        tma_load_3d(s_k, &tmFull, bar_k, W + (item % P.H) * 64, 0, item / P.H);
      };
      auto load_v = [&](int item) {
        mbar_arrive_expect_tx(bar_v, full_bytes);
        tma_load_3d(s_v, &tmFull, bar_v, 2 * W + (item % P.H) * 64, 0, item / P.H);
      };
      auto load_q = [&](int item, int g) {
        mbar_arrive_expect_tx(&bar_q[g], 16384);
        tma_load_3d(s_q + g * 16384, &tmTile, &bar_q[g], (item % P.H) * 64, g * 128, item / P.H);
      };
      const uint32_t idS = idesc_ab(f, f, 128, kvm, 0, 0);
      int it = 0;
      if ((int)blockIdx.x < n_items) {
        load_k(blockIdx.x);
        load_v(blockIdx.x);
        for (int g = 0; g < ng; ++g) load_q(blockIdx.x, g);
      }
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int next = item + gridDim.x;
        mbar_wait(bar_k, ph, 30);
        for (int g = 0; g < ng; ++g) {
          mbar_wait(&bar_q[g], ph, 31);
          if (it > 0) mbar_wait(&bar_ofree[g], ph ^ 1, 32);
          tc_fence_after();
          const uint32_t qa = smem_u32(s_q + g * 16384), ka = smem_u32(s_k);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem + g * 256, make_smem_desc_sw128(qa + k * 32, 16, 1024),
                     make_smem_desc_sw128(ka + k * 32, 16, 1024), idS, k > 0);
          umma_commit(&bar_s[g]);
        }
        for (int g = 0; g < ng; ++g) mbar_wait(&bar_s[g], ph, 33);   // K consumed by the MMAs
        if (P.tail) mbar_wait(bar_tail, ph, 37);                     // ... and by the peeled row
        if (next < n_items) load_k(next);
        mbar_wait(bar_v, ph, 34);
        for (int g = 0; g < ng; ++g) {
          mbar_wait(&bar_p[g], ph, 35);                              // P written, Q consumed
          tc_fence_after();
          if (next < n_items) load_q(next, g);
          issue_mma2(tmem + g * 256, smem_u32(s_p + g * P.nblk * 16384), smem_u32(s_v), kvm, f, f);
          umma_commit(&bar_o[g]);
        }
        for (int g = 0; g < ng; ++g) mbar_wait(&bar_o[g], ph, 36);   // V consumed
        if (next < n_items) load_v(next);
      }
    }
  } else {
    // ===================== row workers: group g = warp / 4, thread = one row =====================
    const int g = warp >> 2, r = threadIdx.x & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t ts = tmem + lane_off + g * 256;
    const float c2 = 0.125f * 1.4426950408889634f;
    const int n32 = kvm >> 5, nch = n32 + ((kvm & 31) >> 4);
    uint8_t* my_p = s_p + g * P.nblk * 16384;
    const uint16_t* kt = s_tail + g * 128;       // k_tail[64]
    const uint16_t* vt = kt + 64;                // v_tail[64]
    if (g < ng) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int b = item / P.H, h = item % P.H;
        const int row = g * 128 + r;
        const bool valid = row < P.tm;
        const bool warp_valid = (g * 128 + (warp & 3) * 32) < P.tm;
        if (P.tail) {
          if (r < 16) {  // this group's copy of the peeled key / value row (index 256)
            const uint16_t* src = P.qkv + ((size_t)b * T + 256) * 3 * W + h * 64 +
                                  (r < 8 ? W : 2 * W) + (r & 7) * 8;
            *reinterpret_cast<uint4*>(s_tail + g * 128 + r * 8) =
                *reinterpret_cast<const uint4*>(src);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        }
        if (P.tail && g == ng - 1) {
          // ---- the peeled QUERY row (token 256) on the CUDA cores, in the shadow of the S
          // MMAs: scores against all 257 keys, softmax, P V.  Scratch = this group's P buffer.
          float* sc = reinterpret_cast<float*>(my_p);   // [272] scores / probabilities
          float* red = sc + 288;                        // [8] warp partials
          float* oh = red + 8;                          // [2][64] output halves
          if (r < 8)
            *reinterpret_cast<uint4*>(s_qtail + r * 8) = *reinterpret_cast<const uint4*>(
                P.qkv + ((size_t)b * T + 256) * 3 * W + h * 64 + r * 8);
          mbar_wait(bar_k, ph, 43);
          mbar_wait(bar_v, ph, 44);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          float qv[64];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 qu = *reinterpret_cast<const uint4*>(s_qtail + c * 8);
            const uint32_t* qa = &qu.x;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 x = unpack2(qa[i], P.fmt);
              qv[c * 8 + 2 * i] = x.x;
              qv[c * 8 + 2 * i + 1] = x.y;
            }
          }
          float lmax = -INFINITY;
          for (int kv = r; kv < T; kv += 128) {
            float acc = 0.f;
            if (kv < 256) {
              const uint8_t* krow = s_k + kv * 128;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 ku = *reinterpret_cast<const uint4*>(krow + ((c ^ (kv & 7)) << 4));
                const uint32_t* ka = &ku.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 y = unpack2(ka[i], P.fmt);
                  acc = fmaf(qv[c * 8 + 2 * i], y.x, fmaf(qv[c * 8 + 2 * i + 1], y.y, acc));
                }
              }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 ku = *reinterpret_cast<const uint4*>(kt + c * 8);
                const uint32_t* ka = &ku.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 y = unpack2(ka[i], P.fmt);
                  acc = fmaf(qv[c * 8 + 2 * i], y.x, fmaf(qv[c * 8 + 2 * i + 1], y.y, acc));
                }
              }
            }
            sc[kv] = acc;
            lmax = fmaxf(lmax, acc);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffff, lmax, o));
          if (lane == 0) red[warp & 3] = lmax;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          const float tmc = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * c2;
          float lsum = 0.f;
          for (int kv = r; kv < T; kv += 128) {
            const float e = fast_ex2(fmaf(sc[kv], c2, -tmc));
            sc[kv] = e;
            lsum += e;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffff, lsum, o);
          if (lane == 0) red[4 + (warp & 3)] = lsum;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          const float tsum = (red[4] + red[5]) + (red[6] + red[7]);
          {  // o[d] over this thread's half of the keys
            const int d = r & 63, hf = r >> 6;
            const int k0 = hf * 128, k1 = hf ? 256 : 128;
            float acc = 0.f;
            for (int kv = k0; kv < k1; ++kv) {
              const uint16_t hv = *reinterpret_cast<const uint16_t*>(
                  s_v + kv * 128 + (((d >> 3) ^ (kv & 7)) << 4) + (d & 7) * 2);
              const float v = P.fmt ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&hv))
                                    : __half2float(*reinterpret_cast<const __half*>(&hv));
              acc = fmaf(sc[kv], v, acc);
            }
            if (hf == 1) {
              const uint16_t hv = vt[d];
              const float v = P.fmt ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&hv))
                                    : __half2float(*reinterpret_cast<const __half*>(&hv));
              acc = fmaf(sc[256], v, acc);
            }
            oh[hf * 64 + d] = acc;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          if (r < 32) {
            const float inv = 1.f / tsum;
            const float o0 = (oh[2 * r] + oh[64 + 2 * r]) * inv;
            const float o1 = (oh[2 * r + 1] + oh[64 + 2 * r + 1]) * inv;
            reinterpret_cast<uint32_t*>(P.o_out + ((size_t)b * T + 256) * W + h * 64)[r] =
                pack2(o0, o1, P.fmt);
            if (r == 0)
              P.lse_out[((size_t)b * P.H + h) * T + 256] = (tmc + log2f(tsum)) * 0.6931471805599453f;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");  // scratch free again
          mbar_arrive(bar_tail);
        }
        mbar_wait(&bar_q[g], ph, 42);  // (complete long ago) makes the TMA-written Q tile visible
        mbar_wait(&bar_s[g], ph, 40);
        tc_fence_after();
        float mx = -INFINITY, s_t = 0.f, sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
        float p_t = 0.f;
        if (warp_valid) {
          if (P.tail) {  // score against the peeled key: q_row . k_tail from swizzled smem
            const uint8_t* qrow = s_q + g * 16384 + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 qu = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
              const uint4 ku = *reinterpret_cast<const uint4*>(kt + c * 8);
              const uint32_t* qa = &qu.x;
              const uint32_t* ka = &ku.x;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 x = unpack2(qa[i], P.fmt), y = unpack2(ka[i], P.fmt);
                s_t = fmaf(x.x, y.x, fmaf(x.y, y.y, s_t));
              }
            }
            mx = s_t;
          }
          for (int c = 0; c < nch; ++c) {
            uint32_t v[32];
            ld_chunk(ts, c, n32, v);
            const int nv = c < n32 ? 32 : 16;
            if (c * 32 + nv <= P.tm) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nv) mx = fmaxf(mx, __uint_as_float(v[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nv && c * 32 + i < P.tm) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
          }
          const float mc = mx * c2;
          if (P.tail) p_t = fast_ex2(fmaf(s_t, c2, -mc));
          for (int c = 0; c < nch; ++c) {
            uint32_t v[32];
            ld_chunk(ts, c, n32, v);
            const int nv = c < n32 ? 32 : 16;
            float p[32];
            const bool full = c * 32 + nv <= P.tm;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              p[i] = (i < nv && (full || c * 32 + i < P.tm))
                         ? fast_ex2(fmaf(__uint_as_float(v[i]), c2, -mc))
                         : 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              sum0 += p[i]; sum1 += p[i + 1]; sum2 += p[i + 2]; sum3 += p[i + 3];
            }
#pragma unroll
            for (int q8 = 0; q8 < 4; ++q8) {
              if (q8 * 8 < nv) {
                uint4 u;
                u.x = pack2(p[q8 * 8 + 0], p[q8 * 8 + 1], P.fmt);
                u.y = pack2(p[q8 * 8 + 2], p[q8 * 8 + 3], P.fmt);
                u.z = pack2(p[q8 * 8 + 4], p[q8 * 8 + 5], P.fmt);
                u.w = pack2(p[q8 * 8 + 6], p[q8 * 8 + 7], P.fmt);
                *reinterpret_cast<uint4*>(my_p + p_off(r, c * 32 + q8 * 8)) = u;
              }
            }
          }
          mx = mc;  // keep in log2 units
        }
        const float rsum = (sum0 + sum1) + (sum2 + sum3) + p_t;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&bar_p[g]);
        mbar_wait(&bar_o[g], ph, 41);
        tc_fence_after();
        if (warp_valid) {
          uint32_t v0[32], v1[32];
          tmem_ld32(ts, v0);
          tmem_ld32(ts + 32, v1);
          tmem_ld_wait();
          if (valid) {
            const float inv = 1.f / rsum;
            const size_t grow = (size_t)b * T + row;
            uint4* dst = reinterpret_cast<uint4*>(P.o_out + grow * W + h * 64);
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8) {
              const uint32_t* sv = q8 < 4 ? &v0[q8 * 8] : &v1[(q8 - 4) * 8];
              float o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(sv[i]);
              if (P.tail) {
                const uint4 vu = *reinterpret_cast<const uint4*>(vt + q8 * 8);
                const uint32_t* va = &vu.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 y = unpack2(va[i], P.fmt);
                  o[2 * i] = fmaf(p_t, y.x, o[2 * i]);
                  o[2 * i + 1] = fmaf(p_t, y.y, o[2 * i + 1]);
                }
              }
              uint4 u;
              u.x = pack2(o[0] * inv, o[1] * inv, P.fmt);
              u.y = pack2(o[2] * inv, o[3] * inv, P.fmt);
              u.z = pack2(o[4] * inv, o[5] * inv, P.fmt);
              u.w = pack2(o[6] * inv, o[7] * inv, P.fmt);
              dst[q8] = u;
            }
            P.lse_out[((size_t)b * P.H + h) * T + row] = (mx + log2f(rsum)) * 0.6931471805599453f;
          }
        }
        tc_fence_before();
        mbar_arrive(&bar_ofree[g]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AT_CTRL_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

struct AttnPlan {
  AttnParams P;
  size_t smem;
  CUtensorMap tmTile, tmFull, tmDoTile, tmDoFull;
};

static int make_plan(AttnPlan& pl, const void* qkv, const void* d_o, int B, int T, int H, int fmt,
                     int fmt_o, int fmt_g) {
  if (B <= 0 || T <= 0 || T > 272 || H <= 0) return RVLM_ERR_UNSUPPORTED;
  AttnParams& P = pl.P;
  P.B = B; P.T = T; P.H = H; P.W = H * 64;
  P.kvp = (T + 15) / 16 * 16;
  P.nblk = (P.kvp + 63) / 64;
  P.ntile = (T + 127) / 128;
  P.nbox = P.kvp > 256 ? 2 : 1;
  P.boxrows = P.kvp / P.nbox;
  P.fmt = fmt; P.fmt_o = fmt_o; P.fmt_g = fmt_g;
  pl.smem = (size_t)2 * P.kvp * 128 + 4 * 16384 + (size_t)P.nblk * 16384 + (size_t)2 * P.kvp * 4 +
            2048 + 128 + 1024;
  const uint64_t W3 = 3ull * P.W;
  int rc = make_tmap_3d(&pl.tmTile, qkv, W3, T, B, W3, (uint64_t)T * W3, 64, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&pl.tmFull, qkv, W3, T, B, W3, (uint64_t)T * W3, 64, P.boxrows);
  if (rc) return rc;
  const void* dptr = d_o ? d_o : qkv;  // unused maps still need a valid encoding
  const uint64_t Wd = d_o ? (uint64_t)P.W : W3;
  rc = make_tmap_3d(&pl.tmDoTile, dptr, Wd, T, B, Wd, (uint64_t)T * Wd, 64, 128);
  if (rc) return rc;
  return make_tmap_3d(&pl.tmDoFull, dptr, Wd, T, B, Wd, (uint64_t)T * Wd, 64, P.boxrows);
}

template <int MODE>
static int launch_attn(const AttnPlan& pl, cudaStream_t st) {
  auto kern = attn_tc_kernel<MODE>;
  static size_t set = 0;
  if (pl.smem > set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pl.smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    set = pl.smem;
  }
  int grid = sm_count();
  if (pl.P.B * pl.P.H < grid) grid = pl.P.B * pl.P.H;
  kern<<<grid, AT_THREADS, pl.smem, st>>>(pl.tmTile, pl.tmFull, pl.tmDoTile, pl.tmDoFull, pl.P);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

}  // namespace rvlm

using namespace rvlm;

extern "C" int rvlm_attn_fwd_simt_rows(const void* qkv16, void* o16, float* lse, int B, int T,
                                       int H, int fmt, int t_begin, int t_end, void* stream);

extern "C" int rvlm_attn_fwd(const void* qkv16, void* o16, float* lse, int B, int T, int H,
                             int fmt, void* stream) {
  if (B <= 0 || T <= 0 || T > 257 || H <= 0) return RVLM_ERR_UNSUPPORTED;
  AttnFwd2Params P{};
  P.B = B; P.T = T; P.H = H; P.W = H * 64;
  P.tm = T < 256 ? T : 256;
  P.kvm = (P.tm + 15) / 16 * 16;
  P.nblk = (P.kvm + 63) / 64;
  P.ntile = (P.tm + 127) / 128;
  P.tail = T > 256 ? 1 : 0;
  P.fmt = fmt;
  P.qkv = reinterpret_cast<const uint16_t*>(qkv16);
  P.o_out = reinterpret_cast<uint16_t*>(o16);
  P.lse_out = lse;
  CUtensorMap tmTile, tmFull;
  const uint64_t W3 = 3ull * P.W;
  int rc = make_tmap_3d(&tmTile, qkv16, W3, T, B, W3, (uint64_t)T * W3, 64, 128);
  if (rc) return rc;
  rc = make_tmap_3d(&tmFull, qkv16, W3, T, B, W3, (uint64_t)T * W3, 64, P.kvm);
  if (rc) return rc;
  const size_t smem = (size_t)2 * P.kvm * 128 + 2 * 16384 + (size_t)2 * P.nblk * 16384 + 512 +
                      128 + 128 + 1024;
  static size_t set = 0;
  if (smem > set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd2_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    set = smem;
  }
  int grid = sm_count();
  if (B * H < grid) grid = B * H;
  attn_fwd2_kernel<<<grid, AT_THREADS, smem, (cudaStream_t)stream>>>(tmTile, tmFull, P);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

extern "C" int rvlm_attn_bwd(const void* qkv16, const void* o16, const void* do16,
                             float* lse_tmp, void* dqkv16, int B, int T, int H, int fmt_qkv,
                             int fmt_o, int fmt_g, void* stream) {
  // every MMA operand of the backward (qkv copy, dO, P, dS) must share one 16-bit format
  if (fmt_qkv != fmt_g) return RVLM_ERR_UNSUPPORTED;
  if (!lse_tmp) return RVLM_ERR_INVAL;
  AttnPlan pl{};
  int rc = make_plan(pl, qkv16, do16, B, T, H, fmt_g, fmt_o, fmt_g);
  if (rc) return rc;
  pl.P.o = reinterpret_cast<const uint16_t*>(o16);
  pl.P.d_o = reinterpret_cast<const uint16_t*>(do16);
  pl.P.lse_tmp = lse_tmp;
  pl.P.dqkv = reinterpret_cast<uint16_t*>(dqkv16);
  rc = launch_attn<AT_DQ>(pl, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_attn<AT_DKV>(pl, (cudaStream_t)stream);
}
