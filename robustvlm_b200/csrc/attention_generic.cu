// tcgen05 attention kernels for GENERIC token counts (everything that is not the ViT-L/14
// T = 256 / 257 case served by attention.cu): ViT-B/32 (T = 50), ViT-B/16 (T = 197), the test
// shapes.  Three kernels:
//   attn_fwd2_kernel   forward, two 128-row tiles of one (image, head) on two warp groups
//   attn_bwd2_kernel   backward as a dQ kernel + a dK/dV kernel (kvp in {64, 256})
//   attn_tc_kernel     the first-generation one-skeleton kernel (DQ / DKV modes), kept for token
//                      counts the others do not tile (kvp not in {64, 256}, e.g. T = 197)
// Selected by rvlm_attn_fwd / rvlm_attn_bwd in attention.cu.
#include "attention_common.cuh"

namespace rvlm {

template <int MODE>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmTile,   // qkv, box 64 x 128
               const __grid_constant__ CUtensorMap tmFull,   // qkv, box 64 x boxrows
               const __grid_constant__ CUtensorMap tmDoTile, // dO,  box 64 x 128
               const __grid_constant__ CUtensorMap tmDoFull, // dO,  box 64 x boxrows
               AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array, not an integer round trip: keeps the
  // address space visible to the compiler -> LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
template <int FMT>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmTile,   // qkv, box 64 x 128
                 const __grid_constant__ CUtensorMap tmFull,   // qkv, box 64 x kvm
                 AttnFwd2Params P) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array, not an integer round trip: keeps the
  // address space visible to the compiler -> LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int kvm = P.kvm, T = P.T, W = P.W;
  const int full_bytes = kvm * 128;
  uint8_t* s_k = smem;
  uint8_t* s_v = s_k + full_bytes;
  uint8_t* s_q = s_v + full_bytes;                 // [2][16384]
  uint8_t* s_p = s_q + 2 * 16384;                  // [2][nblk * 16384]
  uint16_t* s_tail = reinterpret_cast<uint16_t*>(s_p + 2 * P.nblk * 16384);  // [2][2][64]
  uint16_t* s_qtail = s_tail + 256;                // [2][64] the peeled query row, per group
  float* s_part = reinterpret_cast<float*>(s_qtail + 128);   // [2][68] group 0's half of that row
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 136);
  uint64_t* bar_k = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_q = bars + 2;       // [2]
  uint64_t* bar_s = bars + 4;       // [2]
  uint64_t* bar_p = bars + 6;       // [2] 128 arrivals
  uint64_t* bar_o = bars + 8;       // [2]
  uint64_t* bar_ofree = bars + 10;  // [2] 128 arrivals
  uint64_t* bar_tail = bars + 12;   // 256 arrivals: both halves of the peeled query row done with K / V
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_tail, 256);
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
  const uint32_t f = (uint32_t)FMT;
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
          // NOTE: single-buffered -- a warp that runs ahead into the next item rewrites these
          // rows while a slower warp of the group may still read v_tail in its epilogue; the
          // global-load latency in front of the store has always covered the skew (results are
          // bit-stable run to run), but attn_fwd3 double-buffers them by item parity instead.
          if (r < 16) {  // this group's copy of the peeled key / value row (index 256)
            const uint16_t* src = P.qkv + ((size_t)b * T + 256) * 3 * W + h * 64 +
                                  (r < 8 ? W : 2 * W) + (r & 7) * 8;
            *reinterpret_cast<uint4*>(s_tail + g * 128 + r * 8) =
                *reinterpret_cast<const uint4*>(src);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        }
        if (P.tail) {
          // ---- the peeled QUERY row (token 256) on the CUDA cores, in the shadow of the S
          // MMAs, SPLIT between the two row groups (T = 257 always runs both): group g scores
          // keys [g*128, g*128+128) (the last group also key 256), forms its own softmax
          // partials (max, sum) and partial P V; group 0 hands {o[64], max, sum} to group 1
          // through a double-buffered 68-float slot (bar.arrive / bar.sync on barrier 3), and
          // group 1 merges the two halves flash-style and writes the row.  Scratch = this
          // group's P buffer.
          float* sc = reinterpret_cast<float*>(my_p);   // [129] this group's scores / probabilities
          float* red = sc + 136;                        // [8] warp partials
          float* oh = red + 8;                          // [2][64] output halves (two key quarters)
          float* part = s_part + (it & 1) * 68;         // group 0 -> group 1
          uint16_t* qt = s_qtail + g * 64;              // this group's copy of the peeled query
          if (r < 8)
            *reinterpret_cast<uint4*>(qt + r * 8) = *reinterpret_cast<const uint4*>(
                P.qkv + ((size_t)b * T + 256) * 3 * W + h * 64 + r * 8);
          mbar_wait(bar_k, ph, 43);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          float qv[64];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 qu = *reinterpret_cast<const uint4*>(qt + c * 8);
            const uint32_t* qa = &qu.x;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 x = unpackc<FMT>(qa[i]);
              qv[c * 8 + 2 * i] = x.x;
              qv[c * 8 + 2 * i + 1] = x.y;
            }
          }
          float lmax;
          {
            const int kv = g * 128 + r;
            float acc = 0.f;
            const uint8_t* krow = s_k + kv * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 ku = *reinterpret_cast<const uint4*>(krow + ((c ^ (kv & 7)) << 4));
              const uint32_t* ka = &ku.x;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 y = unpackc<FMT>(ka[i]);
                acc = fmaf(qv[c * 8 + 2 * i], y.x, fmaf(qv[c * 8 + 2 * i + 1], y.y, acc));
              }
            }
            sc[r] = acc;
            lmax = acc;
          }
          if (g == 1 && r == 127) {   // key 256 (this group's k_tail copy)
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 ku = *reinterpret_cast<const uint4*>(kt + c * 8);
              const uint32_t* ka = &ku.x;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 y = unpackc<FMT>(ka[i]);
                acc = fmaf(qv[c * 8 + 2 * i], y.x, fmaf(qv[c * 8 + 2 * i + 1], y.y, acc));
              }
            }
            sc[128] = acc;
            lmax = fmaxf(lmax, acc);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffff, lmax, o));
          if (lane == 0) red[warp & 3] = lmax;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          const float tmc = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * c2;
          float lsum;
          {
            const float e = fast_ex2(fmaf(sc[r], c2, -tmc));
            sc[r] = e;
            lsum = e;
          }
          if (g == 1 && r == 127) {
            const float e = fast_ex2(fmaf(sc[128], c2, -tmc));
            sc[128] = e;
            lsum += e;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffff, lsum, o);
          if (lane == 0) red[4 + (warp & 3)] = lsum;
          mbar_wait(bar_v, ph, 44);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          const float tsum = (red[4] + red[5]) + (red[6] + red[7]);
          {  // o[d] over one quarter of the keys (this thread's half of the group's keys)
            const int d = r & 63, hf = r >> 6;
            const int j0 = hf * 64;
            const uint8_t* vbase = s_v + (size_t)(g * 128) * 128 + (d & 7) * 2;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
            for (int j = j0; j < j0 + 64; j += 2) {
              const uint16_t h0 = *reinterpret_cast<const uint16_t*>(
                  vbase + j * 128 + (((d >> 3) ^ (j & 7)) << 4));
              const uint16_t h1 = *reinterpret_cast<const uint16_t*>(
                  vbase + (j + 1) * 128 + (((d >> 3) ^ ((j + 1) & 7)) << 4));
              acc0 = fmaf(sc[j], ld1<FMT>(&h0), acc0);
              acc1 = fmaf(sc[j + 1], ld1<FMT>(&h1), acc1);
            }
            float acc = acc0 + acc1;
            if (g == 1 && hf == 1) {
              const uint16_t hv = vt[d];
              acc = fmaf(sc[128], ld1<FMT>(&hv), acc);
            }
            oh[hf * 64 + d] = acc;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          if (g == 0) {
            if (r < 64) part[r] = oh[r] + oh[64 + r];
            if (r == 64) part[64] = tmc;
            if (r == 65) part[65] = tsum;
            __threadfence_block();
            asm volatile("bar.arrive 3, 256;" ::: "memory");
          } else {
            asm volatile("bar.sync 3, 256;" ::: "memory");
            if (r < 32) {
              const float m0 = part[64], l0 = part[65];
              const float mm = fmaxf(m0, tmc);
              const float a0 = fast_ex2(m0 - mm), a1 = fast_ex2(tmc - mm);
              const float ltot = fmaf(l0, a0, tsum * a1);
              const float inv = 1.f / ltot;
              const float o0 = fmaf(part[2 * r], a0, (oh[2 * r] + oh[64 + 2 * r]) * a1) * inv;
              const float o1 =
                  fmaf(part[2 * r + 1], a0, (oh[2 * r + 1] + oh[64 + 2 * r + 1]) * a1) * inv;
              reinterpret_cast<uint32_t*>(P.o_out + ((size_t)b * T + 256) * W + h * 64)[r] =
                  packc<FMT>(o0, o1);
              if (r == 0)
                P.lse_out[((size_t)b * P.H + h) * T + 256] =
                    (mm + log2f(ltot)) * 0.6931471805599453f;
            }
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
                const float2 x = unpackc<FMT>(qa[i]), y = unpackc<FMT>(ka[i]);
                s_t = fmaf(x.x, y.x, fmaf(x.y, y.y, s_t));
              }
            }
            mx = s_t;
          }
          // TMEM sweeps in 16-column pieces with the next tcgen05.ld in flight
          const int nq = kvm >> 4;
          const bool allcols = P.tm == kvm;          // ViT-L/14: every column valid
          uint8_t* prow = my_p + r * 128;
          const uint32_t rx = (uint32_t)(r & 7) << 4;
          {
            float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            auto pmax = [&](const uint32_t (&v)[16], int q) {
              if (allcols || q * 16 + 16 <= P.tm) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  m0 = fmaxf(m0, __uint_as_float(v[i]));
                  m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                  m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
                  m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (q * 16 + i < P.tm) m0 = fmaxf(m0, __uint_as_float(v[i]));
              }
            };
            uint32_t va[16], vb[16];
            tmem_ld16(ts, va);
            for (int q = 0; q < nq; q += 2) {
              tmem_ld_wait();
              if (q + 1 < nq) tmem_ld16(ts + (q + 1) * 16, vb);
              pmax(va, q);
              if (q + 1 < nq) {
                tmem_ld_wait();
                if (q + 2 < nq) tmem_ld16(ts + (q + 2) * 16, va);
                pmax(vb, q + 1);
              }
            }
            mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          }
          const float mc = mx * c2;
          if (P.tail) p_t = fast_ex2(fmaf(s_t, c2, -mc));
          {
            auto pexp = [&](const uint32_t (&v)[16], int q) {
              float e[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) e[i] = fast_ex2(fmaf(__uint_as_float(v[i]), c2, -mc));
              if (!(allcols || q * 16 + 16 <= P.tm)) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (q * 16 + i >= P.tm) e[i] = 0.f;
              }
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                sum0 += e[i]; sum1 += e[i + 1]; sum2 += e[i + 2]; sum3 += e[i + 3];
              }
#pragma unroll
              for (int q8 = 0; q8 < 2; ++q8) {
                const int cc = q * 16 + q8 * 8;
                *reinterpret_cast<uint4*>(prow + (cc >> 6) * 16384 +
                                          (((uint32_t)((cc & 63) >> 3) << 4) ^ rx)) =
                    make_uint4(packc<FMT>(e[q8 * 8 + 0], e[q8 * 8 + 1]),
                               packc<FMT>(e[q8 * 8 + 2], e[q8 * 8 + 3]),
                               packc<FMT>(e[q8 * 8 + 4], e[q8 * 8 + 5]),
                               packc<FMT>(e[q8 * 8 + 6], e[q8 * 8 + 7]));
              }
            };
            uint32_t va[16], vb[16];
            tmem_ld16(ts, va);
            for (int q = 0; q < nq; q += 2) {
              tmem_ld_wait();
              if (q + 1 < nq) tmem_ld16(ts + (q + 1) * 16, vb);
              pexp(va, q);
              if (q + 1 < nq) {
                tmem_ld_wait();
                if (q + 2 < nq) tmem_ld16(ts + (q + 2) * 16, va);
                pexp(vb, q + 1);
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
                  const float2 y = unpackc<FMT>(va[i]);
                  o[2 * i] = fmaf(p_t, y.x, o[2 * i]);
                  o[2 * i + 1] = fmaf(p_t, y.y, o[2 * i + 1]);
                }
              }
              uint4 u;
              u.x = packc<FMT>(o[0] * inv, o[1] * inv);
              u.y = packc<FMT>(o[2] * inv, o[3] * inv);
              u.z = packc<FMT>(o[4] * inv, o[5] * inv);
              u.w = packc<FMT>(o[6] * inv, o[7] * inv);
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


// =====================================================================================
// Backward, version 2 (ViT-L/14: T = 257, ViT-B/32: T = 50).
//   * the 257th token is peeled exactly as in the forward: tensor-core tiles cover rows and
//     columns [0, 256); the peeled key / query COLUMN is a pair of 64-MAC dot products per
//     thread-row folded into the epilogue as a rank-1 update, and the peeled ROW is computed
//     by all worker threads on the CUDA cores in the shadow of the first tile's MMAs;
//   * S and dP are issued back to back into separate TMEM regions (2 x 256 columns), the
//     64-column output accumulators alias the consumed S region: three issuer<->worker
//     hand-offs per tile instead of five, two tiles per item instead of three;
//   * each thread keeps its 32*NCH probabilities in registers between the two passes (one
//     TMEM sweep of S, one of dP); operand format is a template parameter.
// =====================================================================================
template <int MODE, int NCH, int FMT, bool TAIL, bool FULL>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmTile,   // qkv, box 64 x 128
                 const __grid_constant__ CUtensorMap tmFull,   // qkv, box 64 x kvm
                 const __grid_constant__ CUtensorMap tmDoTile, // dO,  box 64 x 128
                 const __grid_constant__ CUtensorMap tmDoFull, // dO,  box 64 x kvm
                 AttnBwd2Params P) {
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array, not an integer round trip: keeps the
  // address space visible to the compiler -> LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int KVM = NCH * 64;          // columns on the tensor cores (64 or 256)
  constexpr int NBLK = KVM / 64;
  const int T = P.T, W = P.W;
  constexpr int full_bytes = KVM * 128;
  uint8_t* s_full0 = smem;                   // DQ: K      DKV: Q
  uint8_t* s_full1 = s_full0 + full_bytes;   // DQ: V      DKV: dO
  uint8_t* s_tile0 = s_full1 + full_bytes;   // [2][16384] DQ: Q_i   DKV: K_j
  uint8_t* s_tile1 = s_tile0 + 2 * 16384;    // [2][16384] DQ: dO_i  DKV: V_j
  uint8_t* s_p = s_tile1 + 2 * 16384;        // [NBLK][16384] P / dS operand
  float* s_lse = reinterpret_cast<float*>(s_p + NBLK * 16384);  // [272] (DKV) per-q lse
  float* s_dvec = s_lse + 272;                                  // [272] (DKV) per-q D
  float* s_red = s_dvec + 272;    // [256]
  float* s_red2 = s_red + 256;    // [256]
  float* s_tt = s_red2 + 256;     // [256] per-row tail terms: p_t | ds_t
  float* s_misc = s_tt + 256;     // [16]
  uint16_t* s_tv = reinterpret_cast<uint16_t*>(s_misc + 16);  // [4][64] peeled vectors
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tv + 256);
  uint64_t* bar_full = bars + 0;
  uint64_t* bar_tile = bars + 1;      // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_dp = bars + 4;
  uint64_t* bar_p = bars + 5;         // 256 arrivals, 1 (DQ) or 2 (DKV) uses per tile
  uint64_t* bar_dv = bars + 6;        // DKV: dV MMA retired (P may be overwritten)
  uint64_t* bar_acc = bars + 7;
  uint64_t* bar_acc_free = bars + 8;  // 256 arrivals
  uint64_t* bar_trow = bars + 9;      // 256 arrivals: peeled row done with the full matrices
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(&bar_tile[0], 1);
    mbar_init(&bar_tile[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_dp, 1);
    mbar_init(bar_p, 256);
    mbar_init(bar_dv, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc_free, 256);
    mbar_init(bar_trow, 256);
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
  const int colQ = 0, colK = W, colV = 2 * W;
  constexpr uint32_t TS = 0, TDP = 256, TACC = 0, TACC2 = 64;

  if (warp == AT_CTRL_WARP) {
    if (lane == 0) {
      // ===================== TMA + MMA issuer =====================
      auto load_full = [&](int item) {
        const int b = item / P.H, h = item % P.H;
        mbar_arrive_expect_tx(bar_full, 2 * full_bytes);
        if (MODE == AT_DKV) {
          tma_load_3d(s_full0, &tmFull, bar_full, colQ + h * 64, 0, b);
          tma_load_3d(s_full1, &tmDoFull, bar_full, h * 64, 0, b);
        } else {
          tma_load_3d(s_full0, &tmFull, bar_full, colK + h * 64, 0, b);
          tma_load_3d(s_full1, &tmFull, bar_full, colV + h * 64, 0, b);
        }
      };
      auto load_tile = [&](int item, int j, uint32_t gg) {
        const int b = item / P.H, h = item % P.H, buf = gg & 1;
        mbar_arrive_expect_tx(&bar_tile[buf], 32768);
        if (MODE == AT_DKV) {
          tma_load_3d(s_tile0 + buf * 16384, &tmTile, &bar_tile[buf], colK + h * 64, j * 128, b);
          tma_load_3d(s_tile1 + buf * 16384, &tmTile, &bar_tile[buf], colV + h * 64, j * 128, b);
        } else {
          tma_load_3d(s_tile0 + buf * 16384, &tmTile, &bar_tile[buf], colQ + h * 64, j * 128, b);
          tma_load_3d(s_tile1 + buf * 16384, &tmDoTile, &bar_tile[buf], h * 64, j * 128, b);
        }
      };
      const uint32_t idS = idesc_ab(FMT, FMT, 128, KVM, 0, 0);
      uint32_t g = 0;
      int it = 0;
      if ((int)blockIdx.x < n_items) {
        load_full(blockIdx.x);
        load_tile(blockIdx.x, 0, 0);
      }
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int next = item + gridDim.x;
        mbar_wait(bar_full, it & 1, 50);
        for (int j = 0; j < P.ntile; ++j, ++g) {
          const int buf = g & 1;
          if (j + 1 < P.ntile) load_tile(item, j + 1, g + 1);
          else if (next < n_items) load_tile(next, 0, g + 1);
          mbar_wait(&bar_tile[buf], (g >> 1) & 1, 51);
          if (g > 0) mbar_wait(bar_acc_free, (g - 1) & 1, 52);   // accumulators alias S
          tc_fence_after();
          const uint32_t t0 = smem_u32(s_tile0 + buf * 16384), t1 = smem_u32(s_tile1 + buf * 16384);
          const uint32_t f0 = smem_u32(s_full0), f1 = smem_u32(s_full1), pa = smem_u32(s_p);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // S = Q_i K^T   |  S^T = K_j Q^T
            umma_f16(tmem + TS, make_smem_desc_sw128(t0 + k * 32, 16, 1024),
                     make_smem_desc_sw128(f0 + k * 32, 16, 1024), idS, k > 0);
          umma_commit(bar_s);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // dP = dO_i V^T |  dP^T = V_j dO^T
            umma_f16(tmem + TDP, make_smem_desc_sw128(t1 + k * 32, 16, 1024),
                     make_smem_desc_sw128(f1 + k * 32, 16, 1024), idS, k > 0);
          umma_commit(bar_dp);
          if (MODE == AT_DKV) {
            mbar_wait(bar_p, (2 * g) & 1, 53);                   // P^T written, S^T consumed
            tc_fence_after();
            issue_mma2(tmem + TACC, pa, f1, KVM, FMT, FMT);      // dV = P^T dO
            umma_commit(bar_dv);
            mbar_wait(bar_p, (2 * g + 1) & 1, 54);               // dS^T written
            tc_fence_after();
            issue_mma2(tmem + TACC2, pa, f0, KVM, FMT, FMT);     // dK = dS^T Q
          } else {
            mbar_wait(bar_p, g & 1, 53);                         // dS written, S / dP consumed
            tc_fence_after();
            issue_mma2(tmem + TACC, pa, f0, KVM, FMT, FMT);      // dQ = dS K
          }
          umma_commit(bar_acc);
        }
        mbar_wait(bar_acc, (g - 1) & 1, 55);                     // every MMA of this item retired
        if (TAIL) mbar_wait(bar_trow, it & 1, 56);
        if (next < n_items) load_full(next);
      }
    }
  } else {
    // ===================== row workers =====================
    const int r = threadIdx.x & 127, hsel = threadIdx.x >> 7, tid = threadIdx.x;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const float c2 = 0.125f * 1.4426950408889634f;
    const uint32_t ts = tmem + lane_off;
    constexpr int NH = NCH;                 // 32-column chunks per thread (half of the columns)
    const int c0 = hsel * NH;
    const uint16_t* tv0 = s_tv;             // DQ: k_256   DKV: q_256
    const uint16_t* tv1 = s_tv + 64;        // DQ: v_256   DKV: dO_256
    const uint16_t* tv2 = s_tv + 128;       // DQ: q_256   DKV: k_256
    const uint16_t* tv3 = s_tv + 192;       // DQ: dO_256  DKV: v_256
    uint32_t g = 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / P.H, h = item % P.H;
      const size_t lse_base = ((size_t)b * P.H + h) * T;
      // ---- stage the peeled vectors and (DKV) the per-query statistics
      if (TAIL && tid < 32) {
        const int which = tid >> 3, c = tid & 7;
        const size_t rowq = ((size_t)b * T + 256) * 3 * W + h * 64 + c * 8;
        const size_t rowd = ((size_t)b * T + 256) * W + h * 64 + c * 8;
        const uint16_t* src;
        if (MODE == AT_DQ)
          src = which == 0 ? P.qkv + rowq + colK : which == 1 ? P.qkv + rowq + colV
              : which == 2 ? P.qkv + rowq + colQ : P.d_o + rowd;
        else
          src = which == 0 ? P.qkv + rowq + colQ : which == 1 ? P.d_o + rowd
              : which == 2 ? P.qkv + rowq + colK : P.qkv + rowq + colV;
        *reinterpret_cast<uint4*>(s_tv + which * 64 + c * 8) = *reinterpret_cast<const uint4*>(src);
      }
      if (MODE == AT_DKV) {
        for (int q = tid; q < 272; q += 256) {
          float L = INFINITY, Dv = 0.f;
          if (q < T) {
            L = P.lse_tmp[lse_base + q];
            const size_t ro = ((size_t)b * T + q) * W + h * 64;
            Dv = dot_rows16(P.o + ro, P.fmt_o, P.d_o + ro, FMT);
          }
          s_lse[q] = L;
          s_dvec[q] = Dv * 0.125f;   // pre-scaled: dS = P (dP/8 - D/8)
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (TAIL) {
        // ================= the peeled ROW (token 256) on the CUDA cores =================
        float* sc = reinterpret_cast<float*>(s_p);   // [272] p / weights (scratch: P buffer)
        float* sc2 = sc + 272;                       // [272] ds
        float* part = sc2 + 272;                     // [2][4][64] partial sums
        mbar_wait(bar_full, it & 1, 60);
        if (MODE == AT_DQ) {
          // dq_256 = sum_kv ds[kv] K[kv],  ds = softmax(q_256 K^T / 8) (dO_256 V^T - D) / 8
          float sv = -INFINITY, dpv = 0.f;
          if (tid < 256) {
            sv = dot_swz_row<FMT>(s_full0, tid, tv2);
            dpv = dot_swz_row<FMT>(s_full1, tid, tv3);
          }
          float s_last = dot_vec<FMT>(tv0, tv2), dp_last = dot_vec<FMT>(tv1, tv3);
          const float mx = blk256_max(fmaxf(sv, s_last), s_misc, warp, lane) * c2;
          const float e = fast_ex2(fmaf(sv, c2, -mx));
          const float e_last = fast_ex2(fmaf(s_last, c2, -mx));
          const float sum = blk256_sum(e, s_misc, warp, lane) + e_last;
          const float inv = 1.f / sum;
          const size_t ro = ((size_t)b * T + 256) * W + h * 64;
          const float D = dot_rows16(P.o + ro, P.fmt_o, P.d_o + ro, FMT);
          sc2[tid] = e * inv * (dpv - D) * 0.125f;
          if (tid == 0) {
            sc2[256] = e_last * inv * (dp_last - D) * 0.125f;
            P.lse_tmp[lse_base + 256] = mx + log2f(sum);
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          {
            const int d = tid & 63, qt = tid >> 6;
            float acc = 0.f;
            for (int kv = qt * 64; kv < qt * 64 + 64; ++kv)
              acc = fmaf(sc2[kv], swz_at<FMT>(s_full0, kv, d), acc);
            if (qt == 3) acc = fmaf(sc2[256], ld1<FMT>(tv0 + d), acc);
            part[qt * 64 + d] = acc;
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          if (tid < 32) {
            const float a0 = part[2 * tid] + part[64 + 2 * tid] + part[128 + 2 * tid] +
                             part[192 + 2 * tid];
            const float a1 = part[2 * tid + 1] + part[64 + 2 * tid + 1] + part[128 + 2 * tid + 1] +
                             part[192 + 2 * tid + 1];
            reinterpret_cast<uint32_t*>(P.dqkv + ((size_t)b * T + 256) * 3 * W + colQ + h * 64)[tid] =
                packc<FMT>(a0, a1);
          }
        } else {
          // dv_256 = sum_q p[q] dO[q],  dk_256 = sum_q ds[q] Q[q]   (column 256 of P, dS)
          float pv = 0.f, dsv = 0.f;
          {
            const float sT = dot_swz_row<FMT>(s_full0, tid, tv2);   // q_tid . k_256
            const float dpT = dot_swz_row<FMT>(s_full1, tid, tv3);  // dO_tid . v_256
            pv = fast_ex2(fmaf(sT, c2, -s_lse[tid]));
            dsv = pv * fmaf(dpT, 0.125f, -s_dvec[tid]);
          }
          sc[tid] = pv;
          sc2[tid] = dsv;
          if (tid == 0) {
            const float sT = dot_vec<FMT>(tv0, tv2), dpT = dot_vec<FMT>(tv1, tv3);
            const float pl = fast_ex2(fmaf(sT, c2, -s_lse[256]));
            sc[256] = pl;
            sc2[256] = pl * fmaf(dpT, 0.125f, -s_dvec[256]);
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          {
            const int d = tid & 63, qt = tid >> 6;
            float av = 0.f, ak = 0.f;
            for (int q = qt * 64; q < qt * 64 + 64; ++q) {
              av = fmaf(sc[q], swz_at<FMT>(s_full1, q, d), av);
              ak = fmaf(sc2[q], swz_at<FMT>(s_full0, q, d), ak);
            }
            if (qt == 3) {
              av = fmaf(sc[256], ld1<FMT>(tv1 + d), av);
              ak = fmaf(sc2[256], ld1<FMT>(tv0 + d), ak);
            }
            part[qt * 64 + d] = av;
            part[256 + qt * 64 + d] = ak;
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
          if (tid < 64) {
            const int which = tid >> 5, l2 = tid & 31;   // 0: dv, 1: dk
            const float* pp = part + which * 256;
            const float a0 = pp[2 * l2] + pp[64 + 2 * l2] + pp[128 + 2 * l2] + pp[192 + 2 * l2];
            const float a1 = pp[2 * l2 + 1] + pp[64 + 2 * l2 + 1] + pp[128 + 2 * l2 + 1] +
                             pp[192 + 2 * l2 + 1];
            reinterpret_cast<uint32_t*>(P.dqkv + ((size_t)b * T + 256) * 3 * W +
                                        (which ? colK : colV) + h * 64)[l2] = packc<FMT>(a0, a1);
          }
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");   // scratch (P buffer) free again
        mbar_arrive(bar_trow);
      }
      for (int j = 0; j < P.ntile; ++j, ++g) {
        const int buf = g & 1;
        const int row = j * 128 + r;
        const bool valid = FULL || row < P.tm;
        const bool warp_valid = FULL || (j * 128 + (warp & 3) * 32) < P.tm;
        const size_t grow = (size_t)b * T + row;
        mbar_wait(&bar_tile[buf], (g >> 1) & 1, 61);     // tile rows visible to generic loads
        float Dr = 0.f, a_t = 0.f, b_t = 0.f;
        if (MODE == AT_DQ && valid)
          Dr = dot_rows16(P.o + grow * W + h * 64, P.fmt_o, P.d_o + grow * W + h * 64, FMT);
        if (TAIL && hsel == 1 && warp_valid) {
          a_t = dot_swz_row<FMT>(s_tile0 + buf * 16384, r, tv0);   // DQ: q_r.k_256  DKV: k_r.q_256
          b_t = dot_swz_row<FMT>(s_tile1 + buf * 16384, r, tv1);   // DQ: dO_r.v_256 DKV: v_r.dO_256
        }
        // ---------------- pass 1: probabilities ----------------
        // TMEM is swept in 16-column pieces, the next piece's tcgen05.ld in flight while the
        // current one is processed.  Probabilities stay PACKED (operand format) in registers
        // between the passes: pass 2 uses exactly the values the tensor core sees in P.
        mbar_wait(bar_s, g & 1, 62);
        tc_fence_after();
        constexpr int NQ = NH * 2;                 // 16-column pieces owned by this thread
        const uint32_t tcol = ts + c0 * 32;        // first TMEM column of this thread's half
        const int col0 = c0 * 32;
        uint8_t* prow = s_p + (size_t)(col0 >> 6) * 16384 + r * 128;   // this row, first block
        const uint32_t rx = (uint32_t)(r & 7) << 4;
        // byte offset of 8 columns starting at (thread-relative) column cc (cc % 8 == 0)
        auto poff = [&](int cc) -> uint32_t {
          return (uint32_t)(((col0 & 63) + cc) >> 6) * 16384u +
                 ((((uint32_t)(((col0 & 63) + cc) & 63) >> 3) << 4) ^ rx);
        };
        uint32_t pk[NQ * 8];
        float mc = 0.f, rinv = 1.f, p_t = 0.f;
        if (MODE == AT_DQ) {
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          if (warp_valid) {
            uint32_t v[2][16];
            tmem_ld16(tcol + TS, v[0]);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              tmem_ld_wait();
              if (q + 1 < NQ) tmem_ld16(tcol + TS + (q + 1) * 16, v[(q + 1) & 1]);
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                if (FULL || col0 + q * 16 + i + 3 < P.tm) {
                  m0 = fmaxf(m0, __uint_as_float(v[q & 1][i]));
                  m1 = fmaxf(m1, __uint_as_float(v[q & 1][i + 1]));
                  m2 = fmaxf(m2, __uint_as_float(v[q & 1][i + 2]));
                  m3 = fmaxf(m3, __uint_as_float(v[q & 1][i + 3]));
                } else {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    if (col0 + q * 16 + i + u < P.tm) m0 = fmaxf(m0, __uint_as_float(v[q & 1][i + u]));
                }
              }
            }
            if (TAIL && hsel == 1) m0 = fmaxf(m0, a_t);
          }
          s_red[hsel * 128 + r] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          asm volatile("bar.sync 2, 256;" ::: "memory");
          mc = fmaxf(s_red[r], s_red[128 + r]) * c2;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          if (warp_valid) {
            uint32_t v[2][16];
            tmem_ld16(tcol + TS, v[0]);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              tmem_ld_wait();
              if (q + 1 < NQ) tmem_ld16(tcol + TS + (q + 1) * 16, v[(q + 1) & 1]);
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                float e[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  e[u] = fast_ex2(fmaf(__uint_as_float(v[q & 1][i + u]), c2, -mc));
                  if (!FULL && !(valid && col0 + q * 16 + i + u < P.tm)) e[u] = 0.f;
                }
                s0 += e[0]; s1 += e[1]; s2 += e[2]; s3 += e[3];
                pk[q * 8 + i / 2] = packc<FMT>(e[0], e[1]);
                pk[q * 8 + i / 2 + 1] = packc<FMT>(e[2], e[3]);
              }
            }
            if (TAIL && hsel == 1 && valid) p_t = fast_ex2(fmaf(a_t, c2, -mc));
          }
          s_red2[hsel * 128 + r] = (s0 + s1) + (s2 + s3) + p_t;
          asm volatile("bar.sync 2, 256;" ::: "memory");
          const float rsum = s_red2[r] + s_red2[128 + r];
          rinv = 1.f / rsum;
          if (valid && hsel == 0) P.lse_tmp[lse_base + row] = mc + log2f(rsum);
        } else {
          if (warp_valid) {
            uint32_t v[2][16];
            tmem_ld16(tcol + TS, v[0]);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              tmem_ld_wait();
              if (q + 1 < NQ) tmem_ld16(tcol + TS + (q + 1) * 16, v[(q + 1) & 1]);
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 L4 = *reinterpret_cast<const float4*>(s_lse + col0 + q * 16 + i);
                float e0 = fast_ex2(fmaf(__uint_as_float(v[q & 1][i]), c2, -L4.x));
                float e1 = fast_ex2(fmaf(__uint_as_float(v[q & 1][i + 1]), c2, -L4.y));
                float e2 = fast_ex2(fmaf(__uint_as_float(v[q & 1][i + 2]), c2, -L4.z));
                float e3 = fast_ex2(fmaf(__uint_as_float(v[q & 1][i + 3]), c2, -L4.w));
                if (!FULL && !valid) { e0 = 0.f; e1 = 0.f; e2 = 0.f; e3 = 0.f; }
                pk[q * 8 + i / 2] = packc<FMT>(e0, e1);
                pk[q * 8 + i / 2 + 1] = packc<FMT>(e2, e3);
              }
              // P^T -> shared memory (A operand of dV = P^T dO)
#pragma unroll
              for (int q8 = 0; q8 < 2; ++q8)
                *reinterpret_cast<uint4*>(prow + poff(q * 16 + q8 * 8)) =
                    make_uint4(pk[q * 8 + q8 * 4], pk[q * 8 + q8 * 4 + 1], pk[q * 8 + q8 * 4 + 2],
                               pk[q * 8 + q8 * 4 + 3]);
            }
            if (TAIL && hsel == 1 && valid) p_t = fast_ex2(fmaf(a_t, c2, -s_lse[256]));
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(bar_p);
        }
        // ---------------- pass 2: dS = P (dP - D) / 8 ----------------
        // d = p * fma(dP, sc, -D*sc); DKV keeps D pre-multiplied by 1/8 in s_dvec.
        mbar_wait(bar_dp, g & 1, 63);
        tc_fence_after();
        if (MODE == AT_DKV) mbar_wait(bar_dv, g & 1, 64);   // dV no longer reads the P buffer
        float ds_t = 0.f;
        if (warp_valid) {
          const float sc = 0.125f * rinv;
          const float nDsc = -Dr * sc;
          uint32_t v[2][16];
          tmem_ld16(tcol + TDP, v[0]);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            tmem_ld_wait();
            if (q + 1 < NQ) tmem_ld16(tcol + TDP + (q + 1) * 16, v[(q + 1) & 1]);
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
              float d[8];
#pragma unroll
              for (int u = 0; u < 8; u += 4) {
                const int i = q8 * 8 + u;
                const float2 pa = unpackc<FMT>(pk[q * 8 + i / 2]);
                const float2 pb = unpackc<FMT>(pk[q * 8 + i / 2 + 1]);
                float4 nd = make_float4(nDsc, nDsc, nDsc, nDsc);
                if (MODE == AT_DKV) {
                  const float4 D4 = *reinterpret_cast<const float4*>(s_dvec + col0 + q * 16 + i);
                  nd = make_float4(-D4.x, -D4.y, -D4.z, -D4.w);
                }
                d[u] = pa.x * fmaf(__uint_as_float(v[q & 1][i]), sc, nd.x);
                d[u + 1] = pa.y * fmaf(__uint_as_float(v[q & 1][i + 1]), sc, nd.y);
                d[u + 2] = pb.x * fmaf(__uint_as_float(v[q & 1][i + 2]), sc, nd.z);
                d[u + 3] = pb.y * fmaf(__uint_as_float(v[q & 1][i + 3]), sc, nd.w);
              }
              *reinterpret_cast<uint4*>(prow + poff(q * 16 + q8 * 8)) =
                  make_uint4(packc<FMT>(d[0], d[1]), packc<FMT>(d[2], d[3]),
                             packc<FMT>(d[4], d[5]), packc<FMT>(d[6], d[7]));
            }
          }
          if (TAIL && hsel == 1) {
            const float nd = (MODE == AT_DQ) ? nDsc : -s_dvec[256];
            ds_t = p_t * fmaf(b_t, sc, nd);
            s_tt[r] = p_t * rinv;
            s_tt[128 + r] = ds_t;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_p);
        // ---------------- epilogue ----------------
        mbar_wait(bar_acc, g & 1, 65);
        tc_fence_after();
        if (warp_valid) {
          const float pt = TAIL ? s_tt[r] : 0.f, dst = TAIL ? s_tt[128 + r] : 0.f;
          if (MODE == AT_DKV) {
            uint32_t v0[32], v1[32];
            const uint32_t col = hsel ? TACC2 : TACC;
            tmem_ld32(ts + col, v0);
            tmem_ld32(ts + col + 32, v1);
            tmem_ld_wait();
            if (valid) {
              // hsel 0: dV_r += p_t dO_256     hsel 1: dK_r += ds_t q_256
              const float w = hsel ? dst : pt;
              const uint16_t* tvx = hsel ? tv0 : tv1;
              uint4* dst4 = reinterpret_cast<uint4*>(P.dqkv + grow * 3 * W + (hsel ? colK : colV) +
                                                     h * 64);
#pragma unroll
              for (int q8 = 0; q8 < 8; ++q8) {
                const uint32_t* sv = q8 < 4 ? &v0[q8 * 8] : &v1[(q8 - 4) * 8];
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(sv[i]);
                if (TAIL) {
                  const uint4 tu = *reinterpret_cast<const uint4*>(tvx + q8 * 8);
                  const uint32_t* ta = &tu.x;
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 y = unpackc<FMT>(ta[i]);
                    o[2 * i] = fmaf(w, y.x, o[2 * i]);
                    o[2 * i + 1] = fmaf(w, y.y, o[2 * i + 1]);
                  }
                }
                uint4 u;
                u.x = packc<FMT>(o[0], o[1]); u.y = packc<FMT>(o[2], o[3]);
                u.z = packc<FMT>(o[4], o[5]); u.w = packc<FMT>(o[6], o[7]);
                dst4[q8] = u;
              }
            }
          } else {
            uint32_t v0[32];
            tmem_ld32(ts + TACC + hsel * 32, v0);
            tmem_ld_wait();
            if (valid) {   // dQ_r += ds_t k_256 ; each half stores 32 of the 64 columns
              uint4* dst4 = reinterpret_cast<uint4*>(P.dqkv + grow * 3 * W + colQ + h * 64 +
                                                     hsel * 32);
#pragma unroll
              for (int q8 = 0; q8 < 4; ++q8) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v0[q8 * 8 + i]);
                if (TAIL) {
                  const uint4 tu = *reinterpret_cast<const uint4*>(tv0 + hsel * 32 + q8 * 8);
                  const uint32_t* ta = &tu.x;
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 y = unpackc<FMT>(ta[i]);
                    o[2 * i] = fmaf(dst, y.x, o[2 * i]);
                    o[2 * i + 1] = fmaf(dst, y.y, o[2 * i + 1]);
                  }
                }
                uint4 u;
                u.x = packc<FMT>(o[0], o[1]); u.y = packc<FMT>(o[2], o[3]);
                u.z = packc<FMT>(o[4], o[5]); u.w = packc<FMT>(o[6], o[7]);
                dst4[q8] = u;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar_acc_free);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AT_CTRL_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int MODE, int NCH, int FMT, bool TAIL, bool FULL>
static int launch_bwd2(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                       const CUtensorMap& d, const AttnBwd2Params& P, cudaStream_t st) {
  auto kern = attn_bwd2_kernel<MODE, NCH, FMT, TAIL, FULL>;
  const size_t smem = (size_t)2 * NCH * 64 * 128 + 4 * 16384 + (size_t)NCH * 16384 +
                      (272 * 2 + 256 * 3 + 16) * 4 + 512 + 128 + 1024;
  static DeviceOnce set;
  if (!set.done()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    set.mark();
  }
  int grid = sm_count();
  if (P.B * P.H < grid) grid = P.B * P.H;
  kern<<<grid, AT_THREADS, smem, st>>>(a, b, c, d, P);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
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
  static DeviceMax set;
  if (set.need(pl.smem)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pl.smem);
    if (e != cudaSuccess) return -(int)e - 1000;
  }
  int grid = sm_count();
  if (pl.P.B * pl.P.H < grid) grid = pl.P.B * pl.P.H;
  kern<<<grid, AT_THREADS, pl.smem, st>>>(pl.tmTile, pl.tmFull, pl.tmDoTile, pl.tmDoFull, pl.P);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

int attn_fwd2_launch(const CUtensorMap& tmTile, const CUtensorMap& tmFull, const AttnFwd2Params& P,
                     cudaStream_t st) {
  const int fmt = P.fmt;
  const size_t smem = (size_t)2 * P.kvm * 128 + 2 * 16384 + (size_t)2 * P.nblk * 16384 + 512 +
                      256 + 544 + 128 + 1024;   // + q_tail[2][64] + part[2][68] + barriers + align
  auto kern = fmt ? attn_fwd2_kernel<1> : attn_fwd2_kernel<0>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e - 1000;
  int grid = sm_count();
  if (P.B * P.H < grid) grid = P.B * P.H;
  kern<<<grid, AT_THREADS, smem, st>>>(tmTile, tmFull, P);
  count_launch(1);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

// dQ + dK/dV kernel pair (version 2) for kvp in {64, 256}, the first-generation kernel otherwise
int attn_bwd_generic(const void* qkv16, const void* o16, const void* do16, float* lse_tmp,
                     void* dqkv16, int B, int T, int H, int fmt_o, int fmt_g, cudaStream_t st) {
  const int tm = T < 256 ? T : 256;
  const int kvm = (tm + 15) / 16 * 16;
  if (T <= 257 && (kvm == 64 || kvm == 256)) {
    AttnBwd2Params P{};
    P.B = B; P.T = T; P.H = H; P.W = H * 64;
    P.tm = tm; P.kvm = kvm; P.nblk = kvm / 64; P.ntile = (tm + 127) / 128;
    P.tail = T > 256 ? 1 : 0;
    P.fmt_o = fmt_o;
    P.qkv = reinterpret_cast<const uint16_t*>(qkv16);
    P.o = reinterpret_cast<const uint16_t*>(o16);
    P.d_o = reinterpret_cast<const uint16_t*>(do16);
    P.lse_tmp = lse_tmp;
    P.dqkv = reinterpret_cast<uint16_t*>(dqkv16);
    CUtensorMap tmTile, tmFull, tmDoTile, tmDoFull;
    const uint64_t W3 = 3ull * P.W, Wd = P.W;
    int rc = make_tmap_3d(&tmTile, qkv16, W3, T, B, W3, (uint64_t)T * W3, 64, 128);
    if (!rc) rc = make_tmap_3d(&tmFull, qkv16, W3, T, B, W3, (uint64_t)T * W3, 64, kvm);
    if (!rc) rc = make_tmap_3d(&tmDoTile, do16, Wd, T, B, Wd, (uint64_t)T * Wd, 64, 128);
    if (!rc) rc = make_tmap_3d(&tmDoFull, do16, Wd, T, B, Wd, (uint64_t)T * Wd, 64, kvm);
    if (rc) return rc;
    // ViT-L/14 (T = 257): every tile row / column valid, one peeled token -> fully static code
#define RVLM_BWD2_ARGS tmTile, tmFull, tmDoTile, tmDoFull, P, st
#define RVLM_BWD2(MODE)                                                                          \
    (T == 257 ? (fmt_g ? launch_bwd2<MODE, 4, 1, true, true>(RVLM_BWD2_ARGS)                      \
                       : launch_bwd2<MODE, 4, 0, true, true>(RVLM_BWD2_ARGS))                     \
     : kvm == 256 ? (fmt_g ? launch_bwd2<MODE, 4, 1, false, false>(RVLM_BWD2_ARGS)                \
                           : launch_bwd2<MODE, 4, 0, false, false>(RVLM_BWD2_ARGS))               \
                  : (fmt_g ? launch_bwd2<MODE, 1, 1, false, false>(RVLM_BWD2_ARGS)                \
                           : launch_bwd2<MODE, 1, 0, false, false>(RVLM_BWD2_ARGS)))
    rc = RVLM_BWD2(AT_DQ);
    if (rc) return rc;
    return RVLM_BWD2(AT_DKV);
#undef RVLM_BWD2
#undef RVLM_BWD2_ARGS
  }
  AttnPlan pl{};
  int rc = make_plan(pl, qkv16, do16, B, T, H, fmt_g, fmt_o, fmt_g);
  if (rc) return rc;
  pl.P.o = reinterpret_cast<const uint16_t*>(o16);
  pl.P.d_o = reinterpret_cast<const uint16_t*>(do16);
  pl.P.lse_tmp = lse_tmp;
  pl.P.dqkv = reinterpret_cast<uint16_t*>(dqkv16);
  rc = launch_attn<AT_DQ>(pl, st);
  if (rc) return rc;
  return launch_attn<AT_DKV>(pl, st);
}

}  // namespace rvlm
