// Split-operand ("fp32-class") attention forward for the CLIP ViT (head dim 64, T <= 257,
// no mask): softmax(q k^T / 8) v with q, k, v, the probabilities and the output carried as
// fp16 hi + lo pairs, every tensor-core product formed as hi*hi + lo*hi + hi*lo (three
// tcgen05 MMAs into the same TMEM accumulator; the dropped lo*lo term is ~2^-22 relative).
//
// Why: FARE's loss ||phi(x + delta) - phi_orig(x)||^2 (train/adversarial_training_clip.py:509-521)
// is a cancellation site -- a 4e-4 relative forward error (fp16 operands) becomes a ~4 % gradient
// error and ~1 % flipped signs of delta when model_orig == model (SURVEY.md Appendix E;
// tools/precision_probe.py reproduces the measured numbers and shows that splitting the linear
// layers alone leaves 0.35 %, q / k / v too 0.11 %, the probabilities too 0.005 %).  This kernel
// is the attention part of the engine's `split` precision mode; the linear layers use the SPLIT
// instantiation of gemm16_2cta_kernel.
//
// One work item = (image, head); persistent CTAs, one per SM; see the kernel comment below.
#include "gemm.cuh"
#include <cstdlib>
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);

namespace {

constexpr int SP_THREADS = 160;
constexpr int SP_CTRL = 4;

__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn,
                                                 uint32_t b_mn) {
  return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// =====================================================================================
// T <= 256, or T = 257 with the last token peeled.  (A first version -- one work item per query
// tile, K then V through one buffer, single P buffer, the 257th token as a third tile -- measured
// 894 us / layer at B = 256 with every TMA load and every P V hand-off exposed,
// profiles/r02_kernels_ncu.txt; this one 612 us.)  One work item = (image, head); K and V (256
// tokens, hi + lo) stay
// RESIDENT for its two query tiles, the next item's K / V / Q are requested as soon as their
// buffers are free, P travels in 64-column chunks through TWO buffers (the exp pass of chunk c + 1
// runs under the P V MMAs of chunk c), and the 257th token is peeled as in attn_fwd3: its key is a
// 64-MAC dot product per row folded into the softmax, its query row is done on the CUDA cores
// in the shadow of the first S MMA -- with hi + lo operands reconstructed exactly in fp32.
// Shared memory: Q 32 KB + K 64 KB + V 64 KB + P 2 x 32 KB = 224 KB.
// =====================================================================================
struct Split2Params {
  int B, T, H, W;
  int kvp;       // min(T, 256) rounded up to 16
  int ntile;     // query tiles of 128 rows over the first min(T, 256) tokens
  int nchunk;    // 64-column chunks of P
  const uint16_t* qkv_hi;
  const uint16_t* qkv_lo;
  uint16_t* o_hi;
  uint16_t* o_lo;
  float* lse;
};

// x = hi + lo reconstructed in fp32 (exact: 11 + 11 mantissa bits)
__device__ __forceinline__ void unpack_hl8(const uint4& h, const uint4& l, float (&x)[8]) {
  const uint32_t* ph = &h.x;
  const uint32_t* pl = &l.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = unpack2(ph[i], 0), b = unpack2(pl[i], 0);
    x[2 * i] = a.x + b.x;
    x[2 * i + 1] = a.y + b.y;
  }
}
// dot product of row r of a swizzled [rows x 64] hi / lo tile pair with a 64-float vector
__device__ __forceinline__ float dot_swz_hl(const uint8_t* thi, const uint8_t* tlo, int r,
                                            const float* vec) {
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
    float x[8];
    unpack_hl8(*reinterpret_cast<const uint4*>(thi + off), *reinterpret_cast<const uint4*>(tlo + off), x);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      acc0 = fmaf(x[i], vec[c * 8 + i], acc0);
      acc1 = fmaf(x[i + 1], vec[c * 8 + i + 1], acc1);
    }
  }
  return acc0 + acc1;
}
__device__ __forceinline__ float swz_hl_at(const uint8_t* thi, const uint8_t* tlo, int r, int d) {
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)(((d >> 3) ^ (r & 7)) << 4) + (d & 7) * 2;
  return __half2float(*reinterpret_cast<const __half*>(thi + off)) +
         __half2float(*reinterpret_cast<const __half*>(tlo + off));
}

template <bool PEEL>
__global__ void __launch_bounds__(SP_THREADS, 1)
attn_fwd_split2_kernel(const __grid_constant__ CUtensorMap tmQhi,   // box 64 x 128
                       const __grid_constant__ CUtensorMap tmQlo,
                       const __grid_constant__ CUtensorMap tmKVhi,  // box 64 x kvp
                       const __grid_constant__ CUtensorMap tmKVlo,
                       Split2Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_q = smem;                  // [2 (hi, lo)][128][64]
  uint8_t* s_k = s_q + 2 * 16384;       // [2][256][64]
  uint8_t* s_v = s_k + 2 * 32768;       // [2][256][64]
  uint8_t* s_p = s_v + 2 * 32768;       // [2 buffers][2 (hi, lo)][128][64]: 64-column chunks of P
  float* s_tail = reinterpret_cast<float*>(s_p + 4 * 16384);   // [3][64] k_256 | v_256 | q_256 (fp32 = hi + lo)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tail + 192);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;       // [2] 128 arrivals: chunk written into P buffer i
  uint64_t* bar_pv = bars + 6;      // [2] that chunk's P V MMAs retired
  uint64_t* bar_free = bars + 8;    // 128 arrivals: O and S read out
  uint64_t* bar_tail = bars + 9;    // 128 arrivals: peeled query row done with K / V (PEEL)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(&bar_p[0], 128);
    mbar_init(&bar_p[1], 128);
    mbar_init(&bar_pv[0], 1);
    mbar_init(&bar_pv[1], 1);
    mbar_init(bar_free, 128);
    mbar_init(bar_tail, 128);
    fence_barrier_init();
  }
  if (warp == SP_CTRL) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t TO = 384;   // O accumulator [384, 448); S in [0, kvp)
  const int n_items = P.B * P.H;
  const int W = P.W, T = P.T, kvp = P.kvp, ntile = P.ntile, nchunk = P.nchunk;
  const uint32_t kv_bytes = (uint32_t)kvp * 128u;

  if (warp == SP_CTRL) {
    if (lane == 0) {
      // ===================== TMA + MMA issuer =====================
      const uint32_t qa = smem_u32(s_q), ka = smem_u32(s_k), va = smem_u32(s_v), pa = smem_u32(s_p);
      const uint32_t idS = idesc_f16(128, kvp, 0, 0);
      const uint32_t idO = idesc_f16(128, 64, 0, 1);
      auto load_q = [&](int item, int t) {
        mbar_arrive_expect_tx(bar_q, 2 * 16384);
        tma_load_3d(s_q, &tmQhi, bar_q, (item % P.H) * 64, t * 128, item / P.H);
        tma_load_3d(s_q + 16384, &tmQlo, bar_q, (item % P.H) * 64, t * 128, item / P.H);
      };
      auto load_kv = [&](uint8_t* dst, uint64_t* bar, int col0, int item) {
        mbar_arrive_expect_tx(bar, 2 * kv_bytes);
        tma_load_3d(dst, &tmKVhi, bar, col0 + (item % P.H) * 64, 0, item / P.H);
        tma_load_3d(dst + 32768, &tmKVlo, bar, col0 + (item % P.H) * 64, 0, item / P.H);
      };
      uint32_t g = 0, n = 0;
      int it = 0;
      if ((int)blockIdx.x < n_items) {
        load_kv(s_k, bar_k, W, blockIdx.x);
        load_q(blockIdx.x, 0);
        load_kv(s_v, bar_v, 2 * W, blockIdx.x);
      }
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int next = item + gridDim.x;
        for (int t = 0; t < ntile; ++t, ++g) {
          const bool last = t + 1 == ntile;
          if (t == 0) mbar_wait(bar_k, it & 1, 300);
          mbar_wait(bar_q, g & 1, 301);
          if (g > 0) mbar_wait(bar_free, (g - 1) & 1, 302);   // TMEM drained by the previous tile
          tc_fence_after();
          // S = Qhi Khi^T + Qlo Khi^T + Qhi Klo^T
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            const uint32_t a = qa + (pr == 1 ? 16384u : 0u);
            const uint32_t bb = ka + (pr == 2 ? 32768u : 0u);
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_f16(tmem, make_smem_desc_sw128(a + s * 32, 16, 1024),
                       make_smem_desc_sw128(bb + s * 32, 16, 1024), idS, (pr | s) != 0);
          }
          umma_commit(bar_s);
          if (t == 0) mbar_wait(bar_v, it & 1, 303);
          for (int c = 0; c < nchunk; ++c, ++n) {
            const uint32_t buf = n & 1;
            mbar_wait(&bar_p[buf], (n >> 1) & 1, 304);
            if (c == 0) {
              // the workers are through with their Q rows (peeled key) and S is complete: Q may go
              mbar_wait(bar_s, g & 1, 305);
              if (!last) load_q(item, t + 1);
              else if (next < n_items) {
                load_q(next, 0);
                if (PEEL) mbar_wait(bar_tail, it & 1, 306);   // peeled query row done with K
                load_kv(s_k, bar_k, W, next);
              }
            }
            tc_fence_after();
            const int cols = min(64, kvp - c * 64);
            const int steps = cols >> 4;
            // O += Phi Vhi + Plo Vhi + Phi Vlo over this chunk's tokens
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
              const uint32_t a = pa + buf * 32768u + (pr == 1 ? 16384u : 0u);
              const uint32_t bb = va + (pr == 2 ? 32768u : 0u) + c * 8192;
              for (int s = 0; s < steps; ++s)
                umma_f16(tmem + TO, make_smem_desc_sw128(a + s * 32, 16, 1024),
                         make_smem_desc_sw128(bb + s * 2048, 8192, 1024), idO, (c | pr | s) != 0);
            }
            umma_commit(&bar_pv[buf]);
          }
          if (last && next < n_items) {   // V free once the last chunk's MMAs have retired
            const uint32_t nl = n - 1;
            mbar_wait(&bar_pv[nl & 1], (nl >> 1) & 1, 307);
            load_kv(s_v, bar_v, 2 * W, next);
          }
        }
      }
    }
  } else {
    // ===================== row workers =====================
    const int r = threadIdx.x;   // 0 .. 127
    const uint32_t ts = tmem + ((uint32_t)(warp * 32) << 16);
    const float c2 = 0.125f * 1.4426950408889634f;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    uint32_t g = 0, n = 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / P.H, h = item % P.H;
      if (PEEL) {
        // token 256 of this (image, head): k, v, q rows (hi + lo -> fp32) into shared memory
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the previous item's readers are done
        if (r < 24) {
          const int which = r >> 3, c = r & 7;           // 0: k, 1: v, 2: q
          const size_t off = ((size_t)b * T + 256) * 3 * W + h * 64 + (which == 0 ? W : which == 1 ? 2 * W : 0) + c * 8;
          float x[8];
          unpack_hl8(*reinterpret_cast<const uint4*>(P.qkv_hi + off),
                     *reinterpret_cast<const uint4*>(P.qkv_lo + off), x);
#pragma unroll
          for (int i = 0; i < 8; ++i) s_tail[which * 64 + c * 8 + i] = x[i];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // ---- the peeled QUERY row on the CUDA cores, in the shadow of the first S MMA; scratch in
        // P buffer 1 (idle until this tile's second chunk)
        float* sc = reinterpret_cast<float*>(s_p + 32768);   // [257] scores / probabilities
        float* red = sc + 264;                               // [8]
        float* oh = red + 8;                                 // [2][64]
        const float* kt = s_tail;
        const float* vt = s_tail + 64;
        const float* qt = s_tail + 128;
        mbar_wait(bar_k, it & 1, 320);
        float lmax = -INFINITY;
        for (int kv = r; kv < T; kv += 128) {
          float acc;
          if (kv < 256) acc = dot_swz_hl(s_k, s_k + 32768, kv, qt);
          else {
            acc = 0.f;
            for (int d = 0; d < 64; ++d) acc = fmaf(qt[d], kt[d], acc);
          }
          sc[kv] = acc;
          lmax = fmaxf(lmax, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffff, lmax, o));
        if (lane == 0) red[warp] = lmax;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const float tmc = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) * c2;
        float lsum = 0.f;
        for (int kv = r; kv < T; kv += 128) {
          const float e = ex2f(fmaf(sc[kv], c2, -tmc));
          sc[kv] = e;
          lsum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffff, lsum, o);
        if (lane == 0) red[4 + warp] = lsum;
        mbar_wait(bar_v, it & 1, 321);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const float tsum = (red[4] + red[5]) + (red[6] + red[7]);
        {
          const int d = r & 63, hf = r >> 6;
          float acc0 = 0.f, acc1 = 0.f;
          for (int j = hf * 128; j < hf * 128 + 128; j += 2) {
            acc0 = fmaf(sc[j], swz_hl_at(s_v, s_v + 32768, j, d), acc0);
            acc1 = fmaf(sc[j + 1], swz_hl_at(s_v, s_v + 32768, j + 1, d), acc1);
          }
          float acc = acc0 + acc1;
          if (hf == 1) acc = fmaf(sc[256], vt[d], acc);
          oh[hf * 64 + d] = acc;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (r < 32) {
          const float inv = 1.f / tsum;
          const float o0 = (oh[2 * r] + oh[64 + 2 * r]) * inv, o1 = (oh[2 * r + 1] + oh[64 + 2 * r + 1]) * inv;
          const uint32_t hi = pack2_f16(o0, o1);
          const float2 hf2 = unpack2(hi, 0);
          const size_t off = ((size_t)b * T + 256) * W + h * 64;
          reinterpret_cast<uint32_t*>(P.o_hi + off)[r] = hi;
          if (P.o_lo) reinterpret_cast<uint32_t*>(P.o_lo + off)[r] = pack2_f16(o0 - hf2.x, o1 - hf2.y);
          if (r == 0 && P.lse)
            P.lse[((size_t)b * P.H + h) * T + 256] = (tmc + log2f(tsum)) * 0.6931471805599453f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // scratch (P buffer 1) free again
        mbar_arrive(bar_tail);
      }
      for (int t = 0; t < ntile; ++t, ++g) {
        const int row = t * 128 + r;
        const bool wvalid = t * 128 + warp * 32 < T;   // warp-uniform: a fully padded warp skips the sweeps
        float s_t = 0.f;
        mbar_wait(bar_q, g & 1, 322);
        if (PEEL) s_t = dot_swz_hl(s_q, s_q + 16384, r, s_tail);   // q_r . k_256
        mbar_wait(bar_s, g & 1, 323);
        tc_fence_after();
        float mx = PEEL ? s_t : -INFINITY;
        for (int q = 0; wvalid && q < kvp; q += 16) {
          uint32_t v[16];
          tmem_ld16(ts + q, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (q + i < T) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float mc = mx * c2;
        const float p_t = PEEL ? ex2f(fmaf(s_t, c2, -mc)) : 0.f;
        float sum = p_t;
        for (int c = 0; c < nchunk; ++c, ++n) {
          const uint32_t buf = n & 1;
          if (n >= 2) mbar_wait(&bar_pv[buf], ((n >> 1) - 1) & 1, 324);   // this buffer's previous chunk retired
          uint8_t* prow = s_p + buf * 32768 + r * 128;
          const int cols = min(64, kvp - c * 64);
          for (int q = 0; wvalid && q < cols; q += 16) {
            uint32_t v[16];
            tmem_ld16(ts + c * 64 + q, v);
            tmem_ld_wait();
            float e[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              e[i] = ex2f(fmaf(__uint_as_float(v[i]), c2, -mc));
              if (c * 64 + q + i >= T) e[i] = 0.f;
              sum += e[i];
            }
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float e0 = e[q8 * 8 + 2 * i], e1 = e[q8 * 8 + 2 * i + 1];
                hi[i] = pack2_f16(e0, e1);
                const float2 hf = unpack2(hi[i], 0);
                lo[i] = pack2_f16(e0 - hf.x, e1 - hf.y);
              }
              const uint32_t off = (((uint32_t)((q >> 3) + q8)) << 4) ^ rx;
              *reinterpret_cast<uint4*>(prow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(prow + 16384 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&bar_p[buf]);
        }
        // ---- epilogue: (O + p_t v_256) / rowsum -> hi / lo
        {
          const uint32_t nl = n - 1;
          mbar_wait(&bar_pv[nl & 1], (nl >> 1) & 1, 325);
        }
        tc_fence_after();
        if (wvalid) {
          uint32_t v0[32], v1[32];
          tmem_ld32(ts + TO, v0);
          tmem_ld32(ts + TO + 32, v1);
          tmem_ld_wait();
          if (row < T) {
            const float inv = 1.f / sum;
            const size_t off = ((size_t)b * T + row) * W + h * 64;
            uint4* dhi = reinterpret_cast<uint4*>(P.o_hi + off);
            uint4* dlo = P.o_lo ? reinterpret_cast<uint4*>(P.o_lo + off) : nullptr;
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8) {
              const uint32_t* sv = q8 < 4 ? &v0[q8 * 8] : &v1[(q8 - 4) * 8];
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float o0 = __uint_as_float(sv[2 * i]), o1 = __uint_as_float(sv[2 * i + 1]);
                if (PEEL) {
                  o0 = fmaf(p_t, s_tail[64 + q8 * 8 + 2 * i], o0);
                  o1 = fmaf(p_t, s_tail[64 + q8 * 8 + 2 * i + 1], o1);
                }
                o0 *= inv;
                o1 *= inv;
                hi[i] = pack2_f16(o0, o1);
                const float2 hf = unpack2(hi[i], 0);
                lo[i] = pack2_f16(o0 - hf.x, o1 - hf.y);
              }
              dhi[q8] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              if (dlo) dlo[q8] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (P.lse)
              P.lse[((size_t)b * P.H + h) * T + row] = (mc + log2f(sum)) * 0.6931471805599453f;
          }
        }
        tc_fence_before();
        mbar_arrive(bar_free);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == SP_CTRL) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <bool PEEL>
int launch_split2(const CUtensorMap* tq, const CUtensorMap* tkv, const Split2Params& P,
                  cudaStream_t st) {
  auto kern = attn_fwd_split2_kernel<PEEL>;
  const size_t smem = 2 * 16384 + 4 * 32768 + 4 * 16384 + 192 * 4 + 128 + 1024;
  static DeviceOnce attr;
  if (!attr.done()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    attr.mark();
  }
  int grid = sm_count();
  if (P.B * P.H < grid) grid = P.B * P.H;
  kern<<<grid, SP_THREADS, smem, st>>>(tq[0], tq[1], tkv[0], tkv[1], P);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

}  // namespace

// q | k | v as fp16 hi + lo pairs in two [B*T, 3W] buffers; o as hi + lo in two [B*T, W] buffers
// (o_lo may be null); lse [B, H, T] may be null.
int attn_fwd_split(const void* qkv_hi, const void* qkv_lo, void* o_hi, void* o_lo, float* lse,
                   int B, int T, int H, cudaStream_t st) {
  if (!qkv_hi || !qkv_lo || !o_hi || B <= 0 || T <= 0 || T > 272 || H <= 0)
    return RVLM_ERR_UNSUPPORTED;
  if (T > 257) return RVLM_ERR_UNSUPPORTED;   // one peeled token at most
  {
    Split2Params P2{};
    const int tm = T < 256 ? T : 256;
    P2.B = B; P2.T = T; P2.H = H; P2.W = H * 64;
    P2.kvp = (tm + 15) / 16 * 16;
    P2.ntile = (tm + 127) / 128;
    P2.nchunk = (P2.kvp + 63) / 64;
    P2.qkv_hi = reinterpret_cast<const uint16_t*>(qkv_hi);
    P2.qkv_lo = reinterpret_cast<const uint16_t*>(qkv_lo);
    P2.o_hi = reinterpret_cast<uint16_t*>(o_hi);
    P2.o_lo = reinterpret_cast<uint16_t*>(o_lo);
    P2.lse = lse;
    CUtensorMap tq2[2], tkv2[2];
    const uint64_t W3b = 3ull * P2.W;
    const void* src2[2] = {qkv_hi, qkv_lo};
    for (int i = 0; i < 2; ++i) {
      int rc = make_tmap_3d(&tq2[i], src2[i], W3b, T, B, W3b, (uint64_t)T * W3b, 64, 128);
      if (!rc) rc = make_tmap_3d(&tkv2[i], src2[i], W3b, T, B, W3b, (uint64_t)T * W3b, 64, P2.kvp);
      if (rc) return rc;
    }
    return T == 257 ? launch_split2<true>(tq2, tkv2, P2, st) : launch_split2<false>(tq2, tkv2, P2, st);
  }
}

}  // namespace rvlm

// C ABI (include/rvlm.h §3)
extern "C" int rvlm_attn_fwd_split(const void* qkv_hi, const void* qkv_lo, void* o_hi, void* o_lo,
                                   float* lse, int B, int T, int H, void* stream) {
  return rvlm::attn_fwd_split(qkv_hi, qkv_lo, o_hi, o_lo, lse, B, T, H, (cudaStream_t)stream);
}
