// Split-operand ("fp32-class") attention forward for the CLIP ViT (head dim 64, T <= 272,
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
// One work item = (image, head, 128-row query tile); persistent CTAs, one per SM:
//   warp 4 lane 0 : TMA (Q, K hi / lo; then V hi / lo into the K buffers once S is done) and MMA
//                   issuer:  S[128 x kvp] = Q K^T  (N = 256 [+ 16 for the 257th token])
//                            O[128 x 64] += P_c V_c for each 128-column chunk c of P
//   warps 0-3     : thread = one query row: row max over TMEM, then per chunk exp2 ->
//                   hi / lo fp16 -> swizzled smem (the K-major A operand of the P V MMAs);
//                   finally O / rowsum -> o hi / lo, log-sum-exp.
// Rows / columns beyond T are zero-filled by the per-image 3-D tensor maps and masked in the
// softmax; kvp = T rounded up to 16.  Correctness-first: single-buffered, ~1.7x the time of the
// fp16 kernel (attn_fwd3) for three times the tensor work.
#include "gemm.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);

namespace {

constexpr int SP_THREADS = 160;
constexpr int SP_CTRL = 4;
constexpr int SP_KV_BYTES = 272 * 128;   // one K / V half (hi or lo), up to 272 tokens

struct SplitParams {
  int B, T, H, W;
  int kvp;       // T rounded up to 16
  int n0;        // min(kvp, 256): first UMMA N extent; the rest (16) is the tail extent
  int ntile;     // query tiles of 128 rows
  int nchunk;    // 128-column chunks of P (the last may be short)
  uint16_t* o_hi;
  uint16_t* o_lo;
  float* lse;    // may be null
};

__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn,
                                                 uint32_t b_mn) {
  return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(SP_THREADS, 1)
attn_fwd_split_kernel(const __grid_constant__ CUtensorMap tmQhi,   // box 64 x 128
                      const __grid_constant__ CUtensorMap tmQlo,
                      const __grid_constant__ CUtensorMap tmKVhi,  // box 64 x n0
                      const __grid_constant__ CUtensorMap tmKVlo,
                      const __grid_constant__ CUtensorMap tmTlhi,  // box 64 x 16 (tokens 256..271)
                      const __grid_constant__ CUtensorMap tmTllo,
                      SplitParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_q = smem;                         // [2 (hi, lo)][128][64]
  uint8_t* s_kv = s_q + 2 * 16384;             // [2 (hi, lo)][272][64]  K, then V
  uint8_t* s_p = s_kv + 2 * SP_KV_BYTES;       // [2 (hi, lo)][2][128][64]  one 128-column chunk of P
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + 4 * 16384);
  uint64_t* bar_qk = bars + 0;      // Q and K landed
  uint64_t* bar_s = bars + 1;       // S complete (Q, K consumed)
  uint64_t* bar_v = bars + 2;       // V landed
  uint64_t* bar_p = bars + 3;       // 128 arrivals: a chunk of P written
  uint64_t* bar_pv = bars + 4;      // the chunk's P V MMAs retired (P buffer free; last chunk: O done)
  uint64_t* bar_free = bars + 5;    // 128 arrivals: O and S read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_pv, 1);
    mbar_init(bar_free, 128);
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
  constexpr uint32_t TO = 384;   // O accumulator columns [384, 448); S in [0, kvp)
  const int n_work = P.B * P.H * P.ntile;
  const int W = P.W, T = P.T, kvp = P.kvp, n0 = P.n0, ntl = kvp - n0;
  const uint32_t kv_bytes = (uint32_t)kvp * 128u;

  if (warp == SP_CTRL) {
    if (lane == 0) {
      const uint32_t qa = smem_u32(s_q), ka = smem_u32(s_kv), pa = smem_u32(s_p);
      const uint32_t idS0 = idesc_f16(128, n0, 0, 0);
      const uint32_t idS1 = idesc_f16(128, ntl > 0 ? ntl : 16, 0, 0);
      const uint32_t idO = idesc_f16(128, 64, 0, 1);
      uint32_t nchunks_done = 0;
      int k = 0;
      for (int idx = blockIdx.x; idx < n_work; idx += gridDim.x, ++k) {
        const uint32_t ph = k & 1;
        const int tile = idx % P.ntile, item = idx / P.ntile;
        const int b = item / P.H, h = item % P.H;
        // the previous item's O / S are read out and its last MMAs have retired
        if (k > 0) mbar_wait(bar_free, ph ^ 1, 200);
        mbar_arrive_expect_tx(bar_qk, 2 * 16384 + 2 * kv_bytes);
        tma_load_3d(s_q, &tmQhi, bar_qk, h * 64, tile * 128, b);
        tma_load_3d(s_q + 16384, &tmQlo, bar_qk, h * 64, tile * 128, b);
        tma_load_3d(s_kv, &tmKVhi, bar_qk, W + h * 64, 0, b);
        tma_load_3d(s_kv + SP_KV_BYTES, &tmKVlo, bar_qk, W + h * 64, 0, b);
        if (ntl > 0) {
          tma_load_3d(s_kv + 32768, &tmTlhi, bar_qk, W + h * 64, 256, b);
          tma_load_3d(s_kv + SP_KV_BYTES + 32768, &tmTllo, bar_qk, W + h * 64, 256, b);
        }
        mbar_wait(bar_qk, ph, 201);
        tc_fence_after();
        // S = Qhi Khi^T + Qlo Khi^T + Qhi Klo^T
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          const uint32_t a = qa + (pr == 1 ? 16384u : 0u);
          const uint32_t bb = ka + (pr == 2 ? (uint32_t)SP_KV_BYTES : 0u);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const uint64_t ad = make_smem_desc_sw128(a + s * 32, 16, 1024);
            umma_f16(tmem, ad, make_smem_desc_sw128(bb + s * 32, 16, 1024), idS0, (pr | s) != 0);
            if (ntl > 0)
              umma_f16(tmem + 256, ad, make_smem_desc_sw128(bb + 32768 + s * 32, 16, 1024), idS1,
                       (pr | s) != 0);
          }
        }
        umma_commit(bar_s);
        mbar_wait(bar_s, ph, 202);   // K consumed: its buffers take V
        mbar_arrive_expect_tx(bar_v, 2 * kv_bytes);
        tma_load_3d(s_kv, &tmKVhi, bar_v, 2 * W + h * 64, 0, b);
        tma_load_3d(s_kv + SP_KV_BYTES, &tmKVlo, bar_v, 2 * W + h * 64, 0, b);
        if (ntl > 0) {
          tma_load_3d(s_kv + 32768, &tmTlhi, bar_v, 2 * W + h * 64, 256, b);
          tma_load_3d(s_kv + SP_KV_BYTES + 32768, &tmTllo, bar_v, 2 * W + h * 64, 256, b);
        }
        mbar_wait(bar_v, ph, 203);
        for (int c = 0; c < P.nchunk; ++c, ++nchunks_done) {
          mbar_wait(bar_p, nchunks_done & 1, 204);
          tc_fence_after();
          const int cols = min(128, kvp - c * 128);
          const int steps = cols >> 4;
          // O += Phi Vhi + Plo Vhi + Phi Vlo over this chunk's tokens
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            const uint32_t a = pa + (pr == 1 ? 32768u : 0u);
            const uint32_t bb = ka + (pr == 2 ? (uint32_t)SP_KV_BYTES : 0u) + c * 16384;
            for (int s = 0; s < steps; ++s)
              umma_f16(tmem + TO, make_smem_desc_sw128(a + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024),
                       make_smem_desc_sw128(bb + s * 2048, 8192, 1024), idO, (c | pr | s) != 0);
          }
          umma_commit(bar_pv);
        }
      }
    }
  } else {
    // ===================== row workers =====================
    const int r = threadIdx.x;   // 0 .. 127
    const uint32_t ts = tmem + ((uint32_t)(warp * 32) << 16);
    const float c2 = 0.125f * 1.4426950408889634f;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    uint8_t* prow = s_p + r * 128;
    uint32_t nchunks_done = 0;
    int k = 0;
    for (int idx = blockIdx.x; idx < n_work; idx += gridDim.x, ++k) {
      const uint32_t ph = k & 1;
      const int tile = idx % P.ntile, item = idx / P.ntile;
      const int b = item / P.H, h = item % P.H;
      const int row = tile * 128 + r;
      // a warp whose 32 rows all lie beyond T (the last tile of T = 257 has ONE valid row) keeps
      // the barrier protocol but skips the sweeps: its P rows are never read back
      const bool wvalid = tile * 128 + warp * 32 < T;
      mbar_wait(bar_s, ph, 210);
      tc_fence_after();
      // ---- row max over the valid columns
      float mx = -INFINITY;
      for (int q = 0; wvalid && q < kvp; q += 16) {
        uint32_t v[16];
        tmem_ld16(ts + q, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (q + i < T) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float mc = mx * c2;
      float sum = 0.f;
      // ---- exp pass, one 128-column chunk of P at a time
      for (int c = 0; c < P.nchunk; ++c, ++nchunks_done) {
        if (nchunks_done > 0) mbar_wait(bar_pv, (nchunks_done - 1) & 1, 211);   // P buffer free
        const int cols = min(128, kvp - c * 128);
        for (int q = 0; wvalid && q < cols; q += 16) {
          uint32_t v[16];
          tmem_ld16(ts + c * 128 + q, v);
          tmem_ld_wait();
          float e[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            e[i] = ex2f(fmaf(__uint_as_float(v[i]), c2, -mc));
            if (c * 128 + q + i >= T) e[i] = 0.f;
            sum += e[i];
          }
#pragma unroll
          for (int q8 = 0; q8 < 2; ++q8) {
            const int cc = q + q8 * 8;   // column within the chunk
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float e0 = e[q8 * 8 + 2 * i], e1 = e[q8 * 8 + 2 * i + 1];
              hi[i] = pack2_f16(e0, e1);
              const float2 hf = unpack2(hi[i], 0);
              lo[i] = pack2_f16(e0 - hf.x, e1 - hf.y);
            }
            const uint32_t off = (uint32_t)(cc >> 6) * 16384u + ((((uint32_t)(cc & 63) >> 3) << 4) ^ rx);
            *reinterpret_cast<uint4*>(prow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(prow + 32768 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_p);
      }
      // ---- epilogue: O / rowsum -> hi / lo
      mbar_wait(bar_pv, (nchunks_done - 1) & 1, 212);
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
              const float o0 = __uint_as_float(sv[2 * i]) * inv, o1 = __uint_as_float(sv[2 * i + 1]) * inv;
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

  tc_fence_before();
  __syncthreads();
  if (warp == SP_CTRL) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

// q | k | v as fp16 hi + lo pairs in two [B*T, 3W] buffers; o as hi + lo in two [B*T, W] buffers
// (o_lo may be null); lse [B, H, T] may be null.
int attn_fwd_split(const void* qkv_hi, const void* qkv_lo, void* o_hi, void* o_lo, float* lse,
                   int B, int T, int H, cudaStream_t st) {
  if (!qkv_hi || !qkv_lo || !o_hi || B <= 0 || T <= 0 || T > 272 || H <= 0)
    return RVLM_ERR_UNSUPPORTED;
  SplitParams P{};
  P.B = B; P.T = T; P.H = H; P.W = H * 64;
  P.kvp = (T + 15) / 16 * 16;
  P.n0 = P.kvp > 256 ? 256 : P.kvp;
  P.ntile = (T + 127) / 128;
  P.nchunk = (P.kvp + 127) / 128;
  P.o_hi = reinterpret_cast<uint16_t*>(o_hi);
  P.o_lo = reinterpret_cast<uint16_t*>(o_lo);
  P.lse = lse;
  CUtensorMap tq[2], tkv[2], ttl[2];
  const uint64_t W3 = 3ull * P.W;
  const void* src[2] = {qkv_hi, qkv_lo};
  for (int i = 0; i < 2; ++i) {
    int rc = make_tmap_3d(&tq[i], src[i], W3, T, B, W3, (uint64_t)T * W3, 64, 128);
    if (!rc) rc = make_tmap_3d(&tkv[i], src[i], W3, T, B, W3, (uint64_t)T * W3, 64, P.n0);
    if (!rc) rc = make_tmap_3d(&ttl[i], src[i], W3, T, B, W3, (uint64_t)T * W3, 64, 16);
    if (rc) return rc;
  }
  const size_t smem = 2 * 16384 + 2 * (size_t)SP_KV_BYTES + 4 * 16384 + 128 + 1024;
  static DeviceOnce attr;
  if (!attr.done()) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_split_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
    attr.mark();
  }
  const int n_work = B * H * P.ntile;
  int grid = sm_count();
  if (n_work < grid) grid = n_work;
  attn_fwd_split_kernel<<<grid, SP_THREADS, smem, st>>>(tq[0], tq[1], tkv[0], tkv[1], ttl[0], ttl[1], P);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

}  // namespace rvlm

// C ABI (include/rvlm.h §3)
extern "C" int rvlm_attn_fwd_split(const void* qkv_hi, const void* qkv_lo, void* o_hi, void* o_lo,
                                   float* lse, int B, int T, int H, void* stream) {
  return rvlm::attn_fwd_split(qkv_hi, qkv_lo, o_hi, o_lo, lse, B, T, H, (cudaStream_t)stream);
}
