// Fused GEMM epilogues.
//
// The GEMM's epilogue warps read 32 accumulator columns per thread-row out of TMEM,
// transpose them through a swizzled warp-private shared-memory tile and then hand each
// lane a float4 of FOUR CONSECUTIVE COLUMNS OF ONE ROW, 8 lanes per row: a warp-wide
// access covers 4 rows x 128 B (fp32) or 4 rows x 64 B (16-bit) of contiguous global
// memory, i.e. whole cache lines instead of 32 scattered 16-byte pieces.
//
// Functor interface (col % 4 == 0, row < M, col < N guaranteed by the caller, N % 32 == 0):
//   Aux  load(row, col)                       HBM operands of the epilogue, requested early
//   Col  col(col)                             per-column operand (bias), once per 32-column chunk
//   void operator()(row, col, acc, aux, col)  math + stores
#pragma once
#include "ptx.cuh"

namespace rvlm {

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint2 pack4(const float4& v, int fmt) {
  uint2 u;
  u.x = pack2(v.x, v.y, fmt);
  u.y = pack2(v.z, v.w, fmt);
  return u;
}
__device__ __forceinline__ float4 unpack4(const uint2& u, int fmt) {
  const float2 a = unpack2(u.x, fmt), b = unpack2(u.y, fmt);
  return make_float4(a.x, a.y, b.x, b.y);
}
// fp16 hi + lo pair of four values: hi = fp16(v), lo = fp16(v - hi)  (split-operand forward)
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  hi = pack4(v, 0);
  const float4 h = unpack4(hi, 0);
  lo = pack4(make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w), 0);
}
// per-column operand of an epilogue (the bias of 4 consecutive columns), loaded once per chunk
__device__ __forceinline__ float4 load_bias(const float* bias, int col) {
  return bias ? __ldg(reinterpret_cast<const float4*>(bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ---- TMA-store epilogue pieces (lane = accumulator row) -----------------------------------------
// The 16-bit outputs of a 32-column chunk go through a per-warp staging tile [32 rows x 64 bytes]
// in the 64B-swizzled layout of make_tmap_2d_store16 and leave with ONE cp.async.bulk.tensor store
// per output, issued by lane 0: no read-back from shared memory, no per-lane global stores or
// address arithmetic, rows past M clipped by the tensor map.
int make_tmap_2d_store16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t ld);
// RVLM_EPI_TMA: 2 (default) = every 16-bit-output epilogue, 1 = only those with two outputs or an
// input tile, 0 = off (the transposing epilogue everywhere).  Measured in-step on one B200 (ms per
// step, levels 0 / 1 / 2): fc + QuickGELU 134.9 / 123.3 / 123.2, its dgrad 123.6 / 118.0 / 115.9,
// the three plain 16-bit dgrad GEMMs 210.1 / 213.9 / 207.5, step 235.1 / 236.1 / 237.7 img/s.
int epi_tma_level();

constexpr int EPI_TMA_TILE_BYTES = 32 * 64;

// the bias of a chunk's 32 columns (the same addresses in every lane: broadcast loads)
__device__ __forceinline__ void load_bias32(const float* bias, int col, float (&b)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = load_bias(bias, col + 4 * i);
    b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
  }
}
// lane's row of a staging tile: 32 packed 16-bit values = four 16-byte groups, group index XORed
// with bits 7-8 of the byte address (CU_TENSOR_MAP_SWIZZLE_64B)
__device__ __forceinline__ void sts_row64(uint8_t* tile, int lane, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(tile + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}
__device__ __forceinline__ void lds_row64(const uint8_t* tile, int lane, uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 t =
        *reinterpret_cast<const uint4*>(tile + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
    pk[4 * j] = t.x; pk[4 * j + 1] = t.y; pk[4 * j + 2] = t.z; pk[4 * j + 3] = t.w;
  }
}
// the staging tiles of this warp are free again (the previous chunk's stores have read them)
__device__ __forceinline__ void tma_stage_acquire(int lane) {
  if (lane == 0) tma_store_wait_read0();
  __syncwarp();
}
// every lane's writes to the staging tiles become visible to the async proxy; then lane 0 may issue
__device__ __forceinline__ void tma_stage_publish() {
  fence_proxy_async_smem();
  __syncwarp();
}

// C16[row, col] = acc + bias[col]
struct NoAux {};

struct EpiStore16 {
  static constexpr const char* kTag = "gemm.store16";
  uint16_t* C;
  int ldc;
  const float* bias;  // may be null
  int fmt;
  // TMA-store path (filled by prepare() on the host, just before the launch)
  int tma = 0;
  CUtensorMap tmC;
  int prepare(int M, int N) {
    tma = epi_tma_level() >= 2 && make_tmap_2d_store16(&tmC, C, M, N, ldc) == 0;
    return 0;
  }
  __device__ __forceinline__ bool tma_on() const { return tma != 0; }
  __device__ __forceinline__ void tma_chunk(int row0, int lane, int col, const uint32_t (&r)[32],
                                            uint8_t* stg, TmaCtx&) const {
    float b[32];
    load_bias32(bias, col, b);
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      pk[i] = pack2(__uint_as_float(r[2 * i]) + b[2 * i], __uint_as_float(r[2 * i + 1]) + b[2 * i + 1],
                    fmt);
    tma_stage_acquire(lane);
    sts_row64(stg, lane, pk);
    tma_stage_publish();
    if (lane == 0) {
      tma_store_2d(&tmC, stg, col, row0);
      tma_store_commit();
    }
  }
  using Aux = NoAux;
  using Col = float4;
  __device__ __forceinline__ Aux load(int, int) const { return {}; }
  __device__ __forceinline__ Col col(int c) const { return load_bias(bias, c); }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux&,
                                             const Col& b) const {
    *reinterpret_cast<uint2*>(C + (size_t)row * ldc + col) = pack4(add4(acc, b), fmt);
  }
};

// Same value written twice in two 16-bit formats: C (fmt) feeds the forward attention,
// C2 (fmt2) is the copy the backward keeps, in the gradient operand format -- a
// tcgen05.mma needs both operands in one format, and the backward's other operand is
// a gradient.
struct EpiStore16Dual {
  static constexpr const char* kTag = "gemm.store16dual";
  uint16_t* C;
  uint16_t* C2;  // may be null
  int ldc;
  const float* bias;
  int fmt, fmt2;
  uint16_t* Clo;  // may be null: fp16 lo half of C (split-operand forward; fmt == 0)
  // TMA-store path: one or two outputs (C, and C2 or Clo); three (split + second format) fall back
  int tma = 0;
  CUtensorMap tmC, tmX;
  int prepare(int M, int N) {
    tma = 0;
    uint16_t* x = C2 ? C2 : Clo;
    if (epi_tma_level() < (x ? 1 : 2) || (C2 && Clo)) return 0;
    if (make_tmap_2d_store16(&tmC, C, M, N, ldc)) return 0;
    if (x && make_tmap_2d_store16(&tmX, x, M, N, ldc)) return 0;
    tma = 1;
    return 0;
  }
  __device__ __forceinline__ bool tma_on() const { return tma != 0; }
  __device__ __forceinline__ void tma_chunk(int row0, int lane, int col, const uint32_t (&r)[32],
                                            uint8_t* stg, TmaCtx&) const {
    float v[32];
    load_bias32(bias, col, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r[i]);
    uint32_t pk[16], px[16];
    const bool second = C2 || Clo;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      pk[i] = pack2(v[2 * i], v[2 * i + 1], Clo ? 0 : fmt);
      if (Clo) {
        const float2 h = unpack2(pk[i], 0);
        px[i] = pack2(v[2 * i] - h.x, v[2 * i + 1] - h.y, 0);
      } else {
        px[i] = pack2(v[2 * i], v[2 * i + 1], fmt2);
      }
    }
    tma_stage_acquire(lane);
    sts_row64(stg, lane, pk);
    if (second) sts_row64(stg + EPI_TMA_TILE_BYTES, lane, px);
    tma_stage_publish();
    if (lane == 0) {
      tma_store_2d(&tmC, stg, col, row0);
      if (second) tma_store_2d(&tmX, stg + EPI_TMA_TILE_BYTES, col, row0);
      tma_store_commit();
    }
  }
  using Aux = NoAux;
  using Col = float4;
  __device__ __forceinline__ Aux load(int, int) const { return {}; }
  __device__ __forceinline__ Col col(int c) const { return load_bias(bias, c); }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux&,
                                             const Col& b) const {
    const float4 v = add4(acc, b);
    if (Clo) {
      uint2 hi, lo;
      split4(v, hi, lo);
      *reinterpret_cast<uint2*>(C + (size_t)row * ldc + col) = hi;
      *reinterpret_cast<uint2*>(Clo + (size_t)row * ldc + col) = lo;
    } else {
      *reinterpret_cast<uint2*>(C + (size_t)row * ldc + col) = pack4(v, fmt);
    }
    if (C2) *reinterpret_cast<uint2*>(C2 + (size_t)row * ldc + col) = pack4(v, fmt2);
  }
};

// C32[row, col] = acc + bias[col] + resid[row, col]   (fp32 residual stream)
struct EpiStore32 {
  static constexpr const char* kTag = "gemm.store32";
  float* C;
  int ldc;
  const float* bias;   // may be null
  const float* resid;  // may be null
  int ldr;
  using Aux = float4;
  using Col = float4;
  __device__ __forceinline__ Aux load(int row, int col) const {
    return resid ? *reinterpret_cast<const float4*>(resid + (size_t)row * ldr + col)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ Col col(int c) const { return load_bias(bias, c); }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux& r,
                                             const Col& b) const {
    *reinterpret_cast<float4*>(C + (size_t)row * ldc + col) = add4(add4(acc, b), r);
  }
};

// QuickGELU (open_clip QuickGELU: x * sigmoid(1.702 x)), with
// MUFU ex2 / rcp approximations (relative error ~1e-7, far below the 16-bit output).
__device__ __forceinline__ float quick_sigmoid(float x) {
  return fast_rcp(1.f + fast_ex2(-2.4554669595930157f * x));  // 1.702 * log2(e)
}

// A16 = quick_gelu(h), h = acc + bias; G16 = quick_gelu'(h) - 1/2, the only thing the backward needs of
// the pre-activation: s = sigmoid(1.702 h) is at hand here, so the derivative costs three FMAs
// in this epilogue and the dgrad epilogue (EpiGeluBwd) is left with one multiply -- no MUFU work
// and no second evaluation of the sigmoid in the backward.  Measured in-step (B200, config 2,
// ms per step): forward GEMM 121.5 -> 128, dgrad GEMM 139.7 -> 113.
struct EpiGelu {
  static constexpr const char* kTag = "gemm.gelu";
  uint16_t* G;  // may be null
  uint16_t* A;
  int ld;
  const float* bias;
  int fmt;
  uint16_t* Alo;  // may be null: fp16 lo half of A (split-operand forward; fmt == 0)
  // TMA-store path: A and (if kept) G; the split forward's third output falls back
  int tma = 0;
  CUtensorMap tmA, tmG;
  int prepare(int M, int N) {
    tma = 0;
    if (epi_tma_level() < 1 || Alo) return 0;
    if (make_tmap_2d_store16(&tmA, A, M, N, ld)) return 0;
    if (G && make_tmap_2d_store16(&tmG, G, M, N, ld)) return 0;
    tma = 1;
    return 0;
  }
  __device__ __forceinline__ bool tma_on() const { return tma != 0; }
  __device__ __forceinline__ void tma_chunk(int row0, int lane, int col, const uint32_t (&r)[32],
                                            uint8_t* stg, TmaCtx&) const {
    float h[32];
    load_bias32(bias, col, h);
    uint32_t pa[16], pg[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float h0 = h[2 * i] + __uint_as_float(r[2 * i]);
      const float h1 = h[2 * i + 1] + __uint_as_float(r[2 * i + 1]);
      const float s0 = quick_sigmoid(h0), s1 = quick_sigmoid(h1);
      const float a0 = h0 * s0, a1 = h1 * s1;
      pa[i] = pack2(a0, a1, fmt);
      pg[i] = pack2(fmaf(1.702f * a0, 1.f - s0, s0 - 0.5f), fmaf(1.702f * a1, 1.f - s1, s1 - 0.5f),
                    fmt);
    }
    tma_stage_acquire(lane);
    sts_row64(stg, lane, pa);
    if (G) sts_row64(stg + EPI_TMA_TILE_BYTES, lane, pg);
    tma_stage_publish();
    if (lane == 0) {
      tma_store_2d(&tmA, stg, col, row0);
      if (G) tma_store_2d(&tmG, stg + EPI_TMA_TILE_BYTES, col, row0);
      tma_store_commit();
    }
  }
  using Aux = NoAux;
  using Col = float4;
  __device__ __forceinline__ Aux load(int, int) const { return {}; }
  __device__ __forceinline__ Col col(int c) const { return load_bias(bias, c); }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux&,
                                             const Col& b) const {
    const float4 h = add4(acc, b);
    const float4 s = make_float4(quick_sigmoid(h.x), quick_sigmoid(h.y), quick_sigmoid(h.z),
                                 quick_sigmoid(h.w));
    const float4 a = make_float4(h.x * s.x, h.y * s.y, h.z * s.z, h.w * s.w);
    if (G) {
      // gelu'(h) = s (1 + 1.702 h (1 - s)) = s + 1.702 a (1 - s), in [-0.1, 1.1]; stored as
      // gelu' - 1/2: |value| <= 0.6, so the 16-bit rounding error is <= 1.5e-4 absolute (fp16) --
      // what rounding the pre-activation itself cost the derivative before
      const float4 g = make_float4(fmaf(1.702f * a.x, 1.f - s.x, s.x - 0.5f),
                                   fmaf(1.702f * a.y, 1.f - s.y, s.y - 0.5f),
                                   fmaf(1.702f * a.z, 1.f - s.z, s.z - 0.5f),
                                   fmaf(1.702f * a.w, 1.f - s.w, s.w - 0.5f));
      *reinterpret_cast<uint2*>(G + (size_t)row * ld + col) = pack4(g, fmt);
    }
    if (Alo) {
      uint2 hi, lo;
      split4(a, hi, lo);
      *reinterpret_cast<uint2*>(A + (size_t)row * ld + col) = hi;
      *reinterpret_cast<uint2*>(Alo + (size_t)row * ld + col) = lo;
    } else {
      *reinterpret_cast<uint2*>(A + (size_t)row * ld + col) = pack4(a, fmt);
    }
  }
};

// out16 = acc * (G16 + 1/2), G = quick_gelu'(pre-activation) - 1/2 as stored by EpiGelu   (dgrad
// through the MLP activation)
struct EpiGeluBwd {
  static constexpr const char* kTag = "gemm.gelu_bwd";
  uint16_t* out;
  int ld;
  const uint16_t* G;
  int fmt_g;
  int fmt_out;
  // TMA path: the G tile of a chunk is fetched into the warp's second staging tile by a TMA load
  // (requested one chunk ahead, the first chunk of a tile before its MMAs retire), the output
  // leaves by a TMA store
  int tma = 0;
  CUtensorMap tmO, tmG;
  int prepare(int M, int N) {
    tma = 0;
    if (epi_tma_level() < 1) return 0;
    if (make_tmap_2d_store16(&tmO, out, M, N, ld)) return 0;
    if (make_tmap_2d_store16(&tmG, G, M, N, ld)) return 0;
    tma = 1;
    return 0;
  }
  __device__ __forceinline__ bool tma_on() const { return tma != 0; }
  __device__ __forceinline__ void tma_request(int row0, int col, uint8_t* stg, TmaCtx& c) const {
    mbar_arrive_expect_tx(c.bar, EPI_TMA_TILE_BYTES);
    tma_load_2d(stg + EPI_TMA_TILE_BYTES, &tmG, c.bar, col, row0);
  }
  __device__ __forceinline__ void tma_prefetch(int row0, int lane, int col, uint8_t* stg,
                                               TmaCtx& c) const {
    if (lane == 0) tma_request(row0, col, stg, c);
  }
  __device__ __forceinline__ void tma_chunk(int row0, int lane, int col, const uint32_t (&r)[32],
                                            uint8_t* stg, TmaCtx& c) const {
    mbar_wait(c.bar, c.phase, 9);
    c.phase ^= 1u;
    uint32_t g[16], pk[16];
    lds_row64(stg + EPI_TMA_TILE_BYTES, lane, g);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 t = unpack2(g[i], fmt_g);   // gelu' - 1/2
      pk[i] = pack2(__uint_as_float(r[2 * i]) * (t.x + 0.5f),
                    __uint_as_float(r[2 * i + 1]) * (t.y + 0.5f), fmt_out);
    }
    __syncwarp();   // every lane has consumed its row of the G tile: it may be overwritten
    if (lane == 0 && c.next_col >= 0) tma_request(row0, c.next_col, stg, c);
    tma_stage_acquire(lane);
    sts_row64(stg, lane, pk);
    tma_stage_publish();
    if (lane == 0) {
      tma_store_2d(&tmO, stg, col, row0);
      tma_store_commit();
    }
  }
  using Aux = uint2;
  using Col = NoAux;
  __device__ __forceinline__ Aux load(int row, int col) const {
    return *reinterpret_cast<const uint2*>(G + (size_t)row * ld + col);
  }
  __device__ __forceinline__ Col col(int) const { return {}; }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux& gu,
                                             const Col&) const {
    const float4 g = unpack4(gu, fmt_g);   // gelu' - 1/2
    const float4 o = make_float4(acc.x * (g.x + 0.5f), acc.y * (g.y + 0.5f), acc.z * (g.z + 0.5f),
                                 acc.w * (g.w + 0.5f));
    *reinterpret_cast<uint2*>(out + (size_t)row * ld + col) = pack4(o, fmt_out);
  }
};

}  // namespace rvlm
