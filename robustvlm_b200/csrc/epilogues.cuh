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

// ---- direct (row-per-lane) epilogue pieces: 16 consecutive columns of one row -------------------
// 256-bit global accesses: one full 32-byte sector per lane and instruction
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                 "=r"(v[7])
               : "l"(p)
               : "memory");
}
// the bias of 16 consecutive columns (the same addresses in every lane: broadcast loads)
__device__ __forceinline__ void load_bias16(const float* bias, int col, float (&b)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = load_bias(bias, col + 4 * i);
    b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
  }
}
// acc + bias of 16 accumulator columns
__device__ __forceinline__ void acc_bias16(const uint32_t (&r)[16], const float* bias, int col,
                                           float (&v)[16]) {
  float b[16];
  load_bias16(bias, col, b);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + b[i];
}
__device__ __forceinline__ void pack16(const float (&v)[16], int fmt, uint32_t (&pk)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack2(v[2 * i], v[2 * i + 1], fmt);
}
// fp16 hi + lo pairs of 16 values
__device__ __forceinline__ void split16(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hi[i] = pack2(v[2 * i], v[2 * i + 1], 0);
    const float2 h = unpack2(hi[i], 0);
    lo[i] = pack2(v[2 * i] - h.x, v[2 * i + 1] - h.y, 0);
  }
}
struct NoAux16 {};

// C16[row, col] = acc + bias[col]
struct NoAux {};

struct EpiStore16 {
  static constexpr const char* kTag = "gemm.store16";
  uint16_t* C;
  int ldc;
  const float* bias;  // may be null
  int fmt;
  using Aux = NoAux;
  using Col = float4;
  __device__ __forceinline__ Aux load(int, int) const { return {}; }
  __device__ __forceinline__ Col col(int c) const { return load_bias(bias, c); }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux&,
                                             const Col& b) const {
    *reinterpret_cast<uint2*>(C + (size_t)row * ldc + col) = pack4(add4(acc, b), fmt);
  }
  static constexpr bool kDirect = true;
  using Aux16 = NoAux16;
  __device__ __forceinline__ Aux16 load16(int, int) const { return {}; }
  __device__ __forceinline__ void direct(int row, int col, const uint32_t (&r)[16],
                                         const Aux16&) const {
    float v[16];
    acc_bias16(r, bias, col, v);
    uint32_t pk[8];
    pack16(v, fmt, pk);
    st_global_v8(C + (size_t)row * ldc + col, pk);
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
  static constexpr bool kDirect = true;
  using Aux16 = NoAux16;
  __device__ __forceinline__ Aux16 load16(int, int) const { return {}; }
  __device__ __forceinline__ void direct(int row, int col, const uint32_t (&r)[16],
                                         const Aux16&) const {
    float v[16];
    acc_bias16(r, bias, col, v);
    uint32_t pk[8];
    const size_t off = (size_t)row * ldc + col;
    if (Clo) {
      uint32_t lo[8];
      split16(v, pk, lo);
      st_global_v8(C + off, pk);
      st_global_v8(Clo + off, lo);
    } else {
      pack16(v, fmt, pk);
      st_global_v8(C + off, pk);
    }
    if (C2) {
      pack16(v, fmt2, pk);
      st_global_v8(C2 + off, pk);
    }
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
  static constexpr bool kDirect = true;
  struct Aux16 { uint32_t a[8], b[8]; };
  __device__ __forceinline__ Aux16 load16(int row, int col) const {
    Aux16 x;
    if (resid) {
      const float* p = resid + (size_t)row * ldr + col;
      ld_global_v8(p, x.a);
      ld_global_v8(p + 8, x.b);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x.a[i] = x.b[i] = 0u;
    }
    return x;
  }
  __device__ __forceinline__ void direct(int row, int col, const uint32_t (&r)[16],
                                         const Aux16& x) const {
    float v[16];
    acc_bias16(r, bias, col, v);
    uint32_t o0[8], o1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o0[i] = __float_as_uint(v[i] + __uint_as_float(x.a[i]));
      o1[i] = __float_as_uint(v[8 + i] + __uint_as_float(x.b[i]));
    }
    float* p = C + (size_t)row * ldc + col;
    st_global_v8(p, o0);
    st_global_v8(p + 8, o1);
  }
};

// QuickGELU (open_clip QuickGELU: x * sigmoid(1.702 x)), with
// MUFU ex2 / rcp approximations (relative error ~1e-7, far below the 16-bit output).
__device__ __forceinline__ float quick_sigmoid(float x) {
  return fast_rcp(1.f + fast_ex2(-2.4554669595930157f * x));  // 1.702 * log2(e)
}

// RVLM_GELU_DERIV (build-time, A/B): 1 = the forward stores QuickGELU'(h); 0 = it stores h and the
// dgrad epilogue evaluates the derivative (the earlier scheme)
#ifndef RVLM_GELU_DERIV
#define RVLM_GELU_DERIV 1
#endif

// A16 = quick_gelu(h), h = acc + bias; G16 = quick_gelu'(h), the only thing the backward needs of
// the pre-activation: s = sigmoid(1.702 h) is at hand here, so the derivative costs three FMAs
// in this epilogue and the dgrad epilogue (EpiGeluBwd) is left with one multiply -- no MUFU work
// and no second evaluation of the sigmoid on the backward's critical path.
struct EpiGelu {
  static constexpr const char* kTag = "gemm.gelu";
  uint16_t* G;  // may be null
  uint16_t* A;
  int ld;
  const float* bias;
  int fmt;
  uint16_t* Alo;  // may be null: fp16 lo half of A (split-operand forward; fmt == 0)
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
      // gelu'(h) = s (1 + 1.702 h (1 - s)) = s + 1.702 a (1 - s)
#if !RVLM_GELU_DERIV
      const float4 g = h;
#else
      const float4 g = make_float4(fmaf(1.702f * a.x, 1.f - s.x, s.x), fmaf(1.702f * a.y, 1.f - s.y, s.y),
                                   fmaf(1.702f * a.z, 1.f - s.z, s.z), fmaf(1.702f * a.w, 1.f - s.w, s.w));
#endif
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
  static constexpr bool kDirect = true;
  using Aux16 = NoAux16;
  __device__ __forceinline__ Aux16 load16(int, int) const { return {}; }
  __device__ __forceinline__ void direct(int row, int col, const uint32_t (&r)[16],
                                         const Aux16&) const {
    float h[16], a[16];
    acc_bias16(r, bias, col, h);
    uint32_t pk[8];
    const size_t off = (size_t)row * ld + col;
#if RVLM_GELU_DERIV
    float g[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float s = quick_sigmoid(h[i]);
      a[i] = h[i] * s;
      g[i] = fmaf(1.702f * a[i], 1.f - s, s);
    }
    if (G) {
      pack16(g, fmt, pk);
      st_global_v8(G + off, pk);
    }
#else
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = h[i] * quick_sigmoid(h[i]);
    if (G) {
      pack16(h, fmt, pk);
      st_global_v8(G + off, pk);
    }
#endif
    if (Alo) {
      uint32_t lo[8];
      split16(a, pk, lo);
      st_global_v8(A + off, pk);
      st_global_v8(Alo + off, lo);
    } else {
      pack16(a, fmt, pk);
      st_global_v8(A + off, pk);
    }
  }
};

// out16 = acc * G16, G = quick_gelu'(pre-activation) as stored by EpiGelu   (dgrad through the
// MLP activation)
struct EpiGeluBwd {
  static constexpr const char* kTag = "gemm.gelu_bwd";
  uint16_t* out;
  int ld;
  const uint16_t* G;
  int fmt_g;
  int fmt_out;
  using Aux = uint2;
  using Col = NoAux;
  __device__ __forceinline__ Aux load(int row, int col) const {
    return *reinterpret_cast<const uint2*>(G + (size_t)row * ld + col);
  }

  __device__ __forceinline__ Col col(int) const { return {}; }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux& gu,
                                             const Col&) const {
#if RVLM_GELU_DERIV
    const float4 g = unpack4(gu, fmt_g);
#else
    const float4 h = unpack4(gu, fmt_g);
    const float4 sg = make_float4(quick_sigmoid(h.x), quick_sigmoid(h.y), quick_sigmoid(h.z),
                                  quick_sigmoid(h.w));
    const float4 g = make_float4(sg.x * (1.f + 1.702f * h.x * (1.f - sg.x)),
                                 sg.y * (1.f + 1.702f * h.y * (1.f - sg.y)),
                                 sg.z * (1.f + 1.702f * h.z * (1.f - sg.z)),
                                 sg.w * (1.f + 1.702f * h.w * (1.f - sg.w)));
#endif
    const float4 o = make_float4(acc.x * g.x, acc.y * g.y, acc.z * g.z, acc.w * g.w);
    *reinterpret_cast<uint2*>(out + (size_t)row * ld + col) = pack4(o, fmt_out);
  }
  static constexpr bool kDirect = true;
  struct Aux16 { uint32_t g[8]; };
  __device__ __forceinline__ Aux16 load16(int row, int col) const {
    Aux16 x;
    ld_global_v8(G + (size_t)row * ld + col, x.g);
    return x;
  }
  __device__ __forceinline__ void direct(int row, int col, const uint32_t (&r)[16],
                                         const Aux16& x) const {
    float v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 t = unpack2(x.g[i], fmt_g);
#if RVLM_GELU_DERIV
      const float g0 = t.x, g1 = t.y;
#else
      const float s0 = quick_sigmoid(t.x), s1 = quick_sigmoid(t.y);
      const float g0 = s0 * (1.f + 1.702f * t.x * (1.f - s0));
      const float g1 = s1 * (1.f + 1.702f * t.y * (1.f - s1));
#endif
      v[2 * i] = __uint_as_float(r[2 * i]) * g0;
      v[2 * i + 1] = __uint_as_float(r[2 * i + 1]) * g1;
    }
    uint32_t pk[8];
    pack16(v, fmt_out, pk);
    st_global_v8(out + (size_t)row * ld + col, pk);
  }
};

}  // namespace rvlm
