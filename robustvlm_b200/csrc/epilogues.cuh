// Fused GEMM epilogues.  Each functor receives one output row's 32 consecutive
// fp32 accumulators (the thread owns the row: TMEM lane == tile row) and writes
// them to HBM after the fused elementwise work.  Per thread the 32 columns are
// contiguous in memory, so every store instruction fills whole 32-byte sectors.
#pragma once
#include "ptx.cuh"

namespace rvlm {

__device__ __forceinline__ void store16x32(uint16_t* dst, const float (&v)[32], int fmt) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack2(v[8 * j + 0], v[8 * j + 1], fmt);
    u.y = pack2(v[8 * j + 2], v[8 * j + 3], fmt);
    u.z = pack2(v[8 * j + 4], v[8 * j + 5], fmt);
    u.w = pack2(v[8 * j + 6], v[8 * j + 7], fmt);
    d[j] = u;
  }
}

__device__ __forceinline__ void load16x32(const uint16_t* src, float (&v)[32], int fmt) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u = s[j];
    float2 a = unpack2(u.x, fmt), b = unpack2(u.y, fmt), c = unpack2(u.z, fmt),
           d = unpack2(u.w, fmt);
    v[8 * j + 0] = a.x; v[8 * j + 1] = a.y; v[8 * j + 2] = b.x; v[8 * j + 3] = b.y;
    v[8 * j + 4] = c.x; v[8 * j + 5] = c.y; v[8 * j + 6] = d.x; v[8 * j + 7] = d.y;
  }
}

// C16[row, col] = acc + bias[col]
struct EpiStore16 {
  uint16_t* C;
  int ldc;
  const float* bias;  // may be null
  int fmt;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32]) const {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = acc[j] + (bias ? __ldg(bias + col0 + j) : 0.f);
    store16x32(C + (size_t)row * ldc + col0, v, fmt);
  }
};

// Same value written twice in two 16-bit formats: C (fmt) feeds the forward attention,
// C2 (fmt2) is the copy the backward keeps, in the gradient operand format -- a
// tcgen05.mma needs both operands in one format, and the backward's other operand is
// a gradient.
struct EpiStore16Dual {
  uint16_t* C;
  uint16_t* C2;  // may be null
  int ldc;
  const float* bias;
  int fmt, fmt2;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32]) const {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = acc[j] + (bias ? __ldg(bias + col0 + j) : 0.f);
    store16x32(C + (size_t)row * ldc + col0, v, fmt);
    if (C2) store16x32(C2 + (size_t)row * ldc + col0, v, fmt2);
  }
};

// C32[row, col] = acc + bias[col] + resid[row, col]   (fp32 residual stream)
struct EpiStore32 {
  float* C;
  int ldc;
  const float* bias;   // may be null
  const float* resid;  // may be null
  int ldr;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32]) const {
    float4* d = reinterpret_cast<float4*>(C + (size_t)row * ldc + col0);
    const float4* r = resid ? reinterpret_cast<const float4*>(resid + (size_t)row * ldr + col0)
                            : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 o;
      o.x = acc[4 * j + 0]; o.y = acc[4 * j + 1]; o.z = acc[4 * j + 2]; o.w = acc[4 * j + 3];
      if (bias) {
        o.x += __ldg(bias + col0 + 4 * j + 0); o.y += __ldg(bias + col0 + 4 * j + 1);
        o.z += __ldg(bias + col0 + 4 * j + 2); o.w += __ldg(bias + col0 + 4 * j + 3);
      }
      if (r) {
        float4 rr = r[j];
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
      d[j] = o;
    }
  }
};

// QuickGELU (open_clip QuickGELU: x * sigmoid(1.702 x))
__device__ __forceinline__ float quick_gelu(float x) {
  return x / (1.f + __expf(-1.702f * x));
}
__device__ __forceinline__ float quick_gelu_grad(float x) {
  const float s = 1.f / (1.f + __expf(-1.702f * x));
  return s * (1.f + 1.702f * x * (1.f - s));
}

// H16 = acc + bias (pre-activation, kept for the backward); A16 = quick_gelu(H)
struct EpiGelu {
  uint16_t* H;
  uint16_t* A;
  int ld;
  const float* bias;
  int fmt;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32]) const {
    float h[32], a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      h[j] = acc[j] + (bias ? __ldg(bias + col0 + j) : 0.f);
      a[j] = quick_gelu(h[j]);
    }
    if (H) store16x32(H + (size_t)row * ld + col0, h, fmt);
    store16x32(A + (size_t)row * ld + col0, a, fmt);
  }
};

// out16 = acc * quick_gelu'(H16)   (dgrad through the MLP activation)
struct EpiGeluBwd {
  uint16_t* out;
  int ld;
  const uint16_t* H;
  int fmt_h;
  int fmt_out;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32]) const {
    float h[32], o[32];
    load16x32(H + (size_t)row * ld + col0, h, fmt_h);
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = acc[j] * quick_gelu_grad(h[j]);
    store16x32(out + (size_t)row * ld + col0, o, fmt_out);
  }
};

}  // namespace rvlm
