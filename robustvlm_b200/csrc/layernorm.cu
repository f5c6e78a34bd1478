// LayerNorm forward / input-gradient backward for the ViT residual stream
// (open_clip VisionTransformer ln_pre / ln_1 / ln_2 / ln_post, eps = 1e-5;
// SURVEY.md Appendix A.3-5; torch.nn.LayerNorm in the reference's model,
// train/adversarial_training_clip.py:253-257).
//
// HBM-bound: one warp per row, the whole row lives in registers (W/128 float4 per
// lane), two-pass mean/variance in fp32, 128-bit coalesced loads and stores.  The
// forward emits the 16-bit GEMM operand directly; the backward fuses the residual
// add and emits both the fp32 residual gradient and its 16-bit GEMM-operand copy.
#include "gemm.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);
constexpr int LN_WARPS = 8;

// Row pitches in elements (W everywhere for a dense [M, W] matrix).  The last-block shortcut of
// the engine runs LayerNorm on the B class-token rows of a [B*T, W] stream: pitch T*W.
struct LnStrides {
  long long x;    // input rows (forward: x; backward: x)
  long long y;    // forward outputs y16 / y32; backward outputs dx32 / dx16
  long long dy;   // backward: incoming gradient
  long long res;  // backward: residual gradient
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  return v;
}

template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, uint16_t* __restrict__ y16,
              uint16_t* __restrict__ y16lo, float* __restrict__ y32,
              float2* __restrict__ stats, int M, int fmt, LnStrides sd) {
  constexpr int W = NV * 128;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * sd.x);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / W);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / W) + 1e-5f);
  if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = reinterpret_cast<const float4*>(gamma)[i * 32 + lane];
    const float4 b = reinterpret_cast<const float4*>(beta)[i * 32 + lane];
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (y32) reinterpret_cast<float4*>(y32 + (size_t)row * sd.y)[i * 32 + lane] = o;
    if (y16) {
      uint2 p;
      p.x = pack2(o.x, o.y, fmt);
      p.y = pack2(o.z, o.w, fmt);
      reinterpret_cast<uint2*>(y16 + (size_t)row * sd.y)[i * 32 + lane] = p;
      if (y16lo) {   // split-operand forward: lo = fp16(y - hi)   (fmt == fp16)
        const float2 a = unpack2(p.x, 0), b = unpack2(p.y, 0);
        uint2 q;
        q.x = pack2(o.x - a.x, o.y - a.y, 0);
        q.y = pack2(o.z - b.x, o.w - b.y, 0);
        reinterpret_cast<uint2*>(y16lo + (size_t)row * sd.y)[i * 32 + lane] = q;
      }
    }
  }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// drop_T > 0: rows with (row % drop_T == 0) (the CLS token of each image) are not
// written and the remaining rows are compacted (used by the ln_pre backward, whose
// consumer is the patch-embedding dgrad over patch rows only).
// DY16: dy is a 16-bit matrix in `fmt` (the dgrad GEMM's epilogue output on the attack path:
// 2 B/elem written + 2 read instead of 4 + 4).
template <int NV, bool DY16>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x,
              const float2* __restrict__ stats, const float* __restrict__ gamma,
              const float* dres, float* dx32,
              uint16_t* __restrict__ dx16, int M, int fmt, int drop_T, LnStrides sd) {
  constexpr int W = NV * 128;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  int orow = row;
  if (drop_T > 0) {
    const int b = row / drop_T, t = row - b * drop_T;
    if (t == 0) return;
    orow = row - b - 1;
  }
  const float2 st = stats[row];
  const float mean = st.x, rstd = st.y;
  float4 g[NV], xh[NV], res[NV];
  float s1 = 0.f, s2 = 0.f;
  // every HBM read of the row (dy, x, residual gradient) is requested before the first
  // reduction, so one round of memory latency covers all three streams
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (DY16) {
      const uint2 u =
          reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(dy_) + (size_t)row * sd.dy)[i * 32 + lane];
      const float2 a = unpack2(u.x, fmt), b = unpack2(u.y, fmt);
      g[i] = make_float4(a.x, a.y, b.x, b.y);
    } else {
      g[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)row * sd.dy)[i * 32 + lane];
    }
    xh[i] = reinterpret_cast<const float4*>(x + (size_t)row * sd.x)[i * 32 + lane];
  }
  if (dres) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      res[i] = reinterpret_cast<const float4*>(dres + (size_t)row * sd.res)[i * 32 + lane];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 d = g[i];
    const float4 xv = xh[i];
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    g[i].x = d.x * gm.x; g[i].y = d.y * gm.y; g[i].z = d.z * gm.z; g[i].w = d.w * gm.w;
    xh[i].x = (xv.x - mean) * rstd; xh[i].y = (xv.y - mean) * rstd;
    xh[i].z = (xv.z - mean) * rstd; xh[i].w = (xv.w - mean) * rstd;
    s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
  }
  const float c1 = warp_sum(s1) * (1.f / W);
  const float c2 = warp_sum(s2) * (1.f / W);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 o;
    o.x = rstd * (g[i].x - c1 - xh[i].x * c2);
    o.y = rstd * (g[i].y - c1 - xh[i].y * c2);
    o.z = rstd * (g[i].z - c1 - xh[i].z * c2);
    o.w = rstd * (g[i].w - c1 - xh[i].w * c2);
    if (dres) {
      const float4 r = res[i];
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (dx32) reinterpret_cast<float4*>(dx32 + (size_t)orow * sd.y)[i * 32 + lane] = o;
    if (dx16) {
      uint2 p;
      p.x = pack2(o.x, o.y, fmt);
      p.y = pack2(o.z, o.w, fmt);
      reinterpret_cast<uint2*>(dx16 + (size_t)orow * sd.y)[i * 32 + lane] = p;
    }
  }
}

}  // namespace rvlm

using namespace rvlm;

namespace rvlm {
// strided entry points (engine-internal); pitches in elements, 0 = dense (W)
int ln_fwd_strided(const float* x, long long ldx, const float* gamma, const float* beta, void* y16,
                   float* y32, long long ldy, float* stats, int M, int W, int fmt,
                   cudaStream_t st, void* y16lo_ = nullptr) {
  if (M <= 0) return RVLM_ERR_INVAL;
  if (y16lo_ && fmt != 0) return RVLM_ERR_INVAL;
  const dim3 grid((M + LN_WARPS - 1) / LN_WARPS), block(LN_WARPS * 32);
  uint16_t* y = reinterpret_cast<uint16_t*>(y16);
  uint16_t* ylo = reinterpret_cast<uint16_t*>(y16lo_);
  float2* s = reinterpret_cast<float2*>(stats);
  const LnStrides sd{ldx ? ldx : W, ldy ? ldy : W, 0, 0};
  KernelProfScope prof(st, "ln_fwd");
  switch (W) {
    case 128: ln_fwd_kernel<1><<<grid, block, 0, st>>>(x, gamma, beta, y, ylo, y32, s, M, fmt, sd); break;
    case 256: ln_fwd_kernel<2><<<grid, block, 0, st>>>(x, gamma, beta, y, ylo, y32, s, M, fmt, sd); break;
    case 768: ln_fwd_kernel<6><<<grid, block, 0, st>>>(x, gamma, beta, y, ylo, y32, s, M, fmt, sd); break;
    case 1024: ln_fwd_kernel<8><<<grid, block, 0, st>>>(x, gamma, beta, y, ylo, y32, s, M, fmt, sd); break;
    default: return RVLM_ERR_UNSUPPORTED;
  }
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

template <bool DY16>
int ln_bwd_strided_t(const void* dy, long long lddy, const float* x, long long ldx,
                     const float* stats, const float* gamma, const float* dres, long long ldres,
                     float* dx32, void* dx16, long long ldo, int M, int W, int fmt, int drop_T,
                     cudaStream_t st) {
  if (M <= 0) return RVLM_ERR_INVAL;
  const dim3 grid((M + LN_WARPS - 1) / LN_WARPS), block(LN_WARPS * 32);
  uint16_t* d16 = reinterpret_cast<uint16_t*>(dx16);
  const float2* s = reinterpret_cast<const float2*>(stats);
  const LnStrides sd{ldx ? ldx : W, ldo ? ldo : W, lddy ? lddy : W, ldres ? ldres : W};
  KernelProfScope prof(st, "ln_bwd");
  switch (W) {
    case 128: ln_bwd_kernel<1, DY16><<<grid, block, 0, st>>>(dy, x, s, gamma, dres, dx32, d16, M, fmt, drop_T, sd); break;
    case 256: ln_bwd_kernel<2, DY16><<<grid, block, 0, st>>>(dy, x, s, gamma, dres, dx32, d16, M, fmt, drop_T, sd); break;
    case 768: ln_bwd_kernel<6, DY16><<<grid, block, 0, st>>>(dy, x, s, gamma, dres, dx32, d16, M, fmt, drop_T, sd); break;
    case 1024: ln_bwd_kernel<8, DY16><<<grid, block, 0, st>>>(dy, x, s, gamma, dres, dx32, d16, M, fmt, drop_T, sd); break;
    default: return RVLM_ERR_UNSUPPORTED;
  }
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}
int ln_bwd16_strided(const void* dy16, long long lddy, const float* x, long long ldx,
                     const float* stats, const float* gamma, const float* dres, long long ldres,
                     float* dx32, void* dx16, long long ldo, int M, int W, int fmt,
                     cudaStream_t st) {
  return ln_bwd_strided_t<true>(dy16, lddy, x, ldx, stats, gamma, dres, ldres, dx32, dx16, ldo, M,
                                W, fmt, 0, st);
}
}  // namespace rvlm

extern "C" int rvlm_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y16,
                                  float* y32, float* stats, int M, int W, int fmt, void* stream) {
  return ln_fwd_strided(x, 0, gamma, beta, y16, y32, 0, stats, M, W, fmt, (cudaStream_t)stream);
}

extern "C" int rvlm_layernorm_bwd(const float* dy, const float* x, const float* stats,
                                  const float* gamma, const float* dres, float* dx32, void* dx16,
                                  int M, int W, int fmt, int drop_T, void* stream) {
  return ln_bwd_strided_t<false>(dy, 0, x, 0, stats, gamma, dres, 0, dx32, dx16, 0, M, W, fmt,
                                 drop_T, (cudaStream_t)stream);
}

extern "C" int rvlm_layernorm_bwd16(const void* dy16, const float* x, const float* stats,
                                    const float* gamma, const float* dres, float* dx32, void* dx16,
                                    int M, int W, int fmt, int drop_T, void* stream) {
  return ln_bwd_strided_t<true>(dy16, 0, x, 0, stats, gamma, dres, 0, dx32, dx16, 0, M, W, fmt,
                                drop_T, (cudaStream_t)stream);
}
