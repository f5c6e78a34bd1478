// CUDA-core (SIMT, fp32 math) multi-head attention forward / backward.
//
// NOT the production path: this is the simple, obviously-correct on-device
// checker for the tcgen05 attention kernels (attention.cu) and the bring-up path
// of the engine (engine flag RVLM_ATTN_SIMT).  Same I/O contract as attention.cu:
//   qkv16 [B*T, 3W]  (q | k | v per token, head h at columns h*64..h*64+63)
//   o16   [B*T, W],  lse [B, H, T] (natural log of the softmax denominator, scaled scores)
// Semantics: softmax(q k^T / sqrt(64)) v per head, no mask, no dropout
// (nn.MultiheadAttention inside open_clip's ResidualAttentionBlock).
#include "gemm.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);
constexpr int HD = 64;       // head dim
constexpr int ROWP = 66;     // padded smem row (16-bit elements) -> conflict-free column reads

__device__ __forceinline__ float ld16(const uint16_t* p, int fmt) {
  if (fmt) return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
  return __half2float(*reinterpret_cast<const __half*>(p));
}
__device__ __forceinline__ uint16_t st16(float v, int fmt) {
  if (fmt) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffff, v, o));
  return v;
}

// load a [T,64] head slice (row pitch ld) into padded smem
__device__ void load_head(uint16_t* dst, const uint16_t* src, int T, int ld) {
  for (int i = threadIdx.x; i < T * (HD / 2); i += blockDim.x) {
    const int t = i / (HD / 2), c = i % (HD / 2);
    reinterpret_cast<uint32_t*>(dst + t * ROWP)[c] =
        reinterpret_cast<const uint32_t*>(src + (size_t)t * ld)[c];
  }
}

__global__ void attn_fwd_simt_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ o,
                                     float* __restrict__ lse, int T, int H, int fmt, int t_begin,
                                     int t_end) {
  extern __shared__ uint8_t sm[];
  const int W = H * HD;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  uint16_t* sK = reinterpret_cast<uint16_t*>(sm);
  uint16_t* sV = sK + T * ROWP;
  float* sP = reinterpret_cast<float*>(sV + T * ROWP);  // [nwarps][T]
  float* sQ = sP + (blockDim.x >> 5) * T;               // [nwarps][64]
  const uint16_t* base = qkv + (size_t)b * T * 3 * W + h * HD;
  load_head(sK, base + W, T, 3 * W);
  load_head(sV, base + 2 * W, T, 3 * W);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* p = sP + warp * T;
  float* q = sQ + warp * HD;
  const float scale = 0.125f;
  for (int t = t_begin + warp; t < t_end; t += nw) {
    const uint16_t* qp = base + (size_t)t * 3 * W;
    q[lane] = ld16(qp + lane, fmt);
    q[lane + 32] = ld16(qp + lane + 32, fmt);
    __syncwarp();
    float mx = -INFINITY;
    for (int kv = lane; kv < T; kv += 32) {
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) s += q[d] * ld16(sK + kv * ROWP + d, fmt);
      s *= scale;
      p[kv] = s;
      mx = fmaxf(mx, s);
    }
    mx = wmax(mx);
    float sum = 0.f;
    for (int kv = lane; kv < T; kv += 32) {
      const float e = __expf(p[kv] - mx);
      p[kv] = e;
      sum += e;
    }
    sum = wsum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int kv = 0; kv < T; ++kv) {
      const float pv = p[kv];
      o0 += pv * ld16(sV + kv * ROWP + 2 * lane, fmt);
      o1 += pv * ld16(sV + kv * ROWP + 2 * lane + 1, fmt);
    }
    const float inv = 1.f / sum;
    uint16_t* op = o + ((size_t)b * T + t) * W + h * HD;
    op[2 * lane] = st16(o0 * inv, fmt);
    op[2 * lane + 1] = st16(o1 * inv, fmt);
    if (lane == 0) lse[((size_t)b * H + h) * T + t] = mx + logf(sum);
    __syncwarp();
  }
}

// dQ, dK, dV from Q, K, V, O, dO, lse.  Two deterministic passes (row-wise for dQ,
// column-wise for dK / dV), scores recomputed in each.
__global__ void attn_bwd_simt_kernel(const uint16_t* __restrict__ qkv,
                                     const uint16_t* __restrict__ o,
                                     const uint16_t* __restrict__ d_o, uint16_t* __restrict__ dqkv,
                                     int T, int H, int fmt, int fmt_o, int fmt_g) {
  extern __shared__ uint8_t sm[];
  const int W = H * HD;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  uint16_t* sQ = reinterpret_cast<uint16_t*>(sm);
  uint16_t* sK = sQ + T * ROWP;
  uint16_t* sV = sK + T * ROWP;
  uint16_t* sdO = sV + T * ROWP;
  float* sL = reinterpret_cast<float*>(sdO + T * ROWP);  // [T] lse
  float* sD = sL + T;                                     // [T] rowsum(dO * O)
  float* sP = sD + T;                                     // [nwarps][2][T]
  float* sR = sP + (blockDim.x >> 5) * 2 * T;             // [nwarps][128] row vectors
  const uint16_t* base = qkv + (size_t)b * T * 3 * W + h * HD;
  load_head(sQ, base, T, 3 * W);
  load_head(sK, base + W, T, 3 * W);
  load_head(sV, base + 2 * W, T, 3 * W);
  load_head(sdO, d_o + (size_t)b * T * W + h * HD, T, W);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  const float scale = 0.125f;
  float* pp = sP + warp * 2 * T;
  float* ds = pp + T;
  float* ra = sR + warp * 128;
  float* rb = ra + 64;
  // softmax statistics recomputed from the operands given (self-consistent with P below)
  for (int t = warp; t < T; t += nw) {
    const uint16_t* op = o + ((size_t)b * T + t) * W + h * HD;
    float s = ld16(op + lane, fmt_o) * ld16(sdO + t * ROWP + lane, fmt_g) +
              ld16(op + lane + 32, fmt_o) * ld16(sdO + t * ROWP + lane + 32, fmt_g);
    s = wsum(s);
    if (lane == 0) sD[t] = s;
    ra[lane] = ld16(sQ + t * ROWP + lane, fmt);
    ra[lane + 32] = ld16(sQ + t * ROWP + lane + 32, fmt);
    __syncwarp();
    float mx = -INFINITY;
    for (int kv = lane; kv < T; kv += 32) {
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) a += ra[d] * ld16(sK + kv * ROWP + d, fmt);
      pp[kv] = a * scale;
      mx = fmaxf(mx, a * scale);
    }
    mx = wmax(mx);
    float sum = 0.f;
    for (int kv = lane; kv < T; kv += 32) sum += __expf(pp[kv] - mx);
    sum = wsum(sum);
    if (lane == 0) sL[t] = mx + logf(sum);
    __syncwarp();
  }
  __syncthreads();
  // ---- pass A: one q row per warp -> dQ
  for (int t = warp; t < T; t += nw) {
    ra[lane] = ld16(sQ + t * ROWP + lane, fmt);
    ra[lane + 32] = ld16(sQ + t * ROWP + lane + 32, fmt);
    rb[lane] = ld16(sdO + t * ROWP + lane, fmt_g);
    rb[lane + 32] = ld16(sdO + t * ROWP + lane + 32, fmt_g);
    __syncwarp();
    const float L = sL[t], Dv = sD[t];
    for (int kv = lane; kv < T; kv += 32) {
      float s = 0.f, dp = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) {
        s += ra[d] * ld16(sK + kv * ROWP + d, fmt);
        dp += rb[d] * ld16(sV + kv * ROWP + d, fmt);
      }
      const float pr = __expf(s * scale - L);
      ds[kv] = pr * (dp - Dv) * scale;
    }
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int kv = 0; kv < T; ++kv) {
      const float w = ds[kv];
      a0 += w * ld16(sK + kv * ROWP + 2 * lane, fmt);
      a1 += w * ld16(sK + kv * ROWP + 2 * lane + 1, fmt);
    }
    uint16_t* dq = dqkv + ((size_t)b * T + t) * 3 * W + h * HD;
    dq[2 * lane] = st16(a0, fmt_g);
    dq[2 * lane + 1] = st16(a1, fmt_g);
    __syncwarp();
  }
  // ---- pass B: one kv row per warp -> dK, dV
  for (int kv = warp; kv < T; kv += nw) {
    ra[lane] = ld16(sK + kv * ROWP + lane, fmt);
    ra[lane + 32] = ld16(sK + kv * ROWP + lane + 32, fmt);
    rb[lane] = ld16(sV + kv * ROWP + lane, fmt);
    rb[lane + 32] = ld16(sV + kv * ROWP + lane + 32, fmt);
    __syncwarp();
    for (int t = lane; t < T; t += 32) {
      float s = 0.f, dp = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) {
        s += ra[d] * ld16(sQ + t * ROWP + d, fmt);
        dp += rb[d] * ld16(sdO + t * ROWP + d, fmt_g);
      }
      const float pr = __expf(s * scale - sL[t]);
      pp[t] = pr;
      ds[t] = pr * (dp - sD[t]) * scale;
    }
    __syncwarp();
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int t = 0; t < T; ++t) {
      const float w = ds[t], pr = pp[t];
      k0 += w * ld16(sQ + t * ROWP + 2 * lane, fmt);
      k1 += w * ld16(sQ + t * ROWP + 2 * lane + 1, fmt);
      v0 += pr * ld16(sdO + t * ROWP + 2 * lane, fmt_g);
      v1 += pr * ld16(sdO + t * ROWP + 2 * lane + 1, fmt_g);
    }
    uint16_t* dk = dqkv + ((size_t)b * T + kv) * 3 * W + W + h * HD;
    uint16_t* dv = dk + W;
    dk[2 * lane] = st16(k0, fmt_g);
    dk[2 * lane + 1] = st16(k1, fmt_g);
    dv[2 * lane] = st16(v0, fmt_g);
    dv[2 * lane + 1] = st16(v1, fmt_g);
    __syncwarp();
  }
}

}  // namespace rvlm

using namespace rvlm;

extern "C" int rvlm_attn_fwd_simt_rows(const void* qkv16, void* o16, float* lse, int B, int T,
                                       int H, int fmt, int t_begin, int t_end, void* stream);

extern "C" int rvlm_attn_fwd_simt(const void* qkv16, void* o16, float* lse, int B, int T, int H,
                                  int fmt, void* stream) {
  return rvlm_attn_fwd_simt_rows(qkv16, o16, lse, B, T, H, fmt, 0, T, stream);
}

// query rows [t_begin, t_end) only (all keys)
extern "C" int rvlm_attn_fwd_simt_rows(const void* qkv16, void* o16, float* lse, int B, int T,
                                       int H, int fmt, int t_begin, int t_end, void* stream) {
  const int threads = 128, nw = threads / 32;
  const size_t smem = (size_t)2 * T * ROWP * 2 + (size_t)nw * T * 4 + (size_t)nw * HD * 4;
  static DeviceMax set;
  if (set.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_simt_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
  }
  attn_fwd_simt_kernel<<<B * H, threads, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint16_t*>(qkv16), reinterpret_cast<uint16_t*>(o16), lse, T, H, fmt,
      t_begin, t_end);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

extern "C" int rvlm_attn_bwd_simt(const void* qkv16, const void* o16, const void* do16,
                                  float* lse_tmp, void* dqkv16, int B, int T, int H, int fmt,
                                  int fmt_o, int fmt_g, void* stream) {
  (void)lse_tmp;
  const int threads = 256, nw = threads / 32;
  const size_t smem = (size_t)4 * T * ROWP * 2 + (size_t)2 * T * 4 + (size_t)nw * 2 * T * 4 +
                      (size_t)nw * 128 * 4;
  static DeviceMax set;
  if (set.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_simt_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
  }
  attn_bwd_simt_kernel<<<B * H, threads, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint16_t*>(qkv16), reinterpret_cast<const uint16_t*>(o16),
      reinterpret_cast<const uint16_t*>(do16), reinterpret_cast<uint16_t*>(dqkv16), T, H, fmt,
      fmt_o, fmt_g);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}
