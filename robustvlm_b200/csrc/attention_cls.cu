// Class-token attention of the LAST transformer block (forward and input-gradient backward).
//
// The vision tower's output is ln_post(x_L[:, 0, :]) @ proj (open_clip VisionTransformer,
// SURVEY.md Appendix A.6; reference call site train/adversarial_training_clip.py:253-257), so
// of the last block only the class-token row of the attention output -- and, in the backward,
// only the gradient arriving at that row -- is ever used.  The keys / values of ALL tokens
// still matter.  These two kernels evaluate exactly that row per (image, head) on the CUDA
// cores (257 x 64 MACs each way; the work is ~1/257 of the full attention), which lets the
// engine skip the last block's out-proj / MLP GEMMs and LayerNorms for the other 256 rows
// (declared shortcut, SURVEY.md §8d: the roofline still counts the full algorithmic FLOPs).
//
//   forward : o[b*T + 0, h*64 ..] = softmax(q_0 K^T / 8) V
//   backward: given dO_0:  p = softmax(q_0 K^T / 8),  dP = dO_0 V^T,  D = p . dP,
//             dS = p (dP - D) / 8;   dV[kv] = p[kv] dO_0,  dK[kv] = dS[kv] q_0,
//             dQ[0] = dS K,  dQ[t > 0] = 0      (written densely into dqkv [B*T, 3W])
#include "ptx.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);

namespace {

constexpr int CLS_THREADS = 128;
constexpr int CLS_TMAX = 288;

__device__ __forceinline__ void load_vec64(const uint16_t* p, int fmt, float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = reinterpret_cast<const uint4*>(p)[c];
    const uint32_t* a = &u.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack2(a[i], fmt);
      v[c * 8 + 2 * i] = x.x;
      v[c * 8 + 2 * i + 1] = x.y;
    }
  }
}
__device__ __forceinline__ float dot_row64(const uint16_t* p, int fmt, const float (&v)[64]) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = reinterpret_cast<const uint4*>(p)[c];
    const uint32_t* a = &u.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack2(a[i], fmt);
      a0 = fmaf(x.x, v[c * 8 + 2 * i], a0);
      a1 = fmaf(x.y, v[c * 8 + 2 * i + 1], a1);
    }
  }
  return a0 + a1;
}
__device__ __forceinline__ void store_scaled64(uint16_t* p, int fmt, float w, const float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 u;
    u.x = pack2(w * v[c * 8 + 0], w * v[c * 8 + 1], fmt);
    u.y = pack2(w * v[c * 8 + 2], w * v[c * 8 + 3], fmt);
    u.z = pack2(w * v[c * 8 + 4], w * v[c * 8 + 5], fmt);
    u.w = pack2(w * v[c * 8 + 6], w * v[c * 8 + 7], fmt);
    reinterpret_cast<uint4*>(p)[c] = u;
  }
}
__device__ __forceinline__ float ld1f(const uint16_t* p, int fmt) {
  return fmt ? __uint_as_float((uint32_t)*p << 16)
             : __half2float(*reinterpret_cast<const __half*>(p));
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffff, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const float t = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  return t;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const float t = (red[0] + red[1]) + (red[2] + red[3]);
  __syncthreads();
  return t;
}

// qkv_lo / o_lo non-null: split-operand forward (fp16 hi + lo pairs; x = hi + lo in fp32)
__global__ void __launch_bounds__(CLS_THREADS)
attn_cls_fwd_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ qkv_lo,
                    uint16_t* __restrict__ o, uint16_t* __restrict__ o_lo, int T, int H, int fmt) {
  __shared__ float sc[CLS_TMAX];
  __shared__ float red[4];
  __shared__ float oh[2][64];
  const int W = H * 64, tid = threadIdx.x;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const size_t boff = (size_t)b * T * 3 * W + h * 64;
  const uint16_t* base = qkv + boff;
  float q[64];
  load_vec64(base, fmt, q);
  if (qkv_lo) {
    float ql[64];
    load_vec64(qkv_lo + boff, 0, ql);
#pragma unroll
    for (int i = 0; i < 64; ++i) q[i] += ql[i];
  }
  float lmax = -INFINITY;
  for (int kv = tid; kv < T; kv += CLS_THREADS) {
    float s = dot_row64(base + (size_t)kv * 3 * W + W, fmt, q);
    if (qkv_lo) s += dot_row64(qkv_lo + boff + (size_t)kv * 3 * W + W, 0, q);
    s *= 0.125f;
    sc[kv] = s;
    lmax = fmaxf(lmax, s);
  }
  const float mx = block_max(lmax, red);
  float lsum = 0.f;
  for (int kv = tid; kv < T; kv += CLS_THREADS) {
    const float e = __expf(sc[kv] - mx);
    sc[kv] = e;
    lsum += e;
  }
  const float sum = block_sum(lsum, red);   // (its barriers also publish sc[])
  const int d = tid & 63, hf = tid >> 6;
  const int half = (T + 1) / 2;
  const int k0 = hf * half, k1 = min(T, k0 + half);
  const uint16_t* vp = base + 2 * W + d;
  float a0 = 0.f, a1 = 0.f;
  int kv = k0;
  for (; kv + 1 < k1; kv += 2) {
    a0 = fmaf(sc[kv], ld1f(vp + (size_t)kv * 3 * W, fmt), a0);
    a1 = fmaf(sc[kv + 1], ld1f(vp + (size_t)(kv + 1) * 3 * W, fmt), a1);
  }
  if (kv < k1) a0 = fmaf(sc[kv], ld1f(vp + (size_t)kv * 3 * W, fmt), a0);
  if (qkv_lo) {
    const uint16_t* vl = qkv_lo + boff + 2 * W + d;
    for (kv = k0; kv < k1; ++kv) a1 = fmaf(sc[kv], ld1f(vl + (size_t)kv * 3 * W, 0), a1);
  }
  oh[hf][d] = a0 + a1;
  __syncthreads();
  if (tid < 32) {
    const float inv = 1.f / sum;
    const float o0 = (oh[0][2 * tid] + oh[1][2 * tid]) * inv;
    const float o1 = (oh[0][2 * tid + 1] + oh[1][2 * tid + 1]) * inv;
    const uint32_t hi = pack2(o0, o1, fmt);
    reinterpret_cast<uint32_t*>(o + (size_t)b * T * W + h * 64)[tid] = hi;
    if (o_lo) {
      const float2 hf2 = unpack2(hi, 0);
      reinterpret_cast<uint32_t*>(o_lo + (size_t)b * T * W + h * 64)[tid] =
          pack2(o0 - hf2.x, o1 - hf2.y, 0);
    }
  }
}

__global__ void __launch_bounds__(CLS_THREADS)
attn_cls_bwd_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ do_cls,
                    uint16_t* __restrict__ dqkv, int T, int H, int fmt) {
  __shared__ float sp[CLS_TMAX];    // scores -> probabilities
  __shared__ float sds[CLS_TMAX];   // dP -> dS
  __shared__ float red[4];
  __shared__ float oh[2][64];
  const int W = H * 64, tid = threadIdx.x;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const uint16_t* base = qkv + (size_t)b * T * 3 * W + h * 64;
  uint16_t* obase = dqkv + (size_t)b * T * 3 * W + h * 64;
  float q[64], g[64];
  load_vec64(base, fmt, q);
  load_vec64(do_cls + (size_t)b * W + h * 64, fmt, g);
  float lmax = -INFINITY;
  for (int kv = tid; kv < T; kv += CLS_THREADS) {
    const uint16_t* row = base + (size_t)kv * 3 * W;
    const float s = dot_row64(row + W, fmt, q) * 0.125f;
    sp[kv] = s;
    sds[kv] = dot_row64(row + 2 * W, fmt, g);
    lmax = fmaxf(lmax, s);
  }
  const float mx = block_max(lmax, red);
  float lsum = 0.f, ldd = 0.f;
  for (int kv = tid; kv < T; kv += CLS_THREADS) {
    const float e = __expf(sp[kv] - mx);
    sp[kv] = e;
    lsum += e;
    ldd = fmaf(e, sds[kv], ldd);
  }
  const float inv = 1.f / block_sum(lsum, red);
  const float D = block_sum(ldd, red) * inv;
  for (int kv = tid; kv < T; kv += CLS_THREADS) {
    const float p = sp[kv] * inv;
    const float ds = p * (sds[kv] - D) * 0.125f;
    sp[kv] = p;
    sds[kv] = ds;
    uint16_t* orow = obase + (size_t)kv * 3 * W;
    store_scaled64(orow + W, fmt, ds, q);        // dK[kv] = dS[kv] q_0
    store_scaled64(orow + 2 * W, fmt, p, g);     // dV[kv] = p[kv] dO_0
    if (kv > 0) {                                // dQ[kv > 0] = 0
#pragma unroll
      for (int c = 0; c < 8; ++c) reinterpret_cast<uint4*>(orow)[c] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  const int d = tid & 63, hf = tid >> 6;
  const int half = (T + 1) / 2;
  const int k0 = hf * half, k1 = min(T, k0 + half);
  const uint16_t* kp = base + W + d;
  float a0 = 0.f, a1 = 0.f;
  int kv = k0;
  for (; kv + 1 < k1; kv += 2) {
    a0 = fmaf(sds[kv], ld1f(kp + (size_t)kv * 3 * W, fmt), a0);
    a1 = fmaf(sds[kv + 1], ld1f(kp + (size_t)(kv + 1) * 3 * W, fmt), a1);
  }
  if (kv < k1) a0 = fmaf(sds[kv], ld1f(kp + (size_t)kv * 3 * W, fmt), a0);
  oh[hf][d] = a0 + a1;
  __syncthreads();
  if (tid < 32)                                  // dQ[0] = dS K
    reinterpret_cast<uint32_t*>(obase)[tid] =
        pack2(oh[0][2 * tid] + oh[1][2 * tid], oh[0][2 * tid + 1] + oh[1][2 * tid + 1], fmt);
}

}  // namespace

// o16[b*T, :] (class-token rows only; the other rows of o16 are left untouched)
int attn_cls_fwd(const void* qkv16, void* o16, int B, int T, int H, int fmt, cudaStream_t st,
                 const void* qkv16lo, void* o16lo) {
  if (B <= 0 || T <= 0 || T > CLS_TMAX || H <= 0) return RVLM_ERR_INVAL;
  if ((qkv16lo || o16lo) && fmt != 0) return RVLM_ERR_INVAL;
  attn_cls_fwd_kernel<<<B * H, CLS_THREADS, 0, st>>>(reinterpret_cast<const uint16_t*>(qkv16),
                                                    reinterpret_cast<const uint16_t*>(qkv16lo),
                                                    reinterpret_cast<uint16_t*>(o16),
                                                    reinterpret_cast<uint16_t*>(o16lo), T, H, fmt);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

// dqkv16 [B*T, 3W] (dense) from the class-token output gradient do_cls [B, W]
int attn_cls_bwd(const void* qkv16, const void* do_cls, void* dqkv16, int B, int T, int H, int fmt,
                 cudaStream_t st) {
  if (B <= 0 || T <= 0 || T > CLS_TMAX || H <= 0) return RVLM_ERR_INVAL;
  attn_cls_bwd_kernel<<<B * H, CLS_THREADS, 0, st>>>(reinterpret_cast<const uint16_t*>(qkv16),
                                                    reinterpret_cast<const uint16_t*>(do_cls),
                                                    reinterpret_cast<uint16_t*>(dqkv16), T, H, fmt);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

}  // namespace rvlm

// C ABI (tests): class-token attention forward / backward of one block
extern "C" int rvlm_attn_cls_fwd(const void* qkv16, void* o16, int B, int T, int H, int fmt,
                                 void* stream) {
  if (!qkv16 || !o16) return RVLM_ERR_INVAL;
  return rvlm::attn_cls_fwd(qkv16, o16, B, T, H, fmt, (cudaStream_t)stream, nullptr, nullptr);
}
extern "C" int rvlm_attn_cls_bwd(const void* qkv16, const void* do_cls16, void* dqkv16, int B, int T,
                                 int H, int fmt, void* stream) {
  if (!qkv16 || !do_cls16 || !dqkv16) return RVLM_ERR_INVAL;
  return rvlm::attn_cls_bwd(qkv16, do_cls16, dqkv16, B, T, H, fmt, (cudaStream_t)stream);
}
