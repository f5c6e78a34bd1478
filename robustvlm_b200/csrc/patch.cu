// Patch-layout plumbing around the patch-embedding GEMM (open_clip conv1 with
// kernel = stride = patch, no bias == GEMM over im2row'd patches):
// the attack state (x, delta, velocity, gradient) lives in patch-major layout
// [B*np, 3*p*p] for the whole inner loop, so the GEMM operand / epilogue accesses are
// row-contiguous; NCHW <-> patch-major conversion happens once per attack call.
#include "engine_internal.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

// one thread per (b, c, y, x4): float4 along x stays inside one patch row when p % 4 == 0
// (p = 32) or is handled element-wise (p = 14).
template <bool TO_PATCH, bool ADD>
__global__ void patch_layout_kernel(const float* __restrict__ src, const float* __restrict__ src2,
                                    float* __restrict__ dst, int B, int image, int p, int grid,
                                    int kp, const float* __restrict__ unit = nullptr) {
  const long long total = (long long)B * 3 * image * image;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % image);
    long long r = i / image;
    const int y = (int)(r % image);
    r /= image;
    const int c = (int)(r % 3);
    const int b = (int)(r / 3);
    const int py = y / p, ky = y - py * p, px = x / p, kx = x - px * p;
    const long long pm = ((long long)b * grid * grid + py * grid + px) * kp + c * p * p + ky * p + kx;
    if (TO_PATCH) {
      dst[pm] = src[i];
    } else {
      float v = src[pm];
      if (ADD) v = __fadd_rn(v, src2[pm]);
      if (unit) v = v / unit[b];   // undo the per-sample rescaling of a 16-bit-gradient backward
      dst[i] = v;
    }
  }
}

__global__ void patch_operand_kernel(const float* __restrict__ xp, const float* __restrict__ dp,
                                     uint16_t* __restrict__ a16, uint16_t* __restrict__ a16lo,
                                     int lda, long long rows, int kp,
                                     int pp, float m0, float m1, float m2, float s0, float s1,
                                     float s2, int fmt) {
  const int k4 = kp >> 2;
  const long long total = rows * k4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / k4;
    const int k = (int)(i - r * k4) * 4;
    float4 x = *reinterpret_cast<const float4*>(xp + r * kp + k);
    if (dp) {
      const float4 d = *reinterpret_cast<const float4*>(dp + r * kp + k);
      x.x = __fadd_rn(x.x, d.x); x.y = __fadd_rn(x.y, d.y);
      x.z = __fadd_rn(x.z, d.z); x.w = __fadd_rn(x.w, d.w);
    }
    const float* xv = &x.x;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (k + j) / pp;
      const float m = c == 0 ? m0 : (c == 1 ? m1 : m2);
      const float s = c == 0 ? s0 : (c == 1 ? s1 : s2);
      o[j] = (xv[j] - m) * s;
    }
    uint2 pk;
    pk.x = pack2(o[0], o[1], fmt);
    pk.y = pack2(o[2], o[3], fmt);
    *reinterpret_cast<uint2*>(a16 + r * lda + k) = pk;
    if (a16lo) {   // split-operand forward (fmt == fp16): lo = fp16(v - hi)
      const float2 h0 = unpack2(pk.x, 0), h1 = unpack2(pk.y, 0);
      uint2 lo;
      lo.x = pack2(o[0] - h0.x, o[1] - h0.y, 0);
      lo.y = pack2(o[2] - h1.x, o[3] - h1.y, 0);
      *reinterpret_cast<uint2*>(a16lo + r * lda + k) = lo;
    }
  }
}

__global__ void cls_rows_kernel(float* __restrict__ tok, const float* __restrict__ cls,
                                const float* __restrict__ pos, int B, int T, int W) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * W; i += gridDim.x * blockDim.x) {
    const int b = i / W, w = i - b * W;
    tok[(size_t)b * T * W + w] = cls[w] + pos[w];
  }
}

static inline unsigned nblk(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}
void count_launch(int n);
static inline int rc() {
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

int launch_patchify(const float* img, float* pm, int B, const PatchGeom& g, cudaStream_t st) {
  const long long n = (long long)B * 3 * g.image * g.image;
  patch_layout_kernel<true, false><<<nblk(n, 256), 256, 0, st>>>(img, nullptr, pm, B, g.image,
                                                                 g.patch, g.grid, g.kp);
  return rc();
}
int launch_unpatchify(const float* pm, float* img, int B, const PatchGeom& g, cudaStream_t st,
                      const float* unit) {
  const long long n = (long long)B * 3 * g.image * g.image;
  patch_layout_kernel<false, false><<<nblk(n, 256), 256, 0, st>>>(pm, nullptr, img, B, g.image,
                                                                  g.patch, g.grid, g.kp, unit);
  return rc();
}
int launch_unpatchify_add(const float* xp, const float* dp, float* img, int B, const PatchGeom& g,
                          cudaStream_t st) {
  const long long n = (long long)B * 3 * g.image * g.image;
  patch_layout_kernel<false, true><<<nblk(n, 256), 256, 0, st>>>(xp, dp, img, B, g.image, g.patch,
                                                                 g.grid, g.kp);
  return rc();
}
int launch_patch_operand(const float* xp, const float* dp, uint16_t* a16, int lda, long long rows,
                         const PatchGeom& g, int fmt, cudaStream_t st, uint16_t* a16lo) {
  const long long n = rows * (g.kp >> 2);
  patch_operand_kernel<<<nblk(n, 256), 256, 0, st>>>(xp, dp, a16, a16lo, lda, rows, g.kp,
                                                     g.patch * g.patch, g.mean[0], g.mean[1],
                                                     g.mean[2], g.inv_std[0], g.inv_std[1],
                                                     g.inv_std[2], fmt);
  return rc();
}
int launch_cls_rows(float* tok, const float* cls, const float* pos, int B, int T, int W,
                    cudaStream_t st) {
  cls_rows_kernel<<<nblk((long long)B * W, 256), 256, 0, st>>>(tok, cls, pos, B, T, W);
  return rc();
}

// ---------------------------------------------------------------------------
// weight conversion (fp32 open_clip layout -> 16-bit GEMM operands)
// ---------------------------------------------------------------------------
// fmt: 0 = fp16, 1 = bf16, 2 = the fp16 LO half of a split operand: fp16(v - float(fp16(v)))
__device__ __forceinline__ uint16_t cvt16(float v, int fmt) {
  if (fmt == 1) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __half h = __float2half_rn(v);
  if (fmt == 2) h = __float2half_rn(v - __half2float(h));
  return *reinterpret_cast<uint16_t*>(&h);
}

__global__ void convert16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                 int R, int C, int ldd, int fmt) {
  const long long total = (long long)R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    dst[r * ldd + c] = cvt16(src[i], fmt);
  }
}

__global__ void transpose_convert16_kernel(const float* __restrict__ src,
                                           uint16_t* __restrict__ dst, int R, int C, int ldd,
                                           int fmt) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = by + j, c = bx + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = bx + j, r = by + threadIdx.x;  // dst row = c, dst col = r
    if (c < C && r < R) {
      dst[(size_t)c * ldd + r] = cvt16(tile[threadIdx.x][j], fmt);
    }
  }
}

// 16-bit -> 16-bit format change (fp16 <-> bf16) of a contiguous buffer, 8 elements per thread:
// the outer training step keeps a gradient-format copy of every forward GEMM operand for the
// weight-gradient GEMMs when its forward runs in another format (fp16 forward, bf16 gradients)
__global__ void recast16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                long long n8, int fmt_src, int fmt_dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = src[i];
    const uint32_t* a = &u.x;
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack2(a[j], fmt_src);
      o[j] = pack2(f.x, f.y, fmt_dst);
    }
    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
int launch_recast16(const uint16_t* src, uint16_t* dst, long long n, int fmt_src, int fmt_dst,
                    cudaStream_t st) {
  if (n & 7) return RVLM_ERR_INVAL;
  recast16_kernel<<<nblk(n >> 3, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(src),
                                                     reinterpret_cast<uint4*>(dst), n >> 3,
                                                     fmt_src, fmt_dst);
  return rc();
}

int launch_convert16(const float* src, uint16_t* dst, long long n, int fmt, cudaStream_t st) {
  convert16_kernel<<<nblk(n, 256), 256, 0, st>>>(src, dst, 1, (int)n, (int)n, fmt);
  return rc();
}
int launch_convert16_pitched(const float* src, uint16_t* dst, int R, int C, int ldd, int fmt,
                             cudaStream_t st) {
  convert16_kernel<<<nblk((long long)R * C, 256), 256, 0, st>>>(src, dst, R, C, ldd, fmt);
  return rc();
}
int launch_transpose_convert16(const float* src, uint16_t* dst, int R, int C, int ldd, int fmt,
                               cudaStream_t st) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose_convert16_kernel<<<grid, block, 0, st>>>(src, dst, R, C, ldd, fmt);
  return rc();
}

}  // namespace rvlm
