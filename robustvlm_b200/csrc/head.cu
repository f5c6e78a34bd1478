// Embedding head + attack loss, forward and backward in one launch.
//
//   pooled = ln_post(x_L[:, 0]);  emb = pooled @ proj;  optional F.normalize(emb)
//       (open_clip VisionTransformer tail + ClipVisionModel.forward,
//        train/adversarial_training_clip.py:253-257)
//   FARE   : loss_b = sum_d (emb - emb_orig)^2                     (:509-521)
//   TeCoA  : loss_b = CE(emb @ (logit_scale * T), y)               (:500-504,523-528)
//            (also the zero-shot classifier of the evaluation APGD,
//             CLIP_eval/clip_robustbench.py:50-69, with normalisation on)
//   pred_b = [argmax(out_b) == y_b], out = class logits or the embedding itself
//            (apgd_train's quirk, train/apgd_train.py:191,301)
// and the gradient of loss_b w.r.t. the CLS row of the last residual stream, written
// into row b*T of the [B*T, W] gradient buffers (all other rows are zero: only the
// CLS token reaches the loss).
//
// Tiny work (B rows): one CTA of 1024 threads per sample, fp32 SIMT.  The four matrix-vector
// products (proj, text and their transposes; 3 MB each, streamed from L2 by every CTA) are
// latency-bound, so they are organised for memory-level parallelism: the forward products split
// the reduction dimension over four 256-thread slices with up to four independent accumulators
// per thread (partial sums merged through shared memory in a fixed order), the transposed ones
// run one warp per output with four independent partial sums per lane.
#include "gemm.cuh"
#include "engine_internal.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);
constexpr int HEAD_THREADS = 1024;
constexpr int HEAD_SLICES = HEAD_THREADS / 256;   // reduction slices of the forward products

__device__ __forceinline__ float blk_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float blk_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffff, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, red[i]);
  return t;
}
// first index attaining the maximum (torch.max semantics)
__device__ __forceinline__ int blk_argmax(const float* v, int n, float* red, int* redi) {
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[i];
    if (x > best || (x == best && i < bi)) { best = x; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffff, best, o);
    const int oi = __shfl_xor_sync(0xffffffff, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = best; redi[threadIdx.x >> 5] = bi; }
  __syncthreads();
  float fb = -INFINITY;
  int fi = 0x7fffffff;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i)
    if (red[i] > fb || (red[i] == fb && redi[i] < fi)) { fb = red[i]; fi = redi[i]; }
  return fi;
}

// out[j] = sum_i v[i] * Mx[i * ldm + j], j < n_out, i < n_in  (Mx row-major [n_in, n_out]).
// Thread t: columns (t & 255) + 256 k, rows of slice t >> 8.  `part` = [HEAD_SLICES][n_out] scratch.
__device__ __forceinline__ void matvec_fwd(const float* __restrict__ v, const float* __restrict__ Mx,
                                           int n_in, int n_out, float scale, float* part,
                                           float* out) {
  const int col0 = threadIdx.x & 255, slice = threadIdx.x >> 8;
  const int per = (n_in + HEAD_SLICES - 1) / HEAD_SLICES;
  const int i0 = slice * per, i1 = min(n_in, i0 + per);
  for (int cb = 0; cb < n_out; cb += 1024) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ok[k] = cb + col0 + 256 * k < n_out;
#pragma unroll 4
    for (int i = i0; i < i1; ++i) {
      const float vi = v[i];
      const float* row = Mx + (size_t)i * n_out + cb + col0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) acc[k] = fmaf(vi, __ldg(row + 256 * k), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) part[slice * n_out + cb + col0 + 256 * k] = acc[k];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_out; j += HEAD_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int sl = 0; sl < HEAD_SLICES; ++sl) t += part[sl * n_out + j];
    out[j] = t * scale;
  }
  __syncthreads();
}

// out[i] = scale * sum_j Mx[i * n_out + j] * v[j]: one warp per row i, four partial sums per lane.
template <typename F>
__device__ __forceinline__ void matvec_bwd(const float* __restrict__ v, const float* __restrict__ Mx,
                                           int n_in, int n_out, F&& emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < n_in; i += HEAD_THREADS / 32) {
    const float* row = Mx + (size_t)i * n_out;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = lane;
    for (; j + 96 < n_out; j += 128) {
      a0 = fmaf(__ldg(row + j), v[j], a0);
      a1 = fmaf(__ldg(row + j + 32), v[j + 32], a1);
      a2 = fmaf(__ldg(row + j + 64), v[j + 64], a2);
      a3 = fmaf(__ldg(row + j + 96), v[j + 96], a3);
    }
    for (; j < n_out; j += 32) a0 = fmaf(__ldg(row + j), v[j], a0);
    float acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, o);
    if (lane == 0) emit(i, acc);
  }
}

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(HeadArgs A) {
  extern __shared__ float sm[];
  const int W = A.W, D = A.D, C = A.C;
  float* xh = sm;               // [W] normalised row
  float* pooled = xh + W;       // [W]
  float* e = pooled + W;        // [D] embedding (normalised if requested)
  float* de = e + D;            // [D]
  float* lg = de + D;           // [C] logits / dlogits
  float* part = lg + (C > 0 ? C : 1);   // [HEAD_SLICES][max(D, C)] partial sums
  __shared__ float red[32];
  __shared__ int redi[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xr = A.xL + (size_t)b * A.T * W;

  // ---- ln_post
  float s = 0.f;
  for (int i = tid; i < W; i += HEAD_THREADS) { xh[i] = xr[i]; s += xh[i]; }
  const float mean = blk_sum(s, red) / W;
  float q = 0.f;
  for (int i = tid; i < W; i += HEAD_THREADS) { const float d = xh[i] - mean; q += d * d; }
  const float rstd = 1.f / sqrtf(blk_sum(q, red) / W + 1e-5f);
  for (int i = tid; i < W; i += HEAD_THREADS) {
    xh[i] = (xh[i] - mean) * rstd;
    pooled[i] = xh[i] * A.ln_g[i] + A.ln_b[i];
  }
  __syncthreads();
  // ---- projection
  matvec_fwd(pooled, A.proj, W, D, 1.f, part, e);
  float nrm = 1.f;
  if (A.normalize) {
    float ss = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) ss += e[d] * e[d];
    nrm = fmaxf(sqrtf(blk_sum(ss, red)), 1e-12f);
    for (int d = tid; d < D; d += HEAD_THREADS) e[d] = e[d] / nrm;
    __syncthreads();
  }
  if (A.emb_out)
    for (int d = tid; d < D; d += HEAD_THREADS) A.emb_out[(size_t)b * D + d] = e[d];

  const long long yb = A.y ? A.y[b] : -1;
  float loss = 0.f;
  int pred = 0;
  if (A.mode == 1) {
    float ss = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) {
      const float df = e[d] - A.emb_orig[(size_t)b * D + d];
      ss += df * df;
      de[d] = 2.f * df;
    }
    loss = blk_sum(ss, red);
  } else if (A.mode == 3) {
    // upstream gradient w.r.t. the (normalised) embedding supplied by the caller
    for (int d = tid; d < D; d += HEAD_THREADS) de[d] = A.emb_orig[(size_t)b * D + d];
  } else if (A.mode == 2) {
    matvec_fwd(e, A.text, D, C, A.logit_scale, part, lg);
    if (A.logits_out)
      for (int c = tid; c < C; c += HEAD_THREADS) A.logits_out[(size_t)b * C + c] = lg[c];
    if (A.pred_on_logits) pred = (blk_argmax(lg, C, red, redi) == (int)yb) ? 1 : 0;
    float mx = -INFINITY;
    for (int c = tid; c < C; c += HEAD_THREADS) mx = fmaxf(mx, lg[c]);
    mx = blk_max(mx, red);
    float se = 0.f;
    for (int c = tid; c < C; c += HEAD_THREADS) se += expf(lg[c] - mx);
    se = blk_sum(se, red);
    const float lz = mx + logf(se);
    loss = (yb >= 0 && yb < C) ? lz - lg[yb] : 0.f;
    __syncthreads();
    for (int c = tid; c < C; c += HEAD_THREADS)
      lg[c] = expf(lg[c] - lz) - ((long long)c == yb ? 1.f : 0.f);
    __syncthreads();
    // d e = logit_scale * text @ dlogits
    matvec_bwd(lg, A.text, D, C, [&](int d, float acc) { de[d] = acc * A.logit_scale; });
  } else if (A.mode == 4 || A.mode == 5) {
    // DLR (autoattack/autopgd_base.py:195-201) and targeted DLR (:616-621) on the class logits:
    //   dlr   = -(z_y - (argmax == y ? z_(2) : z_(1))) / (z_(1) - z_(3) + 1e-12)
    //   dlr-t = -(z_y - z_t) / (z_(1) - (z_(3) + z_(4)) / 2 + 1e-12)
    // z_(k) = k-th largest logit.  The gradient w.r.t. the logits has at most five non-zero
    // entries, so d e = logit_scale * sum_j w_j text[:, idx_j].
    matvec_fwd(e, A.text, D, C, A.logit_scale, part, lg);
    if (A.logits_out)
      for (int c = tid; c < C; c += HEAD_THREADS) A.logits_out[(size_t)b * C + c] = lg[c];
    __syncthreads();
    int ti[4];
    float tv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {     // top-4 by repeated arg-max (first index on ties)
      ti[k] = blk_argmax(lg, C, red, redi);
      tv[k] = lg[ti[k]];
      __syncthreads();
      if (tid == 0) lg[ti[k]] = -INFINITY;
      __syncthreads();
    }
    if (tid == 0)
      for (int k = 0; k < 4; ++k) lg[ti[k]] = tv[k];
    __syncthreads();
    pred = (ti[0] == (int)yb) ? 1 : 0;
    const int yi = (yb >= 0 && yb < C) ? (int)yb : 0;
    int idx[5];
    float wv[5];
    if (A.mode == 4) {
      const bool top = ti[0] == yi;
      const int other = top ? ti[1] : ti[0];
      const float Dn = tv[0] - tv[2] + 1e-12f;
      const float N = lg[yi] - lg[other];
      loss = -N / Dn;
      idx[0] = yi;    wv[0] = -1.f / Dn;
      idx[1] = other; wv[1] = 1.f / Dn;
      idx[2] = ti[0]; wv[2] = N / (Dn * Dn);
      idx[3] = ti[2]; wv[3] = -N / (Dn * Dn);
      idx[4] = 0;     wv[4] = 0.f;
    } else {
      const long long tb = A.y_target ? A.y_target[b] : 0;
      const int ti_ = (tb >= 0 && tb < C) ? (int)tb : 0;
      const float Dn = tv[0] - 0.5f * (tv[2] + tv[3]) + 1e-12f;
      const float N = lg[yi] - lg[ti_];
      loss = -N / Dn;
      idx[0] = yi;    wv[0] = -1.f / Dn;
      idx[1] = ti_;   wv[1] = 1.f / Dn;
      idx[2] = ti[0]; wv[2] = N / (Dn * Dn);
      idx[3] = ti[2]; wv[3] = -0.5f * N / (Dn * Dn);
      idx[4] = ti[3]; wv[4] = -0.5f * N / (Dn * Dn);
    }
    for (int d = tid; d < D; d += HEAD_THREADS) {
      const float* trow = A.text + (size_t)d * C;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 5; ++j) acc = fmaf(wv[j], __ldg(trow + idx[j]), acc);
      de[d] = acc * A.logit_scale;
    }
  }
  __syncthreads();
  if ((A.mode == 1 || A.mode == 2) && !(A.mode == 2 && A.pred_on_logits))
    pred = (blk_argmax(e, D, red, redi) == (int)yb) ? 1 : 0;   // (DLR modes: always the logits)
  if (tid == 0) {
    if (A.loss) A.loss[b] = loss;
    if (A.pred) A.pred[b] = pred;
  }
  if (A.mode == 0 || !A.dx32) return;

  // ---- backward: normalise
  if (A.normalize) {
    float dot = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) dot += e[d] * de[d];
    dot = blk_sum(dot, red);
    for (int d = tid; d < D; d += HEAD_THREADS) de[d] = (de[d] - e[d] * dot) / nrm;
  }
  __syncthreads();
  if (A.t_demb) {  // outer training step: keep what the head's parameter gradients need
    for (int d = tid; d < D; d += HEAD_THREADS)
      A.t_demb[(size_t)b * D + d] = de[d] * A.grad_scale;
    for (int i = tid; i < W; i += HEAD_THREADS) {
      A.t_xhat[(size_t)b * W + i] = xh[i];
      A.t_pooled[(size_t)b * W + i] = pooled[i];
    }
    __syncthreads();
  }
  // ---- d pooled = proj @ d emb  (reuse `pooled` as g = dpooled * gamma)
  matvec_bwd(de, A.proj, W, D, [&](int w, float acc) {
    if (A.t_dpooled) A.t_dpooled[(size_t)b * W + w] = acc * A.grad_scale;
    pooled[w] = acc * A.ln_g[w] * A.grad_scale;
  });
  __syncthreads();
  // ---- ln_post backward
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid; i < W; i += HEAD_THREADS) { s1 += pooled[i]; s2 += pooled[i] * xh[i]; }
  const float c1 = blk_sum(s1, red) / W;
  const float c2 = blk_sum(s2, red) / W;
  float mxabs = 0.f;
  for (int i = tid; i < W; i += HEAD_THREADS) {
    const float v = rstd * (pooled[i] - c1 - xh[i] * c2);
    pooled[i] = v;
    mxabs = fmaxf(mxabs, fabsf(v));
  }
  float unit = 1.f;
  if (A.grad_unit) {
    mxabs = blk_max(mxabs, red);
    // max-abs of the sample's gradient -> 64: fp16 keeps its full 11 bits down to 1e-6 of that
    // (normal range), and the gradient may still grow 1000x on its way down before the packs saturate
    unit = (mxabs > 0.f && mxabs < INFINITY) ? 64.f / mxabs : 1.f;
  }
  if (A.unit_out && tid == 0) A.unit_out[b] = unit;
  __syncthreads();
  float* d32 = A.dx32 + (size_t)b * A.T * W;
  uint16_t* d16 = A.dx16 ? A.dx16 + (size_t)b * A.T * W : nullptr;
  for (int i = tid; i < W; i += HEAD_THREADS) {
    const float v = pooled[i] * unit;
    d32[i] = v;
    if (d16) {
      if (A.fmt_g) { __nv_bfloat16 h = __float2bfloat16_rn(v); d16[i] = *reinterpret_cast<uint16_t*>(&h); }
      else { __half h = __float2half_rn(v); d16[i] = *reinterpret_cast<uint16_t*>(&h); }
    }
  }
}

int launch_head(const HeadArgs& A, int B, cudaStream_t st) {
  const int mx = A.D > A.C ? A.D : A.C;
  const size_t smem = (size_t)(2 * A.W + 2 * A.D + (A.C > 0 ? A.C : 1) + HEAD_SLICES * mx) *
                      sizeof(float);
  static DeviceMax set;
  if (smem > 48 * 1024 && set.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return -(int)e - 1000;
  }
  head_kernel<<<B, HEAD_THREADS, smem, st>>>(A);
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

}  // namespace rvlm
