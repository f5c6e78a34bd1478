// Reductions and the optimiser of the OUTER training step (the FARE / TeCoA model update,
// train/adversarial_training_clip.py:340-366): bias gradients (column sums of the output
// gradients), LayerNorm affine gradients, positional / class embedding gradients, the
// head's parameter gradients and a fused AdamW.  The weight gradients themselves are
// tcgen05 GEMMs with MN-major operands (vit_engine.cu, gemm.cuh).
//
// Column reductions over the M = B*T token rows are two-stage and deterministic:
// RSPLIT partial sums per column, then a fixed-order final sum.
#include "engine_internal.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

void count_launch(int n);

// ---------------------------------------------------------------------------------
// stage 1: partial[rs][col] over rows rs, rs + RSPLIT, ...   (block = 32 cols x 8 rows)
// ---------------------------------------------------------------------------------
// MODE 0: sum of a 16-bit matrix                     -> out0
// MODE 1: LayerNorm: sum dy * xhat -> out0 (dgamma), sum dy -> out1 (dbeta)
template <int MODE>
__global__ void __launch_bounds__(256)
col_reduce_kernel(const void* __restrict__ a, const float* __restrict__ x,
                  const float2* __restrict__ stats, int M, int N, int fmt,
                  float* __restrict__ part0, float* __restrict__ part1) {
  __shared__ float sh0[8][33], sh1[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const int rs = blockIdx.y, nrs = gridDim.y;
  float s0 = 0.f, s1 = 0.f;
  if (col < N) {
    for (int r = rs * 8 + ty; r < M; r += nrs * 8) {
      if (MODE == 0) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(a) + (size_t)r * N + col;
        float v;
        if (fmt) v = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
        else v = __half2float(*reinterpret_cast<const __half*>(p));
        s0 += v;
      } else {
        const float dy = reinterpret_cast<const float*>(a)[(size_t)r * N + col];
        const float2 st = stats[r];
        s0 += dy * (x[(size_t)r * N + col] - st.x) * st.y;
        s1 += dy;
      }
    }
  }
  sh0[ty][tx] = s0;
  sh1[ty][tx] = s1;
  __syncthreads();
  if (ty == 0 && col < N) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += sh0[i][tx]; t1 += sh1[i][tx]; }
    part0[(size_t)rs * N + col] = t0;
    if (MODE == 1) part1[(size_t)rs * N + col] = t1;
  }
}

__global__ void col_reduce_final_kernel(const float* __restrict__ part, int nrs, int N,
                                        float* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float s = 0.f;
  for (int i = 0; i < nrs; ++i) s += part[(size_t)i * N + col];
  out[col] = s;
}

// out[i] = sum_b in[b, i]      (positional-embedding gradient: sum over the batch)
__global__ void batch_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                 long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += in[(size_t)b * n + i];
    out[i] = s;
  }
}

// Head parameter gradients from the per-sample vectors the head kernel saved:
//   dproj[w, d] = sum_b pooled[b, w] * demb[b, d]
//   dgamma[w]   = sum_b dpooled[b, w] * xhat[b, w];   dbeta[w] = sum_b dpooled[b, w]
__global__ void head_param_grad_kernel(const float* __restrict__ pooled,
                                       const float* __restrict__ demb,
                                       const float* __restrict__ dpooled,
                                       const float* __restrict__ xhat, int B, int W, int D,
                                       float* __restrict__ dproj, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
  const int w = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += pooled[(size_t)b * W + w] * demb[(size_t)b * D + d];
    dproj[(size_t)w * D + d] = s;
  }
  if (threadIdx.x == 0) {
    float g = 0.f, bb = 0.f;
    for (int b = 0; b < B; ++b) {
      const float dp = dpooled[(size_t)b * W + w];
      g += dp * xhat[(size_t)b * W + w];
      bb += dp;
    }
    dgamma[w] = g;
    dbeta[w] = bb;
  }
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction), one fused pass.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                             float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                             float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

// torch.optim.SGD semantics (coupled weight decay, momentum without dampening / Nesterov).
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                           float* __restrict__ buf, long long n, float lr, float momentum,
                           float wd, int first) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i] + wd * pi;
    if (momentum != 0.f) {
      gi = first ? gi : momentum * buf[i] + gi;
      buf[i] = gi;
    }
    p[i] = pi - lr * gi;
  }
}

static inline int rc(int n = 1) {
  count_launch(n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

int launch_colsum16(const uint16_t* a, int M, int N, int fmt, float* scratch, float* out,
                    cudaStream_t st) {
  dim3 grid((N + 31) / 32, TRAIN_RSPLIT);
  col_reduce_kernel<0><<<grid, 256, 0, st>>>(a, nullptr, nullptr, M, N, fmt, scratch, nullptr);
  col_reduce_final_kernel<<<(N + 255) / 256, 256, 0, st>>>(scratch, TRAIN_RSPLIT, N, out);
  return rc(2);
}

int launch_ln_param_grads(const float* dy, const float* x, const float* stats, int M, int W,
                          float* scratch, float* dgamma, float* dbeta, cudaStream_t st) {
  dim3 grid((W + 31) / 32, TRAIN_RSPLIT);
  float* p1 = scratch + (size_t)TRAIN_RSPLIT * W;
  col_reduce_kernel<1><<<grid, 256, 0, st>>>(dy, x, reinterpret_cast<const float2*>(stats), M, W,
                                             0, scratch, p1);
  col_reduce_final_kernel<<<(W + 255) / 256, 256, 0, st>>>(scratch, TRAIN_RSPLIT, W, dgamma);
  col_reduce_final_kernel<<<(W + 255) / 256, 256, 0, st>>>(p1, TRAIN_RSPLIT, W, dbeta);
  return rc(3);
}

int launch_batch_sum(const float* in, float* out, int B, long long n, cudaStream_t st) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  batch_sum_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, B, n);
  return rc();
}

int launch_head_param_grads(const float* pooled, const float* demb, const float* dpooled,
                            const float* xhat, int B, int W, int D, float* dproj, float* dgamma,
                            float* dbeta, cudaStream_t st) {
  head_param_grad_kernel<<<W, 256, 0, st>>>(pooled, demb, dpooled, xhat, B, W, D, dproj, dgamma,
                                            dbeta);
  return rc();
}

}  // namespace rvlm

using namespace rvlm;

extern "C" int rvlm_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               long long n, double lr, double beta1, double beta2, double eps,
                               double weight_decay, int step, void* stream) {
  if (n <= 0 || step < 1) return RVLM_ERR_INVAL;
  // hyper-parameters arrive as the Python doubles torch.optim holds; the bias corrections are
  // formed in double as torch does (1 - beta ** step), then rounded once for the kernel
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, (float)lr, (float)beta1, (float)beta2, (float)eps,
      (float)weight_decay, (float)bc1, (float)sqrt(bc2));
  return rc();
}

extern "C" int rvlm_sgd_step(float* param, const float* grad, float* momentum_buf, long long n,
                             double lr, double momentum, double weight_decay, int step,
                             void* stream) {
  if (n <= 0 || step < 1 || (momentum != 0.0 && !momentum_buf)) return RVLM_ERR_INVAL;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sgd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      param, grad, momentum_buf, n, (float)lr, (float)momentum, (float)weight_decay, step == 1);
  return rc();
}
