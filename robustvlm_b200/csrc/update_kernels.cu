// Perturbation-update kernels: the arithmetic between two model evaluations in
// PGD / APGD.  HBM-bound, per-sample, layout-agnostic: every tensor is a
// contiguous [B, n] fp32 array (n = 3*H*W per image, in whatever element order the
// caller uses consistently), so the same kernels serve the NCHW tensors of the
// generic autograd path and the patch-major buffers of the fused ViT engine.
//
// Bit-exactness: the ell_inf paths reproduce the reference's fp32 rounding
// sequence exactly (every mul/add is an explicit round-to-nearest intrinsic so
// nvcc cannot contract them into FMAs; PyTorch eager never fuses them).
//   pgd ell_inf      train/pgd_train.py:37-59 + vlm_eval/attacks/utils.py:8-26
//   pgd ell_2        same, F.normalize / torch.renorm semantics
//   apgd step        train/apgd_train.py:205-254 == autoattack/autopgd_base.py:328-350
//   apgd bookkeeping train/apgd_train.py:301-355 == autoattack/autopgd_base.py:388-446
// Per-sample ell_2 norms use a fixed two-level reduction (RED_BLOCKS partial sums
// per sample, summed in index order) so results are run-to-run deterministic.
#include "ptx.cuh"
#include "../../include/rvlm.h"

namespace rvlm {

constexpr int UPD_THREADS = 256;
constexpr int RED_BLOCKS = 32;  // partial sums per sample

__device__ __forceinline__ float fsign(float v) {
  return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);  // sign(0) = 0, sign(nan) handled by caller
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffff, t, o);
  }
  return t;  // valid in warp 0
}

// Sum of the RED_BLOCKS partial sums of sample b: one value per lane, xor-shuffle tree (a fixed
// order, so every pass of every launch forms bitwise the same norm).  Must be called by full
// warps before any divergence.
__device__ __forceinline__ float sample_sum(const float* partial, int b) {
  static_assert(RED_BLOCKS == 32, "one partial sum per lane");
  float v = partial[b * RED_BLOCKS + (threadIdx.x & 31)];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  return v;
}

// Element loops of the multi-pass kernels: float4 accesses when n % 4 == 0 (every image shape),
// scalar otherwise.  `body(idx, k)` sees k = 4 or 1 consecutive elements starting at base + idx.
template <int V> struct Pack { float v[V]; };
template <int V> __device__ __forceinline__ Pack<V> ldp(const float* p, long long idx) {
  Pack<V> r;
  if constexpr (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + idx);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = p[idx];
  }
  return r;
}
template <int V> __device__ __forceinline__ void stp(float* p, long long idx, const Pack<V>& r) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p + idx) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
    p[idx] = r.v[0];
  }
}

// ---------------------------------------------------------------------------
// PGD ell_inf: one fused pass.  grid = (blocks, B)
// ---------------------------------------------------------------------------
// `mask` (nullable) multiplies the gradient first (vlm_eval/attacks/pgd.py:39); image_space = 0
// skips the [0,1] clamp (vlm_eval/attacks/pgd.py:57).
__global__ void pgd_update_linf_kernel(float* __restrict__ delta, float* __restrict__ vel,
                                       const float* __restrict__ grad, const float* __restrict__ x,
                                       const float* __restrict__ mask, int image_space,
                                       long long n, float stepsize, float eps, float dir,
                                       float momentum) {
  const long long base = (long long)blockIdx.y * n;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 g = reinterpret_cast<const float4*>(grad + base)[i];
    float4 d = reinterpret_cast<const float4*>(delta + base)[i];
    float4 v = reinterpret_cast<const float4*>(vel + base)[i];
    float4 xx = reinterpret_cast<const float4*>(x + base)[i];
    float4 mk = mask ? reinterpret_cast<const float4*>(mask + base)[i]
                     : make_float4(1.f, 1.f, 1.f, 1.f);
    float* gp = &g.x; float* dp = &d.x; float* vp = &v.x; float* xp = &xx.x; float* mp = &mk.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = mask ? __fmul_rn(mp[j], gp[j]) : gp[j];
      if (gj != gj) gj = 0.f;                                  // NaN scrub (:40-42)
      const float gs = fsign(gj);                              // normalize_grad (:44)
      const float m = __fadd_rn(__fmul_rn(momentum, vp[j]), gs);  // :46
      const float vn = fsign(m);                               // :47
      float dn = __fadd_rn(dp[j], __fmul_rn(__fmul_rn(stepsize, vn), dir));  // :49-52
      dn = fminf(fmaxf(dn, -eps), eps);                        // project (:56)
      if (image_space) {
        const float xa = fminf(fmaxf(__fadd_rn(xp[j], dn), 0.f), 1.f);
        dn = __fsub_rn(xa, xp[j]);                             // :57-59
      }
      dp[j] = dn;
      vp[j] = vn;
    }
    reinterpret_cast<float4*>(delta + base)[i] = d;
    reinterpret_cast<float4*>(vel + base)[i] = v;
  }
}

// ---------------------------------------------------------------------------
// PGD ell_2: four passes, three deterministic per-sample reductions.
// ---------------------------------------------------------------------------
// pass 1: partial[b][blk] = sum g^2 (NaN -> 0)
template <int V>
__global__ void l2_sumsq_kernel(const float* __restrict__ a, const float* __restrict__ bsub,
                                long long n, float* __restrict__ partial, int scrub_nan,
                                const float* __restrict__ mask) {
  __shared__ float sh[32];
  const long long base = (long long)blockIdx.y * n;
  float s = 0.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < n;
       i += (long long)gridDim.x * blockDim.x * V) {
    Pack<V> va = ldp<V>(a, base + i), vb, vm;
    if (bsub) vb = ldp<V>(bsub, base + i);
    if (mask) vm = ldp<V>(mask, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float v = va.v[j];
      if (bsub) v = __fsub_rn(v, vb.v[j]);
      if (mask) v = __fmul_rn(vm.v[j], v);
      if (scrub_nan && v != v) v = 0.f;
      s = __fadd_rn(s, __fmul_rn(v, v));
    }
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.y * RED_BLOCKS + blockIdx.x] = t;
}

// pass 2: m = momentum*v + g / max(||g||, 1e-12)  -> vel (unnormalised); partial = sum m^2
template <int V>
__global__ void pgd_l2_momentum_kernel(float* __restrict__ vel, const float* __restrict__ grad,
                                       const float* __restrict__ mask, long long n, float momentum,
                                       const float* __restrict__ gpart,
                                       float* __restrict__ mpart) {
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const long long base = (long long)b * n;
  const float gn = fmaxf(sqrtf(sample_sum(gpart, b)), 1e-12f);
  float s = 0.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < n;
       i += (long long)gridDim.x * blockDim.x * V) {
    Pack<V> g = ldp<V>(grad, base + i), v = ldp<V>(vel, base + i), vm;
    if (mask) vm = ldp<V>(mask, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float gj = g.v[j];
      if (mask) gj = __fmul_rn(vm.v[j], gj);
      if (gj != gj) gj = 0.f;
      const float m = __fadd_rn(__fmul_rn(momentum, v.v[j]), __fdiv_rn(gj, gn));
      v.v[j] = m;
      s = __fadd_rn(s, __fmul_rn(m, m));
    }
    stp<V>(vel, base + i, v);
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) mpart[b * RED_BLOCKS + blockIdx.x] = t;
}

// pass 3: v = m / max(||m||,1e-12); delta += dir*step*v; partial = sum delta^2
template <int V>
__global__ void pgd_l2_step_kernel(float* __restrict__ delta, float* __restrict__ vel, long long n,
                                   float stepsize, float dir, const float* __restrict__ mpart,
                                   float* __restrict__ dpart) {
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const long long base = (long long)b * n;
  const float mn = fmaxf(sqrtf(sample_sum(mpart, b)), 1e-12f);
  float s = 0.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < n;
       i += (long long)gridDim.x * blockDim.x * V) {
    Pack<V> v = ldp<V>(vel, base + i), d = ldp<V>(delta, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float vj = __fdiv_rn(v.v[j], mn);
      v.v[j] = vj;
      const float dj = __fadd_rn(d.v[j], __fmul_rn(__fmul_rn(stepsize, vj), dir));
      d.v[j] = dj;
      s = __fadd_rn(s, __fmul_rn(dj, dj));
    }
    stp<V>(vel, base + i, v);
    stp<V>(delta, base + i, d);
  }
  const float t = block_sum(s, sh);
  if (threadIdx.x == 0) dpart[b * RED_BLOCKS + blockIdx.x] = t;
}

// pass 4: renorm (torch.renorm: rows with norm > eps scaled by eps/(norm+1e-7)), image clamp
template <int V>
__global__ void pgd_l2_project_kernel(float* __restrict__ delta, const float* __restrict__ x,
                                      int image_space, long long n, float eps,
                                      const float* __restrict__ dpart) {
  const int b = blockIdx.y;
  const long long base = (long long)b * n;
  const float dn = sqrtf(sample_sum(dpart, b));
  const float scale = dn > eps ? __fdiv_rn(eps, __fadd_rn(dn, 1e-7f)) : 1.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < n;
       i += (long long)gridDim.x * blockDim.x * V) {
    Pack<V> d = ldp<V>(delta, base + i), xv;
    if (image_space) xv = ldp<V>(x, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float dj = __fmul_rn(d.v[j], scale);
      if (image_space) {
        const float xa = fminf(fmaxf(__fadd_rn(xv.v[j], dj), 0.f), 1.f);
        dj = __fsub_rn(xa, xv.v[j]);
      }
      d.v[j] = dj;
    }
    stp<V>(delta, base + i, d);
  }
}

// ---------------------------------------------------------------------------
// APGD step, ell_inf: fused single pass.
// ---------------------------------------------------------------------------
template <int V>
__global__ void apgd_step_linf_kernel(float* __restrict__ x_adv, float* __restrict__ x_adv_old,
                                      const float* __restrict__ x, const float* __restrict__ grad,
                                      const float* __restrict__ step_size, long long n, float eps,
                                      float a) {
  const int b = blockIdx.y;
  const long long base = (long long)b * n;
  const float s = step_size[b];
  const float one_minus_a = (float)(1.0 - (double)a);
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < n;
       i += (long long)gridDim.x * blockDim.x * V) {
    Pack<V> pa = ldp<V>(x_adv, base + i), po = ldp<V>(x_adv_old, base + i);
    const Pack<V> pc = ldp<V>(x, base + i), pg = ldp<V>(grad, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xa = pa.v[j], xo = po.v[j], xc = pc.v[j], g = pg.v[j];
      const float grad2 = __fsub_rn(xa, xo);
      const float lo = __fsub_rn(xc, eps), hi = __fadd_rn(xc, eps);
      float sg = fsign(g);
      if (g != g) sg = g;  // torch.sign(nan) = nan (the reference does not scrub in APGD)
      float z = __fadd_rn(xa, __fmul_rn(s, sg));
      z = fminf(fmaxf(fminf(fmaxf(z, lo), hi), 0.f), 1.f);
      float z2 = __fadd_rn(__fadd_rn(xa, __fmul_rn(__fsub_rn(z, xa), a)),
                           __fmul_rn(grad2, one_minus_a));
      z2 = fminf(fmaxf(fminf(fmaxf(z2, lo), hi), 0.f), 1.f);
      po.v[j] = xa;
      pa.v[j] = z2;
    }
    stp<V>(x_adv_old, base + i, po);
    stp<V>(x_adv, base + i, pa);
  }
}

// ---------------------------------------------------------------------------
// APGD step, ell_2: 4 passes.
//   u  = x_adv + s * g/(||g||+1e-12)
//   z  = clamp(x + (u-x)/(||u-x||+1e-12) * min(eps, ||u-x||), 0, 1)
//   w  = x_adv + (z-x_adv)*a + grad2*(1-a)
//   x' = clamp(x + (w-x)/(||w-x||+1e-12) * min(eps, ||w-x||), 0, 1)
// Intermediates are recomputed from the same inputs in each pass (bitwise the same
// values), so only x_adv / x_adv_old are written, in the last pass.
// ---------------------------------------------------------------------------
struct ApgdL2Args {
  float* x_adv;
  float* x_adv_old;
  const float* x;
  const float* grad;
  const float* step_size;
  long long n;
  float eps, a;
  const float* p_g;   // partial sums of g^2
  float* p_u;         // partial sums of (u-x)^2
  float* p_w;         // partial sums of (w-x)^2
};

__device__ __forceinline__ float apgd_l2_u(float xa, float g, float s, float gden) {
  return __fadd_rn(xa, __fdiv_rn(__fmul_rn(s, g), gden));
}
__device__ __forceinline__ float apgd_l2_proj(float xc, float u, float uden, float umin) {
  const float t = __fmul_rn(__fdiv_rn(__fsub_rn(u, xc), uden), umin);
  return fminf(fmaxf(__fadd_rn(xc, t), 0.f), 1.f);
}
__device__ __forceinline__ float apgd_l2_w(float xa, float xo, float xc, float g, float s,
                                           float gden, float uden, float umin, float a,
                                           float one_minus_a) {
  const float u = apgd_l2_u(xa, g, s, gden);
  const float z = apgd_l2_proj(xc, u, uden, umin);
  const float grad2 = __fsub_rn(xa, xo);
  return __fadd_rn(__fadd_rn(xa, __fmul_rn(__fsub_rn(z, xa), a)), __fmul_rn(grad2, one_minus_a));
}

template <int V>
__global__ void apgd_l2_pass_u(ApgdL2Args A) {
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const long long base = (long long)b * A.n;
  const float gden = __fadd_rn(sqrtf(sample_sum(A.p_g, b)), 1e-12f);
  const float s = A.step_size[b];
  float acc = 0.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < A.n;
       i += (long long)gridDim.x * blockDim.x * V) {
    const Pack<V> pa = ldp<V>(A.x_adv, base + i), pg = ldp<V>(A.grad, base + i),
                  pc = ldp<V>(A.x, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float d = __fsub_rn(apgd_l2_u(pa.v[j], pg.v[j], s, gden), pc.v[j]);
      acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) A.p_u[b * RED_BLOCKS + blockIdx.x] = t;
}

template <int V>
__global__ void apgd_l2_pass_w(ApgdL2Args A) {
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const long long base = (long long)b * A.n;
  const float gden = __fadd_rn(sqrtf(sample_sum(A.p_g, b)), 1e-12f);
  const float un = sqrtf(sample_sum(A.p_u, b));
  const float uden = __fadd_rn(un, 1e-12f), umin = fminf(A.eps, un);
  const float s = A.step_size[b];
  const float oma = (float)(1.0 - (double)A.a);
  float acc = 0.f;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < A.n;
       i += (long long)gridDim.x * blockDim.x * V) {
    const Pack<V> pa = ldp<V>(A.x_adv, base + i), po = ldp<V>(A.x_adv_old, base + i),
                  pg = ldp<V>(A.grad, base + i), pc = ldp<V>(A.x, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float w = apgd_l2_w(pa.v[j], po.v[j], pc.v[j], pg.v[j], s, gden, uden, umin, A.a, oma);
      const float d = __fsub_rn(w, pc.v[j]);
      acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) A.p_w[b * RED_BLOCKS + blockIdx.x] = t;
}

template <int V>
__global__ void apgd_l2_pass_final(ApgdL2Args A) {
  const int b = blockIdx.y;
  const long long base = (long long)b * A.n;
  const float gden = __fadd_rn(sqrtf(sample_sum(A.p_g, b)), 1e-12f);
  const float un = sqrtf(sample_sum(A.p_u, b));
  const float wn = sqrtf(sample_sum(A.p_w, b));
  const float uden = __fadd_rn(un, 1e-12f), umin = fminf(A.eps, un);
  const float wden = __fadd_rn(wn, 1e-12f), wmin = fminf(A.eps, wn);
  const float s = A.step_size[b];
  const float oma = (float)(1.0 - (double)A.a);
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V; i < A.n;
       i += (long long)gridDim.x * blockDim.x * V) {
    Pack<V> pa = ldp<V>(A.x_adv, base + i), po = ldp<V>(A.x_adv_old, base + i);
    const Pack<V> pg = ldp<V>(A.grad, base + i), pc = ldp<V>(A.x, base + i);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xa = pa.v[j];
      const float w = apgd_l2_w(xa, po.v[j], pc.v[j], pg.v[j], s, gden, uden, umin, A.a, oma);
      po.v[j] = xa;
      pa.v[j] = apgd_l2_proj(pc.v[j], w, wden, wmin);
    }
    stp<V>(A.x_adv_old, base + i, po);
    stp<V>(A.x_adv, base + i, pa);
  }
}

// ---------------------------------------------------------------------------
// APGD bookkeeping.  Small per-sample kernel computes masks; a wide kernel applies
// them to the image-sized state.  The check-point schedule (k, counter3) is data
// independent, so the host passes `k_check` (> 0 only on check-point iterations).
// ---------------------------------------------------------------------------
struct ApgdBookArgs {
  int B, n_iter, iter, k_check;
  float k_thr;                // (float)(k * rho), computed by the host in double
  const float* loss_indiv;   // [B] loss of the new x_adv
  const int* pred;           // [B] 1 if argmax(logits) == y
  float* loss_steps;         // [n_iter, B]
  float* loss_best;          // [B]
  float* loss_best_last;     // [B]
  float* reduced_last;       // [B]
  float* step_size;          // [B]
  int* acc;                  // [B]
  int* flags;                // [B] bit0: copy x_adv->x_best_adv; bit1: new best; bit2: restart
};

__global__ void apgd_book_small_kernel(ApgdBookArgs A) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= A.B) return;
  const int pred = A.pred[b];
  A.acc[b] = A.acc[b] & pred;                                  // acc = min(acc, pred)
  int f = pred ? 0 : 1;                                        // x_best_adv[pred==0] = x_adv
  const float y1 = A.loss_indiv[b];
  A.loss_steps[(size_t)A.iter * A.B + b] = y1;
  float lb = A.loss_best[b];
  if (y1 > lb) { f |= 2; lb = y1; A.loss_best[b] = lb; }
  if (A.k_check > 0) {
    const int k = A.k_check, j = A.iter, n = A.n_iter;
    float t = 0.f;
    for (int c = 0; c < k; ++c) {
      const int i0 = ((j - c) % n + n) % n, i1 = ((j - c - 1) % n + n) % n;
      t += (A.loss_steps[(size_t)i0 * A.B + b] > A.loss_steps[(size_t)i1 * A.B + b]) ? 1.f : 0.f;
    }
    const float osc = (t <= A.k_thr) ? 1.f : 0.f;
    const float noimpr = (1.f - A.reduced_last[b]) * ((A.loss_best_last[b] >= lb) ? 1.f : 0.f);
    const float flag = fmaxf(osc, noimpr);
    A.reduced_last[b] = flag;
    A.loss_best_last[b] = lb;
    if (flag > 0.f) {
      f |= 4;
      A.step_size[b] = A.step_size[b] / 2.0f;
    }
  }
  A.flags[b] = f;
}

__global__ void apgd_book_apply_kernel(float* __restrict__ x_adv, float* __restrict__ grad,
                                       float* __restrict__ x_best, float* __restrict__ grad_best,
                                       float* __restrict__ x_best_adv,
                                       const int* __restrict__ flags, long long n) {
  const int b = blockIdx.y;
  const int f = flags[b];
  if (f == 0) return;
  const long long base = (long long)b * n;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 xa = reinterpret_cast<float4*>(x_adv + base)[i];
    float4 g = reinterpret_cast<float4*>(grad + base)[i];
    if (f & 1) reinterpret_cast<float4*>(x_best_adv + base)[i] = xa;
    if (f & 2) {
      reinterpret_cast<float4*>(x_best + base)[i] = xa;
      reinterpret_cast<float4*>(grad_best + base)[i] = g;
    }
    if (f & 4) {
      // restart from the best point; if it was just updated above, that is x_adv itself
      if (!(f & 2)) {
        reinterpret_cast<float4*>(x_adv + base)[i] = reinterpret_cast<float4*>(x_best + base)[i];
        reinterpret_cast<float4*>(grad + base)[i] = reinterpret_cast<float4*>(grad_best + base)[i];
      }
    }
  }
}

// out = a + b   (x_adv = x + delta, train/pgd_train.py:68)
__global__ void add_kernel(float* __restrict__ out, const float* __restrict__ a,
                           const float* __restrict__ b, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    x.x = __fadd_rn(x.x, y.x); x.y = __fadd_rn(x.y, y.y);
    x.z = __fadd_rn(x.z, y.z); x.w = __fadd_rn(x.w, y.w);
    reinterpret_cast<float4*>(out)[i] = x;
  }
}

// out = a * b   (grad *= mask, vlm_eval/attacks/apgd.py:225,297)
__global__ void mul_kernel(float* __restrict__ out, const float* __restrict__ a,
                           const float* __restrict__ b, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    x.x = __fmul_rn(x.x, y.x); x.y = __fmul_rn(x.y, y.y);
    x.z = __fmul_rn(x.z, y.z); x.w = __fmul_rn(x.w, y.w);
    reinterpret_cast<float4*>(out)[i] = x;
  }
}

static inline dim3 upd_grid(long long work_items, int B) {
  long long blocks = (work_items + UPD_THREADS - 1) / UPD_THREADS;
  long long cap = (148LL * 8 + B - 1) / B;  // ~8 CTAs per SM in total
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dim3((unsigned)blocks, (unsigned)B);
}

void count_launch(int n);
static inline int cuda_rc(int n = 1) {
  count_launch(n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

}  // namespace rvlm

using namespace rvlm;

extern "C" size_t rvlm_update_workspace_bytes(int B) {
  return (size_t)B * RED_BLOCKS * sizeof(float) * 3;
}

extern "C" int rvlm_pgd_update_linf_masked(float* delta, float* velocity, const float* grad,
                                           const float* x, const float* mask, int image_space,
                                           int B, long long n, float stepsize, float eps,
                                           int maximize, float momentum, void* stream) {
  if (B <= 0 || n <= 0 || (n & 3)) return RVLM_ERR_INVAL;
  pgd_update_linf_kernel<<<upd_grid(n >> 2, B), UPD_THREADS, 0, (cudaStream_t)stream>>>(
      delta, velocity, grad, x, mask, image_space, n, stepsize, eps, maximize ? 1.f : -1.f,
      momentum);
  return cuda_rc();
}

extern "C" int rvlm_pgd_update_linf(float* delta, float* velocity, const float* grad,
                                    const float* x, int B, long long n, float stepsize, float eps,
                                    int maximize, float momentum, void* stream) {
  return rvlm_pgd_update_linf_masked(delta, velocity, grad, x, nullptr, 1, B, n, stepsize, eps,
                                     maximize, momentum, stream);
}

extern "C" int rvlm_pgd_update_l2(float* delta, float* velocity, const float* grad, const float* x,
                                  int B, long long n, float stepsize, float eps, int maximize,
                                  float momentum, float* workspace, void* stream) {
  return rvlm_pgd_update_l2_masked(delta, velocity, grad, x, nullptr, 1, B, n, stepsize, eps,
                                   maximize, momentum, workspace, stream);
}

extern "C" int rvlm_pgd_update_l2_masked(float* delta, float* velocity, const float* grad,
                                         const float* x, const float* mask, int image_space,
                                         int B, long long n, float stepsize, float eps,
                                         int maximize, float momentum, float* workspace,
                                         void* stream) {
  if (B <= 0 || n <= 0 || !workspace) return RVLM_ERR_INVAL;
  cudaStream_t st = (cudaStream_t)stream;
  float* pg = workspace;
  float* pm = workspace + (size_t)B * RED_BLOCKS;
  float* pd = workspace + (size_t)2 * B * RED_BLOCKS;
  const dim3 grid(RED_BLOCKS, B);
  const float dir = maximize ? 1.f : -1.f;
  if ((n & 3) == 0) {
    l2_sumsq_kernel<4><<<grid, UPD_THREADS, 0, st>>>(grad, nullptr, n, pg, 1, mask);
    pgd_l2_momentum_kernel<4><<<grid, UPD_THREADS, 0, st>>>(velocity, grad, mask, n, momentum, pg, pm);
    pgd_l2_step_kernel<4><<<grid, UPD_THREADS, 0, st>>>(delta, velocity, n, stepsize, dir, pm, pd);
    pgd_l2_project_kernel<4><<<grid, UPD_THREADS, 0, st>>>(delta, x, image_space, n, eps, pd);
  } else {
    l2_sumsq_kernel<1><<<grid, UPD_THREADS, 0, st>>>(grad, nullptr, n, pg, 1, mask);
    pgd_l2_momentum_kernel<1><<<grid, UPD_THREADS, 0, st>>>(velocity, grad, mask, n, momentum, pg, pm);
    pgd_l2_step_kernel<1><<<grid, UPD_THREADS, 0, st>>>(delta, velocity, n, stepsize, dir, pm, pd);
    pgd_l2_project_kernel<1><<<grid, UPD_THREADS, 0, st>>>(delta, x, image_space, n, eps, pd);
  }
  return cuda_rc(4);
}

extern "C" int rvlm_apgd_step_linf(float* x_adv, float* x_adv_old, const float* x,
                                   const float* grad, const float* step_size, int B, long long n,
                                   float eps, float a, void* stream) {
  if (B <= 0 || n <= 0) return RVLM_ERR_INVAL;
  if ((n & 3) == 0)
    apgd_step_linf_kernel<4><<<upd_grid(n >> 2, B), UPD_THREADS, 0, (cudaStream_t)stream>>>(
        x_adv, x_adv_old, x, grad, step_size, n, eps, a);
  else
    apgd_step_linf_kernel<1><<<upd_grid(n, B), UPD_THREADS, 0, (cudaStream_t)stream>>>(
        x_adv, x_adv_old, x, grad, step_size, n, eps, a);
  return cuda_rc();
}

extern "C" int rvlm_apgd_step_l2(float* x_adv, float* x_adv_old, const float* x, const float* grad,
                                 const float* step_size, int B, long long n, float eps, float a,
                                 float* workspace, void* stream) {
  if (B <= 0 || n <= 0 || !workspace) return RVLM_ERR_INVAL;
  cudaStream_t st = (cudaStream_t)stream;
  ApgdL2Args A{x_adv, x_adv_old, x, grad, step_size, n, eps, a, workspace,
               workspace + (size_t)B * RED_BLOCKS, workspace + (size_t)2 * B * RED_BLOCKS};
  const dim3 grid(RED_BLOCKS, B);
  if ((n & 3) == 0) {
    l2_sumsq_kernel<4><<<grid, UPD_THREADS, 0, st>>>(grad, nullptr, n, workspace, 0, nullptr);
    apgd_l2_pass_u<4><<<grid, UPD_THREADS, 0, st>>>(A);
    apgd_l2_pass_w<4><<<grid, UPD_THREADS, 0, st>>>(A);
    apgd_l2_pass_final<4><<<grid, UPD_THREADS, 0, st>>>(A);
  } else {
    l2_sumsq_kernel<1><<<grid, UPD_THREADS, 0, st>>>(grad, nullptr, n, workspace, 0, nullptr);
    apgd_l2_pass_u<1><<<grid, UPD_THREADS, 0, st>>>(A);
    apgd_l2_pass_w<1><<<grid, UPD_THREADS, 0, st>>>(A);
    apgd_l2_pass_final<1><<<grid, UPD_THREADS, 0, st>>>(A);
  }
  return cuda_rc(4);
}

extern "C" int rvlm_apgd_bookkeep(float* x_adv, float* grad, float* x_best, float* grad_best,
                                  float* x_best_adv, const float* loss_indiv, const int* pred,
                                  float* loss_steps, float* loss_best, float* loss_best_last,
                                  float* reduced_last, float* step_size, int* acc, int* flags,
                                  int B, long long n, int n_iter, int iter, int k_check, float k_thr,
                                  void* stream) {
  if (B <= 0 || n <= 0 || (n & 3) || iter < 0 || iter >= n_iter) return RVLM_ERR_INVAL;
  cudaStream_t st = (cudaStream_t)stream;
  ApgdBookArgs A{B, n_iter, iter, k_check, k_thr, loss_indiv, pred, loss_steps, loss_best,
                 loss_best_last, reduced_last, step_size, acc, flags};
  apgd_book_small_kernel<<<(B + 127) / 128, 128, 0, st>>>(A);
  apgd_book_apply_kernel<<<upd_grid(n >> 2, B), UPD_THREADS, 0, st>>>(x_adv, grad, x_best,
                                                                      grad_best, x_best_adv, flags,
                                                                      n);
  return cuda_rc(2);
}

extern "C" int rvlm_mul_f32(float* out, const float* a, const float* b, long long n, void* stream) {
  if (n <= 0 || (n & 3)) return RVLM_ERR_INVAL;
  long long blocks = ((n >> 2) + UPD_THREADS - 1) / UPD_THREADS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mul_kernel<<<(unsigned)blocks, UPD_THREADS, 0, (cudaStream_t)stream>>>(out, a, b, n >> 2);
  return cuda_rc();
}

extern "C" int rvlm_add_f32(float* out, const float* a, const float* b, long long n, void* stream) {
  if (n <= 0 || (n & 3)) return RVLM_ERR_INVAL;
  long long blocks = ((n >> 2) + UPD_THREADS - 1) / UPD_THREADS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  add_kernel<<<(unsigned)blocks, UPD_THREADS, 0, (cudaStream_t)stream>>>(out, a, b, n >> 2);
  return cuda_rc();
}
