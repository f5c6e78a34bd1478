// Internal (non-ABI) declarations shared between the engine and its kernels.
#pragma once
#include "ptx.cuh"

namespace rvlm {

// ---- head.cu -------------------------------------------------------------
struct HeadArgs {
  const float* xL;        // [B*T, W]
  int T, W, D, C;
  const float* ln_g;
  const float* ln_b;
  const float* proj;      // [W, D]
  int mode;               // 0 none, 1 l2 (FARE), 2 ce (TeCoA / zero-shot classifier),
                          // 3 external d(loss)/d(emb) passed in emb_orig,
                          // 4 DLR, 5 targeted DLR on the class logits (AutoAttack's APGD losses)
  int normalize;          // F.normalize(emb) before the loss / output
  int pred_on_logits;     // 1: pred from class logits, 0: from the embedding
  const float* emb_orig;  // [B, D]      (mode 1)
  const float* text;      // [D, C]      (mode 2)
  float logit_scale;
  const long long* y;     // [B]         (may be null -> pred = 0)
  const long long* y_target;  // [B]     (mode 5)
  float* emb_out;         // [B, D]      (may be null)
  float* logits_out;      // [B, C]      (may be null)
  float* loss;            // [B]         (may be null)
  int* pred;              // [B]         (may be null)
  float* dx32;            // [B*T, W]    (may be null -> no backward)
  uint16_t* dx16;         // [B*T, W]    (may be null)
  int fmt_g;
  float grad_scale;       // multiplies d loss_b / d emb
  int grad_unit;          // 1: rescale each sample's CLS-row gradient to max-abs 64 (fp16 range)
  float* unit_out;        // [B] the factor applied (1 when grad_unit == 0); may be null
  // outer training step only (may be null): per-sample vectors for the head's parameter grads
  float* t_xhat;          // [B, W] normalised CLS row
  float* t_pooled;        // [B, W] ln_post output
  float* t_dpooled;       // [B, W] d loss / d pooled
  float* t_demb;          // [B, D] d loss / d (un-normalised) embedding
};
int launch_head(const HeadArgs& A, int B, cudaStream_t st);

// ---- train_kernels.cu ------------------------------------------------------
constexpr int TRAIN_RSPLIT = 64;  // row splits of the two-stage column reductions
// out[N] = column sums of a 16-bit [M, N] matrix; scratch: TRAIN_RSPLIT * N floats
int launch_colsum16(const uint16_t* a, int M, int N, int fmt, float* scratch, float* out,
                    cudaStream_t st);
// LayerNorm affine gradients; scratch: 2 * TRAIN_RSPLIT * W floats
int launch_ln_param_grads(const float* dy, const float* x, const float* stats, int M, int W,
                          float* scratch, float* dgamma, float* dbeta, cudaStream_t st);
int launch_batch_sum(const float* in, float* out, int B, long long n, cudaStream_t st);
int launch_head_param_grads(const float* pooled, const float* demb, const float* dpooled,
                            const float* xhat, int B, int W, int D, float* dproj, float* dgamma,
                            float* dbeta, cudaStream_t st);

// GEMM epilogue: conv1 weight gradient [W, kp] (the GEMM's N is kp rounded up to 32)
struct EpiWgradPatch {
  static constexpr const char* kTag = "gemm.wgrad_patch";
  float* G;
  int kp;
  struct Aux {};
  struct Col {};
  __device__ __forceinline__ Aux load(int, int) const { return {}; }
  __device__ __forceinline__ Col col(int) const { return {}; }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux&,
                                             const Col&) const {
    if (col < kp) *reinterpret_cast<float4*>(G + (size_t)row * kp + col) = acc;
  }
};

// ---- patch.cu ------------------------------------------------------------
struct PatchGeom {
  int image, patch, grid, np, kp;  // np = grid*grid patches / image, kp = 3*patch*patch
  float mean[3], inv_std[3];
};
// NCHW fp32 image batch <-> patch-major [B*np, kp] fp32
int launch_patchify(const float* img, float* pm, int B, const PatchGeom& g, cudaStream_t st);
// unit (nullable, [B]): img = pm / unit[b] (undoes the head's per-sample gradient rescaling)
int launch_unpatchify(const float* pm, float* img, int B, const PatchGeom& g, cudaStream_t st,
                      const float* unit = nullptr);
// A16[r, k] = 16-bit((x[r,k] + delta[r,k] - mean_c) / std_c); delta may be null
int launch_patch_operand(const float* xp, const float* dp, uint16_t* a16, int lda, long long rows,
                         const PatchGeom& g, int fmt, cudaStream_t st, uint16_t* a16lo = nullptr);
// tok[b, 0, :] = cls + pos[0]
int launch_cls_rows(float* tok, const float* cls, const float* pos, int B, int T, int W,
                    cudaStream_t st);
// out = x + delta (patch-major) written as NCHW image
int launch_unpatchify_add(const float* xp, const float* dp, float* img, int B, const PatchGeom& g,
                          cudaStream_t st);

// GEMM epilogue: patch-embedding forward.  Row r = b*np + p  ->  token row b*T + 1 + p.
struct EpiPatchFwd {
  static constexpr const char* kTag = "gemm.patch_fwd";
  float* tok;        // [B*T, W]
  const float* pos;  // [T, W]
  int np, T, W;
  struct Aux {};
  struct Col {};
  __device__ __forceinline__ Aux load(int, int) const { return {}; }
  __device__ __forceinline__ Col col(int) const { return {}; }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux&,
                                             const Col&) const {
    const int b = row / np, p = row - b * np;
    const float4 q = __ldg(reinterpret_cast<const float4*>(pos + (size_t)(1 + p) * W + col));
    *reinterpret_cast<float4*>(tok + ((size_t)b * T + 1 + p) * W + col) =
        make_float4(acc.x + q.x, acc.y + q.y, acc.z + q.z, acc.w + q.w);
  }
};

// GEMM epilogue: patch-embedding dgrad -> input gradient in patch-major layout,
// G[r, k] = acc / std_c  (chain rule through Normalize, train/adversarial_training_clip.py:254)
struct EpiPatchGrad {
  static constexpr const char* kTag = "gemm.patch_grad";
  float* G;  // [rows, kp]
  int kp, pp;  // pp = patch*patch; kp % 4 == 0
  float inv_std[3];
  struct Aux {};
  struct Col {};
  __device__ __forceinline__ Aux load(int, int) const { return {}; }
  __device__ __forceinline__ Col col(int) const { return {}; }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux&,
                                             const Col&) const {
    if (col >= kp) return;
    // one channel per piece (patch * patch and col are multiples of 4); no run-time indexing of
    // the functor's arrays
    const int ch = col / pp;
    const float is = ch == 0 ? inv_std[0] : (ch == 1 ? inv_std[1] : inv_std[2]);
    *reinterpret_cast<float4*>(G + (size_t)row * kp + col) =
        make_float4(acc.x * is, acc.y * is, acc.z * is, acc.w * is);
  }
};

// GEMM epilogue: patch-embedding dgrad with the ell_inf PGD update fused in
// (train/pgd_train.py:37-59): NaN scrub, sign, momentum, +-stepsize, clamp(+-eps),
// delta = clamp(x+delta,0,1)-x, and the next iteration's normalised 16-bit GEMM operand,
// so no separate elementwise kernel runs per PGD step.  sign(acc / std) == sign(acc).
struct EpiPatchPgdLinf {
  static constexpr const char* kTag = "gemm.patch_pgd_linf";
  static constexpr int kPiecesPerBatch = 4;   // 13 live registers per piece on top of the accumulator
  const float* xp;   // [rows, kp]
  float* dp;         // [rows, kp] delta (in place)
  float* vp;         // [rows, kp] velocity (in place)
  uint16_t* a16;     // [rows, lda] next forward operand
  uint16_t* a16lo;   // may be null: its fp16 lo half (split-operand forward)
  int kp, pp, lda, fmt;
  float stepsize, eps, dir, momentum;
  float mean[3], inv_std[3];
  // delta and velocity (read-modify-write) are requested EARLY through the Aux mechanism -- inside
  // operator() their loads could not be hoisted above the previous piece's stores to the same
  // arrays, and eight dependent load -> store round trips per chunk made this GEMM 516 us
  // (profiles/r02_kernels_ncu.txt) against 86 us for the plain patch dgrad; x is read-only (ld.nc)
  // (inside rvlm_pgd_run the velocity is 0 at the start and sign(...) afterwards, i.e. always in
  // {-1, 0, +1}: it is carried through the Aux registers as four 2-bit codes)
  // (x too: read inside operator() its eight loads per granule sat behind the early-exit branch,
  // one exposed round trip per piece)
  struct Aux { float4 d, x; uint32_t vs; };
  struct Col {};
  __device__ __forceinline__ Aux load(int row, int col) const {
    // unconditional loads (pad columns read the row's last valid piece; operator() ignores them):
    // a conditionally initialised Aux array ends up in local memory, and every load is then
    // followed by a store that waits for it
    const size_t off = (size_t)row * kp + (col < kp ? col : kp - 4);
    Aux a;
    a.d = *reinterpret_cast<const float4*>(dp + off);
    a.x = __ldg(reinterpret_cast<const float4*>(xp + off));
    const float4 v = *reinterpret_cast<const float4*>(vp + off);
    a.vs = (v.x > 0.f ? 1u : (v.x < 0.f ? 2u : 0u)) | (v.y > 0.f ? 4u : (v.y < 0.f ? 8u : 0u)) |
           (v.z > 0.f ? 16u : (v.z < 0.f ? 32u : 0u)) | (v.w > 0.f ? 64u : (v.w < 0.f ? 128u : 0u));
    return a;
  }
  __device__ __forceinline__ Col col(int) const { return {}; }
  __device__ __forceinline__ void operator()(int row, int col, const float4& acc, const Aux& aux,
                                             const Col&) const {
    if (col >= kp) return;
    const size_t off = (size_t)row * kp + col;
    struct { float4 x, d, v; } a;
    a.x = aux.x;
    a.d = aux.d;
    {
      const uint32_t c = aux.vs;
      a.v = make_float4((c & 1u) ? 1.f : ((c & 2u) ? -1.f : 0.f), (c & 4u) ? 1.f : ((c & 8u) ? -1.f : 0.f),
                        (c & 16u) ? 1.f : ((c & 32u) ? -1.f : 0.f),
                        (c & 64u) ? 1.f : ((c & 128u) ? -1.f : 0.f));
    }
    const float4 x4 = a.x;
    float4 d4 = a.d;
    float4 v4 = a.v;
    const float* gg = &acc.x;
    const float* xx = &x4.x;
    float* dd = &d4.x;
    float* vv = &v4.x;
    float nx[4];
    // the channel of all four columns (patch * patch is a multiple of 4, col too); selected
    // without indexing the functor's arrays at run time -- that would copy it to local memory
    const int ch = col / pp;
    const float mean_c = ch == 0 ? mean[0] : (ch == 1 ? mean[1] : mean[2]);
    const float istd_c = ch == 0 ? inv_std[0] : (ch == 1 ? inv_std[1] : inv_std[2]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float g = gg[i];
      if (g != g) g = 0.f;
      const float gs = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);
      const float m = __fadd_rn(__fmul_rn(momentum, vv[i]), gs);
      const float vn = (m > 0.f) ? 1.f : ((m < 0.f) ? -1.f : 0.f);
      float dn = __fadd_rn(dd[i], __fmul_rn(__fmul_rn(stepsize, vn), dir));
      dn = fminf(fmaxf(dn, -eps), eps);
      const float xa = fminf(fmaxf(__fadd_rn(xx[i], dn), 0.f), 1.f);
      dd[i] = __fsub_rn(xa, xx[i]);
      vv[i] = vn;
      nx[i] = (__fadd_rn(xx[i], dd[i]) - mean_c) * istd_c;
    }
    *reinterpret_cast<float4*>(dp + off) = d4;
    *reinterpret_cast<float4*>(vp + off) = v4;
    uint2 pk;
    pk.x = pack2(nx[0], nx[1], fmt);
    pk.y = pack2(nx[2], nx[3], fmt);
    *reinterpret_cast<uint2*>(a16 + (size_t)row * lda + col) = pk;
    if (a16lo) {
      const float2 h0 = unpack2(pk.x, 0), h1 = unpack2(pk.y, 0);
      uint2 lo;
      lo.x = pack2(nx[0] - h0.x, nx[1] - h0.y, 0);
      lo.y = pack2(nx[2] - h1.x, nx[3] - h1.y, 0);
      *reinterpret_cast<uint2*>(a16lo + (size_t)row * lda + col) = lo;
    }
  }
};

// ---- misc elementwise (convert.cu) ----------------------------------------
// fmt 0 / 1 = fp16 / bf16; fmt 2 = the fp16 lo half of a split operand, fp16(v - float(fp16(v)))
int launch_convert16(const float* src, uint16_t* dst, long long n, int fmt, cudaStream_t st);
// dst = fmt_dst(src) for a 16-bit buffer in fmt_src (n % 8 == 0)
int launch_recast16(const uint16_t* src, uint16_t* dst, long long n, int fmt_src, int fmt_dst,
                    cudaStream_t st);
// dst[c, r] (pitch ldd) = 16-bit(src[r, c]) for src [R, C] row-major
int launch_transpose_convert16(const float* src, uint16_t* dst, int R, int C, int ldd, int fmt,
                               cudaStream_t st);
// dst[r, c] (pitch ldd >= C) = 16-bit(src[r, c]); pad columns untouched
int launch_convert16_pitched(const float* src, uint16_t* dst, int R, int C, int ldd, int fmt,
                             cudaStream_t st);

}  // namespace rvlm
