// The fused CLIP-ViT attack engine: forward, input-gradient backward and the whole
// PGD inner loop of the reference (train/pgd_train.py:29-68 around
// ClipVisionModel.forward, train/adversarial_training_clip.py:253-257) as a stream
// of hand-written sm_100a kernels with no host synchronisation.
//
// Data layout in HBM (M = B*T token rows, W = width):
//   residual stream x_l, x_mid_l       fp32  [M, W]      (kept per layer for the backward)
//   GEMM operands / activations        16-bit [M, *]     (fmt_fwd forward, fmt_bwd gradients)
//   attack state x, delta, velocity    fp32  [B*np, 3*p*p]  patch-major for the whole loop
//   weights                            16-bit, [N, K] row-major (K-major UMMA operands):
//                                      forward copies in fmt_fwd, transposed dgrad copies
//                                      in fmt_bwd
// Backward is dgrad-only (torch.autograd.grad(loss, perturbation), train/pgd_train.py:38).
#include "gemm.cuh"
#include "epilogues.cuh"
#include "engine_internal.cuh"
#include "rvlm_debug.h"

#include <cstdlib>
#include <vector>
#include <new>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost a null-pointer test unless a tool is attached

using namespace rvlm;

extern "C" int rvlm_attn_fwd_simt(const void*, void*, float*, int, int, int, int, void*);
extern "C" int rvlm_attn_bwd_simt(const void*, const void*, const void*, float*, void*, int, int,
                                  int, int, int, int, void*);
extern "C" int rvlm_attn_fwd(const void*, void*, float*, int, int, int, int, void*);
extern "C" int rvlm_attn_bwd(const void*, const void*, const void*, float*, void*, int, int, int,
                             int, int, int, void*);
extern "C" int rvlm_attn_bwd_lse(const void*, const void*, const void*, float*, void*, int, int,
                                 int, int, int, int, void*);
extern "C" int rvlm_layernorm_fwd(const float*, const float*, const float*, void*, float*, float*,
                                  int, int, int, void*);
extern "C" int rvlm_layernorm_bwd(const float*, const float*, const float*, const float*,
                                  const float*, float*, void*, int, int, int, int, void*);
extern "C" int rvlm_layernorm_bwd16(const void*, const float*, const float*, const float*,
                                    const float*, float*, void*, int, int, int, int, void*);

namespace rvlm {
int attn_cls_fwd(const void*, void*, int, int, int, int, cudaStream_t, const void*, void*);
int attn_fwd_split(const void*, const void*, void*, void*, float*, int, int, int, cudaStream_t);
int attn_cls_bwd(const void*, const void*, void*, int, int, int, int, cudaStream_t);
int ln_fwd_strided(const float*, long long, const float*, const float*, void*, float*, long long,
                   float*, int, int, int, cudaStream_t, void* y16lo = nullptr);
int ln_bwd16_strided(const void*, long long, const float*, long long, const float*, const float*,
                     const float*, long long, float*, void*, long long, int, int, int,
                     cudaStream_t);
}  // namespace rvlm

namespace {

// NVTX range around a phase of the step (visible in Nsight Systems / ncu --nvtx): forward, head +
// loss, input-gradient backward, delta update, outer backward.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct LayerDev {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *bqkv, *bo, *bfc, *bpr;
  uint16_t *wqkv, *wo, *wfc, *wpr;      // forward operands [N, K], fmt_fwd
  uint16_t *wqkv_lo, *wo_lo, *wfc_lo, *wpr_lo;   // their fp16 lo halves (split forward), else null
  uint16_t *wqkvT, *woT, *wfcT, *wprT;  // dgrad operands   [K_fwd, N_fwd] = [N', K'], fmt_bwd
};

struct Engine {
  rvlm_vit_config cfg;
  int T, np, kp, kld, kp32, W, D, L, H, F;
  PatchGeom geom;
  uint8_t* slab = nullptr;
  size_t slab_bytes = 0;
  float *cls, *pos, *lnpre_g, *lnpre_b, *lnpost_g, *lnpost_b, *proj;
  uint16_t *conv_w, *conv_wT;
  uint16_t* conv_w_lo = nullptr;
  std::vector<LayerDev> layers;
  int ff = 0;        // 16-bit format of the forward operands (fp16 when split)
  int split = 0;     // cfg.fmt_fwd == RVLM_FMT_F16X2: forward operands are fp16 hi + lo pairs
  // forward and backward multiply the same q | k | v bits (one 16-bit format, forward not split):
  // ONE q|k|v copy serves both, and the backward takes the forward's softmax statistics as they
  // are (rvlm_attn_bwd_lse) instead of recomputing them
  int same_qkv = 0;
  int attn_simt = 0;
  int cls_last = 1;  // last block: class-token row only (flag RVLM_VIT_FULL_LAST_BLOCK / env RVLM_CLS_LAST=0 disable)
  int ln_dz16 = 1;   // attack-path dgrad GEMMs hand LayerNorm' a 16-bit gradient (RVLM_LN_DZ16=0: fp32)
};

// ---- workspace carving --------------------------------------------------------
struct Ws {
  // attack state (patch-major)
  float *xp, *dp, *vp, *gp;
  uint16_t* a16;
  // forward
  float *tok32, *stats_pre;
  float *x32;     // [(L+1)] or [2] x [M, W]
  float *xm32;    // [L] or [1]
  float *stats1, *stats2;  // [L] x [M, 2]
  // [L] or [1]; qkv16 is the backward's copy (fmt_bwd); h16 holds QuickGELU'(pre-activation) - 1/2, the
  // factor the MLP dgrad multiplies by (the pre-activation itself is never stored)
  uint16_t *qkv16, *o16, *h16;
  uint16_t* qkvf16;             // [M, 3W] transient forward-format q|k|v
  float* lse;     // [L] x [B, H, T]
  uint16_t *y16, *act16;
  // split forward: the fp16 lo halves of the transient forward operands
  uint16_t *a16lo, *y16lo, *act16lo, *qkvf16lo, *o16lo;
  // outer training step (stash == 2): per-layer GEMM inputs kept for the weight gradients
  uint16_t *ty16, *tz16, *tact16;   // [L] x [M, W], [M, W], [M, F]   (gradient format)
  uint16_t *to16, *ta16;            // mixed formats only: gradient-format copies of o [L][M, W] and the patch operand
  float *red;                        // column-reduction scratch
  float *t_xhat, *t_pooled, *t_dpooled, *t_demb;
  // backward
  float *dx32, *dz32;
  uint16_t *dx16, *dh16, *do16, *dqkv16, *dtok16;
  // head
  float *emb, *loss, *gunit;   // gunit [B]: per-sample factor the head applied to the gradient
  int* pred;
  float* upd;     // update-kernel reduction workspace
  size_t total;
  int stash;
  size_t M;
};

template <class Tp>
Tp* take(uint8_t*& p, size_t& used, size_t n_elems) {
  const size_t bytes = align_up(n_elems * sizeof(Tp), 1024);
  Tp* r = reinterpret_cast<Tp*>(p);
  if (p) p += bytes;
  used += bytes;
  return r;
}

Ws carve(const Engine& E, uint8_t* base, int B, int stash) {
  Ws w{};
  uint8_t* p = base;
  size_t used = 0;
  const size_t M = (size_t)B * E.T, W = E.W, F = E.F, L = E.L;
  const size_t R = (size_t)B * E.np;
  const size_t nl = stash ? L : 1;
  w.M = M;
  w.stash = stash;
  w.xp = take<float>(p, used, R * E.kp);
  w.dp = take<float>(p, used, R * E.kp);
  w.vp = take<float>(p, used, R * E.kp);
  w.gp = take<float>(p, used, R * E.kp);
  w.a16 = take<uint16_t>(p, used, R * E.kld);
  w.tok32 = take<float>(p, used, M * W);
  w.stats_pre = take<float>(p, used, M * 2);
  w.x32 = take<float>(p, used, (stash ? L + 1 : 2) * M * W);
  w.xm32 = take<float>(p, used, nl * M * W);
  w.stats1 = take<float>(p, used, nl * M * 2);
  w.stats2 = take<float>(p, used, nl * M * 2);
  w.qkv16 = take<uint16_t>(p, used, nl * M * 3 * W);
  w.qkvf16 = take<uint16_t>(p, used, M * 3 * W);
  w.o16 = take<uint16_t>(p, used, nl * M * W);
  w.h16 = take<uint16_t>(p, used, nl * M * F);
  w.lse = take<float>(p, used, nl * (size_t)B * E.H * E.T);
  w.y16 = take<uint16_t>(p, used, M * W);
  w.act16 = take<uint16_t>(p, used, M * F);
  w.emb = take<float>(p, used, (size_t)B * E.D);
  w.loss = take<float>(p, used, B);
  w.gunit = take<float>(p, used, B);
  w.pred = take<int>(p, used, B);
  w.upd = take<float>(p, used, (size_t)B * 32 * 3);
  if (E.split) {
    w.a16lo = take<uint16_t>(p, used, R * E.kld);
    w.y16lo = take<uint16_t>(p, used, M * W);
    w.act16lo = take<uint16_t>(p, used, M * F);
    w.qkvf16lo = take<uint16_t>(p, used, M * 3 * W);
    w.o16lo = take<uint16_t>(p, used, M * W);
  }
  if (stash == 2) {
    w.ty16 = take<uint16_t>(p, used, L * M * W);
    w.tz16 = take<uint16_t>(p, used, L * M * W);
    w.tact16 = take<uint16_t>(p, used, L * M * F);
    if (E.ff != E.cfg.fmt_bwd) {
      w.to16 = take<uint16_t>(p, used, L * M * W);
      w.ta16 = take<uint16_t>(p, used, R * E.kld);
    }
    w.red = take<float>(p, used, (size_t)2 * TRAIN_RSPLIT * (3 * W > F ? 3 * W : F));
    w.t_xhat = take<float>(p, used, (size_t)B * W);
    w.t_pooled = take<float>(p, used, (size_t)B * W);
    w.t_dpooled = take<float>(p, used, (size_t)B * W);
    w.t_demb = take<float>(p, used, (size_t)B * E.D);
  }
  if (stash) {
    w.dx32 = take<float>(p, used, M * W);
    w.dz32 = take<float>(p, used, M * W);
    w.dx16 = take<uint16_t>(p, used, M * W);
    w.dh16 = take<uint16_t>(p, used, M * F);
    w.do16 = take<uint16_t>(p, used, M * W);
    w.dqkv16 = take<uint16_t>(p, used, M * 3 * W);
    w.dtok16 = take<uint16_t>(p, used, R * W);
  }
  w.total = used;
  return w;
}

#define RV(call)                 \
  do {                           \
    int _rc = (call);            \
    if (_rc) return _rc;         \
  } while (0)

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : -(int)e - 1000; }

// ---- forward --------------------------------------------------------------------
// Precondition: w.a16 holds the normalised 16-bit patch operand of the batch.
int run_forward(const Engine& E, const Ws& w, int B, cudaStream_t st) {
  NvtxRange nvtx_("rvlm.vit_forward");
  const int M = B * E.T, W = E.W, F = E.F, R = B * E.np;
  const int ff = E.ff;
  const size_t MW = (size_t)M * W;
  void* stv = (void*)st;
  // split forward: every GEMM operand has an fp16 lo half (null pointers otherwise)
  const bool sp = E.split != 0;
  RV(launch_cls_rows(w.tok32, E.cls, E.pos, B, E.T, W, st));
  {
    EpiPatchFwd epi{w.tok32, E.pos, E.np, E.T, W};
    RV((launch_gemm_fwd(w.a16, w.a16lo, E.kld, E.conv_w, E.conv_w_lo, E.kld, R, W, E.kp, ff, epi,
                        st)));
  }
  RV(rvlm_layernorm_fwd(w.tok32, E.lnpre_g, E.lnpre_b, nullptr, w.x32, w.stats_pre, M, W, ff, stv));
  for (int l = 0; l < E.L; ++l) {
    const LayerDev& Lw = E.layers[l];
    const size_t s = w.stash ? (size_t)l : 0;
    float* x_in = w.stash ? w.x32 + (size_t)l * MW : w.x32 + (size_t)(l & 1) * MW;
    float* x_out = w.stash ? w.x32 + (size_t)(l + 1) * MW : w.x32 + (size_t)((l + 1) & 1) * MW;
    float* xm = w.xm32 + s * MW;
    uint16_t* qkv = w.qkv16 + s * MW * 3;
    uint16_t* o = w.o16 + s * MW;
    uint16_t* h = w.h16 + s * (size_t)M * F;
    float* lse = w.lse + s * (size_t)B * E.H * E.T;
    float* st1 = w.stats1 + s * (size_t)M * 2;
    float* st2 = w.stats2 + s * (size_t)M * 2;
    const bool train = w.stash == 2;
    // outer training step: the GEMM inputs are kept for the weight gradients, in the GRADIENT
    // format (one MMA = one format).  Same formats: the kept copy is the operand itself; mixed
    // (fp16 forward for a <= 1e-3 embedding, bf16 gradients): the forward runs on the transient
    // forward-format buffers and a recast copy is kept.
    const bool mixed = train && ff != E.cfg.fmt_bwd;
    uint16_t* y1 = (train && !mixed) ? w.ty16 + (size_t)l * MW : w.y16;
    uint16_t* z1 = (train && !mixed) ? w.tz16 + (size_t)l * MW : w.y16;
    uint16_t* act = (train && !mixed) ? w.tact16 + (size_t)l * M * F : w.act16;
    uint16_t* y1lo = sp ? w.y16lo : nullptr;
    uint16_t* qf = w.qkvf16;   // q|k|v the forward attention reads
    RV(ln_fwd_strided(x_in, 0, Lw.ln1_g, Lw.ln1_b, y1, nullptr, 0, st1, M, W, ff, st, y1lo));
    {  // q|k|v in the forward format (transient) + the backward's copy in the gradient format
      // (same_qkv: the kept copy IS the forward operand -- one store of 2 * 3W bytes per row less)
      const bool one = E.same_qkv && w.stash;
      qf = one ? qkv : w.qkvf16;
      EpiStore16Dual epi{qf, (w.stash && !one) ? qkv : nullptr, 3 * W, Lw.bqkv, ff, E.cfg.fmt_bwd,
                         sp ? w.qkvf16lo : nullptr};
      RV((launch_gemm_fwd(y1, y1lo, W, Lw.wqkv, Lw.wqkv_lo, W, M, 3 * W, W, ff, epi, st)));
    }
    if (mixed) RV(launch_recast16(y1, w.ty16 + (size_t)l * MW, (long long)MW, ff, E.cfg.fmt_bwd, st));
    if (l == E.L - 1 && E.cls_last && !train) {
      // ---- last block, class-token rows only (row pitch T*W inside the [B*T, W] streams; the
      // per-row operands of the MLP live compactly in the first B rows of their buffers)
      const long long TW = (long long)E.T * W;
      RV(attn_cls_fwd(qf, o, B, E.T, E.H, ff, st, sp ? w.qkvf16lo : nullptr,
                      sp ? w.o16lo : nullptr));
      {
        EpiStore32 epi{xm, (int)TW, Lw.bo, x_in, (int)TW};
        RV((launch_gemm_fwd(o, sp ? w.o16lo : nullptr, (int)TW, Lw.wo, Lw.wo_lo, W, B, W, W, ff, epi,
                            st)));
      }
      RV(ln_fwd_strided(xm, TW, Lw.ln2_g, Lw.ln2_b, w.y16, nullptr, 0, st2, B, W, ff, st,
                        sp ? w.y16lo : nullptr));
      {
        EpiGelu epi{w.stash ? h : nullptr, w.act16, F, Lw.bfc, ff, sp ? w.act16lo : nullptr};
        RV((launch_gemm_fwd(w.y16, sp ? w.y16lo : nullptr, W, Lw.wfc, Lw.wfc_lo, W, B, F, W, ff, epi,
                            st)));
      }
      {
        EpiStore32 epi{x_out, (int)TW, Lw.bpr, xm, (int)TW};
        RV((launch_gemm_fwd(w.act16, sp ? w.act16lo : nullptr, F, Lw.wpr, Lw.wpr_lo, F, B, W, F, ff,
                            epi, st)));
      }
      break;
    }
    if (sp) RV(attn_fwd_split(qf, w.qkvf16lo, o, w.o16lo, lse, B, E.T, E.H, st));
    else if (E.attn_simt) RV(rvlm_attn_fwd_simt(qf, o, lse, B, E.T, E.H, ff, stv));
    else RV(rvlm_attn_fwd(qf, o, lse, B, E.T, E.H, ff, stv));
    {
      EpiStore32 epi{xm, W, Lw.bo, x_in, W};
      RV((launch_gemm_fwd(o, sp ? w.o16lo : nullptr, W, Lw.wo, Lw.wo_lo, W, M, W, W, ff, epi, st)));
    }
    if (mixed) RV(launch_recast16(o, w.to16 + (size_t)l * MW, (long long)MW, ff, E.cfg.fmt_bwd, st));
    uint16_t* z1lo = sp ? w.y16lo : nullptr;
    RV(ln_fwd_strided(xm, 0, Lw.ln2_g, Lw.ln2_b, z1, nullptr, 0, st2, M, W, ff, st, z1lo));
    {
      EpiGelu epi{w.stash ? h : nullptr, act, F, Lw.bfc, ff, sp ? w.act16lo : nullptr};
      RV((launch_gemm_fwd(z1, z1lo, W, Lw.wfc, Lw.wfc_lo, W, M, F, W, ff, epi, st)));
    }
    {
      EpiStore32 epi{x_out, W, Lw.bpr, xm, W};
      RV((launch_gemm_fwd(act, sp ? w.act16lo : nullptr, F, Lw.wpr, Lw.wpr_lo, F, M, W, F, ff, epi,
                          st)));
    }
    if (mixed) {
      RV(launch_recast16(z1, w.tz16 + (size_t)l * MW, (long long)MW, ff, E.cfg.fmt_bwd, st));
      RV(launch_recast16(act, w.tact16 + (size_t)l * M * F, (long long)M * F, ff, E.cfg.fmt_bwd, st));
    }
  }
  return 0;
}

const float* final_x(const Engine& E, const Ws& w, int B) {
  const size_t MW = (size_t)B * E.T * E.W;
  return w.stash ? w.x32 + (size_t)E.L * MW : w.x32 + (size_t)(E.L & 1) * MW;
}

int run_head(const Engine& E, const Ws& w, int B, const rvlm_loss_config* lc, int normalize,
             float* emb_out, float* logits_out, bool backward, cudaStream_t st,
             float grad_scale = 1.f) {
  NvtxRange nvtx_("rvlm.head_loss");
  HeadArgs A{};
  A.xL = final_x(E, w, B);
  A.T = E.T; A.W = E.W; A.D = E.D;
  A.ln_g = E.lnpost_g; A.ln_b = E.lnpost_b; A.proj = E.proj;
  A.mode = lc ? lc->mode : 0;
  A.normalize = lc ? lc->output_normalize : normalize;
  A.C = (lc && (lc->mode == 2 || lc->mode == 4 || lc->mode == 5)) ? lc->n_classes : 0;
  A.y_target = lc ? (const long long*)lc->y_target : nullptr;
  A.pred_on_logits = lc ? lc->pred_on_logits : 0;
  A.emb_orig = lc ? lc->emb_orig : nullptr;
  A.text = lc ? lc->text : nullptr;
  A.logit_scale = lc ? lc->logit_scale : 1.f;
  A.y = lc ? (const long long*)lc->targets : nullptr;
  A.emb_out = emb_out ? emb_out : w.emb;
  A.logits_out = logits_out;
  A.loss = w.loss;
  A.pred = w.pred;
  A.fmt_g = E.cfg.fmt_bwd;
  A.grad_scale = grad_scale;
  // fp16 gradient operands: every sample's gradient is rescaled to max-abs 64 at the head (range);
  // consumers that need the true scale divide by w.gunit again (launch_unpatchify)
  A.grad_unit = (E.cfg.fmt_bwd == RVLM_FMT_F16 && w.stash != 2) ? 1 : 0;
  A.unit_out = w.gunit;
  if (w.stash == 2) {
    A.t_xhat = w.t_xhat; A.t_pooled = w.t_pooled; A.t_dpooled = w.t_dpooled; A.t_demb = w.t_demb;
  }
  if (backward) {
    const size_t MW = (size_t)B * E.T * E.W;
    cudaError_t e = cudaMemsetAsync(w.dx32, 0, MW * sizeof(float), st);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaMemsetAsync(w.dx16, 0, MW * sizeof(uint16_t), st);
    if (e != cudaSuccess) return cuda_rc(e);
    A.dx32 = w.dx32;
    A.dx16 = w.dx16;
  }
  return launch_head(A, B, st);
}

// ---- backward (dgrad only) down to the token gradient of the patch rows -----------
int run_backward(const Engine& E, const Ws& w, int B, cudaStream_t st) {
  NvtxRange nvtx_("rvlm.vit_backward_input");
  const int M = B * E.T, W = E.W, F = E.F;
  const int ff = E.ff, fg = E.cfg.fmt_bwd;
  const size_t MW = (size_t)M * W;
  void* stv = (void*)st;
  for (int l = E.L - 1; l >= 0; --l) {
    const LayerDev& Lw = E.layers[l];
    const float* x_in = w.x32 + (size_t)l * MW;
    const float* xm = w.xm32 + (size_t)l * MW;
    const uint16_t* qkv = w.qkv16 + (size_t)l * MW * 3;
    const uint16_t* o = w.o16 + (size_t)l * MW;
    const uint16_t* h = w.h16 + (size_t)l * M * F;
    float* lse = w.lse + (size_t)l * B * E.H * E.T;  // scratch for the recomputed statistics
    const float* st1 = w.stats1 + (size_t)l * M * 2;
    const float* st2 = w.stats2 + (size_t)l * M * 2;
    if (l == E.L - 1 && E.cls_last) {
      // ---- last block: the incoming gradient is non-zero on the class-token rows only
      const long long TW = (long long)E.T * W;
      {
        EpiGeluBwd epi{w.dh16, F, h, ff, fg};
        RV((launch_gemm_kmajor(w.dx16, (int)TW, Lw.wprT, W, B, F, W, fg, epi, st)));
      }
      {
        EpiStore16 epi{w.do16, W, nullptr, fg};
        RV((launch_gemm_kmajor(w.dh16, F, Lw.wfcT, F, B, W, F, fg, epi, st)));
      }
      RV(ln_bwd16_strided(w.do16, 0, xm, TW, st2, Lw.ln2_g, w.dx32, TW, w.dx32, w.dx16, TW, B, W,
                          fg, st));
      {
        EpiStore16 epi{w.do16, W, nullptr, fg};
        RV((launch_gemm_kmajor(w.dx16, (int)TW, Lw.woT, W, B, W, W, fg, epi, st)));
      }
      RV(attn_cls_bwd(qkv, w.do16, w.dqkv16, B, E.T, E.H, fg, st));
      {
        EpiStore16 epi{w.do16, W, nullptr, fg};
        RV((launch_gemm_kmajor(w.dqkv16, 3 * W, Lw.wqkvT, 3 * W, M, W, 3 * W, fg, epi, st)));
      }
      RV(rvlm_layernorm_bwd16(w.do16, x_in, st1, Lw.ln1_g, w.dx32, w.dx32, w.dx16, M, W, fg, 0, stv));
      continue;
    }
    {  // d act -> d h (through QuickGELU)
      EpiGeluBwd epi{w.dh16, F, h, ff, fg};
      RV((launch_gemm_kmajor(w.dx16, W, Lw.wprT, W, M, F, W, fg, epi, st)));
    }
    // d ln_2 output.  The gradient w.r.t. a LayerNorm output is handed over in the 16-bit
    // gradient format (w.do16 is free at both hand-overs): every other link of this chain is
    // already a 16-bit GEMM operand, and the LayerNorm' kernel is HBM-bound (18 -> 16 B/elem).
    if (E.ln_dz16) {
      EpiStore16 epi{w.do16, W, nullptr, fg};
      RV((launch_gemm_kmajor(w.dh16, F, Lw.wfcT, F, M, W, F, fg, epi, st)));
      RV(rvlm_layernorm_bwd16(w.do16, xm, st2, Lw.ln2_g, w.dx32, w.dx32, w.dx16, M, W, fg, 0, stv));
    } else {
      EpiStore32 epi{w.dz32, W, nullptr, nullptr, 0};
      RV((launch_gemm_kmajor(w.dh16, F, Lw.wfcT, F, M, W, F, fg, epi, st)));
      RV(rvlm_layernorm_bwd(w.dz32, xm, st2, Lw.ln2_g, w.dx32, w.dx32, w.dx16, M, W, fg, 0, stv));
    }
    {  // d attention output
      EpiStore16 epi{w.do16, W, nullptr, fg};
      RV((launch_gemm_kmajor(w.dx16, W, Lw.woT, W, M, W, W, fg, epi, st)));
    }
    if (E.attn_simt)
      RV(rvlm_attn_bwd_simt(qkv, o, w.do16, lse, w.dqkv16, B, E.T, E.H, fg, ff, fg, stv));
    else if (E.same_qkv)   // `lse` holds the forward's statistics of exactly these q | k bits
      RV(rvlm_attn_bwd_lse(qkv, o, w.do16, lse, w.dqkv16, B, E.T, E.H, fg, ff, fg, stv));
    else RV(rvlm_attn_bwd(qkv, o, w.do16, lse, w.dqkv16, B, E.T, E.H, fg, ff, fg, stv));
    if (E.ln_dz16) {  // d ln_1 output
      EpiStore16 epi{w.do16, W, nullptr, fg};
      RV((launch_gemm_kmajor(w.dqkv16, 3 * W, Lw.wqkvT, 3 * W, M, W, 3 * W, fg, epi, st)));
      RV(rvlm_layernorm_bwd16(w.do16, x_in, st1, Lw.ln1_g, w.dx32, w.dx32, w.dx16, M, W, fg, 0, stv));
    } else {
      EpiStore32 epi{w.dz32, W, nullptr, nullptr, 0};
      RV((launch_gemm_kmajor(w.dqkv16, 3 * W, Lw.wqkvT, 3 * W, M, W, 3 * W, fg, epi, st)));
      RV(rvlm_layernorm_bwd(w.dz32, x_in, st1, Lw.ln1_g, w.dx32, w.dx32, w.dx16, M, W, fg, 0, stv));
    }
  }
  // ln_pre backward, compacted to patch rows (the CLS row's input is a constant)
  RV(rvlm_layernorm_bwd(w.dx32, w.tok32, w.stats_pre, E.lnpre_g, nullptr, nullptr, w.dtok16, M, W,
                        fg, E.T, stv));
  return 0;
}

// Weight gradient dW[Nw, Kw] = dY^T X with dY [Mtok, Nw] and X [Mtok, Kw], both consumed
// MN-major (the token dimension is the contraction), fp32 output.
int wgrad(const uint16_t* dY, int Nw, const uint16_t* X, int Kw, int Mtok, int fmt, float* dW,
          cudaStream_t st) {
  EpiStore32 epi{dW, Kw, nullptr, nullptr, 0};
  return launch_gemm16<128, true, true>(dY, Nw, X, Kw, Nw, Kw, Mtok, fmt, epi, st);
}

// Full backward of the outer training step: dgrad chain as in run_backward plus every
// parameter gradient, written (overwritten) into `g` (open_clip layout, fp32).
// `events` (nullable): cudaEvent_t[L + 1]; events[l] is recorded once every parameter gradient of
// block l has been enqueued, events[L] at the very end (stem + head) -- the caller's hook for
// starting the all-reduce of a block's gradient bucket while the backward continues.
int run_backward_train(const Engine& E, const Ws& w, int B, const rvlm_vit_grads* g,
                       cudaStream_t st, void* const* events) {
  NvtxRange nvtx_("rvlm.vit_backward_params");
  const int M = B * E.T, W = E.W, F = E.F, R = B * E.np;
  const int ff = E.ff, fg = E.cfg.fmt_bwd;
  const size_t MW = (size_t)M * W;
  void* stv = (void*)st;
  RV(launch_head_param_grads(w.t_pooled, w.t_demb, w.t_dpooled, w.t_xhat, B, W, E.D, g->proj,
                             g->ln_post_weight, g->ln_post_bias, st));
  for (int l = E.L - 1; l >= 0; --l) {
    const LayerDev& Lw = E.layers[l];
    const float* x_in = w.x32 + (size_t)l * MW;
    const float* xm = w.xm32 + (size_t)l * MW;
    const uint16_t* qkv = w.qkv16 + (size_t)l * MW * 3;
    const uint16_t* o = w.o16 + (size_t)l * MW;
    const uint16_t* h = w.h16 + (size_t)l * M * F;
    float* lse = w.lse + (size_t)l * B * E.H * E.T;
    const float* st1 = w.stats1 + (size_t)l * M * 2;
    const float* st2 = w.stats2 + (size_t)l * M * 2;
    const uint16_t* y1 = w.ty16 + (size_t)l * MW;
    const uint16_t* z1 = w.tz16 + (size_t)l * MW;
    const uint16_t* act = w.tact16 + (size_t)l * M * F;
    // mlp.c_proj: x_{l+1} = xm + act W_pr^T + b_pr ; dY = dx
    RV(wgrad(w.dx16, W, act, F, M, fg, g->mlp_c_proj_weight[l], st));
    RV(launch_colsum16(w.dx16, M, W, fg, w.red, g->mlp_c_proj_bias[l], st));
    {
      EpiGeluBwd epi{w.dh16, F, h, ff, fg};
      RV((launch_gemm_kmajor(w.dx16, W, Lw.wprT, W, M, F, W, fg, epi, st)));
    }
    // mlp.c_fc: dY = dh
    RV(wgrad(w.dh16, F, z1, W, M, fg, g->mlp_c_fc_weight[l], st));
    RV(launch_colsum16(w.dh16, M, F, fg, w.red, g->mlp_c_fc_bias[l], st));
    {
      EpiStore32 epi{w.dz32, W, nullptr, nullptr, 0};
      RV((launch_gemm_kmajor(w.dh16, F, Lw.wfcT, F, M, W, F, fg, epi, st)));
    }
    RV(launch_ln_param_grads(w.dz32, xm, st2, M, W, w.red, g->ln_2_weight[l], g->ln_2_bias[l], st));
    RV(rvlm_layernorm_bwd(w.dz32, xm, st2, Lw.ln2_g, w.dx32, w.dx32, w.dx16, M, W, fg, 0, stv));
    // attn.out_proj: xm = x + o W_o^T + b_o ; dY = dx (now the gradient w.r.t. xm)
    RV(wgrad(w.dx16, W, ff != fg ? w.to16 + (size_t)l * MW : o, W, M, fg, g->attn_out_proj_weight[l], st));
    RV(launch_colsum16(w.dx16, M, W, fg, w.red, g->attn_out_proj_bias[l], st));
    {
      EpiStore16 epi{w.do16, W, nullptr, fg};
      RV((launch_gemm_kmajor(w.dx16, W, Lw.woT, W, M, W, W, fg, epi, st)));
    }
    if (E.attn_simt)
      RV(rvlm_attn_bwd_simt(qkv, o, w.do16, lse, w.dqkv16, B, E.T, E.H, fg, ff, fg, stv));
    else if (E.same_qkv)
      RV(rvlm_attn_bwd_lse(qkv, o, w.do16, lse, w.dqkv16, B, E.T, E.H, fg, ff, fg, stv));
    else RV(rvlm_attn_bwd(qkv, o, w.do16, lse, w.dqkv16, B, E.T, E.H, fg, ff, fg, stv));
    // attn.in_proj: dY = dqkv
    RV(wgrad(w.dqkv16, 3 * W, y1, W, M, fg, g->attn_in_proj_weight[l], st));
    RV(launch_colsum16(w.dqkv16, M, 3 * W, fg, w.red, g->attn_in_proj_bias[l], st));
    {
      EpiStore32 epi{w.dz32, W, nullptr, nullptr, 0};
      RV((launch_gemm_kmajor(w.dqkv16, 3 * W, Lw.wqkvT, 3 * W, M, W, 3 * W, fg, epi,
                                           st)));
    }
    RV(launch_ln_param_grads(w.dz32, x_in, st1, M, W, w.red, g->ln_1_weight[l], g->ln_1_bias[l], st));
    if (events) RV(cuda_rc(cudaEventRecord((cudaEvent_t)events[l], st)));
    RV(rvlm_layernorm_bwd(w.dz32, x_in, st1, Lw.ln1_g, w.dx32, w.dx32, w.dx16, M, W, fg, 0, stv));
  }
  // ln_pre: parameter grads, fp32 token gradient (positional / class embedding), 16-bit
  // compact patch rows (conv1 weight gradient)
  RV(launch_ln_param_grads(w.dx32, w.tok32, w.stats_pre, M, W, w.red, g->ln_pre_weight,
                           g->ln_pre_bias, st));
  RV(rvlm_layernorm_bwd(w.dx32, w.tok32, w.stats_pre, E.lnpre_g, nullptr, w.dz32, nullptr, M, W,
                        fg, 0, stv));
  RV(rvlm_layernorm_bwd(w.dx32, w.tok32, w.stats_pre, E.lnpre_g, nullptr, nullptr, w.dtok16, M, W,
                        fg, E.T, stv));
  RV(launch_batch_sum(w.dz32, g->positional_embedding, B, (long long)E.T * W, st));
  RV(cuda_rc(cudaMemcpyAsync(g->class_embedding, g->positional_embedding, W * sizeof(float),
                             cudaMemcpyDeviceToDevice, st)));
  {
    EpiWgradPatch epi{g->conv1_weight, E.kp};
    RV((launch_gemm16<128, true, true>(w.dtok16, W, ff != fg ? w.ta16 : w.a16, E.kld, W, E.kp32, R, fg,
                                       epi, st)));
  }
  if (events) RV(cuda_rc(cudaEventRecord((cudaEvent_t)events[E.L], st)));
  return 0;
}

int run_patch_grad(const Engine& E, const Ws& w, int B, cudaStream_t st) {
  EpiPatchGrad epi{w.gp, E.kp, E.cfg.patch * E.cfg.patch,
                   {E.geom.inv_std[0], E.geom.inv_std[1], E.geom.inv_std[2]}};
  return launch_gemm_kmajor(w.dtok16, E.W, E.conv_wT, E.W, B * E.np, E.kp32, E.W,
                                          E.cfg.fmt_bwd, epi, st);
}

int run_patch_pgd_linf(const Engine& E, const Ws& w, int B, float stepsize, float eps,
                       int maximize, float momentum, cudaStream_t st) {
  EpiPatchPgdLinf epi{w.xp, w.dp, w.vp, w.a16, E.split ? w.a16lo : nullptr, E.kp,
                      E.cfg.patch * E.cfg.patch, E.kld,
                      E.ff, stepsize, eps, maximize ? 1.f : -1.f, momentum,
                      {E.geom.mean[0], E.geom.mean[1], E.geom.mean[2]},
                      {E.geom.inv_std[0], E.geom.inv_std[1], E.geom.inv_std[2]}};
  return launch_gemm_kmajor(w.dtok16, E.W, E.conv_wT, E.W, B * E.np, E.kp32, E.W,
                                          E.cfg.fmt_bwd, epi, st);
}

int load_weights(Engine& E, const rvlm_vit_weights* wt, cudaStream_t st) {
  const int W = E.W, F = E.F, D = E.D, ff = E.ff, fg = E.cfg.fmt_bwd;
  auto cp = [&](float* dst, const float* src, size_t n) {
    return cuda_rc(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  };
  RV(cp(E.cls, wt->class_embedding, W));
  RV(cp(E.pos, wt->positional_embedding, (size_t)E.T * W));
  RV(cp(E.lnpre_g, wt->ln_pre_weight, W));
  RV(cp(E.lnpre_b, wt->ln_pre_bias, W));
  RV(cp(E.lnpost_g, wt->ln_post_weight, W));
  RV(cp(E.lnpost_b, wt->ln_post_bias, W));
  RV(cp(E.proj, wt->proj, (size_t)W * D));
  // conv1.weight [W, 3, p, p] == [W, kp] row-major
  RV(launch_convert16_pitched(wt->conv1_weight, E.conv_w, W, E.kp, E.kld, ff, st));
  if (E.split) RV(launch_convert16_pitched(wt->conv1_weight, E.conv_w_lo, W, E.kp, E.kld, 2, st));
  RV(launch_transpose_convert16(wt->conv1_weight, E.conv_wT, W, E.kp, W, fg, st));
  for (int l = 0; l < E.L; ++l) {
    LayerDev& Lw = E.layers[l];
    RV(cp(Lw.ln1_g, wt->ln_1_weight[l], W));
    RV(cp(Lw.ln1_b, wt->ln_1_bias[l], W));
    RV(cp(Lw.ln2_g, wt->ln_2_weight[l], W));
    RV(cp(Lw.ln2_b, wt->ln_2_bias[l], W));
    RV(cp(Lw.bqkv, wt->attn_in_proj_bias[l], 3 * (size_t)W));
    RV(cp(Lw.bo, wt->attn_out_proj_bias[l], W));
    RV(cp(Lw.bfc, wt->mlp_c_fc_bias[l], F));
    RV(cp(Lw.bpr, wt->mlp_c_proj_bias[l], W));
    RV(launch_convert16(wt->attn_in_proj_weight[l], Lw.wqkv, (long long)3 * W * W, ff, st));
    RV(launch_convert16(wt->attn_out_proj_weight[l], Lw.wo, (long long)W * W, ff, st));
    RV(launch_convert16(wt->mlp_c_fc_weight[l], Lw.wfc, (long long)F * W, ff, st));
    RV(launch_convert16(wt->mlp_c_proj_weight[l], Lw.wpr, (long long)W * F, ff, st));
    if (E.split) {   // fp16 lo halves: fp16(w - float(fp16(w)))
      RV(launch_convert16(wt->attn_in_proj_weight[l], Lw.wqkv_lo, (long long)3 * W * W, 2, st));
      RV(launch_convert16(wt->attn_out_proj_weight[l], Lw.wo_lo, (long long)W * W, 2, st));
      RV(launch_convert16(wt->mlp_c_fc_weight[l], Lw.wfc_lo, (long long)F * W, 2, st));
      RV(launch_convert16(wt->mlp_c_proj_weight[l], Lw.wpr_lo, (long long)W * F, 2, st));
    }
    RV(launch_transpose_convert16(wt->attn_in_proj_weight[l], Lw.wqkvT, 3 * W, W, 3 * W, fg, st));
    RV(launch_transpose_convert16(wt->attn_out_proj_weight[l], Lw.woT, W, W, W, fg, st));
    RV(launch_transpose_convert16(wt->mlp_c_fc_weight[l], Lw.wfcT, F, W, F, fg, st));
    RV(launch_transpose_convert16(wt->mlp_c_proj_weight[l], Lw.wprT, W, F, W, fg, st));
  }
  return 0;
}

int prep_state(const Engine& E, const Ws& w, int B, const float* x_img, const float* delta_img,
               bool zero_velocity, cudaStream_t st) {
  RV(launch_patchify(x_img, w.xp, B, E.geom, st));
  const size_t n = (size_t)B * E.np * E.kp;
  if (delta_img) RV(launch_patchify(delta_img, w.dp, B, E.geom, st));
  else RV(cuda_rc(cudaMemsetAsync(w.dp, 0, n * sizeof(float), st)));
  if (zero_velocity) RV(cuda_rc(cudaMemsetAsync(w.vp, 0, n * sizeof(float), st)));
  RV(launch_patch_operand(w.xp, delta_img ? w.dp : nullptr, w.a16, E.kld, (long long)B * E.np,
                          E.geom, E.ff, st, E.split ? w.a16lo : nullptr));
  return 0;
}

}  // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" int rvlm_vit_create(const rvlm_vit_config* cfg, const rvlm_vit_weights* weights,
                               void* stream, rvlm_vit_handle* out) {
  if (!cfg || !weights || !out) return RVLM_ERR_INVAL;
  if (cfg->width % 128 || cfg->width != cfg->heads * 64 || cfg->mlp % 32 || cfg->image % cfg->patch)
    return RVLM_ERR_UNSUPPORTED;
  if (cfg->width != 128 && cfg->width != 256 && cfg->width != 768 && cfg->width != 1024)
    return RVLM_ERR_UNSUPPORTED;
  Engine* E = new (std::nothrow) Engine();
  if (!E) return RVLM_ERR_NOMEM;
  E->cfg = *cfg;
  if (cfg->fmt_fwd != RVLM_FMT_F16 && cfg->fmt_fwd != RVLM_FMT_BF16 && cfg->fmt_fwd != RVLM_FMT_F16X2) {
    delete E;
    return RVLM_ERR_INVAL;
  }
  if (cfg->fmt_bwd != RVLM_FMT_F16 && cfg->fmt_bwd != RVLM_FMT_BF16) { delete E; return RVLM_ERR_INVAL; }
  E->split = cfg->fmt_fwd == RVLM_FMT_F16X2 ? 1 : 0;
  E->ff = E->split ? RVLM_FMT_F16 : cfg->fmt_fwd;
  E->same_qkv = (!E->split && cfg->fmt_fwd == cfg->fmt_bwd) ? 1 : 0;
  E->W = cfg->width; E->D = cfg->out_dim; E->L = cfg->layers; E->H = cfg->heads; E->F = cfg->mlp;
  const int g = cfg->image / cfg->patch;
  E->np = g * g;
  E->T = E->np + 1;
  E->kp = 3 * cfg->patch * cfg->patch;
  E->kld = (int)align_up(E->kp, 32);  // row pitch of the patch operand / conv weight
  E->kp32 = (int)align_up(E->kp, 32);
  E->geom.image = cfg->image; E->geom.patch = cfg->patch; E->geom.grid = g;
  E->geom.np = E->np; E->geom.kp = E->kp;
  for (int c = 0; c < 3; ++c) { E->geom.mean[c] = cfg->mean[c]; E->geom.inv_std[c] = 1.f / cfg->std[c]; }
  if (E->T > 272 || (E->kp & 3) || (E->split && E->T > 257)) { delete E; return RVLM_ERR_UNSUPPORTED; }
  E->attn_simt = (cfg->flags & RVLM_VIT_ATTN_SIMT) ? 1 : 0;
  {
    const char* e = getenv("RVLM_CLS_LAST");
    E->cls_last = ((cfg->flags & RVLM_VIT_FULL_LAST_BLOCK) || (e && e[0] == '0')) ? 0 : 1;
  }
  {
    const char* e = getenv("RVLM_LN_DZ16");
    E->ln_dz16 = (e && e[0] == '0') ? 0 : 1;
  }
  // size the slab
  const size_t W = E->W, F = E->F, D = E->D, L = E->L;
  size_t used = 0;
  uint8_t* p = nullptr;
  auto plan = [&](uint8_t* base) {
    p = base; used = 0;
    E->cls = take<float>(p, used, W); E->pos = take<float>(p, used, (size_t)E->T * W);
    E->lnpre_g = take<float>(p, used, W); E->lnpre_b = take<float>(p, used, W);
    E->lnpost_g = take<float>(p, used, W); E->lnpost_b = take<float>(p, used, W);
    E->proj = take<float>(p, used, W * D);
    E->conv_w = take<uint16_t>(p, used, W * E->kld);
    E->conv_wT = take<uint16_t>(p, used, (size_t)E->kp32 * W);
    if (E->split) E->conv_w_lo = take<uint16_t>(p, used, W * E->kld);
    E->layers.resize(L);
    for (size_t l = 0; l < L; ++l) {
      LayerDev& Lw = E->layers[l];
      Lw.ln1_g = take<float>(p, used, W); Lw.ln1_b = take<float>(p, used, W);
      Lw.ln2_g = take<float>(p, used, W); Lw.ln2_b = take<float>(p, used, W);
      Lw.bqkv = take<float>(p, used, 3 * W); Lw.bo = take<float>(p, used, W);
      Lw.bfc = take<float>(p, used, F); Lw.bpr = take<float>(p, used, W);
      Lw.wqkv = take<uint16_t>(p, used, 3 * W * W); Lw.wo = take<uint16_t>(p, used, W * W);
      Lw.wfc = take<uint16_t>(p, used, F * W); Lw.wpr = take<uint16_t>(p, used, W * F);
      Lw.wqkvT = take<uint16_t>(p, used, 3 * W * W); Lw.woT = take<uint16_t>(p, used, W * W);
      Lw.wfcT = take<uint16_t>(p, used, F * W); Lw.wprT = take<uint16_t>(p, used, W * F);
      Lw.wqkv_lo = Lw.wo_lo = Lw.wfc_lo = Lw.wpr_lo = nullptr;
      if (E->split) {
        Lw.wqkv_lo = take<uint16_t>(p, used, 3 * W * W); Lw.wo_lo = take<uint16_t>(p, used, W * W);
        Lw.wfc_lo = take<uint16_t>(p, used, F * W); Lw.wpr_lo = take<uint16_t>(p, used, W * F);
      }
    }
  };
  plan(nullptr);
  E->slab_bytes = used;
  cudaError_t e = cudaMalloc(&E->slab, E->slab_bytes);
  if (e != cudaSuccess) { delete E; return cuda_rc(e); }
  plan(E->slab);
  cudaStream_t st = (cudaStream_t)stream;
  e = cudaMemsetAsync(E->slab, 0, E->slab_bytes, st);
  if (e != cudaSuccess) { cudaFree(E->slab); delete E; return cuda_rc(e); }
  int rc = load_weights(*E, weights, st);
  if (rc) { cudaFree(E->slab); delete E; return rc; }
  *out = E;
  return 0;
}

extern "C" int rvlm_vit_update_weights(rvlm_vit_handle h, const rvlm_vit_weights* weights,
                                       void* stream) {
  if (!h || !weights) return RVLM_ERR_INVAL;
  return load_weights(*reinterpret_cast<Engine*>(h), weights, (cudaStream_t)stream);
}

extern "C" int rvlm_vit_destroy(rvlm_vit_handle h) {
  if (!h) return 0;
  Engine* E = reinterpret_cast<Engine*>(h);
  cudaFree(E->slab);
  delete E;
  return 0;
}

extern "C" size_t rvlm_vit_workspace_bytes(rvlm_vit_handle h, int B, int with_backward) {
  if (!h || B <= 0) return 0;
  return carve(*reinterpret_cast<Engine*>(h), nullptr, B, with_backward).total;
}

extern "C" int rvlm_vit_forward(rvlm_vit_handle h, void* workspace, size_t workspace_bytes,
                                const float* x, const float* delta, int B, int output_normalize,
                                int keep_for_backward, float* emb_out, void* stream) {
  if (!h || !workspace || !x || !emb_out || B <= 0) return RVLM_ERR_INVAL;
  const Engine& E = *reinterpret_cast<Engine*>(h);
  Ws w = carve(E, reinterpret_cast<uint8_t*>(workspace), B, keep_for_backward ? 1 : 0);
  if (w.total > workspace_bytes) return RVLM_ERR_NOMEM;
  cudaStream_t st = (cudaStream_t)stream;
  RV(prep_state(E, w, B, x, delta, false, st));
  RV(run_forward(E, w, B, st));
  return run_head(E, w, B, nullptr, output_normalize, emb_out, nullptr, false, st);
}

extern "C" int rvlm_vit_backward_input(rvlm_vit_handle h, void* workspace,
                                       size_t workspace_bytes, int B, const float* d_emb,
                                       int output_normalize, float* grad_out, void* stream) {
  if (!h || !workspace || !d_emb || !grad_out || B <= 0) return RVLM_ERR_INVAL;
  const Engine& E = *reinterpret_cast<Engine*>(h);
  Ws w = carve(E, reinterpret_cast<uint8_t*>(workspace), B, 1);
  if (w.total > workspace_bytes) return RVLM_ERR_NOMEM;
  cudaStream_t st = (cudaStream_t)stream;
  rvlm_loss_config lc{};
  lc.mode = 3;
  lc.output_normalize = output_normalize;
  lc.emb_orig = d_emb;
  RV(run_head(E, w, B, &lc, 0, nullptr, nullptr, true, st));
  RV(run_backward(E, w, B, st));
  RV(run_patch_grad(E, w, B, st));
  return launch_unpatchify(w.gp, grad_out, B, E.geom, st, w.gunit);   // true scale (a VJP)
}

extern "C" int rvlm_vit_loss_grad(rvlm_vit_handle h, void* workspace, size_t workspace_bytes,
                                  const float* x, const float* delta, int B,
                                  const rvlm_loss_config* loss, float* grad_out, float* loss_out,
                                  int* pred_out, float* emb_out, float* logits_out, int want_grad,
                                  void* stream) {
  if (!h || !workspace || !x || !loss || B <= 0) return RVLM_ERR_INVAL;
  if (loss->mode != RVLM_LOSS_L2 && loss->mode != RVLM_LOSS_CE && loss->mode != RVLM_LOSS_DLR &&
      loss->mode != RVLM_LOSS_DLR_TARGETED)
    return RVLM_ERR_INVAL;
  if (loss->mode == RVLM_LOSS_DLR_TARGETED && !loss->y_target) return RVLM_ERR_INVAL;
  if ((loss->mode == RVLM_LOSS_DLR || loss->mode == RVLM_LOSS_DLR_TARGETED) &&
      (!loss->text || !loss->targets || loss->n_classes < 4))
    return RVLM_ERR_INVAL;
  const Engine& E = *reinterpret_cast<Engine*>(h);
  Ws w = carve(E, reinterpret_cast<uint8_t*>(workspace), B, 1);
  if (w.total > workspace_bytes) return RVLM_ERR_NOMEM;
  cudaStream_t st = (cudaStream_t)stream;
  RV(prep_state(E, w, B, x, delta, false, st));
  RV(run_forward(E, w, B, st));
  RV(run_head(E, w, B, loss, 0, emb_out, logits_out, want_grad != 0, st));
  if (loss_out)
    RV(cuda_rc(cudaMemcpyAsync(loss_out, w.loss, B * sizeof(float), cudaMemcpyDeviceToDevice, st)));
  if (pred_out)
    RV(cuda_rc(cudaMemcpyAsync(pred_out, w.pred, B * sizeof(int), cudaMemcpyDeviceToDevice, st)));
  if (want_grad) {
    if (!grad_out) return RVLM_ERR_INVAL;
    RV(run_backward(E, w, B, st));
    RV(run_patch_grad(E, w, B, st));
    RV(launch_unpatchify(w.gp, grad_out, B, E.geom, st, w.gunit));   // true scale
  }
  return 0;
}

extern "C" int rvlm_vit_train_grads(rvlm_vit_handle h, void* workspace, size_t workspace_bytes,
                                    const float* x, int B, const rvlm_loss_config* loss,
                                    float loss_scale, const rvlm_vit_grads* grads,
                                    float* loss_out, float* emb_out, void* stream) {
  return rvlm_vit_train_grads_ev(h, workspace, workspace_bytes, x, B, loss, loss_scale, grads,
                                 loss_out, emb_out, nullptr, 0, stream);
}

extern "C" int rvlm_vit_train_grads_ev(rvlm_vit_handle h, void* workspace, size_t workspace_bytes,
                                       const float* x, int B, const rvlm_loss_config* loss,
                                       float loss_scale, const rvlm_vit_grads* grads,
                                       float* loss_out, float* emb_out, void* const* events,
                                       int n_events, void* stream) {
  if (!h || !workspace || !x || !loss || !grads || B <= 0) return RVLM_ERR_INVAL;
  if (events && n_events != reinterpret_cast<Engine*>(h)->L + 1) return RVLM_ERR_INVAL;
  if (loss->mode != 1 && loss->mode != 2) return RVLM_ERR_INVAL;
  const Engine& E = *reinterpret_cast<Engine*>(h);
  if (E.split) return RVLM_ERR_UNSUPPORTED;   // (mixed fp16 / bf16 formats keep recast copies for the wgrads)
  Ws w = carve(E, reinterpret_cast<uint8_t*>(workspace), B, 2);
  if (w.total > workspace_bytes) return RVLM_ERR_NOMEM;
  cudaStream_t st = (cudaStream_t)stream;
  RV(prep_state(E, w, B, x, nullptr, false, st));
  if (E.ff != E.cfg.fmt_bwd)   // gradient-format copy of the patch operand for the conv1 weight gradient
    RV(launch_patch_operand(w.xp, nullptr, w.ta16, E.kld, (long long)B * E.np, E.geom, E.cfg.fmt_bwd, st));
  RV(run_forward(E, w, B, st));
  RV(run_head(E, w, B, loss, 0, emb_out, nullptr, true, st, loss_scale));
  if (loss_out)
    RV(cuda_rc(cudaMemcpyAsync(loss_out, w.loss, B * sizeof(float), cudaMemcpyDeviceToDevice, st)));
  return run_backward_train(E, w, B, grads, st, events);
}

extern "C" int rvlm_pgd_run(rvlm_vit_handle h, void* workspace, size_t workspace_bytes,
                            const float* x, const float* delta0, int B,
                            const rvlm_loss_config* loss, int norm, float eps, float stepsize,
                            int iterations, int maximize, float momentum, float* x_adv_out,
                            float* loss_trace, void* stream) {
  if (!h || !workspace || !x || !loss || !x_adv_out || B <= 0 || iterations < 0)
    return RVLM_ERR_INVAL;
  if (norm != RVLM_NORM_LINF && norm != RVLM_NORM_L2) return RVLM_ERR_INVAL;
  const Engine& E = *reinterpret_cast<Engine*>(h);
  Ws w = carve(E, reinterpret_cast<uint8_t*>(workspace), B, 1);
  if (w.total > workspace_bytes) return RVLM_ERR_NOMEM;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)E.np * E.kp;
  RV(prep_state(E, w, B, x, delta0, true, st));
  NvtxRange nvtx_run("rvlm.pgd_run");
  for (int it = 0; it < iterations; ++it) {
    NvtxRange nvtx_it("rvlm.pgd_iteration");
    RV(run_forward(E, w, B, st));
    RV(run_head(E, w, B, loss, 0, nullptr, nullptr, true, st));
    if (loss_trace)
      RV(cuda_rc(cudaMemcpyAsync(loss_trace + (size_t)it * B, w.loss, B * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st)));
    RV(run_backward(E, w, B, st));
    NvtxRange nvtx_up("rvlm.delta_update");
    if (norm == RVLM_NORM_LINF) {
      // delta-update fused into the last backward kernel (also emits the next operand)
      RV(run_patch_pgd_linf(E, w, B, stepsize, eps, maximize, momentum, st));
    } else {
      RV(run_patch_grad(E, w, B, st));
      RV(rvlm_pgd_update_l2(w.dp, w.vp, w.gp, w.xp, B, n, stepsize, eps, maximize, momentum, w.upd,
                            (void*)st));
      RV(launch_patch_operand(w.xp, w.dp, w.a16, E.kld, (long long)B * E.np, E.geom, E.ff, st,
                              E.split ? w.a16lo : nullptr));
    }
  }
  return launch_unpatchify_add(w.xp, w.dp, x_adv_out, B, E.geom, st);
}

// Debug / test access to internal buffers of the last call (same B / with_backward).
extern "C" int rvlm_vit_debug_buffer(rvlm_vit_handle h, void* workspace, int B, int with_backward,
                                     int which, int layer, void** ptr_out) {
  if (!h || !workspace || !ptr_out) return RVLM_ERR_INVAL;
  const Engine& E = *reinterpret_cast<Engine*>(h);
  Ws w = carve(E, reinterpret_cast<uint8_t*>(workspace), B, with_backward);
  const size_t MW = (size_t)B * E.T * E.W;
  const size_t s = with_backward ? (size_t)layer : 0;
  switch (which) {
    case 0: *ptr_out = w.tok32; break;
    case 1: *ptr_out = w.x32 + (with_backward ? (size_t)layer : (size_t)(layer & 1)) * MW; break;
    case 2: *ptr_out = w.xm32 + s * MW; break;
    case 3: *ptr_out = w.qkv16 + s * MW * 3; break;
    case 4: *ptr_out = w.o16 + s * MW; break;
    case 5: *ptr_out = w.h16 + s * (size_t)B * E.T * E.F; break;
    case 6: *ptr_out = w.dx32; break;
    case 7: *ptr_out = w.gp; break;
    case 8: *ptr_out = w.dp; break;
    case 9: *ptr_out = w.a16; break;
    case 10: *ptr_out = w.lse + s * (size_t)B * E.H * E.T; break;
    case 11: *ptr_out = w.dqkv16; break;
    case 12: *ptr_out = w.dtok16; break;
    default: return RVLM_ERR_INVAL;
  }
  return 0;
}
