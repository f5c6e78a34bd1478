"""Model-side call surface of the reference, backed by the fused CUDA engine.

  VisionTransformer     parameter container in open_clip `VisionTransformer` state-dict
                        layout -- what `model.visual` is in the
                        reference (train/adversarial_training_clip.py:183,188).
  ClipVisionModel       train/adversarial_training_clip.py:246-257 (same ctor / forward)
  ClassificationModel   CLIP_eval/clip_robustbench.py:50-69 (zero-shot logits)
  ComputeLossWrapper, compute_loss, l2, ce, compute_acc
                        train/adversarial_training_clip.py:260-274,488-528

`ClipVisionModel.forward` runs the engine; when the input requires grad it returns a
tensor whose backward is the engine's input-gradient pass, so the reference's generic
`torch.autograd.grad(loss, perturbation)` (train/pgd_train.py:38) works with any loss
callable.  The tiny [B, D] loss arithmetic of `l2` / `ce` stays in torch here; the
fused attack paths (attacks.py) evaluate them inside the CUDA head kernel instead.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import CLIP_MEAN, CLIP_STD, FMT_BF16, FMT_F16, PRECISIONS, VitEngine

ARCHS = {  # `--clip_model_name` values (train/adversarial_training_clip.py:32,175-178)
    "ViT-L-14": dict(image=224, patch=14, width=1024, layers=24, heads=16, mlp=4096, out_dim=768),
    "ViT-B-32": dict(image=224, patch=32, width=768, layers=12, heads=12, mlp=3072, out_dim=512),
    "tiny-14": dict(image=224, patch=14, width=256, layers=2, heads=4, mlp=1024, out_dim=128),
    "tiny-32": dict(image=224, patch=32, width=128, layers=2, heads=2, mlp=512, out_dim=64),
}


class _Attn(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * w, w))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * w))
        self.out_proj = nn.Linear(w, w)


class _Mlp(nn.Module):
    def __init__(self, w, m):
        super().__init__()
        self.c_fc = nn.Linear(w, m)
        self.c_proj = nn.Linear(m, w)


class _Block(nn.Module):
    def __init__(self, w, m):
        super().__init__()
        self.ln_1 = nn.LayerNorm(w)
        self.attn = _Attn(w)
        self.ln_2 = nn.LayerNorm(w)
        self.mlp = _Mlp(w, m)


class _Transformer(nn.Module):
    def __init__(self, w, layers, m):
        super().__init__()
        self.resblocks = nn.ModuleList([_Block(w, m) for _ in range(layers)])


class VisionTransformer(nn.Module):
    """Holds the fp32 master parameters; all arithmetic happens in the CUDA engine."""

    def __init__(self, arch="ViT-L-14", seed=None, fmt_fwd=FMT_F16, fmt_bwd=FMT_F16, flags=0,
                 precision=None):
        super().__init__()
        if precision is not None:                    # a named mode of engine.PRECISIONS
            fmt_fwd, fmt_bwd = PRECISIONS[precision]
        a = ARCHS[arch] if isinstance(arch, str) else dict(arch)
        self.arch = a
        w, m = a["width"], a["mlp"]
        tokens = (a["image"] // a["patch"]) ** 2 + 1
        self.conv1 = nn.Conv2d(3, w, a["patch"], a["patch"], bias=False)
        self.class_embedding = nn.Parameter(torch.empty(w))
        self.positional_embedding = nn.Parameter(torch.empty(tokens, w))
        self.ln_pre = nn.LayerNorm(w)
        self.transformer = _Transformer(w, a["layers"], m)
        self.ln_post = nn.LayerNorm(w)
        self.proj = nn.Parameter(torch.empty(w, a["out_dim"]))
        self.output_tokens = False
        self._engine = None
        if os.environ.get("RVLM_ATTN_SIMT", "0") == "1":   # debug: CUDA-core attention checker
            flags |= 1
        self._engine_opts = dict(fmt_fwd=fmt_fwd, fmt_bwd=fmt_bwd, flags=flags)
        self.reset_parameters(seed)

    def reset_parameters(self, seed=None):
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        w = self.arch["width"]
        with torch.no_grad():
            for n, p in self.named_parameters():
                if n.endswith("bias"):
                    p.copy_(0.02 * torch.randn(p.shape, generator=g))
                elif "ln_" in n:
                    p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
                elif p.dim() >= 2:
                    fan_in = p[0].numel()
                    p.copy_(torch.randn(p.shape, generator=g) * fan_in ** -0.5)
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * w ** -0.5)
        self._engine = None

    # -- engine management -------------------------------------------------------------
    def engine(self, mean=CLIP_MEAN, std=CLIP_STD) -> VitEngine:
        key = (tuple(float(v) for v in mean), tuple(float(v) for v in std))
        if self._engine is None or self._engine[0] != key:
            sd = {k: v for k, v in self.state_dict().items()}
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("robustvlm_b200 runs on CUDA only: move the model with .cuda()")
            self._engine = (key, VitEngine(sd, mean=mean, std=std, device=dev,
                                           **self._engine_opts))
        return self._engine[1]

    def refresh_engine(self):
        """Re-read the parameters (after an optimiser step or load_state_dict)."""
        if self._engine is not None:
            self._engine[1].update_weights(self.state_dict())

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.refresh_engine()
        return r

    def forward(self, x_normalized):  # noqa: D401
        raise RuntimeError("call through ClipVisionModel (normalisation is fused into the engine)")


def _norm_constants(normalize):
    if normalize is None:
        return CLIP_MEAN, CLIP_STD
    mean, std = getattr(normalize, "mean", None), getattr(normalize, "std", None)
    if mean is None or std is None:
        raise TypeError("normalize must expose .mean / .std (torchvision.transforms.Normalize)")
    to = lambda v: tuple(float(t) for t in torch.as_tensor(v).flatten())  # noqa: E731
    return to(mean), to(std)


class _EngineVJP(torch.autograd.Function):
    """The activations of the forward stay in the engine's workspace; `stamp` identifies them, and
    the backward refuses to run if any other engine call has touched the workspace since (a second
    forward before the backward, another attack, a larger batch): recompute instead of returning a
    gradient of stale activations."""

    @staticmethod
    def forward(ctx, vision, eng, output_normalize):
        ctx.eng, ctx.norm, ctx.B = eng, output_normalize, vision.shape[0]
        out = eng.forward(vision.detach(), output_normalize=output_normalize,
                          keep_for_backward=True)
        ctx.stamp = eng.kept_stamp()
        ctx.save_for_backward(vision)
        return out

    @staticmethod
    def backward(ctx, d_emb):
        eng = ctx.eng
        if eng.kept_stamp() != ctx.stamp:
            # the workspace was reused in between: redo this node's forward, then differentiate
            (vision,) = ctx.saved_tensors
            eng.forward(vision.detach(), output_normalize=ctx.norm, keep_for_backward=True)
        return eng.backward_input(ctx.B, d_emb, ctx.norm), None, None


class ClipVisionModel(nn.Module):
    """train/adversarial_training_clip.py:246-257."""

    def __init__(self, model, args=None, normalize=None):
        super().__init__()
        self.model = model
        self.args = args
        self.normalize = normalize

    def engine(self) -> VitEngine:
        if not isinstance(self.model, VisionTransformer):
            raise TypeError("fused path needs a robustvlm_b200.model.VisionTransformer")
        mean, std = _norm_constants(self.normalize)
        return self.model.engine(mean, std)

    def forward(self, vision, output_normalize):
        if isinstance(self.model, VisionTransformer):
            eng = self.engine()
            if torch.is_grad_enabled() and vision.requires_grad:
                return _EngineVJP.apply(vision.float().contiguous(), eng, bool(output_normalize))
            return eng.forward(vision.detach().float(), output_normalize=bool(output_normalize))
        # foreign tower (e.g. a real open_clip module): the reference's own arithmetic
        embedding = self.model(self.normalize(vision))
        if output_normalize:
            embedding = F.normalize(embedding, dim=-1)
        return embedding


class ClassificationModel(nn.Module):
    """CLIP_eval/clip_robustbench.py:50-69: logits = normalize(phi(resizer(x))) @ T * logit_scale.

    `model` is the vision tower (ours, or a ClipVisionModel around it).  `logit_scale` as in the
    reference: True multiplies by `model.logit_scale.exp()` when the model carries the CLIP
    parameter (a full CLIP module) and by 100 -- its value for every OpenAI checkpoint,
    SURVEY.md §8a a11 -- when only the tower is given; False leaves the cosine logits; a number is
    used as is.  `resizer` is applied to the images before the tower (:56,62); with a resizer the
    fused attack path cannot be used (the gradient must flow through it) and autograd takes over."""

    def __init__(self, model, text_embedding, args=None, input_normalize=None, resizer=None,
                 logit_scale=True):
        super().__init__()
        self.visual = model if isinstance(model, ClipVisionModel) else ClipVisionModel(
            model, args, input_normalize)
        self.args = args
        self.resizer = resizer
        self.register_buffer("text_embedding", text_embedding)  # [D, n_classes]
        if logit_scale is True:
            ls = getattr(model, "logit_scale", None)
            self.logit_scale_value = float(ls.exp()) if torch.is_tensor(ls) else 100.0
        else:
            self.logit_scale_value = 1.0 if logit_scale is False else float(logit_scale)

    def forward(self, vision, output_normalize=True):
        assert output_normalize                                               # :60
        if self.resizer is not None:
            vision = self.resizer(vision)
        emb = self.visual(vision, output_normalize=True)
        return (emb @ self.text_embedding) * self.logit_scale_value


# ------------------------------------------------------------------------------------
# losses (train/adversarial_training_clip.py:495-528)
# ------------------------------------------------------------------------------------
def l2(out, targets, reduction="none"):
    assert out.shape == targets.shape, f"{out.shape} != {targets.shape}"
    assert out.shape[0] > 1
    sq = F.mse_loss(out, targets, reduction="none").sum(dim=1)
    return sq.mean() if reduction == "mean" else sq


def ce(out, targets, reduction="mean"):
    assert out.shape[0] == targets.shape[0], (out.shape, targets.shape)
    assert out.shape[0] > 1
    return F.cross_entropy(out, targets, reduction=reduction)


def compute_loss(loss_str, embedding, targets, embedding_orig, logit_scale,
                 embedding_text_labels_norm=None, reduction="mean"):
    if loss_str == "l2":
        return l2(out=embedding, targets=embedding_orig, reduction=reduction)
    if loss_str == "ce":
        return ce(out=embedding @ (logit_scale * embedding_text_labels_norm), targets=targets,
                  reduction=reduction)
    raise ValueError(f"loss {loss_str} not supported")


class ComputeLossWrapper:
    def __init__(self, embedding_orig, embedding_text_labels_norm, reduction="mean", loss=None,
                 logit_scale=100.0):
        self.embedding_orig = embedding_orig
        self.embedding_text_labels_norm = embedding_text_labels_norm
        self.reduction = reduction
        self.loss_str = loss
        self.logit_scale = logit_scale

    def __call__(self, embedding, targets):
        return compute_loss(loss_str=self.loss_str, embedding=embedding, targets=targets,
                            embedding_orig=self.embedding_orig, logit_scale=self.logit_scale,
                            embedding_text_labels_norm=self.embedding_text_labels_norm,
                            reduction=self.reduction)


@torch.no_grad()
def compute_acc(logits, targets):
    preds = logits.max(dim=1)[1].detach()
    return (preds.eq(targets).sum() / targets.shape[0]).item() * 100
