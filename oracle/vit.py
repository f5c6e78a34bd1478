"""CPU oracle of the CLIP vision tower (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates the semantics of open_clip 2.19.0 `VisionTransformer` (external pip
dependency of the reference, requirements.txt:91; used at
train/adversarial_training_clip.py:95-97,183,188 and CLIP_eval/eval_utils.py:32-34)
from plain `torch.nn` building blocks, so PyTorch supplies the arithmetic:

  Normalize -> conv1 (k = s = patch, no bias) -> [cls ; patches] + pos -> ln_pre
  -> L x { x += MHA(ln_1(x)) ; x += c_proj(QuickGELU(c_fc(ln_2(x)))) }
  -> ln_post(x[:, 0]) @ proj

Parameter names follow the open_clip state-dict layout
because released FARE/TeCoA checkpoints are `visual.state_dict()` dumps
(train/adversarial_training_clip.py:239,470; README.md:57-67).
"""
from __future__ import annotations

import dataclasses
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

# train/adversarial_training_clip.py:116 (OpenAI CLIP normalisation constants)
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclasses.dataclass(frozen=True)
class VitConfig:
    image_size: int = 224
    patch: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    out_dim: int = 768

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1


# `--clip_model_name` values the reference trains (train/adversarial_training_clip.py:32,175-178)
CONFIGS = {
    "ViT-L-14": VitConfig(224, 14, 1024, 24, 16, 4096, 768),
    "ViT-B-32": VitConfig(224, 32, 768, 12, 12, 3072, 512),
    # small shapes for fast tests (same head_dim = 64 as the real models)
    "tiny-14": VitConfig(224, 14, 256, 2, 4, 1024, 128),
    "tiny-32": VitConfig(224, 32, 128, 2, 2, 512, 64),
}


class QuickGELU(nn.Module):
    # OpenAI-pretrained CLIP uses x * sigmoid(1.702 x)
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class Block(nn.Module):
    def __init__(self, width: int, heads: int, mlp: int):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.MultiheadAttention(width, heads)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(width, mlp)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(mlp, width)),
        ]))

    def forward(self, x):  # x: [T, B, W] (sequence-first, like open_clip 2.19)
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False)[0]
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, mlp):
        super().__init__()
        self.resblocks = nn.ModuleList([Block(width, heads, mlp) for _ in range(layers)])

    def forward(self, x):
        for blk in self.resblocks:
            x = blk(x)
        return x


class VisionTransformer(nn.Module):
    """Input: already-normalised image batch [B,3,H,W]; output: [B,out_dim]."""

    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.cfg = cfg
        w = cfg.width
        scale = w ** -0.5
        self.conv1 = nn.Conv2d(3, w, kernel_size=cfg.patch, stride=cfg.patch, bias=False)
        self.class_embedding = nn.Parameter(scale * torch.randn(w))
        self.positional_embedding = nn.Parameter(scale * torch.randn(cfg.tokens, w))
        self.ln_pre = nn.LayerNorm(w)
        self.transformer = Transformer(w, cfg.layers, cfg.heads, cfg.mlp)
        self.ln_post = nn.LayerNorm(w)
        self.proj = nn.Parameter(scale * torch.randn(w, cfg.out_dim))
        self.output_tokens = False

    def forward(self, x):
        x = self.conv1(x)                                   # [B, W, g, g]
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)  # [B, N, W]
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
        pooled = self.ln_post(x[:, 0])
        return pooled @ self.proj


class Normalize(nn.Module):
    """torchvision.transforms.Normalize equivalent (kept inside the model so eps is
    in pixel space; train/adversarial_training_clip.py:105-107,116)."""

    def __init__(self, mean=CLIP_MEAN, std=CLIP_STD):
        super().__init__()
        self.register_buffer("mean", torch.tensor(mean).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor(std).view(1, 3, 1, 1))

    def forward(self, x):
        return (x - self.mean.to(x.dtype)) / self.std.to(x.dtype)


class ClipVisionModel(nn.Module):
    """train/adversarial_training_clip.py:246-257."""

    def __init__(self, model, args=None, normalize=None):
        super().__init__()
        self.model = model
        self.args = args
        self.normalize = normalize if normalize is not None else Normalize()

    def forward(self, vision, output_normalize):
        emb = self.model(self.normalize(vision))
        if output_normalize:
            emb = F.normalize(emb, dim=-1)
        return emb


class ClassificationModel(nn.Module):
    """Zero-shot classifier used by the evaluation APGD
    (CLIP_eval/clip_robustbench.py:50-69): logits = normalize(phi(x)) @ T * logit_scale."""

    def __init__(self, visual, text_embedding, logit_scale=100.0, normalize=None):
        super().__init__()
        self.visual = visual
        self.normalize = normalize if normalize is not None else Normalize()
        self.register_buffer("text_embedding", text_embedding)  # [D, n_classes]
        self.logit_scale = logit_scale

    def forward(self, x):
        e = F.normalize(self.visual(self.normalize(x)), dim=-1)
        return (e @ self.text_embedding) * self.logit_scale


def build(name_or_cfg, seed: int = 0, dtype=torch.float32) -> VisionTransformer:
    """Random-init tower (no checkpoints offline).  Deterministic in `seed`.

    LayerNorm affine parameters and all biases are perturbed away from their
    (1, 0) defaults so that parity tests exercise them."""
    cfg = CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
    g = torch.Generator().manual_seed(seed)
    rng_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = VisionTransformer(cfg)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n.endswith("bias"):
                    p.copy_(0.02 * torch.randn(p.shape, generator=g))
                elif "ln_" in n and n.endswith("weight"):
                    p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
    finally:
        torch.random.set_rng_state(rng_state)
    return m.to(dtype).eval()


def to_hf_state_dict(m: VisionTransformer) -> dict:
    """Weight mapping onto HF `CLIPVisionModelWithProjection`, used
    only by the independent cross-check in tests/test_oracle_vit.py."""
    sd = m.state_dict()
    w = m.cfg.width
    out = {
        "vision_model.embeddings.class_embedding": sd["class_embedding"],
        "vision_model.embeddings.position_embedding.weight": sd["positional_embedding"],
        "vision_model.embeddings.patch_embedding.weight": sd["conv1.weight"],
        "vision_model.pre_layrnorm.weight": sd["ln_pre.weight"],
        "vision_model.pre_layrnorm.bias": sd["ln_pre.bias"],
        "vision_model.post_layernorm.weight": sd["ln_post.weight"],
        "vision_model.post_layernorm.bias": sd["ln_post.bias"],
        "visual_projection.weight": sd["proj"].t().contiguous(),
    }
    for i in range(m.cfg.layers):
        p = f"transformer.resblocks.{i}."
        q = f"vision_model.encoder.layers.{i}."
        ipw, ipb = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            out[q + f"self_attn.{nm}.weight"] = ipw[j * w:(j + 1) * w]
            out[q + f"self_attn.{nm}.bias"] = ipb[j * w:(j + 1) * w]
        out[q + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
        out[q + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
        out[q + "layer_norm1.weight"] = sd[p + "ln_1.weight"]
        out[q + "layer_norm1.bias"] = sd[p + "ln_1.bias"]
        out[q + "layer_norm2.weight"] = sd[p + "ln_2.weight"]
        out[q + "layer_norm2.bias"] = sd[p + "ln_2.bias"]
        out[q + "mlp.fc1.weight"] = sd[p + "mlp.c_fc.weight"]
        out[q + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.bias"]
        out[q + "mlp.fc2.weight"] = sd[p + "mlp.c_proj.weight"]
        out[q + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.bias"]
    return out
