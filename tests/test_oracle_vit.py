"""The oracle's restated open_clip ViT vs an independent implementation
(HF transformers CLIPVisionModelWithProjection, quick_gelu).  open_clip itself is an
un-vendored dependency of the reference and is absent offline."""
import pytest
import torch

from oracle import vit

transformers = pytest.importorskip("transformers")


# (ViT-L-14 / ViT-B-32: the full-size towers of BASELINE.json's configs, one image on the CPU)
@pytest.mark.parametrize("name", ["tiny-32", "tiny-14", "ViT-B-32", "ViT-L-14"])
def test_restated_vit_matches_hf(name):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = vit.CONFIGS[name]
    m = vit.build(name, seed=5)
    hf_cfg = CLIPVisionConfig(hidden_size=cfg.width, intermediate_size=cfg.mlp,
                              projection_dim=cfg.out_dim, num_hidden_layers=cfg.layers,
                              num_attention_heads=cfg.heads, image_size=cfg.image_size,
                              patch_size=cfg.patch, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                              attention_dropout=0.0)
    hf = CLIPVisionModelWithProjection(hf_cfg).eval()
    missing, unexpected = hf.load_state_dict(vit.to_hf_state_dict(m), strict=False)
    assert not [k for k in missing if "position_ids" not in k], missing
    assert not unexpected, unexpected
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2 if name.startswith("tiny") else 1, 3, cfg.image_size, cfg.image_size,
                    generator=g)
    with torch.no_grad():
        a = m(x)
        b = hf(pixel_values=x).image_embeds
    assert float((a - b).norm() / b.norm()) < 1e-5
    # the input gradient too (what the attack path differentiates): the same scalar of the embedding
    w = torch.randn(cfg.out_dim, generator=g)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    (ga,) = torch.autograd.grad((m(xa) * w).sum(), xa)
    (gb,) = torch.autograd.grad((hf(pixel_values=xb).image_embeds * w).sum(), xb)
    assert float((ga - gb).norm() / gb.norm()) < 1e-4


def test_param_counts_match_reference_checkpoints():
    # README.md:44 "approx. 1.1 GB" == 303 966 208 fp32 parameters
    def count(c):
        w, t = c.width, c.tokens
        per_layer = 4 * w + 3 * w * w + 3 * w + w * w + w + 2 * w * c.mlp + c.mlp + w
        return 3 * c.patch ** 2 * w + w + t * w + 4 * w + c.layers * per_layer + w * c.out_dim
    assert count(vit.CONFIGS["ViT-L-14"]) == 303_966_208
    assert count(vit.CONFIGS["ViT-B-32"]) == 87_849_216
