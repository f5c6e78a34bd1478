#!/usr/bin/env python
"""CPU emulation of the engine's rounding sites: which 16-bit roundings (forward operands,
attention-core operands, backward operands) cost how much of the FARE gradient's sign agreement
with the fp32 reference.  Design tool for the split-operand ("fp32-class") forward; every site
where the CUDA engine hands a 16-bit tensor to a tensor-core MMA is rounded here in the same
place -- to fp16, bf16, a mantissa-only 11-bit rounding ("m11": fp16 precision without its
range, what the unit-rescaled fp16 backward sees) or a hi + lo fp16 pair ("split": the 3-MMA
product hi*hi + hi*lo + lo*hi).

    python tools/precision_probe.py [--model ViT-B-32] [--batch 4] [--same-weights 1]

--same-weights 1 is the FARE situation: model_orig == model, so
e_adv - e_orig is a cancellation site; 0 uses two different random towers.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit as ovit  # noqa: E402


def _m11(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _split(x):
    hi = x.half().float()
    return hi + (x - hi).half().float()


ROUND = {"f32": lambda x: x, "f16": lambda x: x.half().float(), "bf16": lambda x: x.bfloat16().float(),
         "m11": _m11, "split": _split}


class Lin(torch.autograd.Function):
    """y = rf(a) rf(W)^T ; da = rb(dy) rb(W)"""

    @staticmethod
    def forward(ctx, a, W, rf, rb):
        ctx.save_for_backward(W)
        ctx.rb = rb
        rfa, rfw = rf.split("/") if "/" in rf else (rf, rf)    # "split/f16": activations / weights
        return ROUND[rfa](a) @ ROUND[rfw](W).t()

    @staticmethod
    def backward(ctx, dy):
        (W,) = ctx.saved_tensors
        return ROUND[ctx.rb](dy) @ ROUND[ctx.rb](W), None, None, None


class GradRound(torch.autograd.Function):
    """identity whose gradient is rounded (a 16-bit gradient hand-over between kernels)"""

    @staticmethod
    def forward(ctx, x, rb):
        ctx.rb = rb
        return x

    @staticmethod
    def backward(ctx, g):
        return ROUND[ctx.rb](g), None


class Attn(torch.autograd.Function):
    """softmax(q k^T / 8) v per head with the engine's operand roundings: forward q, k, v, P in
    `ra` (P un-normalised, as the kernel stores it), backward everything in `rb` with the
    probabilities recomputed from the backward's copy of q | k | v."""

    @staticmethod
    def forward(ctx, q, k, v, ra, rp, rb):
        r = ROUND[ra]
        s = (r(q) @ r(k).transpose(-1, -2)) * 0.125
        e = torch.exp(s - s.max(-1, keepdim=True)[0])
        o = (ROUND[rp](e) @ r(v)) / e.sum(-1, keepdim=True)
        ctx.save_for_backward(q, k, v, o)
        ctx.rb = rb
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o = ctx.saved_tensors
        r = ROUND[ctx.rb]
        q, k, v, do = r(q), r(k), r(v), r(do)
        p = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1)
        dv = r(p).transpose(-1, -2) @ do
        dp = do @ v.transpose(-1, -2)
        D = (o * do).sum(-1, keepdim=True)
        ds = r(p * (dp - D) * 0.125)
        return ds @ k, ds.transpose(-1, -2) @ q, dv, None, None, None


def forward(m, xn, lin, att, pm, rb):
    cfg = m.cfg
    B = xn.shape[0]
    p = cfg.patch
    a = xn.unfold(2, p, p).unfold(3, p, p).permute(0, 2, 3, 1, 4, 5).reshape(B, -1, 3 * p * p)
    x = Lin.apply(a, m.conv1.weight.reshape(cfg.width, -1), lin, rb)
    x = torch.cat([m.class_embedding.expand(B, 1, -1), x], 1) + m.positional_embedding
    x = m.ln_pre(x)
    H = cfg.heads
    ra_o = ROUND[att]
    for blk in m.transformer.resblocks:
        y = GradRound.apply(blk.ln_1(x), rb)
        qkv = GradRound.apply(Lin.apply(y, blk.attn.in_proj_weight, lin, rb) + blk.attn.in_proj_bias, rb)
        q, k, v = (t.view(B, -1, H, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
        o = Attn.apply(q, k, v, att, pm, rb).transpose(1, 2).reshape(B, -1, cfg.width)
        o = o + (ra_o(o) - o).detach()          # o16 is stored in the attention's output format
        x = x + Lin.apply(o, blk.attn.out_proj.weight, lin, rb) + blk.attn.out_proj.bias
        z = GradRound.apply(blk.ln_2(x), rb)
        h = GradRound.apply(Lin.apply(z, blk.mlp.c_fc.weight, lin, rb) + blk.mlp.c_fc.bias, rb)
        act = h * torch.sigmoid(1.702 * h)
        x = x + Lin.apply(act, blk.mlp.c_proj.weight, lin, rb) + blk.mlp.c_proj.bias
    return m.ln_post(x[:, 0]) @ m.proj


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ViT-B-32")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--same-weights", type=int, default=1)
    ap.add_argument("--e-orig-engine", type=int, default=0,
                    help="1: e_orig computed with the same rounded forward (caller uses our model_orig)")
    args = ap.parse_args()
    eps = 4 / 255
    m = ovit.build(args.model, seed=1)
    mo = m if args.same_weights else ovit.build(args.model, seed=2)
    for p_ in list(m.parameters()) + list(mo.parameters()):
        p_.requires_grad_(False)
    norm = ovit.Normalize()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(args.batch, 3, 224, 224, generator=g)
    d0 = (torch.rand(args.batch, 3, 224, 224, generator=g) * 2 - 1) * eps
    with torch.no_grad():
        e_orig = mo(norm(x))

    def grad(lin, att, pm, rb):
        d = d0.clone().requires_grad_(True)
        e = forward(m, norm(x + d), lin, att, pm, rb)
        eo = e_orig
        if args.e_orig_engine:
            with torch.no_grad():
                eo = forward(mo, norm(x), lin, att, pm, "f32")
        loss = F.mse_loss(e, eo, reduction="none").sum(1).mean()
        return e.detach(), torch.autograd.grad(loss, d)[0]

    # reference: the oracle itself in fp64
    m64, d64 = ovit.build(args.model, seed=1, dtype=torch.float64), d0.double().requires_grad_(True)
    e64 = m64(ovit.Normalize().double()(x.double() + d64))
    g_ref = torch.autograd.grad(F.mse_loss(e64, e_orig.double(), reduction="none").sum(1).mean(), d64)[0].float()
    e_ref = e64.detach().float()
    big = g_ref.abs() > 0.05 * g_ref.abs().median()
    print(f"{args.model} B={args.batch} same_weights={args.same_weights}: |e_adv - e_orig| / |e_orig| = "
          f"{float((e_ref - e_orig).norm() / e_orig.norm()):.3e}")
    print(f"{'linear':>7} {'attn':>6} {'P':>6} {'bwd':>5} | emb rel err | grad rel err | sign mismatch | "
          f"mismatch |g|>0.05 med")
    rows = [("f32", "f32", "f32", "f32"), ("f16", "f16", "f16", "f32"), ("f16", "f16", "f16", "bf16"),
            ("f16", "f16", "f16", "m11"), ("split", "f16", "f16", "f32"), ("split", "split", "f16", "f32"),
            ("split", "split", "split", "f32"), ("split", "split", "split", "bf16"),
            ("split", "split", "split", "m11"), ("split", "split", "f16", "m11"),
            ("split", "f16", "f16", "m11"), ("split/f16", "split", "split", "f32"),
            ("f16/split", "split", "split", "f32"), ("bf16", "bf16", "bf16", "bf16")]
    for lin, att, pm, rb in rows:
        e, gr = grad(lin, att, pm, rb)
        mis = gr.sign() != g_ref.sign()
        print(f"{lin:>9} {att:>6} {pm:>6} {rb:>5} | {float((e - e_ref).norm() / e_ref.norm()):11.3e} | "
              f"{float((gr - g_ref).norm() / g_ref.norm()):12.3e} | {float(mis.float().mean()) * 100:12.4f}% | "
              f"{float((mis & big).float().sum() / big.float().sum()) * 100:.4f}%", flush=True)


if __name__ == "__main__":
    main()
