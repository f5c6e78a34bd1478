"""The outer FARE / TeCoA training step on the CUDA engine
(`train_one_epoch`, train/adversarial_training_clip.py:289-366), one process per GPU.

    embedding_orig = model_orig(data)                     (:296-297, engine forward)
    data_adv       = pgd(...) | apgd(...)                 (:307-333, fused attack engine)
    loss           = l2(model(data_adv), embedding_orig) | ce(...)   (:349-360)
    loss.backward(); optimizer.step(); scheduler(step)    (:362-366)

The model's fp32 master parameters, their gradients and the AdamW moments live in flat
buffers, so the step needs ONE all-reduce (NCCL over NVLink; parallel.allreduce_mean_) and
ONE fused AdamW launch.  The attack runs on an fp16-forward / bf16-backward engine, the
outer forward/backward on a bf16 engine (weight-gradient GEMMs need a single operand
format); both read the same master parameters.
"""
from __future__ import annotations

import math

import torch

from . import _native as nat
from . import attacks, parallel
from ._native import c_float, c_int, c_int64
from .engine import FMT_BF16, LossSpec, VitEngine
from .model import ClipVisionModel, ComputeLossWrapper, VisionTransformer


def cosine_lr(base_lr, warmup_length, steps):
    """open_clip `training.scheduler.cosine_lr` semantics (train/adversarial_training_clip.py:18,211;
    an equivalent copy is vendored at CLIP_benchmark/clip_benchmark/metrics/linear_probe.py:13-32):
    linear warm-up then half-cosine; returns lr(step)."""
    def lr_at(step):
        if step < warmup_length:
            return base_lr * (step + 1) / warmup_length
        e, es = step - warmup_length, steps - warmup_length
        return 0.5 * (1 + math.cos(math.pi * e / es)) * base_lr
    return lr_at


class AdversarialTrainer:
    def __init__(self, model: ClipVisionModel, model_orig: ClipVisionModel, *, lr=1e-5, wd=1e-4,
                 betas=(0.9, 0.95), eps_adam=1e-8, steps=20000, warmup=1400, loss="l2",
                 inner_loss="l2", norm="linf", eps=4 / 255, stepsize=1 / 255, iterations=10,
                 attack="pgd", output_normalize=False, text_embedding=None, logit_scale=100.0):
        vt = model.model
        assert isinstance(vt, VisionTransformer)
        self.model, self.model_orig, self.vt = model, model_orig, vt
        self.dev = next(vt.parameters()).device
        # flat fp32 master parameters / grads / moments; parameters become views
        names, shapes = zip(*[(k, v.shape) for k, v in vt.state_dict().items()])
        sizes = [math.prod(s) for s in shapes]
        n = sum(sizes)
        self.flat_p = torch.empty(n, device=self.dev, dtype=torch.float32)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        self.grads, off = {}, 0
        params = dict(vt.named_parameters())
        for k, shp, sz in zip(names, shapes, sizes):
            view = self.flat_p[off:off + sz].view(shp)
            view.copy_(params[k].data)
            params[k].data = view
            self.grads[k] = self.flat_g[off:off + sz].view(shp)
            off += sz
        self.n = n
        self.lr_at = cosine_lr(lr, warmup, steps)
        self.wd, self.betas, self.eps_adam = wd, betas, eps_adam
        self.cfg = dict(loss=loss, inner_loss=inner_loss, norm=norm, eps=eps, stepsize=stepsize,
                        iterations=iterations, attack=attack, output_normalize=output_normalize)
        self.text, self.logit_scale = text_embedding, logit_scale
        self.step_total = 0
        mean, std = model.engine().cfg.mean, model.engine().cfg.std
        self.train_engine = VitEngine(vt.state_dict(), mean=tuple(mean), std=tuple(std),
                                      fmt_fwd=FMT_BF16, fmt_bwd=FMT_BF16, device=self.dev)

    def train_step(self, data, targets, global_batch=None):
        """One step of HOT LOOP A (train/adversarial_training_clip.py:289-366) on this rank's
        shard.  Returns dict(loss, emb_adv, data_adv)."""
        c = self.cfg
        data = data.to(self.dev, torch.float32)
        targets = targets.to(self.dev)
        B = data.shape[0]
        global_batch = global_batch or B * (torch.distributed.get_world_size()
                                           if torch.distributed.is_initialized() else 1)
        with torch.no_grad():
            emb_orig = self.model_orig(data, output_normalize=c["output_normalize"])   # :296-297
        inner = ComputeLossWrapper(emb_orig, self.text, "mean" if c["attack"] == "pgd" else "none",
                                   c["inner_loss"], self.logit_scale)                    # :300-304
        self.model.eval()                                                               # :305
        if c["attack"] == "pgd":
            d0 = torch.empty_like(data).uniform_(-c["eps"], c["eps"])                   # :318
            data_adv = attacks.pgd(self.model, inner, data, targets, c["norm"], c["eps"],
                                   c["iterations"], c["stepsize"], c["output_normalize"],
                                   perturbation=d0, mode="max")                        # :308-321
        else:
            data_adv = attacks.apgd_train(self.model, data, targets, c["norm"], c["eps"],
                                          n_iter=c["iterations"], loss_fn=inner)       # :324-333
        self.model.train()                                                              # :338
        spec = LossSpec(c["loss"], c["output_normalize"], emb_orig=emb_orig, text=self.text,
                        logit_scale=self.logit_scale, targets=targets)
        loss_b, emb_adv = self.train_engine.train_grads(data_adv, spec, self.grads,
                                                        1.0 / B)                        # :349-362
        parallel.allreduce_mean_(self.flat_g, B, global_batch)
        self.step_total += 1
        lr = self.lr_at(self.step_total - 1)                                            # :366 (scheduler)
        nat.call("rvlm_adamw_step", nat.ptr(self.flat_p), nat.ptr(self.flat_g),
                 nat.ptr(self.flat_m), nat.ptr(self.flat_v), c_int64(self.n), c_float(lr),
                 c_float(self.betas[0]), c_float(self.betas[1]), c_float(self.eps_adam),
                 c_float(self.wd), c_int(self.step_total), nat.stream_ptr())            # :363
        sd = self.vt.state_dict()
        self.vt.refresh_engine()                      # attack engine re-reads the new weights
        self.train_engine.update_weights(sd)
        return dict(loss=loss_b.mean(), emb_adv=emb_adv, data_adv=data_adv, lr=lr)
