"""The outer FARE / TeCoA training step on the CUDA engine
(`train_one_epoch`, train/adversarial_training_clip.py:289-366), one process per GPU.

    embedding_orig = model_orig(data)                     (:296-297, engine forward)
    data_adv       = pgd(...) | apgd(...)                 (:307-333, fused attack engine)
    loss           = l2(model(data_adv), embedding_orig) | ce(...)   (:349-360)
    loss.backward(); optimizer.step(); scheduler(step)    (:362-366)

The model's fp32 master parameters, their gradients and the AdamW moments live in flat
buffers, so the step needs ONE all-reduce (NCCL over NVLink; parallel.allreduce_mean_) and
ONE fused AdamW launch.  The attack runs on the fp16 / fp16 engine, the outer forward / backward
on an fp16-forward / bf16-gradient engine (the weight-gradient GEMMs contract a forward
activation with a gradient in ONE format: recast copies of the GEMM inputs are kept for them);
both read the same master parameters.
"""
from __future__ import annotations

import math

import torch

from . import _native as nat
from . import attacks, checkpoint, parallel
from ._native import c_double, c_int, c_int64
from .engine import FMT_BF16, FMT_F16, LossSpec, VitEngine
from .model import ClipVisionModel, ComputeLossWrapper, VisionTransformer, compute_acc


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
    """`main()` + `train_one_epoch()` of train/adversarial_training_clip.py (:190-245, 276-486)
    for one rank: optimizer (`--opt adamw|sgd`, torch defaults as the reference constructs them,
    :196-204), the `cosine_lr` schedule with the reference's call order (the FIRST step runs at
    the constructor lr -- the scheduler is first called after it, :366), the attack
    (`--attack pgd|apgd|none`), `--clean_weight` / `--loss_clean` / `--trades`, the logged
    metrics (:368-387), the in-training APGD evaluation (:390-425) and checkpoints (:238-239,
    467-479; checkpoint.py)."""

    def __init__(self, model: ClipVisionModel, model_orig: ClipVisionModel, *, lr=1e-5, wd=1e-4,
                 opt="adamw", betas=(0.9, 0.999), eps_adam=1e-8, momentum_sgd=0.9, steps=20000,
                 warmup=1400, loss="l2", loss_clean="none", clean_weight=0.0, trades=False,
                 inner_loss="l2", norm="linf", eps=4 / 255, stepsize=1 / 255, iterations=10,
                 attack="pgd", output_normalize=False, text_embedding=None, logit_scale=100.0,
                 start_step=0, train_precision="mixed"):
        vt = model.model
        assert isinstance(vt, VisionTransformer)
        if opt not in ("adamw", "sgd"):
            raise ValueError(f"Optimizer {opt} not supported.")                         # :206
        if attack not in ("pgd", "apgd", "none"):
            raise ValueError(f"attack {attack} not supported")
        self.model, self.model_orig, self.vt = model, model_orig, vt
        self.dev = next(vt.parameters()).device
        # flat fp32 master parameters / grads / moments; parameters become views.  Order =
        # model.parameters() (what the reference hands to the optimizer, :194)
        self.layout = checkpoint.param_layout(vt)
        n = sum(sz for *_, sz in self.layout)
        self.flat_p = torch.empty(n, device=self.dev, dtype=torch.float32)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p) if opt == "adamw" else None
        self.flat_g2 = torch.zeros_like(self.flat_p) if clean_weight > 0 else None
        self.grads, self.grads2 = {}, {}
        params = dict(vt.named_parameters())
        for k, shp, off, sz in self.layout:
            view = self.flat_p[off:off + sz].view(shp)
            view.copy_(params[k].data)
            params[k].data = view
            self.grads[k] = self.flat_g[off:off + sz].view(shp)
            if self.flat_g2 is not None:
                self.grads2[k] = self.flat_g2[off:off + sz].view(shp)
        self.n = n
        self.opt, self.base_lr, self.lr = opt, lr, lr            # lr: what the NEXT step uses
        self.lr_at = cosine_lr(lr, warmup, steps)
        self.wd, self.betas, self.eps_adam, self.momentum_sgd = wd, betas, eps_adam, momentum_sgd
        self.steps = steps
        self.cfg = dict(loss=loss, loss_clean=loss_clean, clean_weight=clean_weight, trades=trades,
                        inner_loss=inner_loss, norm=norm, eps=eps, stepsize=stepsize,
                        iterations=iterations, attack=attack, output_normalize=output_normalize)
        self.text, self.logit_scale = text_embedding, logit_scale
        self.step_total = start_step                                                    # :219
        self.opt_steps = 0                       # optimizer.step() calls (AdamW bias correction)
        mean, std = model.engine().cfg.mean, model.engine().cfg.std
        # outer forward / backward: "mixed" = fp16 forward operands (embedding within 1e-3 of fp32,
        # as the attack path) + bf16 gradient operands, with gradient-format copies of the GEMM
        # inputs kept for the weight-gradient GEMMs; "bf16" = bf16 everywhere (round 1)
        if train_precision not in ("mixed", "bf16"):
            raise ValueError(f"train_precision {train_precision} not supported")
        self.train_engine = VitEngine(vt.state_dict(), mean=tuple(mean), std=tuple(std),
                                      fmt_fwd=FMT_F16 if train_precision == "mixed" else FMT_BF16,
                                      fmt_bwd=FMT_BF16, device=self.dev)
        self._bucketed = None            # parallel.BucketedAllReduce, built on first use

    # -- optimizer state / checkpoints (reference formats, checkpoint.py) ---------------
    def optimizer_state_dict(self):
        if self.opt == "adamw":
            return checkpoint.adamw_state_dict(self.layout, self.flat_m, self.flat_v,
                                               self.opt_steps, lr=self.lr, betas=self.betas,
                                               eps=self.eps_adam, weight_decay=self.wd)
        return checkpoint.sgd_state_dict(self.layout, self.flat_m, self.opt_steps, lr=self.lr,
                                         momentum=self.momentum_sgd, weight_decay=self.wd)

    def load_optimizer_state_dict(self, sd):
        """`optimizer.load_state_dict(torch.load(args.optimizer_state))` (:207-208)"""
        self.opt_steps, group = checkpoint.load_optimizer_state_dict(sd, self.layout, self.flat_m,
                                                                     self.flat_v)
        self.lr = group["lr"]
        self.wd = group.get("weight_decay", self.wd)
        if self.opt == "adamw":
            self.betas, self.eps_adam = tuple(group.get("betas", self.betas)), group.get(
                "eps", self.eps_adam)
        else:
            self.momentum_sgd = group.get("momentum", self.momentum_sgd)

    def load_model_state_dict(self, sd):
        """`model.visual.load_state_dict(...)` (:93-103): copies into the flat master buffer
        (the parameters are views of it) and refreshes both engines."""
        self.vt.load_state_dict(checkpoint.load_visual_state_dict(sd))
        self._refresh()

    def save(self, output_dir, stem):
        """`final.pt` / `final_opt.pt` (:238-239).  Rank 0 writes; every rank returns after it has."""
        if checkpoint.is_writer_rank():
            checkpoint.save_pair(self.vt.state_dict(), self.optimizer_state_dict(),
                                 f"{output_dir}/checkpoints", stem)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier()

    def _refresh(self):
        sd = self.vt.state_dict()
        self.vt.refresh_engine()                      # attack engine re-reads the new weights
        self.train_engine.update_weights(sd)

    # -- one step ------------------------------------------------------------------------
    def _optimizer_step(self):
        self.opt_steps += 1
        if self.opt == "adamw":
            nat.call("rvlm_adamw_step", nat.ptr(self.flat_p), nat.ptr(self.flat_g),
                     nat.ptr(self.flat_m), nat.ptr(self.flat_v), c_int64(self.n),
                     c_double(self.lr), c_double(self.betas[0]), c_double(self.betas[1]),
                     c_double(self.eps_adam), c_double(self.wd), c_int(self.opt_steps),
                     nat.stream_ptr())                                                  # :363
        else:
            nat.call("rvlm_sgd_step", nat.ptr(self.flat_p), nat.ptr(self.flat_g),
                     nat.ptr(self.flat_m), c_int64(self.n), c_double(self.lr),
                     c_double(self.momentum_sgd), c_double(self.wd), c_int(self.opt_steps),
                     nat.stream_ptr())

    def train_step(self, data, targets, global_batch=None, output_dir=None, save_checkpoints=True):
        """One iteration of HOT LOOP A (train/adversarial_training_clip.py:289-366, logging
        :368-388, saving :467-479) on this rank's shard.  `targets` may be None (non-classification
        data, :290).  Returns the reference's log fields."""
        c = self.cfg
        data = data.to(self.dev, torch.float32)
        is_classification = isinstance(targets, torch.Tensor)                           # :290
        targets = targets.to(self.dev) if is_classification else None
        B = data.shape[0]
        global_batch = global_batch or B * (torch.distributed.get_world_size()
                                           if torch.distributed.is_initialized() else 1)
        with torch.no_grad():
            emb_orig = self.model_orig(data, output_normalize=c["output_normalize"])   # :296-297
        self.model.eval()                                                               # :305
        if c["attack"] == "pgd":
            inner = ComputeLossWrapper(emb_orig, self.text, "mean", c["inner_loss"],
                                       self.logit_scale)                                 # :300-304
            d0 = torch.empty_like(data).uniform_(-c["eps"], c["eps"])                   # :318
            data_adv = attacks.pgd(self.model, inner, data, targets, c["norm"], c["eps"],
                                   c["iterations"], c["stepsize"], c["output_normalize"],
                                   perturbation=d0, mode="max")                        # :308-321
        elif c["attack"] == "apgd":
            inner = ComputeLossWrapper(emb_orig, self.text, "none", c["inner_loss"],
                                       self.logit_scale)
            data_adv = attacks.apgd_train(self.model, data, targets, c["norm"], c["eps"],
                                          n_iter=c["iterations"], loss_fn=inner)       # :324-333
        else:
            data_adv = data                                                             # :334-335
        self.model.train()                                                              # :338
        cw = float(c["clean_weight"])
        loss_clean = None
        if cw > 0.0:                                                                    # :340-345
            spec_c = LossSpec(c["loss_clean"], c["output_normalize"], emb_orig=emb_orig, text=None,
                              logit_scale=self.logit_scale, targets=targets)
            loss_clean_b, emb_clean = self.train_engine.train_grads(data, spec_c, self.grads2,
                                                                    cw / B)
            loss_clean = loss_clean_b.mean()
        else:
            with torch.no_grad():
                emb_clean = self.train_engine.forward(data,
                                                      output_normalize=c["output_normalize"])
        ref_emb = emb_clean.detach().clone() if c["trades"] else emb_orig               # :352-360
        spec = LossSpec(c["loss"], c["output_normalize"], emb_orig=ref_emb, text=self.text,
                        logit_scale=self.logit_scale, targets=targets)
        dist = torch.distributed
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # per-block gradient buckets all-reduced under the rest of the backward (even shards,
        # a single backward); otherwise one blocking all-reduce after it
        overlap = (world > 1 and cw == 0.0 and B * world == global_batch
                   and self.dev.type == "cuda" and self._bucketed is not False)
        works = None
        if overlap:
            if self._bucketed is None:
                self._bucketed = parallel.BucketedAllReduce(self.layout, self.vt.arch["layers"],
                                                            self.dev)
            loss_b, emb_adv = self.train_engine.train_grads(data_adv, spec, self.grads,
                                                            (1.0 - cw) / B,
                                                            events=self._bucketed.events)
            works = self._bucketed.launch(self.flat_g)
        else:
            loss_b, emb_adv = self.train_engine.train_grads(data_adv, spec, self.grads,
                                                            (1.0 - cw) / B)             # :349-362
        loss = loss_b.mean()
        if cw > 0.0:
            nat.call("rvlm_add_f32", nat.ptr(self.flat_g), nat.ptr(self.flat_g),
                     nat.ptr(self.flat_g2), c_int64(self.n & ~3), nat.stream_ptr())
            if self.n & 3:
                self.flat_g[self.n & ~3:] += self.flat_g2[self.n & ~3:]
            loss_total = cw * loss_clean + (1.0 - cw) * loss                            # :361
        else:
            loss_total = loss
        if works is not None:
            self._bucketed.finish(works, self.flat_g)
        else:
            parallel.allreduce_mean_(self.flat_g, B, global_batch)
        lr_used = self.lr
        self._optimizer_step()                                                          # :363
        self.step_total += 1                                                            # :365
        self.lr = self.lr_at(self.step_total)                                           # :366
        self._refresh()
        with torch.no_grad():                                                           # :368-385
            cos = torch.nn.functional.cosine_similarity
            log = dict(loss=loss, loss_total=loss_total, loss_clean=loss_clean, emb_adv=emb_adv,
                       data_adv=data_adv,
                       lr=lr_used, step=self.step_total,
                       cos_sim_clean=cos(emb_clean, emb_orig, dim=1).mean(),
                       cos_sim=cos(emb_adv, emb_orig, dim=1).mean(), acc=None, racc=None)
            if is_classification and self.text is not None:
                log["racc"] = compute_acc(emb_adv @ self.text, targets)
                log["acc"] = compute_acc(torch.nn.functional.normalize(emb_clean, dim=1) @ self.text,
                                         targets)
        if output_dir is not None:                                                      # :467-479
            log["saved"] = checkpoint.save_periodic(
                self.vt.state_dict(), self.optimizer_state_dict, output_dir, self.step_total,
                self.steps, save_checkpoints)             # rank 0 only (checkpoint.is_writer_rank)
        return log

    def evaluate(self, data_eval, targets_eval, n_iter=50):
        """The in-training evaluation (:390-425): clean and robust accuracy against a supervised
        50-step APGD on one validation batch, and the clean / adversarial cosine similarity."""
        c = self.cfg
        assert self.text is not None, "evaluation needs the class text embeddings"
        self.model.eval()
        data_eval = data_eval.to(self.dev, torch.float32)
        targets_eval = targets_eval.to(self.dev)
        wrap = ComputeLossWrapper(None, self.text, "none", "ce", 100.0)                 # :395-398
        adv = attacks.apgd_train(self.model, data_eval, targets_eval, c["norm"], c["eps"], n_iter=n_iter,
                                 loss_fn=wrap, initial_stepsize=0.05 * c["eps"]
                                 if c["clean_weight"] > 0 else None)                    # :399-409
        with torch.no_grad():
            e_adv = self.model(adv, output_normalize=True)
            e_clean = self.model(data_eval, output_normalize=True)
            out = {"eval/racc": compute_acc(e_adv @ self.text, targets_eval),
                   "eval/acc": compute_acc(e_clean @ self.text, targets_eval),
                   "eval/cos-sim": torch.nn.functional.cosine_similarity(e_adv, e_clean,
                                                                          dim=1).mean()}
        self.model.train()
        return out
