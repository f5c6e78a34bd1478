"""Checkpoint I/O in the reference's on-disk formats.

The reference writes, per save point (train/adversarial_training_clip.py:238-239,468-479):

    {output_dir}/checkpoints/step_{N}.pt          unwrap_model(model).model.state_dict()
    {output_dir}/checkpoints/step_{N}_opt.pt      optimizer.state_dict()
    .../fallback_{N}.pt, fallback_{N}_opt.pt      every 200 steps, older fallbacks removed
    .../final.pt, final_opt.pt                    at the end

i.e. the vision tower in open_clip `VisionTransformer` key layout (what
CLIP_eval/eval_utils.py:45-48 and README.md:57-67 load back with
`model.visual.load_state_dict(...)`), and a `torch.optim.AdamW` / `SGD` state dict whose
per-parameter entries follow `model.parameters()` order.  The trainer here keeps parameters,
gradients and optimizer moments in flat fp32 buffers; the functions below convert between
those buffers and the reference's dictionaries, so checkpoints move in both directions
(`--pretrained <file>` / `--optimizer_state <file>`, :93-103,207-208).  Pure tensor code: no
kernel, works on any device (host-side format logic, tested on CPU).
"""
from __future__ import annotations

import math
import os
import re

import torch


def param_layout(module: torch.nn.Module):
    """[(name, shape, offset, numel)] in `module.parameters()` order -- the index order of a
    torch optimizer state dict built from `model.parameters()` (:194-197)."""
    out, off = [], 0
    for name, p in module.named_parameters():
        n = math.prod(p.shape)
        out.append((name, tuple(p.shape), off, n))
        off += n
    return out


def adamw_state_dict(layout, flat_m, flat_v, step, *, lr, betas=(0.9, 0.999), eps=1e-8,
                     weight_decay=1e-2):
    """`torch.optim.AdamW(...).state_dict()` from flat moment buffers.  Before the first step
    torch holds no per-parameter state (empty `state`)."""
    state = {}
    if step > 0:
        for i, (_, shape, off, n) in enumerate(layout):
            state[i] = {"step": torch.tensor(float(step)),
                        "exp_avg": flat_m[off:off + n].view(shape).clone(),
                        "exp_avg_sq": flat_v[off:off + n].view(shape).clone()}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay,
             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
             "differentiable": False, "fused": None, "decoupled_weight_decay": True,
             "params": list(range(len(layout)))}
    return {"state": state, "param_groups": [group]}


def sgd_state_dict(layout, flat_buf, step, *, lr, momentum=0.9, weight_decay=1e-4):
    state = {}
    if step > 0 and momentum != 0:
        for i, (_, shape, off, n) in enumerate(layout):
            state[i] = {"momentum_buffer": flat_buf[off:off + n].view(shape).clone()}
    group = {"lr": lr, "momentum": momentum, "dampening": 0, "weight_decay": weight_decay,
             "nesterov": False, "maximize": False, "foreach": None, "differentiable": False,
             "fused": None, "params": list(range(len(layout)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(sd, layout, flat_a, flat_b=None):
    """Inverse of the two functions above (`optimizer.load_state_dict(torch.load(...))`, :207-208).
    Fills the flat buffers in place; returns (step, param_group) -- `param_group['lr']` is the
    learning rate the next step uses, as in the reference after a resume."""
    groups = sd["param_groups"]
    if len(groups) != 1 or len(groups[0]["params"]) != len(layout):
        raise ValueError("optimizer state does not match the model: expected one param group of "
                         f"{len(layout)} parameters, got {[len(g['params']) for g in groups]}")
    step = 0
    for i, (name, shape, off, n) in enumerate(layout):
        st = sd["state"].get(i)
        if st is None:
            continue
        if "exp_avg" in st:
            a, b = st["exp_avg"], st["exp_avg_sq"]
            step = max(step, int(float(st["step"])))
        else:
            a, b = st["momentum_buffer"], None
            step = max(step, 1)
        if tuple(a.shape) != shape:
            raise ValueError(f"optimizer state of parameter {i} ({name}): shape {tuple(a.shape)} "
                             f"!= {shape}")
        flat_a[off:off + n].copy_(a.reshape(-1))
        if b is not None and flat_b is not None:
            flat_b[off:off + n].copy_(b.reshape(-1))
    return step, dict(groups[0])


def _atomic_save(obj, path):
    """write to a temporary file in the same directory, then rename: a reader (or a crashed
    writer) never sees a torn checkpoint"""
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def is_writer_rank() -> bool:
    """One process per GPU here (the reference is single-process DataParallel,
    train/adversarial_training_clip.py:184-190): only rank 0 touches the checkpoint directory."""
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def save_pair(model_state, opt_state, ckpt_dir, stem):
    os.makedirs(ckpt_dir, exist_ok=True)
    _atomic_save({k: v.detach().cpu() for k, v in model_state.items()},
                 os.path.join(ckpt_dir, f"{stem}.pt"))
    _atomic_save(opt_state, os.path.join(ckpt_dir, f"{stem}_opt.pt"))


def save_periodic(model_state, opt_state_fn, output_dir, step_total, total_steps,
                  save_checkpoints=True):
    """The in-loop saving rules of train_one_epoch (:467-479): ten evenly spaced `step_N`
    checkpoints, plus a rolling `fallback_N` every 200 steps.  Returns the stems written."""
    ckpt_dir = os.path.join(output_dir, "checkpoints")
    written = []
    if not is_writer_rank():     # replicas hold identical parameters: rank 0 writes for all
        return written
    if save_checkpoints and total_steps >= 10 and step_total % (total_steps // 10) == 0:
        save_pair(model_state, opt_state_fn(), ckpt_dir, f"step_{step_total}")
        written.append(f"step_{step_total}")
    if step_total % 200 == 0:
        save_pair(model_state, opt_state_fn(), ckpt_dir, f"fallback_{step_total}")
        written.append(f"fallback_{step_total}")
        for f in os.listdir(ckpt_dir):          # :476-478 (substring test, as the reference)
            if f.startswith("fallback") and str(step_total) not in f:
                try:
                    os.remove(os.path.join(ckpt_dir, f))
                except FileNotFoundError:       # removed concurrently: nothing left to do
                    pass
    return written


def load_visual_state_dict(path_or_sd, prefix_candidates=("visual.", "module.visual.", "model.")):
    """A vision-tower state dict from a FARE / TeCoA checkpoint (the tower alone, as the
    reference saves it), from the `{'vision_encoder_state_dict': ...}` wrapper of the original TeCoA
    files, or from a full CLIP state dict (keys under `visual.`)."""
    sd = torch.load(path_or_sd, map_location="cpu") if isinstance(path_or_sd, (str, os.PathLike)) \
        else path_or_sd
    if "vision_encoder_state_dict" in sd:        # the original TeCoA release, CLIP_eval/eval_utils.py:45-48
        sd = sd["vision_encoder_state_dict"]
    if "conv1.weight" in sd:
        return dict(sd)
    for pre in prefix_candidates:
        sub = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
        if "conv1.weight" in sub:
            return sub
    raise KeyError("no open_clip VisionTransformer keys (conv1.weight ...) in the checkpoint")


def step_of(filename):
    m = re.search(r"(?:step|fallback)_(\d+)(?:_opt)?\.pt$", filename)
    return int(m.group(1)) if m else None
