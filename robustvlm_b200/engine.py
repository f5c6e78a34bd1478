"""Python handle on the fused CUDA ViT attack engine (include/rvlm.h §4).

`VitEngine` owns a `rvlm_vit_handle` built from the fp32 parameters of a CLIP
vision tower in open_clip `VisionTransformer.state_dict()` layout (the layout of
the reference's checkpoints, train/adversarial_training_clip.py:239,470), plus a
caller-side workspace tensor.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, byref, c_float, c_int, c_size_t, c_void_p

import torch

from . import _native as nat

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # train/adversarial_training_clip.py:116
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

FMT_F16, FMT_BF16, FMT_F16X2 = 0, 1, 2
# Engine precision modes: (forward operand format, gradient operand format).
#   fast    fp16 forward operands (tf32-class), bf16 gradients            -- round-1 default
#   fast16  fp16 forward operands, fp16 gradients rescaled per sample (11-bit mantissa)
#   split   forward operands as fp16 hi + lo pairs, three tcgen05 MMAs per product
#           (fp32-class forward for FARE's cancellation site), fp16 gradients
#   bf16    bf16 / bf16 (the outer training step: weight-gradient GEMMs need one format)
PRECISIONS = {"fast": (FMT_F16, FMT_BF16), "fast16": (FMT_F16, FMT_F16),
              "split": (FMT_F16X2, FMT_F16), "bf16": (FMT_BF16, FMT_BF16)}
LOSS_L2, LOSS_CE, LOSS_DLR, LOSS_DLR_TARGETED = 1, 2, 4, 5
NORM_LINF, NORM_L2 = 0, 2
FLAG_ATTN_SIMT = 1


class VitConfigC(Structure):
    _fields_ = [("image", c_int), ("patch", c_int), ("width", c_int), ("layers", c_int),
                ("heads", c_int), ("mlp", c_int), ("out_dim", c_int),
                ("mean", c_float * 3), ("std", c_float * 3),
                ("fmt_fwd", c_int), ("fmt_bwd", c_int), ("flags", c_int)]


_PP = POINTER(c_void_p)


class VitWeightsC(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "conv1_weight", "class_embedding", "positional_embedding", "ln_pre_weight", "ln_pre_bias",
        "ln_post_weight", "ln_post_bias", "proj")] + [(n, _PP) for n in (
            "ln_1_weight", "ln_1_bias", "attn_in_proj_weight", "attn_in_proj_bias",
            "attn_out_proj_weight", "attn_out_proj_bias", "ln_2_weight", "ln_2_bias",
            "mlp_c_fc_weight", "mlp_c_fc_bias", "mlp_c_proj_weight", "mlp_c_proj_bias")]


class VitGradsC(Structure):
    _fields_ = VitWeightsC._fields_


class LossConfigC(Structure):
    _fields_ = [("mode", c_int), ("output_normalize", c_int), ("pred_on_logits", c_int),
                ("emb_orig", c_void_p), ("text", c_void_p), ("n_classes", c_int),
                ("logit_scale", c_float), ("targets", c_void_p), ("y_target", c_void_p)]


_LAYER_KEYS = {
    "ln_1_weight": "ln_1.weight", "ln_1_bias": "ln_1.bias",
    "attn_in_proj_weight": "attn.in_proj_weight", "attn_in_proj_bias": "attn.in_proj_bias",
    "attn_out_proj_weight": "attn.out_proj.weight", "attn_out_proj_bias": "attn.out_proj.bias",
    "ln_2_weight": "ln_2.weight", "ln_2_bias": "ln_2.bias",
    "mlp_c_fc_weight": "mlp.c_fc.weight", "mlp_c_fc_bias": "mlp.c_fc.bias",
    "mlp_c_proj_weight": "mlp.c_proj.weight", "mlp_c_proj_bias": "mlp.c_proj.bias",
}
_TOP_KEYS = {
    "conv1_weight": "conv1.weight", "class_embedding": "class_embedding",
    "positional_embedding": "positional_embedding", "ln_pre_weight": "ln_pre.weight",
    "ln_pre_bias": "ln_pre.bias", "ln_post_weight": "ln_post.weight",
    "ln_post_bias": "ln_post.bias", "proj": "proj",
}


def infer_config(sd: dict, image_size: int = 224) -> dict:
    """Architecture hyper-parameters from an open_clip visual state dict."""
    w, _, p, _ = sd["conv1.weight"].shape
    layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    mlp = sd["transformer.resblocks.0.mlp.c_fc.weight"].shape[0]
    out_dim = sd["proj"].shape[1]
    tokens = sd["positional_embedding"].shape[0]
    grid = int(round((tokens - 1) ** 0.5))
    assert grid * grid + 1 == tokens
    return dict(image=grid * p, patch=p, width=w, layers=layers, heads=w // 64, mlp=mlp,
                out_dim=out_dim)


class LossSpec:
    """Host-side description of the attack loss (mirrors ComputeLossWrapper,
    train/adversarial_training_clip.py:260-274)."""

    def __init__(self, mode, output_normalize, emb_orig=None, text=None, logit_scale=100.0,
                 targets=None, pred_on_logits=False, y_target=None):
        self.mode = {"l2": LOSS_L2, "ce": LOSS_CE, "dlr": LOSS_DLR,
                     "dlr-targeted": LOSS_DLR_TARGETED}.get(mode, mode)
        self.y_target = y_target
        self.output_normalize = bool(output_normalize)
        self.emb_orig, self.text, self.targets = emb_orig, text, targets
        self.logit_scale = float(logit_scale)
        self.pred_on_logits = bool(pred_on_logits)
        self._keep = []

    def to_c(self, lo=None, hi=None) -> LossConfigC:
        def sl(t):
            if t is None:
                return None
            t = t if lo is None else t[lo:hi]
            return t

        eo = sl(self.emb_orig)
        tg = sl(self.targets)
        yt = sl(self.y_target)
        if eo is not None:
            eo = eo.detach().float().contiguous()
        if tg is not None:
            tg = tg.detach().long().contiguous()
        if yt is not None:
            yt = yt.detach().long().contiguous()
        tx = self.text.detach().float().contiguous() if self.text is not None else None
        self._keep = [eo, tg, tx, yt]
        return LossConfigC(self.mode, int(self.output_normalize), int(self.pred_on_logits),
                           nat.ptr(eo).value, nat.ptr(tx).value,
                           tx.shape[1] if tx is not None else 0, self.logit_scale,
                           nat.ptr(tg).value, nat.ptr(yt).value)


class VitEngine:
    def __init__(self, state_dict: dict, *, image_size=224, mean=CLIP_MEAN, std=CLIP_STD,
                 fmt_fwd=FMT_F16, fmt_bwd=FMT_F16, flags=0, device="cuda"):
        if not torch.cuda.is_available():
            raise nat.RvlmError("VitEngine needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.arch = infer_config(state_dict, image_size)
        self.cfg = VitConfigC(self.arch["image"], self.arch["patch"], self.arch["width"],
                              self.arch["layers"], self.arch["heads"], self.arch["mlp"],
                              self.arch["out_dim"], (c_float * 3)(*mean), (c_float * 3)(*std),
                              fmt_fwd, fmt_bwd, flags)
        self.handle = c_void_p(0)
        self._ws = None
        self._gen = 0          # bumped by every call that writes the workspace
        self._kept = None      # (generation, B) of the forward whose activations it holds
        wc, keep = self._weights_struct(state_dict)
        with torch.cuda.device(self.device):
            nat.call("rvlm_vit_create", byref(self.cfg), byref(wc), nat.stream_ptr(),
                     byref(self.handle))
            torch.cuda.current_stream().synchronize()
        del keep

    # -- weights ------------------------------------------------------------------
    def _weights_struct(self, sd):
        keep = []

        def dev(t):
            t = t.detach().to(self.device, torch.float32).contiguous()
            keep.append(t)
            return t

        wc = VitWeightsC()
        for f, k in _TOP_KEYS.items():
            setattr(wc, f, dev(sd[k]).data_ptr())
        L = self.arch["layers"]
        for f, k in _LAYER_KEYS.items():
            arr = (c_void_p * L)(*[dev(sd[f"transformer.resblocks.{i}.{k}"]).data_ptr()
                                   for i in range(L)])
            keep.append(arr)
            setattr(wc, f, ctypes.cast(arr, _PP))
        return wc, keep

    def update_weights(self, state_dict):
        wc, keep = self._weights_struct(state_dict)
        with torch.cuda.device(self.device):
            nat.call("rvlm_vit_update_weights", self.handle, byref(wc), nat.stream_ptr())
            torch.cuda.current_stream().synchronize()
        del keep

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                nat.lib().rvlm_vit_destroy(self.handle)
                self.handle = c_void_p(0)
        except Exception:  # noqa: BLE001
            pass

    # -- workspace ----------------------------------------------------------------
    def workspace_bytes(self, B, with_backward):
        fn = nat.lib().rvlm_vit_workspace_bytes
        fn.restype = c_size_t
        return int(fn(self.handle, c_int(B), c_int(int(with_backward))))

    def _workspace(self, B, with_backward, touch=True):
        need = self.workspace_bytes(B, with_backward)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._kept = None
        if touch:
            self._gen += 1
            self._kept = None
        return self._ws, need

    def kept_stamp(self):
        """identifies the activations a `forward(keep_for_backward=True)` left in the workspace;
        None once any other call has overwritten them"""
        return self._kept

    def max_batch(self, with_backward=True, budget_bytes=None):
        """Largest chunk size whose workspace fits the memory budget."""
        if budget_bytes is None:
            free, _ = torch.cuda.mem_get_info(self.device)
            budget_bytes = int(free * 0.85) + (self._ws.numel() if self._ws is not None else 0)
        per = self.workspace_bytes(64, with_backward) / 64.0
        return max(1, int(budget_bytes / per))

    # -- calls ---------------------------------------------------------------------
    def _chk(self, x):
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3
        assert x.shape[2] == x.shape[3] == self.arch["image"], x.shape
        return x.contiguous()

    def forward(self, x, delta=None, output_normalize=False, keep_for_backward=False):
        x = self._chk(x)
        delta = self._chk(delta) if delta is not None else None
        B = x.shape[0]
        ws, nbytes = self._workspace(B, keep_for_backward)
        emb = torch.empty(B, self.arch["out_dim"], device=self.device, dtype=torch.float32)
        nat.call("rvlm_vit_forward", self.handle, nat.ptr(ws), c_size_t(nbytes), nat.ptr(x),
                 nat.ptr(delta), c_int(B), c_int(int(output_normalize)),
                 c_int(int(keep_for_backward)), nat.ptr(emb), nat.stream_ptr())
        if keep_for_backward:
            self._kept = (self._gen, B)
        return emb

    def backward_input(self, B, d_emb, output_normalize=False):
        """VJP of the forward just run with keep_for_backward=True (raises if another call has
        overwritten its activations since)."""
        if self._kept is None or self._kept[1] != B:
            raise nat.RvlmError("backward_input: no forward(keep_for_backward=True) of this batch "
                                "size is resident in the workspace (another engine call ran in "
                                "between)")
        ws, nbytes = self._workspace(B, True, touch=False)
        d_emb = d_emb.detach().float().contiguous()
        img = self.arch["image"]
        grad = torch.empty(B, 3, img, img, device=self.device, dtype=torch.float32)
        nat.call("rvlm_vit_backward_input", self.handle, nat.ptr(ws), c_size_t(nbytes), c_int(B),
                 nat.ptr(d_emb), c_int(int(output_normalize)), nat.ptr(grad), nat.stream_ptr())
        return grad

    def loss_grad(self, x, loss: LossSpec, delta=None, want_grad=True, want_logits=False):
        """-> dict(loss[B], pred[B] bool, emb[B,D], grad[B,3,H,W] or None, logits or None)"""
        x = self._chk(x)
        delta = self._chk(delta) if delta is not None else None
        B = x.shape[0]
        ws, nbytes = self._workspace(B, True)
        lc = loss.to_c()
        out_loss = torch.empty(B, device=self.device, dtype=torch.float32)
        pred = torch.empty(B, device=self.device, dtype=torch.int32)
        emb = torch.empty(B, self.arch["out_dim"], device=self.device, dtype=torch.float32)
        grad = torch.empty_like(x) if want_grad else None
        logits = (torch.empty(B, lc.n_classes, device=self.device, dtype=torch.float32)
                  if want_logits and lc.mode in (LOSS_CE, LOSS_DLR, LOSS_DLR_TARGETED) else None)
        nat.call("rvlm_vit_loss_grad", self.handle, nat.ptr(ws), c_size_t(nbytes), nat.ptr(x),
                 nat.ptr(delta), c_int(B), byref(lc), nat.ptr(grad), nat.ptr(out_loss),
                 nat.ptr(pred), nat.ptr(emb), nat.ptr(logits), c_int(int(want_grad)),
                 nat.stream_ptr())
        return dict(loss=out_loss, pred=pred.bool(), emb=emb, grad=grad, logits=logits)

    def pgd_run(self, x, loss: LossSpec, *, norm, eps, stepsize, iterations, maximize,
                momentum=0.9, delta0=None, chunk=None, trace=False):
        """The whole pgd() loop on device; returns x + delta (and the per-iteration
        per-sample loss if trace)."""
        x = self._chk(x)
        delta0 = self._chk(delta0) if delta0 is not None else None
        B = x.shape[0]
        chunk = B if chunk is None else min(chunk, B)
        out = torch.empty_like(x)
        tr = (torch.empty(iterations, B, device=self.device, dtype=torch.float32)
              if trace else None)
        ws, nbytes = self._workspace(chunk, True)
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            lc = loss.to_c(lo, hi)
            trc = (torch.empty(iterations, hi - lo, device=self.device, dtype=torch.float32)
                   if trace else None)
            nat.call("rvlm_pgd_run", self.handle, nat.ptr(ws), c_size_t(nbytes), nat.ptr(x[lo:hi]),
                     nat.ptr(delta0[lo:hi]) if delta0 is not None else c_void_p(0),
                     c_int(hi - lo), byref(lc), c_int(norm), c_float(eps), c_float(stepsize),
                     c_int(iterations), c_int(int(maximize)), c_float(momentum),
                     nat.ptr(out[lo:hi]), nat.ptr(trc), nat.stream_ptr())
            if trace:
                tr[:, lo:hi] = trc
        return (out, tr) if trace else out

    def train_grads(self, x, loss: LossSpec, grads: dict, loss_scale: float, events=None):
        """Outer training step: forward(x) + loss + full backward.  `grads` maps every
        state-dict key to a contiguous fp32 CUDA tensor of the parameter's shape; they are
        overwritten with d(loss_scale * sum_b loss_b)/d(param).  Returns (loss[B], emb[B,D]).
        `events`: layers + 1 torch.cuda.Event objects; events[l] is recorded when block l's
        parameter gradients are enqueued, events[-1] at the end (gradient bucketing)."""
        x = self._chk(x)
        B = x.shape[0]
        ws, nbytes = self._workspace(B, 2)
        gc, keep = VitGradsC(), []
        for f, k in _TOP_KEYS.items():
            setattr(gc, f, grads[k].data_ptr())
        L = self.arch["layers"]
        for f, k in _LAYER_KEYS.items():
            arr = (c_void_p * L)(*[grads[f"transformer.resblocks.{i}.{k}"].data_ptr()
                                   for i in range(L)])
            keep.append(arr)
            setattr(gc, f, ctypes.cast(arr, _PP))
        lc = loss.to_c()
        out_loss = torch.empty(B, device=self.device, dtype=torch.float32)
        emb = torch.empty(B, self.arch["out_dim"], device=self.device, dtype=torch.float32)
        if events is None:
            nat.call("rvlm_vit_train_grads", self.handle, nat.ptr(ws), c_size_t(nbytes), nat.ptr(x),
                     c_int(B), byref(lc), c_float(loss_scale), byref(gc), nat.ptr(out_loss),
                     nat.ptr(emb), nat.stream_ptr())
        else:
            assert len(events) == L + 1
            evs = (c_void_p * (L + 1))(*[e.cuda_event for e in events])
            nat.call("rvlm_vit_train_grads_ev", self.handle, nat.ptr(ws), c_size_t(nbytes),
                     nat.ptr(x), c_int(B), byref(lc), c_float(loss_scale), byref(gc),
                     nat.ptr(out_loss), nat.ptr(emb), evs, c_int(L + 1), nat.stream_ptr())
        del keep
        return out_loss, emb

    def debug_buffer(self, B, with_backward, which, layer, shape, dtype):
        ws, _ = self._workspace(B, with_backward)
        p = c_void_p(0)
        nat.call("rvlm_vit_debug_buffer", self.handle, nat.ptr(ws), c_int(B),
                 c_int(int(with_backward)), c_int(which), c_int(layer), byref(p))
        n = 1
        for s in shape:
            n *= s
        off = p.value - ws.data_ptr()
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return ws[off:off + nbytes].view(dtype).view(*shape)
