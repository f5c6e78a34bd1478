"""Time rvlm_attn_fwd (and optionally bwd) alone at the ViT-L/14 bench shape (B = 256, T = 257,
H = 16) with CUDA events; checks the forward against a plain fp32 PyTorch reference on a slice.
RVLM_ATTN_FWD_ONE=0 selects the two-sweep variant of the forward."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustvlm_b200 import _native as nat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--fmt", type=int, default=0)
    args = ap.parse_args()
    B, T, H = args.batch, 257, 16
    W = H * 64
    torch.manual_seed(0)
    dt = torch.half if args.fmt == 0 else torch.bfloat16
    qkv = (torch.randn(B * T, 3 * W, device="cuda") * 1.5).to(dt)
    o = torch.empty(B * T, W, device="cuda", dtype=dt)
    lse = torch.empty(B, H, T, device="cuda")
    call = lambda: nat.call("rvlm_attn_fwd", nat.ptr(qkv), nat.ptr(o), nat.ptr(lse), nat.c_int(B),  # noqa: E731
                            nat.c_int(T), nat.c_int(H), nat.c_int(args.fmt), nat.stream_ptr())
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.reps
    nb = min(B, 4)
    q, k, v = qkv[: nb * T].float().view(nb, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    o_ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(nb * T, W)
    lse_ref = torch.logsumexp(s, -1)
    # last images too (the tail of the work list)
    q2, k2, v2 = qkv[-T:].float().view(1, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s2 = (q2 @ k2.transpose(-1, -2)) * 0.125
    o_ref2 = (s2.softmax(-1) @ v2).permute(0, 2, 1, 3).reshape(T, W)
    res = {"fwd_one_sweep": os.environ.get("RVLM_ATTN_FWD_ONE", "1"), "B": B, "us": us,
           "tflops": 4.0 * B * H * T * T * 64 / (us * 1e-6) / 1e12,
           "err_o": float((o[: nb * T].float() - o_ref).abs().max()),
           "err_o_last": float((o[-T:].float() - o_ref2).abs().max()),
           "err_lse": float((lse[:nb] - lse_ref).abs().max())}
    # backward at the same shape: version 4 (forward lse, 16 worker warps), version 3 (self-normalising)
    d_o = (torch.randn(B * T, W, device="cuda") * 0.1).to(dt)
    dqkv = torch.empty(B * T, 3 * W, device="cuda", dtype=dt)

    def bwd(fn, env):
        os.environ["RVLM_ATTN_BWD4"] = env
        lse_in = lse.clone()
        c = lambda: nat.call(fn, nat.ptr(qkv), nat.ptr(o), nat.ptr(d_o), nat.ptr(lse_in),  # noqa: E731
                             nat.ptr(dqkv), nat.c_int(B), nat.c_int(T), nat.c_int(H),
                             nat.c_int(args.fmt), nat.c_int(args.fmt), nat.c_int(args.fmt),
                             nat.stream_ptr())
        for _ in range(3):
            c()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            c()
        b.record()
        torch.cuda.synchronize()
        os.environ.pop("RVLM_ATTN_BWD4", None)
        return a.elapsed_time(b) * 1e3 / args.reps, dqkv.float().clone()
    us4, g4 = bwd("rvlm_attn_bwd_lse", "1")
    us3, g3 = bwd("rvlm_attn_bwd_lse", "0")
    res.update(bwd4_us=us4, bwd3_us=us3, bwd4_vs_bwd3_rel=float((g4 - g3).norm() / g3.norm()),
               bwd4_tflops=10.0 * B * H * T * T * 64 / (us4 * 1e-6) / 1e12)
    # split-operand forward (fp16 hi + lo, three MMAs per product) at the same shape
    if args.fmt == 0:
        x32 = torch.randn(B * T, 3 * W, device="cuda") * 1.5
        qh = x32.half()
        ql = (x32 - qh.float()).half()
        oh, ol = torch.empty(B * T, W, device="cuda", dtype=torch.half), torch.empty(B * T, W, device="cuda", dtype=torch.half)
        cs = lambda: nat.call("rvlm_attn_fwd_split", nat.ptr(qh), nat.ptr(ql), nat.ptr(oh), nat.ptr(ol),  # noqa: E731
                              nat.ptr(lse), nat.c_int(B), nat.c_int(T), nat.c_int(H), nat.stream_ptr())
        for key in ("split_us",):
            for _ in range(2):
                cs()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.reps):
                cs()
            b.record()
            torch.cuda.synchronize()
            res[key] = a.elapsed_time(b) * 1e3 / args.reps
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
