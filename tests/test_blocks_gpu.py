"""Building-block kernels vs plain PyTorch fp32 references of the same op."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _nat():
    from robustvlm_b200 import _native as nat
    return nat


def _log(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "test_metrics.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **kw}) + "\n")


def _dt(fmt):
    return torch.bfloat16 if fmt else torch.float16


@pytest.mark.parametrize("M,N,K,fmt,flags", [
    (128, 256, 64, 0, 0), (1000, 768, 1024, 0, 0), (514, 640, 1024, 1, 0),
    (4112, 3072, 1024, 0, 0), (256, 512, 512, 0, 1), (256, 256, 512, 0, 6), (300, 96, 192, 1, 0),
    # 2-CTA kernel (tcgen05.mma.cta_group::2)
    (256, 256, 64, 0, 8), (512, 512, 1024, 0, 8), (1000, 768, 1024, 1, 8), (4112, 3072, 1024, 0, 8),
    (300, 96, 192, 1, 8),
])
@pytest.mark.parametrize("tma", ["2", "0"])   # TMA-store epilogue / transposing epilogue
def test_gemm16(M, N, K, fmt, flags, tma, monkeypatch):
    nat = _nat()
    monkeypatch.setenv("RVLM_EPI_TMA", tma)
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").to(_dt(fmt))
    B = torch.randn(N, K, device="cuda").to(_dt(fmt))
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ B.float().t() + bias
    A_in = A.t().contiguous() if flags & 2 else A
    B_in = B.t().contiguous() if flags & 1 else B
    C = torch.empty(M, N, device="cuda", dtype=_dt(fmt))
    nat.call("rvlm_gemm16_out16", nat.ptr(A_in), nat.c_int(A_in.stride(0)), nat.ptr(B_in),
             nat.c_int(B_in.stride(0)), nat.ptr(C), nat.c_int(N), nat.ptr(bias), nat.c_int(M),
             nat.c_int(N), nat.c_int(K), nat.c_int(fmt), nat.c_int(flags), nat.stream_ptr())
    torch.cuda.synchronize()
    tol = (2e-2 if fmt else 3e-3)
    err = (C.float() - ref).abs().max().item()
    _log("gemm16", M=M, N=N, K=K, fmt=fmt, flags=flags, tma=tma, err=err)
    assert err <= tol * ref.abs().max().item()


@pytest.mark.parametrize("tma", ["2", "1", "0"])
def test_gemm_fp32_residual_and_gelu(tma, monkeypatch):
    nat = _nat()
    monkeypatch.setenv("RVLM_EPI_TMA", tma)
    torch.manual_seed(0)
    M, N, K = 700, 512, 256
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    B = (torch.randn(N, K, device="cuda") * 0.1).half()
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    ref = A.float() @ B.float().t() + bias
    C = torch.empty(M, N, device="cuda")
    nat.call("rvlm_gemm16_out32", nat.ptr(A), nat.c_int(K), nat.ptr(B), nat.c_int(K), nat.ptr(C),
             nat.c_int(N), nat.ptr(bias), nat.ptr(resid), nat.c_int(N), nat.c_int(M), nat.c_int(N),
             nat.c_int(K), nat.c_int(0), nat.c_int(0), nat.stream_ptr())
    Gh = torch.empty(M, N, device="cuda", dtype=torch.half)   # QuickGELU'(pre-activation) - 1/2
    Act = torch.empty(M, N, device="cuda", dtype=torch.half)
    nat.call("rvlm_gemm16_gelu", nat.ptr(A), nat.c_int(K), nat.ptr(B), nat.c_int(K), nat.ptr(Gh),
             nat.ptr(Act), nat.c_int(N), nat.ptr(bias), nat.c_int(M), nat.c_int(N), nat.c_int(K),
             nat.c_int(0), nat.stream_ptr())
    # gelu backward: out = (G @ Bt^T) * Gh, G [M, K2] bf16, Bt [N, K2], Gh the stored derivative
    K2 = 128
    G = torch.randn(M, K2, device="cuda").bfloat16()
    Bt = (torch.randn(N, K2, device="cuda") * 0.1).bfloat16()
    dH = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    nat.call("rvlm_gemm16_gelu_bwd", nat.ptr(G), nat.c_int(K2), nat.ptr(Bt), nat.c_int(K2),
             nat.ptr(dH), nat.ptr(Gh), nat.c_int(N), nat.c_int(M), nat.c_int(N), nat.c_int(K2),
             nat.c_int(1), nat.c_int(0), nat.stream_ptr())
    torch.cuda.synchronize()
    assert (C - (ref + resid)).abs().max() < 2e-3
    sg = torch.sigmoid(1.702 * ref)
    gel = ref * sg
    assert (Act.float() - gel).abs().max() < 2e-2
    gprime = sg * (1 + 1.702 * ref * (1 - sg))
    assert (Gh.float() + 0.5 - gprime).abs().max() < 1e-3     # fp16 of a value in [-0.6, 0.6]
    dref = (G.float() @ Bt.float().t()) * gprime
    assert (dH.float() - dref).abs().max() <= 2e-2 * dref.abs().max()


@pytest.mark.parametrize("W", [128, 768, 1024])
def test_layernorm_fwd_bwd(W):
    nat = _nat()
    torch.manual_seed(W)
    M = 515
    x = (torch.randn(M, W, device="cuda") * 2 + 0.3).requires_grad_(True)
    g = torch.randn(W, device="cuda") * 0.1 + 1
    b = torch.randn(W, device="cuda") * 0.1
    y_ref = F.layer_norm(x, (W,), g, b, 1e-5)
    dy = torch.randn(M, W, device="cuda")
    dres = torch.randn(M, W, device="cuda")
    (dx_ref,) = torch.autograd.grad(y_ref, x, dy)
    y16 = torch.empty(M, W, device="cuda", dtype=torch.half)
    y32 = torch.empty(M, W, device="cuda")
    stats = torch.empty(M, 2, device="cuda")
    nat.call("rvlm_layernorm_fwd", nat.ptr(x.detach()), nat.ptr(g), nat.ptr(b), nat.ptr(y16),
             nat.ptr(y32), nat.ptr(stats), nat.c_int(M), nat.c_int(W), nat.c_int(0),
             nat.stream_ptr())
    dx32 = dres.clone()
    dx16 = torch.empty(M, W, device="cuda", dtype=torch.bfloat16)
    nat.call("rvlm_layernorm_bwd", nat.ptr(dy), nat.ptr(x.detach()), nat.ptr(stats), nat.ptr(g),
             nat.ptr(dx32), nat.ptr(dx32), nat.ptr(dx16), nat.c_int(M), nat.c_int(W), nat.c_int(1),
             nat.c_int(0), nat.stream_ptr())
    torch.cuda.synchronize()
    assert (y32 - y_ref).abs().max() < 1e-5
    assert (y16.float() - y_ref).abs().max() < 5e-3
    assert (dx32 - (dx_ref + dres)).abs().max() < 2e-5
    assert (dx16.float() - (dx_ref + dres)).abs().max() < 5e-2
    # compaction: drop every T-th row
    T = 103
    dxc = torch.zeros(M - M // T, W, device="cuda", dtype=torch.bfloat16)
    nat.call("rvlm_layernorm_bwd", nat.ptr(dy), nat.ptr(x.detach()), nat.ptr(stats), nat.ptr(g),
             nat.c_void_p(0), nat.c_void_p(0), nat.ptr(dxc), nat.c_int(M), nat.c_int(W),
             nat.c_int(1), nat.c_int(T), nat.stream_ptr())
    torch.cuda.synchronize()
    keep = torch.arange(M, device="cuda") % T != 0
    assert (dxc.float() - dx_ref[keep]).abs().max() < 5e-2
    # 16-bit dy (the attack path's hand-over format): exact w.r.t. the rounded dy
    dy16 = dy.to(torch.bfloat16)
    (dx_ref16,) = torch.autograd.grad(F.layer_norm(x, (W,), g, b, 1e-5), x, dy16.float())
    dx32b = dres.clone()
    nat.call("rvlm_layernorm_bwd16", nat.ptr(dy16), nat.ptr(x.detach()), nat.ptr(stats), nat.ptr(g),
             nat.ptr(dx32b), nat.ptr(dx32b), nat.ptr(dx16), nat.c_int(M), nat.c_int(W),
             nat.c_int(1), nat.c_int(0), nat.stream_ptr())
    torch.cuda.synchronize()
    assert (dx32b - (dx_ref16 + dres)).abs().max() < 2e-5


def _attn_ref(qkv, B, T, H):
    W = H * 64
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = s.softmax(-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * T, W)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("impl", ["simt", "tc", "tc3", "tc_fwd_two_sweep"])
@pytest.mark.parametrize("B,T,H", [(2, 50, 2), (3, 257, 4), (1, 257, 16), (2, 256, 4), (40, 257, 8),
                                   (37, 257, 16)])
def test_attention_fwd_bwd(impl, B, T, H, monkeypatch):
    nat = _nat()
    # forward: T = 257 runs attn_fwd3 (one tile per CTA, two CTAs per SM; one sweep over S by
    # default, RVLM_ATTN_FWD_ONE=0: the max pass + exp pass variant), other T the generic kernel
    monkeypatch.setenv("RVLM_ATTN_FWD_ONE", "0" if impl == "tc_fwd_two_sweep" else "1")
    if impl == "tc_fwd_two_sweep" and T != 257:
        pytest.skip("attn_fwd3 serves T = 257 only")
    # backward: T in {256, 257} runs the fused dQ / dK / dV kernel ("tc3"), other token counts the
    # generic dQ + dK/dV pair ("tc")
    if impl == "tc" and T >= 256:
        pytest.skip("T in {256, 257}: covered by tc3")
    if impl == "tc3" and T < 256:
        pytest.skip("fused backward serves T in {256, 257}")
    impl = "simt" if impl == "simt" else "tc"
    torch.manual_seed(B * T + H)
    W = H * 64
    qkv = (torch.randn(B * T, 3 * W, device="cuda") * 1.5).half()
    o = torch.empty(B * T, W, device="cuda", dtype=torch.half)
    lse = torch.empty(B, H, T, device="cuda")
    sfx = "_simt" if impl == "simt" else ""
    nat.call("rvlm_attn_fwd" + sfx, nat.ptr(qkv), nat.ptr(o), nat.ptr(lse), nat.c_int(B),
             nat.c_int(T), nat.c_int(H), nat.c_int(0), nat.stream_ptr())
    torch.cuda.synchronize()
    qf = qkv.float().requires_grad_(True)
    o_ref, lse_ref = _attn_ref(qf, B, T, H)
    e_o = (o.float() - o_ref).abs().max().item()
    e_l = (lse - lse_ref).abs().max().item()
    _log("attn_fwd", impl=impl, B=B, T=T, H=H, err_o=e_o, err_lse=e_l)
    assert e_o < 4e-3 and e_l < 2e-3
    # backward: every operand in the gradient format (bf16 copy of q|k|v), statistics recomputed
    d_o = (torch.randn(B * T, W, device="cuda") * 0.1).bfloat16()
    qkv_b = qkv.float().bfloat16()
    qb = qkv_b.float().requires_grad_(True)
    o_ref_b, _ = _attn_ref(qb, B, T, H)
    (dqkv_ref,) = torch.autograd.grad(o_ref_b, qb, d_o.float())
    dqkv = torch.empty(B * T, 3 * W, device="cuda", dtype=torch.bfloat16)
    lse_tmp = torch.empty(B, H, T, device="cuda")
    nat.call("rvlm_attn_bwd" + sfx, nat.ptr(qkv_b), nat.ptr(o), nat.ptr(d_o), nat.ptr(lse_tmp),
             nat.ptr(dqkv), nat.c_int(B), nat.c_int(T), nat.c_int(H), nat.c_int(1), nat.c_int(0),
             nat.c_int(1), nat.stream_ptr())
    torch.cuda.synchronize()
    err = (dqkv.float() - dqkv_ref).abs().max().item()
    scale = dqkv_ref.abs().max().item()
    _log("attn_bwd", impl=impl, B=B, T=T, H=H, err=err, scale=scale)
    assert err <= 3e-2 * scale


@pytest.mark.parametrize("one", ["1", "0"])
def test_attention_fwd_uneven_halves(one, monkeypatch):
    """attn_fwd3 normalises each 128-key half of a row by its own maximum and joins the halves in
    the epilogue: rows whose two halves (and the peeled 257th key) differ by many e-folds in
    either direction must come out like the fp32 reference."""
    nat = _nat()
    monkeypatch.setenv("RVLM_ATTN_FWD_ONE", one)
    B, T, H = 6, 257, 4
    W = H * 64
    torch.manual_seed(5)
    x = torch.randn(B, T, 3, H, 64, device="cuda")
    # per image a different key-scale pattern: (first half, second half, token 256)
    for b, (s0, s1, st) in enumerate([(4, .2, 1), (.2, 4, 1), (1, 1, 6), (5, 5, .1), (.05, .05, .05),
                                      (3, .5, 3)]):
        x[b, :128, 1] *= s0
        x[b, 128:256, 1] *= s1
        x[b, 256, 1] *= st
    qkv = x.reshape(B * T, 3 * W).half()
    o = torch.empty(B * T, W, device="cuda", dtype=torch.half)
    lse = torch.empty(B, H, T, device="cuda")
    nat.call("rvlm_attn_fwd", nat.ptr(qkv), nat.ptr(o), nat.ptr(lse), nat.c_int(B), nat.c_int(T),
             nat.c_int(H), nat.c_int(0), nat.stream_ptr())
    torch.cuda.synchronize()
    o_ref, lse_ref = _attn_ref(qkv.float(), B, T, H)
    e_o = (o.float() - o_ref).abs().max().item()
    e_l = (lse - lse_ref).abs().max().item()
    _log("attn_fwd_uneven", one=one, err_o=e_o, err_lse=e_l)
    assert torch.isfinite(o.float()).all() and e_o < 4e-3 and e_l < 2e-3


@pytest.mark.parametrize("B,T,H", [(3, 257, 4), (2, 50, 2), (5, 257, 16)])
def test_attention_cls_row(B, T, H):
    """Class-token-only attention of the last block: forward row 0 and the dense dqkv it
    induces, vs autograd of the full attention with an output gradient on row 0 only."""
    nat = _nat()
    torch.manual_seed(B + T + H)
    W = H * 64
    qkv = (torch.randn(B * T, 3 * W, device="cuda") * 1.5).half()
    o = torch.full((B * T, W), 7.0, device="cuda", dtype=torch.half)
    nat.call("rvlm_attn_cls_fwd", nat.ptr(qkv), nat.ptr(o), nat.c_int(B), nat.c_int(T), nat.c_int(H),
             nat.c_int(0), nat.stream_ptr())
    torch.cuda.synchronize()
    o_ref, _ = _attn_ref(qkv.float(), B, T, H)
    cls = torch.arange(B, device="cuda") * T
    assert (o[cls].float() - o_ref[cls]).abs().max().item() < 4e-3
    rest = torch.ones(B * T, dtype=torch.bool, device="cuda")
    rest[cls] = False
    assert bool((o[rest] == 7.0).all())                       # other rows untouched
    qkv_b = qkv.float().bfloat16()
    qb = qkv_b.float().requires_grad_(True)
    o_b, _ = _attn_ref(qb, B, T, H)
    d_cls = (torch.randn(B, W, device="cuda") * 0.1).bfloat16()
    d_full = torch.zeros(B * T, W, device="cuda")
    d_full[cls] = d_cls.float()
    (dqkv_ref,) = torch.autograd.grad(o_b, qb, d_full)
    dqkv = torch.full((B * T, 3 * W), float("nan"), device="cuda", dtype=torch.bfloat16)
    nat.call("rvlm_attn_cls_bwd", nat.ptr(qkv_b), nat.ptr(d_cls), nat.ptr(dqkv), nat.c_int(B),
             nat.c_int(T), nat.c_int(H), nat.c_int(1), nat.stream_ptr())
    torch.cuda.synchronize()
    err = (dqkv.float() - dqkv_ref).abs().max().item()
    scale = dqkv_ref.abs().max().item()
    _log("attn_cls_bwd", B=B, T=T, H=H, err=err, scale=scale)
    assert err <= 1e-2 * scale


@pytest.mark.parametrize("fmt", [0, 1])
def test_ts_mma_operand_packing(fmt):
    """tcgen05.mma with A in TMEM: A[i,k] at lane i, 32-bit column k/2 (low half = even k)."""
    nat = _nat()
    torch.manual_seed(fmt)
    A = torch.randn(128, 64, device="cuda").to(_dt(fmt))
    B = torch.randn(64, 64, device="cuda").to(_dt(fmt))
    C = torch.zeros(128, 64, device="cuda")
    nat.call("rvlm_selftest_ts_mma", nat.ptr(A), nat.ptr(B), nat.ptr(C), nat.c_int(fmt),
             nat.stream_ptr())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (C - ref).abs().max().item()
    _log("ts_mma", fmt=fmt, err=err)
    assert err < 1e-2


# ------------------------------------------------------------------------------------------
# split-operand ("fp32-class") kernels: fp16 hi + lo pairs, hi*hi + lo*hi + hi*lo
# ------------------------------------------------------------------------------------------
def _split16(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi.contiguous(), lo.contiguous()


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 768, 1024), (4112, 1024, 4096),
                                   (300, 96, 192), (4, 512, 1024)])
def test_gemm16_split(M, N, K):
    """C = A B^T + bias + resid with both operands split: fp32-class result (the plain fp16
    kernel is ~3e-4 relative on the same data)"""
    nat = _nat()
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") * K ** -0.5
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    ref = (A.double() @ B.double().t() + bias.double() + resid.double())
    (Ah, Al), (Bh, Bl) = _split16(A), _split16(B)
    C = torch.empty(M, N, device="cuda")
    nat.call("rvlm_gemm16_split_out32", nat.ptr(Ah), nat.ptr(Al), nat.c_int(K), nat.ptr(Bh),
             nat.ptr(Bl), nat.c_int(K), nat.ptr(C), nat.c_int(N), nat.ptr(bias), nat.ptr(resid),
             nat.c_int(N), nat.c_int(M), nat.c_int(N), nat.c_int(K), nat.stream_ptr())
    C1 = torch.empty(M, N, device="cuda")
    nat.call("rvlm_gemm16_out32", nat.ptr(Ah), nat.c_int(K), nat.ptr(Bh), nat.c_int(K), nat.ptr(C1),
             nat.c_int(N), nat.ptr(bias), nat.ptr(resid), nat.c_int(N), nat.c_int(M), nat.c_int(N),
             nat.c_int(K), nat.c_int(0), nat.c_int(8), nat.stream_ptr())
    torch.cuda.synchronize()
    prod = (A.double() @ B.double().t())
    err = float(((C.double() - ref).norm() / prod.norm()))
    err1 = float(((C1.double() - ref).norm() / prod.norm()))
    _log("gemm16_split", M=M, N=N, K=K, rel_err=err, rel_err_fp16=err1)
    # the tensor core's fp32 accumulation truncates, so the error grows with the number of
    # accumulation steps: measured 3.5e-7 (K = 64), 3.7e-6 (K = 1024), 1.4e-5 (K = 4096)
    assert err < 7e-9 * K + 1e-6 and err < 0.1 * err1


@pytest.mark.parametrize("B,T,H", [(3, 257, 4), (2, 50, 2), (37, 257, 1), (1, 256, 2), (2, 129, 2)])
def test_attention_fwd_split(B, T, H):
    """softmax(q k^T / 8) v with hi + lo operands vs torch in fp64"""
    nat = _nat()
    torch.manual_seed(B * 1000 + T + H)
    W = 64 * H
    qkv = torch.randn(B * T, 3 * W, device="cuda") * 1.5
    qh, ql = _split16(qkv)
    oh = torch.zeros(B * T, W, device="cuda", dtype=torch.half)
    ol = torch.zeros_like(oh)
    lse = torch.zeros(B, H, T, device="cuda")
    nat.call("rvlm_attn_fwd_split", nat.ptr(qh), nat.ptr(ql), nat.ptr(oh), nat.ptr(ol), nat.ptr(lse),
             nat.c_int(B), nat.c_int(T), nat.c_int(H), nat.stream_ptr())
    torch.cuda.synchronize()
    x = (qh.double() + ql.double()).view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (x[0] @ x[1].transpose(-1, -2)) * 0.125
    ref = (torch.softmax(s, -1) @ x[2]).permute(0, 2, 1, 3).reshape(B * T, W)
    o = oh.double() + ol.double()
    err = float((o - ref).norm() / ref.norm())
    err_hi = float((oh.double() - ref).norm() / ref.norm())
    lse_err = float((lse.double() - torch.logsumexp(s, -1)).abs().max())
    _log("attn_fwd_split", B=B, T=T, H=H, rel_err=err, rel_err_hi_only=err_hi, lse_err=lse_err)
    assert err < 5e-6 and lse_err < 1e-4


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("B,T,H", [(3, 257, 4), (1, 257, 16), (2, 256, 4), (40, 257, 8), (37, 257, 16)])
def test_attention_bwd_with_forward_lse(B, T, H, fmt):
    """rvlm_attn_bwd_lse (version 4: no statistics recomputed, 16 worker warps) on the same q | k | v
    bits the forward used, vs autograd through the fp32 reference and vs the self-normalising
    version 3 on identical inputs."""
    nat = _nat()
    torch.manual_seed(B * T + H + fmt)
    W = H * 64
    dt = _dt(fmt)
    qkv = (torch.randn(B * T, 3 * W, device="cuda") * 1.5).to(dt)
    o = torch.empty(B * T, W, device="cuda", dtype=dt)
    lse = torch.empty(B, H, T, device="cuda")
    nat.call("rvlm_attn_fwd", nat.ptr(qkv), nat.ptr(o), nat.ptr(lse), nat.c_int(B), nat.c_int(T),
             nat.c_int(H), nat.c_int(fmt), nat.stream_ptr())
    d_o = (torch.randn(B * T, W, device="cuda") * 0.1).to(dt)
    qf = qkv.float().requires_grad_(True)
    o_ref, _ = _attn_ref(qf, B, T, H)
    (dqkv_ref,) = torch.autograd.grad(o_ref, qf, d_o.float())
    out = {}
    for tag, fn, env in (("v4", "rvlm_attn_bwd_lse", "1"), ("v3", "rvlm_attn_bwd_lse", "0")):
        os.environ["RVLM_ATTN_BWD4"] = env
        dqkv = torch.full((B * T, 3 * W), float("nan"), device="cuda", dtype=dt)
        lse_in = lse.clone()
        try:
            nat.call(fn, nat.ptr(qkv), nat.ptr(o), nat.ptr(d_o), nat.ptr(lse_in), nat.ptr(dqkv),
                     nat.c_int(B), nat.c_int(T), nat.c_int(H), nat.c_int(fmt), nat.c_int(fmt),
                     nat.c_int(fmt), nat.stream_ptr())
            torch.cuda.synchronize()
        finally:
            os.environ.pop("RVLM_ATTN_BWD4", None)
        out[tag] = dqkv.float()
    scale = dqkv_ref.abs().max().item()
    e4 = (out["v4"] - dqkv_ref).abs().max().item()
    e3 = (out["v3"] - dqkv_ref).abs().max().item()
    rel4 = float((out["v4"] - dqkv_ref).norm() / dqkv_ref.norm())
    rel3 = float((out["v3"] - dqkv_ref).norm() / dqkv_ref.norm())
    _log("attn_bwd_lse", B=B, T=T, H=H, fmt=fmt, err_v4=e4, err_v3=e3, rel_v4=rel4, rel_v3=rel3,
         scale=scale)
    assert torch.isfinite(out["v4"]).all()
    assert e4 <= (3e-2 if fmt else 6e-3) * scale and rel4 <= 1.3 * rel3 + 1e-4
