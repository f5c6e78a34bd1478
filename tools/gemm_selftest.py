"""GPU self-test + micro-benchmark of the tcgen05 GEMM family (run under gpurun).

Writes a JSON report to gpurun_out/gemm_selftest.json.  Diagnostics are verbose
on purpose: there is no GPU in the authoring container, so a failing case must
tell us *how* it fails (which rows / K-slices are wrong) in a single round trip.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustvlm_b200 import _native as nat  # noqa: E402
from robustvlm_b200._native import c_int, c_void_p  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda:0")
report = {"cases": [], "bench": []}


def dt(fmt):
    return torch.bfloat16 if fmt else torch.float16


def run_out16(A, B, M, N, K, fmt, flags, bias=None):
    C = torch.full((M, N), float("nan"), device=dev, dtype=dt(fmt))
    lda = A.stride(0)
    ldb = B.stride(0)
    nat.call("rvlm_gemm16_out16", nat.ptr(A), c_int(lda), nat.ptr(B), c_int(ldb), nat.ptr(C),
             c_int(N), nat.ptr(bias), c_int(M), c_int(N), c_int(K), c_int(fmt), c_int(flags),
             nat.stream_ptr())
    torch.cuda.synchronize()
    return C


def diag(out, ref, name):
    err = (out.float() - ref).abs()
    bad = ~(err <= 2e-2 * ref.abs().clamp(min=1.0))
    info = {
        "name": name,
        "max_abs_err": float(err[~err.isnan()].max()) if (~err.isnan()).any() else None,
        "nan_count": int(out.isnan().sum()),
        "bad_frac": float(bad.float().mean()),
        "ref_absmax": float(ref.abs().max()),
    }
    if bad.any():
        M, N = ref.shape
        rows_bad = bad.any(dim=1).nonzero().flatten()
        cols_bad = bad.any(dim=0).nonzero().flatten()
        info["rows_bad_first"] = rows_bad[:16].tolist()
        info["rows_bad_count"] = int(rows_bad.numel())
        info["cols_bad_first"] = cols_bad[:16].tolist()
        info["cols_bad_count"] = int(cols_bad.numel())
        info["sample_out"] = out.float()[:4, :8].tolist()
        info["sample_ref"] = ref[:4, :8].tolist()
        # bad fraction by row-within-8 and col-within-64 to expose swizzle errors
        rb = bad.float().view(M // 8 if M % 8 == 0 else -1, 8, N).mean(dim=(0, 2)).tolist() if M % 8 == 0 else None
        info["bad_by_row_mod8"] = rb
    return info


def case(M, N, K, fmt, flags, with_bias=False, tag=""):
    torch.manual_seed(M * 7 + N * 3 + K + fmt + flags)
    a_mn = bool(flags & 2)
    b_mn = bool(flags & 1)
    A_log = torch.randn(M, K, device=dev)
    B_log = torch.randn(N, K, device=dev)
    A16 = A_log.to(dt(fmt))
    B16 = B_log.to(dt(fmt))
    ref = A16.float() @ B16.float().t()
    bias = torch.randn(N, device=dev) if with_bias else None
    if with_bias:
        ref = ref + bias
    A_in = A16.t().contiguous() if a_mn else A16.contiguous()
    B_in = B16.t().contiguous() if b_mn else B16.contiguous()
    name = f"{tag}M{M}_N{N}_K{K}_fmt{fmt}_flags{flags}_bias{int(with_bias)}"
    try:
        out = run_out16(A_in, B_in, M, N, K, fmt, flags, bias)
        info = diag(out, ref, name)
    except Exception as e:  # noqa: BLE001
        info = {"name": name, "exception": repr(e)}
    print(json.dumps(info), flush=True)
    report["cases"].append(info)
    return info


def bench(M, N, K, fmt, flags, iters=20):
    A = torch.randn(M, K, device=dev).to(dt(fmt))
    B = torch.randn(N, K, device=dev).to(dt(fmt))
    C = torch.empty(M, N, device=dev, dtype=dt(fmt))

    def go():
        nat.call("rvlm_gemm16_out16", nat.ptr(A), c_int(K), nat.ptr(B), c_int(K), nat.ptr(C),
                 c_int(N), c_void_p(0), c_int(M), c_int(N), c_int(K), c_int(fmt), c_int(flags),
                 nat.stream_ptr())

    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for comparison
    for _ in range(3):
        torch.matmul(A, B.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(A, B.t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    info = {"M": M, "N": N, "K": K, "fmt": fmt, "flags": flags, "ms": ms, "tflops": tf,
            "cublas_ms": ms_ref, "cublas_tflops": 2.0 * M * N * K / ms_ref / 1e9}
    print(json.dumps(info), flush=True)
    report["bench"].append(info)


if __name__ == "__main__":
    t0 = time.time()
    print(torch.cuda.get_device_name(0), flush=True)
    # smallest possible: one tile, one k-block
    case(128, 256, 64, 0, 0, tag="one_tile_")
    case(128, 256, 64, 1, 0, tag="one_tile_bf16_")
    case(128, 256, 256, 0, 0, tag="multi_k_")
    case(128, 128, 64, 0, 4, tag="bn128_")
    case(256, 512, 1024, 0, 0, tag="multi_tile_")
    case(1000, 768, 1024, 0, 0, True, tag="ragged_m_bias_")
    case(514, 640, 1024, 1, 0, tag="ragged_n_")
    case(4112, 3072, 1024, 0, 0, True, tag="qkv_b16_")
    # MN-major operands (experimental: wgrad / attention backward)
    case(128, 256, 64, 0, 1, tag="bmn_one_tile_")
    case(256, 512, 512, 0, 1, tag="bmn_")
    case(128, 128, 64, 0, 2 | 4, tag="amn_one_tile_")
    case(256, 256, 512, 0, 2 | 4, tag="amn_")
    case(256, 256, 512, 0, 3 | 4, tag="abmn_")
    case(256, 256, 64, 0, 8, tag="2cta_one_tile_")
    case(1000, 768, 1024, 0, 8, True, tag="2cta_ragged_")
    case(4112, 3072, 1024, 0, 8, True, tag="2cta_qkv_")
    ok = all(c.get("bad_frac", 1.0) == 0.0 for c in report["cases"][:8])
    report["kmajor_ok"] = ok
    if ok:
        for (M, N, K) in [(65792, 3072, 1024), (65792, 1024, 1024), (65792, 4096, 1024),
                          (65792, 1024, 4096)]:
            bench(M, N, K, 0, 0)
        bench(65792, 1024, 1024, 0, 4)
        for (M, N, K) in [(65792, 3072, 1024), (65792, 1024, 1024), (65792, 4096, 1024),
                          (65792, 1024, 4096)]:
            bench(M, N, K, 0, 8)
    report["seconds"] = time.time() - t0
    with open(os.path.join(OUT, "gemm_selftest.json"), "w") as f:
        json.dump(report, f, indent=1)
    print("KMAJOR_OK" if ok else "KMAJOR_FAIL")
