"""Host-side logic of bench.py that needs no GPU: the per-kernel breakdown table, the FLOP constants
of BASELINE.json's workload, and the shape of the reference arm's JSON contract."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_kernel_breakdown_table():
    b = _bench()
    raw = {"gemm.gelu N=4096 K=1024": {"ms": 240.0, "launches": 480, "gflop": 264000.0},
           "ln_bwd": {"ms": 90.0, "launches": 980, "gflop": 0.0}}
    out = b.kernel_breakdown(raw, ms_total=400.0, steps=2)
    g = out["gemm.gelu N=4096 K=1024"]
    assert g["ms_per_step"] == 120.0 and g["launches_per_step"] == 240
    assert g["us_per_launch"] == 500.0 and g["share"] == 0.6 and g["tflops"] == 1100.0
    assert "tflops" not in out["ln_bwd"] and out["ln_bwd"]["share"] == 0.225
    assert out["other"]["ms_per_step"] == 35.0 and abs(out["other"]["share"] - 0.175) < 1e-9
    assert list(out)[0] == "gemm.gelu N=4096 K=1024"          # sorted by time


def test_workload_constants_match_baseline_json():
    b = _bench()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "ViT-L" in b.METRIC and "PGD" in b.METRIC
    text = json.dumps(base)
    assert "ViT-L" in text
    # 10-step ell_inf PGD, eps 4/255, step 1/255 (BASELINE.json configs[1])
    assert b.ITERS == 10 and abs(b.EPS - 4 / 255) < 1e-12 and abs(b.STEPSIZE - 1 / 255) < 1e-12
    # 330.54 GFLOP per image and PGD step (forward + input-gradient backward)
    tot = b.FWD_GF["ViT-L-14"] + b.BWD_GF["ViT-L-14"]
    assert abs(tot - 330.54) < 0.5
