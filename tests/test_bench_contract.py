"""bench.py pieces that run without a GPU: the peak-file parser and the reference arm's JSON line."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    return m


def test_peak_file_parser_accepts_flat_nested_and_raw_units():
    m = _bench()
    want = (1386.8, 6445.3)
    assert m.parse_peaks({"hbm_gbs": 6445.3, "bf16_tflops_sustained": 1386.8, "bf16_tflops_burst": 1642.7}) == want
    assert m.parse_peaks({"hbm": {"copy_GBps": 6445.3}, "bf16": {"dense_tflops": {"burst": 1642.7, "sustained": 1386.8}}}) == want
    tf, gbs = m.parse_peaks({"hbm_copy_bytes_per_s": 6.4453e12, "cublas_bf16_flops": 1.3868e15})
    assert abs(tf - 1386.8) < 1e-6 and abs(gbs - 6445.3) < 1e-6
    assert m.parse_peaks({"note": "nothing numeric", "n_sms": 148}) == (None, None)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-points", "300"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and "workload" in d["config"]
