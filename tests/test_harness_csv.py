"""The harness's benchmark_results.csv must stay consumable by the reference's plot_results.py
(SURVEY.md section 8 row f1).  tests/golden/harness_benchmark_results.csv is a real output of
harness/flash_attn on a B200.  The parser rules are restated from plot_results.py:16-39; when the
reference is mounted, its unmodified script is run on the file as well."""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(ROOT, "tests", "golden", "harness_benchmark_results.csv")
HEADER10 = "N,Naive(ms),Flash(ms),FlashV2(ms),FlashV3(ms),FlashV4(ms),SpeedupV1,SpeedupV2,SpeedupV3,SpeedupV4"


def parse_like_plot_results(path):
    rows, started = [], False
    for line in open(path):
        line = line.strip()
        if "N,Naive(ms)" in line:          # plot_results.py:16-18
            started = True
            continue
        if not started or not line:
            continue
        f = line.split(",")
        if len(f) < 8:                      # plot_results.py:22
            continue
        n = int(f[0])
        naive, v1, v2, v3, v4 = (float(x) for x in f[1:6])   # plot_results.py:25-32
        if naive <= 0:                      # plot_results.py:34
            continue
        rows.append((n, naive / v1, naive / v2, naive / v3, naive / v4))
    return rows


def test_csv_header_keeps_the_reference_columns_first():
    first = open(CSV).readline().strip()
    assert first.startswith(HEADER10)          # main.mm:598-606
    src = open(os.path.join(ROOT, "harness", "main.cpp")).read()
    assert HEADER10 in src.replace('"\n      "', "")


def test_csv_rows_parse_with_the_reference_rules():
    rows = parse_like_plot_results(CSV)
    assert len(rows) >= 2                      # plot_results.py divides by zero on fewer
    assert [r[0] for r in rows] == sorted(r[0] for r in rows)
    assert max(max(r[1:]) for r in rows) > 0
    # on B200 every optimised variant beats the naive kernel from N=256 up
    for n, s1, s2, s3, s4 in rows:
        if n >= 256:
            assert min(s1, s2, s3, s4) > 1.0


def test_unmodified_reference_plot_script_accepts_the_csv(tmp_path):
    script = "/root/reference/plot_results.py"
    if not os.path.exists(script):
        import pytest

        pytest.skip("reference not mounted")
    shutil.copy(CSV, tmp_path / "benchmark_results.csv")
    subprocess.check_call([sys.executable, script], cwd=tmp_path, stdout=subprocess.DEVNULL)
    svg = (tmp_path / "speedup_plot.svg").read_text()
    assert svg.lstrip().startswith("<svg") or "<svg" in svg
    assert len(re.findall(r"<(polyline|path|line|circle)", svg)) > 4


def test_tflops_plot_script_reads_the_same_csv(tmp_path):
    """harness/plot_tflops.py (the second plot of row f1: TFLOP/s and % of peak vs N) accepts the harness CSV and
    also a CSV that only has the reference's ten columns."""
    script = os.path.join(ROOT, "harness", "plot_tflops.py")
    out = tmp_path / "tflops.svg"
    subprocess.check_call([sys.executable, script, CSV, "-o", str(out)], stdout=subprocess.DEVNULL)
    svg = out.read_text()
    assert svg.lstrip().startswith("<svg") and svg.count("<polyline") >= 3
    ten = tmp_path / "benchmark_results_causal.csv"
    lines = open(CSV).read().splitlines()
    ten.write_text("\n".join(",".join(l.split(",")[:10]) for l in lines) + "\n")
    subprocess.check_call([sys.executable, script, str(ten), "-o", str(out)], stdout=subprocess.DEVNULL)
    assert out.read_text().count("<polyline") >= 3


def test_harness_config_1_reproduces_the_reference_cpu_anchors():
    """`flash_attn --config 1` (BASELINE config 1: the reference's CPU verifier alone, N=128, main.mm:128-159 loop
    order) needs no GPU; its output carries the known-answer values of SURVEY.md section 8c."""
    exe = os.path.join(ROOT, "harness", "flash_attn")
    if not os.path.exists(exe):
        import pytest

        pytest.skip("harness not built")
    out = subprocess.run([exe, "--config", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    m = re.search(r"O\[0\] = (-?[0-9.e+-]+), sum\(O\) = (-?[0-9.e+-]+)", out.stdout)
    assert m, out.stdout
    assert abs(float(m.group(1)) - (-0.0598809421)) < 1e-6 and abs(float(m.group(2)) - 0.307174703) < 1e-4
