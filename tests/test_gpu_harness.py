"""GPU test of the C++ harness (harness/flash_attn, the main.mm replacement): --quick run must
print the reference's PASSED lines, write a CSV the reference's plot script can parse, and exit 0."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "harness", "flash_attn")


def test_harness_quick_run(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "harness")])
    csv = tmp_path / "benchmark_results.csv"
    out = subprocess.run([BIN, "--quick", "--max-n", "1024", "--csv", str(csv)], capture_output=True, text=True, timeout=600)
    text = out.stdout
    assert out.returncode == 0, text[-2000:] + out.stderr[-2000:]
    for line in ("Naive Kernel PASSED", "V1 PASSED", "V2 PASSED", "V3 PASSED", "V4 PASSED", "CAUSAL PASSED",
                 "Backward Pass PASSED", "--- Benchmarking ---", "--- High Occupancy Benchmark (B=16, H=8) ---"):
        assert line in text, line
    assert "FAILED" not in text
    rows = [l for l in csv.read_text().splitlines() if l and l[0].isdigit()]
    assert [int(r.split(",")[0]) for r in rows] == [128, 256, 512, 1024]
    assert all(len(r.split(",")) >= 10 for r in rows)
