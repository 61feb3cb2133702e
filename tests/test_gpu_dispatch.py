"""The CTA dispatch order (csrc/sched.cuh: heads in L2-sized groups, heaviest blocks first) is a pure
scheduling choice: every output must be bit-identical whatever the group size."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("d", [64, 128])
def test_outputs_do_not_depend_on_the_dispatch_order(d):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    digests = {}
    # 0 MB -> one head per group (per-head order); 1 MB -> groups of 1-3 heads with padding CTAs in the
    # last group; default 48 MB and 4096 MB -> all ten heads in one group
    for mb in ("0", "1", "", "4096"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dispatch_digest.py"), str(d)] + ([mb] if mb else []),
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
        digests[mb] = [l for l in out.stdout.splitlines() if l.startswith("digest")][0]
    assert len(set(digests.values())) == 1, digests
