"""Multi-GPU tests of ring attention (skipped with fewer than 2 GPUs): tests/ring_check.py under torchrun
(one process per GPU, every transport, forward and backward, causal and not) and tests/mgpu_check.py (one
process driving all GPUs through fa_mgpu_*).  Both scripts compare with the single-GPU kernels and exit
non-zero when a tolerance is exceeded."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    torch = pytest.importorskip("torch")
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run(cmd, timeout=600):
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, (out.stdout[-3000:] + "\n" + out.stderr[-3000:])
    return out.stdout


def torchrun(world, args, port):
    return run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "ring_check.py")] + args)


@pytest.mark.parametrize("transport", ["peer", "nccl", "gather", "auto"])
@pytest.mark.parametrize("causal", [0, 1])
def test_ring_forward_backward_multi_process(transport, causal):
    world = n_gpus()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(world, 8)
    bwd = 0 if transport == "gather" else 1  # the all-gather mode is forward only
    out = torchrun(world, ["--n-total", str(1024 * world), "--heads", "3", "--hdim", "128", "--causal", str(causal), "--check",
                           "1", "--reps", "2", "--bwd", str(bwd), "--transport", transport], 29531 + causal)
    assert '"failures": []' in out


def test_ring_hdim64_ragged_chunks():
    world = n_gpus()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(world, 8)
    # n_local = 328: chunks of 164 rows, not a multiple of the 128-row tiles
    out = torchrun(world, ["--n-total", str(328 * world), "--heads", "2", "--hdim", "64", "--causal", "1", "--check", "1",
                           "--reps", "1", "--bwd", "1", "--transport", "peer"], 29541)
    assert '"failures": []' in out


@pytest.mark.parametrize("causal", [0, 1])
def test_mgpu_single_process(causal):
    world = n_gpus()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    out = run([sys.executable, os.path.join(ROOT, "tests", "mgpu_check.py"), "--n-total", str(1024 * min(world, 8)), "--heads",
               "5", "--causal", str(causal), "--gpus", str(min(world, 8))])
    assert '"failures": []' in out
