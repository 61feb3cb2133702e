"""GPU tests (single device) of the pieces ring attention is built from: the rectangular forward,
the ring driver at world = 1 (NCCL communicator of one rank: exercises the block launch + merge
kernels end to end), and an emulation of worlds 2..4 on one GPU that replays fa_ring_plan with the
device kernel per block.  The real multi-GPU path is exercised by tests/ring_check.py under torchrun."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def fa():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import flash_attention_metal_b200 as fa

    fa.lib()
    return fa


def dev(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def bf16(x):
    b = oracle.to_half_bits(x, oracle.BF16)
    return b, oracle.from_half_bits(b, oracle.BF16)


def rect_reference(q, k, v, scale):
    s = (q.astype(np.float64) @ k.astype(np.float64).T) * scale
    m = s.max(1, keepdims=True)
    p = np.exp(s - m)
    l = p.sum(1, keepdims=True)
    return p @ v.astype(np.float64) / l, m[:, 0] + np.log(l[:, 0])


# the last three shapes launch fewer CTAs than half the SMs and many key tiles: the forward splits the
# keys of each row block over a cluster of CTAs (4 by default; test_cluster_of_eight forces 8) and
# merges the parts through distributed shared memory
@pytest.mark.parametrize("nq,nk,d", [(128, 256, 64), (300, 77, 64), (64, 1000, 128), (513, 129, 128), (1, 5, 64),
                                     (200, 5000, 64), (130, 4500, 128), (700, 2100, 128)])
def test_rectangular_forward(fa, nq, nk, d):
    import torch

    scale = float(d ** -0.5)
    H = 3
    qb, qf = bf16(oracle.init_random(H * nq * d, 1).reshape(H, nq, d))
    kb, kf = bf16(oracle.init_random(H * nk * d, 2).reshape(H, nk, d))
    vb, vf = bf16(oracle.init_random(H * nk * d, 3).reshape(H, nk, d))
    O = torch.zeros((H, nq, d), dtype=torch.int16, device="cuda")
    L = torch.zeros((H, nq), device="cuda")
    fa.flash_attention_v4_half_rect(dev(qb.view(np.int16)), dev(kb.view(np.int16)), dev(vb.view(np.int16)), O, nq, nk, d,
                                    scale, H * nq * d, nq * d, H * nk * d, nk * d, L, 1, H, fa.BF16)
    got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), oracle.BF16)
    for h in range(H):
        want, want_l = rect_reference(qf[h], kf[h], vf[h], scale)
        assert np.abs(got[h] - want).max() <= TOL
        assert np.abs(L[h].cpu().numpy() - want_l).max() <= 5e-3


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("n,d", [(256, 64), (1000, 128)])
def test_ring_world_one_matches_oracle(fa, causal, n, d):
    import torch

    H, scale = 2, float(d ** -0.5)
    bits, f = zip(*(bf16(oracle.init_random(H * n * d, s).reshape(H, n, d)) for s in (4, 5, 6)))
    ring = fa.Ring(fa.ring_unique_id(), 0, 1, 0)
    try:
        O = torch.zeros((H, n, d), dtype=torch.int16, device="cuda")
        L = torch.zeros((H, n), device="cuda")
        wsb = ring.workspace_bytes(n, d, H, fa.BF16)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        ring.forward(*(dev(b.view(np.int16)) for b in bits), O, L, n, d, H, scale, causal, fa.BF16, ws, wsb)
        torch.cuda.synchronize()
    finally:
        ring.close()
    want, want_l = oracle.forward_batched(*f, scale, causal)
    got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), oracle.BF16)
    assert np.abs(got - want).max() <= TOL
    assert np.abs(L.cpu().numpy() - want_l).max() <= 5e-3


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("causal", [False, True])
def test_ring_schedule_emulated_on_one_gpu(fa, world, causal):
    """Replay every rank's schedule with the device kernel per block and the merge rule on the host."""
    import torch

    n_local, d = 256, 64
    n, scale = world * n_local, 0.125
    bits, f = zip(*(bf16(oracle.init_random(n * d, s).reshape(n, d)) for s in (7, 8, 9)))
    want, want_l = oracle.forward(*f, scale, causal)
    rows_of = lambda r: np.concatenate([np.arange(a, a + c) for a, c in fa.ring_local_rows(r, world, n_local, causal)])
    for rank in range(world):
        rows = rows_of(rank)
        Q = dev(bits[0][rows].view(np.int16))
        o_acc = np.zeros((n_local, d))
        l_acc = np.full(n_local, -np.inf)
        for step in range(world):
            src, q_off, q_rows, k_off, k_rows, bc = fa.ring_plan(rank, world, step, n_local, causal)
            srows = rows_of(src)
            K = dev(bits[1][srows].view(np.int16))
            V = dev(bits[2][srows].view(np.int16))
            O = torch.zeros((q_rows, d), dtype=torch.int16, device="cuda")
            L = torch.zeros((q_rows,), device="cuda")
            q_ptr = Q.data_ptr() + q_off * d * 2
            k_ptr, v_ptr = K.data_ptr() + k_off * d * 2, V.data_ptr() + k_off * d * 2
            if bc:
                fa.flash_attention_v4_half(q_ptr, k_ptr, v_ptr, O, q_rows, d, scale, q_rows * d, q_rows * d, L, True, 1, 1, fa.BF16)
            else:
                fa.flash_attention_v4_half_rect(q_ptr, k_ptr, v_ptr, O, q_rows, k_rows, d, scale, q_rows * d, q_rows * d,
                                                k_rows * d, k_rows * d, L, 1, 1, fa.BF16)
            o = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), oracle.BF16).astype(np.float64)
            l = L.cpu().numpy().astype(np.float64)
            sl = slice(q_off, q_off + q_rows)
            l_new = np.logaddexp(l_acc[sl], l)
            o_acc[sl] = o_acc[sl] * np.exp(l_acc[sl] - l_new)[:, None] + o * np.exp(l - l_new)[:, None]
            l_acc[sl] = l_new
        assert np.abs(o_acc - want[rows]).max() <= TOL
        assert np.abs(l_acc - want_l[rows]).max() <= 5e-3


@pytest.mark.parametrize("causal", [False, True])
def test_ring_backward_world_one_matches_oracle(fa, causal):
    """fa_ring_attention_backward with a one-rank communicator: delta + rectangular backward kernels."""
    import torch

    H, n, d = 2, 384, 64
    scale = float(d ** -0.5)
    bits, f = zip(*(bf16(oracle.init_random(H * n * d, s).reshape(H, n, d)) for s in (11, 12, 13, 14)))
    Q, K, V, dO = (dev(b.view(np.int16)) for b in bits)
    ring = fa.Ring(fa.ring_unique_id(), 0, 1, 0)
    try:
        O = torch.zeros((H, n, d), dtype=torch.int16, device="cuda")
        L = torch.zeros((H, n), device="cuda")
        wsb = ring.workspace_bytes(n, d, H, fa.BF16)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        ring.forward(Q, K, V, O, L, n, d, H, scale, causal, fa.BF16, ws, wsb)
        grads = [torch.full((H, n, d), float("nan"), device="cuda") for _ in range(3)]
        bwsb = ring.workspace_bytes_backward(n, d, H, fa.BF16)
        bws = torch.empty(bwsb, dtype=torch.uint8, device="cuda")
        ring.backward(Q, K, V, O, dO, L, *grads, n, d, H, scale, causal, fa.BF16, bws, bwsb)
        torch.cuda.synchronize()
    finally:
        ring.close()
    for h in range(H):
        want = oracle.backward(f[0][h], f[1][h], f[2][h], f[3][h], scale, causal)
        for g, w in zip(grads, want):
            err = np.abs(g[h].cpu().numpy() - w).max()
            assert err <= 2e-2 and err <= 1e-2 * np.abs(w).max()


def test_rectangular_backward_blocks_sum_to_full_gradients(fa):
    """The ring's building block: gradients of attention over two key chunks, each computed by the
    rectangular backward from the full-row L and delta, add up to the gradients of the whole."""
    import torch

    n, d, scale = 512, 64, 0.125
    ring = fa.Ring(fa.ring_unique_id(), 0, 1, 0)
    ring.close()  # only here to make sure NCCL loading does not interfere; kernels are called directly
    bits, f = zip(*(bf16(oracle.init_random(n * d, s).reshape(n, d)) for s in (21, 22, 23, 24)))
    want = oracle.backward(*f, scale, False)
    Q, K, V, dO = (dev(b.view(np.int16)) for b in bits)
    O = torch.zeros((n, d), dtype=torch.int16, device="cuda")
    L = torch.zeros((n,), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, n * d, n * d, L, False, 1, 1, fa.BF16)
    got = [torch.zeros((n, d), device="cuda") for _ in range(3)]
    wsb = fa.workspace_bytes_backward(n, d, 1, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    fa.flash_attention_backward(Q, K, V, O, dO, L, *got, n, d, scale, n * d, n * d, False, 1, 1, fa.BF16, ws, wsb)
    for g, w in zip(got, want):
        assert np.abs(g.cpu().numpy() - w).max() <= 1e-2 * np.abs(w).max()


def test_cluster_of_eight():
    """The key split also works with clusters of 8 CTAs (not the default: measured slower than 4)."""
    import subprocess, sys, os
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FA_FWD_SPLIT_MAX="8")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_ring.py"), "-m", "gpu", "-q",
                          "-k", "test_rectangular_forward"], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-1000:]
