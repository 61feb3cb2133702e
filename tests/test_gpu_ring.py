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


@pytest.mark.parametrize("nq,d,cuts", [(300, 64, (0, 128, 1000)), (256, 128, (0, 520, 777, 1200)), (130, 128, (0, 64, 128, 192, 600))])
def test_fused_merge_over_key_chunks(fa, nq, d, cuts):
    """The ring's forward step on one GPU: the keys are processed in chunks, each launch folding its
    partial into the running fp32 (O_acc, L_acc) in the kernel epilogue (first / middle / last); rows
    below half_rows stop one chunk early (as the first zig-zag chunk does).  Equals attention over the
    keys each row range has seen."""
    import ctypes
    import torch

    fn = fa.lib().fa_debug_forward_partial
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    fn.argtypes = [vp] * 7 + [i32, i32, i32, ctypes.c_float, i32, i32, i32, i32, i32, i32, vp]
    H, scale, nk = 2, float(d ** -0.5), cuts[-1]
    half = 96  # rows [0, half) see every chunk but the last one
    qb, qf = bf16(oracle.init_random(H * nq * d, 41).reshape(H, nq, d))
    kb, kf = bf16(oracle.init_random(H * nk * d, 42).reshape(H, nk, d))
    vb, vf = bf16(oracle.init_random(H * nk * d, 43).reshape(H, nk, d))
    Q = dev(qb.view(np.int16))
    O = torch.zeros((H, nq, d), dtype=torch.int16, device="cuda")
    L = torch.zeros((H, nq), device="cuda")
    o_acc = torch.full((H, nq, d), float("nan"), device="cuda")
    l_acc = torch.full((H, nq), float("nan"), device="cuda")
    n_chunks = len(cuts) - 1
    mode = lambda first, last: 0 if first and last else 1 if first else 3 if last else 2
    for c in range(n_chunks):
        k0, k1 = cuts[c], cuts[c + 1]
        K = dev(np.ascontiguousarray(kb[:, k0:k1]).view(np.int16))
        V = dev(np.ascontiguousarray(vb[:, k0:k1]).view(np.int16))
        hi = mode(c == 0, c == n_chunks - 1)
        if c < n_chunks - 1:   # both row ranges take part
            lo = mode(c == 0, c == n_chunks - 2)
            rc = fn(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), L.data_ptr(), o_acc.data_ptr(), l_acc.data_ptr(),
                    nq, k1 - k0, d, scale, H, 0, lo, hi, half, fa.BF16, None)
        else:                  # last chunk: rows [half, nq) only -- every head at once needs the head stride of the full tensor,
            rc = 0             # so go head by head with pointer offsets
            for h in range(H):
                off = (h * nq + half)
                rc |= fn(Q.data_ptr() + off * d * 2, K.data_ptr() + h * (k1 - k0) * d * 2, V.data_ptr() + h * (k1 - k0) * d * 2,
                         O.data_ptr() + off * d * 2, L.data_ptr() + off * 4, o_acc.data_ptr() + off * d * 4,
                         l_acc.data_ptr() + off * 4, nq - half, k1 - k0, d, scale, 1, 0, hi, hi, 0, fa.BF16, None)
        assert rc == 0, fa.lib().fa_last_error()
    torch.cuda.synchronize()
    got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), oracle.BF16)
    for h in range(H):
        want_lo, l_lo = rect_reference(qf[h][:half], kf[h][:cuts[-2]], vf[h][:cuts[-2]], scale)
        want_hi, l_hi = rect_reference(qf[h][half:], kf[h], vf[h], scale)
        assert np.abs(got[h][:half] - want_lo).max() <= TOL
        assert np.abs(got[h][half:] - want_hi).max() <= TOL
        assert np.abs(L[h].cpu().numpy() - np.concatenate([l_lo, l_hi])).max() <= 5e-3


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


@pytest.mark.parametrize("world,n_local,d", [(2, 328, 64), (3, 200, 128), (4, 256, 64)])
@pytest.mark.parametrize("causal", [False, True])
def test_ring_forward_and_backward_replayed_on_one_gpu(fa, world, n_local, d, causal):
    """Every rank's ring schedule replayed on one GPU with the kernels and the fused merge exactly as
    fa_ring_attention_forward / _backward drive them (same block offsets, same merge modes per zig-zag
    half, dQ accumulated over the steps, dK/dV of a chunk summed over the ranks) -- only the transport is
    left out.  Chunks of 164 / 100 rows are not multiples of the 128-row tiles."""
    import ctypes
    import torch

    fn = fa.lib().fa_debug_forward_partial
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    fn.argtypes = [vp] * 7 + [i32, i32, i32, ctypes.c_float, i32, i32, i32, i32, i32, i32, vp]
    H, n, scale = 2, world * n_local, float(d ** -0.5)
    c = n_local // 2
    bits, f = zip(*(bf16(oracle.init_random(H * n * d, s).reshape(H, n, d)) for s in (51, 52, 53, 54)))
    rows_of = lambda r: np.concatenate([np.arange(a, a + cnt) for a, cnt in fa.ring_local_rows(r, world, n_local, causal)])
    want_o, want_l = oracle.forward_batched(f[0], f[1], f[2], scale, causal)
    want_g = [np.stack(x) for x in zip(*(oracle.backward(f[0][h], f[1][h], f[2][h], f[3][h], scale, causal) for h in range(H)))]
    mode = lambda first, last: 0 if first and last else 1 if first else 3 if last else 2
    loc = lambda t, r: dev(np.ascontiguousarray(t[:, rows_of(r)]).view(np.int16))
    Kl = [loc(bits[1], r) for r in range(world)]
    Vl = [loc(bits[2], r) for r in range(world)]
    dK_sum = [torch.zeros((H, n_local, d), device="cuda") for _ in range(world)]  # per OWNER rank
    dV_sum = [torch.zeros((H, n_local, d), device="cuda") for _ in range(world)]
    hs = n_local * d
    for rank in range(world):
        rows = rows_of(rank)
        Q, dO = loc(bits[0], rank), loc(bits[3], rank)
        O = torch.zeros((H, n_local, d), dtype=torch.int16, device="cuda")
        L = torch.zeros((H, n_local), device="cuda")
        o_acc = torch.full((H, n_local, d), float("nan"), device="cuda")
        l_acc = torch.full((H, n_local), float("nan"), device="cuda")
        plan = [fa.ring_plan(rank, world, s, n_local, causal) for s in range(world)]
        for s, (src, q_off, q_rows, k_off, k_rows, bc) in enumerate(plan):
            if not causal:
                lo = hi = mode(s == 0, s == world - 1)
                half = 0
            elif q_off == 0:
                lo, hi, half = mode(s == 0, s == rank), mode(s == 0, s == world - 1), c
            else:
                lo = hi = mode(s == 0, s == world - 1)
                half = 0
            # the kernel addresses heads with the strides of the full local tensors: call head by head with offsets
            for h in range(H):
                qo, ko = (h * n_local + q_off) * d, (h * n_local + k_off) * d
                rc = fn(Q.data_ptr() + qo * 2, Kl[src].data_ptr() + ko * 2, Vl[src].data_ptr() + ko * 2, O.data_ptr() + qo * 2,
                        L.data_ptr() + (h * n_local + q_off) * 4, o_acc.data_ptr() + qo * 4, l_acc.data_ptr() + (h * n_local + q_off) * 4,
                        q_rows, k_rows, d, scale, 1, bc, lo, hi, half, fa.BF16, None)
                assert rc == 0, fa.lib().fa_last_error()
        torch.cuda.synchronize()
        got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), oracle.BF16)
        assert np.abs(got - want_o[:, rows]).max() <= TOL
        assert np.abs(L.cpu().numpy() - want_l[:, rows]).max() <= 5e-3
        # ---- backward of the same schedule ----
        delta = torch.zeros((H, n_local), device="cuda")
        fa.rowsum_delta(O, dO, delta, n_local, d, H * hs, hs, 1, H, fa.BF16)
        dQ = torch.full((H, n_local, d), float("nan"), device="cuda")
        for s, (src, q_off, q_rows, k_off, k_rows, bc) in enumerate(plan):
            tmp_k = torch.zeros((H, n_local, d), device="cuda")
            tmp_v = torch.zeros((H, n_local, d), device="cuda")
            for h in range(H):
                qo, ko = (h * n_local + q_off) * d, (h * n_local + k_off) * d
                args = (Q.data_ptr() + qo * 2, Kl[src].data_ptr() + ko * 2, Vl[src].data_ptr() + ko * 2, dO.data_ptr() + qo * 2,
                        L.data_ptr() + (h * n_local + q_off) * 4, delta.data_ptr() + (h * n_local + q_off) * 4,
                        dQ.data_ptr() + qo * 4, tmp_k.data_ptr() + ko * 4, tmp_v.data_ptr() + ko * 4)
                if bc:  # the local causal block (step 0): the square backward entry point computes delta itself
                    wsb = fa.workspace_bytes_backward(q_rows, d, 1, 1)
                    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
                    fa.flash_attention_backward(args[0], args[1], args[2], O.data_ptr() + qo * 2, args[3], args[4], args[6], args[7],
                                                args[8], q_rows, d, scale, q_rows * d, q_rows * d, True, 1, 1, fa.BF16, ws, wsb)
                else:
                    fa.flash_attention_backward_rect(*args, q_rows, k_rows, d, scale, q_rows * d, q_rows * d, k_rows * d, k_rows * d,
                                                     s > 0, 1, 1, fa.BF16)
            dK_sum[src] += tmp_k
            dV_sum[src] += tmp_v
        torch.cuda.synchronize()
        err = np.abs(dQ.cpu().numpy() - want_g[0][:, rows]).max()
        assert err <= 1e-2 * np.abs(want_g[0]).max(), ("dQ", rank, err)
    for r in range(world):
        rows = rows_of(r)
        for got, want, name in ((dK_sum[r], want_g[1], "dK"), (dV_sum[r], want_g[2], "dV")):
            err = np.abs(got.cpu().numpy() - want[:, rows]).max()
            assert err <= 1e-2 * np.abs(want).max(), (name, r, err)


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


def rect_backward_reference(q, k, v, do, scale):
    """fp64 dense gradients of non-causal attention with Nq != Nk (the formulas of kernels.metal:983-990,
    1082-1089, 1160-1169 without the mask)."""
    q, k, v, do = (x.astype(np.float64) for x in (q, k, v, do))
    s = q @ k.T * scale
    p = np.exp(s - s.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    o = p @ v
    dv = p.T @ do
    dp = do @ v.T
    ds = p * (dp - (do * o).sum(1, keepdims=True)) * scale
    return ds @ k, ds.T @ q, dv


@pytest.mark.parametrize("nq,nk,d", [(200, 456, 64), (384, 130, 128), (1, 300, 64), (513, 64, 128)])
def test_cross_attention_backward(fa, nq, nk, d):
    """flash_attention_backward_rect: gradients of attention with Nq != Nk against a dense fp64 reference."""
    import torch

    H, scale = 2, float(d ** -0.5)
    qb, qf = bf16(oracle.init_random(H * nq * d, 31).reshape(H, nq, d))
    kb, kf = bf16(oracle.init_random(H * nk * d, 32).reshape(H, nk, d))
    vb, vf = bf16(oracle.init_random(H * nk * d, 33).reshape(H, nk, d))
    gb, gf = bf16(oracle.init_random(H * nq * d, 34).reshape(H, nq, d))
    Q, K, V, dO = (dev(b.view(np.int16)) for b in (qb, kb, vb, gb))
    O = torch.zeros((H, nq, d), dtype=torch.int16, device="cuda")
    L = torch.zeros((H, nq), device="cuda")
    fa.flash_attention_v4_half_rect(Q, K, V, O, nq, nk, d, scale, H * nq * d, nq * d, H * nk * d, nk * d, L, 1, H, fa.BF16)
    delta = torch.zeros((H, nq), device="cuda")
    fa.rowsum_delta(O, dO, delta, nq, d, H * nq * d, nq * d, 1, H, fa.BF16)
    dQ = torch.full((H, nq, d), float("nan"), device="cuda")
    dK, dV = (torch.full((H, nk, d), float("nan"), device="cuda") for _ in range(2))
    fa.flash_attention_backward_rect(Q, K, V, dO, L, delta, dQ, dK, dV, nq, nk, d, scale, H * nq * d, nq * d, H * nk * d, nk * d,
                                     False, 1, H, fa.BF16)
    for h in range(H):
        want = rect_backward_reference(qf[h], kf[h], vf[h], gf[h], scale)
        for g, w in zip((dQ, dK, dV), want):
            err = np.abs(g[h].cpu().numpy() - w).max()
            assert err <= 2e-2 and err <= 1e-2 * np.abs(w).max()


def test_rectangular_backward_blocks_sum_to_full_gradients(fa):
    """The ring's building block: the keys are cut into two ragged chunks; each chunk's rectangular
    backward uses the L and delta of the FULL softmax rows; dQ accumulates over the chunks (acc_dq),
    dK/dV of a chunk are complete after its own call.  Together they equal the gradients of the whole."""
    import torch

    n, d, scale = 512, 64, 0.125
    bits, f = zip(*(bf16(oracle.init_random(n * d, s).reshape(n, d)) for s in (21, 22, 23, 24)))
    want = oracle.backward(*f, scale, False)
    Q, K, V, dO = (dev(b.view(np.int16)) for b in bits)
    O = torch.zeros((n, d), dtype=torch.int16, device="cuda")
    L = torch.zeros((n,), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, n * d, n * d, L, False, 1, 1, fa.BF16)
    delta = torch.zeros((n,), device="cuda")
    fa.rowsum_delta(O, dO, delta, n, d, n * d, n * d, 1, 1, fa.BF16)
    dQ, dK, dV = (torch.full((n, d), float("nan"), device="cuda") for _ in range(3))
    cut = 200  # chunks of 200 and 312 keys: neither is a multiple of the 128-key tiles
    for k0, k1, acc in ((0, cut, False), (cut, n, True)):
        kp, vp = K.data_ptr() + k0 * d * 2, V.data_ptr() + k0 * d * 2
        fa.flash_attention_backward_rect(Q, kp, vp, dO, L, delta, dQ, dK.data_ptr() + k0 * d * 4, dV.data_ptr() + k0 * d * 4,
                                         n, k1 - k0, d, scale, n * d, n * d, (k1 - k0) * d, (k1 - k0) * d, acc, 1, 1, fa.BF16)
    for g, w in zip((dQ, dK, dV), want):
        assert np.abs(g.cpu().numpy() - w).max() <= 1e-2 * np.abs(w).max()


def test_cluster_of_eight(fa):
    """The key split also works with clusters of 8 CTAs (not the default: measured slower than 4)."""
    import ctypes

    setter = fa.lib().fa_debug_set_fwd_split_max
    setter.argtypes = [ctypes.c_int]
    setter(8)
    try:
        for nq, nk, d in [(200, 5000, 64), (130, 4500, 128), (700, 2100, 128)]:
            test_rectangular_forward(fa, nq, nk, d)
    finally:
        setter(4)
