"""CPU tests of the ring-attention host logic (no GPU): the schedule the library's ring driver
executes (fa_ring_plan / fa_ring_local_rows, the same functions csrc/ring.cu calls) is replayed
with the oracle as the per-block attention and the library's merge rule in numpy, and must
reproduce full attention.  World sizes 1..8 are simulated in-process; world_size 2 is also run as
two real processes that rotate K/V with torch.distributed (gloo) send/recv."""
import os
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = 32
SCALE = float(1.0 / np.sqrt(D))


@pytest.fixture(scope="module")
def fa():
    import flash_attention_metal_b200 as fa

    if not os.path.exists(fa.LIB_PATH):
        fa.build()
    return fa


def block_attention(q, k, v, causal):
    """Oracle attention of a rectangular block -> (O, L).  Rectangular = pad-free: the oracle's
    causal flag is only used for square blocks, exactly as the ring driver does."""
    if causal:
        assert q.shape[0] == k.shape[0]
        return oracle.forward(q, k, v, SCALE, True)
    s = (q.astype(np.float64) @ k.astype(np.float64).T) * SCALE
    m = s.max(1, keepdims=True)
    p = np.exp(s - m)
    l = p.sum(1, keepdims=True)
    return (p @ v.astype(np.float64) / l).astype(np.float32), (m[:, 0] + np.log(l[:, 0])).astype(np.float32)


def merge(o_acc, l_acc, o_part, l_part):
    """kernels.metal:784-791 rule, as ring_merge_kernel applies it."""
    l_new = np.logaddexp(l_acc, l_part)
    return o_acc * np.exp(l_acc - l_new)[:, None] + o_part * np.exp(l_part - l_new)[:, None], l_new


def local_rows(fa, rank, world, n_local, causal):
    idx = []
    for first, rows in fa.ring_local_rows(rank, world, n_local, causal):
        idx.extend(range(first, first + rows))
    return np.array(idx)


def run_rank(fa, rank, world, n_local, causal, q, kv_of):
    """Replay the schedule for one rank; kv_of(src) returns the (K, V) chunk rank `src` owns."""
    o_acc = np.zeros((n_local, D), np.float64)
    l_acc = np.full(n_local, -np.inf)
    for step in range(world):
        src, q_off, q_rows, k_off, k_rows, bc = fa.ring_plan(rank, world, step, n_local, causal)
        assert src == (rank - step) % world
        k, v = kv_of(src)
        o, l = block_attention(q[q_off:q_off + q_rows], k[k_off:k_off + k_rows], v[k_off:k_off + k_rows], bool(bc))
        sl = slice(q_off, q_off + q_rows)
        o_acc[sl], l_acc[sl] = merge(o_acc[sl], l_acc[sl], o, l)
    return o_acc, l_acc


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("causal", [False, True])
def test_schedule_reproduces_full_attention(fa, world, causal):
    n_local = 12
    n = world * n_local
    q, k, v = (oracle.init_random(n * D, 50 + i).reshape(n, D) for i in range(3))
    want, want_l = oracle.forward(q, k, v, SCALE, causal)
    seen = np.zeros(n, bool)
    for rank in range(world):
        rows = local_rows(fa, rank, world, n_local, causal)
        assert not seen[rows].any()
        seen[rows] = True
        kv_of = lambda src: (k[local_rows(fa, src, world, n_local, causal)], v[local_rows(fa, src, world, n_local, causal)])
        o, l = run_rank(fa, rank, world, n_local, causal, q[rows], kv_of)
        assert np.abs(o - want[rows]).max() < 2e-6
        assert np.abs(l - want_l[rows]).max() < 2e-6
    assert seen.all()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_causal_work_is_balanced_and_masked_blocks_are_skipped(fa, world):
    n_local = 16
    c = n_local // 2
    for rank in range(world):
        for step in range(world):
            src, q_off, q_rows, k_off, k_rows, bc = fa.ring_plan(rank, world, step, n_local, True)
            work = q_rows * k_rows * (0.5 if bc else 1.0)
            # two c x c blocks of unmasked work at every step, on every rank (zig-zag)
            assert work == 2 * c * c
            if step == 0:
                assert (q_rows, k_rows, bc) == (n_local, n_local, 1)
            elif src < rank:
                assert (q_off, q_rows, k_off, k_rows, bc) == (0, n_local, 0, c, 0)
            else:
                assert (q_off, q_rows, k_off, k_rows, bc) == (c, c, 0, n_local, 0)


def test_plan_rejects_bad_arguments(fa):
    with pytest.raises(fa.FlashAttnError):
        fa.ring_plan(2, 2, 0, 16, True)
    with pytest.raises(fa.FlashAttnError):
        fa.ring_plan(0, 2, 0, 15, True)  # causal zig-zag needs an even n_local


def _gloo_worker(rank, world, port, causal, ret):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import flash_attention_metal_b200 as fa

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = 16
    n = world * n_local
    q, k, v = (oracle.init_random(n * D, 70 + i).reshape(n, D) for i in range(3))
    rows = local_rows(fa, rank, world, n_local, causal)
    ql, cur_k, cur_v = q[rows], torch.from_numpy(k[rows].copy()), torch.from_numpy(v[rows].copy())
    o_acc, l_acc = np.zeros((n_local, D)), np.full(n_local, -np.inf)
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    for step in range(world):
        reqs = []
        if step + 1 < world:  # rotate K/V exactly as the library does: send to next, receive from previous
            rk, rv = torch.empty_like(cur_k), torch.empty_like(cur_v)
            reqs = [dist.isend(cur_k, nxt), dist.isend(cur_v, nxt), dist.irecv(rk, prv), dist.irecv(rv, prv)]
        src, q_off, q_rows, k_off, k_rows, bc = fa.ring_plan(rank, world, step, n_local, causal)
        kk, vv = cur_k.numpy(), cur_v.numpy()
        o, l = block_attention(ql[q_off:q_off + q_rows], kk[k_off:k_off + k_rows], vv[k_off:k_off + k_rows], bool(bc))
        sl = slice(q_off, q_off + q_rows)
        o_acc[sl], l_acc[sl] = merge(o_acc[sl], l_acc[sl], o, l)
        for r in reqs:
            r.wait()
        if step + 1 < world:
            cur_k, cur_v = rk, rv
    want, want_l = oracle.forward(q, k, v, SCALE, causal)
    ret[rank] = float(max(np.abs(o_acc - want[rows]).max(), np.abs(l_acc - want_l[rows]).max()))
    dist.destroy_process_group()


@pytest.mark.parametrize("causal", [False, True])
def test_two_process_gloo_ring(fa, causal):
    mp = pytest.importorskip("torch.multiprocessing")
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, causal, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert max(ret.values()) < 2e-6


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("causal", [False, True])
def test_merge_plan_gives_every_row_range_one_first_and_one_last(fa, world, causal):
    """fa_ring_merge_plan (the function the ring driver asks how a step's launch folds its partial into the running
    (O, L) inside the kernel epilogue): over the steps of a call every local row must be written exactly as
    'the only partial' once, or as first, middle ..., last in that order -- and rows a step's block does not
    cover must not be touched by it."""
    n_local = 64
    for rank in range(world):
        state = np.zeros(n_local, dtype=np.int64)  # 0 untouched, 1 running pair started, 2 finished
        for step in range(world):
            src, q_off, q_rows, k_off, k_rows, bc = fa.ring_plan(rank, world, step, n_local, causal)
            lo, hi, half = fa.ring_merge_plan(rank, world, step, n_local, causal)
            for r in range(q_rows):
                mode = lo if r < half else hi
                row = q_off + r
                if mode == 0:      # the only partial of this row
                    assert state[row] == 0
                    state[row] = 2
                elif mode == 1:    # first: starts the running pair
                    assert state[row] == 0
                    state[row] = 1
                elif mode == 2:    # middle: needs a running pair, leaves one
                    assert state[row] == 1
                else:              # last: needs a running pair, writes the final O and L
                    assert mode == 3 and state[row] == 1
                    state[row] = 2
        assert (state == 2).all(), (rank, state)


def test_ring_api_argument_errors_without_gpu(fa):
    with pytest.raises(fa.FlashAttnError, match="bad ring plan"):
        fa.ring_merge_plan(2, 2, 0, 64, True)
    with pytest.raises(fa.FlashAttnError, match="bad ring plan"):
        fa.ring_plan(0, 2, 0, 63, True)  # causal needs an even n_local
    with pytest.raises(fa.FlashAttnError, match="unknown backward algorithm"):
        fa.set_backward_algorithm(7)
    assert fa.get_backward_algorithm() == fa.BWD_TWO_KERNEL
    L = fa.lib()
    # forward workspace: the all-gather transport needs room for every rank's K/V, the others two slots
    small = L.fa_ring_workspace_bytes_ex(8, fa.TRANSPORT_PEER, 1024, 128, 4, fa.BF16)
    assert small == L.fa_ring_workspace_bytes_ex(8, fa.TRANSPORT_NCCL, 1024, 128, 4, fa.BF16) == L.fa_ring_workspace_bytes(1024, 128, 4, fa.BF16)
    assert L.fa_ring_workspace_bytes_ex(8, fa.TRANSPORT_NCCL_GATHER, 1024, 128, 4, fa.BF16) == small + 6 * 2 * (4 * 1024 * 128 * 2)
    assert L.fa_ring_workspace_bytes_backward(1024, 128, 4, fa.BF16) > small
    # the backward workspace covers the fused kernel's ordering counters and delta
    assert fa.workspace_bytes_backward(1000, 128, 2, 3) >= 2 * 3 * 1000 * 4 + 2 * 3 * 8 * 4


def fold_like_the_kernel_epilogue(mode, o_part, l_part, o_acc, l_acc):
    """What fwd_tc_kernel's epilogue does with a launch's partial (csrc/fwd_tc.cu store_rows): returns
    (final O or None, final L or None, new o_acc, new l_acc)."""
    if mode == 0:                                   # the only partial: straight to O and L
        return o_part, l_part, o_acc, l_acc
    if mode == 1:                                   # first: the running pair starts
        return None, None, o_part.copy(), l_part.copy()
    o_new, l_new = merge(o_acc, l_acc, o_part, l_part)
    if mode == 2:                                   # middle: folded into the running pair
        return None, None, o_new, l_new
    return o_new, l_new, o_acc, l_acc               # last: folded and written out


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("causal", [False, True])
def test_schedule_with_fused_merge_modes_reproduces_full_attention(fa, world, causal):
    """The forward exactly as the ring driver runs it since round 2: one launch per step whose epilogue folds the
    partial into a running (O, L) pair according to fa_ring_merge_plan, per zig-zag half -- replayed with the
    oracle as the per-block kernel.  The running pair starts as NaN, as uninitialised workspace might: a mode
    that reads it before a 'first' wrote it poisons the result."""
    n_local = 16
    n = world * n_local
    q, k, v = (oracle.init_random(n * D, 70 + i).reshape(n, D) for i in range(3))
    want, want_l = oracle.forward(q, k, v, SCALE, causal)
    for rank in range(world):
        rows = local_rows(fa, rank, world, n_local, causal)
        ql = q[rows]
        o_acc = np.full((n_local, D), np.nan)
        l_acc = np.full(n_local, np.nan)
        o_out = np.full((n_local, D), np.nan)
        l_out = np.full(n_local, np.nan)
        for step in range(world):
            src, q_off, q_rows, k_off, k_rows, bc = fa.ring_plan(rank, world, step, n_local, causal)
            lo, hi, half = fa.ring_merge_plan(rank, world, step, n_local, causal)
            srows = local_rows(fa, src, world, n_local, causal)
            o, l = block_attention(ql[q_off:q_off + q_rows], k[srows][k_off:k_off + k_rows], v[srows][k_off:k_off + k_rows], bool(bc))
            for r0, r1, mode in ((0, min(half, q_rows), lo), (min(half, q_rows), q_rows, hi)):
                if r1 <= r0:
                    continue
                sl = slice(q_off + r0, q_off + r1)
                fo, fl, o_acc[sl], l_acc[sl] = fold_like_the_kernel_epilogue(mode, o[r0:r1].astype(np.float64), l[r0:r1].astype(np.float64),
                                                                           o_acc[sl], l_acc[sl])
                if fo is not None:
                    assert np.isnan(o_out[sl]).all()   # every row is written out exactly once
                    o_out[sl], l_out[sl] = fo, fl
        assert np.abs(o_out - want[rows]).max() < 2e-6
        assert np.abs(l_out - want_l[rows]).max() < 2e-6
