"""GPU parity tests for flash_attention_backward (C ABI) against the oracle's backward
(main.mm:1091-1179 formulas, fp16/bf16 decoded correctly).  Bars: max-abs <= 2e-2
(BASELINE.json, 16-bit) and, because gradients are small numbers, also <= 1 % (bf16) /
0.2 % (fp16) of the largest reference gradient.  Gradients must be bit-identical from
run to run (no float atomics, unlike kernels.metal:1227, 1243)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

TOL_ABS = 2e-2
TOL_REL = {oracle.FP16: 2e-3, oracle.BF16: 1e-2}


@pytest.fixture(scope="module", params=["fused", "two_kernel"])
def fa(request):
    """Every test of this file runs on both implementations of the backward: the two-kernel form
    (csrc/bwd_tc.cu, the default) and the fused five-GEMM kernel (csrc/bwd_fused.cu), selected with
    fa_set_backward_algorithm."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import flash_attention_metal_b200 as fa

    fa.set_backward_algorithm(fa.BWD_FUSED if request.param == "fused" else fa.BWD_TWO_KERNEL)
    yield fa
    fa.set_backward_algorithm(fa.BWD_TWO_KERNEL)


def dev(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def run_backward(fa, bits, n, d, scale, causal, dtype, B=1, H=1):
    import torch

    qb, kb, vb, dob = bits
    Q, K, V, dO = (dev(t.view(np.int16)) for t in (qb, kb, vb, dob))
    shape = (B, H, n, d)
    O = torch.zeros(shape, dtype=torch.int16, device="cuda")
    L = torch.zeros((B, H, n), device="cuda")
    hs, bs = n * d, H * n * d
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, bs, hs, L, causal, B, H, dtype)
    grads = [torch.full(shape, float("nan"), device="cuda") for _ in range(3)]
    wsb = fa.workspace_bytes_backward(n, d, B, H)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    fa.flash_attention_backward(Q, K, V, O, dO, L, *grads, n, d, scale, bs, hs, causal, B, H, dtype, ws, wsb)
    torch.cuda.synchronize()
    return [g.cpu().numpy() for g in grads]


def make_bits(n, d, dtype, heads=(), seeds=(1, 2, 3, 4), mul=1.0):
    size = int(np.prod(heads, dtype=np.int64)) * n * d if heads else n * d
    return [oracle.to_half_bits(oracle.init_random(size, s).reshape(*heads, n, d) * np.float32(mul), dtype) for s in seeds]


def check(got, want, dtype):
    for g, w, name in zip(got, want, ("dQ", "dK", "dV")):
        assert np.isfinite(g).all(), name
        err = np.abs(g - w).max()
        assert err <= TOL_ABS, (name, err)
        assert err <= TOL_REL[dtype] * np.abs(w).max(), (name, err, np.abs(w).max())


@pytest.mark.parametrize("dtype", [oracle.FP16, oracle.BF16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("n,d", [(128, 64), (64, 64), (1, 64), (100, 64), (257, 64), (1024, 64),
                                 (128, 128), (200, 128), (640, 128), (1024, 128)])
def test_backward_matches_oracle(fa, dtype, causal, n, d):
    scale = float(1.0 / np.sqrt(d))
    bits = make_bits(n, d, dtype)
    f = [oracle.from_half_bits(t, dtype) for t in bits]
    want = oracle.backward(*f, scale, causal)
    got = run_backward(fa, bits, n, d, scale, causal, dtype)
    check([g[0, 0] for g in got], want, dtype)


@pytest.mark.parametrize("dtype", [oracle.FP16, oracle.BF16])
def test_backward_reference_harness_inputs(fa, golden, dtype):
    """The reference's own backward check (main.mm:946-967, 1087-1195): N=128, D=64,
    Q = K = V = dO = 0.01 * initRandom.  fp16 is compared with the frozen output of the
    reference's CPU loops (decoded correctly).  The bar is relative because the signal is
    ~1e-7 (the reference's absolute 1e-1, main.mm:1191, is vacuous).  With these inputs dS is
    ~1e-6, i.e. *subnormal* in fp16 (spacing 6e-8), so an fp16 operand path -- the reference's
    kernel included -- cannot do better than a few percent here; bf16 has the range and meets 1 %."""
    n, d = 128, 64
    if dtype == oracle.FP16:
        qb = golden["bwd128_qbits"]
        want = [golden[k] for k in ("bwd128_dq", "bwd128_dk", "bwd128_dv")]
        tol = 6e-2
    else:
        qb = oracle.to_half_bits(oracle.init_random(n * d).reshape(n, d) * np.float32(0.01), dtype)
        qf = oracle.from_half_bits(qb, dtype)
        want = oracle.backward(qf, qf, qf, qf, 0.125, False)
        tol = 1e-2
    got = run_backward(fa, (qb, qb, qb, qb), n, d, 0.125, False, dtype)
    for g, w in zip(got, want):
        assert np.abs(g[0, 0] - w).max() <= tol * np.abs(w).max() + 1e-12


@pytest.mark.parametrize("causal", [False, True])
def test_backward_batched_strided(fa, causal):
    B, H, n, d = 2, 3, 300, 64
    dtype = oracle.BF16
    bits = make_bits(n, d, dtype, heads=(B, H), seeds=(5, 6, 7, 8))
    f = [oracle.from_half_bits(t, dtype) for t in bits]
    got = run_backward(fa, bits, n, d, 0.125, causal, dtype, B, H)
    for b in range(B):
        for h in range(H):
            want = oracle.backward(*(t[b, h] for t in f), 0.125, causal)
            check([g[b, h] for g in got], want, dtype)


def test_backward_is_deterministic(fa):
    n, d, dtype = 1024, 128, oracle.BF16
    bits = make_bits(n, d, dtype, heads=(1, 4))
    a = run_backward(fa, bits, n, d, 0.0884, True, dtype, 1, 4)
    b = run_backward(fa, bits, n, d, 0.0884, True, dtype, 1, 4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_backward_causal_structure(fa):
    """Causal: dK/dV of the last key depend on the last query only; dQ of row 0 sees key 0 only,
    where P = 1 and dS = P (dP - D) = 0 exactly."""
    n, d, dtype = 384, 64, oracle.BF16
    bits = make_bits(n, d, dtype)
    dq, dk, dv = (g[0, 0] for g in run_backward(fa, bits, n, d, 0.125, True, dtype))
    assert np.abs(dq[0]).max() <= 1e-6
    dof = oracle.from_half_bits(bits[3], dtype)
    # last key is seen by the last query only: dV[n-1] = P[n-1, n-1] * dO[n-1]
    ratio = dv[-1] / dof[-1]
    assert np.ptp(ratio) <= 2e-2 * abs(ratio.mean()) and 0 < ratio.mean() <= 1.0


def test_backward_workspace_and_argument_errors(fa):
    import torch

    n, d = 128, 64
    t = torch.zeros((n, d), dtype=torch.int16, device="cuda")
    g = torch.zeros((n, d), device="cuda")
    L = torch.zeros((n,), device="cuda")
    with pytest.raises(fa.FlashAttnError, match="workspace"):
        fa.flash_attention_backward(t, t, t, t, t, L, g, g, g, n, d, 0.125, n * d, n * d, False, 1, 1, fa.BF16, None, 0)
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")
    with pytest.raises(fa.FlashAttnError, match="workspace"):
        fa.flash_attention_backward(t, t, t, t, t, L, g, g, g, n, d, 0.125, n * d, n * d, False, 1, 1, fa.BF16, ws, 16)


@pytest.mark.parametrize("causal", [True, False])
def test_backward_flagship_shape_sampled(fa, causal):
    """BASELINE config 3 (bf16, H=16, N=16384, d=128): sampled rows of dQ and sampled keys of
    dK/dV recomputed in fp64 from the same inputs."""
    import torch

    B, H, n, d = 1, 16, 16384, 128
    scale = float(1.0 / np.sqrt(d))
    gen = torch.Generator(device="cuda").manual_seed(7)
    Q, K, V, dO = (torch.rand((B, H, n, d), device="cuda", generator=gen).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(4))
    O = torch.empty_like(Q)
    L = torch.empty((B, H, n), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, fa.BF16)
    dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
    wsb = fa.workspace_bytes_backward(n, d, B, H)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, H * n * d, n * d, causal, B, H, fa.BF16, ws, wsb)
    torch.cuda.synchronize()
    for h in (0, 9):
        q, k, v, do = (t[0, h].double() for t in (Q, K, V, dO))
        s = (q @ k.T) * scale
        if causal:
            s = s.masked_fill(torch.ones(n, n, dtype=torch.bool, device="cuda").triu(1), float("-inf"))
        p = torch.softmax(s, dim=1)
        dp = do @ v.T
        ds = p * (dp - (dp * p).sum(1, keepdim=True)) * scale
        for got, want in ((dQ[0, h], ds @ k), (dK[0, h], ds.T @ q), (dV[0, h], p.T @ do)):
            err = (got.double() - want).abs().max().item()
            assert err <= TOL_ABS and err <= TOL_REL[oracle.BF16] * want.abs().max().item()
        del s, p, dp, ds


def test_host_buffer_fwd_bwd_entry_point(fa):
    """fa_host_attention_fwd_bwd_half: host tensors in, host tensors out, pipelined over head groups;
    must equal the device-pointer path bit for bit."""
    import torch

    B, H, n, d, dtype = 2, 5, 512, 64, oracle.BF16
    bits = make_bits(n, d, dtype, heads=(B, H), seeds=(11, 12, 13, 14))
    want = run_backward(fa, bits, n, d, 0.125, True, dtype, B, H)
    o = np.empty((B, H, n, d), np.uint16)
    l = np.empty((B, H, n), np.float32)
    grads = [np.empty((B, H, n, d), np.float32) for _ in range(3)]
    fa.host_attention_fwd_bwd_half(bits[0], bits[1], bits[2], bits[3], o, l, *grads, n, d, 0.125, True, B, H, dtype)
    for g, w in zip(grads, want):
        assert np.array_equal(g, w)
    f = [oracle.from_half_bits(t, dtype) for t in bits]
    ow, lw = oracle.forward_batched(f[0], f[1], f[2], 0.125, True)
    assert np.abs(oracle.from_half_bits(o, dtype) - ow).max() <= 2e-2
    assert np.abs(l - lw).max() <= 5e-3
    fa.host_release()


@pytest.mark.parametrize("gdt", [oracle.BF16, oracle.FP16])
def test_host_buffer_call_with_16_bit_gradients(fa, gdt):
    """fa_host_attention_fwd_bwd_half_ex with grad_dtype: the gradients leave the device rounded to 16 bits
    (round to nearest even) -- bit for bit the fp32 gradients of the ordinary call, rounded."""
    B, H, n, d, dtype = 1, 3, 384, 64, oracle.BF16
    bits = make_bits(n, d, dtype, heads=(B, H), seeds=(21, 22, 23, 24))
    o = np.empty((B, H, n, d), np.uint16)
    l = np.empty((B, H, n), np.float32)
    g32 = [np.empty((B, H, n, d), np.float32) for _ in range(3)]
    fa.host_attention_fwd_bwd_half(bits[0], bits[1], bits[2], bits[3], o, l, *g32, n, d, 0.125, True, B, H, dtype)
    g16 = [np.zeros((B, H, n, d), np.uint16) for _ in range(3)]
    o2 = np.empty_like(o)
    fa.host_attention_fwd_bwd_half_ex(bits[0], bits[1], bits[2], bits[3], o2, l, *g16, n, d, 0.125, True, B, H, dtype, gdt)
    assert np.array_equal(o, o2)
    for a, b in zip(g16, g32):
        assert np.array_equal(a, oracle.to_half_bits(b, gdt))
    fa.host_release()


def test_no_writes_outside_the_output_tensors(fa):
    """compute-sanitizer is closed on this pool, so bounds are checked with guard bands: every output
    sits inside a larger buffer filled with a sentinel, N is ragged (not a multiple of any tile), and
    the bands must come back untouched."""
    import torch

    for n, d, causal in ((333, 64, True), (129, 128, False), (1, 64, False)):
        H, guard = 2, 4096
        scale = float(d ** -0.5)
        mk16 = lambda: torch.randn((H, n, d), device="cuda").to(torch.bfloat16)
        Q, K, V, dO = mk16(), mk16(), mk16(), mk16()

        def guarded(numel, dtype, fill):
            buf = torch.full((numel + 2 * guard,), fill, dtype=dtype, device="cuda")
            return buf, buf[guard:guard + numel]

        ob, O = guarded(H * n * d, torch.int16, 0x5A5A)
        lb, L = guarded(H * n, torch.float32, 12345.0)
        fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, 1, H, fa.BF16)
        gb = [guarded(H * n * d, torch.float32, 12345.0) for _ in range(3)]
        wsb = fa.workspace_bytes_backward(n, d, 1, H)
        wb, ws = guarded(wsb, torch.uint8, 0x5A)
        fa.flash_attention_backward(Q, K, V, O, dO, L, gb[0][1], gb[1][1], gb[2][1], n, d, scale, H * n * d, n * d,
                                    causal, 1, H, fa.BF16, ws, wsb)
        fb, Of = guarded(n * d, torch.float32, 12345.0)
        qf = torch.randn((n, d), device="cuda")
        for f in (fa.naive_attention, fa.flash_attention, fa.flash_attention_v2):
            f(qf, qf, qf, Of, n, d, scale, causal)
        torch.cuda.synchronize()
        for buf, fill in [(ob, 0x5A5A), (lb, 12345.0), (wb, 0x5A), (fb, 12345.0)] + [(g[0], 12345.0) for g in gb]:
            assert bool((buf[:guard] == fill).all()) and bool((buf[-guard:] == fill).all()), (n, d)
        assert torch.isfinite(O.view(torch.bfloat16).float()).all()
        assert all(torch.isfinite(g[1]).all() for g in gb)


def test_backward_gpt2_shape_all_heads(fa):
    """BASELINE config 4 (bf16, B=8, H=12, N=4096, d=64, causal): dQ, dK, dV of every head against
    dense fp64 math (the formulas of oracle_backward / main.mm:1091-1179 with the causal mask)."""
    import torch

    B, H, n, d = 8, 12, 4096, 64
    scale = float(1.0 / np.sqrt(d))
    gen = torch.Generator(device="cuda").manual_seed(11)
    Q, K, V, dO = (torch.rand((B, H, n, d), device="cuda", generator=gen).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(4))
    O = torch.empty_like(Q)
    L = torch.empty((B, H, n), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, True, B, H, fa.BF16)
    dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
    wsb = fa.workspace_bytes_backward(n, d, B, H)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, H * n * d, n * d, True, B, H, fa.BF16, ws, wsb)
    torch.cuda.synchronize()
    mask = torch.ones(n, n, dtype=torch.bool, device="cuda").triu(1)
    for b in range(B):
        for h in range(H):
            q, k, v, do = (t[b, h].double() for t in (Q, K, V, dO))
            p = torch.softmax(((q @ k.T) * scale).masked_fill(mask, float("-inf")), dim=1)
            dp = do @ v.T
            ds = p * (dp - (dp * p).sum(1, keepdim=True)) * scale
            for got, want in ((dQ[b, h], ds @ k), (dK[b, h], ds.T @ q), (dV[b, h], p.T @ do)):
                err = (got.double() - want).abs().max().item()
                assert err <= TOL_ABS and err <= TOL_REL[oracle.BF16] * want.abs().max().item(), (b, h, err)
