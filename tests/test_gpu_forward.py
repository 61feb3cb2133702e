"""GPU parity tests (run with -m gpu on a B200): every forward variant, called through
the C ABI, against the oracle (oracle/cpu_ref.c, the restated reference CPU verifier).

Tolerances are BASELINE.json's: max-abs <= 1e-4 for fp32, <= 2e-2 for fp16/bf16
(vs. the fp32 oracle evaluated on the same 16-bit-rounded inputs), causal pattern exact.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4
TOL_HALF = 2e-2
TOL_LSE = 5e-3


@pytest.fixture(scope="module")
def fa():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import flash_attention_metal_b200 as fa

    fa.lib()  # fails loudly if the extension is not built
    return fa


def _torch():
    import torch

    return torch


def dev(x):
    return _torch().from_numpy(np.ascontiguousarray(x)).cuda()


def ref_inputs(n, d):
    """The reference's quirk: initRandom re-seeds per call, so Q == K == V (main.mm:117-119)."""
    q = oracle.init_random(n * d).reshape(n, d)
    return q, q.copy(), q.copy()


def indep_inputs(n, d, seed=0, heads=None):
    shape = (n, d) if heads is None else (*heads, n, d)
    size = int(np.prod(shape))
    return tuple(oracle.init_random(size, seed=seed + 42 + i).reshape(shape) for i in range(3))


FP32_FUNCS = ["naive_attention", "flash_attention", "flash_attention_v2"]


# ------------------------------------------------------------------ fp32 ----
@pytest.mark.parametrize("func", FP32_FUNCS)
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("n,d,inputs", [
    (128, 64, "ref"), (1024, 64, "ref"), (128, 64, "indep"), (100, 64, "indep"), (333, 64, "indep"),
    (1, 64, "indep"), (65, 128, "indep"), (512, 128, "indep"), (1000, 128, "ref"),
])
def test_fp32_variants_match_oracle(fa, func, causal, n, d, inputs):
    q, k, v = ref_inputs(n, d) if inputs == "ref" else indep_inputs(n, d, seed=n)
    scale = float(1.0 / np.sqrt(d))
    want, _ = oracle.forward(q, k, v, scale, causal)
    Q, K, V = dev(q), dev(k), dev(v)
    O = _torch().full((n, d), float("nan"), device="cuda")
    getattr(fa, func)(Q, K, V, O, n, d, scale, causal)
    got = O.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= TOL_F32


@pytest.mark.parametrize("func", FP32_FUNCS)
def test_fp32_golden_reference_vectors(fa, func, golden):
    """N=128, D=64 outputs of the reference's own CPU loops (tests/golden)."""
    q, k, v = ref_inputs(128, 64)
    Q = dev(q)
    for causal, key in ((False, "fwd128"), (True, "causal128")):
        O = _torch().empty((128, 64), device="cuda")
        getattr(fa, func)(Q, Q, Q, O, 128, 64, 0.125, causal)
        assert np.abs(O.cpu().numpy() - golden[key]).max() <= TOL_F32


@pytest.mark.parametrize("func", FP32_FUNCS)
def test_fp32_causal_pattern_exact(fa, func):
    """Row 0 attends to key 0 only: softmax weight is exactly 1, O[0] == V[0] bit for bit."""
    n, d = 256, 64
    q, k, v = indep_inputs(n, d, seed=9)
    O = _torch().empty((n, d), device="cuda")
    getattr(fa, func)(dev(q), dev(k), dev(v), O, n, d, 0.125, True)
    got = O.cpu().numpy()
    assert np.array_equal(got[0], v[0])
    # changing keys/values j > i must not change row i
    k2, v2 = k.copy(), v.copy()
    k2[100:], v2[100:] = 7.0, -3.0
    O2 = _torch().empty((n, d), device="cuda")
    getattr(fa, func)(dev(q), dev(k2), dev(v2), O2, n, d, 0.125, True)
    assert np.array_equal(O2.cpu().numpy()[:100], got[:100])


def test_fp32_batched_v2(fa):
    B, H, n, d = 2, 3, 200, 64
    q, k, v = indep_inputs(n, d, seed=3, heads=(B, H))
    want, _ = oracle.forward_batched(q, k, v, 0.125, True)
    O = _torch().empty((B, H, n, d), device="cuda")
    fa.flash_attention_v2_batched(dev(q), dev(k), dev(v), O, n, d, 0.125, H * n * d, n * d, True, B, H)
    assert np.abs(O.cpu().numpy() - want).max() <= TOL_F32


# ---------------------------------------------------------------- 16-bit ----
def half_case(n, d, dtype, inputs, heads=None, seed=0, scale_in=1.0):
    if inputs == "ref":
        q, k, v = ref_inputs(n, d)
    else:
        q, k, v = indep_inputs(n, d, seed=seed, heads=heads)
    qb, kb, vb = (oracle.to_half_bits(t * np.float32(scale_in), dtype) for t in (q, k, v))
    qf, kf, vf = (oracle.from_half_bits(t, dtype) for t in (qb, kb, vb))
    return (qb, kb, vb), (qf, kf, vf)


@pytest.mark.parametrize("dtype", [oracle.FP16, oracle.BF16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("n,d,inputs", [
    (128, 64, "ref"), (1024, 64, "ref"), (128, 64, "indep"), (256, 64, "indep"), (200, 64, "indep"),
    (1, 64, "indep"), (129, 64, "indep"), (1000, 64, "indep"), (2048, 64, "indep"),
    (128, 128, "indep"), (384, 128, "indep"), (777, 128, "indep"), (2048, 128, "indep"),
])
def test_v4_half_matches_oracle(fa, dtype, causal, n, d, inputs):
    bits, f = half_case(n, d, dtype, inputs, seed=n + d)
    scale = float(1.0 / np.sqrt(d))
    want, want_l = oracle.forward(*f, scale, causal)
    Q, K, V = (dev(b.view(np.int16)) for b in bits)
    O = _torch().full((n, d), -1, dtype=_torch().int16, device="cuda")
    L = _torch().full((n,), float("nan"), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, n * d, n * d, L, causal, 1, 1, dtype)
    got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), dtype)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= TOL_HALF
    assert np.abs(L.cpu().numpy() - want_l).max() <= TOL_LSE


@pytest.mark.parametrize("dtype", [oracle.FP16, oracle.BF16])
def test_v4_half_golden_reference_vectors(fa, dtype, golden):
    """The reference's own checks: V4 vs fp32 results at N=128 causal (main.mm:549-594, tol 1e-2
    there) -- here against the frozen outputs of its CPU loops on unrounded inputs."""
    bits, _ = half_case(128, 64, dtype, "ref")
    Q = dev(bits[0].view(np.int16))
    for causal, key in ((False, "fwd128"), (True, "causal128")):
        O = _torch().empty((128, 64), dtype=_torch().int16, device="cuda")
        fa.flash_attention_v4_half(Q, Q, Q, O, 128, 64, 0.125, 8192, 8192, None, causal, 1, 1, dtype)
        got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), dtype)
        assert np.abs(got - golden[key]).max() <= TOL_HALF


@pytest.mark.parametrize("dtype", [oracle.FP16, oracle.BF16])
def test_v3_simd_matches_oracle(fa, dtype):
    n, d = 1024, 64
    bits, f = half_case(n, d, dtype, "ref")
    want, _ = oracle.forward(*f, 0.125, False)
    Q = dev(bits[0].view(np.int16))
    O = _torch().empty((n, d), dtype=_torch().int16, device="cuda")
    fa.flash_attention_simd(Q, Q, Q, O, n, d, 0.125, dtype)
    got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), dtype)
    assert np.abs(got - want).max() <= TOL_HALF


@pytest.mark.parametrize("dtype", [oracle.FP16, oracle.BF16])
@pytest.mark.parametrize("d", [64, 128])
def test_v4_half_causal_pattern_exact(fa, dtype, d):
    n = 512
    bits, f = half_case(n, d, dtype, "indep", seed=5)
    scale = float(1.0 / np.sqrt(d))
    Q, K, V = (dev(b.view(np.int16)) for b in bits)
    O = _torch().empty((n, d), dtype=_torch().int16, device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, n * d, n * d, None, True, 1, 1, dtype)
    got = O.cpu().numpy().view(np.uint16)
    assert np.array_equal(got[0], bits[2][0])  # row 0 == V[0] exactly
    kb2, vb2 = bits[1].copy(), bits[2].copy()
    kb2[300:], vb2[300:] = oracle.to_half_bits(np.float32([5.0]), dtype)[0], oracle.to_half_bits(np.float32([-2.0]), dtype)[0]
    O2 = _torch().empty((n, d), dtype=_torch().int16, device="cuda")
    fa.flash_attention_v4_half(Q, dev(kb2.view(np.int16)), dev(vb2.view(np.int16)), O2, n, d, scale, n * d, n * d,
                               None, True, 1, 1, dtype)
    assert np.array_equal(O2.cpu().numpy().view(np.uint16)[:300], got[:300])


@pytest.mark.parametrize("dtype", [oracle.FP16, oracle.BF16])
@pytest.mark.parametrize("causal", [False, True])
def test_v4_half_batched_with_strides(fa, dtype, causal):
    """B=2, H=3 with padded head/batch strides; L is [B, H, N] at offset/D (kernels.metal:622-623)."""
    B, H, n, d = 2, 3, 300, 64
    bits, f = half_case(n, d, dtype, "indep", heads=(B, H), seed=11)
    want, want_l = oracle.forward_batched(*f, 0.125, causal)
    hs = n * d
    bs = H * hs
    Q, K, V = (dev(b.view(np.int16)) for b in bits)
    O = _torch().zeros((B, H, n, d), dtype=_torch().int16, device="cuda")
    L = _torch().zeros((B, H, n), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, 0.125, bs, hs, L, causal, B, H, dtype)
    got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), dtype)
    assert np.abs(got - want).max() <= TOL_HALF
    assert np.abs(L.cpu().numpy() - want_l).max() <= TOL_LSE


def test_host_buffer_entry_points(fa):
    n, d = 256, 64
    q, k, v = indep_inputs(n, d, seed=21)
    want, want_l = oracle.forward(q, k, v, 0.125, True)
    o = np.empty_like(q)
    for variant in (fa.NAIVE, fa.V1, fa.V2):
        o[:] = np.nan
        fa.host_attention_f32(variant, q, k, v, o, n, d, 0.125, True)
        assert np.abs(o - want).max() <= TOL_F32
    bits, f = half_case(n, d, oracle.BF16, "indep", seed=21)
    want, want_l = oracle.forward(*f, 0.125, True)
    ob = np.empty((n, d), np.uint16)
    lb = np.empty(n, np.float32)
    fa.host_attention_half(bits[0], bits[1], bits[2], ob, lb, n, d, 0.125, True, 1, 1, oracle.BF16)
    assert np.abs(oracle.from_half_bits(ob, oracle.BF16) - want).max() <= TOL_HALF
    assert np.abs(lb - want_l).max() <= TOL_LSE


def test_large_magnitude_scores_do_not_overflow(fa):
    """Scores far from zero and a late, much larger maximum exercise the conditional rescale."""
    n, d = 1024, 64
    q, k, v = indep_inputs(n, d, seed=31)
    k = k.copy()
    k[900:] *= 6.0  # the row maximum jumps by >> 2^8 in the last tile
    q = q * 4.0
    for dtype in (oracle.FP16, oracle.BF16):
        qb, kb, vb = (oracle.to_half_bits(t, dtype) for t in (q, k, v))
        f = tuple(oracle.from_half_bits(t, dtype) for t in (qb, kb, vb))
        want, want_l = oracle.forward(*f, 0.125, False)
        O = _torch().empty((n, d), dtype=_torch().int16, device="cuda")
        L = _torch().empty((n,), device="cuda")
        fa.flash_attention_v4_half(dev(qb.view(np.int16)), dev(kb.view(np.int16)), dev(vb.view(np.int16)), O, n, d,
                                   0.125, n * d, n * d, L, False, 1, 1, dtype)
        got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), dtype)
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() <= TOL_HALF
        assert np.abs(L.cpu().numpy() - want_l).max() <= 2e-2


# ------------------------------------------------ BASELINE full-size checks ----
@pytest.mark.parametrize("causal", [True, False])
def test_flagship_shape_sampled_rows(fa, causal):
    """BASELINE config 3: bf16, B=1, H=16, N=16384, d=128.  The oracle cannot finish the whole
    problem in seconds, so sampled query rows are recomputed in fp64 numpy from the same
    rounded inputs, plus size-independent properties (row 0 == V[0]; L == logsumexp)."""
    torch = _torch()
    B, H, n, d = 1, 16, 16384, 128
    scale = float(1.0 / np.sqrt(d))
    g = torch.Generator(device="cuda").manual_seed(1234)
    Q, K, V = (torch.rand((B, H, n, d), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(3))
    O = torch.empty_like(Q)
    L = torch.empty((B, H, n), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, fa.BF16)
    torch.cuda.synchronize()
    rng = np.random.default_rng(0)
    rows = sorted({0, 1, 127, 128, 255, 256, n - 1, *rng.integers(0, n, 24).tolist()})
    for h in (0, 7, 15):
        k64 = K[0, h].double()
        v64 = V[0, h].double()
        for i in rows:
            nk = i + 1 if causal else n
            s = (k64[:nk] @ Q[0, h, i].double()) * scale
            lse = torch.logsumexp(s, 0)
            want = torch.softmax(s, 0) @ v64[:nk]
            assert (O[0, h, i].double() - want).abs().max().item() <= TOL_HALF
            assert abs(L[0, h, i].item() - lse.item()) <= TOL_LSE
    if causal:
        assert torch.equal(O[0, :, 0], V[0, :, 0])


def _dense_reference(Q, K, V, scale, causal):
    """fp64 dense attention of one head on the GPU (same formulas as the oracle's forward:
    oracle/cpu_ref.c oracle_forward / main.mm:128-159, 549-578), used where the CPU oracle cannot
    finish: it is first checked against the oracle itself in test_dense_reference_matches_oracle."""
    torch = _torch()
    s = (Q.double() @ K.double().T) * scale
    if causal:
        n = s.shape[0]
        s = s.masked_fill(torch.ones(n, n, dtype=torch.bool, device=s.device).triu(1), float("-inf"))
    return torch.softmax(s, dim=1) @ V.double(), torch.logsumexp(s, dim=1)


@pytest.mark.parametrize("causal", [False, True])
def test_dense_reference_matches_oracle(fa, causal):
    n, d, scale = 256, 64, 0.125
    q, k, v = indep_inputs(n, d)
    want, want_l = oracle.forward(q, k, v, scale, causal)
    got, got_l = _dense_reference(dev(q), dev(k), dev(v), scale, causal)
    assert np.abs(got.cpu().numpy() - want).max() <= 2e-6
    assert np.abs(got_l.cpu().numpy() - want_l).max() <= 2e-5


@pytest.mark.parametrize("causal", [True, False])
def test_flagship_shape_full_heads(fa, causal):
    """BASELINE config 3 (bf16, B=1, H=16, N=16384, d=128): every row of two heads against the dense
    fp64 reference."""
    torch = _torch()
    B, H, n, d = 1, 16, 16384, 128
    scale = float(1.0 / np.sqrt(d))
    g = torch.Generator(device="cuda").manual_seed(99)
    Q, K, V = (torch.rand((B, H, n, d), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(3))
    O = torch.empty_like(Q)
    L = torch.empty((B, H, n), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, fa.BF16)
    torch.cuda.synchronize()
    for h in (3, 15):
        want, want_l = _dense_reference(Q[0, h], K[0, h], V[0, h], scale, causal)
        assert (O[0, h].double() - want).abs().max().item() <= TOL_HALF
        assert (L[0, h].double() - want_l).abs().max().item() <= TOL_LSE
        del want, want_l


def test_gpt2_shape_all_heads(fa):
    """BASELINE config 4 (bf16, B=8, H=12, N=4096, d=64, causal): every row of every head against
    the dense fp64 reference."""
    torch = _torch()
    B, H, n, d = 8, 12, 4096, 64
    scale = float(1.0 / np.sqrt(d))
    g = torch.Generator(device="cuda").manual_seed(5)
    Q, K, V = (torch.rand((B, H, n, d), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(3))
    O = torch.empty_like(Q)
    L = torch.empty((B, H, n), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, True, B, H, fa.BF16)
    torch.cuda.synchronize()
    worst = 0.0
    for b in range(B):
        for h in range(H):
            want, want_l = _dense_reference(Q[b, h], K[b, h], V[b, h], scale, True)
            worst = max(worst, (O[b, h].double() - want).abs().max().item())
            assert (L[b, h].double() - want_l).abs().max().item() <= TOL_LSE
    assert worst <= TOL_HALF
