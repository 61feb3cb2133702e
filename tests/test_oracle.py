"""CPU tests of the oracle (oracle/cpu_ref.c) -- no GPU.

The oracle is pinned three ways:
  1. against tests/golden/reference_cpu.npz, outputs of the reference's own CPU
     verifier loops (main.mm:24-30, 128-159, 550-578, 1092-1179) frozen by
     tests/golden/make_golden.py;
  2. against that same reference build live, when oracle/_ref/libref_cpu.so exists;
  3. against libstdc++'s std::mt19937 / uniform_real_distribution for the inputs.
Plus its own accuracy (fp32 vs fp64) and a finite-difference check of the backward.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle

D = 64
SCALE = np.float32(1.0 / np.sqrt(64.0))


def _qkv(n, d=D):
    return oracle.init_random(n * d).reshape(n, d)


# ---- 1. golden vectors ------------------------------------------------------
def test_init_random_matches_golden(golden):
    x = oracle.init_random(128 * D)
    assert np.array_equal(x[:8].view(np.uint32), golden["init8"].view(np.uint32))
    assert float(x.astype(np.float64).sum()) == float(golden["init_sum"])
    # SURVEY.md section 8 a9 anchors
    assert x[:4].view(np.uint32).tolist() == [0xBE80788E, 0x3F17D47C, 0x3F66C406, 0xBF2214D6]


def test_forward_n128_bit_exact_vs_golden(golden):
    q = _qkv(128)
    o_faithful, _ = oracle.forward(q, q, q, SCALE, faithful=True)
    o_hoisted, _ = oracle.forward(q, q, q, SCALE)
    assert np.array_equal(o_faithful, golden["fwd128"])
    assert np.array_equal(o_hoisted, golden["fwd128"])


def test_forward_n1024_vs_golden(golden):
    q = _qkv(1024)
    o, _ = oracle.forward(q, q, q, SCALE)
    assert np.array_equal(o[[0, 1, 511, 1023]], golden["fwd1024_rows"])
    assert float(o.astype(np.float64).sum()) == float(golden["fwd1024_sum"])


def test_causal_n128_bit_exact_vs_golden(golden):
    q = _qkv(128)
    o, _ = oracle.forward(q, q, q, SCALE, causal=True)
    assert np.array_equal(o, golden["causal128"])
    # causal pattern: row 0 attends to key 0 only -> equals V[0] bit for bit
    assert np.array_equal(o[0], q[0])
    # last row sees every key -> equals the non-causal last row
    assert np.array_equal(o[-1], golden["fwd128"][-1])


def test_backward_n128_bit_exact_vs_golden(golden):
    qf = oracle.from_half_bits(golden["bwd128_qbits"], oracle.FP16)
    dq, dk, dv = oracle.backward(qf, qf, qf, qf, SCALE)
    assert np.array_equal(dq, golden["bwd128_dq"])
    assert np.array_equal(dk, golden["bwd128_dk"])
    assert np.array_equal(dv, golden["bwd128_dv"])


# ---- 2. live reference build -------------------------------------------------
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("n", [16, 100, 128, 256])
def test_forward_matches_reference_loops(n):
    R = oracle.ref()
    rng = np.random.default_rng(n)
    q, k, v = (rng.uniform(-1, 1, (n, D)).astype(np.float32) for _ in range(3))
    want = np.empty_like(q)
    R.ref_forward(q, k, v, want, n, D, SCALE)
    got, _ = oracle.forward(q, k, v, SCALE)
    assert np.array_equal(got, want)
    R.ref_forward_causal(q, k, v, want, n, D, SCALE)
    got, _ = oracle.forward(q, k, v, SCALE, causal=True)
    assert np.array_equal(got, want)


@needs_ref
def test_backward_matches_reference_loops_and_shows_its_bug():
    R = oracle.ref()
    n = 64
    qb = oracle.to_half_bits(_qkv(n) * np.float32(0.01), oracle.FP16)
    dob = oracle.to_half_bits(oracle.init_random(n * D, seed=7).reshape(n, D) * np.float32(0.01), oracle.FP16)
    dq, dk, dv = (np.empty((n, D), np.float32) for _ in range(3))
    R.ref_backward(qb.reshape(-1), dob.reshape(-1), dq, dk, dv, n, D, SCALE)
    qf, dof = oracle.from_half_bits(qb, 0), oracle.from_half_bits(dob, 0)
    g = oracle.backward(qf, qf, qf, dof, SCALE)
    for a, b in zip(g, (dq, dk, dv)):
        assert np.array_equal(a, b)
    # the reference's literal decoding (main.mm:1100) yields dQ == 0: its check is vacuous
    dq_bug = np.empty((n, D), np.float32)
    R.ref_backward_buggy(qb.reshape(-1), dob.reshape(-1), dq_bug, n, D, SCALE)
    assert np.abs(dq_bug).max() == 0.0 and np.abs(dq).max() > 0.0


# ---- 3. inputs against libstdc++ --------------------------------------------
def test_init_random_matches_libstdcxx():
    src = r"""
#include <random>
#include <cstdio>
int main() { std::mt19937 gen(42); std::uniform_real_distribution<float> dis(-1.0f, 1.0f);
  for (int i = 0; i < 100000; ++i) { float f = dis(gen); unsigned u; __builtin_memcpy(&u, &f, 4);
    if (i < 16 || i % 997 == 0) std::printf("%u\n", u); } }
"""
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "a.cpp")
        open(p, "w").write(src)
        subprocess.check_call(["g++", "-O1", "-o", os.path.join(td, "a"), p])
        lines = subprocess.check_output([os.path.join(td, "a")]).split()
    want = np.array([int(s) for s in lines], dtype=np.uint32)
    x = oracle.init_random(100000).view(np.uint32)
    idx = [i for i in range(100000) if i < 16 or i % 997 == 0]
    assert np.array_equal(x[idx], want)


# ---- conversions -------------------------------------------------------------
def test_half_conversions_match_numpy_and_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.uniform(-1, 1, 20000), rng.uniform(-70000, 70000, 2000), rng.uniform(-1e-6, 1e-6, 2000),
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 2.0 ** -24, 2.0 ** -25, 3 * 2.0 ** -26, np.inf, -np.inf]),
    ]).astype(np.float32)
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(oracle.to_half_bits(x, oracle.FP16), want)
    assert np.array_equal(oracle.from_half_bits(want, oracle.FP16), want.view(np.float16).astype(np.float32))
    torch = pytest.importorskip("torch")
    tb = torch.from_numpy(x).to(torch.bfloat16)
    assert np.array_equal(oracle.to_half_bits(x, oracle.BF16), tb.view(torch.int16).numpy().view(np.uint16))
    assert np.array_equal(oracle.round_to(x, oracle.BF16), tb.float().numpy())


# ---- accuracy of the fp32 oracle --------------------------------------------
@pytest.mark.parametrize("causal", [False, True])
def test_forward_fp32_close_to_fp64(causal):
    n = 300
    q = oracle.init_random(n * D, 1).reshape(n, D)
    k = oracle.init_random(n * D, 2).reshape(n, D)
    v = oracle.init_random(n * D, 3).reshape(n, D)
    o, l = oracle.forward(q, k, v, SCALE, causal)
    o64, l64 = oracle.forward_f64(q, k, v, SCALE, causal)
    assert np.abs(o - o64).max() < 2e-6
    assert np.abs(l - l64).max() < 2e-6
    # L is the log-sum-exp of the scaled scores (kernels.metal:863)
    s = (q.astype(np.float64) @ k.astype(np.float64).T) * float(SCALE)
    if causal:
        s = np.where(np.tril(np.ones((n, n), bool)), s, -np.inf)
    assert np.abs(np.log(np.exp(s).sum(1)) - l64).max() < 1e-12


@pytest.mark.parametrize("causal", [False, True])
def test_backward_fp32_close_to_fp64_and_streaming(causal):
    n = 96
    q, k, v, do = (oracle.init_random(n * D, s).reshape(n, D) for s in (1, 2, 3, 4))
    g32 = oracle.backward(q, k, v, do, SCALE, causal)
    g64 = oracle.backward_f64(q, k, v, do, SCALE, causal)
    gs = oracle.backward(q, k, v, do, SCALE, causal, streaming=True)
    for a, b, c in zip(g32, g64, gs):
        assert np.abs(a - b).max() < 5e-6
        assert np.abs(c - b).max() < 5e-6


@pytest.mark.parametrize("causal", [False, True])
def test_backward_matches_finite_differences(causal):
    n, d = 12, 8
    rng = np.random.default_rng(5)
    q, k, v, do = (rng.uniform(-1, 1, (n, d)).astype(np.float32) for _ in range(4))
    scale = np.float32(0.35)
    dq, dk, dv = oracle.backward_f64(q, k, v, do, scale, causal)

    def loss(q_, k_, v_):
        s = (q_.astype(np.float64) @ k_.astype(np.float64).T) * float(scale)
        if causal:
            s = np.where(np.tril(np.ones((n, n), bool)), s, -np.inf)
        p = np.exp(s - s.max(1, keepdims=True))
        p /= p.sum(1, keepdims=True)
        return float(((p @ v_.astype(np.float64)) * do).sum())

    eps = 1e-3
    for name, g in (("q", dq), ("k", dk), ("v", dv)):
        for (i, j) in [(0, 0), (3, 5), (n - 1, d - 1), (5, 2)]:
            args = {"q": q.copy(), "k": k.copy(), "v": v.copy()}
            args[name][i, j] += eps
            up = loss(args["q"], args["k"], args["v"])
            args[name][i, j] -= 2 * eps
            dn = loss(args["q"], args["k"], args["v"])
            assert abs((up - dn) / (2 * eps) - g[i, j]) < 2e-4, (name, i, j)
