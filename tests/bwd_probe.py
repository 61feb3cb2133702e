"""Development probe (not a pytest file): one backward case vs the oracle, in its own process.
usage: python tests/bwd_probe.py N D dtype causal [B H]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle, flash_attention_metal_b200 as fa

n, d, dtype, causal = (int(x) for x in sys.argv[1:5])
scale = float(1 / np.sqrt(d))
mk = lambda s: oracle.to_half_bits(oracle.init_random(n * d, s).reshape(n, d), dtype)
qb, kb, vb, dob = mk(1), mk(2), mk(3), mk(4)
qf, kf, vf, dof = (oracle.from_half_bits(t, dtype) for t in (qb, kb, vb, dob))
want = oracle.backward(qf, kf, vf, dof, scale, bool(causal), streaming=n > 2048)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
Q, K, V, dO = (dev(t.view(np.int16)) for t in (qb, kb, vb, dob))
O = torch.zeros((n, d), dtype=torch.int16, device="cuda"); L = torch.zeros((n,), device="cuda")
fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, n * d, n * d, L, causal, 1, 1, dtype)
dQ, dK, dV = (torch.full((n, d), float("nan"), device="cuda") for _ in range(3))
wsb = fa.workspace_bytes_backward(n, d, 1, 1); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, n * d, n * d, causal, 1, 1, dtype, ws, wsb)
torch.cuda.synchronize()
out = f"bwd N={n} D={d} dtype={dtype} causal={causal}:"
for name, g, w in zip(("dQ", "dK", "dV"), (dQ, dK, dV), want):
    g = g.cpu().numpy(); e = np.abs(g - w)
    out += f"  {name} err {np.nanmax(e):.3e} (max|ref| {np.abs(w).max():.3e}, nan {np.isnan(g).sum()})"
    if np.nanmax(e) > 0.05 * np.abs(w).max() or np.isnan(g).any():
        bad = np.where(~(e.max(1) <= 0.05 * np.abs(w).max()))[0]
        out += f"\n    {name} bad rows {len(bad)}: {bad[:12]} got {g[bad[0], :4] if len(bad) else ''} want {w[bad[0], :4] if len(bad) else ''}"
print(out)
