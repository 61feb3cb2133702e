"""Development probe (not a pytest file): run one 16-bit forward case in its own process and
print where it disagrees with the oracle.  usage: python tests/tc_probe.py N D dtype causal"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle, flash_attention_metal_b200 as fa

n, d, dtype, causal = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "rand"
scale = float(1 / np.sqrt(d))
q, k, v = (oracle.init_random(n * d, 42 + i).reshape(n, d) for i in range(3))
if mode == "vid":   # V = first d rows identity-ish: O shows P directly
    v = np.zeros((n, d), np.float32); v[np.arange(n), np.arange(n) % d] = 1
qb, kb, vb = (oracle.to_half_bits(t, dtype) for t in (q, k, v))
qf, kf, vf = (oracle.from_half_bits(t, dtype) for t in (qb, kb, vb))
want, wl = oracle.forward(qf, kf, vf, scale, bool(causal))
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
O = torch.zeros((n, d), dtype=torch.int16, device="cuda"); L = torch.zeros((n,), device="cuda")
fa.flash_attention_v4_half(dev(qb.view(np.int16)), dev(kb.view(np.int16)), dev(vb.view(np.int16)), O, n, d, scale,
                           n * d, n * d, L, causal, 1, 1, dtype)
torch.cuda.synchronize()
got = oracle.from_half_bits(O.cpu().numpy().view(np.uint16), dtype); gl = L.cpu().numpy()
eo = np.abs(got - want); el = np.abs(gl - wl)
print(f"case N={n} D={d} dtype={dtype} causal={causal} mode={mode}: O max-abs {np.nanmax(eo):.4e} nan={np.isnan(got).sum()}  L max-abs {np.nanmax(el):.4e} nan={np.isnan(gl).sum()}")
if not (np.nanmax(eo) <= 2e-2 and np.nanmax(el) <= 5e-3):
    bad_rows = np.where(eo.max(1) > 2e-2)[0]; bad_cols = np.where(eo.max(0) > 2e-2)[0]
    print("  bad rows:", len(bad_rows), bad_rows[:16], " bad cols:", len(bad_cols), bad_cols[:16])
    print("  L bad rows:", np.where(el > 5e-3)[0][:16], "L got/want", gl[:4], wl[:4])
    print("  got[0,:8]", got[0, :8], "\n  want[0,:8]", want[0, :8])
    r = bad_rows[0] if len(bad_rows) else 0
    print(f"  got[{r},:8]", got[r, :8], f"\n  want[{r},:8]", want[r, :8])
