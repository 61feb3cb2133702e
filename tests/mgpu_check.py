"""Single-process multi-GPU check of fa_mgpu_* (run directly; wrapped by tests/test_gpu_ring_multi.py):
  python tests/mgpu_check.py [--gpus P] [--n-total N] [--heads H] [--hdim D] [--causal 0|1] [--reps R]
One process drives P GPUs: ring forward + backward (PEER transport, no NCCL) and the batch x heads
sharded forward + backward are compared with the single-GPU kernels on device 0; exits non-zero when a
tolerance is exceeded.  Prints one JSON line with the timings (host wall clock around a synchronised
group of calls: the group's streams live on different devices)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

TOL_O, TOL_L, TOL_G = 2e-2, 5e-3, 1e-2
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--n-total", type=int, default=8192)
ap.add_argument("--heads", type=int, default=4)
ap.add_argument("--hdim", type=int, default=128)
ap.add_argument("--causal", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
P = a.gpus or torch.cuda.device_count()
assert P >= 1 and torch.cuda.device_count() >= P
H, D, N = a.heads, a.hdim, a.n_total
n_local, scale = N // P, D ** -0.5
failures = []
grp = fa.Mgpu(list(range(P)))
g = torch.Generator(device="cuda:0").manual_seed(3)
Qf, Kf, Vf, dOf = (torch.rand((H, N, D), device="cuda:0", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(4))
# single-GPU results on device 0
torch.cuda.set_device(0)
Of = torch.empty_like(Qf); Lf = torch.empty((H, N), device="cuda:0")
fa.flash_attention_v4_half(Qf, Kf, Vf, Of, N, D, scale, H * N * D, N * D, Lf, a.causal, 1, H, fa.BF16)
gQ, gK, gV = (torch.empty((H, N, D), device="cuda:0") for _ in range(3))
w1 = fa.workspace_bytes_backward(N, D, 1, H); w1b = torch.empty(w1, dtype=torch.uint8, device="cuda:0")
fa.flash_attention_backward(Qf, Kf, Vf, Of, dOf, Lf, gQ, gK, gV, N, D, scale, H * N * D, N * D, a.causal, 1, H, fa.BF16, w1b, w1)
torch.cuda.synchronize()

# ---- ring: device i is rank i ----
rows = [torch.cat([torch.arange(f, f + r) for f, r in fa.ring_local_rows(i, P, n_local, bool(a.causal))]) for i in range(P)]
loc = lambda t, i: t[:, rows[i].to(t.device)].contiguous().to(f"cuda:{i}")
Q, K, V, dO = ([loc(t, i) for i in range(P)] for t in (Qf, Kf, Vf, dOf))
O = [torch.zeros_like(q) for q in Q]
L = [torch.zeros((H, n_local), device=f"cuda:{i}") for i in range(P)]
dQ, dK, dV = ([torch.full((H, n_local, D), float("nan"), device=f"cuda:{i}") for i in range(P)] for _ in range(3))
for i in range(P):
    torch.cuda.synchronize(i)
def ring_step():
    grp.ring_forward(Q, K, V, O, L, n_local, D, H, scale, a.causal, fa.BF16)
    grp.ring_backward(Q, K, V, O, dO, L, dQ, dK, dV, n_local, D, H, scale, a.causal, fa.BF16)
ring_step(); grp.synchronize()
err = max((O[i].float().cpu() - Of[:, rows[i].to(Of.device)].float().cpu()).abs().max().item() for i in range(P))
errl = max((L[i].cpu() - Lf[:, rows[i].to(Of.device)].cpu()).abs().max().item() for i in range(P))
berr = [max((x[i].cpu() - y[:, rows[i].to(Of.device)].cpu()).abs().max().item() for i in range(P)) / y.abs().max().item()
        for x, y in ((dQ, gQ), (dK, gK), (dV, gV))]
if not err <= TOL_O: failures.append(f"ring O max-abs {err}")
if not errl <= TOL_L: failures.append(f"ring L max-abs {errl}")
for nm, x in zip(("dQ", "dK", "dV"), berr):
    if not x <= TOL_G: failures.append(f"ring {nm} relative error {x}")
first = [t.clone() for t in O + dQ + dK + dV]
ring_step(); grp.synchronize()
if not all(torch.equal(x, y) for x, y in zip(first, O + dQ + dK + dV)): failures.append("ring results differ between two calls")
t0 = time.perf_counter()
for _ in range(a.reps): ring_step()
grp.synchronize()
ring_ms = (time.perf_counter() - t0) * 1e3 / a.reps

# ---- batch x heads sharding: the H heads split as evenly as possible ----
heads = [H // P + (1 if i < H % P else 0) for i in range(P)]
h0 = [sum(heads[:i]) for i in range(P)]
sl = lambda t, i: t[h0[i]:h0[i] + heads[i]].contiguous().to(f"cuda:{i}")
sQ, sK, sV, sdO = ([sl(t, i) for i in range(P)] for t in (Qf, Kf, Vf, dOf))
sO = [torch.zeros_like(q) for q in sQ]
sL = [torch.zeros((max(heads[i], 1), N), device=f"cuda:{i}") for i in range(P)]
sg = [[torch.full((max(heads[i], 1), N, D), float("nan"), device=f"cuda:{i}") for i in range(P)] for _ in range(3)]
for i in range(P):
    torch.cuda.synchronize(i)
def shard_step():
    grp.sharded_forward(sQ, sK, sV, sO, sL, N, D, scale, a.causal, heads, fa.BF16)
    grp.sharded_backward(sQ, sK, sV, sO, sdO, sL, sg[0], sg[1], sg[2], N, D, scale, a.causal, heads, fa.BF16)
shard_step(); grp.synchronize()
for i in range(P):
    if heads[i] == 0: continue
    e = (sO[i].float().cpu() - Of[h0[i]:h0[i] + heads[i]].float().cpu()).abs().max().item()
    if not e <= 2 ** -7: failures.append(f"sharded O differs on device {i}: {e}")  # same kernel, maybe another split
    for nm, x, y in zip(("dQ", "dK", "dV"), sg, (gQ, gK, gV)):
        e = (x[i][:heads[i]].cpu() - y[h0[i]:h0[i] + heads[i]].cpu()).abs().max().item() / y.abs().max().item()
        if not e <= TOL_G: failures.append(f"sharded {nm} differs on device {i}: {e}")
t0 = time.perf_counter()
for _ in range(a.reps): shard_step()
grp.synchronize()
shard_ms = (time.perf_counter() - t0) * 1e3 / a.reps
flops = 4.0 * H * N * N * D * (0.5 if a.causal else 1.0) * 3.5
print(json.dumps({"mgpu": True, "gpus": P, "N_total": N, "H": H, "d": D, "causal": a.causal, "ring_fwd_bwd_ms": ring_ms,
                  "ring_tflops_total": flops / ring_ms / 1e9, "sharded_fwd_bwd_ms": shard_ms,
                  "sharded_tflops_total": flops / shard_ms / 1e9, "ring_O_err": err, "ring_L_err": errl,
                  "ring_grad_rel_err": berr, "failures": failures}), flush=True)
grp.close()
if failures:
    print("MGPU CHECK FAILED: " + "; ".join(failures), file=sys.stderr, flush=True)
    sys.exit(1)
