"""Multi-GPU check + timing of ring attention (not a pytest file; run under torchrun):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 \
      tests/ring_check.py [--n-total N] [--heads H] [--hdim D] [--causal 0|1] [--check 0|1] [--reps R]
Every rank builds the same full Q/K/V (same seed), runs the single-GPU kernel on the full problem
(when --check 1) and compares its ring output rows with it; then times the ring forward with CUDA
events (max over ranks) and rank 0 prints one JSON line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import flash_attention_metal_b200 as fa

ap = argparse.ArgumentParser()
ap.add_argument("--n-total", type=int, default=16384)
ap.add_argument("--heads", type=int, default=4)
ap.add_argument("--hdim", type=int, default=128)
ap.add_argument("--causal", type=int, default=1)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--bwd", type=int, default=0)
ap.add_argument("--gather", type=int, default=0, help="1: all-gather forward mode (sets FA_RING_GATHER=1)")
a = ap.parse_args()
if a.gather:
    os.environ["FA_RING_GATHER"] = "1"  # read by the library at its first ring forward
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, D, N = a.heads, a.hdim, a.n_total
n_local = N // world
scale = D ** -0.5
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.frombuffer(bytearray(fa.ring_unique_id()), dtype=torch.uint8).cuda()
dist.broadcast(uid, 0)
ring = fa.Ring(bytes(uid.cpu().numpy().tobytes()), rank, world, local)
rows = torch.cat([torch.arange(f, f + r) for f, r in fa.ring_local_rows(rank, world, n_local, bool(a.causal))]).cuda()
g = torch.Generator(device="cuda").manual_seed(1)
if a.check:
    Qf, Kf, Vf = (torch.rand((H, N, D), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(3))
    Q, K, V = (t[:, rows].contiguous() for t in (Qf, Kf, Vf))
else:
    Q, K, V = (torch.rand((H, n_local, D), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(3))
O = torch.zeros_like(Q); L = torch.zeros((H, n_local), device="cuda")
wsb = ring.workspace_bytes_gather(world, n_local, D, H, fa.BF16) if a.gather else ring.workspace_bytes(n_local, D, H, fa.BF16)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()
run = lambda: ring.forward(Q, K, V, O, L, n_local, D, H, scale, a.causal, fa.BF16, ws, wsb, st)
run(); torch.cuda.synchronize()
err = errl = None
if a.check:
    Of = torch.empty_like(Qf); Lf = torch.empty((H, N), device="cuda")
    fa.flash_attention_v4_half(Qf, Kf, Vf, Of, N, D, scale, H * N * D, N * D, Lf, a.causal, 1, H, fa.BF16)
    torch.cuda.synchronize()
    err = (O.float() - Of[:, rows].float()).abs().max().item()
    errl = (L - Lf[:, rows]).abs().max().item()
    t = torch.tensor([err, errl], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); err, errl = t.tolist()
berr = None
if a.bwd:
    gd = torch.Generator(device="cuda").manual_seed(2)
    if a.check:
        dOf = torch.rand((H, N, D), device="cuda", generator=gd).mul_(2).sub_(1).to(torch.bfloat16)
        dO = dOf[:, rows].contiguous()
    else:
        dO = torch.rand((H, n_local, D), device="cuda", generator=gd).mul_(2).sub_(1).to(torch.bfloat16)
    dQ, dK, dV = (torch.full((H, n_local, D), float("nan"), device="cuda") for _ in range(3))
    bwsb = ring.workspace_bytes_backward(n_local, D, H, fa.BF16); bws = torch.empty(bwsb, dtype=torch.uint8, device="cuda")
    run_b = lambda: ring.backward(Q, K, V, O, dO, L, dQ, dK, dV, n_local, D, H, scale, a.causal, fa.BF16, bws, bwsb, st)
    run_b(); torch.cuda.synchronize()
    if a.check:
        gQ, gK, gV = (torch.empty((H, N, D), device="cuda") for _ in range(3))
        w1 = fa.workspace_bytes_backward(N, D, 1, H); w1b = torch.empty(w1, dtype=torch.uint8, device="cuda")
        fa.flash_attention_backward(Qf, Kf, Vf, Of, dOf, Lf, gQ, gK, gV, N, D, scale, H * N * D, N * D, a.causal, 1, H, fa.BF16, w1b, w1)
        torch.cuda.synchronize()
        errs = [(x - y[:, rows]).abs().max().item() / y.abs().max().item() for x, y in ((dQ, gQ), (dK, gK), (dV, gV))]
        t = torch.tensor(errs, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); berr = t.tolist()
    run_f = run
    run = lambda: (run_f(), run_b())
for _ in range(2): run()
dist.barrier(); torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(st)
for _ in range(a.reps): run()
ev1.record(st); torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1) / a.reps], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
flops = 4.0 * H * N * N * D * (0.5 if a.causal else 1.0) * (3.5 if a.bwd else 1.0)
if rank == 0:
    print(json.dumps({"ring_forward": True, "world": world, "N_total": N, "n_local": n_local, "H": H, "d": D, "causal": a.causal,
                      "ms": ms.item(), "tflops_total": flops / ms.item() / 1e9, "tflops_per_gpu": flops / ms.item() / 1e9 / world,
                      "gather": a.gather, "max_abs_err_vs_single_gpu": err, "max_abs_L_err": errl, "bwd": a.bwd,
                      "bwd_rel_err_dq_dk_dv_vs_single_gpu": berr}), flush=True)
ring.close()
dist.destroy_process_group()
